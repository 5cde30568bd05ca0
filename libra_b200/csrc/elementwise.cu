// HBM-bound element-wise / gather kernels of the Libra hot path:
//   SwiGLU (A15), bias+quick_gelu (A3), row gather (routing permutation), embeddings (A17),
//   LFQ pack/unpack (A7/A8), attention prologue = bridge add + RoPE + un-permute (A10/A12) and its adjoint,
//   attention-backward prepare (delta + dO gather), fused cross-entropy (A18).
// All use 128-bit loads/stores on rows that are multiples of 8 bf16; grids are sized from the row count.
#include <math_constants.h>

#include "common.cuh"

namespace lb {

union V8 {
    uint4 u;
    __nv_bfloat162 h[4];
};
__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    V8 v;
    v.u = u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(v.h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    V8 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) v.h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v.u;
}

// ------------------------------------------------------------------ SwiGLU
__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ gate, const __nv_bfloat16* __restrict__ up,
                                  __nv_bfloat16* __restrict__ out, int64_t rows, int nvec, int64_t ldg, int64_t ldu,
                                  int64_t ldo) {
    const int64_t total = rows * nvec;
    // (row, vector) of the grid-stride index without a 64-bit division per iteration: divide once, then step
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t dr = stride / nvec;
    const int dv = (int)(stride - dr * nvec);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = i / nvec;
    int v = (int)(i - r * nvec);
    for (; i < total; i += stride, r += dr, v += dv) {
        if (v >= nvec) { v -= nvec; ++r; }
        float g[8], u[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(gate + r * ldg) + v), g);
        unpack8(__ldg(reinterpret_cast<const uint4*>(up + r * ldu) + v), u);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            // reference rounds silu(g) to bf16 before the product (act_fn output dtype); keep one rounding here
            const float s = __fdividef(g[j], 1.f + __expf(-g[j]));      // MUFU.RCP: the result is rounded to bf16 right after
            g[j] = s * u[j];
        }
        reinterpret_cast<uint4*>(out + r * ldo)[v] = pack8(g);
    }
}

__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ gate,
                                  const __nv_bfloat16* __restrict__ up, __nv_bfloat16* __restrict__ dgate,
                                  __nv_bfloat16* __restrict__ dup, int64_t rows, int nvec, int64_t ldd, int64_t ldg,
                                  int64_t ldu, int64_t ldgg, int64_t ldgu) {
    const int64_t total = rows * nvec;
    // (row, vector) of the grid-stride index without a 64-bit division per iteration: divide once, then step
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t dr = stride / nvec;
    const int dv = (int)(stride - dr * nvec);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = i / nvec;
    int v = (int)(i - r * nvec);
    for (; i < total; i += stride, r += dr, v += dv) {
        if (v >= nvec) { v -= nvec; ++r; }
        float g[8], u[8], d[8], dg[8], du[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(gate + r * ldg) + v), g);
        unpack8(__ldg(reinterpret_cast<const uint4*>(up + r * ldu) + v), u);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dout + r * ldd) + v), d);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float sg = __fdividef(1.f, 1.f + __expf(-g[j]));
            const float s = g[j] * sg;
            du[j] = d[j] * s;
            dg[j] = d[j] * u[j] * (sg * (1.f + g[j] * (1.f - sg)));
        }
        reinterpret_cast<uint4*>(dgate + r * ldgg)[v] = pack8(dg);
        reinterpret_cast<uint4*>(dup + r * ldgu)[v] = pack8(du);
    }
}

// ------------------------------------------------------- bias + quick_gelu
__global__ void bias_qgelu_fwd_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ bias,
                                      __nv_bfloat16* __restrict__ y, int64_t rows, int nvec) {
    const int64_t total = rows * nvec;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % nvec);
        float f[8], b[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), f);
        if (bias) {
            unpack8(__ldg(reinterpret_cast<const uint4*>(bias) + v), b);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(__float2bfloat16(f[j] + b[j]));  // Linear output is bf16
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = f[j] / (1.f + __expf(-1.702f * f[j]));
        reinterpret_cast<uint4*>(y)[i] = pack8(f);
    }
}
__global__ void bias_qgelu_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x,
                                      const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ dx,
                                      int64_t rows, int nvec) {
    const int64_t total = rows * nvec;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int v = (int)(i % nvec);
        float f[8], b[8], d[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), f);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dy) + i), d);
        if (bias) {
            unpack8(__ldg(reinterpret_cast<const uint4*>(bias) + v), b);
#pragma unroll
            for (int j = 0; j < 8; ++j) f[j] = __bfloat162float(__float2bfloat16(f[j] + b[j]));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float s = 1.f / (1.f + __expf(-1.702f * f[j]));
            d[j] = d[j] * (s + 1.702f * f[j] * s * (1.f - s));
        }
        reinterpret_cast<uint4*>(dx)[i] = pack8(d);
    }
}

// --------------------------------------------------------------- row gather
template <typename IdxT>
__global__ void gather_rows_kernel(const __nv_bfloat16* __restrict__ src, const IdxT* __restrict__ index,
                                   __nv_bfloat16* __restrict__ dst, int64_t rows, int nvec, int64_t ld_dst,
                                   int col0_vec) {
    const int64_t total = rows * nvec;
    // (row, vector) of the grid-stride index without a 64-bit division per iteration: divide once, then step
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    const int64_t dr = stride / nvec;
    const int dv = (int)(stride - dr * nvec);
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    int64_t r = i / nvec;
    int v = (int)(i - r * nvec);
    for (; i < total; i += stride, r += dr, v += dv) {
        if (v >= nvec) { v -= nvec; ++r; }
        const int64_t s = index ? (int64_t)index[r] : r;
        reinterpret_cast<uint4*>(dst + r * ld_dst)[col0_vec + v] =
            __ldg(reinterpret_cast<const uint4*>(src + s * (int64_t)nvec * 8) + v);
    }
}

__global__ void embed_bwd_kernel(const int64_t* __restrict__ ids, const __nv_bfloat16* __restrict__ dy, int64_t ld_dy,
                                 int col0, float* __restrict__ dtable, int64_t rows, int cols) {
    const int64_t total = rows * cols;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / cols;
        const int c = (int)(i - r * cols);
        atomicAdd(dtable + ids[r] * cols + c, __bfloat162float(dy[r * ld_dy + col0 + c]));
    }
}

// ------------------------------------------------------------------- LFQ
template <typename T>
__device__ __forceinline__ float to_f(T v);
template <>
__device__ __forceinline__ float to_f<float>(float v) { return v; }
template <>
__device__ __forceinline__ float to_f<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }

template <typename T>
__global__ void lfq_pack_kernel(const T* __restrict__ h, int64_t n_img, int tokens, int ncb, int bits, int64_t offset,
                                int64_t boi, int64_t eoi, int64_t* __restrict__ ids) {
    const int64_t total = n_img * (tokens + 2) * ncb;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int q = (int)(i % ncb);
        const int64_t it = i / ncb;
        const int t = (int)(it % (tokens + 2));
        const int64_t img = it / (tokens + 2);
        int64_t val;
        if (t == 0) val = boi;
        else if (t == tokens + 1) val = eoi;
        else {
            const T* p = h + ((img * tokens + (t - 1)) * ncb + q) * bits;
            int code = 0;
            for (int d = 0; d < bits; ++d) code = (code << 1) | (to_f<T>(p[d]) > 0.f ? 1 : 0);   // strict >, MSB first
            val = offset + code;
        }
        ids[((int64_t)q * n_img + img) * (tokens + 2) + t] = val;
    }
}

template <typename T>
__global__ void lfq_unpack_kernel(const int64_t* __restrict__ idx, int64_t n, int ncb, int bits, T* __restrict__ codes) {
    const int64_t total = n * ncb * bits;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int d = (int)(i % bits);
        const int64_t e = i / bits;   // (token, codebook)
        const int bit = (int)((idx[e] >> (bits - 1 - d)) & 1);
        codes[i] = (T)(bit ? 1.0f : -1.0f);
    }
}

// ------------------------------------------------ attention prologue (fwd)
// Rank-r bridge folded into the prologue (lb_attn_prep_fwd_bridge: the one-token decode step, where the separate rank-8
// GEMM launch costs more than its arithmetic): kc = bf16(k + bf16(tk . B_k^T)), vc likewise, B picked by the token's modality.
struct PrepBridge {
    const __nv_bfloat16* tk;       // [n, rank] sorted rows (x . A_k^T), or null: kc / vc come precomputed (or absent)
    const __nv_bfloat16* tv;
    const __nv_bfloat16* Bk[2];    // [C, rank]: index 0 language tokens, 1 vision tokens
    const __nv_bfloat16* Bv[2];
    int rank;                      // multiple of 8
};

// bf16(x + bf16(sum_r t[r] * B[c][r])) for 8 consecutive channels c0..c0+7 (the GEMM epilogue's rounding sequence).  The eight
// B rows of a rank chunk are requested together (one L2 round trip per chunk, not one per channel).
__device__ __forceinline__ void bridge_add8(float (&x)[8], const __nv_bfloat16* __restrict__ t, const __nv_bfloat16* __restrict__ B,
                                            int c0, int rank) {
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int r = 0; r < rank; r += 8) {
        uint4 b4[8];
        const uint4 t4 = __ldg(reinterpret_cast<const uint4*>(t + r));
#pragma unroll
        for (int j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const uint4*>(B + (int64_t)(c0 + j) * rank + r));
        float tb[8];
        unpack8(t4, tb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            float bb[8];
            unpack8(b4[j], bb);
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[j] = fmaf(tb[e], bb[e], acc[j]);
        }
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const float kb = __bfloat162float(__float2bfloat16_rn(acc[j]));
        x[j] = __bfloat162float(__float2bfloat16_rn(x[j] + kb));
    }
}

// One CTA per original token.  A "unit" is 8 rotary pairs: elements [d0,d0+8) and [d0+D/2, d0+D/2+8) of one head.
// kc = k + kb and vc = v + vb (the bridged variants) are produced upstream by rank-r GEMMs (beta = 1), or here (PrepBridge).
// FOLD = false is the training prologue (no bridge code in it: folding the branch into one kernel cost the training path
// 106 -> 145 us per launch); FOLD = true is the decode step's variant (lb_attn_prep_fwd_bridge).
template <bool FOLD>
__global__ void __launch_bounds__(256) attn_prep_fwd_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ k, const __nv_bfloat16* __restrict__ kc,
    const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ vc, const uint8_t* __restrict__ flag_sorted,
    const int32_t* __restrict__ sorted_of, const int32_t* __restrict__ pos, const float* __restrict__ cos_t,
    const float* __restrict__ sin_t, __nv_bfloat16* __restrict__ Q, __nv_bfloat16* __restrict__ Kfv,
    __nv_bfloat16* __restrict__ Kfl, __nv_bfloat16* __restrict__ Vfv, __nv_bfloat16* __restrict__ Vfl, int heads, int D,
    const int32_t* __restrict__ kv_row, const PrepBridge br) {
    pdl_trigger();
    pdl_wait();
    const int64_t bt = blockIdx.x;
    const int64_t kr = kv_row ? (int64_t)kv_row[bt] : bt;      // row of the K/V outputs (decode: the token's slot in the KV cache)
    const int64_t s = sorted_of[bt];
    const bool vis = flag_sorted[s] != 0;
    const int C = heads * D, half = D >> 1, upH = D >> 4;      // units per head
    const int p = pos[bt];
    for (int u = threadIdx.x; u < heads * upH; u += blockDim.x) {
        const int h = u / upH, d0 = (u - h * upH) * 8;
        const int c_lo = h * D + d0, c_hi = c_lo + half;
        float cs[8], sn[8];
        {
            const float4* cp = reinterpret_cast<const float4*>(cos_t + (int64_t)p * half + d0);
            const float4* sp = reinterpret_cast<const float4*>(sin_t + (int64_t)p * half + d0);
            float4 a = __ldg(cp), b = __ldg(cp + 1), c = __ldg(sp), d = __ldg(sp + 1);
            cs[0] = a.x; cs[1] = a.y; cs[2] = a.z; cs[3] = a.w; cs[4] = b.x; cs[5] = b.y; cs[6] = b.z; cs[7] = b.w;
            sn[0] = c.x; sn[1] = c.y; sn[2] = c.z; sn[3] = c.w; sn[4] = d.x; sn[5] = d.y; sn[6] = d.z; sn[7] = d.w;
        }
        auto rope = [&](const __nv_bfloat16* src, uint4& lo, uint4& hi, bool bridged) {
            float xl[8], xh[8], o_lo[8], o_hi[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(src + s * C + c_lo)), xl);
            unpack8(__ldg(reinterpret_cast<const uint4*>(src + s * C + c_hi)), xh);
            if (FOLD && bridged) {
                bridge_add8(xl, br.tk + s * br.rank, br.Bk[vis ? 1 : 0], c_lo, br.rank);
                bridge_add8(xh, br.tk + s * br.rank, br.Bk[vis ? 1 : 0], c_hi, br.rank);
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                o_lo[j] = xl[j] * cs[j] - xh[j] * sn[j];
                o_hi[j] = xh[j] * cs[j] + xl[j] * sn[j];
            }
            lo = pack8(o_lo);
            hi = pack8(o_hi);
        };
        uint4 lo, hi;
        rope(q, lo, hi, false);
        *reinterpret_cast<uint4*>(Q + bt * C + c_lo) = lo;
        *reinterpret_cast<uint4*>(Q + bt * C + c_hi) = hi;
        uint4 kp_lo, kp_hi, kc_lo, kc_hi;
        rope(k, kp_lo, kp_hi, false);
        if (FOLD) rope(k, kc_lo, kc_hi, true);
        else if (kc) rope(kc, kc_lo, kc_hi, false);
        else { kc_lo = kp_lo; kc_hi = kp_hi; }
        // vision token: vision queries (fv) see plain, language queries (fl) see bridged; language token: the reverse
        *reinterpret_cast<uint4*>(Kfv + kr * C + c_lo) = vis ? kp_lo : kc_lo;
        *reinterpret_cast<uint4*>(Kfv + kr * C + c_hi) = vis ? kp_hi : kc_hi;
        *reinterpret_cast<uint4*>(Kfl + kr * C + c_lo) = vis ? kc_lo : kp_lo;
        *reinterpret_cast<uint4*>(Kfl + kr * C + c_hi) = vis ? kc_hi : kp_hi;
        const uint4 vp_lo = __ldg(reinterpret_cast<const uint4*>(v + s * C + c_lo));
        const uint4 vp_hi = __ldg(reinterpret_cast<const uint4*>(v + s * C + c_hi));
        uint4 vc_lo = vc ? __ldg(reinterpret_cast<const uint4*>(vc + s * C + c_lo)) : vp_lo;
        uint4 vc_hi = vc ? __ldg(reinterpret_cast<const uint4*>(vc + s * C + c_hi)) : vp_hi;
        if (FOLD) {
            float xl[8], xh[8];
            unpack8(vp_lo, xl);
            unpack8(vp_hi, xh);
            bridge_add8(xl, br.tv + s * br.rank, br.Bv[vis ? 1 : 0], c_lo, br.rank);
            bridge_add8(xh, br.tv + s * br.rank, br.Bv[vis ? 1 : 0], c_hi, br.rank);
            vc_lo = pack8(xl);
            vc_hi = pack8(xh);
        }
        *reinterpret_cast<uint4*>(Vfv + kr * C + c_lo) = vis ? vp_lo : vc_lo;
        *reinterpret_cast<uint4*>(Vfv + kr * C + c_hi) = vis ? vp_hi : vc_hi;
        *reinterpret_cast<uint4*>(Vfl + kr * C + c_lo) = vis ? vc_lo : vp_lo;
        *reinterpret_cast<uint4*>(Vfl + kr * C + c_hi) = vis ? vc_hi : vp_hi;
    }
}

// ------------------------------------------------ attention prologue (bwd)
__global__ void __launch_bounds__(256) attn_prep_bwd_kernel(
    const __nv_bfloat16* __restrict__ dQ, const __nv_bfloat16* __restrict__ dKfv, const __nv_bfloat16* __restrict__ dKfl,
    const __nv_bfloat16* __restrict__ dVfv, const __nv_bfloat16* __restrict__ dVfl,
    const uint8_t* __restrict__ flag_sorted, const int32_t* __restrict__ sorted_of, const int32_t* __restrict__ pos,
    const float* __restrict__ cos_t, const float* __restrict__ sin_t, __nv_bfloat16* __restrict__ dq,
    __nv_bfloat16* __restrict__ dk, __nv_bfloat16* __restrict__ dv, __nv_bfloat16* __restrict__ dkb,
    __nv_bfloat16* __restrict__ dvb, int heads, int D) {
    const int64_t bt = blockIdx.x;
    const int64_t s = sorted_of[bt];
    const bool vis = flag_sorted[s] != 0;
    const int C = heads * D, half = D >> 1, upH = D >> 4;
    const int p = pos[bt];
    for (int u = threadIdx.x; u < heads * upH; u += blockDim.x) {
        const int h = u / upH, d0 = (u - h * upH) * 8;
        const int c_lo = h * D + d0, c_hi = c_lo + half;
        float cs[8], sn[8];
        {
            const float4* cp = reinterpret_cast<const float4*>(cos_t + (int64_t)p * half + d0);
            const float4* sp = reinterpret_cast<const float4*>(sin_t + (int64_t)p * half + d0);
            float4 a = __ldg(cp), b = __ldg(cp + 1), c = __ldg(sp), d = __ldg(sp + 1);
            cs[0] = a.x; cs[1] = a.y; cs[2] = a.z; cs[3] = a.w; cs[4] = b.x; cs[5] = b.y; cs[6] = b.z; cs[7] = b.w;
            sn[0] = c.x; sn[1] = c.y; sn[2] = c.z; sn[3] = c.w; sn[4] = d.x; sn[5] = d.y; sn[6] = d.z; sn[7] = d.w;
        }
        float yl[8], yh[8], o_lo[8], o_hi[8];
        // rope^T: dx_lo = dy_lo c + dy_hi s ; dx_hi = dy_hi c - dy_lo s
        unpack8(__ldg(reinterpret_cast<const uint4*>(dQ + bt * C + c_lo)), yl);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dQ + bt * C + c_hi)), yh);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o_lo[j] = yl[j] * cs[j] + yh[j] * sn[j];
            o_hi[j] = yh[j] * cs[j] - yl[j] * sn[j];
        }
        *reinterpret_cast<uint4*>(dq + s * C + c_lo) = pack8(o_lo);
        *reinterpret_cast<uint4*>(dq + s * C + c_hi) = pack8(o_hi);

        float al[8], ah[8], bl[8], bh[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(dKfv + bt * C + c_lo)), al);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dKfv + bt * C + c_hi)), ah);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dKfl + bt * C + c_lo)), bl);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dKfl + bt * C + c_hi)), bh);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float tl = al[j] + bl[j], th = ah[j] + bh[j];
            o_lo[j] = tl * cs[j] + th * sn[j];
            o_hi[j] = th * cs[j] - tl * sn[j];
        }
        *reinterpret_cast<uint4*>(dk + s * C + c_lo) = pack8(o_lo);
        *reinterpret_cast<uint4*>(dk + s * C + c_hi) = pack8(o_hi);
        if (dkb) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {      // cross variant: vision token -> Kfl, language token -> Kfv
                const float tl = vis ? bl[j] : al[j], th = vis ? bh[j] : ah[j];
                o_lo[j] = tl * cs[j] + th * sn[j];
                o_hi[j] = th * cs[j] - tl * sn[j];
            }
            *reinterpret_cast<uint4*>(dkb + s * C + c_lo) = pack8(o_lo);
            *reinterpret_cast<uint4*>(dkb + s * C + c_hi) = pack8(o_hi);
        }
        unpack8(__ldg(reinterpret_cast<const uint4*>(dVfv + bt * C + c_lo)), al);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dVfv + bt * C + c_hi)), ah);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dVfl + bt * C + c_lo)), bl);
        unpack8(__ldg(reinterpret_cast<const uint4*>(dVfl + bt * C + c_hi)), bh);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            o_lo[j] = al[j] + bl[j];
            o_hi[j] = ah[j] + bh[j];
        }
        *reinterpret_cast<uint4*>(dv + s * C + c_lo) = pack8(o_lo);
        *reinterpret_cast<uint4*>(dv + s * C + c_hi) = pack8(o_hi);
        if (dvb) {
            *reinterpret_cast<uint4*>(dvb + s * C + c_lo) = pack8(vis ? bl : al);
            *reinterpret_cast<uint4*>(dvb + s * C + c_hi) = pack8(vis ? bh : ah);
        }
    }
}

// --------------------------------------- attention backward prepare: delta
// thread t of the CTA covers elements [16t, 16t+16) of the token's H*D row.
__global__ void __launch_bounds__(256) attn_bwd_prepare_kernel(const __nv_bfloat16* __restrict__ O,
                                                               const __nv_bfloat16* __restrict__ dO,
                                                               const int32_t* __restrict__ row_of,
                                                               __nv_bfloat16* __restrict__ dO_orig,
                                                               float* __restrict__ delta, int seqlen, int heads, int D) {
    const int64_t bt = blockIdx.x;
    const int64_t r = row_of ? (int64_t)row_of[bt] : bt;
    const int C = heads * D;
    const int tph = D >> 4;    // threads per head (power of two <= 32)
    const int b = (int)(bt / seqlen), t = (int)(bt - (int64_t)b * seqlen);
    for (int base = 0; base < C; base += blockDim.x * 16) {
        const int c = base + threadIdx.x * 16;
        float acc = 0.f;
        if (c < C) {
            const uint4 o0 = __ldg(reinterpret_cast<const uint4*>(O + r * C + c));
            const uint4 o1 = __ldg(reinterpret_cast<const uint4*>(O + r * C + c + 8));
            const uint4 d0 = __ldg(reinterpret_cast<const uint4*>(dO + r * C + c));
            const uint4 d1 = __ldg(reinterpret_cast<const uint4*>(dO + r * C + c + 8));
            float fo[8], fd[8];
            unpack8(o0, fo);
            unpack8(d0, fd);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += fo[j] * fd[j];
            unpack8(o1, fo);
            unpack8(d1, fd);
#pragma unroll
            for (int j = 0; j < 8; ++j) acc += fo[j] * fd[j];
            if (dO_orig) {
                *reinterpret_cast<uint4*>(dO_orig + bt * C + c) = d0;
                *reinterpret_cast<uint4*>(dO_orig + bt * C + c + 8) = d1;
            }
        }
        for (int o = tph >> 1; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
        if (c < C && (threadIdx.x % tph) == 0) {
            const int h = c / D;
            delta[((int64_t)b * heads + h) * seqlen + t] = acc;
        }
    }
}

// ------------------------------------------------ fused cross-entropy
// One CTA per row; online (max, sum) pass, then gradient written in place.
__global__ void __launch_bounds__(256) cross_entropy_kernel(__nv_bfloat16* __restrict__ logits, int64_t ld,
                                                            const int64_t* __restrict__ labels,
                                                            float* __restrict__ row_loss, int vocab, float grad_scale) {
    __shared__ float red_m[8], red_s[8];
    const int64_t r = blockIdx.x;
    __nv_bfloat16* row = logits + r * ld;
    const int64_t label = labels[r];
    const bool vec = ((ld & 7) == 0) && ((reinterpret_cast<uintptr_t>(logits) & 15) == 0);
    const int nvec = vec ? (vocab >> 3) : 0;
    float m = -CUDART_INF_F, s = 0.f;
    // read the target logit BEFORE any thread overwrites the row with its gradient (the block-wide reduction below
    // orders this read against those writes)
    const float x_label = (label >= 0) ? __bfloat162float(row[label]) : 0.f;
    if (label >= 0) {     // ignored rows only need a zero gradient
        for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
            float f[8];
            unpack8(reinterpret_cast<const uint4*>(row)[v], f);
            float lm = f[0];
#pragma unroll
            for (int j = 1; j < 8; ++j) lm = fmaxf(lm, f[j]);
            if (lm > m) {
                s *= __expf(m - lm);
                m = lm;
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s += __expf(f[j] - m);
        }
        for (int c = nvec * 8 + threadIdx.x; c < vocab; c += blockDim.x) {
            const float f = __bfloat162float(row[c]);
            if (f > m) {
                s *= __expf(m - f);
                m = f;
            }
            s += __expf(f - m);
        }
        // block combine
        float wm = warp_max(m);
        s *= (m == -CUDART_INF_F) ? 0.f : __expf(m - wm);
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) {
            red_m[threadIdx.x >> 5] = wm;
            red_s[threadIdx.x >> 5] = s;
        }
        __syncthreads();
        float gm = red_m[0];
#pragma unroll
        for (int i = 1; i < 8; ++i) gm = fmaxf(gm, red_m[i]);
        float gs = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) gs += red_s[i] * ((red_m[i] == -CUDART_INF_F) ? 0.f : __expf(red_m[i] - gm));
        m = gm;
        s = gs;
    }
    const float lse = m + __logf(s);
    if (threadIdx.x == 0) row_loss[r] = (label >= 0) ? (lse - x_label) : 0.f;
    const float inv = (label >= 0) ? grad_scale / s : 0.f;
    for (int v = threadIdx.x; v < nvec; v += blockDim.x) {
        float f[8];
        unpack8(reinterpret_cast<const uint4*>(row)[v], f);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int c = v * 8 + j;
            f[j] = (label >= 0) ? (__expf(f[j] - m) * inv - ((int64_t)c == label ? grad_scale : 0.f)) : 0.f;
        }
        reinterpret_cast<uint4*>(row)[v] = pack8(f);
    }
    for (int c = nvec * 8 + threadIdx.x; c < vocab; c += blockDim.x) {
        const float f = __bfloat162float(row[c]);
        const float g = (label >= 0) ? (__expf(f - m) * inv - ((int64_t)c == label ? grad_scale : 0.f)) : 0.f;
        row[c] = __float2bfloat16(g);
    }
}

// AdamW's per-element divisions and square root with the hardware approximations (MUFU.SQRT / MUFU.RCP, <= 2 ulp in fp32):
// the parameter and both moments are rounded to bf16 (8 mantissa bits) right after, and with the IEEE sequences (~35 of the
// ~50 instructions per element) the update was ALU-bound at the power-capped clock of a training step: 8.9 ms inside the step
// against 7.3 ms alone for the 8-layer model.
__device__ __forceinline__ float sqrt_approx(float x) {
    float y;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// ------------------------------------------------ fused AdamW over flat bf16 buffers
// torch.optim.AdamW semantics (decoupled weight decay, bias-corrected moments), fp32 maths, bf16 storage of p, m, v.
__global__ void __launch_bounds__(256) adamw_bf16_kernel(__nv_bfloat16* __restrict__ p, const __nv_bfloat16* __restrict__ g,
                                                         __nv_bfloat16* __restrict__ m, __nv_bfloat16* __restrict__ v,
                                                         int64_t nvec, float lr, float beta1, float beta2, float eps,
                                                         float weight_decay, float bc1, float bc2_sqrt) {
    const float inv_bc2_sqrt = 1.f / bc2_sqrt, step_size = lr / bc1;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float fp[8], fg[8], fm[8], fv[8];
        unpack8(reinterpret_cast<const uint4*>(p)[i], fp);
        unpack8(__ldg(reinterpret_cast<const uint4*>(g) + i), fg);
        unpack8(reinterpret_cast<const uint4*>(m)[i], fm);
        unpack8(reinterpret_cast<const uint4*>(v)[i], fv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            fp[j] *= (1.f - lr * weight_decay);
            fm[j] = beta1 * fm[j] + (1.f - beta1) * fg[j];
            fv[j] = beta2 * fv[j] + (1.f - beta2) * fg[j] * fg[j];
            const float denom = sqrt_approx(fv[j]) * inv_bc2_sqrt + eps;
            fp[j] -= step_size * __fdividef(fm[j], denom);
        }
        reinterpret_cast<uint4*>(p)[i] = pack8(fp);
        reinterpret_cast<uint4*>(m)[i] = pack8(fm);
        reinterpret_cast<uint4*>(v)[i] = pack8(fv);
    }
}

// same, the gradient multiplied by a device-resident scalar first (gradient clipping without a pass over the buffer)
__global__ void __launch_bounds__(256) adamw_bf16_scaled_kernel(__nv_bfloat16* __restrict__ p, const __nv_bfloat16* __restrict__ g,
                                                                __nv_bfloat16* __restrict__ m, __nv_bfloat16* __restrict__ v,
                                                                int64_t nvec, float lr, float beta1, float beta2, float eps,
                                                                float weight_decay, float bc1, float bc2_sqrt,
                                                                const float* __restrict__ grad_scale,
                                                                const int64_t* __restrict__ nodecay, int n_nodecay) {
    const float gs = grad_scale ? __ldg(grad_scale) : 1.f;
    const float inv_bc2_sqrt = 1.f / bc2_sqrt, step_size = lr / bc1;
    // weight decay applies outside the sorted [lo, hi) vector ranges of `nodecay` (norm weights, biases).  A thread's index only
    // grows, so it keeps the end of the constant-decay interval it is in and searches the table again only after leaving it
    // (intervals are millions of vectors long; the grid stride is ~600 k vectors)
    float wd = weight_decay;
    int64_t seg_end = n_nodecay > 0 ? -1 : INT64_MAX;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        if (i >= seg_end) {
            int lo = 0, hi = n_nodecay;
            while (lo < hi) {                            // first range whose end is beyond i
                const int mid = (lo + hi) >> 1;
                if (__ldg(nodecay + 2 * mid + 1) <= i) lo = mid + 1; else hi = mid;
            }
            const int64_t r_lo = lo < n_nodecay ? __ldg(nodecay + 2 * lo) : INT64_MAX;
            if (r_lo <= i) {
                wd = 0.f;
                seg_end = __ldg(nodecay + 2 * lo + 1);
            } else {
                wd = weight_decay;
                seg_end = r_lo;
            }
        }
        float fp[8], fg[8], fm[8], fv[8];
        unpack8(reinterpret_cast<const uint4*>(p)[i], fp);
        unpack8(__ldg(reinterpret_cast<const uint4*>(g) + i), fg);
        unpack8(reinterpret_cast<const uint4*>(m)[i], fm);
        unpack8(reinterpret_cast<const uint4*>(v)[i], fv);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const float gj = fg[j] * gs;
            fp[j] *= (1.f - lr * wd);
            fm[j] = beta1 * fm[j] + (1.f - beta1) * gj;
            fv[j] = beta2 * fv[j] + (1.f - beta2) * gj * gj;
            const float denom = sqrt_approx(fv[j]) * inv_bc2_sqrt + eps;
            fp[j] -= step_size * __fdividef(fm[j], denom);
        }
        reinterpret_cast<uint4*>(p)[i] = pack8(fp);
        reinterpret_cast<uint4*>(m)[i] = pack8(fm);
        reinterpret_cast<uint4*>(v)[i] = pack8(fv);
    }
}

// ------------------------------------------------ global gradient norm -> clip factor, all on the device
// two deterministic stages (no atomics): per-CTA partial sums of squares, then one CTA folds them and writes
// out[0] = ||g||_2, out[1] = min(1, max_norm / (||g|| + 1e-6))   (torch.nn.utils.clip_grad_norm_ semantics)
__global__ void __launch_bounds__(256) sumsq_partial_kernel(const __nv_bfloat16* __restrict__ x, int64_t nvec,
                                                            float* __restrict__ partial) {
    __shared__ float red[8];
    float s = 0.f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        float f[8];
        unpack8(__ldg(reinterpret_cast<const uint4*>(x) + i), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s = fmaf(f[j], f[j], s);
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += red[i];
        partial[blockIdx.x] = t;
    }
}
__global__ void __launch_bounds__(256) clip_scale_kernel(const float* __restrict__ partial, int n, float max_norm,
                                                         float* __restrict__ out) {
    __shared__ double red[8];
    double s = 0.0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) s += (double)partial[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; ++i) t += red[i];
        const float norm = (float)sqrt(t);
        out[0] = norm;
        out[1] = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
    }
}

static int ew_grid(int64_t work_items, int threads) {
    int64_t g = (work_items + threads - 1) / threads;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return (int)g;
}

}  // namespace lb

using namespace lb;

#define AL16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" {

int lb_swiglu_fwd(const void* gate, const void* up, void* out, int64_t rows, int cols, int64_t ld_gate, int64_t ld_up,
                  int64_t ld_out, void* stream) {
    LB_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0, LB_EINVAL, "swiglu: cols=%d must be a positive multiple of 8", cols);
    LB_REQUIRE(ld_gate % 8 == 0 && ld_up % 8 == 0 && ld_out % 8 == 0 && AL16(gate) && AL16(up) && AL16(out), LB_EALIGN,
               "swiglu: pointers/pitches must be 16-byte aligned");
    if (rows == 0) return LB_OK;
    const int nvec = cols / 8;
    swiglu_fwd_kernel<<<ew_grid(rows * nvec, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)gate, (const __nv_bfloat16*)up, (__nv_bfloat16*)out, rows, nvec, ld_gate, ld_up, ld_out);
    return check_launch("swiglu_fwd");
}

int lb_swiglu_bwd(const void* dout, const void* gate, const void* up, void* dgate, void* dup, int64_t rows, int cols,
                  int64_t ld_dout, int64_t ld_gate, int64_t ld_up, int64_t ld_dgate, int64_t ld_dup, void* stream) {
    LB_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0, LB_EINVAL, "swiglu: cols=%d must be a positive multiple of 8", cols);
    LB_REQUIRE(ld_dout % 8 == 0 && ld_gate % 8 == 0 && ld_up % 8 == 0 && ld_dgate % 8 == 0 && ld_dup % 8 == 0 &&
                   AL16(dout) && AL16(gate) && AL16(up) && AL16(dgate) && AL16(dup),
               LB_EALIGN, "swiglu_bwd: pointers/pitches must be 16-byte aligned");
    if (rows == 0) return LB_OK;
    const int nvec = cols / 8;
    swiglu_bwd_kernel<<<ew_grid(rows * nvec, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dout, (const __nv_bfloat16*)gate, (const __nv_bfloat16*)up, (__nv_bfloat16*)dgate,
        (__nv_bfloat16*)dup, rows, nvec, ld_dout, ld_gate, ld_up, ld_dgate, ld_dup);
    return check_launch("swiglu_bwd");
}

int lb_bias_quick_gelu_fwd(const void* x, const void* bias, void* y, int64_t rows, int cols, void* stream) {
    LB_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0, LB_EINVAL, "quick_gelu: cols=%d must be a positive multiple of 8", cols);
    LB_REQUIRE(AL16(x) && AL16(y) && (!bias || AL16(bias)), LB_EALIGN, "quick_gelu: pointers must be 16-byte aligned");
    if (rows == 0) return LB_OK;
    const int nvec = cols / 8;
    bias_qgelu_fwd_kernel<<<ew_grid(rows * nvec, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)bias, (__nv_bfloat16*)y, rows, nvec);
    return check_launch("bias_quick_gelu_fwd");
}

int lb_bias_quick_gelu_bwd(const void* dy, const void* x, const void* bias, void* dx, int64_t rows, int cols,
                           void* stream) {
    LB_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0, LB_EINVAL, "quick_gelu: cols=%d must be a positive multiple of 8", cols);
    LB_REQUIRE(AL16(x) && AL16(dy) && AL16(dx) && (!bias || AL16(bias)), LB_EALIGN,
               "quick_gelu_bwd: pointers must be 16-byte aligned");
    if (rows == 0) return LB_OK;
    const int nvec = cols / 8;
    bias_qgelu_bwd_kernel<<<ew_grid(rows * nvec, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)bias, (__nv_bfloat16*)dx, rows, nvec);
    return check_launch("bias_quick_gelu_bwd");
}

int lb_gather_rows(const void* src, const int32_t* index, void* dst, int64_t rows, int cols, void* stream) {
    LB_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0, LB_EINVAL, "gather_rows: cols=%d must be a positive multiple of 8", cols);
    LB_REQUIRE(AL16(src) && AL16(dst), LB_EALIGN, "gather_rows: pointers must be 16-byte aligned");
    if (rows == 0) return LB_OK;
    const int nvec = cols / 8;
    gather_rows_kernel<int32_t><<<ew_grid(rows * nvec, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)src, index, (__nv_bfloat16*)dst, rows, nvec, cols, 0);
    return check_launch("gather_rows");
}

int lb_embed_lang_fwd(const int64_t* ids, const void* table, void* out, int64_t rows, int cols, void* stream) {
    LB_REQUIRE(rows >= 0 && cols > 0 && cols % 8 == 0, LB_EINVAL, "embed: cols=%d must be a positive multiple of 8", cols);
    LB_REQUIRE(AL16(table) && AL16(out), LB_EALIGN, "embed: pointers must be 16-byte aligned");
    if (rows == 0) return LB_OK;
    const int nvec = cols / 8;
    gather_rows_kernel<int64_t><<<ew_grid(rows * nvec, 256), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)table, ids, (__nv_bfloat16*)out, rows, nvec, cols, 0);
    return check_launch("embed_lang_fwd");
}

int lb_embed_vision_cat_fwd(const int64_t* ids0, const int64_t* ids1, const void* table0, const void* table1,
                            const void* signal, const int32_t* signal_row, void* out, int64_t rows, int half,
                            int signal_cols, void* stream) {
    LB_REQUIRE(rows >= 0 && half > 0 && half % 8 == 0 && signal_cols >= 0 && signal_cols % 8 == 0, LB_EINVAL,
               "embed_vision: half=%d signal_cols=%d must be multiples of 8", half, signal_cols);
    LB_REQUIRE(AL16(table0) && AL16(table1) && AL16(out) && (!signal || AL16(signal)), LB_EALIGN,
               "embed_vision: pointers must be 16-byte aligned");
    if (rows == 0) return LB_OK;
    cudaStream_t st = (cudaStream_t)stream;
    const int64_t ld = 2 * (int64_t)half + signal_cols;
    const int hv = half / 8;
    gather_rows_kernel<int64_t><<<ew_grid(rows * hv, 256), 256, 0, st>>>((const __nv_bfloat16*)table0, ids0,
                                                                         (__nv_bfloat16*)out, rows, hv, ld, 0);
    gather_rows_kernel<int64_t><<<ew_grid(rows * hv, 256), 256, 0, st>>>((const __nv_bfloat16*)table1, ids1,
                                                                         (__nv_bfloat16*)out, rows, hv, ld, hv);
    if (signal_cols > 0) {
        const int sv = signal_cols / 8;
        if (signal) {
            gather_rows_kernel<int32_t><<<ew_grid(rows * sv, 256), 256, 0, st>>>(
                (const __nv_bfloat16*)signal, signal_row, (__nv_bfloat16*)out, rows, sv, ld, 2 * hv);
        } else {
            cudaError_t e = cudaMemset2DAsync((char*)out + 4 * (size_t)half, (size_t)ld * 2, 0, (size_t)signal_cols * 2,
                                              (size_t)rows, st);
            if (e != cudaSuccess) return fail(LB_ELAUNCH, "embed_vision memset: %s", cudaGetErrorString(e));
        }
    }
    return check_launch("embed_vision_cat_fwd");
}

int lb_embed_bwd(const int64_t* ids, const void* dy, int64_t ld_dy, int col0, float* dtable, int64_t rows, int cols,
                 void* stream) {
    LB_REQUIRE(rows >= 0 && cols > 0 && ids && dy && dtable, LB_EINVAL, "embed_bwd: bad arguments");
    if (rows == 0) return LB_OK;
    embed_bwd_kernel<<<ew_grid(rows * cols, 256), 256, 0, (cudaStream_t)stream>>>(ids, (const __nv_bfloat16*)dy, ld_dy,
                                                                                 col0, dtable, rows, cols);
    return check_launch("embed_bwd");
}

int lb_lfq_pack(const void* h, int dtype, int64_t n_img, int tokens, int num_codebooks, int bits, int64_t offset,
                int64_t boi, int64_t eoi, int64_t* ids, void* stream) {
    LB_REQUIRE(n_img >= 0 && tokens > 0 && num_codebooks > 0 && bits > 0 && bits < 31 && h && ids, LB_EINVAL,
               "lfq_pack: bad arguments");
    LB_REQUIRE(dtype == LB_DT_BF16 || dtype == LB_DT_F32, LB_EDTYPE, "lfq_pack: dtype %d", dtype);
    if (n_img == 0) return LB_OK;
    const int64_t total = n_img * (tokens + 2) * num_codebooks;
    if (dtype == LB_DT_BF16)
        lfq_pack_kernel<__nv_bfloat16><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
            (const __nv_bfloat16*)h, n_img, tokens, num_codebooks, bits, offset, boi, eoi, ids);
    else
        lfq_pack_kernel<float><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>((const float*)h, n_img, tokens,
                                                                                      num_codebooks, bits, offset, boi,
                                                                                      eoi, ids);
    return check_launch("lfq_pack");
}

int lb_lfq_unpack(const int64_t* idx, int64_t n, int num_codebooks, int bits, void* codes, int dtype, void* stream) {
    LB_REQUIRE(n >= 0 && num_codebooks > 0 && bits > 0 && bits < 31 && idx && codes, LB_EINVAL, "lfq_unpack: bad arguments");
    LB_REQUIRE(dtype == LB_DT_BF16 || dtype == LB_DT_F32, LB_EDTYPE, "lfq_unpack: dtype %d", dtype);
    if (n == 0) return LB_OK;
    const int64_t total = n * num_codebooks * bits;
    if (dtype == LB_DT_BF16)
        lfq_unpack_kernel<__nv_bfloat16><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(
            idx, n, num_codebooks, bits, (__nv_bfloat16*)codes);
    else
        lfq_unpack_kernel<float><<<ew_grid(total, 256), 256, 0, (cudaStream_t)stream>>>(idx, n, num_codebooks, bits,
                                                                                       (float*)codes);
    return check_launch("lfq_unpack");
}

int lb_attn_prep_fwd(const void* q, const void* k, const void* kc, const void* v, const void* vc,
                     const uint8_t* flag_sorted, const int32_t* sorted_of, const int32_t* pos, const float* cos_t,
                     const float* sin_t, void* Q, void* Kfv, void* Kfl, void* Vfv, void* Vfl, int64_t n_tokens, int heads,
                     int head_dim, const int32_t* kv_row, void* stream) {
    LB_REQUIRE(n_tokens >= 0 && heads > 0 && head_dim >= 16 && head_dim % 16 == 0, LB_EINVAL,
               "attn_prep: head_dim=%d must be a multiple of 16", head_dim);
    LB_REQUIRE(q && k && v && flag_sorted && sorted_of && pos && cos_t && sin_t && Q && Kfv && Kfl && Vfv && Vfl,
               LB_EINVAL, "attn_prep: null argument");
    LB_REQUIRE(AL16(q) && AL16(k) && AL16(v) && (!kc || AL16(kc)) && (!vc || AL16(vc)) && AL16(Q) && AL16(Kfv) && AL16(Kfl) &&
                   AL16(Vfv) && AL16(Vfl) && AL16(cos_t) && AL16(sin_t),
               LB_EALIGN, "attn_prep: pointers must be 16-byte aligned");
    if (n_tokens == 0) return LB_OK;
    launch_chain(attn_prep_fwd_kernel<false>, dim3((unsigned)n_tokens), dim3(256), 0, (cudaStream_t)stream,
                 (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)kc, (const __nv_bfloat16*)v,
                 (const __nv_bfloat16*)vc, flag_sorted, sorted_of, pos, cos_t, sin_t, (__nv_bfloat16*)Q, (__nv_bfloat16*)Kfv,
                 (__nv_bfloat16*)Kfl, (__nv_bfloat16*)Vfv, (__nv_bfloat16*)Vfl, heads, head_dim, kv_row, PrepBridge{});
    return check_launch("attn_prep_fwd");
}

int lb_attn_prep_fwd_bridge(const void* q, const void* k, const void* v, const void* tk, const void* tv, const void* Bk_lang,
                            const void* Bk_vis, const void* Bv_lang, const void* Bv_vis, int rank, const uint8_t* flag_sorted,
                            const int32_t* sorted_of, const int32_t* pos, const float* cos_t, const float* sin_t, void* Q, void* Kfv,
                            void* Kfl, void* Vfv, void* Vfl, int64_t n_tokens, int heads, int head_dim, const int32_t* kv_row,
                            void* stream) {
    LB_REQUIRE(n_tokens >= 0 && heads > 0 && head_dim >= 16 && head_dim % 16 == 0, LB_EINVAL,
               "attn_prep_bridge: head_dim=%d must be a multiple of 16", head_dim);
    LB_REQUIRE(rank > 0 && rank % 8 == 0, LB_EINVAL, "attn_prep_bridge: rank=%d must be a positive multiple of 8", rank);
    LB_REQUIRE(q && k && v && tk && tv && Bk_lang && Bk_vis && Bv_lang && Bv_vis && flag_sorted && sorted_of && pos && cos_t &&
                   sin_t && Q && Kfv && Kfl && Vfv && Vfl, LB_EINVAL, "attn_prep_bridge: null argument");
    LB_REQUIRE(AL16(q) && AL16(k) && AL16(v) && AL16(tk) && AL16(tv) && AL16(Bk_lang) && AL16(Bk_vis) && AL16(Bv_lang) &&
                   AL16(Bv_vis) && AL16(Q) && AL16(Kfv) && AL16(Kfl) && AL16(Vfv) && AL16(Vfl) && AL16(cos_t) && AL16(sin_t),
               LB_EALIGN, "attn_prep_bridge: pointers must be 16-byte aligned");
    if (n_tokens == 0) return LB_OK;
    PrepBridge br;
    br.tk = (const __nv_bfloat16*)tk; br.tv = (const __nv_bfloat16*)tv;
    br.Bk[0] = (const __nv_bfloat16*)Bk_lang; br.Bk[1] = (const __nv_bfloat16*)Bk_vis;
    br.Bv[0] = (const __nv_bfloat16*)Bv_lang; br.Bv[1] = (const __nv_bfloat16*)Bv_vis;
    br.rank = rank;
    launch_chain(attn_prep_fwd_kernel<true>, dim3((unsigned)n_tokens), dim3(256), 0, (cudaStream_t)stream,
                 (const __nv_bfloat16*)q, (const __nv_bfloat16*)k, (const __nv_bfloat16*)nullptr, (const __nv_bfloat16*)v,
                 (const __nv_bfloat16*)nullptr, flag_sorted, sorted_of, pos, cos_t, sin_t, (__nv_bfloat16*)Q, (__nv_bfloat16*)Kfv,
                 (__nv_bfloat16*)Kfl, (__nv_bfloat16*)Vfv, (__nv_bfloat16*)Vfl, heads, head_dim, kv_row, br);
    return check_launch("attn_prep_fwd_bridge");
}

int lb_attn_prep_bwd(const void* dQ, const void* dKfv, const void* dKfl, const void* dVfv, const void* dVfl,
                     const uint8_t* flag_sorted, const int32_t* sorted_of, const int32_t* pos, const float* cos_t,
                     const float* sin_t, void* dq, void* dk, void* dv, void* dkb, void* dvb, int64_t n_tokens, int heads,
                     int head_dim, void* stream) {
    LB_REQUIRE(n_tokens >= 0 && heads > 0 && head_dim >= 16 && head_dim % 16 == 0, LB_EINVAL,
               "attn_prep_bwd: head_dim=%d must be a multiple of 16", head_dim);
    LB_REQUIRE(dQ && dKfv && dKfl && dVfv && dVfl && flag_sorted && sorted_of && pos && cos_t && sin_t && dq && dk && dv,
               LB_EINVAL, "attn_prep_bwd: null argument");
    LB_REQUIRE(AL16(dQ) && AL16(dKfv) && AL16(dKfl) && AL16(dVfv) && AL16(dVfl) && AL16(dq) && AL16(dk) && AL16(dv) &&
                   (!dkb || AL16(dkb)) && (!dvb || AL16(dvb)),
               LB_EALIGN, "attn_prep_bwd: pointers must be 16-byte aligned");
    if (n_tokens == 0) return LB_OK;
    attn_prep_bwd_kernel<<<(unsigned)n_tokens, 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)dQ, (const __nv_bfloat16*)dKfv, (const __nv_bfloat16*)dKfl, (const __nv_bfloat16*)dVfv,
        (const __nv_bfloat16*)dVfl, flag_sorted, sorted_of, pos, cos_t, sin_t, (__nv_bfloat16*)dq, (__nv_bfloat16*)dk,
        (__nv_bfloat16*)dv, (__nv_bfloat16*)dkb, (__nv_bfloat16*)dvb, heads, head_dim);
    return check_launch("attn_prep_bwd");
}

int lb_attn_bwd_prepare(const void* O, const void* dO, const int32_t* row_of, void* dO_orig, float* delta, int batch,
                        int seqlen, int heads, int head_dim, void* stream) {
    LB_REQUIRE(batch >= 0 && seqlen > 0 && heads > 0, LB_EINVAL, "attn_bwd_prepare: bad shape");
    LB_REQUIRE(head_dim == 16 || head_dim == 32 || head_dim == 64 || head_dim == 128 || head_dim == 256, LB_EINVAL,
               "attn_bwd_prepare: head_dim=%d unsupported", head_dim);
    LB_REQUIRE(O && dO && delta && AL16(O) && AL16(dO) && (!dO_orig || AL16(dO_orig)), LB_EALIGN,
               "attn_bwd_prepare: null/unaligned argument");
    if (batch == 0) return LB_OK;
    attn_bwd_prepare_kernel<<<(unsigned)((int64_t)batch * seqlen), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)O, (const __nv_bfloat16*)dO, row_of, (__nv_bfloat16*)dO_orig, delta, seqlen, heads, head_dim);
    return check_launch("attn_bwd_prepare");
}

int lb_adamw_bf16(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, int64_t n, float lr, float beta1, float beta2,
                  float eps, float weight_decay, int step, void* stream) {
    LB_REQUIRE(n >= 0 && n % 8 == 0 && step >= 1, LB_EINVAL, "adamw: n=%lld must be a multiple of 8 and step >= 1", (long long)n);
    LB_REQUIRE(param && grad && exp_avg && exp_avg_sq && AL16(param) && AL16(grad) && AL16(exp_avg) && AL16(exp_avg_sq), LB_EALIGN,
               "adamw: null/unaligned pointer");
    if (n == 0) return LB_OK;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    const int64_t nvec = n / 8;
    adamw_bf16_kernel<<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)param, (const __nv_bfloat16*)grad,
                                                                          (__nv_bfloat16*)exp_avg, (__nv_bfloat16*)exp_avg_sq,
                                                                          nvec, lr, beta1, beta2, eps, weight_decay, bc1,
                                                                          bc2_sqrt);
    return check_launch("adamw_bf16");
}

int lb_adamw_bf16_scaled(void* param, const void* grad, void* exp_avg, void* exp_avg_sq, int64_t n, float lr, float beta1,
                         float beta2, float eps, float weight_decay, int step, const float* grad_scale,
                         const int64_t* nodecay_ranges, int n_nodecay, void* stream) {
    LB_REQUIRE(n >= 0 && n % 8 == 0 && step >= 1, LB_EINVAL, "adamw: n=%lld must be a multiple of 8, step >= 1", (long long)n);
    LB_REQUIRE(AL16(param) && AL16(grad) && AL16(exp_avg) && AL16(exp_avg_sq), LB_EALIGN, "adamw: buffers must be 16-byte aligned");
    if (n == 0) return LB_OK;
    const float bc1 = 1.f - powf(beta1, (float)step);
    const float bc2_sqrt = sqrtf(1.f - powf(beta2, (float)step));
    const int64_t nvec = n / 8;
    adamw_bf16_scaled_kernel<<<ew_grid(nvec, 256), 256, 0, (cudaStream_t)stream>>>(
        (__nv_bfloat16*)param, (const __nv_bfloat16*)grad, (__nv_bfloat16*)exp_avg, (__nv_bfloat16*)exp_avg_sq, nvec, lr, beta1,
        beta2, eps, weight_decay, bc1, bc2_sqrt, grad_scale, nodecay_ranges, nodecay_ranges ? n_nodecay : 0);
    return check_launch("adamw_bf16_scaled");
}

int lb_grad_clip_scale(const void* grad, int64_t n, float max_norm, float* workspace, int workspace_floats, float* out2,
                       void* stream) {
    LB_REQUIRE(n >= 0 && n % 8 == 0 && grad && workspace && out2 && workspace_floats >= 64, LB_EINVAL,
               "grad_clip_scale: n=%lld must be a multiple of 8, workspace >= 64 floats", (long long)n);
    LB_REQUIRE(AL16(grad), LB_EALIGN, "grad_clip_scale: gradient buffer must be 16-byte aligned");
    int grid = ew_grid(n / 8, 256);
    if (grid > workspace_floats) grid = workspace_floats;
    sumsq_partial_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)grad, n / 8, workspace);
    clip_scale_kernel<<<1, 256, 0, (cudaStream_t)stream>>>(workspace, grid, max_norm, out2);
    return check_launch("grad_clip_scale");
}

int lb_cross_entropy_fwd_bwd(void* logits, int64_t ld, const int64_t* labels, float* row_loss, int64_t rows, int vocab,
                             float grad_scale, void* stream) {
    LB_REQUIRE(rows >= 0 && vocab > 0 && ld >= vocab && logits && labels && row_loss, LB_EINVAL,
               "cross_entropy: bad arguments");
    if (rows == 0) return LB_OK;
    cross_entropy_kernel<<<(unsigned)rows, 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)logits, ld, labels, row_loss,
                                                                          vocab, grad_scale);
    return check_launch("cross_entropy_fwd_bwd");
}

}

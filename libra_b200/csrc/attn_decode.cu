// KV-cached decode step of the bridge attention (N1; LibraAttention.forward with past_key_value, modeling_libra.py:343-397):
// one new query per sample against the cached keys/values.  HBM-bound: every cached K and V row of the query's variant is
// read exactly once (2 * T * H * D * 2 bytes per sample), so the kernel is laid out for coalesced streaming, not for the
// tensor cores (at q_len = 1 the contraction is a GEMV).
//
// The cache holds, per layer, the same four operand tensors the training kernels consume (see lb_attn_prep_fwd): Kfl/Vfl
// (what LANGUAGE queries see) and Kfv/Vfv (what VISION queries see), [B, capacity, H*D] bf16, token-major.  The new
// token's modality picks the pair -- the bridge select costs nothing here.
//
//   lb_attn_decode:  grid = (B*H, n_split).  A CTA of 256 threads takes a contiguous chunk of the sample's visible keys
//     [kv_start, kv_end): thread = key for the scores (row of 2*D bytes, 16-byte loads, q broadcast from shared memory),
//     block max / sum, then thread = (key quarter, 2 value columns) for P.V (a warp reads 128 contiguous bytes of one V
//     row).  Writes the chunk's (max, sum, unnormalised O[D]) in fp32 to the caller's workspace.
//   combine (same call):  grid = B*H, merges the n_split partials, normalises, writes bf16 O [B, H*D] (row out_row[b]).
#include <math_constants.h>

#include "common.cuh"

namespace lb {
namespace dec {

constexpr int THREADS = 256;
constexpr int CHUNK = 256;                 // keys per CTA pass
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
    const __nv_bfloat16* q;       // [B, H*D]
    const __nv_bfloat16* k[2];    // variant 0 (language queries), 1 (vision queries): [B, capacity, H*D]
    const __nv_bfloat16* v[2];
    const uint8_t* qflag;         // [B] variant of the sample's query, or null (all 0)
    const int32_t* kv_start;      // [B] or null (0)
    const int32_t* kv_end;        // [B] or null (kv_len)
    const int32_t* out_row;       // [B] or null
    float* partial;               // [B*H, n_split, D + 2]
    __nv_bfloat16* out;           // [B, H*D]
    int batch, heads, capacity, kv_len, n_split;
    float scale;
};

__device__ __forceinline__ float block_reduce(float v, float* red, bool is_max) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    v = is_max ? warp_max(v) : warp_sum(v);
    __syncthreads();                       // red may still be read from the previous reduction
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = red[0];
#pragma unroll
    for (int i = 1; i < THREADS / 32; ++i) r = is_max ? fmaxf(r, red[i]) : r + red[i];
    return r;
}

template <int D>
__global__ void __launch_bounds__(THREADS) attn_decode_kernel(const Params p) {
    __shared__ float q_sh[D];
    __shared__ float p_sh[CHUNK];
    __shared__ float red[THREADS / 32];
    constexpr int COLS = D / 2, NG = THREADS / COLS;                 // column pairs; key groups for P.V
    __shared__ float o_sh[NG][D];
    const int bh = blockIdx.x, b = bh / p.heads, h = bh - b * p.heads;
    const int C = p.heads * D;
    const int variant = p.qflag ? (int)p.qflag[b] : 0;
    const int kvs = p.kv_start ? p.kv_start[b] : 0;
    const int kve = p.kv_end ? p.kv_end[b] : p.kv_len;
    const int n_keys = kve > kvs ? kve - kvs : 0;
    // this CTA's share of the visible keys (whole multiples of 8 keys so that splits stay sector aligned)
    const int per = ((n_keys + p.n_split - 1) / p.n_split + 7) & ~7;
    const int j0 = kvs + blockIdx.y * per, j1 = min(kve, j0 + per);
    const __nv_bfloat16* K = p.k[variant] + ((int64_t)b * p.capacity) * C + (int64_t)h * D;
    const __nv_bfloat16* V = p.v[variant] + ((int64_t)b * p.capacity) * C + (int64_t)h * D;
    if (threadIdx.x < D) q_sh[threadIdx.x] = __bfloat162float(p.q[(int64_t)b * C + h * D + threadIdx.x]) * p.scale * LOG2E;
    __syncthreads();

    float m_run = -CUDART_INF_F, l_run = 0.f;
    float acc0 = 0.f, acc1 = 0.f;                                  // my 2 value columns, my quarter of the chunk's keys
    const int grp_d = threadIdx.x / COLS, col_d = (threadIdx.x % COLS) * 2;
    for (int c0 = j0; c0 < j1; c0 += CHUNK) {
        const int j = c0 + (int)threadIdx.x;
        float s = -CUDART_INF_F;
        if (j < j1) {
            const uint4* row = reinterpret_cast<const uint4*>(K + (int64_t)j * C);
            float a = 0.f;
#pragma unroll
            for (int i = 0; i < D / 8; ++i) {
                const uint4 u = row[i];
                const float* qq = q_sh + i * 8;
                a = fmaf(bf16_lo(u.x), qq[0], a); a = fmaf(bf16_hi(u.x), qq[1], a);
                a = fmaf(bf16_lo(u.y), qq[2], a); a = fmaf(bf16_hi(u.y), qq[3], a);
                a = fmaf(bf16_lo(u.z), qq[4], a); a = fmaf(bf16_hi(u.z), qq[5], a);
                a = fmaf(bf16_lo(u.w), qq[6], a); a = fmaf(bf16_hi(u.w), qq[7], a);
            }
            s = a;                                                 // log2 domain, scaled
        }
        const float m_c = block_reduce(s, red, true);
        const float m_new = fmaxf(m_run, m_c);
        const float pj = (j < j1) ? fast_ex2(s - m_new) : 0.f;
        // the reference rounds the probabilities to the activation dtype before P.V (modeling_libra.py:391)
        p_sh[threadIdx.x] = __bfloat162float(__float2bfloat16(pj));
        const float l_c = block_reduce(pj, red, false);            // (also orders the p_sh writes before the reads below)
        const float alpha = (m_run == -CUDART_INF_F) ? 0.f : fast_ex2(m_run - m_new);
        l_run = l_run * alpha + l_c;
        acc0 *= alpha;
        acc1 *= alpha;
        m_run = m_new;
        const int n_here = min(CHUNK, j1 - c0);
        for (int t = grp_d; t < n_here; t += NG) {
            const __nv_bfloat162 vv = *reinterpret_cast<const __nv_bfloat162*>(V + (int64_t)(c0 + t) * C + col_d);
            const float pt = p_sh[t];
            acc0 = fmaf(pt, __bfloat162float(vv.x), acc0);
            acc1 = fmaf(pt, __bfloat162float(vv.y), acc1);
        }
    }
    // fold the key groups
    __syncthreads();
    if (grp_d > 0) {
        o_sh[grp_d][col_d] = acc0;
        o_sh[grp_d][col_d + 1] = acc1;
    }
    __syncthreads();
    if (grp_d == 0) {
        for (int g = 1; g < NG; ++g) {
            acc0 += o_sh[g][col_d];
            acc1 += o_sh[g][col_d + 1];
        }
        float* dst = p.partial + ((int64_t)bh * p.n_split + blockIdx.y) * (D + 2);
        dst[col_d] = acc0;
        dst[col_d + 1] = acc1;
        if (threadIdx.x == 0) {
            dst[D] = m_run;
            dst[D + 1] = l_run;
        }
    }
}

template <int D>
__global__ void __launch_bounds__(D) attn_decode_combine_kernel(const Params p) {
    const int bh = blockIdx.x, b = bh / p.heads, h = bh - b * p.heads;
    const float* src = p.partial + (int64_t)bh * p.n_split * (D + 2);
    float m = -CUDART_INF_F;
    for (int s = 0; s < p.n_split; ++s) m = fmaxf(m, src[s * (D + 2) + D]);
    float l = 0.f, o = 0.f;
    for (int s = 0; s < p.n_split; ++s) {
        const float ms = src[s * (D + 2) + D];
        const float w = (ms == -CUDART_INF_F) ? 0.f : fast_ex2(ms - m);
        l += w * src[s * (D + 2) + D + 1];
        o += w * src[s * (D + 2) + threadIdx.x];
    }
    const int64_t row = p.out_row ? p.out_row[b] : b;
    p.out[row * ((int64_t)p.heads * D) + h * D + threadIdx.x] = __float2bfloat16(l > 0.f ? o / l : 0.f);
}

}  // namespace dec
}  // namespace lb

using namespace lb;

extern "C" int lb_attn_decode_workspace_floats(int batch, int heads, int head_dim, int n_split) {
    return batch * heads * n_split * (head_dim + 2);
}

extern "C" int lb_attn_decode(const void* q, const void* K0, const void* V0, const void* K1, const void* V1,
                              const uint8_t* qflag, const int32_t* kv_start, const int32_t* kv_end, const int32_t* out_row,
                              float* workspace, void* out, int batch, int heads, int head_dim, int capacity, int kv_len,
                              int n_split, float scale, void* stream) {
    LB_REQUIRE(batch > 0 && heads > 0 && capacity > 0 && kv_len >= 0 && kv_len <= capacity && n_split > 0, LB_EINVAL,
               "attn_decode: bad shape");
    LB_REQUIRE(head_dim == 64 || head_dim == 128, LB_EINVAL, "attn_decode: head_dim %d (64 or 128 supported)", head_dim);
    LB_REQUIRE(q && K0 && V0 && workspace && out, LB_EINVAL, "attn_decode: null argument");
    LB_REQUIRE((((uintptr_t)K0 | (uintptr_t)V0 | (uintptr_t)(K1 ? K1 : K0) | (uintptr_t)(V1 ? V1 : V0)) & 15) == 0, LB_EALIGN,
               "attn_decode: cache tensors must be 16-byte aligned");
    int rc = require_sm100();
    if (rc) return rc;
    dec::Params p;
    p.q = (const __nv_bfloat16*)q;
    p.k[0] = (const __nv_bfloat16*)K0; p.v[0] = (const __nv_bfloat16*)V0;
    p.k[1] = (const __nv_bfloat16*)(K1 ? K1 : K0); p.v[1] = (const __nv_bfloat16*)(V1 ? V1 : V0);
    p.qflag = qflag; p.kv_start = kv_start; p.kv_end = kv_end; p.out_row = out_row;
    p.partial = workspace; p.out = (__nv_bfloat16*)out;
    p.batch = batch; p.heads = heads; p.capacity = capacity; p.kv_len = kv_len; p.n_split = n_split; p.scale = scale;
    cudaStream_t st = (cudaStream_t)stream;
    dim3 grid((unsigned)(batch * heads), (unsigned)n_split);
    if (head_dim == 128) {
        dec::attn_decode_kernel<128><<<grid, dec::THREADS, 0, st>>>(p);
        dec::attn_decode_combine_kernel<128><<<batch * heads, 128, 0, st>>>(p);
    } else {
        dec::attn_decode_kernel<64><<<grid, dec::THREADS, 0, st>>>(p);
        dec::attn_decode_combine_kernel<64><<<batch * heads, 64, 0, st>>>(p);
    }
    return check_launch("attn_decode");
}

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
//     [kv_start, kv_end); each of its 8 warps streams every 8th batch of keys with its own online softmax (D/8 lanes per
//     row, 16-byte loads, whole rows per instruction), the warps are merged through shared memory.  Writes the chunk's
//     (max, sum, unnormalised O[D]) in fp32 to the caller's workspace.
//   combine (same call):  grid = B*H, merges the n_split partials, normalises, writes bf16 O [B, H*D] (row out_row[b]).
#include <math_constants.h>

#include "common.cuh"

namespace lb {
namespace dec {

constexpr int THREADS = 256;
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

__device__ __forceinline__ float dot8(const uint4& u, const float (&q)[8]) {
    float a = bf16_lo(u.x) * q[0];
    a = fmaf(bf16_hi(u.x), q[1], a);
    a = fmaf(bf16_lo(u.y), q[2], a); a = fmaf(bf16_hi(u.y), q[3], a);
    a = fmaf(bf16_lo(u.z), q[4], a); a = fmaf(bf16_hi(u.z), q[5], a);
    a = fmaf(bf16_lo(u.w), q[6], a); a = fmaf(bf16_hi(u.w), q[7], a);
    return a;
}

// Each warp streams its own subset of the CTA's keys with an online softmax of its own: D/8 lanes cover one K (or V) row with
// one 16-byte load each, so a warp instruction reads 32/(D/8) whole rows = 512 contiguous-per-row bytes, NB instructions are
// in flight per batch, and nothing inside the loop needs a block-wide barrier.  The warps are merged through shared memory
// once at the end.  (The first version took one key per THREAD: 32 different rows per load instruction, 0.27-0.46 of the
// HBM peak.)
template <int D>
__global__ void __launch_bounds__(THREADS) attn_decode_kernel(const Params p) {
    constexpr int LPK = D / 8, KPI = 32 / LPK, NB = 4, NW = THREADS / 32;      // lanes per key, keys per instruction, batch depth
    __shared__ float m_sh[NW], l_sh[NW];
    __shared__ float o_sh[NW][D];
    pdl_wait();
    const int bh = blockIdx.x, b = bh / p.heads, h = bh - b * p.heads;
    const int C = p.heads * D;
    const int variant = p.qflag ? (int)p.qflag[b] : 0;
    const int kvs = p.kv_start ? p.kv_start[b] : 0;
    const int kve = p.kv_end ? p.kv_end[b] : p.kv_len;
    const int n_keys = kve > kvs ? kve - kvs : 0;
    // this CTA's share of the visible keys
    const int per = ((n_keys + p.n_split - 1) / p.n_split + 7) & ~7;
    const int j0 = kvs + blockIdx.y * per, j1 = min(kve, j0 + per);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, sub = lane / LPK, dc = lane % LPK;
    const __nv_bfloat16* K = p.k[variant] + ((int64_t)b * p.capacity) * C + (int64_t)h * D + dc * 8;
    const __nv_bfloat16* V = p.v[variant] + ((int64_t)b * p.capacity) * C + (int64_t)h * D + dc * 8;
    float qf[8];
    {
        const uint4 u = *reinterpret_cast<const uint4*>(p.q + (int64_t)b * C + h * D + dc * 8);
        const float sc = p.scale * LOG2E;                          // scores in the log2 domain
        qf[0] = bf16_lo(u.x) * sc; qf[1] = bf16_hi(u.x) * sc; qf[2] = bf16_lo(u.y) * sc; qf[3] = bf16_hi(u.y) * sc;
        qf[4] = bf16_lo(u.z) * sc; qf[5] = bf16_hi(u.z) * sc; qf[6] = bf16_lo(u.w) * sc; qf[7] = bf16_hi(u.w) * sc;
    }
    float m_run = -CUDART_INF_F, l_run = 0.f, acc[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[e] = 0.f;
    for (int jb = j0 + warp * (NB * KPI); jb < j1; jb += NW * NB * KPI) {
        uint4 kk[NB], vv[NB];
        float s[NB];
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int j = jb + i * KPI + sub;
            kk[i] = j < j1 ? *reinterpret_cast<const uint4*>(K + (int64_t)j * C) : make_uint4(0u, 0u, 0u, 0u);
        }
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const int j = jb + i * KPI + sub;
            vv[i] = j < j1 ? *reinterpret_cast<const uint4*>(V + (int64_t)j * C) : make_uint4(0u, 0u, 0u, 0u);
        }
        float m_new = m_run;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            float a = dot8(kk[i], qf);
#pragma unroll
            for (int o = LPK / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);     // the key's LPK lanes
            s[i] = (jb + i * KPI + sub < j1) ? a : -CUDART_INF_F;
            m_new = fmaxf(m_new, s[i]);
        }
#pragma unroll
        for (int o = LPK; o < 32; o <<= 1) m_new = fmaxf(m_new, __shfl_xor_sync(0xffffffffu, m_new, o));   // the other keys of the batch
        const float alpha = (m_run == -CUDART_INF_F) ? 0.f : fast_ex2(m_run - m_new);
        l_run *= alpha;
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] *= alpha;
        m_run = m_new;
#pragma unroll
        for (int i = 0; i < NB; ++i) {
            const float pj = (s[i] == -CUDART_INF_F) ? 0.f : fast_ex2(s[i] - m_new);
            l_run += pj;                                           // (every lane of the key's group counts it; only sub-group leaders are summed below)
            // the reference rounds the probabilities to the activation dtype before P.V (modeling_libra.py:391)
            const float pr = __bfloat162float(__float2bfloat16(pj));
            acc[0] = fmaf(pr, bf16_lo(vv[i].x), acc[0]); acc[1] = fmaf(pr, bf16_hi(vv[i].x), acc[1]);
            acc[2] = fmaf(pr, bf16_lo(vv[i].y), acc[2]); acc[3] = fmaf(pr, bf16_hi(vv[i].y), acc[3]);
            acc[4] = fmaf(pr, bf16_lo(vv[i].z), acc[4]); acc[5] = fmaf(pr, bf16_hi(vv[i].z), acc[5]);
            acc[6] = fmaf(pr, bf16_lo(vv[i].w), acc[6]); acc[7] = fmaf(pr, bf16_hi(vv[i].w), acc[7]);
        }
    }
    // successors of the dependent-launch chain (combine, then the o-proj GEMM, whose parked CTAs would take registers away
    // from this bandwidth-bound kernel) may become resident only now, while this CTA merges its warps
    pdl_trigger();
    // fold the key sub-groups of the warp (same running max in every lane)
#pragma unroll
    for (int o = LPK; o < 32; o <<= 1) {
        l_run += __shfl_xor_sync(0xffffffffu, l_run, o);
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] += __shfl_xor_sync(0xffffffffu, acc[e], o);
    }
    if (sub == 0) {
#pragma unroll
        for (int e = 0; e < 8; ++e) o_sh[warp][dc * 8 + e] = acc[e];
    }
    if (lane == 0) {
        m_sh[warp] = m_run;
        l_sh[warp] = l_run;
    }
    __syncthreads();
    if (threadIdx.x < D) {
        float m = -CUDART_INF_F;
#pragma unroll
        for (int w = 0; w < NW; ++w) m = fmaxf(m, m_sh[w]);
        float l = 0.f, o = 0.f;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const float sc = (m_sh[w] == -CUDART_INF_F) ? 0.f : fast_ex2(m_sh[w] - m);
            l += sc * l_sh[w];
            o += sc * o_sh[w][threadIdx.x];
        }
        float* dst = p.partial + ((int64_t)bh * p.n_split + blockIdx.y) * (D + 2);
        dst[threadIdx.x] = o;
        if (threadIdx.x == 0) {
            dst[D] = m;
            dst[D + 1] = l;
        }
    }
}

template <int D>
__global__ void __launch_bounds__(D) attn_decode_combine_kernel(const Params p) {
    pdl_trigger();
    pdl_wait();
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
        launch_chain(dec::attn_decode_kernel<128>, grid, dim3(dec::THREADS), 0, st, p);
        launch_chain(dec::attn_decode_combine_kernel<128>, dim3((unsigned)(batch * heads)), dim3(128), 0, st, p);
    } else {
        launch_chain(dec::attn_decode_kernel<64>, grid, dim3(dec::THREADS), 0, st, p);
        launch_chain(dec::attn_decode_combine_kernel<64>, dim3((unsigned)(batch * heads)), dim3(64), 0, st, p);
    }
    return check_launch("attn_decode");
}

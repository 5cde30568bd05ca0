// Bridge attention backward (A10) on tcgen05/TMEM: two kernels, no atomics, deterministic.
//
//   lb_attn_bwd_dq : CTA = (sample, 128-row q tile, variant, head); loops over 64-wide kv tiles:
//                    S = Q.K^T, dP = dO.V^T (SS MMAs) -> dS = P*(dP-delta)*scale (threads) -> dQ += dS.K (TS MMA, K as MN-major B)
//                    (two CTAs per SM, 256 TMEM columns each)
//   lb_attn_bwd_dkv: attn_bwd_dkv.cu
// P is recomputed from the saved log-sum-exp; delta = rowsum(dO*O) comes from lb_attn_bwd_prepare.
// Gradients w.r.t. the two operand variants (K0/V0 seen by qflag==0 rows, K1/V1 by qflag==1 rows) are written to
// separate buffers; the prologue's adjoint (lb_attn_prep_bwd) folds them into dk, dv and the bridge gradients.
#include <math_constants.h>

#include "common.cuh"

namespace lb {

constexpr int BW_COMPUTE_WARPS = 8;              // two warpgroups, each owns half of the score columns of a tile
constexpr int BW_COMPUTE_THREADS = BW_COMPUTE_WARPS * 32;
constexpr int BW_WARP_TMA = BW_COMPUTE_WARPS, BW_WARP_MMA = BW_COMPUTE_WARPS + 1;
constexpr int BW_THREADS = (BW_COMPUTE_WARPS + 2) * 32;
constexpr float LOG2E_F = 1.4426950408889634f;

struct AttnBwdParams {
    const uint8_t* qflag;      // [B*T] or null
    const uint8_t* qtile_has;  // [B,2,n_qtiles] or null (dK/dV only): q tile holds rows of the variant
    const int32_t* work;       // [n][4]
    const int32_t* kv_start;
    const int32_t* kv_end;
    const float* lse;          // [B,H,T]
    const float* delta;        // [B,H,T]
    __nv_bfloat16* out0;       // dQ            | dK0
    __nv_bfloat16* out1;       //               | dV0
    __nv_bfloat16* out2;       //               | dK1
    __nv_bfloat16* out3;       //               | dV1
    int batch, seqlen, heads;
    int n_work, head_group;
    float scale;
};

// ---- score-tile helpers with the mask as a template parameter (the common unmasked path carries no index arithmetic).
// All TMEM reads of a tile are issued up front and waited for once; results are packed in place.
// dQ kernel: thread = query row; 32 key columns at TMEM `ts` (S) / `tdp` (dP); dS (bf16, 16 columns) is written over S.
template <bool MASK, bool CAUSAL>
__device__ __forceinline__ void dq_tile(uint32_t ts, uint32_t tdp, float sl2, float lse2, float dlt, float scale, int kv0, int qi,
                                        int kvs, int kve) {
    uint32_t s[32], dp[32];
    tmem_ld32(ts, s);
    tmem_ld32(tdp, dp);
    tc_wait_ld();
#pragma unroll
    for (int j = 0; j < 32; j += 2) {
        float p0 = fast_ex2(fmaf(__uint_as_float(s[j]), sl2, -lse2));
        float p1 = fast_ex2(fmaf(__uint_as_float(s[j + 1]), sl2, -lse2));
        if (MASK) {
            const int kj = kv0 + j;
            p0 = ((!CAUSAL || kj <= qi) && kj < kve && kj >= kvs) ? p0 : 0.f;
            p1 = ((!CAUSAL || kj + 1 <= qi) && kj + 1 < kve && kj + 1 >= kvs) ? p1 : 0.f;
        }
        const float d0 = p0 * (__uint_as_float(dp[j]) - dlt) * scale;
        const float d1 = p1 * (__uint_as_float(dp[j + 1]) - dlt) * scale;
        s[j >> 1] = pack_bf16(d0, d1);
    }
    tmem_st16(ts, s);       // dS (bf16) over my own, already consumed S columns
}

// =====================================================================================================
// dQ kernel
// =====================================================================================================
template <int D>
struct DqSmem {
    static constexpr int BN = 64;
    static constexpr int Q_BYTES = 128 * D * 2;
    static constexpr int DO_BYTES = 128 * D * 2;
    static constexpr int K_BYTES = BN * D * 2;
    static constexpr int V_BYTES = BN * D * 2;
    static constexpr int BAR_OFF = Q_BYTES + DO_BYTES + K_BYTES + V_BYTES;
    static constexpr int TOTAL = BAR_OFF + 1024 + 128;
};
enum { DQ_QDO = 0, DQ_KFULL, DQ_KEMPTY, DQ_VFULL, DQ_VEMPTY, DQ_SDP, DQ_DS, DQ_READY, DQ_NBAR };

template <int D, bool CAUSAL>
__global__ void __launch_bounds__(BW_THREADS, 2)
attn_bwd_dq_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                   const __grid_constant__ CUtensorMap tmK0, const __grid_constant__ CUtensorMap tmV0,
                   const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV1,
                   const AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using S = DqSmem<D>;
    constexpr int BN = S::BN;
    uint8_t* sQ = smem;
    uint8_t* sdO = sQ + S::Q_BYTES;
    uint8_t* sK = sdO + S::DO_BYTES;
    uint8_t* sV = sK + S::K_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + DQ_NBAR);
    const uint32_t bar0 = smem_u32(bars);                 // barrier i lives at bar0 + 8 i

    const int warp = threadIdx.x >> 5;
    int item, h;
    attn_cta_order(p.n_work, p.heads, p.head_group, item, h);
    const int b = p.work[item * 4 + 0];
    const int q_tile = p.work[item * 4 + 1];
    const int variant = p.work[item * 4 + 2];
    const int T = p.seqlen;
    const int q0 = q_tile * 128;
    const int kvs = p.kv_start ? p.kv_start[b] : 0;
    const int kve = p.kv_end ? p.kv_end[b] : T;
    const int first_tile = kvs / BN;
    int last_tile = (kve + BN - 1) / BN;
    if (CAUSAL && last_tile > (q0 + 128) / BN) last_tile = (q0 + 128) / BN;
    const int n_tiles = last_tile > first_tile ? last_tile - first_tile : 0;

    constexpr uint32_t TMEM_COLS = 256, COL_S = 0, COL_DP = 64, COL_DQ = 128;

    if (threadIdx.x == 0) {
        for (int i = 0; i < DQ_NBAR; ++i) mbar_init(bars + i, i == DQ_DS ? BW_COMPUTE_THREADS : 1);
        fence_barrier_init();
    }
    if (warp == BW_WARP_MMA) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == BW_WARP_TMA) {
        if (elect_one() && n_tiles > 0) {
            const CUtensorMap* tK = variant ? &tmK1 : &tmK0;
            const CUtensorMap* tV = variant ? &tmV1 : &tmV0;
            const int row_q = b * T + q0;
            mbar_arrive_expect_tx(bars + DQ_QDO, S::Q_BYTES + S::DO_BYTES);
#pragma unroll
            for (int c = 0; c < D / 64; ++c) {
                tma_load_2d(sQ + c * (128 * 128), &tmQ, bars + DQ_QDO, h * D + c * 64, row_q);
                tma_load_2d(sdO + c * (128 * 128), &tmdO, bars + DQ_QDO, h * D + c * 64, row_q);
            }
            for (int it = 0; it < n_tiles; ++it) {
                const int row_k = b * T + (first_tile + it) * BN;
                const uint32_t ph = (uint32_t)it & 1u;
                wait_bar(bar0 + 8 * (DQ_KEMPTY), ph ^ 1u);
                mbar_arrive_expect_tx(bars + DQ_KFULL, S::K_BYTES);
#pragma unroll
                for (int c = 0; c < D / 64; ++c) tma_load_2d(sK + c * (BN * 128), tK, bars + DQ_KFULL, h * D + c * 64, row_k);
                wait_bar(bar0 + 8 * (DQ_VEMPTY), ph ^ 1u);
                mbar_arrive_expect_tx(bars + DQ_VFULL, S::V_BYTES);
#pragma unroll
                for (int c = 0; c < D / 64; ++c) tma_load_2d(sV + c * (BN * 128), tV, bars + DQ_VFULL, h * D + c * 64, row_k);
            }
        }
    } else if (warp == BW_WARP_MMA) {
        if (elect_one() && n_tiles > 0) {
            constexpr uint32_t idesc_s = make_idesc_bf16(128, BN, 0, 0);
            constexpr uint32_t idesc_dq = make_idesc_bf16(128, D, 0, 1);
            const uint32_t dQ0 = desc_lo_kmajor(smem_u32(sQ)), ddO0 = desc_lo_kmajor(smem_u32(sdO));
            const uint32_t dK0 = desc_lo_kmajor(smem_u32(sK)), dV0 = desc_lo_kmajor(smem_u32(sV));
            const uint32_t dKmn0 = desc_lo_mnmajor(smem_u32(sK), BN * 128);
            wait_bar(bar0 + 8 * (DQ_QDO), 0);
            for (int it = 0; it < n_tiles; ++it) {
                const uint32_t ph = (uint32_t)it & 1u;
                wait_bar(bar0 + 8 * (DQ_KFULL), ph);
                tc_fence_after_sync();
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t offA = ((uint32_t)(kk / 4) * (128 * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                    const uint32_t offB = ((uint32_t)(kk / 4) * (BN * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                    umma_ss_lo(tmem_base + COL_S, dQ0 + offA, dK0 + offB, idesc_s, kk ? 1u : 0u);
                }
                wait_bar(bar0 + 8 * (DQ_VFULL), ph);
                tc_fence_after_sync();
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t offA = ((uint32_t)(kk / 4) * (128 * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                    const uint32_t offB = ((uint32_t)(kk / 4) * (BN * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                    umma_ss_lo(tmem_base + COL_DP, ddO0 + offA, dV0 + offB, idesc_s, kk ? 1u : 0u);
                }
                commit_bar(bar0 + 8 * (DQ_VEMPTY));
                commit_bar(bar0 + 8 * (DQ_SDP));
                wait_bar(bar0 + 8 * (DQ_DS), ph);
                tc_fence_after_sync();
#pragma unroll
                for (int kk = 0; kk < BN / 16; ++kk)     // dS of keys 16kk.. lives at column 32*(kk/2) + 8*(kk%2)
                    umma_ts_lo(tmem_base + COL_DQ, tmem_base + COL_S + (uint32_t)(kk / 2) * 32 + (uint32_t)(kk % 2) * 8,
                               dKmn0 + (uint32_t)kk * (2048 >> 4), idesc_dq, (it | kk) ? 1u : 0u);
                commit_bar(bar0 + 8 * (DQ_KEMPTY));
                commit_bar(bar0 + 8 * (DQ_READY));
            }
        }
    } else {
        // ---------------- compute warps: dS = P * (dP - delta) * scale for my 32 of the tile's 64 key columns
        const int half = warp >> 2;
        const int r = (warp & 3) * 32 + (threadIdx.x & 31);
        const int qi = q0 + r;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t colS = COL_S + half * 32, colDP = COL_DP + half * 32;      // my dS (bf16) goes to [colS, colS+16)
        const float sl2 = p.scale * LOG2E_F;
        const int64_t stat_idx = ((int64_t)b * p.heads + h) * T + (qi < T ? qi : 0);
        const bool row_ok = (qi < T) && (!p.qflag || (int)p.qflag[(int64_t)b * T + (qi < T ? qi : 0)] == variant);
        const float lse2 = row_ok ? p.lse[stat_idx] * LOG2E_F : CUDART_INF_F;
        const float dlt = row_ok ? p.delta[stat_idx] : 0.f;
        for (int it = 0; it < n_tiles; ++it) {
            const uint32_t ph = (uint32_t)it & 1u;
            const int kv0 = (first_tile + it) * BN + half * 32;
            const bool need_mask = (CAUSAL && kv0 + 31 > q0) || (kv0 + 32 > kve) || (kv0 < kvs);
            wait_bar(bar0 + 8 * (DQ_SDP), ph);
            tc_fence_after_sync();
            if (need_mask) dq_tile<true, CAUSAL>(lane_addr + colS, lane_addr + colDP, sl2, lse2, dlt, p.scale, kv0, qi, kvs, kve);
            else           dq_tile<false, CAUSAL>(lane_addr + colS, lane_addr + colDP, sl2, lse2, dlt, p.scale, kv0, qi, kvs, kve);
            tc_wait_st();
            tc_fence_before_sync();
            mbar_arrive(bars + DQ_DS);
        }
        if (n_tiles > 0) {
            wait_bar(bar0 + 8 * (DQ_READY), (uint32_t)(n_tiles - 1) & 1u);
            tc_fence_after_sync();
        }
        // dQ leaves through a per-warp staging tile (the Q/dO/K/V tiles are dead now): 32 rows x D/2 columns bf16, 16-byte
        // chunks XOR-swizzled by row, written out as whole row segments (see the dK/dV epilogue in attn_bwd_dkv.cu)
        constexpr int DH = D / 2, CH = DH / 8, RPI = 32 / CH;
        const int lane = threadIdx.x & 31;
        const uint32_t stage_s = smem_u32(smem) + warp * (32 * DH * 2);
        auto swz = [](int row, int chunk) { return CH == 8 ? (chunk ^ (row & 7)) : (chunk ^ ((row >> 1) & 3)); };
#pragma unroll
        for (int c = 0; c < DH / 32; ++c) {
            uint32_t v[32];
            if (n_tiles > 0) {
                tmem_ld32(lane_addr + COL_DQ + half * DH + c * 32, v);
                tc_wait_ld();
            } else {
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = 0u;
            }
#pragma unroll
            for (int j = 0; j < 32; j += 8) {
                uint4 o;
                o.x = pack_bf16(__uint_as_float(v[j + 0]), __uint_as_float(v[j + 1]));
                o.y = pack_bf16(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3]));
                o.z = pack_bf16(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5]));
                o.w = pack_bf16(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7]));
                sts128(stage_s + lane * (DH * 2) + (swz(lane, c * 4 + (j >> 3)) << 4), o);
            }
        }
        __syncwarp();
        const unsigned ok_mask = __ballot_sync(0xffffffffu, row_ok);
        __nv_bfloat16* dst = p.out0 + ((int64_t)b * T + q0 + (warp & 3) * 32) * ((int64_t)p.heads * D) + (int64_t)h * D + half * DH;
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            const int row = i * RPI + lane / CH, chunk = lane % CH;
            const uint4 o = lds128(stage_s + row * (DH * 2) + (swz(row, chunk) << 4));
            if ((ok_mask >> row) & 1u) *reinterpret_cast<uint4*>(dst + (int64_t)row * ((int64_t)p.heads * D) + chunk * 8) = o;
        }
        tc_fence_before_sync();
    }
    __syncthreads();
    if (warp == BW_WARP_MMA) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <typename KernT>
static int configure(KernT kern, int smem, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(LB_ELAUNCH, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return LB_OK;
}

template <int D, bool CAUSAL>
static int launch_dq(const CUtensorMap* tm, const AttnBwdParams& p, int n_work, cudaStream_t st) {
    auto kern = attn_bwd_dq_kernel<D, CAUSAL>;
    static bool configured = false;
    if (!configured) {
        int rc = configure(kern, DqSmem<D>::TOTAL, "attn_bwd_dq");
        if (rc) return rc;
        configured = true;
    }
    kern<<<(unsigned)(n_work * p.heads), BW_THREADS, DqSmem<D>::TOTAL, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4],
                                                                                       tm[5], p);
    return check_launch("attn_bwd_dq");
}

static int make_maps(CUtensorMap* tm, const void* Q, const void* dO, const void* K0, const void* V0, const void* K1,
                     const void* V1, int batch, int seqlen, int heads, int head_dim, uint32_t q_box, uint32_t kv_box) {
    const uint64_t rows = (uint64_t)batch * seqlen, cols = (uint64_t)heads * head_dim;
    const void* ptrs[6] = {Q, dO, K0, V0, K1 ? K1 : K0, V1 ? V1 : V0};
    for (int i = 0; i < 6; ++i) {
        int rc = make_tmap_bf16_2d(&tm[i], ptrs[i], rows, cols, cols, i < 2 ? q_box : kv_box, 64);
        if (rc) return rc;
    }
    return LB_OK;
}

}  // namespace lb

using namespace lb;

extern "C" {

int lb_attn_bwd_dq(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1, const void* dO,
                   const float* lse, const float* delta, const uint8_t* qflag, const int32_t* work, int n_work,
                   const int32_t* kv_start, const int32_t* kv_end, void* dQ, int batch, int seqlen, int heads,
                   int head_dim, int causal, float scale, void* stream) {
    LB_REQUIRE(batch > 0 && seqlen > 0 && heads > 0 && n_work >= 0, LB_EINVAL, "attn_bwd_dq: bad shape");
    LB_REQUIRE(head_dim == 64 || head_dim == 128, LB_EINVAL, "attn_bwd_dq: head_dim %d (64 or 128 supported)", head_dim);
    LB_REQUIRE(Q && K0 && V0 && dO && lse && delta && work && dQ, LB_EINVAL, "attn_bwd_dq: null argument");
    if (n_work == 0) return LB_OK;
    int rc = require_sm100();
    if (rc) return rc;
    CUtensorMap tm[6];
    rc = make_maps(tm, Q, dO, K0, V0, K1, V1, batch, seqlen, heads, head_dim, 128, 64);
    if (rc) return rc;
    AttnBwdParams p{};
    p.qflag = qflag; p.work = work; p.kv_start = kv_start; p.kv_end = kv_end; p.lse = lse; p.delta = delta;
    p.out0 = (__nv_bfloat16*)dQ; p.batch = batch; p.seqlen = seqlen; p.heads = heads; p.scale = scale;
    p.n_work = n_work; p.head_group = attn_head_group();
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 128) return causal ? launch_dq<128, true>(tm, p, n_work, st) : launch_dq<128, false>(tm, p, n_work, st);
    return causal ? launch_dq<64, true>(tm, p, n_work, st) : launch_dq<64, false>(tm, p, n_work, st);
}


}

// Bridge attention backward, dK/dV kernel (A10) on tcgen05/TMEM; the dQ kernel and the shared notes live in attn_bwd.cu.
//
//   lb_attn_bwd_dkv: CTA = (sample, 128-row kv tile, variant, head); loops over the 128-row q tiles that contain rows of the
//                    variant: S^T = K.Q^T, dP^T = V.dO^T -> P^T, dS^T (threads) -> dV += P^T.dO, dK += dS^T.Q (TS MMAs,
//                    dO / Q as MN-major B).  The CTA owns the SM (512 TMEM columns) and is software-pipelined by
//                    64-query halves: while the compute warps of one half turn scores into P^T/dS^T the tensor pipe
//                    runs the other half's MMAs (16 compute warps; measured 381 us vs 479 us unpipelined at
//                    B=4, T=2048, H=32, D=128).
// No atomics, deterministic.  P is recomputed from the saved log-sum-exp; delta = rowsum(dO*O) from lb_attn_bwd_prepare.
#include <math_constants.h>

#include "common.cuh"

namespace lb {
namespace dkv {

constexpr float LOG2E_F = 1.4426950408889634f;
// dK/dV kernel: 16 compute warps (4 warpgroups; two per 64-query half, 32 query columns per thread), 1 CTA per SM
constexpr int KV_COMPUTE_WARPS = 16;
constexpr int KV_WARP_TMA = KV_COMPUTE_WARPS, KV_WARP_MMA = KV_COMPUTE_WARPS + 1;
constexpr int KV_THREADS = (KV_COMPUTE_WARPS + 2) * 32;

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
    long long* cta_log;        // optional [grid][8]: q tiles, clock64 at entry / tile list ready / K,V landed / last MMA issued / all MMAs done / exit
};

// Score-tile helper with the mask as a template parameter (the common unmasked path carries no index arithmetic).
// All TMEM reads of a tile are issued up front and waited for once; results are packed in place.
// dK/dV kernel: thread = key row; 32 query columns at TMEM `ts` (S^T) / `tdp` (dP^T); per-column lse2 / delta in smem.
template <bool MASK, bool CAUSAL>
__device__ __forceinline__ void dkv_tile(uint32_t ts, uint32_t tdp, uint32_t st_s, float sl2, float scale, int kj,
                                         int qbase, bool key_ok) {
    // st_s: shared-window address of this thread's 32 values of -lse2 (-inf on excluded query rows); the 32 values of
    // -delta*scale sit 512 bytes further (LDS.128 instead of one generic load per column).  Packed fp32x2 arithmetic: per PAIR
    // of query columns FFMA2 (exponent), 2 x MUFU.EX2, FFMA2 (dP*scale - delta*scale), FMUL2, two bf16x2 packs.
    uint32_t s[32], dp[32];
    tmem_ld32(ts, s);
    tmem_ld32(tdp, dp);
    tc_wait_ld();
    const uint64_t sl2_2 = f32x2(sl2, sl2), sc_2 = f32x2(scale, scale);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        float4 ls, dl;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ls.x), "=f"(ls.y), "=f"(ls.z), "=f"(ls.w) : "r"(st_s + j * 4));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dl.x), "=f"(dl.y), "=f"(dl.z), "=f"(dl.w) : "r"(st_s + 512 + j * 4));
        const uint64_t nlse[2] = {f32x2(ls.x, ls.y), f32x2(ls.z, ls.w)}, nds[2] = {f32x2(dl.x, dl.y), f32x2(dl.z, dl.w)};
#pragma unroll
        for (int u = 0; u < 4; u += 2) {
            float x0, x1;
            f32x2_unpack(fma_f32x2(f32x2(__uint_as_float(s[j + u]), __uint_as_float(s[j + u + 1])), sl2_2, nlse[u >> 1]), x0, x1);
            float p0 = fast_ex2(x0), p1 = fast_ex2(x1);
            if (MASK) {
                const int qa = qbase + j + u;
                p0 = (key_ok && (!CAUSAL || kj <= qa)) ? p0 : 0.f;
                p1 = (key_ok && (!CAUSAL || kj <= qa + 1)) ? p1 : 0.f;
            }
            const uint64_t t = fma_f32x2(f32x2(__uint_as_float(dp[j + u]), __uint_as_float(dp[j + u + 1])), sc_2, nds[u >> 1]);
            float d0, d1;
            f32x2_unpack(mul_f32x2(f32x2(p0, p1), t), d0, d1);
            s[(j + u) >> 1] = pack_bf16(p0, p1);
            dp[(j + u) >> 1] = pack_bf16(d0, d1);
        }
    }
    tmem_st16(ts, s);        // P^T (bf16, 16 columns) over my own, already consumed S^T columns
    tmem_st16(tdp, dp);      // dS^T over dP^T
}

// =====================================================================================================
// dK/dV kernel
// =====================================================================================================
template <int D>
struct DkvSmem {
    static constexpr int K_BYTES = 128 * D * 2;
    static constexpr int V_BYTES = 128 * D * 2;
    static constexpr int Q_BYTES = 128 * D * 2;      // per stage
    static constexpr int DO_BYTES = 128 * D * 2;     // per stage
    static constexpr int STAGES = 2;
    static constexpr int STAT_OFF = K_BYTES + V_BYTES + STAGES * (Q_BYTES + DO_BYTES);
    static constexpr int STAT_BYTES = 2 * 128 * 12;  // [stage][lse2, delta, ok] x 128
    static constexpr int BAR_OFF = STAT_OFF + STAT_BYTES;
    static constexpr int TOTAL = BAR_OFF + 1024 + 256;
};
// SDP_h / PDS_h: scores ready / probabilities written, per 64-query-column half h (each half is an independent pipeline)
enum { KV_KV = 0, KV_QFULL0, KV_QFULL1, KV_QEMPTY0, KV_QEMPTY1, KV_SDP0, KV_SDP1, KV_PDS0, KV_PDS1, KV_DONE, KV_NBAR };

template <int D, bool CAUSAL>
__global__ void __launch_bounds__(KV_THREADS, 1)
attn_bwd_dkv_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                    const __grid_constant__ CUtensorMap tmK0, const __grid_constant__ CUtensorMap tmV0,
                    const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV1,
                    const AttnBwdParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using S = DkvSmem<D>;
    uint8_t* sK = smem;
    uint8_t* sV = sK + S::K_BYTES;
    uint8_t* sQ0 = sV + S::V_BYTES;                       // stage s: sQ0 + s*(Q+dO), dO right after Q
    float* stats = reinterpret_cast<float*>(smem + S::STAT_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + KV_NBAR);
    const uint32_t bar0 = smem_u32(bars);                 // barrier i lives at bar0 + 8 i

    const int warp = threadIdx.x >> 5;
    long long* clog = p.cta_log ? p.cta_log + (int64_t)blockIdx.x * 8 : nullptr;
    if (clog && threadIdx.x == 0) clog[1] = clock64();
    int item, h;
    attn_cta_order(p.n_work, p.heads, p.head_group, item, h);
    const int b = p.work[item * 4 + 0];
    const int kv_tile = p.work[item * 4 + 1];
    const int variant = p.work[item * 4 + 2];
    const int first_q = p.work[item * 4 + 3];
    const int T = p.seqlen;
    const int kv0 = kv_tile * 128;
    const int kvs = p.kv_start ? p.kv_start[b] : 0;
    const int kve = p.kv_end ? p.kv_end[b] : T;
    const int nqt = (T + 127) / 128;

    constexpr uint32_t TMEM_COLS = 512, COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 256 + D;

    // q tiles that hold rows of this variant (host-computed bitmap; without it every tile is visited and the
    // per-row flags alone do the masking)
    auto tile_has = [&](int qt) -> bool {
        return p.qtile_has ? p.qtile_has[((int64_t)b * 2 + variant) * nqt + qt] != 0 : true;
    };

    if (threadIdx.x == 0) {
        for (int i = 0; i < KV_NBAR; ++i)
            if (i != KV_KV) mbar_init(bars + i, (i == KV_PDS0 || i == KV_PDS1) ? 256 : 1);
        fence_barrier_init();
    }
    if (warp == KV_WARP_TMA && elect_one()) {
        // K and V of this CTA's kv tile go out before anything else (their barrier is initialised right here): the load
        // latency (~2-3 k clk) then overlaps the TMEM allocation, the tile list and the CTA-wide barrier below
        mbar_init(bars + KV_KV, 1);
        fence_barrier_init();
        const CUtensorMap* tK = variant ? &tmK1 : &tmK0;
        const CUtensorMap* tV = variant ? &tmV1 : &tmV0;
        const int row_k = b * T + kv0;
        mbar_arrive_expect_tx(bars + KV_KV, S::K_BYTES + S::V_BYTES);
#pragma unroll
        for (int c = 0; c < D / 64; ++c) {
            tma_load_2d(sK + c * (128 * 128), tK, bars + KV_KV, h * D + c * 64, row_k);
            tma_load_2d(sV + c * (128 * 128), tV, bars + KV_KV, h * D + c * 64, row_k);
        }
    }
    if (warp == KV_WARP_MMA) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    // tile list in shared memory (<= 64 q tiles, T <= 8192): warps 0 and 1 read one flag per lane and compact them with
    // ballots (one thread walking the bitmap cost a dependent global load per q tile: ~10 k clk for the first kv tiles)
    __shared__ int s_tiles[64];
    __shared__ int s_ntiles;
    __shared__ int s_cnt0;
    if (warp < 2) {
        const int qt = first_q + (int)threadIdx.x;
        const bool has = qt < nqt && tile_has(qt);
        const unsigned m = __ballot_sync(0xffffffffu, has);
        if (threadIdx.x == 0) s_cnt0 = __popc(m);
        asm volatile("bar.sync 15, 64;" ::: "memory");
        const int base = warp ? s_cnt0 : 0;
        if (has) s_tiles[base + __popc(m & ((1u << (threadIdx.x & 31)) - 1u))] = qt;
        if (threadIdx.x == 32) s_ntiles = s_cnt0 + __popc(m);
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const int n_tiles = s_ntiles;
    if (clog && threadIdx.x == 0) {
        clog[0] = n_tiles;
        clog[2] = clock64();
    }

    if (warp == KV_WARP_TMA) {
        if (elect_one() && n_tiles > 0) {
            for (int it = 0; it < n_tiles; ++it) {
                const int st = it & 1;
                const uint32_t ph = (uint32_t)(it >> 1) & 1u;
                const int row_q = b * T + s_tiles[it] * 128;
                uint8_t* q = sQ0 + st * (S::Q_BYTES + S::DO_BYTES);
                uint8_t* d = q + S::Q_BYTES;
                wait_bar(bar0 + 8 * (KV_QEMPTY0 + st), ph ^ 1u);
                mbar_arrive_expect_tx(bars + KV_QFULL0 + st, S::Q_BYTES + S::DO_BYTES);
#pragma unroll
                for (int c = 0; c < D / 64; ++c) {
                    tma_load_2d(q + c * (128 * 128), &tmQ, bars + KV_QFULL0 + st, h * D + c * 64, row_q);
                    tma_load_2d(d + c * (128 * 128), &tmdO, bars + KV_QFULL0 + st, h * D + c * 64, row_q);
                }
            }
        }
    } else if (warp == KV_WARP_MMA) {
        if (n_tiles == 0 && elect_one()) wait_bar(bar0 + 8 * (KV_KV), 0);        // never exit with the K/V load in flight
        if (elect_one() && n_tiles > 0) {
            constexpr uint32_t idesc_g = make_idesc_bf16(128, D, 0, 1);
            const uint32_t dK0 = desc_lo_kmajor(smem_u32(sK)), dV0 = desc_lo_kmajor(smem_u32(sV));
            constexpr uint32_t idesc_h = make_idesc_bf16(128, 64, 0, 0);      // S^T / dP^T for one 64-query half
            // Software pipeline over (q tile, half): the tensor pipe computes the scores of the next half while the compute
            // warpgroup of the previous half turns its scores into P^T / dS^T, and the dV/dK MMAs of a half are issued as
            // soon as that half's probabilities are back in TMEM.  Issue order per tile `it`:
            //   S_A dP_A | [dV_B dK_B of tile it-1] | S_B dP_B | dV_A dK_A
            auto issue_scores = [&](int half, uint32_t dQk, uint32_t ddOk) {
                const uint32_t hoff = (uint32_t)(half * 64 * 128) >> 4;           // 64 query rows further down each 16 KB chunk
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = ((uint32_t)(kk / 4) * (128 * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                    umma_ss_lo(tmem_base + COL_S + half * 64, dK0 + off, dQk + off + hoff, idesc_h, kk ? 1u : 0u);
                }
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = ((uint32_t)(kk / 4) * (128 * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                    umma_ss_lo(tmem_base + COL_DP + half * 64, dV0 + off, ddOk + off + hoff, idesc_h, kk ? 1u : 0u);
                }
                commit_bar(bar0 + 8 * (KV_SDP0 + half));
            };
            auto issue_grads = [&](int half, uint32_t dQmn, uint32_t ddOmn, bool first) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)      // 64 queries of this half = 4 K steps; P^T of quarter kk/2 at 64*half + 32*(kk/2) + 8*(kk%2)
                    umma_ts_lo(tmem_base + COL_DV, tmem_base + COL_S + (uint32_t)half * 64 + (uint32_t)(kk / 2) * 32 + (uint32_t)(kk % 2) * 8,
                               ddOmn + (uint32_t)(half * 4 + kk) * (2048 >> 4), idesc_g, (first && kk == 0) ? 0u : 1u);
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                    umma_ts_lo(tmem_base + COL_DK, tmem_base + COL_DP + (uint32_t)half * 64 + (uint32_t)(kk / 2) * 32 + (uint32_t)(kk % 2) * 8,
                               dQmn + (uint32_t)(half * 4 + kk) * (2048 >> 4), idesc_g, (first && kk == 0) ? 0u : 1u);
            };
            wait_bar(bar0 + 8 * (KV_KV), 0);
            if (clog) clog[3] = clock64();
            uint32_t pQmn = 0, pdOmn = 0;            // MN-major descriptors of the previous tile's stage
            for (int it = 0; it < n_tiles; ++it) {
                const int st = it & 1;
                const uint32_t phq = (uint32_t)(it >> 1) & 1u;
                const uint32_t ph = (uint32_t)it & 1u;
                const uint32_t aQ = smem_u32(sQ0 + st * (S::Q_BYTES + S::DO_BYTES));
                const uint32_t adO = aQ + S::Q_BYTES;
                const uint32_t dQk = desc_lo_kmajor(aQ), ddOk = desc_lo_kmajor(adO);
                const uint32_t dQmn = desc_lo_mnmajor(aQ, 128 * 128), ddOmn = desc_lo_mnmajor(adO, 128 * 128);
                wait_bar(bar0 + 8 * (KV_QFULL0 + st), phq);
                tc_fence_after_sync();
                issue_scores(0, dQk, ddOk);
                if (it > 0) {
                    wait_bar(bar0 + 8 * (KV_PDS1), ph ^ 1u);                  // half B of the previous tile
                    tc_fence_after_sync();
                    issue_grads(1, pQmn, pdOmn, false);
                    commit_bar(bar0 + 8 * (KV_QEMPTY0 + (st ^ 1)));             // previous tile's Q/dO stage is free
                }
                issue_scores(1, dQk, ddOk);
                wait_bar(bar0 + 8 * (KV_PDS0), ph);
                tc_fence_after_sync();
                issue_grads(0, dQmn, ddOmn, it == 0);
                pQmn = dQmn;
                pdOmn = ddOmn;
            }
            {
                const int last = n_tiles - 1;
                wait_bar(bar0 + 8 * (KV_PDS1), (uint32_t)last & 1u);
                tc_fence_after_sync();
                issue_grads(1, pQmn, pdOmn, false);
                commit_bar(bar0 + 8 * (KV_QEMPTY0 + (last & 1)));
            }
            commit_bar(bar0 + 8 * (KV_DONE));
            if (clog) clog[4] = clock64();
        }
    } else {
        // ---------------- compute warps: thread <-> kv row (TMEM lane).  Warpgroup g = warp/4 owns query columns [32g, 32g+32);
        // the two warpgroups of a 64-column half share that half's pipeline (statistics buffer, named barrier, SDP/PDS).
        const int wg = warp >> 2, half = wg >> 1, quarter = wg & 1;
        const int r = (warp & 3) * 32 + (threadIdx.x & 31);
        const int kj = kv0 + r;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t colS = COL_S + wg * 32, colDP = COL_DP + wg * 32;
        const float sl2 = p.scale * LOG2E_F;
        const bool key_ok = kj < kve && kj >= kvs;
        const int t256 = threadIdx.x & 255;      // thread within the half's 256-thread group
        for (int it = 0; it < n_tiles; ++it) {
            const uint32_t ph = (uint32_t)it & 1u;
            const int q0 = s_tiles[it] * 128;
            const int qhalf = q0 + half * 64;
            // per-column statistics of this half's 64 query columns: [parity][half][-lse2 x 64 | (-delta*scale at +128) x 64]
            float* sl = stats + (it & 1) * 384 + half * 64;
            if (t256 < 128) {
                const int col = t256 & 63;
                const int qi = qhalf + col;
                const bool ok = qi < T && (!p.qflag || (int)p.qflag[(int64_t)b * T + (qi < T ? qi : 0)] == variant);
                const int64_t si = ((int64_t)b * p.heads + h) * T + (qi < T ? qi : 0);
                if (t256 < 64) sl[col] = ok ? -p.lse[si] * LOG2E_F : -CUDART_INF_F;       // -lse2 (excluded rows: P = 2^-inf = 0)
                else           sl[128 + col] = ok ? -p.delta[si] * p.scale : 0.f;         // -delta * scale
            }
            named_bar_sync(1 + half, 256);
            wait_bar(bar0 + 8 * (KV_SDP0 + half), ph);
            tc_fence_after_sync();
            const int qbase = qhalf + quarter * 32;
            // warp-uniform: tile straddles the key range or the causal diagonal (excluded query rows carry lse = +inf)
            const bool need_mask = (kv0 + 128 > kve) || (kv0 < kvs) || (CAUSAL && kv0 + 127 > qbase);
            const uint32_t st_s = smem_u32(sl + quarter * 32);
            if (need_mask) dkv_tile<true, CAUSAL>(lane_addr + colS, lane_addr + colDP, st_s, sl2, p.scale, kj, qbase, key_ok);
            else           dkv_tile<false, CAUSAL>(lane_addr + colS, lane_addr + colDP, st_s, sl2, p.scale, kj, qbase, key_ok);
            tc_wait_st();
            tc_fence_before_sync();
            mbar_arrive(bars + KV_PDS0 + half);
        }
        if (n_tiles > 0) {
            wait_bar(bar0 + 8 * (KV_DONE), 0);
            tc_fence_after_sync();
            if (clog && threadIdx.x == 0) clog[5] = clock64();
            // warpgroups 0,1 store the two halves of dV (TMEM columns [COL_DV, +D)), warpgroups 2,3 those of dK.  The K/V/Q/dO
            // tiles are dead now: each warp stages its 32 rows x D/2 columns (bf16) in shared memory, 16-byte chunks XOR-
            // swizzled by row, and writes whole 128-byte (D=128) row segments -- one row per lane costs 32 LSU wavefronts
            // per store instruction and made this epilogue ~6 k clk of a ~37 k clk CTA.
            constexpr int DH = D / 2, CH = DH / 8, RPI = 32 / CH;      // 16-byte chunks per row, rows per store instruction
            const int lane = threadIdx.x & 31;
            const uint32_t stage_s = smem_u32(smem) + warp * (32 * DH * 2);
            __nv_bfloat16* dst = (wg < 2 ? (variant ? p.out3 : p.out1) : (variant ? p.out2 : p.out0)) + (int64_t)h * D + (wg & 1) * DH;
            const uint32_t col0 = (wg < 2 ? COL_DV : COL_DK) + (wg & 1) * DH;
            auto swz = [](int row, int chunk) { return CH == 8 ? (chunk ^ (row & 7)) : (chunk ^ ((row >> 1) & 3)); };
#pragma unroll
            for (int c = 0; c < DH / 32; ++c) {
                uint32_t v[32];
                tmem_ld32(lane_addr + col0 + c * 32, v);
                tc_wait_ld();
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
            const int row0 = kv0 + (warp & 3) * 32;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                const int row = i * RPI + lane / CH, chunk = lane % CH;
                const uint4 o = lds128(stage_s + row * (DH * 2) + (swz(row, chunk) << 4));
                if (row0 + row < T)
                    *reinterpret_cast<uint4*>(dst + ((int64_t)b * T + row0 + row) * ((int64_t)p.heads * D) + chunk * 8) = o;
            }
        }
        tc_fence_before_sync();
    }
    __syncthreads();
    if (warp == KV_WARP_MMA) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    if (clog && threadIdx.x == 0) clog[6] = clock64();
}

template <typename KernT>
static int configure(KernT kern, int smem, const char* what) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return fail(LB_ELAUNCH, "%s: cudaFuncSetAttribute: %s", what, cudaGetErrorString(e));
    return LB_OK;
}

template <int D, bool CAUSAL>
static int launch_dkv(const CUtensorMap* tm, const AttnBwdParams& p, int n_work, cudaStream_t st) {
    auto kern = attn_bwd_dkv_kernel<D, CAUSAL>;
    static bool configured = false;
    if (!configured) {
        int rc = configure(kern, DkvSmem<D>::TOTAL, "attn_bwd_dkv");
        if (rc) return rc;
        configured = true;
    }
    kern<<<(unsigned)(n_work * p.heads), KV_THREADS, DkvSmem<D>::TOTAL, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4],
                                                                                        tm[5], p);
    return check_launch("attn_bwd_dkv");
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

}  // namespace dkv
}  // namespace lb

using namespace lb;
using namespace lb::dkv;

static long long* g_dkv_cta_log = nullptr;

extern "C" {

/* diagnostics: every CTA of subsequent lb_attn_bwd_dkv launches logs {q tiles, clock64 at entry, tile list ready, K/V
 * landed, last MMA issued, all MMAs done, exit, -} into buf ([n_work*heads][8] int64, device memory).  NULL = off */
int lb_attn_bwd_dkv_set_cta_log(void* buf) {
    g_dkv_cta_log = (long long*)buf;
    return LB_OK;
}

int lb_attn_bwd_dkv(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1, const void* dO,
                    const float* lse, const float* delta, const uint8_t* qflag, const uint8_t* qtile_has,
                    const int32_t* work_kv, int n_work, const int32_t* kv_start, const int32_t* kv_end, void* dK0,
                    void* dV0, void* dK1, void* dV1, int batch, int seqlen, int heads, int head_dim, int causal,
                    float scale, void* stream) {
    LB_REQUIRE(batch > 0 && seqlen > 0 && heads > 0 && n_work >= 0, LB_EINVAL, "attn_bwd_dkv: bad shape");
    LB_REQUIRE(head_dim == 64 || head_dim == 128, LB_EINVAL, "attn_bwd_dkv: head_dim %d (64 or 128 supported)", head_dim);
    LB_REQUIRE(seqlen <= 8192, LB_EINVAL, "attn_bwd_dkv: seqlen %d > 8192", seqlen);
    LB_REQUIRE(Q && K0 && V0 && dO && lse && delta && work_kv && dK0 && dV0, LB_EINVAL, "attn_bwd_dkv: null argument");
    if (n_work == 0) return LB_OK;
    int rc = require_sm100();
    if (rc) return rc;
    CUtensorMap tm[6];
    rc = make_maps(tm, Q, dO, K0, V0, K1, V1, batch, seqlen, heads, head_dim, 128, 128);
    if (rc) return rc;
    AttnBwdParams p{};
    p.qflag = qflag; p.qtile_has = qtile_has; p.work = work_kv; p.kv_start = kv_start; p.kv_end = kv_end; p.lse = lse; p.delta = delta;
    p.out0 = (__nv_bfloat16*)dK0; p.out1 = (__nv_bfloat16*)dV0;
    p.out2 = (__nv_bfloat16*)(dK1 ? dK1 : dK0); p.out3 = (__nv_bfloat16*)(dV1 ? dV1 : dV0);
    p.batch = batch; p.seqlen = seqlen; p.heads = heads; p.scale = scale;
    p.n_work = n_work; p.head_group = attn_head_group();
    p.cta_log = g_dkv_cta_log;
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 128) return causal ? launch_dkv<128, true>(tm, p, n_work, st) : launch_dkv<128, false>(tm, p, n_work, st);
    return causal ? launch_dkv<64, true>(tm, p, n_work, st) : launch_dkv<64, false>(tm, p, n_work, st);
}

}

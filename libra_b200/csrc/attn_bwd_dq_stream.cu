// Bridge attention backward, dQ: persistent streaming kernel (A10; same maths and operands as attn_bwd.cu's dQ kernel).
//
// What the first dQ kernel (attn_bwd.cu: one CTA per (item, head), two CTAs per SM, 64-key tiles) measured: 36 % tensor pipe.
// A CTA went through the whole chain K/V load -> S, dP -> dS (threads) -> dQ -> K slot free -> next K load for every tile
// with K and V single-buffered (two CTAs x 96 KB left no room), so each tile exposed a TMA round trip.  A first persistent
// version with 64-key tiles (S/dP double-buffered, K ring of 4) reached 243 us from 318: its clock64 trace showed the
// ISSUING THREAD as the limit -- a tcgen05.mma costs ~40 clk to issue next to busy compute warps, an M128 N64 K16 MMA runs
// 32 clk, so the 16 S/dP MMAs of a tile took 660 clk to issue for 512 clk of tensor work, inside the per-buffer dependency
// loop S/dP -> dS -> dQ -> buffer free.  Hence 128-key tiles (N = 128: 64 clk per MMA, 24 MMAs per 128 keys instead of 40):
//   * one CTA per SM walks a host-balanced share of the (work item, head) list as ONE stream of 128x128 tiles (index G),
//     like the persistent forward (attn_fwd_stream.cu); the forward's plan is the balanced split for this kernel too;
//   * TMEM (512 columns): S0 [0,128) S1 [128,256) dP [256,384) dQ [384,384+D).  S is double-buffered: S(G+1) runs during
//     the dS computation of tile G.  dP is single-buffered but free again as soon as the threads hold it in registers (the
//     first thing they do), so dP(G+1) follows S(G+1) directly.  dS (bf16) is written over the S columns of its tile and is
//     the TMEM A operand of dQ += dS.K; S(G+2) waits for that MMA to retire.  Tensor pipe per tile: 1536 clk (3 x 512),
//     thread work ~1000 clk, so in steady state the tensor pipe is the limit;
//   * both compute warpgroups work on the SAME tile, thread = query row x 64 of the 128 keys (64 S + 64 dP values in
//     registers; alternate tiles would need 256), each warpgroup writes its dS into its own S columns;
//   * K ring of 3 whole tiles (a K tile lives until its dQ MMA retired), V ring of 3 HALF tiles (64 keys; dP is issued as
//     two N = 64 halves so that a V slot is 16 KB: 64 + 96 + 48 KB of operands + staging is what 227 KB allows);
//   * two issuing threads (S/dP and dQ) so that neither's waits sit between the other's MMAs;
//   * epilogue: both warpgroups pull their 64 dQ columns out of TMEM (released to the next item right after), then pack
//     and store through a per-warp staging tile; S(0)/dP(0) of the next item run meanwhile.
// CTA = 384 threads: warps 0-3 / 4-7 compute warpgroups (keys 0-63 / 64-127 of a tile), warp 8 TMA K and V, warp 9 tcgen05
// issuer for S and dP (+ TMEM alloc), warp 10 tcgen05 issuer for dQ, warp 11 TMA Q and dO.
// Deterministic: no atomics, every dQ element is written by exactly one thread, accumulation order fixed by the tile order.
#include <math_constants.h>

#include "common.cuh"

namespace lb {
namespace dqs {

constexpr int BM = 128, BN = 128, BH = 64;                       // q rows, keys per tile, keys per V half-tile
constexpr int KST = 3, VST = 3;                                  // K ring (whole tiles), V ring (half tiles)
constexpr int WARP_KV = 8, WARP_SDP = 9, WARP_DQ = 10, WARP_QDO = 11, THREADS = 384;
constexpr int REGS_COMPUTE = 216, REGS_PRODUCER = 64;
constexpr int MAX_ITEMS = 128;
constexpr float LOG2E = 1.4426950408889634f;
constexpr uint32_t COL_S = 0, COL_DP = 256, COL_DQ = 384, TMEM_COLS = 512;

struct Params {
    const uint8_t* qflag;        // [B*T] or null
    const int32_t* work;         // [n_work][4] = {b, q_tile, variant, -}
    const int32_t* kv_start;     // [B] or null
    const int32_t* kv_end;       // [B] or null
    const float* lse;            // [B,H,T]
    const float* delta;          // [B,H,T]
    __nv_bfloat16* dQ;           // [B*T, H*D]
    int batch, seqlen, heads;
    int n_work, head_group, n_items, n_cta;
    const int32_t* plan_items;   // [n_items] list positions grouped by CTA (host-side balanced split), or null: snake split
    const int32_t* plan_off;     // [gridDim.x + 1]
    float scale;
    long long* trace;            // optional [64][32] clock64 stamps of CTA 0, one row per tile (global index) / item
};

#define DQ_TRACE(slot, G)                                                                              \
    do {                                                                                               \
        if (TRACE && blockIdx.x == 0 && (G) < 64) p.trace[(G) * 32 + (slot)] = clock64();              \
    } while (0)

struct __align__(16) Item {
    int b, q_tile, variant, h, kvs, kve, first_tile, n_tiles;       // b < 0 (and n_tiles < 0): end of the CTA's list
};

template <int D>
struct Smem {
    static constexpr int QTILE = BM * D * 2;                          // Q / dO / K tile
    static constexpr int VHALF = BH * D * 2;                          // V half tile (64 keys)
    static constexpr int Q_OFF = 0, DO_OFF = QTILE, K_OFF = 2 * QTILE, V_OFF = K_OFF + KST * QTILE;
    static constexpr int ITEM_OFF = V_OFF + VST * VHALF;              // Item[MAX_ITEMS + 1]
    static constexpr int STAGE_OFF = ITEM_OFF + (MAX_ITEMS + 1) * 32 + 96;     // epilogue staging: 8 warps x 8 rows x 128 B
    static constexpr int BAR_OFF = STAGE_OFF + 8 * 8 * 128;
    static constexpr int NEEDED = BAR_OFF + 256 + 1024;
    static_assert(STAGE_OFF % 128 == 0, "staging alignment");
    static_assert(NEEDED <= 227 * 1024, "shared memory budget");
    static constexpr int TOTAL = NEEDED > 120 * 1024 ? NEEDED : 120 * 1024;     // > half an SM: one CTA per SM (512 TMEM columns)
};

enum {
    B_QDOFULL = 0,                   // Q and dO of an item landed
    B_QDOEMPTY,                      // last S / dP MMA of the item retired
    B_KFULL,                         // [3] K(G) landed in slot G % 3
    B_KEMPTY = B_KFULL + KST,        // [3] dQ MMA of tile G retired: K slot free
    B_VFULL = B_KEMPTY + KST,        // [3] V half landed
    B_VEMPTY = B_VFULL + VST,        // [3] dP MMAs of that half retired: V slot free
    B_SFULL = B_VEMPTY + VST,        // [2] S of tile G landed in S[G & 1]
    B_SFREE = B_SFULL + 2,           // [2] dQ MMA of tile G retired: S[G & 1] (S, dS) reusable
    B_DPFULL = B_SFREE + 2,          // dP of tile G landed
    B_DPFREE,                        // dP of tile G is in registers (256 arrivals)
    B_DSFULL,                        // [2] dS of tile G written over S[G & 1] (256 arrivals)
    B_DQFULL = B_DSFULL + 2,         // the item's last dQ MMA retired
    B_DQFREE,                        // dQ of the item is in registers (256 arrivals)
    B_COUNT
};
// Parity waits follow the rule of attn_fwd_stream.cu: a waiter never waits on a barrier that can be more than one phase
// ahead of it (every barrier's next completion depends, through the pipeline, on the waiter's own progress).

__device__ __forceinline__ int item_of_round(const Params& p, int k) {
    if (p.plan_items) {
        const int i = p.plan_off[blockIdx.x] + k;
        return i < p.plan_off[blockIdx.x + 1] ? p.plan_items[i] : -1;
    }
    const int n_items = p.n_items;
    const int G = (int)gridDim.x, c = (int)blockIdx.x;
    if ((int64_t)k * G >= n_items) return -1;
    const int L = k * G + ((k & 1) ? G - 1 - c : c);
    return L < n_items ? L : -1;
}

template <bool CAUSAL>
__device__ __forceinline__ Item decode_item(const Params& p, int L) {
    Item it;
    if (L < 0) {
        it.b = it.n_tiles = -1;
        it.q_tile = it.variant = it.h = it.kvs = it.kve = it.first_tile = 0;
        return it;
    }
    const int per_group = p.head_group * p.n_work;
    const int g = L / per_group;
    const int rem = L - g * per_group;
    const int gl = min(p.head_group, p.heads - g * p.head_group);
    const int w = rem / gl;
    it.h = g * p.head_group + (rem - w * gl);
    it.b = p.work[w * 4 + 0];
    it.q_tile = p.work[w * 4 + 1];
    it.variant = p.work[w * 4 + 2];
    it.kvs = p.kv_start ? p.kv_start[it.b] : 0;
    it.kve = p.kv_end ? p.kv_end[it.b] : p.seqlen;
    it.first_tile = it.kvs / BN;
    int last = (it.kve + BN - 1) / BN;                               // exclusive, in 64-key tiles
    if (CAUSAL && last > (it.q_tile + 1) * (BM / BN)) last = (it.q_tile + 1) * (BM / BN);
    it.n_tiles = last > it.first_tile ? last - it.first_tile : 0;
    return it;
}

__device__ __forceinline__ Item load_item(const Item* tab, int k) {
    const int4 a = reinterpret_cast<const int4*>(tab + k)[0], b = reinterpret_cast<const int4*>(tab + k)[1];
    Item it;
    it.b = a.x; it.q_tile = a.y; it.variant = a.z; it.h = a.w;
    it.kvs = b.x; it.kve = b.y; it.first_tile = b.z; it.n_tiles = b.w;
    return it;
}

// dS = P * (dP - delta) * scale with P = 2^(S*sl2 - lse2); s[64] in, the 32 packed bf16 pairs come back in s[0..32).
// Per PAIR of elements: FFMA2 (exponent), 2 x MUFU.EX2, FFMA2 (dP*scale - delta*scale), FMUL2, one bf16x2 pack -- the SFU
// (16 ex2 / clk / SM = 1024 clk per 128x128 tile) and the issue slots are what bounds the compute warps.
template <bool MASK, bool CAUSAL>
__device__ __forceinline__ void ds_tile(uint32_t (&s)[64], const uint32_t (&dp)[64], float sl2, float lse2, float dlt, float scale,
                                        int kv0, int qi, int kvs, int kve) {
    const uint64_t sl2_2 = f32x2(sl2, sl2), nlse_2 = f32x2(-lse2, -lse2), sc_2 = f32x2(scale, scale);
    const float nds = -dlt * scale;
    const uint64_t nds_2 = f32x2(nds, nds);
#pragma unroll
    for (int j = 0; j < 64; j += 2) {
        float x0, x1;
        f32x2_unpack(fma_f32x2(f32x2(__uint_as_float(s[j]), __uint_as_float(s[j + 1])), sl2_2, nlse_2), x0, x1);
        float p0 = fast_ex2(x0), p1 = fast_ex2(x1);
        if (MASK) {
            const int kj = kv0 + j;
            p0 = ((!CAUSAL || kj <= qi) && kj < kve && kj >= kvs) ? p0 : 0.f;
            p1 = ((!CAUSAL || kj + 1 <= qi) && kj + 1 < kve && kj + 1 >= kvs) ? p1 : 0.f;
        }
        const uint64_t t = fma_f32x2(f32x2(__uint_as_float(dp[j]), __uint_as_float(dp[j + 1])), sc_2, nds_2);
        float d0, d1;
        f32x2_unpack(mul_f32x2(f32x2(p0, p1), t), d0, d1);
        s[j >> 1] = pack_bf16(d0, d1);
    }
}

template <int D, bool CAUSAL, bool TRACE>
__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_dq_stream_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                          const __grid_constant__ CUtensorMap tmK0, const __grid_constant__ CUtensorMap tmV0,
                          const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV1, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using S = Smem<D>;
    Item* items = reinterpret_cast<Item*>(smem + S::ITEM_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    const int warp = threadIdx.x >> 5;
    const int T = p.seqlen;

    if (threadIdx.x == 0) {
        for (int i = 0; i < B_COUNT; ++i) {
            const bool all = i == B_DPFREE || i == B_DSFULL || i == B_DSFULL + 1 || i == B_DQFREE;
            mbar_init(bars + i, all ? 256 : 1);
        }
        fence_barrier_init();
    }
    for (int k = threadIdx.x; k <= MAX_ITEMS; k += THREADS) items[k] = decode_item<CAUSAL>(p, item_of_round(p, k));
    if (warp == WARP_KV && elect_one()) {
        tma_prefetch_desc(&tmK0);
        tma_prefetch_desc(&tmK1);
        tma_prefetch_desc(&tmV0);
        tma_prefetch_desc(&tmV1);
    }
    if (warp == WARP_QDO && elect_one()) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmdO);
    }
    if (warp == WARP_SDP) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t bar0 = smem_u32(bars);                             // barrier i lives at bar0 + 8 i

    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
        if (warp == WARP_QDO) {
            // ------------------------------------------------------------ TMA producer for Q and dO (one pair per item)
            if (elect_one()) {
                uint32_t iq = 0;
                for (int k = 0;; ++k) {
                    const Item it = load_item(items, k);
                    if (it.b < 0) break;
                    if (it.n_tiles == 0) continue;
                    wait_bar(bar0 + 8 * B_QDOEMPTY, (iq & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bars + B_QDOFULL, 2 * S::QTILE);
                    const int row_q = it.b * T + it.q_tile * BM;
#pragma unroll
                    for (int c = 0; c < D / 64; ++c) {
                        tma_load_2d(smem + S::Q_OFF + c * (BM * 128), &tmQ, bars + B_QDOFULL, it.h * D + c * 64, row_q);
                        tma_load_2d(smem + S::DO_OFF + c * (BM * 128), &tmdO, bars + B_QDOFULL, it.h * D + c * 64, row_q);
                    }
                    ++iq;
                }
            }
        } else if (warp == WARP_KV) {
            // ------------------------------------------------------------ TMA producer for K (whole tiles) and V (64-key halves)
            if (elect_one()) {
                uint32_t ks = 0, kph = 1, vs = 0, vph = 1;            // ring slots, parity of the "slot empty" phase to wait for
                for (int k = 0;; ++k) {
                    const Item it = load_item(items, k);
                    if (it.b < 0) break;
                    const CUtensorMap* tK = it.variant ? &tmK1 : &tmK0;
                    const CUtensorMap* tV = it.variant ? &tmV1 : &tmV0;
                    for (int j = 0; j < it.n_tiles; ++j) {
                        const int row_k = it.b * T + (it.first_tile + j) * BN;
                        wait_bar(bar0 + 8 * (B_KEMPTY + ks), kph);
                        mbar_arrive_expect_tx(bars + B_KFULL + ks, S::QTILE);
#pragma unroll
                        for (int c = 0; c < D / 64; ++c)
                            tma_load_2d(smem + S::K_OFF + ks * S::QTILE + c * (BN * 128), tK, bars + B_KFULL + ks, it.h * D + c * 64, row_k);
                        if (++ks == KST) {
                            ks = 0;
                            kph ^= 1u;
                        }
#pragma unroll
                        for (int hf = 0; hf < 2; ++hf) {
                            wait_bar(bar0 + 8 * (B_VEMPTY + vs), vph);
                            mbar_arrive_expect_tx(bars + B_VFULL + vs, S::VHALF);
#pragma unroll
                            for (int c = 0; c < D / 64; ++c)
                                tma_load_2d(smem + S::V_OFF + vs * S::VHALF + c * (BH * 128), tV, bars + B_VFULL + vs, it.h * D + c * 64,
                                            row_k + hf * BH);
                            if (++vs == VST) {
                                vs = 0;
                                vph ^= 1u;
                            }
                        }
                    }
                }
            }
        } else if (warp == WARP_SDP) {
            // ------------------------------------------------------------ tcgen05 issuer 1: S = Q.K^T and dP = dO.V^T of tile G
            if (elect_one()) {
                constexpr uint32_t idesc_s = make_idesc_bf16(BM, BN, 0, 0);
                constexpr uint32_t idesc_dp = make_idesc_bf16(BM, BH, 0, 0);
                constexpr uint32_t QT16 = (uint32_t)(S::QTILE >> 4), VH16 = (uint32_t)(S::VHALF >> 4);
                const uint32_t dQd = desc_lo_kmajor(smem_u32(smem + S::Q_OFF)), ddO = desc_lo_kmajor(smem_u32(smem + S::DO_OFF));
                const uint32_t dK0 = desc_lo_kmajor(smem_u32(smem + S::K_OFF)), dV0 = desc_lo_kmajor(smem_u32(smem + S::V_OFF));
                const int* ntile = &items[0].n_tiles;                 // stride 8 ints
                int k = 0;
                auto next_count = [&]() {                             // tiles of the next item that has any; -1 at the end
                    int n;
                    do {
                        n = ntile[8 * k];
                        ++k;
                    } while (n == 0);
                    return n;
                };
                int left = next_count();
                bool first = true;
                uint32_t iq = 0, G = 0, ks = 0, kph = 0, vs = 0, vph = 0;
                while (left > 0) {
                    const uint32_t b = G & 1u;
                    DQ_TRACE(0, G);
                    wait_bar(bar0 + 8 * (B_SFREE + b), ((G >> 1) & 1u) ^ 1u);         // dQ(G-2) retired (passes at once for G < 2)
                    DQ_TRACE(1, G);
                    if (first) wait_bar(bar0 + 8 * B_QDOFULL, iq & 1u);
                    wait_bar(bar0 + 8 * (B_KFULL + ks), kph);
                    tc_fence_after_sync();
                    DQ_TRACE(2, G);
                    const uint32_t dK = dK0 + ks * QT16;
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t off = ((uint32_t)(kk / 4) * (BM * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                        umma_ss_lo(tmem_base + COL_S + b * 128, dQd + off, dK + off, idesc_s, kk ? 1u : 0u);
                    }
                    commit_bar(bar0 + 8 * (B_SFULL + b));
                    wait_bar(bar0 + 8 * B_DPFREE, (G & 1u) ^ 1u);                      // dP(G-1) is in registers (passes at once for G = 0)
                    DQ_TRACE(3, G);
#pragma unroll
                    for (int hf = 0; hf < 2; ++hf) {
                        wait_bar(bar0 + 8 * (B_VFULL + vs), vph);
                        tc_fence_after_sync();
                        const uint32_t dV = dV0 + vs * VH16;
#pragma unroll
                        for (int kk = 0; kk < D / 16; ++kk) {
                            const uint32_t offA = ((uint32_t)(kk / 4) * (BM * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                            const uint32_t offB = ((uint32_t)(kk / 4) * (BH * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                            umma_ss_lo(tmem_base + COL_DP + hf * 64, ddO + offA, dV + offB, idesc_dp, kk ? 1u : 0u);
                        }
                        commit_bar(bar0 + 8 * (B_VEMPTY + vs));
                        if (++vs == VST) {
                            vs = 0;
                            vph ^= 1u;
                        }
                    }
                    commit_bar(bar0 + 8 * B_DPFULL);
                    DQ_TRACE(4, G);
                    first = false;
                    if (--left == 0) {                                // Q / dO may be replaced by the next item's
                        commit_bar(bar0 + 8 * B_QDOEMPTY);
                        ++iq;
                        left = next_count();
                        first = true;
                    }
                    if (++ks == KST) {
                        ks = 0;
                        kph ^= 1u;
                    }
                    ++G;
                }
            }
        } else {
            // ------------------------------------------------------------ tcgen05 issuer 2: dQ (+)= dS(G) . K(G)
            if (elect_one()) {
                constexpr uint32_t idesc_dq = make_idesc_bf16(BM, D, 0, 1);
                constexpr uint32_t QT16 = (uint32_t)(S::QTILE >> 4);
                const uint32_t dKmn0 = desc_lo_mnmajor(smem_u32(smem + S::K_OFF), BN * 128);
                const int* ntile = &items[0].n_tiles;
                int k = 0;
                auto next_count = [&]() {
                    int n;
                    do {
                        n = ntile[8 * k];
                        ++k;
                    } while (n == 0);
                    return n;
                };
                int left = next_count();
                bool first = true;
                uint32_t ne = 0, G = 0, ks = 0, kph = 0;
                while (left > 0) {
                    const uint32_t b = G & 1u;
                    DQ_TRACE(5, G);
                    wait_bar(bar0 + 8 * (B_KFULL + ks), kph);                          // (long complete: S(G) read the same slot)
                    if (first && ne > 0) wait_bar(bar0 + 8 * B_DQFREE, (ne - 1) & 1u);  // the previous item's dQ is out of TMEM
                    DQ_TRACE(6, G);
                    wait_bar(bar0 + 8 * (B_DSFULL + b), (G >> 1) & 1u);
                    tc_fence_after_sync();
                    DQ_TRACE(7, G);
                    const uint32_t dKmn = dKmn0 + ks * QT16;
                    const uint32_t a_ds = tmem_base + COL_S + b * 128;
                    const uint32_t acc0 = first ? 0u : 1u;
#pragma unroll
                    for (int kk = 0; kk < BN / 16; ++kk)      // A = dS in TMEM: keys 16kk.. at column 64*(kk/4) + 8*(kk%4); B = K as MN-major (16 key rows = 2048 B)
                        umma_ts_lo(tmem_base + COL_DQ, a_ds + (uint32_t)(kk / 4) * 64 + (uint32_t)(kk % 4) * 8, dKmn + (uint32_t)kk * (2048 >> 4),
                                   idesc_dq, kk ? 1u : acc0);
                    commit_bar(bar0 + 8 * (B_KEMPTY + ks));
                    commit_bar(bar0 + 8 * (B_SFREE + b));
                    DQ_TRACE(8, G);
                    first = false;
                    if (--left == 0) {
                        commit_bar(bar0 + 8 * B_DQFULL);
                        ++ne;
                        left = next_count();
                        first = true;
                    }
                    if (++ks == KST) {
                        ks = 0;
                        kph ^= 1u;
                    }
                    ++G;
                }
            }
        }
    } else {
        // ------------------------------------------------------------ compute warps: thread = query row x 64 keys of every tile
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_COMPUTE));
        const int hf = warp >> 2;                                 // warpgroup = key half of the tile / column half of dQ
        const int r = (warp & 3) * 32 + (threadIdx.x & 31);       // query row in tile == TMEM lane
        const int lane = threadIdx.x & 31;
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const float sl2 = p.scale * LOG2E;
        uint32_t G = 0, ne = 0;                                   // tiles / items with tiles so far
        // per-item row state (destination row, log-sum-exp, delta): two dependent global loads.  The NEXT item's are issued
        // before the current item's epilogue so that their latency (~1.4 k clk in the first version's trace) hides behind it.
        struct RowState {
            int64_t bt;
            bool row_ok;
            float lse2, dlt;
        };
        auto load_row = [&](const Item& it) {
            RowState rs;
            const int qi = it.q_tile * BM + r;
            rs.bt = (int64_t)it.b * T + (qi < T ? qi : 0);
            rs.row_ok = (qi < T) && (!p.qflag || (int)p.qflag[rs.bt] == it.variant);
            const int64_t stat_idx = ((int64_t)it.b * p.heads + it.h) * T + (qi < T ? qi : 0);
            rs.lse2 = rs.row_ok ? p.lse[stat_idx] * LOG2E : CUDART_INF_F;
            rs.dlt = rs.row_ok ? p.delta[stat_idx] : 0.f;
            return rs;
        };
        Item it = load_item(items, 0);
        RowState rs = it.b >= 0 ? load_row(it) : RowState{0, false, 0.f, 0.f};
        for (int ic = 0; it.b >= 0; ++ic) {
            const int n = it.n_tiles;
            const int q0 = it.q_tile * BM, qi = q0 + r;
            const bool row_ok = rs.row_ok;
            const int64_t bt = rs.bt;
            const float lse2 = rs.lse2, dlt = rs.dlt;
            for (int j = 0; j < n; ++j, ++G) {
                const uint32_t b = G & 1u;
                const int kv0 = (it.first_tile + j) * BN + hf * 64;
                const bool need_mask = (CAUSAL && kv0 + 63 > q0) || (kv0 + 64 > it.kve) || (kv0 < it.kvs);
                if ((threadIdx.x & 127) == 0) DQ_TRACE(10 + 5 * hf, G);
                wait_bar(bar0 + 8 * (B_SFULL + b), (G >> 1) & 1u);
                wait_bar(bar0 + 8 * B_DPFULL, G & 1u);
                tc_fence_after_sync();
                if ((threadIdx.x & 127) == 0) DQ_TRACE(11 + 5 * hf, G);
                uint32_t s[64], dp[64];
                tmem_ld32(lane_addr + COL_S + b * 128 + hf * 64, s);
                tmem_ld32(lane_addr + COL_S + b * 128 + hf * 64 + 32, s + 32);
                tmem_ld32(lane_addr + COL_DP + hf * 64, dp);
                tmem_ld32(lane_addr + COL_DP + hf * 64 + 32, dp + 32);
                tc_wait_ld();
                tc_fence_before_sync();
                mbar_arrive(bars + B_DPFREE);                         // dP(G) is in registers: dP(G+1) may overwrite the buffer
                if ((threadIdx.x & 127) == 0) DQ_TRACE(12 + 5 * hf, G);
                if (need_mask) ds_tile<true, CAUSAL>(s, dp, sl2, lse2, dlt, p.scale, kv0, qi, it.kvs, it.kve);
                else           ds_tile<false, CAUSAL>(s, dp, sl2, lse2, dlt, p.scale, kv0, qi, it.kvs, it.kve);
                if ((threadIdx.x & 127) == 0) DQ_TRACE(13 + 5 * hf, G);
                tmem_st32(lane_addr + COL_S + b * 128 + hf * 64, s);  // dS (bf16, 32 columns) over my own, already consumed S columns
                tc_wait_st();
                tc_fence_before_sync();
                mbar_arrive(bars + B_DSFULL + b);
                if ((threadIdx.x & 127) == 0) DQ_TRACE(14 + 5 * hf, G);
            }
            const Item it_next = load_item(items, ic + 1);
            const RowState rs_next = it_next.b >= 0 ? load_row(it_next) : RowState{0, false, 0.f, 0.f};
            const int64_t ld_o = (int64_t)p.heads * D;
            constexpr int DH = D / 2;                                 // dQ columns of this warpgroup
            if (n > 0) {
                uint32_t ov[DH];
                if (threadIdx.x == 0) DQ_TRACE(20, ic);
                wait_bar(bar0 + 8 * B_DQFULL, ne & 1u);
                tc_fence_after_sync();
                if (threadIdx.x == 0) DQ_TRACE(21, ic);
#pragma unroll
                for (int c = 0; c < DH / 32; ++c) tmem_ld32(lane_addr + COL_DQ + hf * DH + c * 32, ov + c * 32);
                tc_wait_ld();
                tc_fence_before_sync();
                mbar_arrive(bars + B_DQFREE);                         // dQ is in registers: the next item's first dQ MMA may start
                if (threadIdx.x == 0) DQ_TRACE(22, ic);
                // dQ leaves through a per-warp staging tile of 8 rows x DH columns (bf16; 16-byte chunks XOR-swizzled by row), four
                // rounds of 8 rows: the 8 lanes owning the rows deposit them, then every store instruction of the warp writes
                // 4 rows x (DH*2 = 128) contiguous bytes -- whole lines.  (The first version wrote 16 rows x 32 bytes per
                // instruction: 2.3 k clk per item, bound by the number of distinct lines / partial-line requests per store.)
                constexpr int CH = DH / 8;                            // 16-byte chunks per row (8 at D = 128, 4 at D = 64)
                constexpr int RPI = 32 / CH;                          // rows per store instruction
                const uint32_t stage_s = smem_u32(smem + S::STAGE_OFF) + (uint32_t)warp * (8 * 128);
                const int32_t dst32 = row_ok ? (int32_t)bt : -1;
                __nv_bfloat16* obase = p.dQ + (int64_t)it.h * D + hf * DH + (lane % CH) * 8;
#pragma unroll
                for (int t = 0; t < 4; ++t) {
                    if ((lane >> 3) == t) {
                        const uint32_t rr = (uint32_t)lane & 7u;
#pragma unroll
                        for (int c = 0; c < CH; ++c) {
                            uint4 o;
                            o.x = pack_bf16(__uint_as_float(ov[c * 8 + 0]), __uint_as_float(ov[c * 8 + 1]));
                            o.y = pack_bf16(__uint_as_float(ov[c * 8 + 2]), __uint_as_float(ov[c * 8 + 3]));
                            o.z = pack_bf16(__uint_as_float(ov[c * 8 + 4]), __uint_as_float(ov[c * 8 + 5]));
                            o.w = pack_bf16(__uint_as_float(ov[c * 8 + 6]), __uint_as_float(ov[c * 8 + 7]));
                            sts128(stage_s + rr * (DH * 2) + ((((uint32_t)c ^ rr) & (CH - 1)) << 4), o);
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 8 / RPI; ++i) {
                        const uint32_t row = (uint32_t)(i * RPI + lane / CH);                  // row within the round
                        const int32_t drow = __shfl_sync(0xffffffffu, dst32, t * 8 + (int)row);
                        const uint4 o = lds128(stage_s + row * (DH * 2) + (((((uint32_t)lane % CH) ^ row) & (CH - 1)) << 4));
                        if (drow >= 0) *reinterpret_cast<uint4*>(obase + (int64_t)drow * ld_o) = o;
                    }
                    __syncwarp();
                }
                if (threadIdx.x == 0) DQ_TRACE(23, ic);
                ++ne;
            } else if (row_ok) {                                      // no visible key at all: the gradient row is zero
                uint4* orow = reinterpret_cast<uint4*>(p.dQ + bt * ld_o + (int64_t)it.h * D + hf * DH);
#pragma unroll
                for (int e = 0; e < DH / 8; ++e) orow[e] = make_uint4(0u, 0u, 0u, 0u);
            }
            it = it_next;
            rs = rs_next;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == WARP_SDP) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

template <int D, bool CAUSAL, bool TRACE>
static int launch_t(const CUtensorMap* tm, const Params& p, cudaStream_t st) {
    using S = Smem<D>;
    auto kern = attn_bwd_dq_stream_kernel<D, CAUSAL, TRACE>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail(LB_ELAUNCH, "attn_bwd_dq_stream: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    kern<<<(unsigned)p.n_cta, THREADS, S::TOTAL, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], p);
    return check_launch("attn_bwd_dq_stream");
}

template <int D, bool CAUSAL>
static int launch(const CUtensorMap* tm, const Params& p, cudaStream_t st) {
    // the clock64 stamps are compiled into a separate instantiation (diagnostics, head_dim 128 causal only)
    if (D == 128 && CAUSAL && p.trace) return launch_t<D, CAUSAL, true>(tm, p, st);
    return launch_t<D, CAUSAL, false>(tm, p, st);
}

}  // namespace dqs
}  // namespace lb

using namespace lb;

static long long* g_dqs_trace = nullptr;

/* diagnostics: CTA 0 of subsequent lb_attn_bwd_dq_stream launches (head_dim 128, causal) writes clock64 stamps into `buf`
 * ([64][32] int64, device; one row per tile: slots 0-4 S/dP issuer, 5-8 dQ issuer, 10-14 / 15-19 thread 0 of compute
 * warpgroup 0 / 1; rows indexed by item: 20-23 epilogue).  NULL = off */
extern "C" int lb_attn_bwd_dq_stream_set_trace(void* buf) {
    g_dqs_trace = (long long*)buf;
    return LB_OK;
}

extern "C" int lb_attn_bwd_dq_stream_max_cta_items(void) { return dqs::MAX_ITEMS; }

extern "C" int lb_attn_bwd_dq_stream(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1, const void* dO,
                                     const float* lse, const float* delta, const uint8_t* qflag, const int32_t* work, int n_work,
                                     const int32_t* plan_items, const int32_t* plan_off, int n_cta, int max_cta_items,
                                     int head_group, const int32_t* kv_start, const int32_t* kv_end, void* dQ, int batch,
                                     int seqlen, int heads, int head_dim, int causal, float scale, void* stream) {
    LB_REQUIRE(batch > 0 && seqlen > 0 && heads > 0 && n_work >= 0, LB_EINVAL, "attn_bwd_dq_stream: bad shape");
    LB_REQUIRE(head_dim == 64 || head_dim == 128, LB_EINVAL, "attn_bwd_dq_stream: head_dim %d (64 or 128 supported)", head_dim);
    LB_REQUIRE(Q && K0 && V0 && dO && lse && delta && work && dQ, LB_EINVAL, "attn_bwd_dq_stream: null argument");
    LB_REQUIRE((plan_items == nullptr) == (plan_off == nullptr), LB_EINVAL, "attn_bwd_dq_stream: plan_items and plan_off go together");
    LB_REQUIRE(!plan_items || (n_cta > 0 && max_cta_items > 0), LB_EINVAL,
               "attn_bwd_dq_stream: a plan needs n_cta (%d) and max_cta_items (%d)", n_cta, max_cta_items);
    LB_REQUIRE(((uintptr_t)dQ & 15) == 0, LB_EALIGN, "attn_bwd_dq_stream: dQ must be 16-byte aligned");
    if (n_work == 0) return LB_OK;
    int rc = require_sm100();
    if (rc) return rc;
    const uint64_t rows = (uint64_t)batch * seqlen, cols = (uint64_t)heads * head_dim;
    CUtensorMap tm[6];
    const void* ptrs[6] = {Q, dO, K0, V0, K1 ? K1 : K0, V1 ? V1 : V0};
    for (int i = 0; i < 6; ++i) {
        rc = make_tmap_bf16_2d(&tm[i], ptrs[i], rows, cols, cols, (i == 3 || i == 5) ? dqs::BH : dqs::BM, 64);      // V: 64-key half tiles
        if (rc) return rc;
    }
    dqs::Params p;
    p.qflag = qflag; p.work = work; p.kv_start = kv_start; p.kv_end = kv_end; p.lse = lse; p.delta = delta;
    p.dQ = (__nv_bfloat16*)dQ; p.batch = batch; p.seqlen = seqlen; p.heads = heads; p.scale = scale;
    p.n_work = n_work; p.head_group = head_group > 0 ? head_group : attn_head_group(); p.n_items = n_work * heads;
    p.plan_items = plan_items; p.plan_off = plan_off;
    p.trace = g_dqs_trace;
    int per_cta;
    if (plan_items) {
        p.n_cta = n_cta;
        per_cta = max_cta_items;
    } else {
        const int sms = sm_count();
        if (sms <= 0) return fail(LB_ELAUNCH, "attn_bwd_dq_stream: no SM count");
        p.n_cta = p.n_items < sms ? p.n_items : sms;
        per_cta = (p.n_items + p.n_cta - 1) / p.n_cta;
    }
    LB_REQUIRE(per_cta <= dqs::MAX_ITEMS, LB_EINVAL,
               "attn_bwd_dq_stream: %d items per CTA exceed the in-kernel table (%d); use lb_attn_bwd_dq for this shape", per_cta,
               dqs::MAX_ITEMS);
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 128) return causal ? dqs::launch<128, true>(tm, p, st) : dqs::launch<128, false>(tm, p, st);
    return causal ? dqs::launch<64, true>(tm, p, st) : dqs::launch<64, false>(tm, p, st);
}

// Bridge attention backward, dK/dV: persistent streaming kernel (A10; same maths, operands and work list as attn_bwd_dkv.cu).
//
// What the one-CTA-per-item kernel measured (attn_bwd_dkv.cu, scripts/attn_dkv_ctalog.py): per CTA 2.1 k clk of prologue
// (TMEM allocation, barrier init, q-tile list) + 1.9 k clk waiting for K/V + 3.0 k clk per q tile (tensor work: 2.0 k) + 2.6 k
// clk of epilogue that nothing overlaps, i.e. ~6.6 k of ~32 k clk per item outside the main loop, and a main loop whose 48
// MMAs per tile (32 of them N = 64: 32 clk of tensor work each, ~40-55 clk to issue next to busy compute warps) come from ONE
// issuing thread.  Here:
//   * one CTA per SM walks a host-balanced share of the (work item, head) list; prologue once per CTA; an item's q-tile list
//     is a 64-bit mask decoded once into shared memory;
//   * K and V of the NEXT item are loaded as soon as the last score MMAs of the current item retired (they are the A operands
//     of S^T and dP^T only), i.e. during the item's last dS computation, gradient MMAs and epilogue;
//   * the epilogue goes through a DEDICATED 32 KB staging tile and TMA stores: the compute warps pull their 64 accumulator
//     columns into registers (dV/dK released to the next item ~200 clk after the last MMA), the dV warpgroups deposit bf16 rows
//     (128-byte swizzle) and one thread stores the tile with two 16 KB TMA boxes; the dK warpgroups follow when the
//     TMA engine has read the staging tile.  The dV warpgroups are the compute warps of query half A and the dK warpgroups
//     those of half B, so the next item's first half-tile is already being computed while dK leaves;
//   * two issuing threads: scores (S^T, dP^T per 64-query half) and gradients (dV, dK per half).
// Pipeline per q tile as in attn_bwd_dkv.cu: two independent 64-query halves; P^T / dS^T (bf16) overwrite the score columns
// of their half and are the TMEM A operands of dV += P^T.dO, dK += dS^T.Q; the scores of the next tile's half wait for those
// MMAs to retire.  TMEM: S^T [0,128) dP^T [128,256) dV [256,256+D) dK [256+D, 256+2D).
// CTA = 640 threads: warps 0-15 compute (4 warpgroups; warpgroup g: half g/2, query columns [32g, 32g+32), thread = key row),
// warp 16 TMA loads, warp 17 tcgen05 issuer for the scores (+ TMEM alloc), warp 18 tcgen05 issuer for the gradients
// (warp 19 idle; setmaxnreg 104 / 56).
// Shared memory is the binding budget: K 32 + V 32 + 2 x (Q 32 + dO 32) + staging 32 KB = 224 KB, leaving 3 KB for the
// statistics, the item table and the barriers; the kernel relies on the 1024-byte alignment of the dynamic shared memory
// window (checked at run time; the host falls back to lb_attn_bwd_dkv if a platform does not provide it).
// Deterministic: no atomics; every dK/dV element is written by exactly one thread, accumulation order fixed by the tile order.
#include <math_constants.h>

#include "common.cuh"

namespace lb {
namespace dks {

constexpr float LOG2E_F = 1.4426950408889634f;
constexpr int WARP_LOAD = 16, WARP_SC = 17, WARP_GR = 18, THREADS = 20 * 32;      // warp 19 idles (setmaxnreg works on whole warpgroups)
constexpr int REGS_COMPUTE = 104, REGS_PRODUCER = 56;      // 512 x 104 + 128 x 56 <= 640 x 96 (what the launch allocated)
constexpr int MAX_ITEMS = 40;
constexpr int SMEM_MAX = 227 * 1024;

struct Params {
    const uint8_t* qflag;      // [B*T] or null
    const uint8_t* qtile_has;  // [B,2,n_qtiles] or null: q tile holds rows of the variant
    const int32_t* work;       // [n][4] = {b, kv_tile, variant, first_q_tile}
    const int32_t* kv_start;
    const int32_t* kv_end;
    const float* lse;          // [B,H,T]
    const float* delta;        // [B,H,T]
    int batch, seqlen, heads;
    int n_work, head_group, n_items;
    const int32_t* plan_items; // [n_items] list positions grouped by CTA, or null: snake split
    const int32_t* plan_off;   // [gridDim.x + 1]
    float scale;
    int smem_bytes;            // dynamic shared memory of the launch
};

// 16 bytes per item: packed = b | h << 12 | variant << 22 | kv_tile << 23 (b < 4096, h < 1024, kv_tile < 256 -- host-checked)
struct __align__(16) Item {
    int packed;                // < 0: end of the CTA's list
    int kvs_kve;               // kvs | kve << 16 (T <= 8192)
    uint32_t mask_lo, mask_hi; // q tiles to visit (bit = q tile index)
};

template <int D>
struct Smem {
    static constexpr int TILE = 128 * D * 2;                          // K / V / Q / dO tile
    static constexpr int K_OFF = 0, V_OFF = TILE, Q_OFF = 2 * TILE;   // stage s: Q at Q_OFF + s * 2 * TILE, dO right after
    static constexpr int STAGE_OFF = Q_OFF + 4 * TILE;                // epilogue staging: 128 rows x D columns bf16
    static constexpr int STAT_OFF = STAGE_OFF + TILE;                 // float [2 parity][2 half][-lse2 x 64 | -delta*scale x 64]
    static constexpr int ITEM_OFF = STAT_OFF + 2 * 2 * 128 * 4;       // Item[MAX_ITEMS + 1]
    static constexpr int BAR_OFF = ITEM_OFF + (MAX_ITEMS + 1) * 16;
    static constexpr int NEEDED = BAR_OFF + 256;
    static_assert(NEEDED <= SMEM_MAX, "shared memory budget");
};

enum {
    B_KVFULL = 0,                    // K and V of an item landed
    B_KVFREE,                        // the item's last score MMAs retired: K / V may be replaced
    B_QFULL,                         // [2] Q and dO of a q tile landed in stage s
    B_QEMPTY = B_QFULL + 2,          // [2] the tile's last gradient MMAs retired
    B_SDP = B_QEMPTY + 2,            // [2] scores of half h landed
    B_PDS = B_SDP + 2,               // [2] P^T / dS^T of half h written (256 arrivals)
    B_SFREE = B_PDS + 2,             // [2] gradient MMAs of half h retired: its score columns are reusable
    B_DONE = B_SFREE + 2,            // the item's last gradient MMAs retired
    B_ACCFREE,                       // dV / dK of the item are in registers (512 arrivals)
    B_STAGE,                         // the TMA engine has read the staging tile (two completions per item: after dV, after dK)
    B_COUNT
};

__device__ __forceinline__ int item_of_round(const Params& p, int k) {
    if (p.plan_items) {
        const int i = p.plan_off[blockIdx.x] + k;
        return i < p.plan_off[blockIdx.x + 1] ? p.plan_items[i] : -1;
    }
    const int G = (int)gridDim.x, c = (int)blockIdx.x;
    if ((int64_t)k * G >= p.n_items) return -1;
    const int L = k * G + ((k & 1) ? G - 1 - c : c);
    return L < p.n_items ? L : -1;
}

__device__ __forceinline__ Item decode_item(const Params& p, int L) {
    Item it;
    it.packed = -1;
    it.kvs_kve = 0;
    it.mask_lo = it.mask_hi = 0u;
    if (L < 0) return it;
    const int per_group = p.head_group * p.n_work;
    const int g = L / per_group;
    const int rem = L - g * per_group;
    const int gl = min(p.head_group, p.heads - g * p.head_group);
    const int w = rem / gl;
    const int h = g * p.head_group + (rem - w * gl);
    const int b = p.work[w * 4 + 0], kv_tile = p.work[w * 4 + 1], variant = p.work[w * 4 + 2], first_q = p.work[w * 4 + 3];
    const int nqt = (p.seqlen + 127) / 128;
    uint64_t m = 0;
    for (int qt = first_q; qt < nqt; ++qt)
        if (!p.qtile_has || p.qtile_has[((int64_t)b * 2 + variant) * nqt + qt]) m |= 1ull << qt;
    it.packed = b | (h << 12) | (variant << 22) | (kv_tile << 23);
    it.kvs_kve = (p.kv_start ? p.kv_start[b] : 0) | ((p.kv_end ? p.kv_end[b] : p.seqlen) << 16);
    it.mask_lo = (uint32_t)m;
    it.mask_hi = (uint32_t)(m >> 32);
    return it;
}

struct ItemV {
    int b, h, variant, kv_tile, kvs, kve;
    uint64_t mask;
};
__device__ __forceinline__ bool load_item(const Item* tab, int k, ItemV& v) {
    const int4 a = *reinterpret_cast<const int4*>(tab + k);
    if (a.x < 0) return false;
    v.b = a.x & 0xfff; v.h = (a.x >> 12) & 0x3ff; v.variant = (a.x >> 22) & 1; v.kv_tile = (a.x >> 23) & 0xff;
    v.kvs = a.y & 0xffff; v.kve = (a.y >> 16) & 0xffff;
    v.mask = (uint64_t)(uint32_t)a.z | ((uint64_t)(uint32_t)a.w << 32);
    return true;
}
__device__ __forceinline__ int pop_tile(uint64_t& m) {      // lowest set bit = next q tile
    const int qt = __ffsll((long long)m) - 1;
    m &= m - 1;
    return qt;
}

// thread = key row; 32 query columns at TMEM `ts` (S^T) / `tdp` (dP^T); per-column -lse2 / -delta*scale in shared memory.
// Packed fp32x2 arithmetic: per PAIR of query columns FFMA2 (exponent), 2 x MUFU.EX2, FFMA2, FMUL2, two bf16x2 packs.
template <bool MASK, bool CAUSAL>
__device__ __forceinline__ void dkv_tile(uint32_t ts, uint32_t tdp, uint32_t st_s, float sl2, float scale, int kj, int qbase,
                                         bool key_ok) {
    uint32_t s[32], dp[32];
    tmem_ld32(ts, s);
    tmem_ld32(tdp, dp);
    tc_wait_ld();
    const uint64_t sl2_2 = f32x2(sl2, sl2), sc_2 = f32x2(scale, scale);
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
        float4 ls, dl;
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(ls.x), "=f"(ls.y), "=f"(ls.z), "=f"(ls.w) : "r"(st_s + j * 4));
        asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(dl.x), "=f"(dl.y), "=f"(dl.z), "=f"(dl.w) : "r"(st_s + 256 + j * 4));
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

__device__ __forceinline__ void tma_store_3d_addr(const CUtensorMap* m, uint32_t smem_src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(reinterpret_cast<uint64_t>(m)),
                 "r"(smem_src), "r"(c0), "r"(c1), "r"(c2)
                 : "memory");
}

template <int D, bool CAUSAL>
__global__ void __launch_bounds__(THREADS, 1)
attn_bwd_dkv_stream_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmdO,
                           const __grid_constant__ CUtensorMap tmK0, const __grid_constant__ CUtensorMap tmV0,
                           const __grid_constant__ CUtensorMap tmK1, const __grid_constant__ CUtensorMap tmV1,
                           const __grid_constant__ CUtensorMap tmdK0, const __grid_constant__ CUtensorMap tmdV0,
                           const __grid_constant__ CUtensorMap tmdK1, const __grid_constant__ CUtensorMap tmdV1, const Params p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using S = Smem<D>;
    if ((int)(smem - smem_raw) + S::NEEDED > p.smem_bytes) __trap();      // the window was not 1024-byte aligned: see the host check
    Item* items = reinterpret_cast<Item*>(smem + S::ITEM_OFF);
    float* stats = reinterpret_cast<float*>(smem + S::STAT_OFF);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);
    const uint32_t bar0 = smem_u32(bars);
    const uint32_t smem_s = smem_u32(smem);

    const int warp = threadIdx.x >> 5;
    const int T = p.seqlen;
    constexpr uint32_t TMEM_COLS = 512, COL_S = 0, COL_DP = 128, COL_DV = 256, COL_DK = 256 + D;

    if (threadIdx.x == 0) {
        for (int i = 0; i < B_COUNT; ++i)
            mbar_init(bars + i, (i == B_PDS || i == B_PDS + 1) ? 256 : (i == B_ACCFREE ? 512 : 1));
        fence_barrier_init();
    }
    for (int k = threadIdx.x; k <= MAX_ITEMS; k += THREADS) items[k] = decode_item(p, item_of_round(p, k));
    if (warp == WARP_LOAD && elect_one()) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmdO); tma_prefetch_desc(&tmK0); tma_prefetch_desc(&tmV0);
        tma_prefetch_desc(&tmK1); tma_prefetch_desc(&tmV1);
    }
    if (warp == WARP_SC) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // (setmaxnreg sits at the top of each side's own branch: a join of paths with different register counts does not compile)
    if (warp >= 16) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
    if (warp == WARP_LOAD) {
        // ------------------------------------------------------------ TMA producer: K/V per item, Q/dO per q tile (ring of 2)
        if (elect_one()) {
            uint32_t ni = 0, nt = 0;                                      // items / tiles issued so far
            ItemV it;
            for (int k = 0; load_item(items, k, it); ++k) {
                if (it.mask == 0) continue;
                const CUtensorMap* tK = it.variant ? &tmK1 : &tmK0;
                const CUtensorMap* tV = it.variant ? &tmV1 : &tmV0;
                const int row_k = it.b * T + it.kv_tile * 128;
                wait_bar(bar0 + 8 * B_KVFREE, (ni & 1u) ^ 1u);
                mbar_arrive_expect_tx(bars + B_KVFULL, 2 * S::TILE);
#pragma unroll
                for (int c = 0; c < D / 64; ++c) {
                    tma_load_2d(smem + S::K_OFF + c * (128 * 128), tK, bars + B_KVFULL, it.h * D + c * 64, row_k);
                    tma_load_2d(smem + S::V_OFF + c * (128 * 128), tV, bars + B_KVFULL, it.h * D + c * 64, row_k);
                }
                ++ni;
                uint64_t m = it.mask;
                while (m) {
                    const int qt = pop_tile(m);
                    const uint32_t st = nt & 1u;
                    uint8_t* q = smem + S::Q_OFF + st * (2 * S::TILE);
                    wait_bar(bar0 + 8 * (B_QEMPTY + st), ((nt >> 1) & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bars + B_QFULL + st, 2 * S::TILE);
#pragma unroll
                    for (int c = 0; c < D / 64; ++c) {
                        tma_load_2d(q + c * (128 * 128), &tmQ, bars + B_QFULL + st, it.h * D + c * 64, it.b * T + qt * 128);
                        tma_load_2d(q + S::TILE + c * (128 * 128), &tmdO, bars + B_QFULL + st, it.h * D + c * 64, it.b * T + qt * 128);
                    }
                    ++nt;
                }
            }
        }
    } else if (warp == WARP_SC) {
        // ------------------------------------------------------------ tcgen05 issuer 1: S^T = K.Q^T, dP^T = V.dO^T per 64-query half
        if (elect_one()) {
            constexpr uint32_t idesc_h = make_idesc_bf16(128, 64, 0, 0);
            const uint32_t dK0 = desc_lo_kmajor(smem_s + S::K_OFF), dV0 = desc_lo_kmajor(smem_s + S::V_OFF);
            uint32_t ni = 0, nt = 0;
            ItemV it;
            for (int k = 0; load_item(items, k, it); ++k) {
                if (it.mask == 0) continue;
                wait_bar(bar0 + 8 * B_KVFULL, ni & 1u);
                int left = __popcll(it.mask);
                for (; left > 0; --left, ++nt) {
                    const uint32_t st = nt & 1u;
                    const uint32_t aQ = smem_s + S::Q_OFF + st * (2 * S::TILE);
                    const uint32_t dQk = desc_lo_kmajor(aQ), ddOk = desc_lo_kmajor(aQ + S::TILE);
                    wait_bar(bar0 + 8 * (B_QFULL + st), (nt >> 1) & 1u);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        wait_bar(bar0 + 8 * (B_SFREE + half), (nt & 1u) ^ 1u);        // gradients of (previous tile, half) retired
                        tc_fence_after_sync();
                        const uint32_t hoff = (uint32_t)(half * 64 * 128) >> 4;        // 64 query rows further down each 16 KB chunk
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
                        commit_bar(bar0 + 8 * (B_SDP + half));
                    }
                }
                commit_bar(bar0 + 8 * B_KVFREE);                                      // K / V of this item are dead
                ++ni;
            }
        }
    } else if (warp == WARP_GR) {
        // ------------------------------------------------------------ tcgen05 issuer 2: dV += P^T.dO, dK += dS^T.Q per half
        if (elect_one()) {
            constexpr uint32_t idesc_g = make_idesc_bf16(128, D, 0, 1);
            uint32_t ni = 0, nt = 0;
            ItemV it;
            for (int k = 0; load_item(items, k, it); ++k) {
                if (it.mask == 0) continue;
                int left = __popcll(it.mask);
                bool first = true;
                if (ni > 0) wait_bar(bar0 + 8 * B_ACCFREE, (ni - 1) & 1u);            // the previous item's dV / dK are out of TMEM
                for (; left > 0; --left, ++nt) {
                    const uint32_t st = nt & 1u;
                    const uint32_t aQ = smem_s + S::Q_OFF + st * (2 * S::TILE);
                    const uint32_t dQmn = desc_lo_mnmajor(aQ, 128 * 128), ddOmn = desc_lo_mnmajor(aQ + S::TILE, 128 * 128);
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        wait_bar(bar0 + 8 * (B_PDS + half), nt & 1u);
                        tc_fence_after_sync();
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)      // 64 queries of this half = 4 K steps; P^T of quarter kk/2 at 64*half + 32*(kk/2) + 8*(kk%2)
                            umma_ts_lo(tmem_base + COL_DV, tmem_base + COL_S + (uint32_t)half * 64 + (uint32_t)(kk / 2) * 32 + (uint32_t)(kk % 2) * 8,
                                       ddOmn + (uint32_t)(half * 4 + kk) * (2048 >> 4), idesc_g, (first && half == 0 && kk == 0) ? 0u : 1u);
#pragma unroll
                        for (int kk = 0; kk < 4; ++kk)
                            umma_ts_lo(tmem_base + COL_DK, tmem_base + COL_DP + (uint32_t)half * 64 + (uint32_t)(kk / 2) * 32 + (uint32_t)(kk % 2) * 8,
                                       dQmn + (uint32_t)(half * 4 + kk) * (2048 >> 4), idesc_g, (first && half == 0 && kk == 0) ? 0u : 1u);
                        commit_bar(bar0 + 8 * (B_SFREE + half));
                    }
                    commit_bar(bar0 + 8 * (B_QEMPTY + st));
                    first = false;
                }
                commit_bar(bar0 + 8 * B_DONE);
                ++ni;
            }
        }
    }
    } else {
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_COMPUTE));
        // ------------------------------------------------------------ compute warps: thread <-> kv row (TMEM lane)
        const int wg = warp >> 2, half = wg >> 1, quarter = wg & 1;
        const int r = (warp & 3) * 32 + (threadIdx.x & 31);
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const uint32_t colS = COL_S + wg * 32, colDP = COL_DP + wg * 32;
        const float sl2 = p.scale * LOG2E_F;
        const int t256 = threadIdx.x & 255;      // thread within the half's 256-thread group
        uint32_t ni = 0, nt = 0;
        ItemV it;
        for (int k = 0; load_item(items, k, it); ++k) {
            const int kv0 = it.kv_tile * 128, kj = kv0 + r;
            const bool key_ok = kj < it.kve && kj >= it.kvs;
            uint64_t m = it.mask;
            if (m == 0) continue;
            while (m) {
                const int qt = pop_tile(m);
                const int q0 = qt * 128, qhalf = q0 + half * 64;
                // per-column statistics of this half's 64 query columns: [parity][half][-lse2 x 64 | -delta*scale x 64]
                float* sl = stats + (nt & 1) * 256 + half * 128;
                if (t256 < 128) {
                    const int col = t256 & 63;
                    const int qi = qhalf + col;
                    const bool ok = qi < T && (!p.qflag || (int)p.qflag[(int64_t)it.b * T + (qi < T ? qi : 0)] == it.variant);
                    const int64_t si = ((int64_t)it.b * p.heads + it.h) * T + (qi < T ? qi : 0);
                    if (t256 < 64) sl[col] = ok ? -p.lse[si] * LOG2E_F : -CUDART_INF_F;
                    else           sl[64 + col] = ok ? -p.delta[si] * p.scale : 0.f;
                }
                named_bar_sync(1 + half, 256);
                wait_bar(bar0 + 8 * (B_SDP + half), nt & 1u);
                tc_fence_after_sync();
                const int qbase = qhalf + quarter * 32;
                const bool need_mask = (kv0 + 128 > it.kve) || (kv0 < it.kvs) || (CAUSAL && kv0 + 127 > qbase);
                const uint32_t st_s = smem_u32(sl + quarter * 32);
                if (need_mask) dkv_tile<true, CAUSAL>(lane_addr + colS, lane_addr + colDP, st_s, sl2, p.scale, kj, qbase, key_ok);
                else           dkv_tile<false, CAUSAL>(lane_addr + colS, lane_addr + colDP, st_s, sl2, p.scale, kj, qbase, key_ok);
                tc_wait_st();
                tc_fence_before_sync();
                mbar_arrive(bars + B_PDS + half);
                ++nt;
            }
            // ---- epilogue: warpgroups 0,1 hold the two column halves of dV, warpgroups 2,3 those of dK
            constexpr int DH = D / 2;
            uint32_t v[DH];
            wait_bar(bar0 + 8 * B_DONE, ni & 1u);
            tc_fence_after_sync();
            const uint32_t col0 = (wg < 2 ? COL_DV : COL_DK) + (wg & 1) * DH;
#pragma unroll
            for (int c = 0; c < DH / 32; ++c) tmem_ld32(lane_addr + col0 + c * 32, v + c * 32);
            tc_wait_ld();
            tc_fence_before_sync();
            mbar_arrive(bars + B_ACCFREE);                                            // the next item's gradients may start
            // staging tile: D/64 chunks of [128 rows x 128 B], 16-byte units XOR-swizzled by row (SWIZZLE_128B); this warp's DH
            // columns are units [u0, u0 + DH/8) of chunk ch of its 32 rows
            const int ch = D == 128 ? (wg & 1) : 0;
            const int u0 = D == 128 ? 0 : (wg & 1) * (DH / 8);
            // the dV warpgroups wait until the previous item's dK left the staging tile, the dK warpgroups until this item's dV did
            const uint32_t sphase = 2u * ni + (wg < 2 ? 0u : 1u);                     // completions of B_STAGE that must have happened
            if (sphase > 0) wait_bar(bar0 + 8 * B_STAGE, (sphase - 1) & 1u);
            const uint32_t row_s = smem_s + S::STAGE_OFF + (uint32_t)ch * (128 * 128) + (uint32_t)r * 128;
#pragma unroll
            for (int j = 0; j < DH / 8; ++j) {
                uint4 o;
                o.x = pack_bf16(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
                o.y = pack_bf16(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
                o.z = pack_bf16(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
                o.w = pack_bf16(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
                sts128(row_s + ((((uint32_t)(u0 + j)) ^ ((uint32_t)r & 7u)) << 4), o);
            }
            fence_proxy_async_smem();
            named_bar_sync(3 + (wg >> 1), 256);                                       // the 8 warps of this tensor deposited
            if ((threadIdx.x & 255) == 0) {
                const CUtensorMap* tm = wg < 2 ? (it.variant ? &tmdV1 : &tmdV0) : (it.variant ? &tmdK1 : &tmdK0);
#pragma unroll
                for (int c = 0; c < D / 64; ++c)
                    tma_store_3d_addr(tm, smem_s + S::STAGE_OFF + c * (128 * 128), it.h * D + c * 64, kv0, it.b);
                tma_store_commit();
                tma_store_wait_read();                                                // the staging tile has been read
                mbar_arrive(bars + B_STAGE);
            }
            ++ni;
        }
        if ((threadIdx.x & 255) == 0) tma_store_wait_all();
        tc_fence_before_sync();
    }
    __syncthreads();
    if (warp == WARP_SC) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
}

__global__ void smem_align_probe_kernel(int* out) {
    extern __shared__ __align__(1024) uint8_t probe_smem[];
    if (threadIdx.x == 0) *out = (int)(smem_u32(probe_smem) & 1023u);
}

// (base of the dynamic shared memory window) mod 1024 for a launch with the kernel's footprint; cached.  < 0: error
static int smem_window_misalignment() {
    static int cached = -2;
    if (cached != -2) return cached;
    int* d = nullptr;
    int h = -1;
    if (cudaMalloc(&d, sizeof(int)) != cudaSuccess) return cached = -1;
    cudaFuncSetAttribute(smem_align_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
    smem_align_probe_kernel<<<1, 32, SMEM_MAX>>>(d);
    if (cudaMemcpy(&h, d, sizeof(int), cudaMemcpyDeviceToHost) != cudaSuccess) h = -1;
    cudaFree(d);
    return cached = h;
}

template <int D, bool CAUSAL>
static int launch(const CUtensorMap* tm, Params& p, int n_cta, cudaStream_t st) {
    auto kern = attn_bwd_dkv_stream_kernel<D, CAUSAL>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_MAX);
        if (e != cudaSuccess) return fail(LB_ELAUNCH, "attn_bwd_dkv_stream: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    p.smem_bytes = SMEM_MAX;
    kern<<<(unsigned)n_cta, THREADS, SMEM_MAX, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], tm[5], tm[6], tm[7], tm[8], tm[9], p);
    return check_launch("attn_bwd_dkv_stream");
}

}  // namespace dks
}  // namespace lb

using namespace lb;

extern "C" {

int lb_attn_bwd_dkv_stream_max_cta_items(void) { return dks::MAX_ITEMS; }

/* 1 when the persistent dK/dV kernel can run on this device (its 224 KB footprint needs the dynamic shared-memory window to start
 * on a 1024-byte boundary; probed once with a one-thread kernel on the first call -- call it outside stream capture) */
int lb_attn_bwd_dkv_stream_supported(void) { return dks::smem_window_misalignment() == 0 ? 1 : 0; }

int lb_attn_bwd_dkv_stream(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1, const void* dO,
                           const float* lse, const float* delta, const uint8_t* qflag, const uint8_t* qtile_has,
                           const int32_t* work_kv, int n_work, const int32_t* plan_items, const int32_t* plan_off, int n_cta,
                           int max_cta_items, int head_group, const int32_t* kv_start, const int32_t* kv_end, void* dK0, void* dV0,
                           void* dK1, void* dV1, int batch, int seqlen, int heads, int head_dim, int causal, float scale,
                           void* stream) {
    LB_REQUIRE(batch > 0 && seqlen > 0 && heads > 0 && n_work >= 0, LB_EINVAL, "attn_bwd_dkv_stream: bad shape");
    LB_REQUIRE(head_dim == 64 || head_dim == 128, LB_EINVAL, "attn_bwd_dkv_stream: head_dim %d (64 or 128 supported)", head_dim);
    LB_REQUIRE(seqlen <= 8192 && batch < 4096 && heads < 1024, LB_EINVAL, "attn_bwd_dkv_stream: seqlen %d <= 8192, batch %d < 4096, heads %d < 1024",
               seqlen, batch, heads);
    LB_REQUIRE(Q && K0 && V0 && dO && lse && delta && work_kv && dK0 && dV0, LB_EINVAL, "attn_bwd_dkv_stream: null argument");
    LB_REQUIRE((plan_items == nullptr) == (plan_off == nullptr), LB_EINVAL, "attn_bwd_dkv_stream: plan_items and plan_off go together");
    LB_REQUIRE(!plan_items || (n_cta > 0 && max_cta_items > 0), LB_EINVAL, "attn_bwd_dkv_stream: a plan needs n_cta and max_cta_items");
    if (n_work == 0) return LB_OK;
    int rc = require_sm100();
    if (rc) return rc;
    LB_REQUIRE(dks::smem_window_misalignment() == 0, LB_EINVAL,
               "attn_bwd_dkv_stream: dynamic shared memory window is not 1024-byte aligned on this platform; use lb_attn_bwd_dkv");
    const uint64_t rows = (uint64_t)batch * seqlen, cols = (uint64_t)heads * head_dim;
    CUtensorMap tm[10];
    const void* ptrs[6] = {Q, dO, K0, V0, K1 ? K1 : K0, V1 ? V1 : V0};
    for (int i = 0; i < 6; ++i) {
        rc = make_tmap_bf16_2d(&tm[i], ptrs[i], rows, cols, cols, 128, 64);
        if (rc) return rc;
    }
    void* outs[4] = {dK0, dV0, dK1 ? dK1 : dK0, dV1 ? dV1 : dV0};
    for (int i = 0; i < 4; ++i) {
        rc = make_tmap_bf16_3d(&tm[6 + i], outs[i], (uint64_t)batch, (uint64_t)seqlen, cols, 128, 64);
        if (rc) return rc;
    }
    dks::Params p{};
    p.qflag = qflag; p.qtile_has = qtile_has; p.work = work_kv; p.kv_start = kv_start; p.kv_end = kv_end; p.lse = lse; p.delta = delta;
    p.batch = batch; p.seqlen = seqlen; p.heads = heads; p.scale = scale;
    p.n_work = n_work; p.head_group = head_group > 0 ? head_group : attn_head_group(); p.n_items = n_work * heads;
    p.plan_items = plan_items; p.plan_off = plan_off;
    int per_cta, grid;
    if (plan_items) {
        grid = n_cta;
        per_cta = max_cta_items;
    } else {
        const int sms = sm_count();
        if (sms <= 0) return fail(LB_ELAUNCH, "attn_bwd_dkv_stream: no SM count");
        // more items than one wave of CTAs can hold: several CTAs per SM in sequence (the later ones start as SMs free up)
        int waves = (p.n_items + sms * dks::MAX_ITEMS - 1) / (sms * dks::MAX_ITEMS);
        if (waves < 1) waves = 1;
        grid = p.n_items < sms * waves ? p.n_items : sms * waves;
        per_cta = (p.n_items + grid - 1) / grid;
    }
    LB_REQUIRE(per_cta <= dks::MAX_ITEMS, LB_EINVAL, "attn_bwd_dkv_stream: %d items per CTA exceed the in-kernel table (%d)", per_cta,
               dks::MAX_ITEMS);
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 128) return causal ? dks::launch<128, true>(tm, p, grid, st) : dks::launch<128, false>(tm, p, grid, st);
    return causal ? dks::launch<64, true>(tm, p, grid, st) : dks::launch<64, false>(tm, p, grid, st);
}

}

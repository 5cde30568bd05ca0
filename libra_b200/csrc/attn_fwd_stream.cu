// Bridge attention forward, persistent streaming kernel (A10; same maths and operand formulation as attn_fwd.cu).
//
// One CTA per SM walks a static share of the (work item, head) list.  What the measurements of the earlier kernels said
// (profiles/r01_attn_fwd_design_notes.md) and how this kernel answers:
//   * per 128x128 score tile the tensor pipe needs 1024 clk (QK^T 512 + PV 512; the SS-mode QK^T is also exactly at the
//     128 B/clk shared-memory limit, so narrower key tiles are slower) and the SFU needs 1024 clk (16 ex2/clk/SM).  With
//     one S buffer the chain  S -> exp -> P -> PV -> next S  serialises them (34 % tensor activity).  Here S is
//     TRIPLE-buffered in TMEM: the issuer runs   wait P(j);  O += P(j).V(j);  S[j%3] = Q.K(j+3)^T   so S(j+2) is
//     complete long before a softmax warpgroup asks for it.
//   * two softmax warpgroups take ALTERNATE kv tiles (thread = query row, all 128 key columns, no cross-thread max/sum
//     exchange inside a tile).  While one warpgroup is in its SFU-bound exp phase the other does its TMEM load + max,
//     so the SFU stays busy.  The only coupling between consecutive tiles is the running row maximum, handed over
//     through shared memory right after the (short) max phase; O is rescaled lazily (FlashAttention-4 rule: only when
//     the maximum grew by more than 2^8), each warpgroup keeps its own partial row sum.
//   * prologue (TMEM alloc, barrier init, tensor-map fetch, first loads) and epilogue cost ~7 k clk per CTA when a CTA
//     handles one item.  The CTA is persistent: TMEM and barriers are set up once, the TMA warp prefetches the next
//     item's Q/K/V as ring slots free up, and the next item's first three QK^T run under the current item's epilogue.
//
// CTA = 384 threads: warps 0-3 softmax warpgroup 0 (even kv tiles), warps 4-7 warpgroup 1 (odd kv tiles), warp 8 TMA,
// warp 9 tcgen05 issuer (+ TMEM alloc); warps 10-11 idle (they complete the producer warpgroup for setmaxnreg: the softmax
// warpgroups run with 216 registers per thread, the producer warpgroup with 64).  TMEM (512 columns): S buffers [0,128) [128,256) [256,384), O [384,384+D).
// P (bf16, 64 columns) is written over the start of its own S buffer and is the TMEM A operand of O += P.V.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"

namespace lb {
namespace fs {

constexpr int BM = 128, BN = 128;
constexpr int KST = 3, VST = 2;                        // K / V ring depth
constexpr int WARP_TMA = 8, WARP_MMA = 9, THREADS = 384;   // warps 10, 11 only complete the third warpgroup (setmaxnreg)
constexpr int REGS_SOFTMAX = 216, REGS_PRODUCER = 64;       // the CTA pool is what the launch allocated: 3 x 168 = 504 >= 2 x 216 + 64
constexpr float LOG2E = 1.4426950408889634f;

struct Params {
    const uint8_t* qflag;        // [B*T] or null
    const int32_t* work;         // [n_work][4] = {b, q_tile, variant, -}
    const int32_t* kv_start;     // [B] or null
    const int32_t* kv_end;       // [B] or null
    const int32_t* out_row;      // [B*T] or null
    __nv_bfloat16* O;
    float* lse;                  // [B,H,T]
    int batch, seqlen, heads;
    int n_work, head_group, n_items, n_cta;
    const int32_t* plan_items;   // [n_items] list positions grouped by CTA (host-side balanced split), or null: snake split
    const int32_t* plan_off;     // [gridDim.x + 1]
    float scale;
    long long* cta_log;          // optional [gridDim.x][8]: smid, items, tiles, clock64 at entry / first Q landed / exit
    long long* trace;            // optional [64][8] clock64 stamps of CTA 0, one row per kv tile (global index)
};

#define FS_TRACE(slot, G)                                                                              \
    do {                                                                                               \
        if (p.trace && blockIdx.x == 0 && (G) < 64) p.trace[(G) * 8 + (slot)] = clock64();             \
    } while (0)

template <int D>
struct Smem {
    static constexpr int TILE = 128 * D * 2;                          // one Q / K / V tile
    static constexpr int Q_OFF = 0, K_OFF = TILE, V_OFF = K_OFF + KST * TILE;
    static constexpr int STAT_OFF = V_OFF + VST * TILE;               // float m_sh[2][128], l_sh[2 item parity][2 wg][128]
    static constexpr int BAR_OFF = STAT_OFF + (2 + 4) * 128 * 4;
    static constexpr int NEEDED = BAR_OFF + 512 + 1024;
    static constexpr int TOTAL = NEEDED > 120 * 1024 ? NEEDED : 120 * 1024;     // > half an SM: one CTA per SM (512 TMEM columns)
};

enum {
    B_QFULL = 0,
    B_QEMPTY,
    B_KFULL,
    B_KEMPTY = B_KFULL + KST,
    B_VFULL = B_KEMPTY + KST,
    B_VEMPTY = B_VFULL + VST,
    B_SFULL = B_VEMPTY + VST,        // [3] scores of a tile landed in S buffer
    B_PFULL = B_SFULL + 3,           // [3] probabilities written (128 arrivals: one warpgroup)
    B_OREADY = B_PFULL + 3,          // [2] PV of a tile done, by parity of the global PV index
    B_OFINAL = B_OREADY + 2,         // all MMAs of an item done
    B_OFREE,                         // the epilogue has read O (256 arrivals)
    B_MPUB,                          // [2] running max of a tile published, by tile parity (128 arrivals)
    B_COUNT = B_MPUB + 2
};

// the items of this CTA: round k takes list position k*G + c, alternating direction (the list is sorted heaviest
// first inside a head group, so the snake keeps the per-CTA sums close)
__device__ __forceinline__ int item_of_round(const Params& p, int k) {
    if (p.plan_items) {
        const int i = p.plan_off[blockIdx.x] + k;
        return i < p.plan_off[blockIdx.x + 1] ? p.plan_items[i] : -1;
    }
    const int n_items = p.n_items;
    const int G = (int)gridDim.x, c = (int)blockIdx.x;
    if (k * G >= n_items) return -1;
    const int L = k * G + ((k & 1) ? G - 1 - c : c);
    return L < n_items ? L : -1;
}

struct Item {
    int b, q_tile, variant, h, kvs, kve, first_tile, n_tiles;
};

template <bool CAUSAL>
__device__ __forceinline__ Item decode_item(const Params& p, int L) {
    Item it;
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
    int last = (it.kve + BN - 1) / BN;                               // exclusive
    if (CAUSAL && last > it.q_tile + 1) last = it.q_tile + 1;
    it.n_tiles = last > it.first_tile ? last - it.first_tile : 0;
    return it;
}

// scores of one row (128 columns at TMEM `ts`) -> registers; returns the row maximum.
// MASK: keys outside [kvs, min(kve-1, qi)] become -inf.  The test is classified per 32-column chunk with warp votes: a
// chunk in which every lane sees all 32 keys needs no work, a chunk no lane sees is set to -inf wholesale, only the
// rest is tested element-wise -- on the causal diagonal tile that is one chunk of four per warp.
template <bool MASK, bool CAUSAL>
__device__ __forceinline__ float load_max(uint32_t ts, uint32_t (&v)[128], int kv0, int qi, int kvs, int kve) {
    tmem_ld32(ts, v);
    tmem_ld32(ts + 32, v + 32);
    tmem_ld32(ts + 64, v + 64);
    tmem_ld32(ts + 96, v + 96);
    tc_wait_ld();
    if (MASK) {
        const int hi_key = CAUSAL ? min(qi, kve - 1) : kve - 1;      // last visible key of this row
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int lo = kvs - (kv0 + 32 * c), hi = hi_key - (kv0 + 32 * c);      // visible columns of the chunk: [lo, hi]
            const bool full = lo <= 0 && hi >= 31, none = hi < 0 || lo > 31 || hi < lo;
            if (__all_sync(0xffffffffu, full)) continue;
            if (__all_sync(0xffffffffu, none)) {
#pragma unroll
                for (int e = 0; e < 32; ++e) v[32 * c + e] = 0xff800000u;
                continue;
            }
#pragma unroll
            for (int e = 0; e < 32; ++e) v[32 * c + e] = (e >= lo && e <= hi) ? v[32 * c + e] : 0xff800000u;
        }
    }
    float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F, mx2 = -CUDART_INF_F, mx3 = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 128; j += 8) {                               // pairs of maxima: one 3-input FMNMX each
        mx0 = fmaxf(mx0, fmaxf(__uint_as_float(v[j]), __uint_as_float(v[j + 1])));
        mx1 = fmaxf(mx1, fmaxf(__uint_as_float(v[j + 2]), __uint_as_float(v[j + 3])));
        mx2 = fmaxf(mx2, fmaxf(__uint_as_float(v[j + 4]), __uint_as_float(v[j + 5])));
        mx3 = fmaxf(mx3, fmaxf(__uint_as_float(v[j + 6]), __uint_as_float(v[j + 7])));
    }
    return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
}

// P = 2^(S*sl2 - m_off), packed to bf16 over the first 64 columns of the S buffer; one exponential in every POLY runs
// on the FMA pipes (poly_ex2), 0 = all on the SFU.  Returns the row sum.
template <int POLY>
__device__ __forceinline__ float exp_store(uint32_t ts, uint32_t (&v)[128], float sl2, float m_off) {
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int j = c * 32; j < c * 32 + 32; j += 2) {
            const float x0 = fmaf(__uint_as_float(v[j]), sl2, -m_off), x1 = fmaf(__uint_as_float(v[j + 1]), sl2, -m_off);
            const float p0 = (POLY > 0 && (j % POLY) == 0) ? poly_ex2(x0) : fast_ex2(x0);
            const float p1 = (POLY > 0 && ((j + 1) % POLY) == 0) ? poly_ex2(x1) : fast_ex2(x1);
            l0 += p0;
            l1 += p1;
            v[j >> 1] = pack_bf16(p0, p1);
        }
        tmem_st16(ts + c * 16, v + c * 16);
    }
    return l0 + l1;
}

template <int D, bool CAUSAL, int POLY>
__global__ void __launch_bounds__(THREADS, 1)
attn_fwd_stream_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
                       const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
                       const __grid_constant__ CUtensorMap tmV1, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using S = Smem<D>;
    float* m_sh = reinterpret_cast<float*>(smem + S::STAT_OFF);       // [2][128]
    float* l_sh = m_sh + 256;                                         // [2][2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    const int warp = threadIdx.x >> 5;
    const long long t_entry = p.cta_log ? clock64() : 0;
    const int T = p.seqlen;
    constexpr uint32_t TMEM_COLS = 512, COL_O = 384;

    if (threadIdx.x == 0) {
        for (int i = 0; i < B_COUNT; ++i) {
            const bool wg = (i >= B_PFULL && i < B_PFULL + 3) || i >= B_MPUB;
            mbar_init(bars + i, wg ? 128 : (i == B_OFREE ? 256 : 1));
        }
        fence_barrier_init();
    }
    if (warp == WARP_TMA && elect_one()) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK0);
        tma_prefetch_desc(&tmV0);
        tma_prefetch_desc(&tmK1);
        tma_prefetch_desc(&tmV1);
    }
    if (warp == WARP_MMA) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    // register budget: the softmax warpgroups hold 128 scores per thread, the producer warpgroup needs almost nothing.
    // (setmaxnreg sits at the top of each role's own branch so that ptxas budgets the branch with it.)
    if (warp >= 8) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
    if (warp == WARP_TMA) {
        // ------------------------------------------------------------ TMA producer (runs ahead across items)
        if (elect_one()) {
            uint32_t kl = 0, vl = 0, ic = 0;                          // K / V tiles loaded, items started
            for (int k = 0;; ++k) {
                const int L = item_of_round(p, k);
                if (L < 0) break;
                const Item it = decode_item<CAUSAL>(p, L);
                const int n = it.n_tiles;
                if (n > 0) {
                    const CUtensorMap* tK = it.variant ? &tmK1 : &tmK0;
                    const CUtensorMap* tV = it.variant ? &tmV1 : &tmV0;
                    auto load_k = [&](int j) {
                        const uint32_t s = kl % KST;
                        mbar_wait(bars + B_KEMPTY + s, ((kl / KST) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(bars + B_KFULL + s, S::TILE);
#pragma unroll
                        for (int c = 0; c < D / 64; ++c)
                            tma_load_2d(smem + S::K_OFF + s * S::TILE + c * (BN * 128), tK, bars + B_KFULL + s, it.h * D + c * 64,
                                        it.b * T + (it.first_tile + j) * BN);
                        ++kl;
                    };
                    mbar_wait(bars + B_QEMPTY, (ic & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bars + B_QFULL, S::TILE);
#pragma unroll
                    for (int c = 0; c < D / 64; ++c)
                        tma_load_2d(smem + S::Q_OFF + c * (BM * 128), &tmQ, bars + B_QFULL, it.h * D + c * 64, it.b * T + it.q_tile * BM);
                    for (int j = 0; j < 3 && j < n; ++j) load_k(j);
                    for (int j = 0; j < n; ++j) {
                        const uint32_t s = vl % VST;
                        mbar_wait(bars + B_VEMPTY + s, ((vl / VST) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(bars + B_VFULL + s, S::TILE);
#pragma unroll
                        for (int c = 0; c < D / 64; ++c)
                            tma_load_2d(smem + S::V_OFF + s * S::TILE + c * (BN * 128), tV, bars + B_VFULL + s, it.h * D + c * 64,
                                        it.b * T + (it.first_tile + j) * BN);
                        ++vl;
                        if (j + 3 < n) load_k(j + 3);
                    }
                    ++ic;                                            // items with n == 0 use no Q slot
                }
            }
        }
    } else if (warp == WARP_MMA) {
        // ------------------------------------------------------------ tcgen05 issuer: three score tiles ahead of the softmax
        if (elect_one()) {
            constexpr uint32_t idesc_qk = make_idesc_bf16(BM, BN, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(BM, D, 0, 1);
            const uint32_t dQ = desc_lo_kmajor(smem_u32(smem + S::Q_OFF));
            const uint32_t dK0 = desc_lo_kmajor(smem_u32(smem + S::K_OFF));
            const uint32_t dV0 = desc_lo_mnmajor(smem_u32(smem + S::V_OFF), BN * 128);
            uint32_t g = 0, kc = 0, vc = 0, ic = 0, iq = 0;          // tiles, K / V tiles consumed, items, items with tiles
            bool logged = false;
            auto issue_qk = [&](uint32_t G, bool last_of_item) {      // S[G%3] = Q . K^T
                const uint32_t s = kc % KST;
                mbar_wait(bars + B_KFULL + s, (kc / KST) & 1u);
                tc_fence_after_sync();
                const uint32_t dK = dK0 + s * (uint32_t)(S::TILE >> 4);
                const uint32_t d_s = tmem_base + (G % 3) * 128;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t off = ((uint32_t)(kk / 4) * (BM * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                    umma_ss_lo(d_s, dQ + off, dK + off, idesc_qk, kk ? 1u : 0u);
                }
                tc_commit(bars + B_SFULL + G % 3);
                tc_commit(bars + B_KEMPTY + s);
                if (last_of_item) tc_commit(bars + B_QEMPTY);        // Q may be replaced by the next item's
                ++kc;
            };
            for (int k = 0;; ++k) {
                const int L = item_of_round(p, k);
                if (L < 0) break;
                const Item it = decode_item<CAUSAL>(p, L);
                const int n = it.n_tiles;
                if (n > 0) {
                    mbar_wait(bars + B_QFULL, iq & 1u);
                    if (p.cta_log && !logged) {
                        p.cta_log[(int64_t)blockIdx.x * 8 + 4] = clock64();
                        logged = true;
                    }
                    for (int j = 0; j < 3 && j < n; ++j) issue_qk(g + j, j == n - 1);      // overlaps the previous epilogue
                    for (int j = 0; j < n; ++j) {
                        const uint32_t G = g + j;
                        FS_TRACE(0, G);                                  // MMA: start waiting for P
                        mbar_wait(bars + B_PFULL + G % 3, (G / 3) & 1u);
                        FS_TRACE(1, G);                                  // MMA: P seen
                        if (j == 0 && ic > 0) mbar_wait(bars + B_OFREE, (ic - 1) & 1u);    // previous epilogue has read O
                        const uint32_t s = vc % VST;
                        mbar_wait(bars + B_VFULL + s, (vc / VST) & 1u);
                        tc_fence_after_sync();
                        const uint32_t dV = dV0 + s * (uint32_t)(S::TILE >> 4);
                        const uint32_t a_p = tmem_base + (G % 3) * 128;
#pragma unroll
                        for (int kk = 0; kk < BN / 16; ++kk) {
                            // A = P in TMEM: keys 16kk.. at column 8kk of the S buffer; B = V as MN-major (16 key rows = 2048 B)
                            umma_ts_lo(tmem_base + COL_O, a_p + (uint32_t)kk * 8, dV + (uint32_t)kk * (2048 >> 4), idesc_pv,
                                       (j | kk) ? 1u : 0u);
                        }
                        tc_commit(bars + B_OREADY + (G & 1u));
                        tc_commit(bars + B_VEMPTY + s);
                        ++vc;
                        if (j + 3 < n) issue_qk(G + 3, j + 3 == n - 1);
                        FS_TRACE(2, G);                                  // MMA: PV + QK(+3) issued
                    }
                    ++iq;
                } else if (ic > 0) {
                    mbar_wait(bars + B_OFREE, (ic - 1) & 1u);        // keep the per-item waits consecutive
                }
                tc_commit(bars + B_OFINAL);
                g += (uint32_t)n;
                ++ic;
            }
        }
    }
    } else {
        // ------------------------------------------------------------ softmax warpgroups: alternate kv tiles, shared running max
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOFTMAX));
        const int w = warp >> 2;                                  // warpgroup = parity of the kv tiles it takes
        const int r = (warp & 3) * 32 + (threadIdx.x & 31);       // query row in tile == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        constexpr int DH = D / 2;
        const float sl2 = p.scale * LOG2E;
        uint32_t g = 0, ic = 0, mpc0 = 0, mpc1 = 0;               // tiles / items so far, completions of MPUB[0], MPUB[1]
        for (int k = 0;; ++k) {
            const int L = item_of_round(p, k);
            if (L < 0) break;
            const Item it = decode_item<CAUSAL>(p, L);
            const int n = it.n_tiles;
            const int q0 = it.q_tile * BM, qi = q0 + r;
            float m_ref = -CUDART_INF_F, l = 0.f;                 // l is relative to m_ref
            // destination of my row: loaded now, used by the epilogue (keeps two dependent global loads off its path)
            const int64_t bt = (int64_t)it.b * T + qi;
            const bool row_ok = (qi < T) && (!p.qflag || (int)p.qflag[qi < T ? bt : 0] == it.variant);
            const int64_t dst = row_ok ? (p.out_row ? (int64_t)p.out_row[bt] : bt) : 0;
            for (int j = w; j < n; j += 2) {
                const uint32_t G = g + (uint32_t)j;
                const uint32_t colS = (G % 3) * 128;
                const int kv0 = (it.first_tile + j) * BN;
                const bool need_mask = (CAUSAL && kv0 + BN - 1 > q0) || (kv0 + BN > it.kve) || (kv0 < it.kvs);
                if ((threadIdx.x & 127) == 0) FS_TRACE(3, G);         // softmax: start waiting for S
                mbar_wait(bars + B_SFULL + G % 3, (G / 3) & 1u);
                tc_fence_after_sync();
                if ((threadIdx.x & 127) == 0) FS_TRACE(4, G);         // softmax: S seen
                uint32_t sv[128];
                const float mx = need_mask ? load_max<true, CAUSAL>(lane_addr + colS, sv, kv0, qi, it.kvs, it.kve)
                                           : load_max<false, CAUSAL>(lane_addr + colS, sv, kv0, qi, it.kvs, it.kve);
                // ---- running max of the row after tile j-1 (the other warpgroup's tile)
                float m_cur;
                if ((threadIdx.x & 127) == 0) FS_TRACE(5, G);         // softmax: scores loaded, max done
                if (j == 0) {
                    m_cur = mx;
                } else {
                    const int pb = (j - 1) & 1;
                    mbar_wait(bars + B_MPUB + pb, ((pb ? mpc1 : mpc0) + (uint32_t)((j - 1) >> 1)) & 1u);
                    const float m_prev = m_sh[pb * 128 + r];
                    const float m_new = fmaxf(m_prev, mx);
                    // lazy correction: rescale O only when the running max moved by more than 2^8
                    const bool grow = (m_new - m_prev) * sl2 > 8.f;       // also true when m_prev == -inf and m_new finite
                    m_cur = grow ? m_new : m_prev;
                    if (__any_sync(0xffffffffu, grow)) {
                        mbar_wait(bars + B_OREADY + ((G - 1) & 1u), ((G - 1) >> 1) & 1u);   // PV of tile j-1 has landed in O
                        tc_fence_after_sync();
                        const float alpha = grow ? ((m_prev == -CUDART_INF_F) ? 0.f : fast_ex2((m_prev - m_new) * sl2)) : 1.f;
#pragma unroll 1
                        for (int c = 0; c < D / 8; ++c) {            // small chunks: the scores stay live in registers
                            uint32_t v[8];
                            tmem_ld8(lane_addr + COL_O + c * 8, v);
                            tc_wait_ld();
#pragma unroll
                            for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
                            tmem_st8(lane_addr + COL_O + c * 8, v);
                        }
                        tc_wait_st();
                    }
                }
                m_sh[(j & 1) * 128 + r] = m_cur;
                mbar_arrive(bars + B_MPUB + (j & 1));
                if ((threadIdx.x & 127) == 0) FS_TRACE(6, G);         // softmax: running max published
                // ---- my partial row sum follows the reference maximum
                if (m_ref != m_cur) {
                    l = (m_ref == -CUDART_INF_F) ? 0.f : l * fast_ex2((m_ref - m_cur) * sl2);
                    m_ref = m_cur;
                }
                const float m_off = (m_cur == -CUDART_INF_F) ? 0.f : m_cur * sl2;
                l += exp_store<POLY>(lane_addr + colS, sv, sl2, m_off);
                tc_wait_st();
                tc_fence_before_sync();
                mbar_arrive(bars + B_PFULL + G % 3);
                if ((threadIdx.x & 127) == 0) FS_TRACE(7, G);         // softmax: P stored, arrived
            }
            // ---- end of the item: final maximum, combine the two partial sums
            float m_fin = -CUDART_INF_F;
            if (n > 0) {
                const int pb = (n - 1) & 1;
                mbar_wait(bars + B_MPUB + pb, ((pb ? mpc1 : mpc0) + (uint32_t)((n - 1) >> 1)) & 1u);
                m_fin = m_sh[pb * 128 + r];
                if (m_ref != m_fin) l = (m_ref == -CUDART_INF_F) ? 0.f : l * fast_ex2((m_ref - m_fin) * sl2);
            }
            float* lbuf = l_sh + (ic & 1u) * 256;
            lbuf[w * 128 + r] = l;
            named_bar_sync(1, 256);
            const float l_tot = lbuf[r] + lbuf[128 + r];
            mbar_wait(bars + B_OFINAL, ic & 1u);
            tc_fence_after_sync();
            const float inv_l = l_tot > 0.f ? 1.f / l_tot : 0.f;
            __nv_bfloat16* orow = p.O + dst * ((int64_t)p.heads * D) + (int64_t)it.h * D + w * DH;
#pragma unroll 1
            for (int c = 0; c < DH / 32; ++c) {
                uint32_t v[32];
                if (n > 0) {
                    tmem_ld32(lane_addr + COL_O + w * DH + c * 32, v);
                    tc_wait_ld();
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) v[e] = 0u;
                }
                if (row_ok) {
#pragma unroll
                    for (int e = 0; e < 32; e += 8) {
                        uint4 o;
                        o.x = pack_bf16(__uint_as_float(v[e + 0]) * inv_l, __uint_as_float(v[e + 1]) * inv_l);
                        o.y = pack_bf16(__uint_as_float(v[e + 2]) * inv_l, __uint_as_float(v[e + 3]) * inv_l);
                        o.z = pack_bf16(__uint_as_float(v[e + 4]) * inv_l, __uint_as_float(v[e + 5]) * inv_l);
                        o.w = pack_bf16(__uint_as_float(v[e + 6]) * inv_l, __uint_as_float(v[e + 7]) * inv_l);
                        *reinterpret_cast<uint4*>(orow + c * 32 + e) = o;
                    }
                }
                __syncwarp();
            }
            tc_fence_before_sync();
            mbar_arrive(bars + B_OFREE);                          // O may be overwritten by the next item
            if (row_ok && p.lse && w == 0) {
                // natural-log LSE of the scaled scores; +inf marks a row with no visible key (P == 0 in backward)
                p.lse[((int64_t)it.b * p.heads + it.h) * T + qi] = l_tot > 0.f ? (m_fin * p.scale + __logf(l_tot)) : CUDART_INF_F;
            }
            g += (uint32_t)n;
            mpc0 += (uint32_t)((n + 1) >> 1);
            mpc1 += (uint32_t)(n >> 1);
            ++ic;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == WARP_MMA) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    if (p.cta_log && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        int items = 0, tiles = 0;
        for (int k = 0;; ++k) {
            const int L = item_of_round(p, k);
            if (L < 0) break;
            ++items;
            tiles += decode_item<CAUSAL>(p, L).n_tiles;
        }
        long long* e = p.cta_log + (int64_t)blockIdx.x * 8;
        e[0] = smid; e[1] = items; e[2] = tiles; e[3] = t_entry; e[5] = clock64();
    }
}

template <int D, bool CAUSAL, int POLY>
static int launch_p(const CUtensorMap* tm, const Params& p, cudaStream_t st) {
    using S = Smem<D>;
    auto kern = attn_fwd_stream_kernel<D, CAUSAL, POLY>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail(LB_ELAUNCH, "attn_fwd_stream: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    kern<<<(unsigned)p.n_cta, THREADS, S::TOTAL, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], p);
    return check_launch("attn_fwd_stream");
}

// fraction of exponentials evaluated on the FMA pipes: 1/POLY (LB_STREAM_EXP_POLY=0|2|3|4 for experiments)
static int poly_mod() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LB_STREAM_EXP_POLY");
        v = e ? atoi(e) : 0;
        if (v != 0 && v != 2 && v != 3 && v != 4) v = 0;
    }
    return v;
}

template <int D, bool CAUSAL>
static int launch(const CUtensorMap* tm, const Params& p, cudaStream_t st) {
    switch (poly_mod()) {
        case 2: return launch_p<D, CAUSAL, 2>(tm, p, st);
        case 3: return launch_p<D, CAUSAL, 3>(tm, p, st);
        case 4: return launch_p<D, CAUSAL, 4>(tm, p, st);
        default: return launch_p<D, CAUSAL, 0>(tm, p, st);
    }
}

}  // namespace fs
}  // namespace lb

using namespace lb;

static long long* g_stream_cta_log = nullptr;
static long long* g_stream_trace = nullptr;

/* diagnostics: CTA 0 of subsequent lb_attn_fwd_stream launches writes clock64 stamps into `buf` ([64][8] int64, device; one
 * row per kv tile: MMA wait-P / P seen / issued, softmax wait-S / S seen / max done / max published / P arrived). NULL = off */
extern "C" int lb_attn_fwd_stream_set_trace(void* buf) {
    g_stream_trace = (long long*)buf;
    return LB_OK;
}

/* diagnostics: every CTA of subsequent lb_attn_fwd_stream launches logs {smid, items, tiles, clock64 at entry, first Q
 * landed, exit, -, -} into `buf` ([number of SMs][8] int64, device); NULL = off */
extern "C" int lb_attn_fwd_stream_set_cta_log(void* buf) {
    g_stream_cta_log = (long long*)buf;
    return LB_OK;
}

extern "C" int lb_attn_fwd_stream(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1,
                                  const uint8_t* qflag, const int32_t* work, int n_work, const int32_t* plan_items,
                                  const int32_t* plan_off, int n_cta, int head_group, const int32_t* kv_start,
                                  const int32_t* kv_end, const int32_t* out_row, void* O, float* lse, int batch, int seqlen,
                                  int heads, int head_dim, int causal, float scale, void* stream) {
    LB_REQUIRE(batch > 0 && seqlen > 0 && heads > 0 && n_work >= 0, LB_EINVAL, "attn_fwd_stream: bad shape");
    LB_REQUIRE(head_dim == 64 || head_dim == 128, LB_EINVAL, "attn_fwd_stream: head_dim %d (64 or 128 supported)", head_dim);
    LB_REQUIRE(Q && K0 && V0 && O && work, LB_EINVAL, "attn_fwd_stream: null argument");
    LB_REQUIRE((plan_items == nullptr) == (plan_off == nullptr), LB_EINVAL, "attn_fwd_stream: plan_items and plan_off go together");
    LB_REQUIRE(!plan_items || n_cta > 0, LB_EINVAL, "attn_fwd_stream: n_cta %d with a plan", n_cta);
    LB_REQUIRE(((uintptr_t)O & 15) == 0, LB_EALIGN, "attn_fwd_stream: O must be 16-byte aligned");
    if (n_work == 0) return LB_OK;
    int rc = require_sm100();
    if (rc) return rc;
    const uint64_t rows = (uint64_t)batch * seqlen, cols = (uint64_t)heads * head_dim;
    CUtensorMap tm[5];
    const void* ptrs[5] = {Q, K0, V0, K1 ? K1 : K0, V1 ? V1 : V0};
    for (int i = 0; i < 5; ++i) {
        rc = make_tmap_bf16_2d(&tm[i], ptrs[i], rows, cols, cols, 128, 64);
        if (rc) return rc;
    }
    fs::Params p;
    p.qflag = qflag; p.work = work; p.kv_start = kv_start; p.kv_end = kv_end; p.out_row = out_row;
    p.O = (__nv_bfloat16*)O; p.lse = lse; p.batch = batch; p.seqlen = seqlen; p.heads = heads; p.scale = scale;
    p.n_work = n_work; p.head_group = head_group > 0 ? head_group : attn_head_group(); p.n_items = n_work * heads;
    p.plan_items = plan_items; p.plan_off = plan_off;
    if (plan_items) {
        p.n_cta = n_cta;
    } else {
        const int sms = sm_count();
        if (sms <= 0) return fail(LB_ELAUNCH, "attn_fwd_stream: no SM count");
        p.n_cta = p.n_items < sms ? p.n_items : sms;
    }
    p.cta_log = g_stream_cta_log;
    p.trace = g_stream_trace;
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 128) return causal ? fs::launch<128, true>(tm, p, st) : fs::launch<128, false>(tm, p, st);
    return causal ? fs::launch<64, true>(tm, p, st) : fs::launch<64, false>(tm, p, st);
}

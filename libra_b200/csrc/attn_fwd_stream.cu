// Bridge attention forward, persistent streaming kernel (A10; same maths and operand formulation as attn_fwd.cu).
//
// One CTA per SM walks a static share of the (work item, head) list as ONE stream of 128x128 score tiles (global tile
// index G), independent of where the items begin and end.  What the measurements of the earlier kernels said
// (profiles/r01_attn_fwd_stream_notes.md) and how this kernel answers:
//   * per score tile the tensor pipe needs 1024 clk (QK^T 512 + PV 512) and the SFU needs 1024 clk (16 ex2/clk/SM), so
//     any dependency loop between the two shows up directly as idle pipe.  With P written over its own S buffer the loop
//     P(G) -> PV(G) -> QK^T(G+3) -> softmax(G+3) was the limit (1650 clk per tile in steady state).  Here S is
//     double-buffered and P has its OWN two TMEM buffers: an S buffer is free as soon as the softmax warpgroup has pulled
//     the scores into registers, so QK^T(G+2) is issued at the START of softmax(G) and the scores of a tile are always
//     complete long before a warpgroup asks for them.  QK^T and PV are issued by two different threads: a single issuing
//     thread executes one dependent instruction every ~5 clk, and its waits / descriptor arithmetic / commits between two
//     batches of MMAs were the critical path (~2100 clk per tile for both products from one general loop).
//   * two softmax warpgroups take alternate tiles of the stream (thread = query row, all 128 key columns, no cross-thread
//     max/sum exchange inside a tile).  The only coupling between consecutive tiles is the running row maximum, handed
//     over through shared memory right after the (short) max phase; O is rescaled lazily (FlashAttention-4 rule: only
//     when the maximum grew by more than 2^8), each warpgroup keeps its own partial row sum.
//   * an item switch cost ~6.5 k clk (4 tiles' worth, 30 % of the kernel at 15 items per CTA) when both softmax
//     warpgroups met at a barrier for a joint, chunked epilogue and every role decoded the next item from global memory
//     (a chain of dependent loads).  Now (a) the CTA's item list is decoded once into shared memory; (b) the warpgroup
//     that took the item's LAST tile writes the item out alone, and it releases O first: all D accumulator columns go
//     TMEM -> registers in one go (the score registers are free between tiles), O is handed back to the PV thread
//     ~200 clk after the last PV retired, and only then is the row normalised, packed and stored; the other warpgroup
//     has deposited its (reference max, partial sum) pair in shared memory and is already on the next item; (c) the
//     loaders and the tcgen05 threads run across item boundaries.  (A separate epilogue warpgroup was tried: with the
//     56 registers the budget leaves it, its chunked TMEM -> staging -> global loop held O for ~3 k clk, longer than the
//     two tiles the P double buffer lets the softmax run ahead.)
//   * V is needed right after P(G) is written (PV is on the critical path P -> PV -> P buffer free), and a TMA load
//     takes ~1.5-2 k clk here: the V ring is 3 deep (load issued when PV(G-3) retired); K is consumed two tiles ahead
//     of the softmax anyway and shares the S double buffer's phase (K slot == S slot).
//
// CTA = 384 threads: warps 0-3 softmax warpgroup 0 (even tiles of the stream), warps 4-7 warpgroup 1 (odd tiles),
// warp 8 TMA for Q and K, warp 9 tcgen05 issuer for QK^T (+ TMEM alloc), warp 10 TMA for V, warp 11 tcgen05 issuer for
// PV.  setmaxnreg: softmax warpgroups 216 registers per thread, the producer warpgroup 64 (2 x 216 + 64 <= 3 x 168).
// TMEM (512 columns): S0 [0,128) S1 [128,256) P0 [256,320) P1 [320,384) O [384,384+D).  P (bf16, 64 columns) is the
// TMEM A operand of O += P.V.
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"

namespace lb {
namespace fs {

constexpr int BM = 128, BN = 128;
constexpr int KST = 2, VST = 3;                                  // K / V ring depth; K slot == S slot (G & 1)
constexpr int WARP_KLOAD = 8, WARP_QK = 9, WARP_VLOAD = 10, WARP_PV = 11, THREADS = 384;
constexpr int REGS_SOFTMAX = 216, REGS_PRODUCER = 64;            // the CTA pool is what the launch allocated: 3 x 168 = 504 >= 2 x 216 + 64
constexpr int MAX_ITEMS = 128;                                   // decoded items per CTA held in shared memory
constexpr float LOG2E = 1.4426950408889634f;
constexpr uint32_t COL_P = 256, COL_O = 384, TMEM_COLS = 512;

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
    long long* trace;            // optional [64][32] clock64 stamps of CTA 0, one row per kv tile (global index)
};

#define FS_TRACE(slot, G)                                                                              \
    do {                                                                                               \
        if (TRACE && blockIdx.x == 0 && (G) < 64) p.trace[(G) * 32 + (slot)] = clock64();              \
    } while (0)

struct __align__(16) Item {
    int b, q_tile, variant, h, kvs, kve, first_tile, n_tiles;       // b < 0 (and n_tiles < 0): end of the CTA's list
};

template <int D>
struct Smem {
    static constexpr int TILE = 128 * D * 2;                          // one Q / K / V tile
    static constexpr int Q_OFF = 0, K_OFF = TILE, V_OFF = K_OFF + KST * TILE;
    static constexpr int ITEM_OFF = V_OFF + VST * TILE;               // Item[MAX_ITEMS + 1]
    static constexpr int M_OFF = ITEM_OFF + (MAX_ITEMS + 1) * 32;     // float m_sh[2 tile parity][128]
    static constexpr int DEP_OFF = M_OFF + 2 * 128 * 4;               // float lm_sh[2: m_ref, l][128] of the depositing warpgroup
    static constexpr int STAGE_OFF = DEP_OFF + 2 * 128 * 4;           // epilogue staging: 2 wg x 4 warps x 32 rows x 64 B
    static constexpr int BAR_OFF = STAGE_OFF + 8 * 32 * 64;
    static constexpr int NEEDED = BAR_OFF + 512 + 1024;
    static_assert(NEEDED <= 227 * 1024, "shared memory budget");
    static constexpr int TOTAL = NEEDED > 120 * 1024 ? NEEDED : 120 * 1024;     // > half an SM: one CTA per SM (512 TMEM columns)
};

enum {
    B_QFULL = 0,                     // Q of an item landed
    B_QEMPTY,                        // last QK^T of the item retired
    B_KFULL,                         // [2] K(G) landed in slot G & 1
    B_VFULL = B_KFULL + KST,         // [3] V(G) landed in slot G % 3
    B_VEMPTY = B_VFULL + VST,        // [3] PV(G) retired: V slot free
    B_SFULL = B_VEMPTY + VST,        // [2] scores of tile G landed in S[G & 1]  (== K slot G & 1 free again)
    B_SFREE = B_SFULL + 2,           // [2] the softmax warpgroup holds S[G & 1] in registers (128 arrivals)
    B_PFULL = B_SFREE + 2,           // [2] probabilities of tile G written to P[G & 1] (128 arrivals)
    B_PFREE = B_PFULL + 2,           // [2] PV of tile G retired: P[G & 1] reusable, O holds tile G
    B_OFREE = B_PFREE + 2,           // the item's owner holds O in registers (128 arrivals; items with tiles only)
    B_MPUB,                          // [2] running max of tile G published in m_sh[G & 1] (128 arrivals)
    B_LDEP = B_MPUB + 2,             // the other warpgroup's (m_ref, l) of an item deposited in lm_sh (128 arrivals)
    B_LFREE,                         // ... and read by the item's owner (128 arrivals)
    B_COUNT
};

// NOTE on parity waits: mbarrier.try_wait.parity(p) is true whenever the barrier's CURRENT phase has parity != p, so a
// waiter that is two phases ahead of the barrier gets a false positive.  Every wait in this kernel is therefore on a
// barrier whose previous phase the waiter (or something it has already synchronised with) is known to have seen
// complete; where the stream would allow a role to run further ahead (the owner of a one-tile item, the rescale path)
// the wait goes through a barrier that cannot be more than one phase behind (PFREE of the role's own tile).

// list position of the CTA's k-th item, or -1.  Without a plan: round k takes list position k*G + c, alternating
// direction (the list is sorted heaviest first inside a head group, so the snake keeps the per-CTA sums close)
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
    if (L < 0) {                                                     // end marker: b < 0 and n_tiles < 0
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
    int last = (it.kve + BN - 1) / BN;                               // exclusive
    if (CAUSAL && last > it.q_tile + 1) last = it.q_tile + 1;
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

// row maximum of the 128 scores in v.  MASK: keys outside [kvs, min(kve-1, qi)] become -inf first (NEED_LO: the tile starts
// before kvs -- left padding only; without it the test is one compare + select per element).  The test is
// classified per 32-column chunk with a warp vote: a chunk in which every lane sees all 32 keys needs no work, the rest
// is tested element-wise (a third, "set the chunk to -inf wholesale" variant made ptxas spill the score registers).
template <bool MASK, bool CAUSAL, bool NEED_LO>
__device__ __forceinline__ float mask_max(uint32_t (&v)[128], int kv0, int qi, int kvs, int kve) {
    if (MASK) {
        const int hi_key = CAUSAL ? min(qi, kve - 1) : kve - 1;      // last visible key of this row
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            const int lo = kvs - (kv0 + 32 * c), hi = hi_key - (kv0 + 32 * c);      // visible columns of the chunk: [lo, hi]
            const bool full = (!NEED_LO || lo <= 0) && hi >= 31;
            if (__all_sync(0xffffffffu, full)) continue;
#pragma unroll
            for (int e = 0; e < 32; ++e) v[32 * c + e] = ((!NEED_LO || e >= lo) && e <= hi) ? v[32 * c + e] : 0xff800000u;
        }
    }
    // eight independent chains of 3-input maxima (16 deep each would be the critical path of the phase with four)
    float mx[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) mx[c] = fmaxf(__uint_as_float(v[2 * c]), __uint_as_float(v[2 * c + 1]));
#pragma unroll
    for (int j = 16; j < 128; j += 16) {
#pragma unroll
        for (int c = 0; c < 8; ++c) mx[c] = fmaxf(mx[c], fmaxf(__uint_as_float(v[j + 2 * c]), __uint_as_float(v[j + 2 * c + 1])));
    }
    return fmaxf(fmaxf(fmaxf(mx[0], mx[1]), fmaxf(mx[2], mx[3])), fmaxf(fmaxf(mx[4], mx[5]), fmaxf(mx[6], mx[7])));
}

// P = 2^(S*sl2 - m_off), packed to bf16 into the 64 columns at TMEM `tp`; one exponential in every POLY runs on the FMA
// pipes (poly_ex2), 0 = all on the SFU.  Returns the row sum.
template <int POLY>
__device__ __forceinline__ float exp_store(uint32_t tp, uint32_t (&v)[128], float sl2, float m_off) {
    // packed fp32x2 arithmetic: per PAIR of keys one FFMA2 (exponent), two MUFU.EX2, one FADD2 (row sum), one bf16x2 pack
    const uint64_t sl2_2 = f32x2(sl2, sl2), nm_2 = f32x2(-m_off, -m_off);
    uint64_t l2 = f32x2(0.f, 0.f);
#pragma unroll
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int j = c * 32; j < c * 32 + 32; j += 2) {
            float x0, x1;
            f32x2_unpack(fma_f32x2(f32x2(__uint_as_float(v[j]), __uint_as_float(v[j + 1])), sl2_2, nm_2), x0, x1);
            const float p0 = (POLY > 0 && (j % POLY) == 0) ? poly_ex2(x0) : fast_ex2(x0);
            const float p1 = (POLY > 0 && ((j + 1) % POLY) == 0) ? poly_ex2(x1) : fast_ex2(x1);
            l2 = add_f32x2(l2, f32x2(p0, p1));
            v[j >> 1] = pack_bf16(p0, p1);
        }
        tmem_st16(tp + c * 16, v + c * 16);
    }
    float l0, l1;
    f32x2_unpack(l2, l0, l1);
    return l0 + l1;
}

template <int D, bool CAUSAL, int POLY, bool TRACE>
__global__ void __launch_bounds__(THREADS, 1)
attn_fwd_stream_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
                       const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
                       const __grid_constant__ CUtensorMap tmV1, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using S = Smem<D>;
    Item* items = reinterpret_cast<Item*>(smem + S::ITEM_OFF);
    float* m_sh = reinterpret_cast<float*>(smem + S::M_OFF);          // [2][128]
    float* lm_sh = reinterpret_cast<float*>(smem + S::DEP_OFF);       // [2][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + B_COUNT);

    const int warp = threadIdx.x >> 5;
    const long long t_entry = p.cta_log ? clock64() : 0;
    const int T = p.seqlen;

    if (threadIdx.x == 0) {
        for (int i = 0; i < B_COUNT; ++i) {
            const bool wg = (i >= B_SFREE && i < B_PFREE) || i == B_OFREE || i >= B_MPUB;
            mbar_init(bars + i, wg ? 128 : 1);
        }
        fence_barrier_init();
    }
    // the CTA's items, decoded once (every role walks this table; an entry with b < 0 ends it)
    for (int k = threadIdx.x; k <= MAX_ITEMS; k += THREADS) items[k] = decode_item<CAUSAL>(p, item_of_round(p, k));
    if (warp == WARP_KLOAD && elect_one()) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(&tmK0);
        tma_prefetch_desc(&tmK1);
    }
    if (warp == WARP_VLOAD && elect_one()) {
        tma_prefetch_desc(&tmV0);
        tma_prefetch_desc(&tmV1);
    }
    if (warp == WARP_QK) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t bar0 = smem_u32(bars);                             // barrier i lives at bar0 + 8 i

    // register budget: the softmax warpgroups hold 128 scores per thread, the producer warpgroup needs almost nothing.
    // (setmaxnreg sits at the top of each role's own branch so that ptxas budgets the branch with it.)
    if (warp >= 8) {
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS_PRODUCER));
        if (warp == WARP_KLOAD) {
            // ------------------------------------------------------------ TMA producer for Q and K (runs ahead across items)
            if (elect_one()) {
                uint32_t G = 0, iq = 0;                               // K tiles loaded, items with tiles started
                for (int k = 0;; ++k) {
                    const Item it = load_item(items, k);
                    if (it.b < 0) break;
                    if (it.n_tiles == 0) continue;
                    const CUtensorMap* tK = it.variant ? &tmK1 : &tmK0;
                    wait_bar(bar0 + 8 * B_QEMPTY, (iq & 1u) ^ 1u);
                    mbar_arrive_expect_tx(bars + B_QFULL, S::TILE);
#pragma unroll
                    for (int c = 0; c < D / 64; ++c)
                        tma_load_2d(smem + S::Q_OFF + c * (BM * 128), &tmQ, bars + B_QFULL, it.h * D + c * 64, it.b * T + it.q_tile * BM);
                    ++iq;
                    for (int j = 0; j < it.n_tiles; ++j, ++G) {
                        const uint32_t s = G & 1u;                    // free again once QK(G-2) retired: that is SFULL's phase
                        wait_bar(bar0 + 8 * (B_SFULL + s), ((G >> 1) & 1u) ^ 1u);
                        mbar_arrive_expect_tx(bars + B_KFULL + s, S::TILE);
#pragma unroll
                        for (int c = 0; c < D / 64; ++c)
                            tma_load_2d(smem + S::K_OFF + s * S::TILE + c * (BN * 128), tK, bars + B_KFULL + s, it.h * D + c * 64,
                                        it.b * T + (it.first_tile + j) * BN);
                    }
                }
            }
        } else if (warp == WARP_VLOAD) {
            // ------------------------------------------------------------ TMA producer for V
            if (elect_one()) {
                uint32_t vs = 0, vph = 1;                             // ring slot, parity of the "slot empty" phase to wait for
                for (int k = 0;; ++k) {
                    const Item it = load_item(items, k);
                    if (it.b < 0) break;
                    const CUtensorMap* tV = it.variant ? &tmV1 : &tmV0;
                    for (int j = 0; j < it.n_tiles; ++j) {
                        wait_bar(bar0 + 8 * (B_VEMPTY + vs), vph);
                        mbar_arrive_expect_tx(bars + B_VFULL + vs, S::TILE);
#pragma unroll
                        for (int c = 0; c < D / 64; ++c)
                            tma_load_2d(smem + S::V_OFF + vs * S::TILE + c * (BN * 128), tV, bars + B_VFULL + vs, it.h * D + c * 64,
                                        it.b * T + (it.first_tile + j) * BN);
                        if (++vs == VST) {
                            vs = 0;
                            vph ^= 1u;
                        }
                    }
                }
            }
        } else if (warp == WARP_QK) {
            // ------------------------------------------------------------ tcgen05 issuer 1: S = Q . K^T, two tiles ahead of PV.
            // Everything an issuing thread executes between two batches of MMAs is serial latency on the tensor pipe's
            // critical path (a dependent instruction every ~5-10 clk while the softmax warps keep the issue ports busy;
            // measured ~2100 clk per tile with ONE thread running a general two-cursor loop for both products).  So QK^T
            // and PV have their own issuing threads (the hardware orders them through the barriers they already wait on),
            // the loops are unrolled over the slot bit G & 1 -- barrier addresses, TMEM columns and K descriptors become
            // constants -- and the items are reduced to their tile counts.
            if (elect_one()) {
                constexpr uint32_t idesc_qk = make_idesc_bf16(BM, BN, 0, 0);
                constexpr uint32_t TILE16 = (uint32_t)(S::TILE >> 4);
                const uint32_t dQ = desc_lo_kmajor(smem_u32(smem + S::Q_OFF));
                const uint32_t dK0 = desc_lo_kmajor(smem_u32(smem + S::K_OFF));
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
                bool first = true, logged = false;
                uint32_t iq = 0, qph0 = 0, qph1 = 0, Gq = 0;          // items started; phase parity of each slot's next use
                auto do_qk = [&](const uint32_t b, uint32_t& qph, const bool wait_sfree) {     // S[b] = Q . K(Gq)^T, b == Gq & 1
                    FS_TRACE(0, Gq);                                  // QK: start waiting for the S buffer
                    if (wait_sfree) wait_bar(bar0 + 8 * (B_SFREE + b), qph ^ 1u);             // S(Gq-2) is in registers
                    FS_TRACE(1, Gq);                                  // QK: S buffer free
                    if (first) {
                        wait_bar(bar0 + 8 * B_QFULL, iq & 1u);
                        if (p.cta_log && !logged) {
                            p.cta_log[(int64_t)blockIdx.x * 8 + 4] = clock64();
                            logged = true;
                        }
                    }
                    wait_bar(bar0 + 8 * (B_KFULL + b), qph);
                    tc_fence_after_sync();
                    FS_TRACE(2, Gq);                                  // QK: Q, K landed
                    const uint32_t dK = dK0 + b * TILE16;
                    const uint32_t d_s = tmem_base + b * 128;
#pragma unroll
                    for (int kk = 0; kk < D / 16; ++kk) {
                        const uint32_t off = ((uint32_t)(kk / 4) * (BM * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                        umma_ss_lo(d_s, dQ + off, dK + off, idesc_qk, kk ? 1u : 0u);
                    }
                    FS_TRACE(3, Gq);                                  // QK: issued
                    commit_bar(bar0 + 8 * (B_SFULL + b));             // scores landed; K slot b free
                    qph ^= 1u;
                    first = false;
                    if (--left == 0) {                                // Q may be replaced by the next item's
                        commit_bar(bar0 + 8 * B_QEMPTY);
                        ++iq;
                        left = next_count();
                        first = true;
                    }
                    FS_TRACE(4, Gq);                                  // QK: committed
                    ++Gq;
                };
                if (left > 0) do_qk(0u, qph0, false);                 // QK(0), QK(1): both S buffers start out free
                if (left > 0) do_qk(1u, qph1, false);
                while (left > 0) {
                    do_qk(0u, qph0, true);
                    if (left <= 0) break;
                    do_qk(1u, qph1, true);
                }
            }
        } else {
            // ------------------------------------------------------------ tcgen05 issuer 2: O (+)= P . V
            if (elect_one()) {
                constexpr uint32_t idesc_pv = make_idesc_bf16(BM, D, 0, 1);
                constexpr uint32_t TILE16 = (uint32_t)(S::TILE >> 4);
                const uint32_t dV0 = desc_lo_mnmajor(smem_u32(smem + S::V_OFF), BN * 128);
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
                uint32_t ne = 0, pph0 = 0, pph1 = 0, Gp = 0;          // items (with tiles) started
                uint32_t vs = 0, vph = 0;                             // V ring slot and its "full" phase parity
                auto do_pv = [&](const uint32_t b, uint32_t& pph) {   // O (+)= P[b] . V(Gp), b == Gp & 1
                    FS_TRACE(5, Gp);                                  // PV: start waiting for V, O, P (the last to arrive last)
                    wait_bar(bar0 + 8 * (B_VFULL + vs), vph);
                    if (first && ne > 0) wait_bar(bar0 + 8 * B_OFREE, (ne - 1) & 1u);        // the previous item's O is out of TMEM
                    FS_TRACE(6, Gp);                                  // PV: V landed, O free
                    wait_bar(bar0 + 8 * (B_PFULL + b), pph);
                    tc_fence_after_sync();
                    FS_TRACE(7, Gp);                                  // PV: P seen
                    const uint32_t dV = dV0 + vs * TILE16;
                    const uint32_t a_p = tmem_base + COL_P + b * 64;
                    const uint32_t acc0 = first ? 0u : 1u;
#pragma unroll
                    for (int kk = 0; kk < BN / 16; ++kk) {
                        // A = P in TMEM: keys 16kk.. at column 8kk of the P buffer; B = V as MN-major (16 key rows = 2048 B)
                        umma_ts_lo(tmem_base + COL_O, a_p + (uint32_t)kk * 8, dV + (uint32_t)kk * (2048 >> 4), idesc_pv,
                                   kk ? 1u : acc0);
                    }
                    FS_TRACE(8, Gp);                                  // PV: issued
                    commit_bar(bar0 + 8 * (B_PFREE + b));             // P[b] free; O holds tile Gp
                    commit_bar(bar0 + 8 * (B_VEMPTY + vs));
                    pph ^= 1u;
                    if (++vs == VST) {
                        vs = 0;
                        vph ^= 1u;
                    }
                    first = false;
                    if (--left == 0) {                                // the commit on PFREE above also tells the item's owner that O is final
                        ++ne;
                        left = next_count();
                        first = true;
                    }
                    FS_TRACE(9, Gp);                                  // PV: committed
                    ++Gp;
                };
                while (left > 0) {
                    do_pv(0u, pph0);
                    if (left <= 0) break;
                    do_pv(1u, pph1);
                }
            }
        }
    } else {
        // ------------------------------------------------------------ softmax warpgroups: alternate tiles of the CTA's tile stream
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS_SOFTMAX));
        const int w = warp >> 2;                                  // warpgroup = parity of the global tile indices it takes
        const int r = (warp & 3) * 32 + (threadIdx.x & 31);       // query row in tile == TMEM lane
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const float sl2 = p.scale * LOG2E;
        const uint32_t m_sh_s = smem_u32(m_sh);
        uint32_t g = 0, dep = 0;                                  // tiles / deposits (items with >= 2 tiles) so far
        for (int ic = 0;; ++ic) {
            const Item it = load_item(items, ic);
            if (it.b < 0) break;
            const int n = it.n_tiles;
            const int q0 = it.q_tile * BM, qi = q0 + r;
            // the warpgroup that takes the item's last tile writes the item out; the other one (if it has tiles at all)
            // deposits its partial row sum and moves on
            const bool owner = n > 0 ? (int)((g + (uint32_t)n - 1u) & 1u) == w : w == 0;
            float m_ref = -CUDART_INF_F, l = 0.f;                 // l is relative to m_ref
            // destination of my row: loaded now, used by the epilogue (keeps two dependent global loads off its path)
            bool row_ok = false;
            int64_t dst = 0;
            if (owner) {
                const int64_t bt = (int64_t)it.b * T + qi;
                row_ok = (qi < T) && (!p.qflag || (int)p.qflag[bt] == it.variant);
                dst = row_ok ? (p.out_row ? (int64_t)p.out_row[bt] : bt) : 0;
            }
            for (int j = (int)((g & 1u) ^ (uint32_t)w); j < n; j += 2) {
                const uint32_t G = g + (uint32_t)j;
                const uint32_t colS = (G & 1u) * 128;
                const int kv0 = (it.first_tile + j) * BN;
                const bool need_mask = (CAUSAL && kv0 + BN - 1 > q0) || (kv0 + BN > it.kve) || (kv0 < it.kvs);
                if ((threadIdx.x & 127) == 0) FS_TRACE(10, G);        // softmax: start waiting for S
                wait_bar(bar0 + 8 * (B_SFULL + (G & 1u)), (G >> 1) & 1u);
                tc_fence_after_sync();
                if ((threadIdx.x & 127) == 0) FS_TRACE(11, G);        // softmax: S seen
                uint32_t sv[128];
                tmem_ld32(lane_addr + colS, sv);
                tmem_ld32(lane_addr + colS + 32, sv + 32);
                tmem_ld32(lane_addr + colS + 64, sv + 64);
                tmem_ld32(lane_addr + colS + 96, sv + 96);
                tc_wait_ld();
                if ((threadIdx.x & 127) == 0) FS_TRACE(12, G);        // softmax: scores in registers
                tc_fence_before_sync();
                mbar_arrive(bars + B_SFREE + (G & 1u));               // S(G) is in registers: QK(G+2) may overwrite the buffer
                float mx;
                if (!need_mask) mx = mask_max<false, CAUSAL, false>(sv, kv0, qi, it.kvs, it.kve);
                else if (kv0 >= it.kvs) mx = mask_max<true, CAUSAL, false>(sv, kv0, qi, it.kvs, it.kve);
                else mx = mask_max<true, CAUSAL, true>(sv, kv0, qi, it.kvs, it.kve);
                if ((threadIdx.x & 127) == 0) FS_TRACE(13, G);        // softmax: max done
                // ---- running max of the row after the previous tile of the stream (the other warpgroup's).  Every tile waits
                // for its predecessor's publication, also across items (value unused then): that keeps the two m_sh slots and
                // the two barriers strictly alternating.
                if (G > 0) wait_bar(bar0 + 8 * (B_MPUB + ((G - 1) & 1u)), ((G - 1) >> 1) & 1u);
                if ((threadIdx.x & 127) == 0) FS_TRACE(14, G);        // softmax: previous tile's max seen
                float m_cur;
                if (j == 0) {
                    m_cur = mx;
                } else {
                    const float m_prev = lds_f32(m_sh_s + (((G - 1) & 1u) * 128 + r) * 4);
                    const float m_new = fmaxf(m_prev, mx);
                    // lazy correction: rescale O only when the running max moved by more than 2^8
                    const bool grow = (m_new - m_prev) * sl2 > 8.f;       // also true when m_prev == -inf and m_new finite
                    m_cur = grow ? m_new : m_prev;
                    if (__any_sync(0xffffffffu, grow)) {
                        // PV of tile j-1 has landed in O.  The other warpgroup's barrier could still be in PV(G-3)'s phase (a
                        // false positive for the wait below, see the parity note); MMAs of one thread retire in order, so
                        // seeing my own PV(G-2) retired first rules that out.  (Waiting for PV(G-3)'s phase instead would
                        // deadlock whenever PV(G-1) has already retired.)
                        if (G >= 2) wait_bar(bar0 + 8 * (B_PFREE + (G & 1u)), ((G - 2) >> 1) & 1u);
                        wait_bar(bar0 + 8 * (B_PFREE + ((G - 1) & 1u)), ((G - 1) >> 1) & 1u);
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
                sts_f32(m_sh_s + ((G & 1u) * 128 + r) * 4, m_cur);
                mbar_arrive(bars + B_MPUB + (G & 1u));
                if ((threadIdx.x & 127) == 0) FS_TRACE(15, G);        // softmax: running max published
                // ---- my partial row sum follows the reference maximum
                if (m_ref != m_cur) {
                    l = (m_ref == -CUDART_INF_F) ? 0.f : l * fast_ex2((m_ref - m_cur) * sl2);
                    m_ref = m_cur;
                }
                const float m_off = (m_cur == -CUDART_INF_F) ? 0.f : m_cur * sl2;
                if (G >= 2) {                                         // PV(G-2) has consumed P[G & 1]
                    wait_bar(bar0 + 8 * (B_PFREE + (G & 1u)), ((G - 2) >> 1) & 1u);
                    tc_fence_after_sync();
                }
                if ((threadIdx.x & 127) == 0) FS_TRACE(16, G);        // softmax: P buffer free
                l += exp_store<POLY>(lane_addr + COL_P + (G & 1u) * 64, sv, sl2, m_off);
                tc_wait_st();
                tc_fence_before_sync();
                mbar_arrive(bars + B_PFULL + (G & 1u));
                if ((threadIdx.x & 127) == 0) FS_TRACE(17, G);        // softmax: P stored, arrived
            }
            if (n >= 2 && !owner) {
                // ---- hand my (reference max, partial sum) to the owner; the slot is free once the previous deposit was read
                if (dep > 0) wait_bar(bar0 + 8 * B_LFREE, (dep - 1) & 1u);
                lm_sh[r] = m_ref;
                lm_sh[128 + r] = l;
                mbar_arrive(bars + B_LDEP);
            }
            if (owner) {
                // ---- end of the item: the last reference maximum is the larger of the two (references only grow); O is
                // relative to it
                if (threadIdx.x == w * 128) FS_TRACE(18, ic);         // epilogue (row = item): start
                float m_fin = m_ref, l_tot = l;
                if (n >= 2) {
                    wait_bar(bar0 + 8 * B_LDEP, dep & 1u);
                    const float m_o = lm_sh[r], l_o = lm_sh[128 + r];
                    mbar_arrive(bars + B_LFREE);
                    m_fin = fmaxf(m_ref, m_o);
                    const float a = (m_ref == -CUDART_INF_F) ? 0.f : l * fast_ex2((m_ref - m_fin) * sl2);
                    const float b = (m_o == -CUDART_INF_F) ? 0.f : l_o * fast_ex2((m_o - m_fin) * sl2);
                    l_tot = a + b;
                }
                const float inv_l = l_tot > 0.f ? 1.f / l_tot : 0.f;
                if (threadIdx.x == w * 128) FS_TRACE(19, ic);         // epilogue: deposit seen
                if (n > 0) {
                    uint32_t ov[D];
                    const uint32_t Gl = g + (uint32_t)n - 1u;             // my last tile: its PV is the item's last MMA
                    wait_bar(bar0 + 8 * (B_PFREE + (Gl & 1u)), (Gl >> 1) & 1u);
                    tc_fence_after_sync();
                    if (threadIdx.x == w * 128) FS_TRACE(20, ic);     // epilogue: O final
#pragma unroll
                    for (int c = 0; c < D / 32; ++c) tmem_ld32(lane_addr + COL_O + c * 32, ov + c * 32);
                    tc_wait_ld();
                    tc_fence_before_sync();
                    mbar_arrive(bars + B_OFREE);                      // O is in registers: the next item's PV may start
                    if (threadIdx.x == w * 128) FS_TRACE(21, ic);     // epilogue: O released
                    // O leaves through a per-warp staging tile (32 rows x 32 columns bf16 at a time, 16-byte chunks XOR-swizzled
                    // by row) so that a store instruction covers 8 rows x 64 contiguous bytes.  Shared-window addresses and the
                    // four destination pointers are set up once per item: the first version of this block compiled to ~2400
                    // instructions of generic LD/ST and 64-bit address arithmetic and held the owner for ~3 k clk per item.
                    const int lane = threadIdx.x & 31;
                    const uint32_t stage_s = smem_u32(smem + S::STAGE_OFF) + (uint32_t)(w * 4 + (warp & 3)) * (32 * 64);
                    const uint32_t my_row_s = stage_s + (uint32_t)lane * 64;
                    const uint32_t my_swz = (uint32_t)(lane >> 1) & 3u;
                    const int32_t dst32 = row_ok ? (int32_t)dst : -1;
                    const int64_t ld_o = (int64_t)p.heads * D;
                    __nv_bfloat16* obase = p.O + (int64_t)it.h * D + (lane & 3) * 8;
                    uint32_t rd_s[4];
                    __nv_bfloat16* gp[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int row = i * 8 + (lane >> 2);
                        const int32_t drow = __shfl_sync(0xffffffffu, dst32, row);
                        rd_s[i] = stage_s + (uint32_t)row * 64 + ((((uint32_t)lane & 3u) ^ ((uint32_t)(row >> 1) & 3u)) << 4);
                        gp[i] = drow >= 0 ? obase + (int64_t)drow * ld_o : nullptr;
                    }
#pragma unroll
                    for (int q = 0; q < D / 32; ++q) {
#pragma unroll
                        for (int c = 0; c < 4; ++c) {
                            const int e = q * 32 + c * 8;
                            uint4 o;
                            o.x = pack_bf16(__uint_as_float(ov[e + 0]) * inv_l, __uint_as_float(ov[e + 1]) * inv_l);
                            o.y = pack_bf16(__uint_as_float(ov[e + 2]) * inv_l, __uint_as_float(ov[e + 3]) * inv_l);
                            o.z = pack_bf16(__uint_as_float(ov[e + 4]) * inv_l, __uint_as_float(ov[e + 5]) * inv_l);
                            o.w = pack_bf16(__uint_as_float(ov[e + 6]) * inv_l, __uint_as_float(ov[e + 7]) * inv_l);
                            sts128(my_row_s + (((uint32_t)c ^ my_swz) << 4), o);
                        }
                        __syncwarp();
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            const uint4 o = lds128(rd_s[i]);
                            if (gp[i]) *reinterpret_cast<uint4*>(gp[i] + q * 32) = o;
                        }
                        __syncwarp();
                    }
                } else if (row_ok) {                                  // no visible key at all: the row is zero
                    uint4* orow = reinterpret_cast<uint4*>(p.O + dst * ((int64_t)p.heads * D) + (int64_t)it.h * D);
#pragma unroll
                    for (int e = 0; e < D / 8; ++e) orow[e] = make_uint4(0u, 0u, 0u, 0u);
                }
                if (row_ok && p.lse) {
                    // natural-log LSE of the scaled scores; +inf marks a row with no visible key (P == 0 in backward)
                    p.lse[((int64_t)it.b * p.heads + it.h) * T + qi] = l_tot > 0.f ? (m_fin * p.scale + __logf(l_tot)) : CUDART_INF_F;
                }
                if (threadIdx.x == w * 128) FS_TRACE(22, ic);         // epilogue: stores issued
            }
            if (n >= 2) ++dep;
            g += (uint32_t)n;
        }
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == WARP_QK) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    if (p.cta_log && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        int n_it = 0, tiles = 0;
#pragma unroll 1
        for (int k = 0; k < MAX_ITEMS && items[k].b >= 0; ++k) {
            ++n_it;
            tiles += items[k].n_tiles;
        }
        long long* e = p.cta_log + (int64_t)blockIdx.x * 8;
        e[0] = smid; e[1] = n_it; e[2] = tiles; e[3] = t_entry; e[5] = clock64();
    }
}

template <int D, bool CAUSAL, int POLY, bool TRACE>
static int launch_t(const CUtensorMap* tm, const Params& p, cudaStream_t st) {
    using S = Smem<D>;
    auto kern = attn_fwd_stream_kernel<D, CAUSAL, POLY, TRACE>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail(LB_ELAUNCH, "attn_fwd_stream: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    kern<<<(unsigned)p.n_cta, THREADS, S::TOTAL, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], p);
    return check_launch("attn_fwd_stream");
}

template <int D, bool CAUSAL, int POLY>
static int launch_p(const CUtensorMap* tm, const Params& p, cudaStream_t st) {
    // the clock64 stamps are compiled into a separate instantiation (diagnostics; only the plain SFU variant has one)
    if (POLY == 0 && p.trace) return launch_t<D, CAUSAL, 0, true>(tm, p, st);
    return launch_t<D, CAUSAL, POLY, false>(tm, p, st);
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

/* diagnostics: CTA 0 of subsequent lb_attn_fwd_stream launches writes clock64 stamps into `buf` ([64][32] int64, device; one
 * row per kv tile; slots 0-9 tcgen05 thread, 10-17 softmax thread 0 of the tile's warpgroup, see FS_TRACE sites). NULL = off */
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

extern "C" int lb_attn_fwd_stream_max_cta_items(void) { return fs::MAX_ITEMS; }

extern "C" int lb_attn_fwd_stream(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1,
                                  const uint8_t* qflag, const int32_t* work, int n_work, const int32_t* plan_items,
                                  const int32_t* plan_off, int n_cta, int max_cta_items, int head_group,
                                  const int32_t* kv_start, const int32_t* kv_end, const int32_t* out_row, void* O, float* lse,
                                  int batch, int seqlen, int heads, int head_dim, int causal, float scale, void* stream) {
    LB_REQUIRE(batch > 0 && seqlen > 0 && heads > 0 && n_work >= 0, LB_EINVAL, "attn_fwd_stream: bad shape");
    LB_REQUIRE(head_dim == 64 || head_dim == 128, LB_EINVAL, "attn_fwd_stream: head_dim %d (64 or 128 supported)", head_dim);
    LB_REQUIRE(Q && K0 && V0 && O && work, LB_EINVAL, "attn_fwd_stream: null argument");
    LB_REQUIRE((plan_items == nullptr) == (plan_off == nullptr), LB_EINVAL, "attn_fwd_stream: plan_items and plan_off go together");
    LB_REQUIRE(!plan_items || (n_cta > 0 && max_cta_items > 0), LB_EINVAL,
               "attn_fwd_stream: a plan needs n_cta (%d) and max_cta_items (%d)", n_cta, max_cta_items);
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
    int per_cta;
    if (plan_items) {
        p.n_cta = n_cta;
        per_cta = max_cta_items;
    } else {
        const int sms = sm_count();
        if (sms <= 0) return fail(LB_ELAUNCH, "attn_fwd_stream: no SM count");
        p.n_cta = p.n_items < sms ? p.n_items : sms;
        per_cta = (p.n_items + p.n_cta - 1) / p.n_cta;
    }
    LB_REQUIRE(per_cta <= fs::MAX_ITEMS, LB_EINVAL,
               "attn_fwd_stream: %d items per CTA exceed the in-kernel table (%d); use lb_attn_fwd for this shape", per_cta,
               fs::MAX_ITEMS);
    p.cta_log = g_stream_cta_log;
    p.trace = g_stream_trace;
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 128) return causal ? fs::launch<128, true>(tm, p, st) : fs::launch<128, false>(tm, p, st);
    return causal ? fs::launch<64, true>(tm, p, st) : fs::launch<64, false>(tm, p, st);
}

// Skinny GEMM for the one-token decode step (N1): C[M <= 32, N] = A[M,K] . W[N,K]^T (+ bias, SwiGLU, addend), bf16.
//
// With M = batch (8) rows a projection is a weight-streaming problem: 2 bytes of W per M*2 FLOP, i.e. HBM-bound by a factor of
// ~50.  The training GEMM (gemm_grouped.cu) tiles the OUTPUT in 256 x 256 blocks -- a 4096 x 4096 weight becomes 16 tiles on 32
// of the 148 SMs, each streaming its weight slab through one TMA ring: the decode step measured 9.9 ms against a 2.7 ms
// streaming floor.  This kernel is the same tcgen05 machinery with the roles swapped and the work cut for bandwidth:
//   * swap-AB: the WEIGHT rows are the M = 128 operand (TMEM lanes = output features), the tokens are N = 16 or 32 columns:
//     one tcgen05.mma (M128, N16, K16) per 16 weight columns -- the tensor pipe idles, as it should;
//   * work unit = (problem, 128-feature tile, K split): ~2 units per SM, all resident at once (two CTAs of ~105 KB per SM: one
//     streams while the other starts up or reduces; a first version with one 200 KB CTA per SM and 160 units on 148 SMs ran
//     two waves and reached 17 % of the HBM rate on a 4096 x 4096 weight); each CTA runs a TMA ring of five 16 KB weight
//     stages (160 KB in flight per SM; the bandwidth-delay product of HBM3e is ~66 KB per SM);
//   * split-K partial sums go to an fp32 workspace; the LAST CTA of a tile (a self-resetting counter) adds the partials in
//     split order -- deterministic -- and applies the epilogue: bias, bf16 rounding, SwiGLU of a (gate, up) weight pair
//     (two accumulators), addend (the residual), with the rounding sequence of gemm_grouped.cu;
//   * several problems per launch (q / k / v of a layer share one launch and one read of x);
//   * what the phase trace (lb_gemm_skinny_set_trace, scripts/gemm_skinny_trace.py) showed about the ~8 us a launch costs
//     beyond its streaming time, and what the code does about it: the problem's fields sit behind a run-time index in the
//     kernel parameters, so every use was an indexed constant load re-issued after each asm barrier (0.3 us per output row)
//     -> read once into registers; the residual addend was requested only after the accumulator barrier (2-4 us of exposed
//     latency) -> requested while the weights stream; the tile's last CTA fetched nine partials one split at a time ->
//     sixteen float4 in flight, and the partials carry only the real tokens (8 of the 16 padded columns at batch 8);
//   * programmatic dependent launch: the first ring pass of WEIGHTS is issued before griddepcontrol.wait, so it streams
//     while the previous kernel of the decode chain drains (lb_set_pdl).
// Replaces for M <= 32 the same reference products as gemm_grouped.cu (modeling_llama.py:185-201, modeling_libra.py:192-238,
// 1018-1052); the chained low-rank form and the backward layouts stay on gemm_grouped.cu.
#include <stdlib.h>

#include "common.cuh"

namespace lb {
namespace sk {

constexpr int BM = 128;                 // weight rows (output features) per tile = TMEM lanes
constexpr int BK = 64;                  // K elements per stage (one 128 B swizzle row)
constexpr int MAXP = 8;                 // problems per launch
constexpr int THREADS = 192;            // warp 0 TMA, warp 1 MMA (+ TMEM alloc), warps 2-5 epilogue
constexpr int W_BYTES = BM * BK * 2;    // 16 KB
constexpr int A_BYTES = 32 * BK * 2;    // 4 KB (slot size; 16-token problems fill half)
// work units per SM the split-K factor aims for (LB_SKINNY_UNITS_X2 = twice the value, for experiments: 2 = one unit per SM)
static int target_units_x2() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LB_SKINNY_UNITS_X2");
        v = e ? atoi(e) : 4;
        if (v < 1 || v > 16) v = 4;
    }
    return v;
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// Split-K reduction of one output row: the partials of 16 / NF4 splits are loaded together (each is an L2 round trip of ~1 us;
// issued one split at a time, nine splits cost ~7 us of the o-proj / down-proj launches), then added in split order, so the
// sum is the same for every launch.  Row layout of a partial: [acc (TOK) | acc2 (TOK, SwiGLU only)] fp32; only the first Q4
// float4 of each half carry tokens (Q4 = ceil(M / 4) rounded up to a power of two): with 8 tokens all of nine splits fit one batch.
template <int Q4, bool TWO>
__device__ __forceinline__ void reduce_splits(const float* __restrict__ src0, int splits, int tok, float (&acc)[32], float (&acc2)[32]) {
    constexpr int NF4 = Q4 * (TWO ? 2 : 1), RB = 16 / NF4 > 9 ? 9 : 16 / NF4;
    const int64_t stride = (int64_t)BM * tok * (TWO ? 2 : 1);
    for (int sp0 = 0; sp0 < splits; sp0 += RB) {
        float4 buf[RB][NF4];
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            const float* src = src0 + (int64_t)min(sp0 + r, splits - 1) * stride;
#pragma unroll
            for (int f = 0; f < Q4; ++f) buf[r][f] = __ldcg(reinterpret_cast<const float4*>(src) + f);
            if (TWO) {
#pragma unroll
                for (int f = 0; f < Q4; ++f) buf[r][Q4 + f] = __ldcg(reinterpret_cast<const float4*>(src + tok) + f);
            }
        }
#pragma unroll
        for (int r = 0; r < RB; ++r) {
            if (sp0 + r < splits) {
#pragma unroll
                for (int f = 0; f < Q4; ++f) {
                    const float4 v = buf[r][f];
                    acc[4 * f] += v.x; acc[4 * f + 1] += v.y; acc[4 * f + 2] += v.z; acc[4 * f + 3] += v.w;
                    if (TWO) {
                        const float4 u = buf[r][Q4 + f];
                        acc2[4 * f] += u.x; acc2[4 * f + 1] += u.y; acc2[4 * f + 2] += u.z; acc2[4 * f + 3] += u.w;
                    }
                }
            }
        }
    }
}

template <bool TWO>
__device__ __forceinline__ void reduce_dispatch(const float* __restrict__ src0, int splits, int tok, int M, float (&acc)[32], float (&acc2)[32]) {
    if (M <= 4) reduce_splits<1, TWO>(src0, splits, tok, acc, acc2);
    else if (M <= 8) reduce_splits<2, TWO>(src0, splits, tok, acc, acc2);
    else if (M <= 16) reduce_splits<4, TWO>(src0, splits, tok, acc, acc2);
    else reduce_splits<8, TWO>(src0, splits, tok, acc, acc2);
}

struct Prob {
    int M, N, tok;                      // tok = 16 or 32 accumulator columns (tokens padded)
    int mred;                           // token columns the split-K partials carry: M (default) or tok (LB_SKINNY_FULLTOK=1, A/B)
    int num_kb, kb_per_split, splits, tiles_n;
    int unit_begin;
    int dual;                           // SwiGLU: second weight, second accumulator
    int a_map, w_map, w2_map;
    __nv_bfloat16* C;
    const __nv_bfloat16* D;
    const __nv_bfloat16* bias;
    long long ldc, ldd;
    float* partial;                     // [tiles_n][splits][128][tok * (1 + dual)] fp32 (splits > 1)
    int* counters;                      // [tiles_n], zero between launches (the last CTA of a tile resets its counter)
};

struct Params {
    CUtensorMap maps[3 * MAXP];
    Prob prob[MAXP];
    int n_prob, total_units;
    long long* trace;                   // debug (lb_gemm_skinny_set_trace): [unit][12] globaltimer ns at the phases below, or null
};

__device__ __forceinline__ long long gtime() {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
#define SK_TRACE(slot) do { if (p.trace) p.trace[(long long)blockIdx.x * 12 + (slot)] = gtime(); } while (0)

// DUAL: the launch holds at least one SwiGLU problem (stage = W | W2 | A, fewer stages); plain problems in it skip W2
// MINB: resident CTAs per SM the kernel is compiled for.  2 (five / three stages, ~105 KB): the stand-alone optimum.  3 (three /
// two stages, <= 74 KB): a kernel's two units per SM leave one slot free, so under a programmatic dependent launch the NEXT
// kernel's CTAs become resident while this one streams, run their prologue and fill their weight ring, and the HBM pipe does
// not drain between the launches of the decode chain.
template <int STAGES, bool DUAL, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) gemm_skinny_kernel(const __grid_constant__ Params p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int A_OFF = (DUAL ? 2 : 1) * W_BYTES;
    constexpr int STAGE_BYTES = A_OFF + A_BYTES;                    // W | (W2) | A
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);      // full[STAGES] empty[STAGES] done
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 1);
    __shared__ int s_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int g = 0;
#pragma unroll 1
    while (g + 1 < p.n_prob && (int)blockIdx.x >= p.prob[g + 1].unit_begin) ++g;
    const Prob& pb = p.prob[g];
    const int local = (int)blockIdx.x - pb.unit_begin;
    const int tile = local / pb.splits, split = local - tile * pb.splits;
    const int kb0 = split * pb.kb_per_split;
    const int kb1 = min(pb.num_kb, kb0 + pb.kb_per_split);
    const int nkb = kb1 - kb0;                                      // >= 1 by construction
    // fields used inside the pipeline loops, read once (indexed constant loads otherwise: see the epilogue)
    const CUtensorMap* const map_w = &p.maps[pb.w_map];
    const CUtensorMap* const map_w2 = &p.maps[pb.w2_map];
    const CUtensorMap* const map_a = &p.maps[pb.a_map];
    const bool pb_dual = pb.dual != 0;
    const int pb_tok = pb.tok;
    constexpr uint32_t TMEM_COLS = 64;

    pdl_trigger();
    if (threadIdx.x == 0) {
        SK_TRACE(0);                                                // CTA start
        for (int i = 0; i < 2 * STAGES + 1; ++i) mbar_init(bars + i, 1);
        fence_barrier_init();
    }
    if (warp == 0 && elect_one()) {
        tma_prefetch_desc(map_w);
        tma_prefetch_desc(map_a);
        if (pb_dual) tma_prefetch_desc(map_w2);
    }
    if (warp == 1) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t bar0 = smem_u32(bars);

    if (warp == 0) {
        if (elect_one()) {
            const uint32_t tx = (uint32_t)((pb_dual ? 2 : 1) * W_BYTES + pb_tok * BK * 2);
            // the first ring pass of WEIGHTS depends on no earlier kernel: under a programmatic dependent launch it streams
            // while the previous kernel of the chain drains; the activations follow once that kernel's writes are visible
            const int pre = min(nkb, STAGES);
            for (int i = 0; i < pre; ++i) {
                uint8_t* st = smem + i * STAGE_BYTES;
                mbar_arrive_expect_tx(bars + i, tx);
                tma_load_2d(st, map_w, bars + i, (kb0 + i) * BK, tile * BM);
                if (DUAL && pb_dual) tma_load_2d(st + W_BYTES, map_w2, bars + i, (kb0 + i) * BK, tile * BM);
            }
            SK_TRACE(1);                                            // prologue done, first weight loads issued
            pdl_wait();
            SK_TRACE(2);                                            // grid dependency resolved
            for (int i = 0; i < pre; ++i)
                tma_load_2d(smem + i * STAGE_BYTES + A_OFF, map_a, bars + i, (kb0 + i) * BK, 0);
            for (int i = pre; i < nkb; ++i) {
                const int s = i % STAGES;
                wait_bar(bar0 + 8 * (STAGES + s), ((uint32_t)(i / STAGES) & 1u) ^ 1u);
                uint8_t* st = smem + s * STAGE_BYTES;
                mbar_arrive_expect_tx(bars + s, tx);
                tma_load_2d(st, map_w, bars + s, (kb0 + i) * BK, tile * BM);
                if (DUAL && pb_dual) tma_load_2d(st + W_BYTES, map_w2, bars + s, (kb0 + i) * BK, tile * BM);
                tma_load_2d(st + A_OFF, map_a, bars + s, (kb0 + i) * BK, 0);
            }
        }
    } else if (warp == 1) {
        if (elect_one()) {
            const uint32_t idesc = make_idesc_bf16(BM, pb_tok, 0, 0);
            for (int i = 0; i < nkb; ++i) {
                const int s = i % STAGES;
                wait_bar(bar0 + 8 * s, (uint32_t)(i / STAGES) & 1u);
                if (i == 0) SK_TRACE(3);                            // first stage landed
                tc_fence_after_sync();
                const uint32_t st = smem_u32(smem + s * STAGE_BYTES);
                const uint32_t dW = desc_lo_kmajor(st), dW2 = desc_lo_kmajor(st + W_BYTES), dA = desc_lo_kmajor(st + A_OFF);
#pragma unroll
                for (int kk = 0; kk < BK / 16; ++kk) {
                    umma_ss_lo(tmem_base, dW + 2 * kk, dA + 2 * kk, idesc, (i | kk) ? 1u : 0u);
                    if (DUAL && pb_dual) umma_ss_lo(tmem_base + 32, dW2 + 2 * kk, dA + 2 * kk, idesc, (i | kk) ? 1u : 0u);
                }
                commit_bar(bar0 + 8 * (STAGES + s));
            }
            commit_bar(bar0 + 8 * (2 * STAGES));
        }
    } else {
        // ---------------- epilogue: thread <-> output feature (TMEM lane), registers <-> tokens
        // The problem's fields live in the kernel parameters behind a run-time index: every use is an indexed constant load,
        // and the asm barriers / fences below keep the compiler from hoisting them -- the traced epilogue spent 0.3 us per
        // output row on them.  Read them ONCE; and request the addend and bias rows now, while the weights stream, instead
        // of after the accumulator barrier (2-4 us of exposed load latency per launch with a residual).
        const int row = (warp & 3) * 32 + lane;                     // warps 2..5 own TMEM lanes 32*(warp%4)..
        const uint32_t taddr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16);
        const int tok = pb.tok, pM = pb.M, pN = pb.N, splits = pb.splits, mred = pb.mred;
        const bool dual = pb.dual != 0;
        __nv_bfloat16* const pC = pb.C;
        const __nv_bfloat16* const pD = pb.D;
        const __nv_bfloat16* const pbias = pb.bias;
        const long long ldc = pb.ldc, ldd = pb.ldd;
        float* const ppartial = pb.partial;
        int* const pcounters = pb.counters;
        const int n = tile * BM + row;
        pdl_wait();                                                 // these threads read the addend and write C / the shared workspace
        if (threadIdx.x == 64) SK_TRACE(8);
        float dd[16];                                               // addend of tokens 0..15 (tokens 16..31 are loaded at the end: registers)
#pragma unroll
        for (int m = 0; m < 16; ++m) dd[m] = (pD && m < pM && n < pN) ? __bfloat162float(__ldg(pD + (int64_t)m * ldd + n)) : 0.f;
        const float bias = (pbias && n < pN) ? __bfloat162float(__ldg(pbias + n)) : 0.f;
        wait_bar(bar0 + 8 * (2 * STAGES), 0);
        if (threadIdx.x == 64) SK_TRACE(4);                         // accumulator complete (all K blocks streamed)
        tc_fence_after_sync();
        float acc[32], acc2[32];
        {
            uint32_t r[32];
            if (tok == 32) tmem_ld32(taddr, r); else tmem_ld16(taddr, r);
            tc_wait_ld();
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = j < tok ? __uint_as_float(r[j]) : 0.f;
            if (dual) {
                if (tok == 32) tmem_ld32(taddr + 32, r); else tmem_ld16(taddr + 32, r);
                tc_wait_ld();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc2[j] = j < tok ? __uint_as_float(r[j]) : 0.f;
            }
        }
        if (threadIdx.x == 64) SK_TRACE(9);                         // accumulator in registers
        bool finalize = true;
        if (splits > 1) {
            const int w = tok * (dual ? 2 : 1);                     // floats per row of a partial
            float* mine = ppartial + (((int64_t)tile * splits + split) * BM + row) * w;
#pragma unroll
            for (int j = 0; j < 32; j += 4)
                if (j < mred) *reinterpret_cast<float4*>(mine + j) = make_float4(acc[j], acc[j + 1], acc[j + 2], acc[j + 3]);
            if (dual) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    if (j < mred) *reinterpret_cast<float4*>(mine + tok + j) = make_float4(acc2[j], acc2[j + 1], acc2[j + 2], acc2[j + 3]);
            }
            __threadfence();
            named_bar_sync(1, 128);
            if (threadIdx.x == 64) s_last = atomicAdd(pcounters + tile, 1) == splits - 1 ? 1 : 0;
            named_bar_sync(1, 128);
            if (threadIdx.x == 64) SK_TRACE(5);                     // partial published, arrival counted
            finalize = s_last != 0;
            if (finalize) {                                         // the tile's last CTA: add the partials in split order
                __threadfence();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[j] = acc2[j] = 0.f;
                const float* src0 = ppartial + ((int64_t)tile * splits * BM + row) * w;
                if (dual) reduce_dispatch<true>(src0, splits, tok, mred, acc, acc2); else reduce_dispatch<false>(src0, splits, tok, mred, acc, acc2);
                if (threadIdx.x == 64) pcounters[tile] = 0;         // ready for the next launch
            }
        }
        if (finalize && n < pN) {
            float dhi[16];                                          // addend of tokens 16..31, requested together
#pragma unroll
            for (int m = 0; m < 16; ++m) dhi[m] = (pD && m + 16 < pM) ? __bfloat162float(__ldg(pD + (int64_t)(m + 16) * ldd + n)) : 0.f;
            __nv_bfloat16* crow = pC + n;
#pragma unroll
            for (int m = 0; m < 32; ++m) {
                if (m < pM) {
                    float v;
                    if (dual) {
                        // reference: silu(gate) rounded to bf16, then * up (modeling_libra.py:232-233); gate / up rounded first (nn.Linear)
                        const float gt = round_bf16(acc[m]), up = round_bf16(acc2[m]);
                        v = round_bf16(round_bf16(silu_f(gt)) * up);
                    } else {
                        v = round_bf16(acc[m] + bias);
                    }
                    if (pD) v = round_bf16(v + (m < 16 ? dd[m & 15] : dhi[m & 15]));      // bf16(bf16(x W^T) + addend)
                    crow[(int64_t)m * ldc] = __float2bfloat16_rn(v);
                    if (m == 0 && threadIdx.x == 64) SK_TRACE(10);  // first output element
                }
            }
        }
        if (threadIdx.x == 64) SK_TRACE(6);                         // output (or nothing, for a non-final split) written
        tc_fence_before_sync();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, TMEM_COLS);
        if (lane == 0) SK_TRACE(7);                                 // CTA end
    }
}

constexpr int STAGES_SINGLE = 5, STAGES_DUAL = 3;       // ~105 KB per CTA: two CTAs per SM, one streaming while the other starts or reduces
constexpr int STAGES_SINGLE3 = 3, STAGES_DUAL3 = 2;     // <= 74 KB per CTA: three CTAs per SM (see MINB)
// LB_SKINNY_SLOTS=2|3 selects the variant; default 2.  Measured on the Libra-11B decode step (B=8, graph replay, same box):
// serial launches 5.94 ms; dependent launches with two slots 5.58 ms; with three slots 5.93 ms (the shallower rings cost the
// streaming phase what the earlier start gains) -- kept selectable for other shapes.
static int slots_per_sm() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LB_SKINNY_SLOTS");
        v = e ? atoi(e) : 0;
        if (v != 2 && v != 3) v = 0;
    }
    return v ? v : 2;
}
template <int STAGES, bool DUAL>
static constexpr int smem_bytes() { return STAGES * ((DUAL ? 2 : 1) * W_BYTES + A_BYTES) + (2 * STAGES + 1) * 8 + 16 + 1024; }

template <int STAGES, bool DUAL, int MINB>
static cudaError_t launch_variant(const Params& local, cudaStream_t st) {
    auto kern = gemm_skinny_kernel<STAGES, DUAL, MINB>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes<STAGES, DUAL>());
        if (e != cudaSuccess) return e;
        configured = true;
    }
    return launch_chain(kern, dim3((unsigned)local.total_units), dim3(THREADS), (size_t)smem_bytes<STAGES, DUAL>(), st, local);
}

static long long* g_trace = nullptr;

constexpr int64_t CTR_BYTES = 64 * 1024;      // tile counters: fixed region at the start of the workspace

struct Plan {
    int64_t partial_floats, counters;
};

// work decomposition shared by the workspace query and the launch
static int plan(const lb_gemm_problem* probs, int n, Params* P, Plan* pl) {
    LB_REQUIRE(probs && n > 0 && n <= MAXP, LB_EINVAL, "gemm_skinny: 1..%d problems per launch, got %d", MAXP, n);
    const int sms = sm_count();
    int tiles_total = 0;
    for (int i = 0; i < n; ++i) tiles_total += ceil_div(probs[i].N, BM);
    int units = 0;
    int64_t pf = 0, ctr = 0;
    for (int i = 0; i < n; ++i) {
        const lb_gemm_problem& q = probs[i];
        LB_REQUIRE(q.M >= 1 && q.M <= 32 && q.N >= 1 && q.K >= 8, LB_EINVAL, "gemm_skinny[%d]: M=%lld must be 1..32, N=%lld, K=%lld", i,
                   (long long)q.M, (long long)q.N, (long long)q.K);
        LB_REQUIRE(!q.trans_a && !q.trans_b && q.wait_on < 0 && !q.alpha && !q.G && !q.U && !(q.flags & LB_GEMM_ACCUMULATE_PREV), LB_EINVAL,
                   "gemm_skinny[%d]: plain x W^T products only (no transposes, chains, alpha, extra outputs)", i);
        LB_REQUIRE(q.epilogue == LB_EPI_NONE || (q.epilogue == LB_EPI_SWIGLU && q.B2 && !q.bias), LB_EINVAL,
                   "gemm_skinny[%d]: epilogue %d unsupported", i, q.epilogue);
        LB_REQUIRE(q.A && q.B && q.C && q.lda % 8 == 0 && q.ldb % 8 == 0 && q.K % 8 == 0, LB_EALIGN,
                   "gemm_skinny[%d]: null operand or lda/ldb/K not multiples of 8", i);
        Prob tmp;
        Prob& g = P ? P->prob[i] : tmp;
        g.M = (int)q.M; g.N = (int)q.N;
        g.tok = q.M <= 16 ? 16 : 32;
        {
            static int fulltok = -1;
            if (fulltok < 0) { const char* e = getenv("LB_SKINNY_FULLTOK"); fulltok = (e && atoi(e) == 1) ? 1 : 0; }
            g.mred = fulltok ? g.tok : g.M;
        }
        g.dual = q.epilogue == LB_EPI_SWIGLU ? 1 : 0;
        g.tiles_n = ceil_div(q.N, BM);
        g.num_kb = ceil_div(q.K, BK);
        int s = (target_units_x2() * sms / 2) / tiles_total;          // rounded DOWN: all units resident at once (2 CTAs per SM)
        if (s > g.num_kb / 4) s = g.num_kb / 4;
        if (s < 1) s = 1;
        g.kb_per_split = ceil_div(g.num_kb, s);
        g.splits = ceil_div(g.num_kb, g.kb_per_split);              // every split has at least one K block
        g.unit_begin = units;
        units += g.tiles_n * g.splits;
        g.C = (__nv_bfloat16*)q.C; g.D = (const __nv_bfloat16*)q.D; g.bias = (const __nv_bfloat16*)q.bias;
        g.ldc = q.ldc; g.ldd = q.ldd;
        g.a_map = 3 * i; g.w_map = 3 * i + 1; g.w2_map = 3 * i + 2;
        if (g.splits > 1) pf += (int64_t)g.tiles_n * g.splits * BM * g.tok * (1 + g.dual);
        ctr += g.tiles_n;
        if (!P) continue;
        int rc = make_tmap_bf16_2d(&P->maps[g.a_map], q.A, (uint64_t)q.M, (uint64_t)q.K, (uint64_t)q.lda, (uint32_t)g.tok, BK);
        if (rc) return rc;
        rc = make_tmap_bf16_2d(&P->maps[g.w_map], q.B, (uint64_t)q.N, (uint64_t)q.K, (uint64_t)q.ldb, BM, BK);
        if (rc) return rc;
        if (g.dual) {
            rc = make_tmap_bf16_2d(&P->maps[g.w2_map], q.B2, (uint64_t)q.N, (uint64_t)q.K, (uint64_t)q.ldb, BM, BK);
            if (rc) return rc;
        }
    }
    if (P) {
        P->n_prob = n;
        P->total_units = units;
        P->trace = g_trace;
    }
    pl->partial_floats = pf;
    pl->counters = ctr;
    return LB_OK;
}

}  // namespace sk
}  // namespace lb

using namespace lb;

extern "C" {

/* bytes of workspace lb_gemm_skinny needs for this problem list: [64 KB of int32 tile counters | fp32 split-K partials].
 * The counter region has a FIXED size so that launches with different problem lists can share one workspace: it must be zero
 * before the first launch and every launch leaves it zero; the partials need no initialisation. */
int64_t lb_gemm_skinny_workspace_bytes(const lb_gemm_problem* problems, int n) {
    sk::Plan pl{};
    if (sk::plan(problems, n, nullptr, &pl)) return -1;
    return sk::CTR_BYTES + pl.partial_floats * 4;
}

/* debug: device buffer of [units][12] int64 the next launches fill with %globaltimer stamps (NULL switches it off) */
int lb_gemm_skinny_set_trace(void* buf) {
    sk::g_trace = (long long*)buf;
    return LB_OK;
}

int lb_gemm_skinny(const lb_gemm_problem* problems, int n, void* workspace, int64_t workspace_bytes, void* stream) {
    int rc = require_sm100();
    if (rc) return rc;
    sk::Plan pl{};
    sk::Params local;
    rc = sk::plan(problems, n, &local, &pl);
    if (rc) return rc;
    const int64_t ctr_bytes = sk::CTR_BYTES;
    LB_REQUIRE(pl.counters * 4 <= ctr_bytes, LB_EINVAL, "gemm_skinny: %lld output tiles exceed the counter region", (long long)pl.counters);
    LB_REQUIRE(workspace && workspace_bytes >= ctr_bytes + pl.partial_floats * 4, LB_EINVAL,
               "gemm_skinny: workspace of %lld bytes needed, %lld given", (long long)(ctr_bytes + pl.partial_floats * 4),
               (long long)workspace_bytes);
    int* counters = (int*)workspace;
    float* partial = (float*)((char*)workspace + ctr_bytes);
    bool any_dual = false;
    for (int i = 0; i < n; ++i) {
        sk::Prob& g = local.prob[i];
        g.counters = counters;
        counters += g.tiles_n;
        g.partial = partial;
        if (g.splits > 1) partial += (int64_t)g.tiles_n * g.splits * sk::BM * g.tok * (1 + g.dual);
        any_dual |= g.dual != 0;
    }
    cudaStream_t st = (cudaStream_t)stream;
    const bool three = sk::slots_per_sm() == 3;
    cudaError_t e = cudaSuccess;
    if (any_dual && three) e = sk::launch_variant<sk::STAGES_DUAL3, true, 3>(local, st);
    else if (any_dual) e = sk::launch_variant<sk::STAGES_DUAL, true, 2>(local, st);
    else if (three) e = sk::launch_variant<sk::STAGES_SINGLE3, false, 3>(local, st);
    else e = sk::launch_variant<sk::STAGES_SINGLE, false, 2>(local, st);
    if (e != cudaSuccess) return fail(LB_ELAUNCH, "gemm_skinny: %s", cudaGetErrorString(e));
    return check_launch("gemm_skinny");
}

}

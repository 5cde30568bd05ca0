// Bridge attention forward, paired-tile kernel (A10; same maths and operand formulation as attn_fwd.cu).
//
// One CTA per SM works on TWO 128-row query tiles ("lanes") of the same (sample, head, variant) at once and streams the
// 64-key K/V tiles they share through TMA rings.  Per lane the score tile is DOUBLE-BUFFERED in TMEM, so the single
// tcgen05 issuer always runs two steps ahead of the softmax:
//
//     step j of lane L:   wait P_L(j)  ->  O_L += P_L(j).V(j)  ;  S_L[j&1] = Q_L.K(j+2)^T
//
// S_L(j+1) was produced while the softmax warpgroup of lane L was still busy with step j, so that warpgroup never waits
// for the tensor pipe: its chain is only load -> max -> exp -> store, back to back, and the two lanes keep the SFU (the
// slowest unit: 16 ex2/clk/SM = 1024 clk per 128x128 scores, the same as the two MMAs) permanently fed.  Measured
// motivation (profiles/r01_attn_pair_trace.md): with one S buffer per lane the chain exp -> P -> PV -> QK -> S
// serialises inside a lane (3660 clk per pair of 128x128 tiles, tensor pipe idle 44 %).
//
// Softmax: one thread per query row (TMEM lane) and all 64 key columns of the step -> the row max and row sum need no
// cross-thread exchange.  P (bf16, 32 columns) is written over the start of its own S buffer and consumed as the TMEM
// A operand of O += P.V.  The running max uses the lazy-rescale rule of attn_fwd.cu.
//
// CTA = 320 threads: warps 0-3 softmax/epilogue of lane A, warps 4-7 of lane B, warp 8 TMA, warp 9 MMA (+TMEM alloc).
// TMEM (512 columns), lane L at 256 L:  S buffer 0 [0,64)   S buffer 1 [64,128)   O [128,128+D).
#include <math_constants.h>
#include <stdlib.h>

#include "common.cuh"

namespace lb {

constexpr int PF_BM = 128, PF_BN = 64;
constexpr int PF_KST = 4, PF_VST = 4;                 // K / V ring depth (64-key tiles)
constexpr int PF_WARP_TMA = 8, PF_WARP_MMA = 9, PF_THREADS = 320;
constexpr float PF_LOG2E = 1.4426950408889634f;

struct AttnPairParams {
    const uint8_t* qflag;        // [B*T] or null
    const int32_t* work;         // [n_work][4] = {b, q_tile of lane A, variant, q_tile of lane B or -1}
    const int32_t* kv_start;     // [B] or null
    const int32_t* kv_end;       // [B] or null
    const int32_t* out_row;      // [B*T] or null
    __nv_bfloat16* O;
    float* lse;                  // [B,H,T]
    int batch, seqlen, heads;
    int n_work, head_group;
    float scale;
    long long* trace;            // optional [64][16] clock64 stamps of CTA 0 (diagnostics, lb_attn_fwd_pair_set_trace)
    long long* cta_log;          // optional [n_cta][8]: smid, steps A, steps B, clock64 at entry / Q landed / last MMA issued / exit
};

#define PF_TRACE(slot, it)                                                                                  \
    do {                                                                                                    \
        if (p.trace && blockIdx.x == 0 && (it) < 64) p.trace[(it) * 16 + (slot)] = clock64();               \
    } while (0)

template <int D>
struct PairSmem {
    static constexpr int QTILE = PF_BM * D * 2;                       // one Q tile
    static constexpr int KVTILE = PF_BN * D * 2;                      // one K or V tile (64 keys)
    static constexpr int Q_OFF = 0, K_OFF = 2 * QTILE, V_OFF = K_OFF + PF_KST * KVTILE, BAR_OFF = V_OFF + PF_VST * KVTILE;
    static constexpr int NEEDED = BAR_OFF + 512 + 1024;
    static constexpr int TOTAL = NEEDED > 120 * 1024 ? NEEDED : 120 * 1024;     // > half an SM: exactly one CTA per SM (512 TMEM columns)
};

// barrier indices: S/P barriers are [lane][buffer]
enum {
    PB_Q = 0,
    PB_KFULL = 1,
    PB_KEMPTY = PB_KFULL + PF_KST,
    PB_VFULL = PB_KEMPTY + PF_KST,
    PB_VEMPTY = PB_VFULL + PF_VST,
    PB_SFULL = PB_VEMPTY + PF_VST,
    PB_PFULL = PB_SFULL + 4,
    PB_OREADY = PB_PFULL + 4,
    PB_COUNT = PB_OREADY + 2
};

// scores of one row (64 columns at TMEM `ts`) -> registers, optionally masked; returns the row maximum
template <bool MASK, bool CAUSAL>
__device__ __forceinline__ float pair_load_max(uint32_t ts, uint32_t (&v)[64], int kv0, int qi, int kvs, int kve) {
    tmem_ld32(ts, v);
    tmem_ld32(ts + 32, v + 32);
    tc_wait_ld();
    float mx0 = -CUDART_INF_F, mx1 = -CUDART_INF_F, mx2 = -CUDART_INF_F, mx3 = -CUDART_INF_F;
#pragma unroll
    for (int j = 0; j < 64; j += 4) {
        if (MASK) {
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const int kj = kv0 + j + e;
                const bool ok = (!CAUSAL || kj <= qi) && kj < kve && kj >= kvs;
                v[j + e] = ok ? v[j + e] : 0xff800000u;          // -inf
            }
        }
        mx0 = fmaxf(mx0, __uint_as_float(v[j]));
        mx1 = fmaxf(mx1, __uint_as_float(v[j + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(v[j + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(v[j + 3]));
    }
    return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
}

// P = 2^(S*sl2 - m_off), packed to bf16 over the first 32 columns of the row's S buffer; one exponential in every POLY
// runs on the FMA pipes (poly_ex2), 0 = all on the SFU.  Returns the row sum.
template <int POLY>
__device__ __forceinline__ float pair_exp_store(uint32_t ts, uint32_t (&v)[64], float sl2, float m_off) {
    float l0 = 0.f, l1 = 0.f;
#pragma unroll
    for (int j = 0; j < 64; j += 2) {
        const float x0 = fmaf(__uint_as_float(v[j]), sl2, -m_off), x1 = fmaf(__uint_as_float(v[j + 1]), sl2, -m_off);
        const float p0 = (POLY > 0 && (j % POLY) == 0) ? poly_ex2(x0) : fast_ex2(x0);
        const float p1 = (POLY > 0 && ((j + 1) % POLY) == 0) ? poly_ex2(x1) : fast_ex2(x1);
        l0 += p0;
        l1 += p1;
        v[j >> 1] = pack_bf16(p0, p1);
    }
    tmem_st32(ts, v);
    return l0 + l1;
}

template <int D, bool CAUSAL, int POLY>
__global__ void __launch_bounds__(PF_THREADS, 1)
attn_fwd_pair_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK0,
                     const __grid_constant__ CUtensorMap tmV0, const __grid_constant__ CUtensorMap tmK1,
                     const __grid_constant__ CUtensorMap tmV1, const AttnPairParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using S = PairSmem<D>;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + PB_COUNT);

    const int warp = threadIdx.x >> 5;
    const long long t_entry = p.cta_log ? clock64() : 0;
    int item, h;
    attn_cta_order(p.n_work, p.heads, p.head_group, item, h);
    const int b = p.work[item * 4 + 0];
    const int variant = p.work[item * 4 + 2];
    const int qt[2] = {p.work[item * 4 + 1], p.work[item * 4 + 3]};
    const int T = p.seqlen;
    const int kvs = p.kv_start ? p.kv_start[b] : 0;
    const int kve = p.kv_end ? p.kv_end[b] : T;
    const int first_step = kvs / PF_BN;
    const int end_step = (kve + PF_BN - 1) / PF_BN;                   // exclusive
    int ns[2];                                                        // 64-key steps per lane
#pragma unroll
    for (int L = 0; L < 2; ++L) {
        int last = end_step;
        if (CAUSAL && last > 2 * (qt[L] + 1)) last = 2 * (qt[L] + 1);
        ns[L] = (qt[L] >= 0 && last > first_step) ? last - first_step : 0;
    }
    const int n_max = ns[0] > ns[1] ? ns[0] : ns[1];

    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t LANE_COLS = 256, COL_O = 128;

    if (threadIdx.x == 0) {
        for (int i = 0; i < PB_COUNT; ++i) mbar_init(bars + i, (i >= PB_PFULL && i < PB_PFULL + 4) ? 128 : 1);
        fence_barrier_init();
    }
    if (warp == PF_WARP_TMA && elect_one()) {
        tma_prefetch_desc(&tmQ);
        tma_prefetch_desc(variant ? &tmK1 : &tmK0);
        tma_prefetch_desc(variant ? &tmV1 : &tmV0);
    }
    if (warp == PF_WARP_MMA) {
        tmem_alloc(tmem_slot, TMEM_COLS);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == PF_WARP_TMA) {
        // ------------------------------------------------------------ TMA producer
        if (elect_one() && n_max > 0) {
            const CUtensorMap* tK = variant ? &tmK1 : &tmK0;
            const CUtensorMap* tV = variant ? &tmV1 : &tmV0;
            mbar_arrive_expect_tx(bars + PB_Q, (uint32_t)S::QTILE * ((ns[0] > 0) + (ns[1] > 0)));
#pragma unroll
            for (int L = 0; L < 2; ++L) {
                if (ns[L] > 0) {
#pragma unroll
                    for (int c = 0; c < D / 64; ++c)
                        tma_load_2d(smem + S::Q_OFF + L * S::QTILE + c * (PF_BM * 128), &tmQ, bars + PB_Q, h * D + c * 64,
                                    b * T + qt[L] * PF_BM);
                }
            }
            auto load_k = [&](int j) {
                const int s = j % PF_KST;
                mbar_wait(bars + PB_KEMPTY + s, ((uint32_t)(j / PF_KST) & 1u) ^ 1u);
                mbar_arrive_expect_tx(bars + PB_KFULL + s, S::KVTILE);
#pragma unroll
                for (int c = 0; c < D / 64; ++c)
                    tma_load_2d(smem + S::K_OFF + s * S::KVTILE + c * (PF_BN * 128), tK, bars + PB_KFULL + s, h * D + c * 64,
                                b * T + (first_step + j) * PF_BN);
            };
            auto load_v = [&](int j) {
                const int s = j % PF_VST;
                mbar_wait(bars + PB_VEMPTY + s, ((uint32_t)(j / PF_VST) & 1u) ^ 1u);
                mbar_arrive_expect_tx(bars + PB_VFULL + s, S::KVTILE);
#pragma unroll
                for (int c = 0; c < D / 64; ++c)
                    tma_load_2d(smem + S::V_OFF + s * S::KVTILE + c * (PF_BN * 128), tV, bars + PB_VFULL + s, h * D + c * 64,
                                b * T + (first_step + j) * PF_BN);
            };
            load_k(0);
            if (n_max > 1) load_k(1);
            for (int j = 0; j < n_max; ++j) {
                load_v(j);
                if (j + 2 < n_max) load_k(j + 2);
            }
        }
    } else if (warp == PF_WARP_MMA) {
        // ------------------------------------------------------------ MMA issuer (both lanes, two steps ahead of the softmax)
        if (elect_one() && n_max > 0) {
            constexpr uint32_t idesc_qk = make_idesc_bf16(PF_BM, PF_BN, 0, 0);
            constexpr uint32_t idesc_pv = make_idesc_bf16(PF_BM, D, 0, 1);
            const uint32_t dQ[2] = {desc_lo_kmajor(smem_u32(smem + S::Q_OFF)), desc_lo_kmajor(smem_u32(smem + S::Q_OFF + S::QTILE))};
            const uint32_t dK0 = desc_lo_kmajor(smem_u32(smem + S::K_OFF));
            const uint32_t dV0 = desc_lo_mnmajor(smem_u32(smem + S::V_OFF), PF_BN * 128);
            auto issue_qk = [&](int L, int j) {                       // S_L[j&1] = Q_L . K(j)^T
                const uint32_t dK = dK0 + (uint32_t)((j % PF_KST) * (S::KVTILE >> 4));
                const uint32_t d_s = tmem_base + (uint32_t)L * LANE_COLS + (uint32_t)(j & 1) * 64;
#pragma unroll
                for (int kk = 0; kk < D / 16; ++kk) {
                    const uint32_t offq = ((uint32_t)(kk / 4) * (PF_BM * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                    const uint32_t offk = ((uint32_t)(kk / 4) * (PF_BN * 128) + (uint32_t)(kk % 4) * 32) >> 4;
                    umma_ss_lo(d_s, dQ[L] + offq, dK + offk, idesc_qk, kk ? 1u : 0u);
                }
                tc_commit(bars + PB_SFULL + L * 2 + (j & 1));
            };
            auto issue_pv = [&](int L, int j) {                       // O_L += P_L(j) . V(j)
                const uint32_t dV = dV0 + (uint32_t)((j % PF_VST) * (S::KVTILE >> 4));
                const uint32_t a_p = tmem_base + (uint32_t)L * LANE_COLS + (uint32_t)(j & 1) * 64;
#pragma unroll
                for (int kk = 0; kk < PF_BN / 16; ++kk) {
                    // A = P in TMEM: keys 16kk.. at column 8kk of the S buffer; B = V as MN-major (16 key rows = 2048 B)
                    umma_ts_lo(tmem_base + (uint32_t)L * LANE_COLS + COL_O, a_p + (uint32_t)kk * 8, dV + (uint32_t)kk * (2048 >> 4),
                               idesc_pv, (j | kk) ? 1u : 0u);
                }
                tc_commit(bars + PB_OREADY + L);
            };
            mbar_wait(bars + PB_Q, 0);
            if (p.cta_log) p.cta_log[(int64_t)blockIdx.x * 8 + 4] = clock64();
            for (int j = 0; j < 2 && j < n_max; ++j) {                // prologue: S(0), S(1) of both lanes
                mbar_wait(bars + PB_KFULL + j, 0);
                tc_fence_after_sync();
                if (j < ns[0]) issue_qk(0, j);
                if (j < ns[1]) issue_qk(1, j);
                tc_commit(bars + PB_KEMPTY + j);
            }
            for (int j = 0; j < n_max; ++j) {
                const uint32_t ph_p = (uint32_t)(j >> 1) & 1u;                       // S/P buffers: every other step
                const int sv = j % PF_VST, sk = (j + 2) % PF_KST;
                bool v_seen = false, k_seen = false;
#pragma unroll
                for (int L = 0; L < 2; ++L) {
                    if (j < ns[L]) {
                        mbar_wait(bars + PB_PFULL + L * 2 + (j & 1), ph_p);
                        PF_TRACE(2 * L, j);                                  // MMA: P of lane L seen
                        if (!v_seen) {
                            mbar_wait(bars + PB_VFULL + sv, (uint32_t)(j / PF_VST) & 1u);
                            v_seen = true;
                        }
                        tc_fence_after_sync();
                        issue_pv(L, j);
                        if (j + 2 < ns[L]) {
                            if (!k_seen) {
                                mbar_wait(bars + PB_KFULL + sk, (uint32_t)((j + 2) / PF_KST) & 1u);
                                tc_fence_after_sync();
                                k_seen = true;
                            }
                            issue_qk(L, j + 2);
                        }
                        PF_TRACE(2 * L + 1, j);                              // MMA: PV (+ QK two steps ahead) of lane L issued
                    }
                }
                tc_commit(bars + PB_VEMPTY + sv);
                if (j + 2 < n_max) tc_commit(bars + PB_KEMPTY + sk);
            }
            if (p.cta_log) p.cta_log[(int64_t)blockIdx.x * 8 + 5] = clock64();
        }
    } else {
        // ------------------------------------------------------------ softmax / correction / epilogue (one lane per warpgroup)
        const int L = warp >> 2;
        const int r = (warp & 3) * 32 + (threadIdx.x & 31);       // query row in tile == TMEM lane
        const int q0 = qt[L] * PF_BM;
        const int qi = q0 + r;
        const int n_steps = ns[L];
        const uint32_t lane_addr = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)L * LANE_COLS;
        const float sl2 = p.scale * PF_LOG2E;
        float m_used = -CUDART_INF_F, l = 0.f;
        for (int j = 0; j < n_steps; ++j) {
            const uint32_t ph = (uint32_t)(j >> 1) & 1u;
            const uint32_t colS = (uint32_t)(j & 1) * 64;
            const int kv0 = (first_step + j) * PF_BN;
            const bool need_mask = (CAUSAL && kv0 + PF_BN - 1 > q0) || (kv0 + PF_BN > kve) || (kv0 < kvs);
            mbar_wait(bars + PB_SFULL + L * 2 + (j & 1), ph);
            tc_fence_after_sync();
            if ((threadIdx.x & 127) == 0) PF_TRACE(4 + 3 * L, j);            // softmax: S seen
            uint32_t sv[64];
            const float mx = need_mask ? pair_load_max<true, CAUSAL>(lane_addr + colS, sv, kv0, qi, kvs, kve)
                                       : pair_load_max<false, CAUSAL>(lane_addr + colS, sv, kv0, qi, kvs, kve);
            const float m_new = fmaxf(m_used, mx);
            if ((threadIdx.x & 127) == 0) PF_TRACE(5 + 3 * L, j);            // softmax: scores loaded, max done
            // lazy correction: rescale O only when the running max moved by more than 2^8
            const bool grow = (m_new - m_used) * sl2 > 8.f;      // also true when m_used == -inf and m_new finite
            if (j == 0) {
                m_used = m_new;
            } else if (__any_sync(0xffffffffu, grow)) {
                mbar_wait(bars + PB_OREADY + L, (uint32_t)(j - 1) & 1u);     // PV of the previous step has landed in O
                tc_fence_after_sync();
                const float alpha = grow ? ((m_used == -CUDART_INF_F) ? 0.f : fast_ex2((m_used - m_new) * sl2)) : 1.f;
                if (grow) {
                    m_used = m_new;
                    l *= alpha;
                }
#pragma unroll 1
                for (int c = 0; c < D / 16; ++c) {
                    uint32_t v[16];
                    tmem_ld16(lane_addr + COL_O + c * 16, v);
                    tc_wait_ld();
#pragma unroll
                    for (int e = 0; e < 16; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
                    tmem_st16(lane_addr + COL_O + c * 16, v);
                }
                tc_wait_st();
            }
            const float m_off = (m_used == -CUDART_INF_F) ? 0.f : m_used * sl2;
            l += pair_exp_store<POLY>(lane_addr + colS, sv, sl2, m_off);
            tc_wait_st();
            tc_fence_before_sync();
            mbar_arrive(bars + PB_PFULL + L * 2 + (j & 1));
            if ((threadIdx.x & 127) == 0) PF_TRACE(6 + 3 * L, j);            // softmax: P stored, arrived
        }
        // ---- epilogue: normalise and write the rows of this variant
        if (qt[L] >= 0) {
            const int64_t bt = (int64_t)b * T + qi;
            const bool row_ok = (qi < T) && (!p.qflag || (int)p.qflag[qi < T ? bt : 0] == variant);
            if (n_steps > 0) {
                mbar_wait(bars + PB_OREADY + L, (uint32_t)(n_steps - 1) & 1u);
                tc_fence_after_sync();
            }
            const float inv_l = l > 0.f ? 1.f / l : 0.f;
            const int64_t dst = row_ok ? (p.out_row ? (int64_t)p.out_row[bt] : bt) : 0;
            __nv_bfloat16* orow = p.O + dst * ((int64_t)p.heads * D) + (int64_t)h * D;
#pragma unroll 1
            for (int c = 0; c < D / 32; ++c) {
                uint32_t v[32];
                if (n_steps > 0) {
                    tmem_ld32(lane_addr + COL_O + c * 32, v);
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
            if (row_ok && p.lse) {
                // natural-log LSE of the scaled scores; +inf marks a row with no visible key (P == 0 in backward)
                p.lse[((int64_t)b * p.heads + h) * T + qi] = l > 0.f ? (m_used * p.scale + __logf(l)) : CUDART_INF_F;
            }
        }
        tc_fence_before_sync();
    }
    __syncthreads();
    if (warp == PF_WARP_MMA) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, TMEM_COLS);
    }
    if (p.cta_log && threadIdx.x == 0) {
        unsigned smid;
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        long long* e = p.cta_log + (int64_t)blockIdx.x * 8;
        e[0] = smid; e[1] = ns[0]; e[2] = ns[1]; e[3] = t_entry; e[6] = clock64();
    }
}

template <int D, bool CAUSAL, int POLY>
static int launch_pair_p(const CUtensorMap* tm, const AttnPairParams& p, cudaStream_t st) {
    using S = PairSmem<D>;
    auto kern = attn_fwd_pair_kernel<D, CAUSAL, POLY>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail(LB_ELAUNCH, "attn_fwd_pair: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    kern<<<(unsigned)(p.n_work * p.heads), PF_THREADS, S::TOTAL, st>>>(tm[0], tm[1], tm[2], tm[3], tm[4], p);
    return check_launch("attn_fwd_pair");
}

// fraction of exponentials evaluated on the FMA pipes: 1/POLY (LB_PAIR_EXP_POLY=0|2|3|4 for experiments)
static int pair_poly_mod() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("LB_PAIR_EXP_POLY");
        v = e ? atoi(e) : 0;
        if (v != 0 && v != 2 && v != 3 && v != 4) v = 0;
    }
    return v;
}

template <int D, bool CAUSAL>
static int launch_pair(const CUtensorMap* tm, const AttnPairParams& p, cudaStream_t st) {
    switch (pair_poly_mod()) {
        case 2: return launch_pair_p<D, CAUSAL, 2>(tm, p, st);
        case 3: return launch_pair_p<D, CAUSAL, 3>(tm, p, st);
        case 4: return launch_pair_p<D, CAUSAL, 4>(tm, p, st);
        default: return launch_pair_p<D, CAUSAL, 0>(tm, p, st);
    }
}

}  // namespace lb

using namespace lb;

static long long* g_pair_trace = nullptr;
static long long* g_pair_cta_log = nullptr;

/* diagnostics: every CTA of subsequent lb_attn_fwd_pair launches logs {smid, steps A, steps B, clock64 at entry, Q landed,
 * last MMA issued, exit, -} into `buf` ([n_work*heads][8] int64, device); NULL = off */
extern "C" int lb_attn_fwd_pair_set_cta_log(void* buf) {
    g_pair_cta_log = (long long*)buf;
    return LB_OK;
}

/* diagnostics: clock64 stamps of CTA 0 of subsequent lb_attn_fwd_pair launches go to `buf` ([64][16] int64, device); NULL = off */
extern "C" int lb_attn_fwd_pair_set_trace(void* buf) {
    g_pair_trace = (long long*)buf;
    return LB_OK;
}

extern "C" int lb_attn_fwd_pair(const void* Q, const void* K0, const void* V0, const void* K1, const void* V1,
                                const uint8_t* qflag, const int32_t* work, int n_work, const int32_t* kv_start,
                                const int32_t* kv_end, const int32_t* out_row, void* O, float* lse, int batch, int seqlen,
                                int heads, int head_dim, int causal, float scale, void* stream) {
    LB_REQUIRE(batch > 0 && seqlen > 0 && heads > 0 && n_work >= 0, LB_EINVAL, "attn_fwd_pair: bad shape");
    LB_REQUIRE(head_dim == 64 || head_dim == 128, LB_EINVAL, "attn_fwd_pair: head_dim %d (64 or 128 supported)", head_dim);
    LB_REQUIRE(Q && K0 && V0 && O && work, LB_EINVAL, "attn_fwd_pair: null argument");
    LB_REQUIRE(((uintptr_t)O & 15) == 0, LB_EALIGN, "attn_fwd_pair: O must be 16-byte aligned");
    if (n_work == 0) return LB_OK;
    int rc = require_sm100();
    if (rc) return rc;
    const uint64_t rows = (uint64_t)batch * seqlen, cols = (uint64_t)heads * head_dim;
    CUtensorMap tm[5];
    const void* ptrs[5] = {Q, K0, V0, K1 ? K1 : K0, V1 ? V1 : V0};
    for (int i = 0; i < 5; ++i) {
        rc = make_tmap_bf16_2d(&tm[i], ptrs[i], rows, cols, cols, i == 0 ? PF_BM : PF_BN, 64);    // Q: 128-row boxes, K/V: 64
        if (rc) return rc;
    }
    AttnPairParams p;
    p.qflag = qflag; p.work = work; p.kv_start = kv_start; p.kv_end = kv_end; p.out_row = out_row;
    p.O = (__nv_bfloat16*)O; p.lse = lse; p.batch = batch; p.seqlen = seqlen; p.heads = heads; p.scale = scale;
    p.n_work = n_work; p.head_group = attn_head_group();
    p.trace = g_pair_trace;
    p.cta_log = g_pair_cta_log;
    cudaStream_t st = (cudaStream_t)stream;
    if (head_dim == 128) return causal ? launch_pair<128, true>(tm, p, st) : launch_pair<128, false>(tm, p, st);
    return causal ? launch_pair<64, true>(tm, p, st) : launch_pair<64, false>(tm, p, st);
}

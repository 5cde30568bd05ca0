// Grouped, persistent tcgen05 GEMM for sm_100a -- every dense product of the Libra decoder / ViT / heads.
//
//   C_g[M_g, N_g] = epilogue_g( op(A_g) . op(B_g) )      for up to GG_MAXG problems in ONE launch
//
// Replaces (reference, PyTorch eager -> cuBLAS): nn.Linear / F.linear of LlamaAttention / LlamaMLP
// (libra/models/llama/modeling_llama.py:185-201), the LibraLinear chain F.linear(F.linear(x, A), B)
// (libra/models/libra/modeling_libra.py:192-199), the routed q/k/v/o + bridge projections (:310-319), the SwiGLU product
// (:232-233), CLIP's biased linears + quick_gelu (libra/models/clip/modeling_clip.py:279-282, 371-378) and the heads
// (modeling_libra.py:1018-1052), forward and both backward products.
//
// Structure (CG = 2, the default: a CTA pair per 256 x tile_n output tile, tcgen05 cta_group::2):
//   * persistent: one CTA pair per SM pair walks the concatenated tile list of all problems (static round-robin, tiles of a
//     problem in 8-row-block supertiles so the tiles in flight share operands in L2);
//   * warp 0 (one lane): TMA producer.  Each CTA loads its own 128 rows of A and its own half of the B tile; both signal the
//     LEADER's full barrier (cp.async.bulk.tensor ... cta_group::2), GG_STAGES-deep ring of 64-wide K blocks;
//   * warp 1 (one lane, leader CTA only): issues tcgen05.mma.cta_group::2 (UMMA 256 x tile_n x 16), releases ring slots in both
//     CTAs with a multicast commit, and commits each finished accumulator to both CTAs' tmem_full barriers;
//   * TMEM: 2 accumulator buffers x 256 columns, so the epilogue of tile i overlaps the main loop of tile i+1;
//   * warps 2..5: epilogue.  tcgen05.ld (thread = accumulator row) -> bias / activation / SwiGLU / addend -> bf16 ->
//     128B-swizzled staging tile in smem -> TMA store (coalesced, clipped at the tensor edge).  An addend (residual,
//     beta = 1 accumulation into a gradient buffer) arrives through the same staging tile by TMA load;
//   * all four operand layouts through UMMA K-major / MN-major descriptors (dgrad and wgrad need no transposes);
//   * chained problems (LibraLinear: mid = x A^T, y = mid B^T): the second problem's producer waits on per-row-block
//     counters that the first problem's epilogue bumps after its stores completed -- the intermediate never leaves L2 and
//     the chain's small tile sets fill the wave tails of the big dense problems of the same launch.
// CG = 1 is the same kernel on single CTAs (128 x tile_n tiles), kept selectable (LB_GEMM_CG=1) as a cross-check.
#ifndef LB_MBAR_TIMEOUT_CLK
#define LB_MBAR_TIMEOUT_CLK (1ll << 33)
#endif
#include <mutex>
#include <stdlib.h>
#include <string.h>

#include "common.cuh"

namespace lb {

constexpr int GG_BM = 128;              // accumulator rows per CTA (TMEM lanes)
constexpr int GG_BK = 64;               // K elements per ring stage (one 128 B swizzle row)
constexpr int GG_MAXG = 24;             // output problems per launch (after merging accumulation chains)
constexpr int GG_MAXIN = 32;            // entries of the caller's list per launch
constexpr int GG_MAXSEG = 9;            // K segments (A_s . B_s products summed into one accumulator) per problem
constexpr int GG_MAXMAPS = 112;         // tensor maps per launch
constexpr int GG_THREADS = 224;             // warp 0 TMA, 1 MMA, 2-5 epilogue, 6 tile scheduler
constexpr int GG_NSLOT = 4;                // claimed-tile ring between the scheduler and the other roles
constexpr int GG_A_BYTES = GG_BM * GG_BK * 2;          // 16 KB
constexpr int GG_CHUNK_BYTES = GG_BM * 64 * 2;         // epilogue staging tile: 128 rows x 64 bf16
constexpr int GG_SUPER_M = 8;                          // row blocks per supertile

template <int CG>
struct GGCfg {
    static constexpr int B_BYTES = (256 / CG) * GG_BK * 2;            // 32 KB (CG 1) / 16 KB (CG 2)
    static constexpr int STAGE_BYTES = GG_A_BYTES + B_BYTES;
    static constexpr int STAGES = CG == 2 ? 6 : 4;
    static constexpr int BAR_BYTES = 512;
    static constexpr int SMEM = STAGES * STAGE_BYTES + 2 * GG_CHUNK_BYTES + BAR_BYTES + 1024 /*align slack*/;
};

enum { GG_EPI_NONE = 0, GG_EPI_QGELU = 1, GG_EPI_SWIGLU = 2 };

struct GGSeg {
    short a_map, b_map;      // indices into GGParams::maps
    short wait_on;           // problem whose C is this segment's A, or -1
    short wait_all;          // 0: wait for the row block this tile reads; 1: wait for the whole problem (A read transposed)
    int num_kb;
};

struct GGProb {
    int M, N;
    int tile_n;              // accumulator columns per tile
    int tiles_m, tiles_n;
    int tile_begin;          // first global tile index of this problem
    int a_mn, b_mn;          // operands are MN-major (stored transposed); the same for every segment
    int epi;
    int nseg;
    short c_map, aux_d, aux_b2, aux_g, aux_u;    // indices into GGParams::maps or -1
    short signal;            // this problem's epilogue bumps counters[counter_off + row block]
    int counter_off;
    int b_part_rows;         // rows (K-major) or columns (MN-major) of B each CTA loads per part
    int b_parts;             // parts per CTA (2: SwiGLU on CG = 1)
    int b_tx_bytes;          // bytes of B per stage per CTA
    const __nv_bfloat16* bias;
    const float* alpha;      // optional device scalar multiplying the accumulator
    GGSeg seg[GG_MAXSEG];
};

struct GGParams {
    CUtensorMap maps[GG_MAXMAPS];
    GGProb prob[GG_MAXG];
    int n_prob, total_tiles;
    int* counters;
};

struct GGTile {
    int g, mt, nt;
};

__device__ __forceinline__ GGTile gg_decode(const GGParams& p, int t) {
    int g = 0;
#pragma unroll 1
    while (g + 1 < p.n_prob && t >= p.prob[g + 1].tile_begin) ++g;
    const GGProb& pb = p.prob[g];
    const int local = t - pb.tile_begin;
    const int per_group = GG_SUPER_M * pb.tiles_n;
    const int grp = local / per_group;
    const int rem = local - grp * per_group;
    const int gm = min(GG_SUPER_M, pb.tiles_m - grp * GG_SUPER_M);
    GGTile r;
    r.g = g;
    r.nt = rem / gm;
    r.mt = grp * GG_SUPER_M + (rem - r.nt * gm);
    return r;
}

__device__ __forceinline__ void gg_wait(uint32_t bar_addr, uint32_t parity, int tag) {
    uint32_t ok;
    asm volatile(
        "{\n\t"
        ".reg .pred P;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P;\n\t"
        "}\n"
        : "=r"(ok)
        : "r"(bar_addr), "r"(parity)
        : "memory");
    if (ok) return;
    const long long t0 = clock64();
    uint32_t spins = 0;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred P;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(bar_addr), "r"(parity)
            : "memory");
        if (LB_MBAR_TIMEOUT_CLK > 0 && (++spins & 1023u) == 0 && clock64() - t0 > LB_MBAR_TIMEOUT_CLK) {
            printf("libra_b200 gemm_grouped: barrier timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x,
                   threadIdx.x, parity);
            __trap();
        }
    } while (!ok);
}

// wait with cluster-scope acquire: the data guarded by the barrier was written by the partner CTA (st.shared::cluster)
__device__ __forceinline__ void gg_wait_cluster(uint32_t bar_addr, uint32_t parity, int tag) {
    uint32_t ok;
    const long long t0 = clock64();
    uint32_t spins = 0;
    for (;;) {
        asm volatile(
            "{\n\t"
            ".reg .pred P;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P, [%1], %2;\n\t"
            "selp.b32 %0, 1, 0, P;\n\t"
            "}\n"
            : "=r"(ok)
            : "r"(bar_addr), "r"(parity)
            : "memory");
        if (ok) return;
        if (LB_MBAR_TIMEOUT_CLK > 0 && (++spins & 1023u) == 0 && clock64() - t0 > LB_MBAR_TIMEOUT_CLK) {
            printf("libra_b200 gemm_grouped: tile-ring timeout tag=%d block=%d thread=%d parity=%u\n", tag, blockIdx.x, threadIdx.x, parity);
            __trap();
        }
    }
}

// consumer side of the claimed-tile ring: every role of both CTAs sees the same sequence of tile ids, ending with -1
struct GGFeed {
    uint32_t slot = 0, phase = 0;
};
template <int CG>
__device__ __forceinline__ int gg_next_tile(GGFeed& f, uint32_t sfull_a, uint32_t ring_a, uint32_t sempty_leader, bool arrive) {
    gg_wait_cluster(sfull_a + 8 * f.slot, f.phase, 600 + (int)f.slot);
    int t;
    asm volatile("ld.shared.s32 %0, [%1];" : "=r"(t) : "r"(ring_a + 4 * f.slot) : "memory");
    if (arrive) {
        if (CG == 2) mbar_arrive_cluster(sempty_leader + 8 * f.slot);
        else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sempty_leader + 8 * f.slot) : "memory");
    }
    if (++f.slot == GG_NSLOT) { f.slot = 0; f.phase ^= 1u; }
    return t;
}

__device__ __forceinline__ int ld_acquire_gpu(const int* p) {
    int v;
    asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ float silu_f(float x) { return x / (1.f + __expf(-x)); }
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

template <int CG>
__global__ void __launch_bounds__(GG_THREADS, 1) gemm_grouped_kernel(const __grid_constant__ GGParams p) {
    using Cfg = GGCfg<CG>;
    constexpr int STAGES = Cfg::STAGES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem = smem_raw + (smem_base - smem_u32(smem_raw));
    const uint32_t stage_base = smem_base;
    const uint32_t chunk_base = smem_base + STAGES * Cfg::STAGE_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES + 2 * GG_CHUNK_BYTES);
    const uint32_t bar_base = chunk_base + 2 * GG_CHUNK_BYTES;
    // barrier layout (8 B each): full[STAGES] empty[STAGES] tmem_full[2] tmem_empty[2] dbar[2], then the TMEM base slot
    const uint32_t full_a = bar_base, empty_a = bar_base + 8 * STAGES, tfull_a = bar_base + 16 * STAGES,
                   tempty_a = tfull_a + 16, dbar_a = tfull_a + 32;
    // tile ring: sfull[GG_NSLOT] sempty[GG_NSLOT] (8 B each), then the TMEM base slot and GG_NSLOT claimed tile ids
    const uint32_t sfull_a = tfull_a + 48, sempty_a = sfull_a + 8 * GG_NSLOT;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 6 + 2 * GG_NSLOT);
    int* ring = reinterpret_cast<int*>(tmem_slot + 2);
    const uint32_t ring_a = sempty_a + 8 * GG_NSLOT + 8;

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const uint32_t rank = CG == 2 ? cluster_ctarank() : 0u;
    // Tiles are claimed dynamically (one atomic per tile by the leader CTA's scheduler warp) and handed to every role of both
    // CTAs through a small ring: CTAs that are resident take work, CTAs that the hardware could not place yet (SMs busy with a
    // concurrent NCCL kernel during the overlapped gradient reduction) simply find nothing left when they start.

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(bars + s, 1);
            mbar_init(bars + STAGES + s, 1);
        }
        mbar_init(bars + 2 * STAGES + 0, 1);            // tmem_full
        mbar_init(bars + 2 * STAGES + 1, 1);
        mbar_init(bars + 2 * STAGES + 2, 4 * CG);       // tmem_empty: one arrival per epilogue warp of every CTA
        mbar_init(bars + 2 * STAGES + 3, 4 * CG);
        mbar_init(bars + 2 * STAGES + 4, 1);            // dbar (addend tile landed)
        mbar_init(bars + 2 * STAGES + 5, 1);
        for (int i = 0; i < GG_NSLOT; ++i) {
            mbar_init(bars + 2 * STAGES + 6 + i, 1);                             // sfull: the scheduler published a tile id
            mbar_init(bars + 2 * STAGES + 6 + GG_NSLOT + i, CG == 2 ? 11 : 6);   // sempty: every consumer role of the pair read it
        }
        fence_barrier_init();
    }
    if (warp == 1) {
        if (CG == 2) {
            tmem_alloc_cg2(tmem_slot, 512);
            tmem_relinquish_cg2();
        } else {
            tmem_alloc(tmem_slot, 512);
            tmem_relinquish();
        }
    }
    tc_fence_before_sync();
    if (CG == 2) cluster_sync_all();
    __syncthreads();      // (the cluster barrier already orders the CTA; compute-sanitizer's racecheck only models bar.sync)
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t sempty_leader = CG == 2 ? mapa_shared(sempty_a, 0) : sempty_a;
    (void)ring;

    if (warp == 6) {
        // ------------------------------------------------------------------ tile scheduler (leader CTA)
        if (rank == 0 && elect_one()) {
            uint32_t slot = 0, phase = 0;
            const uint32_t ring_peer = CG == 2 ? mapa_shared(ring_a, 1) : 0u;
            const uint32_t sfull_peer = CG == 2 ? mapa_shared(sfull_a, 1) : 0u;
            for (;;) {
                gg_wait(sempty_a + 8 * slot, phase ^ 1u, 700 + (int)slot);
                int t = atomicAdd(p.counters, 1);
                if (t >= p.total_tiles) t = -1;
                asm volatile("st.shared.s32 [%0], %1;" ::"r"(ring_a + 4 * slot), "r"(t) : "memory");
                if (CG == 2) {
                    asm volatile("st.shared::cluster.s32 [%0], %1;" ::"r"(ring_peer + 4 * slot), "r"(t) : "memory");
                    mbar_arrive_cluster(sfull_peer + 8 * slot);
                }
                asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(sfull_a + 8 * slot) : "memory");
                if (t < 0) break;
                if (++slot == GG_NSLOT) { slot = 0; phase ^= 1u; }
            }
        }
    } else if (warp == 0) {
        // ------------------------------------------------------------------ TMA producer
        if (elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t full_leader = CG == 2 ? mapa_shared(full_a, 0) : full_a;
            GGFeed feed;
            for (;;) {
                const int t = gg_next_tile<CG>(feed, sfull_a, ring_a, sempty_leader, true);
                if (t < 0) break;
                const GGTile tl = gg_decode(p, t);
                const GGProb& pb = p.prob[tl.g];
                const int m0 = tl.mt * (GG_BM * CG) + (int)rank * GG_BM;
                const bool dual = pb.epi == GG_EPI_SWIGLU;
                const int n_base = dual ? tl.nt * (pb.tile_n / 2) : tl.nt * pb.tile_n + (int)rank * pb.b_part_rows;
                const uint32_t tx = (uint32_t)(GG_A_BYTES + pb.b_tx_bytes) * CG;
                // per-tile copies of the fields the K loop uses: p.prob[] sits behind a run-time index in the kernel parameters,
                // every use is an indexed constant load that the asm barriers below keep the compiler from hoisting
                const bool a_mn = pb.a_mn != 0, b_mn = pb.b_mn != 0;
                const int b_parts = pb.b_parts, b_part_rows = pb.b_part_rows, nseg = pb.nseg;
                const CUtensorMap* tmB2 = &p.maps[pb.aux_b2 >= 0 ? pb.aux_b2 : 0];
                for (int sg = 0; sg < nseg; ++sg) {
                    const GGSeg& sgm = pb.seg[sg];
                    if (sgm.wait_on >= 0) {
                        // the A operand of this segment is another problem's output: wait until its tiles were published
                        const GGProb& src = p.prob[sgm.wait_on];
                        const int target = src.tiles_n * CG;
                        const int r0 = sgm.wait_all ? 0 : tl.mt, r1 = sgm.wait_all ? src.tiles_m : tl.mt + 1;
                        const long long t0 = clock64();
                        for (int r = r0; r < r1; ++r) {
                            const int* ctr = p.counters + src.counter_off + r;
                            while (ld_acquire_gpu(ctr) < target) {
                                __nanosleep(100);
                                if (LB_MBAR_TIMEOUT_CLK > 0 && clock64() - t0 > LB_MBAR_TIMEOUT_CLK) {
                                    printf("libra_b200 gemm_grouped: dependency timeout problem=%d segment=%d row block=%d\n", tl.g, sg, r);
                                    __trap();
                                }
                            }
                        }
                        fence_proxy_async_all();
                    }
                    const CUtensorMap* tmA = &p.maps[sgm.a_map];
                    const CUtensorMap* tmB0 = &p.maps[sgm.b_map];
                    const int seg_kb = sgm.num_kb;
                    for (int kb = 0; kb < seg_kb; ++kb) {
                        gg_wait(empty_a + 8 * stage, phase ^ 1u, 100 + stage);
                        if (rank == 0) {
                            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(full_a + 8 * stage), "r"(tx)
                                         : "memory");
                        }
                        const uint32_t bar = full_leader + 8 * stage;
                        const uint32_t sa = stage_base + stage * Cfg::STAGE_BYTES;
                        const uint32_t sb = sa + GG_A_BYTES;
                        const int k0 = kb * GG_BK;
                        if (!a_mn) {
                            if (CG == 2) tma_load_2d_cg2(sa, tmA, bar, k0, m0); else tma_load_2d_addr(sa, tmA, bar, k0, m0);
                        } else {
#pragma unroll
                            for (int c = 0; c < GG_BM / 64; ++c) {
                                if (CG == 2) tma_load_2d_cg2(sa + c * (GG_BK * 128), tmA, bar, m0 + c * 64, k0);
                                else tma_load_2d_addr(sa + c * (GG_BK * 128), tmA, bar, m0 + c * 64, k0);
                            }
                        }
                        if (!b_mn) {
                            for (int part = 0; part < b_parts; ++part) {
                                const int which = dual ? (CG == 2 ? (int)rank : part) : 0;
                                const CUtensorMap* tmB = which ? tmB2 : tmB0;
                                const uint32_t dst = sb + part * b_part_rows * 128;
                                if (CG == 2) tma_load_2d_cg2(dst, tmB, bar, k0, n_base); else tma_load_2d_addr(dst, tmB, bar, k0, n_base);
                            }
                        } else {
                            const int nchunk = (b_part_rows + 63) >> 6;
                            for (int c = 0; c < nchunk; ++c) {
                                if (CG == 2) tma_load_2d_cg2(sb + c * (GG_BK * 128), tmB0, bar, n_base + c * 64, k0);
                                else tma_load_2d_addr(sb + c * (GG_BK * 128), tmB0, bar, n_base + c * 64, k0);
                            }
                        }
                        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ------------------------------------------------------------------ MMA issuer (leader CTA)
        if (rank == 0 && elect_one()) {
            int stage = 0;
            uint32_t phase = 0;
            int it = 0;
            GGFeed feed;
            for (;; ++it) {
                const int t = gg_next_tile<CG>(feed, sfull_a, ring_a, sempty_leader, true);
                if (t < 0) break;
                const GGTile tl = gg_decode(p, t);
                const GGProb& pb = p.prob[tl.g];
                const int acc = it & 1;
                gg_wait(tempty_a + 8 * acc, (((uint32_t)it >> 1) & 1u) ^ 1u, 200 + acc);
                tc_fence_after_sync();
                const uint32_t idesc = make_idesc_bf16(GG_BM * CG, pb.tile_n, pb.a_mn, pb.b_mn);
                const uint32_t d_tmem = tmem_base + (uint32_t)acc * 256u;
                const bool a_mn = pb.a_mn != 0, b_mn = pb.b_mn != 0;    // per-tile copies (see the producer)
                const uint32_t a_step = a_mn ? 128u : 2u;             // +2048 B / +32 B per 16 K elements, in 16 B units
                const uint32_t b_step = b_mn ? 128u : 2u;
                int total_kb = 0;
                for (int sg = 0; sg < pb.nseg; ++sg) total_kb += pb.seg[sg].num_kb;
                for (int kb = 0; kb < total_kb; ++kb) {
                    gg_wait(full_a + 8 * stage, phase, 300 + stage);
                    tc_fence_after_sync();
                    const uint32_t sa = stage_base + stage * Cfg::STAGE_BYTES;
                    const uint32_t sb = sa + GG_A_BYTES;
                    const uint32_t a_lo = a_mn ? desc_lo_mnmajor(sa, GG_BK * 128) : desc_lo_kmajor(sa);
                    const uint32_t b_lo = b_mn ? desc_lo_mnmajor(sb, GG_BK * 128) : desc_lo_kmajor(sb);
#pragma unroll
                    for (int k = 0; k < GG_BK / 16; ++k) {
                        const uint32_t accum = (kb | k) != 0 ? 1u : 0u;
                        if (CG == 2) umma_ss_lo_cg2(d_tmem, a_lo + k * a_step, b_lo + k * b_step, idesc, accum);
                        else umma_ss_lo(d_tmem, a_lo + k * a_step, b_lo + k * b_step, idesc, accum);
                    }
                    if (CG == 2) commit_bar_cg2(empty_a + 8 * stage, 3); else commit_bar(empty_a + 8 * stage);
                    if (++stage == STAGES) { stage = 0; phase ^= 1u; }
                }
                if (CG == 2) commit_bar_cg2(tfull_a + 8 * acc, 3); else commit_bar(tfull_a + 8 * acc);
            }
        }
    } else if (warp >= 2 && warp <= 5) {
        // ------------------------------------------------------------------ epilogue (warps 2..5)
        const int q = warp & 3;                         // TMEM lane quadrant this warp may read
        const int row = q * 32 + lane;                  // accumulator row of this thread
        const bool t0 = threadIdx.x == 64;              // the thread that issues this CTA's TMA stores / addend loads
        const uint32_t lane_sel = (uint32_t)(q * 32) << 16;
        const uint32_t tempty_leader = CG == 2 ? mapa_shared(tempty_a, 0) : tempty_a;
        const uint32_t row_off = (uint32_t)row * 128u;
        const uint32_t rsw = (uint32_t)(row & 7);
        uint32_t cc = 0;                                // staging chunks emitted so far (buffer = cc & 1)
        uint32_t dpar0 = 0, dpar1 = 0;                  // phase parity of dbar[0], dbar[1]
        int it = 0;

        // write 64 packed bf16 columns of this thread's row into staging buffer (cc & 1) and TMA-store the tile
        auto emit = [&](const uint32_t* pk, const CUtensorMap* map, int col0, int m0) {
            const uint32_t sbuf = chunk_base + (cc & 1u) * GG_CHUNK_BYTES;
            if (t0) tma_store_wait_read1();             // the store that last read this buffer (two emits ago) is done
            named_bar_sync(1, 128);
#pragma unroll
            for (int j = 0; j < 8; ++j)
                sts128(sbuf + row_off + (((uint32_t)j ^ rsw) << 4), make_uint4(pk[4 * j], pk[4 * j + 1], pk[4 * j + 2], pk[4 * j + 3]));
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (t0) {
                tma_store_2d_addr(map, sbuf, col0, m0);
                tma_store_commit();
            }
            ++cc;
        };

        GGFeed feed;
        for (;; ++it) {
            const int t = gg_next_tile<CG>(feed, sfull_a, ring_a, sempty_leader, false);
            __syncwarp();                               // every lane has read the slot before it is handed back
            if (lane == 0) {
                if (CG == 2) mbar_arrive_cluster(sempty_leader + 8 * ((feed.slot + GG_NSLOT - 1) % GG_NSLOT));
                else asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(sempty_leader + 8 * ((feed.slot + GG_NSLOT - 1) % GG_NSLOT)) : "memory");
            }
            if (t < 0) break;
            const GGTile tl = gg_decode(p, t);
            const GGProb& pb = p.prob[tl.g];
            const int acc = it & 1;
            const int m0 = tl.mt * (GG_BM * CG) + (int)rank * GG_BM;
            const uint32_t tacc = tmem_base + (uint32_t)acc * 256u + lane_sel;
            const bool has_d = pb.aux_d >= 0;
            const CUtensorMap* tmC = &p.maps[pb.c_map];
            // per-tile copies of the problem's fields (indexed constant loads otherwise, re-issued after every asm barrier)
            const int epi = pb.epi, pN = pb.N, tile_n = pb.tile_n;
            const float* const alpha_p = pb.alpha;
            const __nv_bfloat16* const bias_p = pb.bias;
            const CUtensorMap* const tmD = &p.maps[has_d ? pb.aux_d : 0];
            const CUtensorMap* const tmG = pb.aux_g >= 0 ? &p.maps[pb.aux_g] : nullptr;
            const CUtensorMap* const tmU = pb.aux_u >= 0 ? &p.maps[pb.aux_u] : nullptr;
            const bool signal = pb.signal != 0;
            const int counter_off = pb.counter_off;

            if (epi == GG_EPI_SWIGLU) {
                // accumulator columns [0,128) = gate, [128,256) = up, of output columns n0 .. n0+127
                const int n0 = tl.nt * 128;
                gg_wait(tfull_a + 8 * acc, ((uint32_t)it >> 1) & 1u, 400 + acc);
                tc_fence_after_sync();
#pragma unroll 1
                for (int c = 0; c < 2; ++c) {
                    uint32_t pg[32], pu[32], ph[32];
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        uint32_t rg[32], ru[32];
                        tmem_ld32(tacc + c * 64 + h * 32, rg);
                        tmem_ld32(tacc + 128 + c * 64 + h * 32, ru);
                        tc_wait_ld();
#pragma unroll
                        for (int j = 0; j < 32; j += 2) {
                            const float g0 = round_bf16(__uint_as_float(rg[j])), g1 = round_bf16(__uint_as_float(rg[j + 1]));
                            const float u0 = round_bf16(__uint_as_float(ru[j])), u1 = round_bf16(__uint_as_float(ru[j + 1]));
                            // reference: silu(gate) rounded to bf16, then * up (modeling_libra.py:232-233)
                            const float s0 = round_bf16(silu_f(g0)), s1 = round_bf16(silu_f(g1));
                            pg[h * 16 + j / 2] = pack_bf16(g0, g1);
                            pu[h * 16 + j / 2] = pack_bf16(u0, u1);
                            ph[h * 16 + j / 2] = pack_bf16(s0 * u0, s1 * u1);
                        }
                    }
                    if (c == 1) {                       // accumulator drained: hand the buffer back to the MMA issuer
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(tempty_leader + 8 * acc); else mbar_arrive(bars + 2 * STAGES + 2 + acc); }
                    }
                    if (tmG) emit(pg, tmG, n0 + c * 64, m0);
                    if (tmU) emit(pu, tmU, n0 + c * 64, m0);
                    emit(ph, tmC, n0 + c * 64, m0);
                }
            } else {
                const int n0 = tl.nt * tile_n;
                const int nch = (tile_n + 63) >> 6;
                if (has_d && t0) {                      // addend tile of chunk 0 -> staging buffer (cc & 1)
                    tma_store_wait_read1();
                    const uint32_t b = cc & 1u;
                    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dbar_a + 8 * b), "r"(GG_CHUNK_BYTES) : "memory");
                    tma_load_2d_addr(chunk_base + b * GG_CHUNK_BYTES, tmD, dbar_a + 8 * b, n0, m0);
                }
                gg_wait(tfull_a + 8 * acc, ((uint32_t)it >> 1) & 1u, 400 + acc);
                tc_fence_after_sync();
#pragma unroll 1
                for (int c = 0; c < nch; ++c) {
                    uint32_t r[64];
                    tmem_ld32(tacc + c * 64, r);
                    tmem_ld32(tacc + c * 64 + 32, r + 32);
                    tc_wait_ld();
                    if (c == nch - 1) {
                        tc_fence_before_sync();
                        __syncwarp();
                        if (lane == 0) { if (CG == 2) mbar_arrive_cluster(tempty_leader + 8 * acc); else mbar_arrive(bars + 2 * STAGES + 2 + acc); }
                    }
                    const int col0 = n0 + c * 64;
                    if (alpha_p) {
                        const float al = __ldg(alpha_p);
#pragma unroll
                        for (int j = 0; j < 64; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) * al);
                    }
                    if (bias_p) {
                        const uint4* bp = reinterpret_cast<const uint4*>(bias_p + col0);
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            uint4 bv = make_uint4(0, 0, 0, 0);
                            if (col0 + j * 8 < pN) bv = __ldg(bp + j);        // the bias buffer holds N rounded up to 8 elements
                            r[8 * j + 0] = __float_as_uint(__uint_as_float(r[8 * j + 0]) + bf16_lo(bv.x));
                            r[8 * j + 1] = __float_as_uint(__uint_as_float(r[8 * j + 1]) + bf16_hi(bv.x));
                            r[8 * j + 2] = __float_as_uint(__uint_as_float(r[8 * j + 2]) + bf16_lo(bv.y));
                            r[8 * j + 3] = __float_as_uint(__uint_as_float(r[8 * j + 3]) + bf16_hi(bv.y));
                            r[8 * j + 4] = __float_as_uint(__uint_as_float(r[8 * j + 4]) + bf16_lo(bv.z));
                            r[8 * j + 5] = __float_as_uint(__uint_as_float(r[8 * j + 5]) + bf16_hi(bv.z));
                            r[8 * j + 6] = __float_as_uint(__uint_as_float(r[8 * j + 6]) + bf16_lo(bv.w));
                            r[8 * j + 7] = __float_as_uint(__uint_as_float(r[8 * j + 7]) + bf16_hi(bv.w));
                        }
                    }
                    uint32_t pk[32];
#pragma unroll
                    for (int j = 0; j < 32; ++j) pk[j] = pack_bf16(__uint_as_float(r[2 * j]), __uint_as_float(r[2 * j + 1]));
                    if (epi == GG_EPI_QGELU) {
                        // pre-activation (rounded to bf16, what nn.Linear returns) is kept for backward when asked for
                        if (tmG) emit(pk, tmG, col0, m0);
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            const float x0 = bf16_lo(pk[j]), x1 = bf16_hi(pk[j]);
                            pk[j] = pack_bf16(x0 / (1.f + __expf(-1.702f * x0)), x1 / (1.f + __expf(-1.702f * x1)));
                        }
                    }
                    if (!has_d) {
                        emit(pk, tmC, col0, m0);
                    } else {
                        const uint32_t b = cc & 1u;
                        const uint32_t sbuf = chunk_base + b * GG_CHUNK_BYTES;
                        gg_wait(dbar_a + 8 * b, b ? dpar1 : dpar0, 500 + (int)b);
                        if (b) dpar1 ^= 1u; else dpar0 ^= 1u;
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            const uint32_t addr = sbuf + row_off + (((uint32_t)j ^ rsw) << 4);
                            const uint4 d = lds128(addr);
                            uint4 o;
                            // bf16(bf16(x W^T) + addend): the rounding sequence of `residual + linear(x)` in eager PyTorch
                            o.x = pack_bf16(bf16_lo(pk[4 * j + 0]) + bf16_lo(d.x), bf16_hi(pk[4 * j + 0]) + bf16_hi(d.x));
                            o.y = pack_bf16(bf16_lo(pk[4 * j + 1]) + bf16_lo(d.y), bf16_hi(pk[4 * j + 1]) + bf16_hi(d.y));
                            o.z = pack_bf16(bf16_lo(pk[4 * j + 2]) + bf16_lo(d.z), bf16_hi(pk[4 * j + 2]) + bf16_hi(d.z));
                            o.w = pack_bf16(bf16_lo(pk[4 * j + 3]) + bf16_lo(d.w), bf16_hi(pk[4 * j + 3]) + bf16_hi(d.w));
                            sts128(addr, o);
                        }
                        fence_proxy_async_smem();
                        named_bar_sync(1, 128);
                        if (t0) {
                            tma_store_2d_addr(tmC, sbuf, col0, m0);
                            tma_store_commit();
                            if (c + 1 < nch) {          // prefetch the next chunk's addend into the other buffer
                                tma_store_wait_read1();
                                const uint32_t nb = b ^ 1u;
                                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dbar_a + 8 * nb), "r"(GG_CHUNK_BYTES) : "memory");
                                tma_load_2d_addr(chunk_base + nb * GG_CHUNK_BYTES, tmD, dbar_a + 8 * nb, col0 + 64, m0);
                            }
                        }
                        ++cc;
                    }
                }
            }
            if (signal) {
                // publish this CTA's part of the tile: stores complete -> visible device-wide -> counter bump
                if (t0) {
                    tma_store_wait_all();
                    fence_proxy_async_all();
                    __threadfence();
                    atomicAdd(p.counters + counter_off + tl.mt, 1);
                }
            }
        }
        if (t0) tma_store_wait_all();
    }

    // ---------------------------------------------------------------------- teardown
    tc_fence_before_sync();
    if (CG == 2) cluster_sync_all(); else __syncthreads();
    if (warp == 1) {
        tc_fence_after_sync();
        if (CG == 2) tmem_dealloc_cg2(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
    }
}

static int cached_tmap_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows,
                          uint32_t box_cols) {
    return make_tmap_bf16_2d(out, base, rows, cols, ld, box_rows, box_cols);      // cached in host.cu
}

static int gemm_cg() {
    static int cg = 0;
    if (cg == 0) {
        const char* e = getenv("LB_GEMM_CG");
        cg = (e && atoi(e) == 1) ? 1 : 2;
    }
    return cg;
}

template <int CG>
static int launch_grouped(const GGParams& P, cudaStream_t st) {
    using Cfg = GGCfg<CG>;
    auto kern = gemm_grouped_kernel<CG>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM);
        if (e != cudaSuccess) return fail(LB_ELAUNCH, "gemm_grouped: cudaFuncSetAttribute(%d B): %s", Cfg::SMEM, cudaGetErrorString(e));
        configured = true;
    }
    const int units = sm_count() / CG;
    const int n = P.total_tiles < units ? P.total_tiles : units;
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)(n * CG));
    cfg.blockDim = dim3(GG_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = CG;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, P);
    if (e != cudaSuccess) return fail(LB_ELAUNCH, "gemm_grouped: launch: %s", cudaGetErrorString(e));
    return check_launch("gemm_grouped");
}

}  // namespace lb

using namespace lb;

extern "C" int lb_gemm_grouped_workspace_bytes(const lb_gemm_problem* probs, int n) {
    int64_t ctr = 1;                                        // the tile counter
    const int cg = gemm_cg();
    for (int i = 0; i < n; ++i) {
        bool signals = false;
        for (int j = 0; j < n; ++j) signals |= probs[j].wait_on == i;
        if (signals) ctr += ceil_div(probs[i].M, GG_BM * cg);
    }
    return (int)(ctr * 4);
}

extern "C" int lb_gemm_tmap_cache_stats(int64_t* hits, int64_t* misses) { return tmap_cache_stats(hits, misses); }

extern "C" int lb_gemm_grouped(const lb_gemm_problem* probs, int n, void* workspace, int64_t workspace_bytes, void* stream) {
    LB_REQUIRE(probs && n > 0 && n <= GG_MAXIN, LB_EINVAL, "gemm_grouped: 1..%d entries per launch, got %d", GG_MAXIN, n);
    int rc = require_sm100();
    if (rc) return rc;
    const int cg = gemm_cg();
    static thread_local GGParams P;
    P.n_prob = 0;
    P.total_tiles = 0;
    P.counters = (int*)workspace;
    int n_maps = 0, ctr_off = 1, n_live = 0;                 // counters[0]: the launch's tile counter
    int remap[GG_MAXIN];
    auto add_map = [&](const void* ptr, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t brow, short* idx) -> int {
        if (n_maps >= GG_MAXMAPS) return fail(LB_EINVAL, "gemm_grouped: more than %d tensor maps in one launch", GG_MAXMAPS);
        int r = cached_tmap_2d(&P.maps[n_maps], ptr, rows, cols, ld, brow, 64);
        if (r) return r;
        *idx = (short)n_maps++;
        return LB_OK;
    };
    for (int i = 0; i < n; ++i) {
        const lb_gemm_problem& q = probs[i];
        remap[i] = -1;
        LB_REQUIRE(q.M >= 0 && q.N >= 0 && q.K >= 0, LB_EINVAL, "gemm_grouped[%d]: bad shape M=%lld N=%lld K=%lld", i,
                   (long long)q.M, (long long)q.N, (long long)q.K);
        const bool chained = (q.flags & LB_GEMM_ACCUMULATE_PREV) != 0;
        if (q.M == 0 || q.N == 0) continue;             // empty modality segment: nothing to do
        LB_REQUIRE(q.A && q.B, LB_EINVAL, "gemm_grouped[%d]: null operand", i);
        LB_REQUIRE(q.lda % 8 == 0 && q.ldb % 8 == 0, LB_EALIGN, "gemm_grouped[%d]: lda=%lld ldb=%lld must be multiples of 8 elements",
                   i, (long long)q.lda, (long long)q.ldb);
        LB_REQUIRE(q.M < (1ll << 31) && q.N < (1ll << 31) && q.K < (1ll << 31), LB_EINVAL, "gemm_grouped[%d]: dimension too large", i);
        if (chained) {
            // another A_s . B_s product summed into the accumulator of the previous entry's problem
            LB_REQUIRE(i > 0 && remap[i - 1] >= 0, LB_EINVAL, "gemm_grouped[%d]: ACCUMULATE_PREV needs a preceding non-empty entry", i);
            GGProb& g = P.prob[remap[i - 1]];
            LB_REQUIRE(g.M == q.M && g.N == q.N && g.a_mn == (q.trans_a ? 1 : 0) && g.b_mn == (q.trans_b ? 1 : 0) && q.K > 0,
                       LB_EINVAL, "gemm_grouped[%d]: ACCUMULATE_PREV entries must share M, N and operand layouts", i);
            LB_REQUIRE(g.nseg < GG_MAXSEG && g.epi != GG_EPI_SWIGLU, LB_EINVAL, "gemm_grouped[%d]: too many accumulation segments", i);
            remap[i] = remap[i - 1];
        } else {
            LB_REQUIRE(n_live < GG_MAXG, LB_EINVAL, "gemm_grouped: more than %d output problems", GG_MAXG);
            LB_REQUIRE(q.K > 0 && q.C, LB_EINVAL, "gemm_grouped[%d]: K = 0 or null C", i);
            LB_REQUIRE(q.ldc % 8 == 0, LB_EALIGN, "gemm_grouped[%d]: ldc=%lld must be a multiple of 8 elements", i, (long long)q.ldc);
            LB_REQUIRE(q.N % 8 == 0 || q.ldc >= (q.N + 7) / 8 * 8, LB_EALIGN,
                       "gemm_grouped[%d]: N=%lld is not a multiple of 8: the row pitch must cover the padded row (TMA stores whole 16-byte units)",
                       i, (long long)q.N);
            GGProb& g = P.prob[n_live];
            memset(&g, 0, sizeof(g));
            g.M = (int)q.M; g.N = (int)q.N;
            g.a_mn = q.trans_a ? 1 : 0;
            g.b_mn = q.trans_b ? 1 : 0;
            g.epi = q.epilogue;
            g.bias = (const __nv_bfloat16*)q.bias;
            g.alpha = (const float*)q.alpha;
            g.c_map = g.aux_d = g.aux_b2 = g.aux_g = g.aux_u = -1;
            LB_REQUIRE(g.epi == GG_EPI_NONE || g.epi == GG_EPI_QGELU || g.epi == GG_EPI_SWIGLU, LB_EINVAL,
                       "gemm_grouped[%d]: unknown epilogue %d", i, g.epi);
            LB_REQUIRE(!g.bias || ((uintptr_t)q.bias & 15) == 0, LB_EALIGN,
                       "gemm_grouped[%d]: bias must be 16-byte aligned (and hold N rounded up to 8 elements)", i);
            const bool dual = g.epi == GG_EPI_SWIGLU;
            if (dual) {
                LB_REQUIRE(q.B2 && !q.trans_b && !q.D && !q.bias && !q.alpha, LB_EINVAL,
                           "gemm_grouped[%d]: SwiGLU needs B2 (up weight), K-major weights, no addend/bias/alpha", i);
                g.tile_n = 256;
                g.tiles_n = ceil_div(q.N, 128);
            } else if (q.N <= 48) {
                g.tile_n = (int)((q.N + 15) / 16 * 16);
                if (cg == 2 && g.tile_n % 32) g.tile_n += 16;     // each CTA of a pair supplies tile_n / 2 rows of B
                g.tiles_n = 1;
            } else {
                g.tile_n = q.N >= 256 ? 256 : (int)((q.N + 63) / 64 * 64);
                g.tiles_n = ceil_div(q.N, g.tile_n);
            }
            g.tiles_m = ceil_div(q.M, GG_BM * cg);
            g.tile_begin = P.total_tiles;
            P.total_tiles += g.tiles_m * g.tiles_n;
            if (dual) {
                g.b_part_rows = 128;
                g.b_parts = cg == 2 ? 1 : 2;
                g.b_tx_bytes = g.b_parts * 128 * 128;
            } else if (!g.b_mn) {
                g.b_part_rows = g.tile_n / cg;
                g.b_parts = 1;
                g.b_tx_bytes = g.b_part_rows * 128;
            } else {
                g.b_part_rows = g.tile_n / cg;
                g.b_parts = 1;
                g.b_tx_bytes = ((g.b_part_rows + 63) / 64) * (GG_BK * 128);
            }
            rc = add_map(q.C, (uint64_t)q.M, (uint64_t)q.N, (uint64_t)q.ldc, GG_BM, &g.c_map);
            if (rc) return rc;
            if (q.D) {
                LB_REQUIRE(q.ldd % 8 == 0, LB_EALIGN, "gemm_grouped[%d]: ldd=%lld must be a multiple of 8", i, (long long)q.ldd);
                rc = add_map(q.D, (uint64_t)q.M, (uint64_t)q.N, (uint64_t)q.ldd, GG_BM, &g.aux_d);
                if (rc) return rc;
            }
            if (dual) {
                rc = add_map(q.B2, (uint64_t)q.N, (uint64_t)q.K, (uint64_t)q.ldb, 128, &g.aux_b2);
                if (rc) return rc;
            }
            if (q.G) {
                LB_REQUIRE(g.epi != GG_EPI_NONE && !q.D, LB_EINVAL, "gemm_grouped[%d]: G output needs an activation epilogue and no addend", i);
                rc = add_map(q.G, (uint64_t)q.M, (uint64_t)q.N, (uint64_t)q.ldc, GG_BM, &g.aux_g);
                if (rc) return rc;
            }
            if (q.U) {
                LB_REQUIRE(dual, LB_EINVAL, "gemm_grouped[%d]: U output is SwiGLU only", i);
                rc = add_map(q.U, (uint64_t)q.M, (uint64_t)q.N, (uint64_t)q.ldc, GG_BM, &g.aux_u);
                if (rc) return rc;
            }
            remap[i] = n_live++;
        }
        // ---- this entry's K segment
        GGProb& g = P.prob[remap[i]];
        GGSeg& sg = g.seg[g.nseg++];
        sg.num_kb = ceil_div(q.K, GG_BK);
        sg.wait_on = -1;
        sg.wait_all = 0;
        if (!g.a_mn) rc = add_map(q.A, (uint64_t)q.M, (uint64_t)q.K, (uint64_t)q.lda, GG_BM, &sg.a_map);
        else         rc = add_map(q.A, (uint64_t)q.K, (uint64_t)q.M, (uint64_t)q.lda, GG_BK, &sg.a_map);
        if (rc) return rc;
        if (!g.b_mn) rc = add_map(q.B, (uint64_t)q.N, (uint64_t)q.K, (uint64_t)q.ldb, (uint32_t)g.b_part_rows, &sg.b_map);
        else         rc = add_map(q.B, (uint64_t)q.K, (uint64_t)q.N, (uint64_t)q.ldb, GG_BK, &sg.b_map);
        if (rc) return rc;
        if (q.wait_on >= 0) {
            LB_REQUIRE(q.wait_on < i && remap[q.wait_on] >= 0 && remap[q.wait_on] != remap[i], LB_EINVAL,
                       "gemm_grouped[%d]: wait_on=%d must name an earlier, non-empty problem", i, q.wait_on);
            GGProb& src = P.prob[remap[q.wait_on]];
            if (!src.signal) {
                src.signal = 1;
                src.counter_off = ctr_off;
                ctr_off += src.tiles_m;
            }
            sg.wait_on = (short)remap[q.wait_on];
            // A read row block by row block (same rows as the producer's C) can start per row block; anything else
            // (A transposed: the contraction runs over the producer's rows) waits for the whole producer
            sg.wait_all = (g.a_mn || src.M != g.M) ? 1 : 0;
        }
    }
    if (n_live == 0) return LB_OK;
    P.n_prob = n_live;
    cudaStream_t st = (cudaStream_t)stream;
    LB_REQUIRE(workspace && workspace_bytes >= (int64_t)ctr_off * 4, LB_EINVAL,
               "gemm_grouped: this launch needs %d bytes of workspace (tile counter + chain counters)", ctr_off * 4);
    cudaError_t e = cudaMemsetAsync(workspace, 0, (size_t)ctr_off * 4, st);
    if (e != cudaSuccess) return fail(LB_ELAUNCH, "gemm_grouped: cudaMemsetAsync: %s", cudaGetErrorString(e));
    return cg == 2 ? launch_grouped<2>(P, st) : launch_grouped<1>(P, st);
}

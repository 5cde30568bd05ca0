// tcgen05 GEMM for sm_100a:  C[M,N] = op(A) . op(B) (+C) (+bias, activation), bf16 in, fp32 accumulate in TMEM.
//
//   warp 4 lane 0 : TMA producer   (cp.async.bulk.tensor, SWIZZLE_128B boxes of 64 contiguous elements)
//   warp 5 lane 0 : MMA issuer     (tcgen05.mma.cta_group::1.kind::f16, UMMA 128 x BN x 16), owns TMEM alloc/free
//   warps 0..3    : epilogue       (tcgen05.ld 32x32b -> registers -> bias/activation -> global)
//
// One 128 x BN output tile per CTA, STAGES-deep smem ring, two CTAs per SM (96 KB smem + 128 TMEM columns each) so
// one CTA's epilogue overlaps the other's main loop.  All four operand layouts are supported through the UMMA
// K-major / MN-major descriptors (common.cuh), which is what the backward GEMMs (dX = dY.W, dW = dY^T.X) need
// without materialising transposes.
#include "common.cuh"

namespace lb {

constexpr int GEMM_BM = 128;
constexpr int GEMM_BK = 64;
constexpr int GEMM_STAGES = 3;
constexpr int GEMM_THREADS = 192;

template <int BN>
struct GemmSmem {
    static constexpr int A_BYTES = GEMM_BM * GEMM_BK * 2;          // 16 KB
    static constexpr int B_BYTES = BN * GEMM_BK * 2;
    static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
    static constexpr int TOTAL = GEMM_STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

struct GemmParams {
    void* C;
    const __nv_bfloat16* bias;
    int64_t M, N, K, ldc;
    int out_f32, accumulate, act;
};

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(GEMM_THREADS, 2)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    using S = GemmSmem<BN>;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + GEMM_STAGES * S::STAGE_BYTES);
    uint64_t* full = bars;
    uint64_t* empty = bars + GEMM_STAGES;
    uint64_t* tmem_full = bars + 2 * GEMM_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * GEMM_STAGES + 1);

    const int warp = threadIdx.x >> 5;
    const int m0 = blockIdx.y * GEMM_BM;
    const int n0 = blockIdx.x * BN;
    const int num_kb = (int)((p.K + GEMM_BK - 1) / GEMM_BK);

    if (threadIdx.x == 0) {
        for (int s = 0; s < GEMM_STAGES; ++s) {
            mbar_init(full + s, 1);
            mbar_init(empty + s, 1);
        }
        mbar_init(tmem_full, 1);
        fence_barrier_init();
    }
    if (warp == 4 && elect_one()) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, BN);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        if (elect_one()) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % GEMM_STAGES;
                const uint32_t ph = (uint32_t)(kb / GEMM_STAGES) & 1u;
                mbar_wait(empty + s, ph ^ 1u);
                mbar_arrive_expect_tx(full + s, S::STAGE_BYTES);
                uint8_t* sa = smem + s * S::STAGE_BYTES;
                uint8_t* sb = sa + S::A_BYTES;
                const int k0 = kb * GEMM_BK;
                if (!A_MN) {
                    tma_load_2d(sa, &tmA, full + s, k0, m0);                 // box {64 k, 128 m}
                } else {
#pragma unroll
                    for (int c = 0; c < GEMM_BM / 64; ++c)                    // boxes {64 m, 64 k}
                        tma_load_2d(sa + c * (GEMM_BK * 128), &tmA, full + s, m0 + c * 64, k0);
                }
                if (!B_MN) {
                    tma_load_2d(sb, &tmB, full + s, k0, n0);                 // box {64 k, BN n}
                } else {
#pragma unroll
                    for (int c = 0; c < BN / 64; ++c)                         // boxes {64 n, 64 k}
                        tma_load_2d(sb + c * (GEMM_BK * 128), &tmB, full + s, n0 + c * 64, k0);
                }
            }
        }
    } else if (warp == 5) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(GEMM_BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % GEMM_STAGES;
                const uint32_t ph = (uint32_t)(kb / GEMM_STAGES) & 1u;
                mbar_wait(full + s, ph);
                tc_fence_after_sync();
                const uint32_t sa = smem_u32(smem + s * S::STAGE_BYTES);
                const uint32_t sb = sa + S::A_BYTES;
                const uint64_t da = A_MN ? desc_mnmajor(sa, GEMM_BK * 128) : desc_kmajor(sa);
                const uint64_t db = B_MN ? desc_mnmajor(sb, GEMM_BK * 128) : desc_kmajor(sb);
#pragma unroll
                for (int k = 0; k < GEMM_BK / 16; ++k) {
                    const uint64_t a = da + (uint64_t)(A_MN ? k * 128 : k * 2);   // +2048 B or +32 B, in 16 B units
                    const uint64_t b = db + (uint64_t)(B_MN ? k * 128 : k * 2);
                    umma_ss(tmem_base, a, b, idesc, (kb | k) != 0 ? 1u : 0u);
                }
                tc_commit(empty + s);
            }
            tc_commit(tmem_full);
        }
    } else {
        // ---------------- epilogue: thread <-> accumulator row (TMEM lane)
        mbar_wait(tmem_full, 0);
        tc_fence_after_sync();
        const int row = m0 + threadIdx.x;
        const uint32_t lane_base = tmem_base + ((uint32_t)(warp * 32) << 16);
        const bool row_ok = row < p.M;
#pragma unroll 1
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t r[32];
            tmem_ld32(lane_base + c0, r);
            tc_wait_ld();
            const int col = n0 + c0;
            if (row_ok && col < p.N) {
            float f[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(r[j]);
            const int ncol = (int)((p.N - col) < 32 ? (p.N - col) : 32);
            if (p.bias) {
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    if (j < ncol) f[j] += __bfloat162float(p.bias[col + j]);
            }
            if (p.act == 1) {
#pragma unroll
                for (int j = 0; j < 32; ++j) f[j] = f[j] / (1.f + __expf(-1.702f * f[j]));
            }
            if (p.out_f32) {
                float* crow = reinterpret_cast<float*>(p.C) + (int64_t)row * p.ldc + col;
                if (ncol == 32 && (p.ldc & 3) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        float4 o = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
                        if (p.accumulate) {
                            const float4 c = *reinterpret_cast<const float4*>(crow + j);
                            o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w;
                        }
                        *reinterpret_cast<float4*>(crow + j) = o;
                    }
                } else {
                    for (int j = 0; j < ncol; ++j) crow[j] = f[j] + (p.accumulate ? crow[j] : 0.f);
                }
            } else {
                __nv_bfloat16* crow = reinterpret_cast<__nv_bfloat16*>(p.C) + (int64_t)row * p.ldc + col;
                if (ncol == 32 && (p.ldc & 7) == 0) {
#pragma unroll
                    for (int j = 0; j < 32; j += 8) {
                        if (p.accumulate) {
                            const uint4 c = *reinterpret_cast<const uint4*>(crow + j);
                            f[j + 0] += bf16_lo(c.x); f[j + 1] += bf16_hi(c.x);
                            f[j + 2] += bf16_lo(c.y); f[j + 3] += bf16_hi(c.y);
                            f[j + 4] += bf16_lo(c.z); f[j + 5] += bf16_hi(c.z);
                            f[j + 6] += bf16_lo(c.w); f[j + 7] += bf16_hi(c.w);
                        }
                        uint4 o;
                        o.x = pack_bf16(f[j + 0], f[j + 1]);
                        o.y = pack_bf16(f[j + 2], f[j + 3]);
                        o.z = pack_bf16(f[j + 4], f[j + 5]);
                        o.w = pack_bf16(f[j + 6], f[j + 7]);
                        *reinterpret_cast<uint4*>(crow + j) = o;
                    }
                } else {
                    for (int j = 0; j < ncol; ++j)
                        crow[j] = __float2bfloat16(f[j] + (p.accumulate ? __bfloat162float(crow[j]) : 0.f));
                }
            }
            }
            __syncwarp();
        }
        tc_fence_before_sync();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, BN);
    }
}

template <int BN, bool A_MN, bool B_MN>
static int launch_gemm(const CUtensorMap& tmA, const CUtensorMap& tmB, const GemmParams& p, cudaStream_t st) {
    using S = GemmSmem<BN>;
    auto kern = gemm_bf16_kernel<BN, A_MN, B_MN>;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, S::TOTAL);
        if (e != cudaSuccess) return fail(LB_ELAUNCH, "gemm: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    dim3 grid((unsigned)ceil_div(p.N, BN), (unsigned)ceil_div(p.M, GEMM_BM));
    kern<<<grid, GEMM_THREADS, S::TOTAL, st>>>(tmA, tmB, p);
    return check_launch("gemm_bf16");
}

}  // namespace lb

using namespace lb;

extern "C" int lb_gemm_bf16(const void* A, const void* B, void* C, const void* bias, int64_t M, int64_t N, int64_t K,
                            int64_t lda, int64_t ldb, int64_t ldc, int trans_a, int trans_b, int out_dtype,
                            int accumulate, int act, void* stream) {
    LB_REQUIRE(M > 0 && N > 0 && K > 0, LB_EINVAL, "gemm: bad shape M=%lld N=%lld K=%lld", (long long)M, (long long)N,
               (long long)K);
    LB_REQUIRE(A && B && C, LB_EINVAL, "gemm: null pointer");
    LB_REQUIRE(out_dtype == LB_DT_BF16 || out_dtype == LB_DT_F32, LB_EDTYPE, "gemm: out dtype %d", out_dtype);
    LB_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, LB_EALIGN, "gemm: lda=%lld ldb=%lld must be multiples of 8", (long long)lda,
               (long long)ldb);
    LB_REQUIRE(((uintptr_t)C & 15) == 0, LB_EALIGN, "gemm: C must be 16-byte aligned");
    int rc = require_sm100();
    if (rc) return rc;
    CUtensorMap tmA, tmB;
    constexpr int BN = 128;
    if (!trans_a) rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)M, (uint64_t)K, (uint64_t)lda, GEMM_BM, 64);
    else          rc = make_tmap_bf16_2d(&tmA, A, (uint64_t)K, (uint64_t)M, (uint64_t)lda, GEMM_BK, 64);
    if (rc) return rc;
    if (!trans_b) rc = make_tmap_bf16_2d(&tmB, B, (uint64_t)N, (uint64_t)K, (uint64_t)ldb, BN, 64);
    else          rc = make_tmap_bf16_2d(&tmB, B, (uint64_t)K, (uint64_t)N, (uint64_t)ldb, GEMM_BK, 64);
    if (rc) return rc;
    GemmParams p;
    p.C = C;
    p.bias = (const __nv_bfloat16*)bias;
    p.M = M; p.N = N; p.K = K; p.ldc = ldc;
    p.out_f32 = out_dtype == LB_DT_F32;
    p.accumulate = accumulate;
    p.act = act;
    cudaStream_t st = (cudaStream_t)stream;
    if (!trans_a && !trans_b) return launch_gemm<BN, false, false>(tmA, tmB, p, st);
    if (!trans_a && trans_b) return launch_gemm<BN, false, true>(tmA, tmB, p, st);
    if (trans_a && !trans_b) return launch_gemm<BN, true, false>(tmA, tmB, p, st);
    return launch_gemm<BN, true, true>(tmA, tmB, p, st);
}

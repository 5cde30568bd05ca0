// A1 -- CLIPVisionEmbeddings (libra/models/clip/modeling_clip.py:193-228) as an im2col-free, TMA-staged tcgen05 tile:
//   emb[b, 0, :]   = class_emb + pos[0]
//   emb[b, 1+p, :] = conv14x14/s14(pixels[b])[p] + pos[1+p]          (no bias)
//
// The 14x14 patch extraction never touches HBM as an im2col matrix.  Per CTA (image b, group of 4 patch rows, 128
// output channels) and per K block (input channel c, 4 kernel rows kh0..kh0+3):
//   1. TMA copies the raw pixel rows {14*(ph0+i)+kh0+j : i,j in 0..3} x full width into shared memory through a 5-D
//      view [b, c, patch_row, kernel_row, w] of the NCHW image (rows past kernel row 13 are out of bounds => zeros);
//   2. the 128 transform threads scatter those pixels into the UMMA K-major SWIZZLE_128B operand tile
//      A[m = patch, k = (kh-kh0)*14 + kw]  (56 of 64 columns used, the rest stay zero);
//   3. TMA brings the matching [128 x 64] block of the K-packed weight; tcgen05.mma accumulates in TMEM.
// The weight is K-packed once by lb_patch_embed_pack_weight: [C, 3*14*14] -> [C, 12*64] (12 = 3 channels x 4 kernel-row
// groups, each group padded from 56 to 64 with zeros).
#include "common.cuh"

namespace lb {

constexpr int PE_THREADS = 192;
constexpr int PE_P = 14;             // patch size
constexpr int PE_KHB = 4;            // kernel rows per K block
constexpr int PE_NKB_C = 4;          // K blocks per input channel (ceil(14/4))
constexpr int PE_NKB = 3 * PE_NKB_C; // 12
constexpr int PE_KP = PE_NKB * 64;   // packed K = 768
constexpr int PE_ROWS = 4;           // patch rows per CTA

__global__ void pack_weight_kernel(const __nv_bfloat16* __restrict__ w, __nv_bfloat16* __restrict__ out, int C) {
    const int64_t total = (int64_t)C * PE_KP;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int n = (int)(i / PE_KP), kp = (int)(i % PE_KP);
        const int kb = kp / 64, kk = kp % 64;
        const int c = kb / PE_NKB_C, kh = (kb % PE_NKB_C) * PE_KHB + kk / PE_P, kw = kk % PE_P;
        __nv_bfloat16 v = __float2bfloat16(0.f);
        if (kk < PE_KHB * PE_P && kh < PE_P) v = w[((int64_t)n * 3 + c) * (PE_P * PE_P) + kh * PE_P + kw];
        out[i] = v;
    }
}

struct PeParams {
    const __nv_bfloat16* class_emb;
    const __nv_bfloat16* pos_emb;
    __nv_bfloat16* emb;
    int batch, S, G, C, n_wbox, wbox;     // image size, grid, channels out, TMA boxes per row, box width
};

struct PeSmem {
    static constexpr int RAW_BYTES = 16 * 512 * 2;       // 16 image rows x <=512 px (S <= 512)
    static constexpr int A_BYTES = 128 * 64 * 2;
    static constexpr int W_BYTES = 128 * 64 * 2;
    static constexpr int STAGE = RAW_BYTES + A_BYTES + W_BYTES;
    static constexpr int BAR_OFF = 2 * STAGE;
    static constexpr int TOTAL = BAR_OFF + 1024 + 128;
};
enum { PE_RAWFULL0 = 0, PE_RAWFULL1, PE_AFULL0, PE_AFULL1, PE_FREE0, PE_FREE1, PE_DONE, PE_NBAR };

__global__ void __launch_bounds__(PE_THREADS, 2)
patch_embed_kernel(const __grid_constant__ CUtensorMap tmPix, const __grid_constant__ CUtensorMap tmW, const PeParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + PeSmem::BAR_OFF);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + PE_NBAR);
    auto sA = [&](int st) { return smem + st * PeSmem::STAGE; };                       // 1024-aligned (STAGE % 1024 == 0)
    auto sW = [&](int st) { return smem + st * PeSmem::STAGE + PeSmem::A_BYTES; };
    auto sR = [&](int st) { return smem + st * PeSmem::STAGE + PeSmem::A_BYTES + PeSmem::W_BYTES; };

    const int warp = threadIdx.x >> 5;
    const int n0 = blockIdx.x * 128;
    const int prg = blockIdx.y;                 // patch-row group
    const int b = blockIdx.z;
    const int G = p.G, S = p.S;
    const int ph0 = prg * PE_ROWS;
    const int n_rows = min(PE_ROWS, G - ph0);   // valid patch rows in this group
    const int m_valid = n_rows * G;

    if (threadIdx.x == 0) {
        for (int i = 0; i < PE_NBAR; ++i) mbar_init(bars + i, (i == PE_AFULL0 || i == PE_AFULL1) ? 128 : 1);
        fence_barrier_init();
    }
    if (warp == 5) {
        tmem_alloc(tmem_slot, 128);
        tmem_relinquish();
    }
    // zero both A stages once: pad rows (m >= m_valid) and pad columns (k >= 56) are never written afterwards
    for (int i = threadIdx.x; i < 2 * PeSmem::A_BYTES / 16; i += PE_THREADS) {
        const int st = i / (PeSmem::A_BYTES / 16), j = i % (PeSmem::A_BYTES / 16);
        reinterpret_cast<uint4*>(sA(st))[j] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;
    const uint32_t raw_tx = (uint32_t)(16 * p.wbox * p.n_wbox * 2);

    if (warp == 4) {
        if (elect_one()) {
            for (int kb = 0; kb < PE_NKB; ++kb) {
                const int st = kb & 1;
                const uint32_t ph = (uint32_t)(kb >> 1) & 1u;
                mbar_wait(bars + PE_FREE0 + st, ph ^ 1u);
                const int c = kb / PE_NKB_C, kh0 = (kb % PE_NKB_C) * PE_KHB;
                mbar_arrive_expect_tx(bars + PE_RAWFULL0 + st, raw_tx + PeSmem::W_BYTES);
                // raw pixels: 5-D view {w, kh, ph, c, b}, box {wbox, 4, 4, 1, 1}
                for (int wb = 0; wb < p.n_wbox; ++wb)
                    asm volatile(
                        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, "
                        "%7}], [%2];" ::"r"(smem_u32(sR(st) + wb * (16 * p.wbox * 2))),
                        "l"(reinterpret_cast<uint64_t>(&tmPix)), "r"(smem_u32(bars + PE_RAWFULL0 + st)), "r"(wb * p.wbox),
                        "r"(kh0), "r"(ph0), "r"(c), "r"(b)
                        : "memory");
                tma_load_2d(sW(st), &tmW, bars + PE_RAWFULL0 + st, kb * 64, n0);
            }
        }
    } else if (warp == 5) {
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
            for (int kb = 0; kb < PE_NKB; ++kb) {
                const int st = kb & 1;
                const uint32_t ph = (uint32_t)(kb >> 1) & 1u;
                mbar_wait(bars + PE_AFULL0 + st, ph);       // implies the weight tile landed (transformers waited RAWFULL)
                tc_fence_after_sync();
                const uint32_t a = smem_u32(sA(st)), w = smem_u32(sW(st));
#pragma unroll
                for (int k = 0; k < 4; ++k) umma_ss(tmem_base, desc_kmajor(a + k * 32), desc_kmajor(w + k * 32), idesc, (kb | k) ? 1u : 0u);
                tc_commit(bars + PE_FREE0 + st);
            }
            tc_commit(bars + PE_DONE);
        }
    } else {
        // ---------------- transform: raw pixel rows -> K-major swizzled A tile
        const int t = threadIdx.x;
        for (int kb = 0; kb < PE_NKB; ++kb) {
            const int st = kb & 1;
            const uint32_t ph = (uint32_t)(kb >> 1) & 1u;
            // the A stage is free once the MMAs of K block kb-2 completed
            if (kb >= 2) mbar_wait(bars + PE_FREE0 + st, ((uint32_t)((kb - 2) >> 1) & 1u));
            mbar_wait(bars + PE_RAWFULL0 + st, ph);
            const __nv_bfloat16* raw = reinterpret_cast<const __nv_bfloat16*>(sR(st));
            uint8_t* A = sA(st);
            const int n_elem = m_valid * (PE_KHB * PE_P);
            for (int e = t; e < n_elem; e += 128) {
                const int m = e / (PE_KHB * PE_P), k = e - m * (PE_KHB * PE_P);
                const int pr = m / G, pw = m - pr * G;
                const int khl = k / PE_P, kw = k - khl * PE_P;
                const int w = pw * PE_P + kw;
                const int wb = w / p.wbox, wl = w - wb * p.wbox;
                const __nv_bfloat16 v = raw[(wb * 16 + pr * 4 + khl) * p.wbox + wl];
                const uint32_t off = (uint32_t)(m >> 3) * 1024u + (uint32_t)(m & 7) * 128u + ((((uint32_t)k >> 3) ^ ((uint32_t)m & 7)) << 4) +
                                     ((uint32_t)k & 7) * 2u;
                *reinterpret_cast<__nv_bfloat16*>(A + off) = v;
            }
            fence_proxy_async_smem();           // generic-proxy writes -> visible to the tensor core (async proxy)
            mbar_arrive(bars + PE_AFULL0 + st);
        }
        // ---------------- epilogue
        mbar_wait(bars + PE_DONE, 0);
        tc_fence_after_sync();
        const int m = t;
        const uint32_t lane_addr = tmem_base + ((uint32_t)(warp * 32) << 16);
        const bool row_ok = m < m_valid;
        const int patch = ph0 * G + m;
        const int64_t out_row = (int64_t)b * (G * G + 1) + 1 + patch;
#pragma unroll 1
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(lane_addr + c0, v);
            tc_wait_ld();
            if (row_ok) {
                const __nv_bfloat16* pe = p.pos_emb + (int64_t)(1 + patch) * p.C + n0 + c0;
                __nv_bfloat16* o = p.emb + out_row * p.C + n0 + c0;
#pragma unroll
                for (int j = 0; j < 32; j += 8) {
                    const uint4 pv = *reinterpret_cast<const uint4*>(pe + j);
                    // conv output is rounded to bf16 by the reference before the position add (two bf16 ops)
                    uint4 ov;
                    ov.x = pack_bf16(bf16_lo(pack_bf16(__uint_as_float(v[j + 0]), 0.f)) + bf16_lo(pv.x),
                                     bf16_lo(pack_bf16(__uint_as_float(v[j + 1]), 0.f)) + bf16_hi(pv.x));
                    ov.y = pack_bf16(bf16_lo(pack_bf16(__uint_as_float(v[j + 2]), 0.f)) + bf16_lo(pv.y),
                                     bf16_lo(pack_bf16(__uint_as_float(v[j + 3]), 0.f)) + bf16_hi(pv.y));
                    ov.z = pack_bf16(bf16_lo(pack_bf16(__uint_as_float(v[j + 4]), 0.f)) + bf16_lo(pv.z),
                                     bf16_lo(pack_bf16(__uint_as_float(v[j + 5]), 0.f)) + bf16_hi(pv.z));
                    ov.w = pack_bf16(bf16_lo(pack_bf16(__uint_as_float(v[j + 6]), 0.f)) + bf16_lo(pv.w),
                                     bf16_lo(pack_bf16(__uint_as_float(v[j + 7]), 0.f)) + bf16_hi(pv.w));
                    *reinterpret_cast<uint4*>(o + j) = ov;
                }
            }
            __syncwarp();
        }
        if (prg == 0) {     // class token row
            const int n = n0 + t;
            p.emb[(int64_t)b * (G * G + 1) * p.C + n] =
                __float2bfloat16(__bfloat162float(p.class_emb[n]) + __bfloat162float(p.pos_emb[n]));
        }
        tc_fence_before_sync();
    }
    __syncthreads();
    if (warp == 5) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, 128);
    }
}

}  // namespace lb

using namespace lb;

extern "C" {

int lb_patch_embed_pack_weight(const void* weight, void* packed, int channels_out, int patch, void* stream) {
    LB_REQUIRE(weight && packed && channels_out > 0, LB_EINVAL, "patch_embed_pack_weight: bad arguments");
    LB_REQUIRE(patch == PE_P, LB_EINVAL, "patch_embed: patch size %d (14 supported)", patch);
    const int64_t total = (int64_t)channels_out * PE_KP;
    pack_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)weight,
                                                                                         (__nv_bfloat16*)packed, channels_out);
    return check_launch("patch_embed_pack_weight");
}

int lb_patch_embed_fwd(const void* pixels, const void* weight_packed, const void* class_emb, const void* pos_emb, void* emb,
                       int batch, int image_size, int patch, int channels_out, void* stream) {
    LB_REQUIRE(pixels && weight_packed && class_emb && pos_emb && emb && batch > 0, LB_EINVAL, "patch_embed: null argument");
    LB_REQUIRE(patch == PE_P, LB_EINVAL, "patch_embed: patch size %d (14 supported)", patch);
    LB_REQUIRE(image_size % patch == 0 && image_size <= 512, LB_EINVAL, "patch_embed: image size %d", image_size);
    LB_REQUIRE(channels_out % 128 == 0, LB_EINVAL, "patch_embed: channels_out %d must be a multiple of 128", channels_out);
    const int G = image_size / patch;
    LB_REQUIRE(G * PE_ROWS <= 128, LB_EINVAL, "patch_embed: grid width %d too large for a 128-row tile", G);
    int rc = require_sm100();
    if (rc) return rc;
    // TMA boxes along the image width: <= 256 elements and a multiple of 16 bytes
    int n_wbox = 0, wbox = 0;
    for (int nb = 1; nb <= 4; ++nb) {
        if (image_size % nb) continue;
        const int wdt = image_size / nb;
        if (wdt <= 256 && (wdt * 2) % 16 == 0) { n_wbox = nb; wbox = wdt; break; }
    }
    LB_REQUIRE(n_wbox > 0, LB_EINVAL, "patch_embed: image width %d cannot be split into TMA boxes", image_size);
    CUtensorMap tmPix, tmW;
    const uint64_t S = (uint64_t)image_size;
    uint64_t dims[5] = {S, (uint64_t)PE_P, (uint64_t)G, 3, (uint64_t)batch};
    uint64_t strides[4] = {S * 2, S * 2 * PE_P, S * S * 2, S * S * 2 * 3};
    uint32_t box[5] = {(uint32_t)wbox, PE_KHB, PE_ROWS, 1, 1};
    rc = make_tmap_bf16_nd(&tmPix, pixels, 5, dims, strides, box, 0);
    if (rc) return rc;
    rc = make_tmap_bf16_2d(&tmW, weight_packed, (uint64_t)channels_out, PE_KP, PE_KP, 128, 64);
    if (rc) return rc;
    static bool configured = false;
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(patch_embed_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, PeSmem::TOTAL);
        if (e != cudaSuccess) return fail(LB_ELAUNCH, "patch_embed: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured = true;
    }
    PeParams p;
    p.class_emb = (const __nv_bfloat16*)class_emb; p.pos_emb = (const __nv_bfloat16*)pos_emb; p.emb = (__nv_bfloat16*)emb;
    p.batch = batch; p.S = image_size; p.G = G; p.C = channels_out; p.n_wbox = n_wbox; p.wbox = wbox;
    dim3 grid((unsigned)(channels_out / 128), (unsigned)((G + PE_ROWS - 1) / PE_ROWS), (unsigned)batch);
    patch_embed_kernel<<<grid, PE_THREADS, PeSmem::TOTAL, (cudaStream_t)stream>>>(tmPix, tmW, p);
    return check_launch("patch_embed_fwd");
}

}

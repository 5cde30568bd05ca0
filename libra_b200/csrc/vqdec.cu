// N2: the vision tokenizer's decode path (ids -> pixels) -- the memory-bound kernels around the tcgen05 GEMM.
//
// Reference: ImageTokenizer.decode (libra/models/libra/image_tokenizer.py:97-124) -> LFQ.indices_to_codes
// (taming/modules/quantization/lookup_free_quantization.py:129-158) -> post_quant_conv -> taming Decoder
// (taming/modules/diffusionmodules/model.py:34-230 Normalize / nonlinearity / Upsample / ResnetBlock / AttnBlock, :474-588 Decoder).
//
// Layout: activations are NHWC bf16.  Everything a 3x3 convolution reads or writes lives in the PADDED ROW layout
//     row(b, y, x) = (b*(H+2) + y+1)*(W+2) + x+1,      rows x C, one zero pixel around every image,
// because there the nine taps of a 3x3 / pad 1 convolution are nine constant ROW SHIFTS of the same matrix:
//     out[m, :] = sum_{dy,dx} in[m + dy*(W+2) + dx, :] . W[dy,dx]^T        for every padded row m,
// i.e. ONE GEMM problem of nine K segments (lb_gemm_grouped, LB_GEMM_ACCUMULATE_PREV) whose A operands are the same buffer at
// nine pointer offsets -- no im2col, no gather, fp32 accumulation over all 9*Cin products like cuDNN's.  Border rows of a
// convolution's OUTPUT hold garbage; the kernels here (group norm, upsample, pad) read interiors only and write the zero
// borders the next convolution needs.  1x1 convolutions are plain GEMMs in either layout.
//
// All kernels: 16-byte accesses (C % 8 == 0), fp32 arithmetic, bf16 rounding at the points PyTorch rounds.
#include <math_constants.h>

#include "common.cuh"

namespace lb {
namespace vq {

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
    f[0] = bf16_lo(v.x); f[1] = bf16_hi(v.x); f[2] = bf16_lo(v.y); f[3] = bf16_hi(v.y);
    f[4] = bf16_lo(v.z); f[5] = bf16_hi(v.z); f[6] = bf16_lo(v.w); f[7] = bf16_hi(v.w);
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
    return make_uint4(pack_bf16(f[0], f[1]), pack_bf16(f[2], f[3]), pack_bf16(f[4], f[5]), pack_bf16(f[6], f[7]));
}
__device__ __forceinline__ float round_bf16(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// row of pixel (b, y, x) in a [B, H, W] image batch, padded (one pixel of border) or compact
__device__ __forceinline__ int64_t pix_row(int b, int y, int x, int H, int W, int padded) {
    return padded ? ((int64_t)b * (H + 2) + y + 1) * (W + 2) + x + 1 : ((int64_t)b * H + y) * W + x;
}

// ------------------------------------------------------------------ ids -> +-1 codes (zero-padded K for the GEMM)
// ids [Q, B, N] (token ids, offset removed here) -> codes [B*N, ld]: column q*bits + d = bit (bits-1-d) of code q ? +1 : -1
__global__ void __launch_bounds__(256) lfq_codes_kernel(const int64_t* __restrict__ ids, int64_t offset, int Q, int64_t BN, int bits,
                                                        __nv_bfloat16* __restrict__ codes, int ld) {
    const int64_t total = BN * ld;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t t = i / ld;
        const int c = (int)(i - t * ld);
        float v = 0.f;
        if (c < Q * bits) {
            const int q = c / bits, d = c - q * bits;
            const int64_t code = ids[(int64_t)q * BN + t] - offset;
            v = ((code >> (bits - 1 - d)) & 1) ? 1.f : -1.f;
        }
        codes[i] = __float2bfloat16_rn(v);
    }
}

// ------------------------------------------------------------------ GroupNorm(32, eps, affine) [+ swish]
// stage 1: per (sample, chunk of pixels) per-CHANNEL sums and sums of squares (fp32), deterministic tree, no atomics
// partial: [B][nchunk][2][C]
__global__ void __launch_bounds__(256) gn_partial_kernel(const __nv_bfloat16* __restrict__ x, int H, int W, int C, int in_padded,
                                                         int nchunk, float* __restrict__ partial) {
    extern __shared__ float sm[];                          // [P][2][C]
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int vpp = C >> 3;                                // 16-byte vectors per pixel
    const int P = 256 / vpp;                               // pixel lanes
    const int pl = threadIdx.x / vpp, v = threadIdx.x - pl * vpp;
    const int npix = H * W;
    const int per = (npix + nchunk - 1) / nchunk;
    const int p0 = chunk * per, p1 = min(npix, p0 + per);
    float s[8], ss[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] = ss[j] = 0.f;
    if (pl < P) {
        for (int pix = p0 + pl; pix < p1; pix += P) {
            const int y = pix / W, xx = pix - y * W;
            const uint4 u = __ldg(reinterpret_cast<const uint4*>(x + pix_row(b, y, xx, H, W, in_padded) * C) + v);
            float f[8];
            unpack8(u, f);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                s[j] += f[j];
                ss[j] = fmaf(f[j], f[j], ss[j]);
            }
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            sm[(pl * 2 + 0) * C + v * 8 + j] = s[j];
            sm[(pl * 2 + 1) * C + v * 8 + j] = ss[j];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < 2 * C; c += 256) {       // c < C: sums, else sums of squares
        const int which = c / C, ch = c - which * C;
        float t = 0.f;
        for (int q = 0; q < P; ++q) t += sm[(q * 2 + which) * C + ch];
        partial[(((int64_t)b * nchunk + chunk) * 2 + which) * C + ch] = t;
    }
}
// stage 2: fold the partials per group (double), normalise, affine, optional swish; interior pixels of the output get the
// result, border pixels (out_padded) zeros.  One thread per 16-byte vector of an OUTPUT row.
// Rounding: group_norm's bf16 output is rounded first, swish = x * sigmoid(x) is evaluated on that bf16 value (two eager ops).
__global__ void __launch_bounds__(256) gn_apply_kernel(const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ gamma,
                                                       const __nv_bfloat16* __restrict__ beta, const float* __restrict__ partial,
                                                       __nv_bfloat16* __restrict__ y, int B, int H, int W, int C, int groups,
                                                       int in_padded, int out_padded, int nchunk, float eps, int swish) {
    __shared__ float s_mean[64], s_rstd[64];
    const int b = blockIdx.y;
    const int cg = C / groups;
    if ((int)threadIdx.x < groups) {
        double s = 0.0, ss = 0.0;
        for (int ch = 0; ch < nchunk; ++ch) {
            const float* ps = partial + (((int64_t)b * nchunk + ch) * 2) * C + threadIdx.x * cg;
            for (int c = 0; c < cg; ++c) {
                s += (double)ps[c];
                ss += (double)ps[C + c];
            }
        }
        const double n = (double)H * W * cg;
        const double mean = s / n;
        double var = ss / n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    const int vpp = C >> 3;
    const int Ho = out_padded ? H + 2 : H, Wo = out_padded ? W + 2 : W;
    const int64_t nvec = (int64_t)Ho * Wo * vpp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t prow = i / vpp;
        const int v = (int)(i - prow * vpp);
        const int yo = (int)(prow / Wo), xo = (int)(prow - (int64_t)yo * Wo);
        const int yy = out_padded ? yo - 1 : yo, xx = out_padded ? xo - 1 : xo;
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (yy >= 0 && yy < H && xx >= 0 && xx < W) {
            float f[8], g[8], bt[8];
            unpack8(__ldg(reinterpret_cast<const uint4*>(x + pix_row(b, yy, xx, H, W, in_padded) * C) + v), f);
            unpack8(__ldg(reinterpret_cast<const uint4*>(gamma) + v), g);
            unpack8(__ldg(reinterpret_cast<const uint4*>(beta) + v), bt);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int grp = (v * 8 + j) / cg;
                float t = round_bf16((f[j] - s_mean[grp]) * s_rstd[grp] * g[j] + bt[j]);
                if (swish) t = t / (1.f + __expf(-t));
                f[j] = t;
            }
            o = pack8(f);
        }
        reinterpret_cast<uint4*>(y + ((int64_t)b * Ho * Wo + prow) * C)[v] = o;
    }
}

// ------------------------------------------------------------------ nearest upsample, padded or compact in -> padded out
// src_y [Ho], src_x [Wo]: source index of every output row / column (computed by the host with PyTorch's formula)
__global__ void __launch_bounds__(256) upsample_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                                       const int32_t* __restrict__ src_y, const int32_t* __restrict__ src_x, int B,
                                                       int H, int W, int C, int Ho, int Wo, int in_padded) {
    const int vpp = C >> 3;
    const int64_t nvec = (int64_t)B * (Ho + 2) * (Wo + 2) * vpp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t prow = i / vpp;
        const int v = (int)(i - prow * vpp);
        const int xo = (int)(prow % (Wo + 2)) - 1;
        const int64_t t = prow / (Wo + 2);
        const int yo = (int)(t % (Ho + 2)) - 1;
        const int b = (int)(t / (Ho + 2));
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (yo >= 0 && yo < Ho && xo >= 0 && xo < Wo)
            o = __ldg(reinterpret_cast<const uint4*>(x + pix_row(b, __ldg(src_y + yo), __ldg(src_x + xo), H, W, in_padded) * C) + v);
        reinterpret_cast<uint4*>(y + prow * C)[v] = o;
    }
}

// ------------------------------------------------------------------ compact [B*H*W, C] (+ padded addend) -> padded, zero borders
__global__ void __launch_bounds__(256) pad_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, const __nv_bfloat16* __restrict__ add,
                                                  __nv_bfloat16* __restrict__ y, int B, int H, int W, int C) {
    const int vpp = C >> 3;
    const int64_t nvec = (int64_t)B * (H + 2) * (W + 2) * vpp;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t prow = i / vpp;
        const int v = (int)(i - prow * vpp);
        const int xo = (int)(prow % (W + 2)) - 1;
        const int64_t t = prow / (W + 2);
        const int yo = (int)(t % (H + 2)) - 1;
        const int b = (int)(t / (H + 2));
        uint4 o = make_uint4(0u, 0u, 0u, 0u);
        if (yo >= 0 && yo < H && xo >= 0 && xo < W) {
            o = __ldg(reinterpret_cast<const uint4*>(x + (((int64_t)b * H + yo) * W + xo) * ldx) + v);
            if (add) {                                       // residual: bf16(a + h), the eager `x + h_`
                float f[8], a[8];
                unpack8(o, f);
                unpack8(__ldg(reinterpret_cast<const uint4*>(add + prow * C) + v), a);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] += a[j];
                o = pack8(f);
            }
        }
        reinterpret_cast<uint4*>(y + prow * C)[v] = o;
    }
}

// ------------------------------------------------------------------ padded NHWC [.., C] -> NCHW [B, Cout, H, W] (Cout <= C)
__global__ void __launch_bounds__(256) to_nchw_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int B, int H,
                                                      int W, int C, int Cout) {
    const int64_t total = (int64_t)B * Cout * H * W;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int xx = (int)(i % W);
        int64_t t = i / W;
        const int yy = (int)(t % H);
        t /= H;
        const int c = (int)(t % Cout);
        const int b = (int)(t / Cout);
        y[i] = x[pix_row(b, yy, xx, H, W, 1) * C + c];
    }
}

// ------------------------------------------------------------------ row softmax in place: x = softmax(bf16(x * scale)) (bf16)
// one warp per row, three passes over the row (L1-resident: rows are a few KB); fp32 maths as torch.softmax on bf16 input
__global__ void __launch_bounds__(256) softmax_rows_kernel(__nv_bfloat16* __restrict__ x, int64_t rows, int cols, int64_t ld,
                                                           float scale) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    __nv_bfloat16* r = x + row * ld;
    float mx = -CUDART_INF_F;
    for (int c = lane; c < cols; c += 32) mx = fmaxf(mx, round_bf16(__bfloat162float(r[c]) * scale));
    mx = warp_max(mx);
    float sum = 0.f;
    for (int c = lane; c < cols; c += 32) sum += __expf(round_bf16(__bfloat162float(r[c]) * scale) - mx);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int c = lane; c < cols; c += 32) r[c] = __float2bfloat16_rn(__expf(round_bf16(__bfloat162float(r[c]) * scale) - mx) * inv);
}

static int grid_for(int64_t items) {
    int64_t g = (items + 255) / 256;
    const int64_t cap = (int64_t)sm_count() * 16;
    if (g > cap) g = cap;
    return (int)(g < 1 ? 1 : g);
}

}  // namespace vq
}  // namespace lb

using namespace lb;

#define VQ_AL16(p) ((reinterpret_cast<uintptr_t>(p) & 15) == 0)

extern "C" {

int lb_vq_codes(const int64_t* ids, int64_t offset, int num_codebooks, int64_t tokens, int bits, void* codes, int ld, void* stream) {
    LB_REQUIRE(ids && codes && num_codebooks > 0 && bits > 0 && bits < 32 && tokens >= 0 && ld >= num_codebooks * bits, LB_EINVAL,
               "vq_codes: bad arguments (ld %d must cover %d code columns)", ld, num_codebooks * bits);
    if (tokens == 0) return LB_OK;
    vq::lfq_codes_kernel<<<vq::grid_for(tokens * ld), 256, 0, (cudaStream_t)stream>>>(ids, offset, num_codebooks, tokens, bits,
                                                                                     (__nv_bfloat16*)codes, ld);
    return check_launch("vq_codes");
}

int lb_vq_groupnorm_chunks(int height, int width) {
    const int64_t npix = (int64_t)height * width;
    int64_t n = (npix + 1023) / 1024;                      // ~1 k pixels per CTA, at most 128 chunks per sample
    if (n > 128) n = 128;
    return (int)(n < 1 ? 1 : n);
}

int lb_vq_groupnorm(const void* x, const void* gamma, const void* beta, void* y, float* workspace, int batch, int height, int width,
                    int channels, int groups, float eps, int swish, int in_padded, int out_padded, void* stream) {
    LB_REQUIRE(x && gamma && beta && y && workspace, LB_EINVAL, "vq_groupnorm: null argument");
    LB_REQUIRE(batch > 0 && height > 0 && width > 0 && channels > 0 && channels % 8 == 0 && channels <= 2048, LB_EINVAL,
               "vq_groupnorm: channels %d must be a multiple of 8, <= 2048", channels);
    LB_REQUIRE(groups > 0 && groups <= 64 && channels % groups == 0, LB_EINVAL, "vq_groupnorm: %d groups over %d channels", groups, channels);
    LB_REQUIRE(VQ_AL16(x) && VQ_AL16(y) && VQ_AL16(gamma) && VQ_AL16(beta), LB_EALIGN, "vq_groupnorm: pointers must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)stream;
    const int nchunk = lb_vq_groupnorm_chunks(height, width);
    const int vpp = channels / 8, P = 256 / vpp;
    LB_REQUIRE(P >= 1, LB_EINVAL, "vq_groupnorm: too many channels");
    const size_t smem = (size_t)P * 2 * channels * sizeof(float);
    static bool configured = false;
    if (!configured) {
        cudaFuncSetAttribute(vq::gn_partial_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
        configured = true;
    }
    LB_REQUIRE(smem <= 128 * 1024, LB_EINVAL, "vq_groupnorm: shared memory");
    vq::gn_partial_kernel<<<dim3(nchunk, batch), 256, smem, st>>>((const __nv_bfloat16*)x, height, width, channels, in_padded, nchunk,
                                                                  workspace);
    int rc = check_launch("vq_groupnorm_partial");
    if (rc) return rc;
    const int Ho = out_padded ? height + 2 : height, Wo = out_padded ? width + 2 : width;
    int gx = vq::grid_for((int64_t)Ho * Wo * vpp);
    vq::gn_apply_kernel<<<dim3(gx, batch), 256, 0, st>>>((const __nv_bfloat16*)x, (const __nv_bfloat16*)gamma, (const __nv_bfloat16*)beta,
                                                         workspace, (__nv_bfloat16*)y, batch, height, width, channels, groups, in_padded,
                                                         out_padded, nchunk, eps, swish);
    return check_launch("vq_groupnorm_apply");
}

int lb_vq_upsample_nearest(const void* x, void* y, const int32_t* src_y, const int32_t* src_x, int batch, int height, int width,
                           int channels, int out_height, int out_width, int in_padded, void* stream) {
    LB_REQUIRE(x && y && src_y && src_x && batch > 0 && height > 0 && width > 0 && out_height > 0 && out_width > 0, LB_EINVAL,
               "vq_upsample: bad arguments");
    LB_REQUIRE(channels > 0 && channels % 8 == 0 && VQ_AL16(x) && VQ_AL16(y), LB_EALIGN, "vq_upsample: channels %% 8, 16-byte pointers");
    const int64_t nvec = (int64_t)batch * (out_height + 2) * (out_width + 2) * (channels / 8);
    vq::upsample_kernel<<<vq::grid_for(nvec), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, src_y, src_x, batch,
                                                                            height, width, channels, out_height, out_width, in_padded);
    return check_launch("vq_upsample");
}

int lb_vq_pad(const void* x, int64_t ldx, const void* addend, void* y, int batch, int height, int width, int channels, void* stream) {
    LB_REQUIRE(x && y && batch > 0 && height > 0 && width > 0, LB_EINVAL, "vq_pad: bad arguments");
    LB_REQUIRE(channels > 0 && channels % 8 == 0 && ldx % 8 == 0 && ldx >= channels && VQ_AL16(x) && VQ_AL16(y) && (!addend || VQ_AL16(addend)),
               LB_EALIGN, "vq_pad: channels %% 8, ldx %% 8, 16-byte pointers");
    const int64_t nvec = (int64_t)batch * (height + 2) * (width + 2) * (channels / 8);
    vq::pad_kernel<<<vq::grid_for(nvec), 256, 0, (cudaStream_t)stream>>>((const __nv_bfloat16*)x, ldx, (const __nv_bfloat16*)addend,
                                                                       (__nv_bfloat16*)y, batch, height, width, channels);
    return check_launch("vq_pad");
}

int lb_vq_to_nchw(const void* x, void* y, int batch, int height, int width, int channels, int channels_out, void* stream) {
    LB_REQUIRE(x && y && batch > 0 && height > 0 && width > 0 && channels_out > 0 && channels_out <= channels, LB_EINVAL,
               "vq_to_nchw: bad arguments");
    vq::to_nchw_kernel<<<vq::grid_for((int64_t)batch * channels_out * height * width), 256, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (__nv_bfloat16*)y, batch, height, width, channels, channels_out);
    return check_launch("vq_to_nchw");
}

int lb_softmax_rows(void* x, int64_t rows, int cols, int64_t ld, float scale, void* stream) {
    LB_REQUIRE(x && rows >= 0 && cols > 0 && ld >= cols, LB_EINVAL, "softmax_rows: bad arguments");
    if (rows == 0) return LB_OK;
    vq::softmax_rows_kernel<<<(unsigned)((rows + 7) / 8), 256, 0, (cudaStream_t)stream>>>((__nv_bfloat16*)x, rows, cols, ld, scale);
    return check_launch("softmax_rows");
}

}

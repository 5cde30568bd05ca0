// Routed RMSNorm (A13) and LayerNorm (A1/A4), forward + backward.
//
// HBM-bound: algorithmic traffic is one read of x (+dy in bwd) and one write of
// y (dx) per element.  One CTA of 256 threads owns a row at a time and keeps
// it in registers (128-bit loads, <= 4 vectors per thread => cols <= 8192), so
// every element crosses HBM once.  The grid is persistent (a multiple of the
// SM count) and strides over rows; weight-gradient partials live in registers
// per CTA and are reduced by a second tiny kernel (no atomics).
#include "common.cuh"

namespace lb {

constexpr int NT = 256;      // threads per CTA
constexpr int MAXV = 4;      // 16-byte vectors per thread
constexpr int CTAS_PER_SM = 4;

union V8 {
    uint4 u;
    __nv_bfloat162 h[4];
};

__device__ __forceinline__ void unpack8(const uint4& u, float* f) {
    V8 v;
    v.u = u;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        float2 t = __bfloat1622float2(v.h[i]);
        f[2 * i] = t.x;
        f[2 * i + 1] = t.y;
    }
}
__device__ __forceinline__ uint4 pack8(const float* f) {
    V8 v;
#pragma unroll
    for (int i = 0; i < 4; ++i) v.h[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
    return v.u;
}

// block-wide sum of up to two values (returns to all threads)
__device__ __forceinline__ float2 block_sum2(float a, float b, float* red) {
    a = warp_sum(a);
    b = warp_sum(b);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();               // protect `red` from the previous use
    if (l == 0) {
        red[w] = a;
        red[w + 8] = b;
    }
    __syncthreads();
    float ra = 0.f, rb = 0.f;
#pragma unroll
    for (int i = 0; i < NT / 32; ++i) {
        ra += red[i];
        rb += red[i + 8];
    }
    return make_float2(ra, rb);
}

// ---------------------------------------------------------------- RMSNorm fwd
__global__ void __launch_bounds__(NT) rmsnorm_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                         const __nv_bfloat16* __restrict__ w_lang,
                                                         const __nv_bfloat16* __restrict__ w_vis,
                                                         const uint8_t* __restrict__ flag, __nv_bfloat16* __restrict__ y,
                                                         float* __restrict__ rstd, int64_t rows, int cols, float eps) {
    __shared__ float red[16];
    pdl_trigger();
    pdl_wait();
    const int nvec = cols >> 3;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + r * cols);
        uint4 xv[MAXV];
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = threadIdx.x + i * NT;
            if (v < nvec) {
                xv[i] = __ldg(xr + v);
                float f[8];
                unpack8(xv[i], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) ss += f[j] * f[j];
            }
        }
        const float tot = block_sum2(ss, 0.f, red).x;
        const float rs = rsqrtf(tot / (float)cols + eps);
        if (threadIdx.x == 0 && rstd) rstd[r] = rs;
        const bool vis = flag ? (flag[r] != 0) : false;
        const uint4* wr = reinterpret_cast<const uint4*>(vis ? w_vis : w_lang);
        uint4* yr = reinterpret_cast<uint4*>(y + r * cols);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = threadIdx.x + i * NT;
            if (v < nvec) {
                float f[8], g[8];
                unpack8(xv[i], f);
                unpack8(__ldg(wr + v), g);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = g[j] * (f[j] * rs);
                yr[v] = pack8(f);
            }
        }
    }
}

// ---------------------------------------------------------------- RMSNorm bwd
// dx = rs * (g - x * rs^2 * mean(g . x)) (+ residual grad),  g = w . dy ;  dw[m] += sum_rows dy . x . rs
// One CTA per row at a time, the row (x, dy) kept in registers: every element crosses HBM once.  The row's mean(g.x) is a
// block reduction (one barrier per row: the scratch is double buffered).  dw: every thread owns fixed columns and adds its
// rows' contributions, in row order, into its own slots of a per-CTA shared-memory accumulator -- no atomics, so the result
// is bit-reproducible run to run (gradient checkpointing's recompute must give identical gradients); the CTA's accumulator
// goes to `partial`, folded in fixed order by reduce_partials_kernel.
constexpr int RB_WARPS = 8;

__global__ void __launch_bounds__(RB_WARPS * 32) rmsnorm_bwd_kernel(
    const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, const __nv_bfloat16* __restrict__ w_lang,
    const __nv_bfloat16* __restrict__ w_vis, const uint8_t* __restrict__ flag, const float* __restrict__ rstd,
    const __nv_bfloat16* __restrict__ resid, __nv_bfloat16* __restrict__ dx, float* __restrict__ partial, int64_t rows,
    int cols) {
    extern __shared__ float dwacc[];                    // [2 modalities][cols], slot (v, j) -> v * 8 + j, owned by thread v % 256
    __shared__ float red[2][RB_WARPS];
    const int nvec = cols >> 3;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 2 * cols; i += blockDim.x) dwacc[i] = 0.f;
    __syncthreads();
    int buf = 0;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x, buf ^= 1) {
        const bool vis = flag ? (flag[r] != 0) : false;
        const uint4* xr = reinterpret_cast<const uint4*>(x + r * cols);
        const uint4* dr = reinterpret_cast<const uint4*>(dy + r * cols);
        const uint4* wr = reinterpret_cast<const uint4*>(vis ? w_vis : w_lang);
        const float rs = rstd[r];
        uint4 xv[MAXV], dv[MAXV], wv[MAXV];
        float dot = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = threadIdx.x + i * NT;
            if (v < nvec) {
                xv[i] = __ldg(xr + v);
                dv[i] = __ldg(dr + v);
                wv[i] = __ldg(wr + v);
            }
        }
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = threadIdx.x + i * NT;
            if (v < nvec) {
                float fx[8], fd[8], fw[8];
                unpack8(xv[i], fx);
                unpack8(dv[i], fd);
                unpack8(wv[i], fw);
#pragma unroll
                for (int j = 0; j < 8; ++j) dot += fw[j] * fd[j] * fx[j];
            }
        }
        dot = warp_sum(dot);
        if (lane == 0) red[buf][warp] = dot;
        __syncthreads();                                 // the other buffer is free again two rows later: one barrier per row
        float tot = 0.f;
#pragma unroll
        for (int i = 0; i < RB_WARPS; ++i) tot += red[buf][i];
        const float coef = tot / (float)cols * rs * rs * rs;
        uint4* oxr = reinterpret_cast<uint4*>(dx + r * cols);
        const uint4* rr = resid ? reinterpret_cast<const uint4*>(resid + r * cols) : nullptr;
        float* acc = dwacc + (vis ? cols : 0);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = threadIdx.x + i * NT;
            if (v < nvec) {
                float fx[8], fd[8], fw[8], o[8];
                unpack8(xv[i], fx);
                unpack8(dv[i], fd);
                unpack8(wv[i], fw);
#pragma unroll
                for (int j = 0; j < 8; ++j) o[j] = rs * fw[j] * fd[j] - fx[j] * coef;
                if (rr) {
                    float fr[8];
                    unpack8(__ldg(rr + v), fr);
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] += fr[j];
                }
                oxr[v] = pack8(o);
                if (partial) {
                    float4* a4 = reinterpret_cast<float4*>(acc + v * 8);
                    float4 a0 = a4[0], a1 = a4[1];
                    a0.x += fd[0] * fx[0] * rs; a0.y += fd[1] * fx[1] * rs; a0.z += fd[2] * fx[2] * rs; a0.w += fd[3] * fx[3] * rs;
                    a1.x += fd[4] * fx[4] * rs; a1.y += fd[5] * fx[5] * rs; a1.z += fd[6] * fx[6] * rs; a1.w += fd[7] * fx[7] * rs;
                    a4[0] = a0;
                    a4[1] = a1;
                }
            }
        }
    }
    __syncthreads();
    if (partial) {
        float* out = partial + (int64_t)blockIdx.x * 2 * cols;      // partial[block][2][cols], natural column order
        for (int i = threadIdx.x; i < 2 * cols; i += blockDim.x) out[i] = dwacc[i];
    }
}

// out[c] += sum_b partial[b][off + c]: CTA = 32 columns x 8 row lanes, rows strided by 8, then a shared-memory fold.
// blockIdx.y selects (off, out) / (off2, out2): both halves of the partial rows in one launch.
__global__ void __launch_bounds__(256) reduce_partials_kernel(const float* __restrict__ partial, int nblocks, int width, int off,
                                                              int stride, float* __restrict__ out, int off2 = 0,
                                                              float* __restrict__ out2 = nullptr) {
    __shared__ float sm[8][33];
    if (blockIdx.y == 1) {
        off = off2;
        out = out2;
    }
    if (!out) return;
    const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + cx;
    float s = 0.f;
    if (c < width)
        for (int b = ry; b < nblocks; b += 8) s += partial[(int64_t)b * stride + off + c];
    sm[ry][cx] = s;
    __syncthreads();
    if (ry == 0 && c < width) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; ++i) t += sm[i][cx];
        out[c] += t;
    }
}

// -------------------------------------------------------------- LayerNorm fwd
__global__ void __launch_bounds__(NT) layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ x,
                                                           const __nv_bfloat16* __restrict__ w,
                                                           const __nv_bfloat16* __restrict__ b,
                                                           __nv_bfloat16* __restrict__ y, float* __restrict__ mean,
                                                           float* __restrict__ rstd, int64_t rows, int cols, float eps) {
    __shared__ float red[16];
    const int nvec = cols >> 3;
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + r * cols);
        uint4 xv[MAXV];
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = threadIdx.x + i * NT;
            if (v < nvec) {
                xv[i] = __ldg(xr + v);
                float f[8];
                unpack8(xv[i], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) s += f[j];
            }
        }
        const float mu = block_sum2(s, 0.f, red).x / (float)cols;
        float ss = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = threadIdx.x + i * NT;
            if (v < nvec) {
                float f[8];
                unpack8(xv[i], f);
#pragma unroll
                for (int j = 0; j < 8; ++j) ss += (f[j] - mu) * (f[j] - mu);
            }
        }
        const float rs = rsqrtf(block_sum2(ss, 0.f, red).x / (float)cols + eps);
        if (threadIdx.x == 0) {
            if (mean) mean[r] = mu;
            if (rstd) rstd[r] = rs;
        }
        const uint4* wr = reinterpret_cast<const uint4*>(w);
        const uint4* br = reinterpret_cast<const uint4*>(b);
        uint4* yr = reinterpret_cast<uint4*>(y + r * cols);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = threadIdx.x + i * NT;
            if (v < nvec) {
                float f[8], g[8], h[8];
                unpack8(xv[i], f);
                unpack8(__ldg(wr + v), g);
                unpack8(__ldg(br + v), h);
#pragma unroll
                for (int j = 0; j < 8; ++j) f[j] = (f[j] - mu) * rs * g[j] + h[j];
                yr[v] = pack8(f);
            }
        }
    }
}

// -------------------------------------------------------------- LayerNorm bwd
// xh = (x-mu)*rs ; g = w.dy ; dx = rs * (g - mean(g) - xh * mean(g.xh)) ; dw += dy.xh ; db += dy
__global__ void __launch_bounds__(NT) layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ dy,
                                                           const __nv_bfloat16* __restrict__ x,
                                                           const __nv_bfloat16* __restrict__ w,
                                                           const float* __restrict__ mean, const float* __restrict__ rstd,
                                                           __nv_bfloat16* __restrict__ dx, float* __restrict__ partial,
                                                           int64_t rows, int cols) {
    __shared__ float red[16];
    const int nvec = cols >> 3;
    float accW[MAXV][8], accB[MAXV][8];
#pragma unroll
    for (int i = 0; i < MAXV; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) accW[i][j] = accB[i][j] = 0.f;
    const uint4* wr = reinterpret_cast<const uint4*>(w);
    for (int64_t r = blockIdx.x; r < rows; r += gridDim.x) {
        const uint4* xr = reinterpret_cast<const uint4*>(x + r * cols);
        const uint4* dr = reinterpret_cast<const uint4*>(dy + r * cols);
        const float mu = mean[r], rs = rstd[r];
        uint4 xv[MAXV], dv[MAXV];
        float s1 = 0.f, s2 = 0.f;
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = threadIdx.x + i * NT;
            if (v < nvec) {
                xv[i] = __ldg(xr + v);
                dv[i] = __ldg(dr + v);
                float fx[8], fd[8], fw[8];
                unpack8(xv[i], fx);
                unpack8(dv[i], fd);
                unpack8(__ldg(wr + v), fw);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (fx[j] - mu) * rs;
                    const float g = fw[j] * fd[j];
                    s1 += g;
                    s2 += g * xh;
                    accW[i][j] += fd[j] * xh;
                    accB[i][j] += fd[j];
                }
            }
        }
        const float2 t = block_sum2(s1, s2, red);
        const float m1 = t.x / (float)cols, m2 = t.y / (float)cols;
        uint4* oxr = reinterpret_cast<uint4*>(dx + r * cols);
#pragma unroll
        for (int i = 0; i < MAXV; ++i) {
            const int v = threadIdx.x + i * NT;
            if (v < nvec) {
                float fx[8], fd[8], fw[8], o[8];
                unpack8(xv[i], fx);
                unpack8(dv[i], fd);
                unpack8(__ldg(wr + v), fw);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float xh = (fx[j] - mu) * rs;
                    o[j] = rs * (fw[j] * fd[j] - m1 - xh * m2);
                }
                oxr[v] = pack8(o);
            }
        }
    }
    float* pw = partial + (int64_t)blockIdx.x * 2 * cols;
    float* pb = pw + cols;
#pragma unroll
    for (int i = 0; i < MAXV; ++i) {
        const int v = threadIdx.x + i * NT;
        if (v < nvec) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                pw[v * 8 + j] = accW[i][j];
                pb[v * 8 + j] = accB[i][j];
            }
        }
    }
}

static int norm_grid(int64_t rows) {
    int64_t g = (int64_t)sm_count() * CTAS_PER_SM;
    return (int)(rows < g ? (rows > 0 ? rows : 1) : g);
}

static int rms_bwd_grid(int64_t rows) {
    int64_t g = (int64_t)sm_count() * 4;
    const int64_t need = rows;                       // one CTA per row at a time
    return (int)(need < g ? (need > 0 ? need : 1) : g);
}

static int check_norm_args(const void* x, const void* y, int64_t rows, int cols) {
    LB_REQUIRE(rows >= 0 && cols > 0, LB_EINVAL, "norm: bad shape rows=%lld cols=%d", (long long)rows, cols);
    LB_REQUIRE(cols % 8 == 0 && cols <= NT * MAXV * 8, LB_EINVAL, "norm: cols=%d must be a multiple of 8 and <= %d", cols,
               NT * MAXV * 8);
    LB_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)y & 15) == 0, LB_EALIGN, "norm: pointers must be 16-byte aligned");
    return LB_OK;
}

}  // namespace lb

using namespace lb;

extern "C" {

int lb_rmsnorm_fwd(const void* x, const void* w_lang, const void* w_vis, const uint8_t* flag, void* y, float* rstd,
                   int64_t rows, int cols, float eps, void* stream) {
    int rc = check_norm_args(x, y, rows, cols);
    if (rc) return rc;
    LB_REQUIRE(w_lang && (w_vis || !flag), LB_EINVAL, "rmsnorm_fwd: missing weight");
    if (rows == 0) return LB_OK;
    launch_chain(rmsnorm_fwd_kernel, dim3(norm_grid(rows)), dim3(NT), 0, (cudaStream_t)stream,
                 (const __nv_bfloat16*)x, (const __nv_bfloat16*)w_lang, (const __nv_bfloat16*)(w_vis ? w_vis : w_lang), flag,
                 (__nv_bfloat16*)y, rstd, rows, cols, eps);
    return check_launch("rmsnorm_fwd");
}

int64_t lb_rmsnorm_bwd_workspace(int64_t rows, int cols) {
    const int64_t g = norm_grid(rows) > rms_bwd_grid(rows) ? norm_grid(rows) : rms_bwd_grid(rows);
    return g * 2 * cols * (int64_t)sizeof(float);
}

int lb_rmsnorm_bwd(const void* dy, const void* x, const void* w_lang, const void* w_vis, const uint8_t* flag,
                   const float* rstd, const void* residual_grad, void* dx, float* dw_lang, float* dw_vis, void* partial,
                   int64_t rows, int cols, void* stream) {
    int rc = check_norm_args(x, dx, rows, cols);
    if (rc) return rc;
    LB_REQUIRE(dy && rstd && partial && w_lang, LB_EINVAL, "rmsnorm_bwd: null argument");
    if (rows == 0) return LB_OK;
    const int grid = rms_bwd_grid(rows);
    cudaStream_t st = (cudaStream_t)stream;
    const int smem = 2 * cols * (int)sizeof(float);
    static int configured_smem = 0;
    if (smem > configured_smem) {
        cudaError_t e = cudaFuncSetAttribute(rmsnorm_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        if (e != cudaSuccess) return fail(LB_ELAUNCH, "rmsnorm_bwd: cudaFuncSetAttribute: %s", cudaGetErrorString(e));
        configured_smem = smem;
    }
    const bool need_dw = dw_lang || dw_vis;
    rmsnorm_bwd_kernel<<<grid, RB_WARPS * 32, smem, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x,
                                                          (const __nv_bfloat16*)w_lang,
                                                          (const __nv_bfloat16*)(w_vis ? w_vis : w_lang), flag, rstd,
                                                          (const __nv_bfloat16*)residual_grad, (__nv_bfloat16*)dx,
                                                          need_dw ? (float*)partial : nullptr, rows, cols);
    rc = check_launch("rmsnorm_bwd");
    if (rc) return rc;
    const int tb = 256, gb = ceil_div(cols, 32);
    if (dw_lang || dw_vis)
        reduce_partials_kernel<<<dim3(gb, 2), tb, 0, st>>>((const float*)partial, grid, cols, 0, 2 * cols, dw_lang, cols, dw_vis);
    return check_launch("rmsnorm_bwd_reduce");
}

int lb_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, int64_t rows,
                     int cols, float eps, void* stream) {
    int rc = check_norm_args(x, y, rows, cols);
    if (rc) return rc;
    LB_REQUIRE(w && b, LB_EINVAL, "layernorm_fwd: missing affine parameters");
    if (rows == 0) return LB_OK;
    layernorm_fwd_kernel<<<norm_grid(rows), NT, 0, (cudaStream_t)stream>>>(
        (const __nv_bfloat16*)x, (const __nv_bfloat16*)w, (const __nv_bfloat16*)b, (__nv_bfloat16*)y, mean, rstd, rows,
        cols, eps);
    return check_launch("layernorm_fwd");
}

int64_t lb_layernorm_bwd_workspace(int64_t rows, int cols) { return lb_rmsnorm_bwd_workspace(rows, cols); }

int lb_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx,
                     float* dw, float* db, void* partial, int64_t rows, int cols, void* stream) {
    int rc = check_norm_args(x, dx, rows, cols);
    if (rc) return rc;
    LB_REQUIRE(dy && w && mean && rstd && partial, LB_EINVAL, "layernorm_bwd: null argument");
    if (rows == 0) return LB_OK;
    const int grid = norm_grid(rows);
    cudaStream_t st = (cudaStream_t)stream;
    layernorm_bwd_kernel<<<grid, NT, 0, st>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (const __nv_bfloat16*)w,
                                              mean, rstd, (__nv_bfloat16*)dx, (float*)partial, rows, cols);
    rc = check_launch("layernorm_bwd");
    if (rc) return rc;
    const int tb = 256, gb = ceil_div(cols, 32);
    if (dw || db)
        reduce_partials_kernel<<<dim3(gb, 2), tb, 0, st>>>((const float*)partial, grid, cols, 0, 2 * cols, dw, cols, db);
    return check_launch("layernorm_bwd_reduce");
}

}

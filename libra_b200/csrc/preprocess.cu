// N4 (second half): CLIP image preprocessing on the GPU -- raw uint8 RGB images in, normalised pixel_values out.
//
// Replaces, bit for bit, what the reference runs on the CPU for every training / inference image:
//   libra/data/processors/libra_processor.py:44-60   Expand2Square (paste on a square canvas of the mean colour; eval processor)
//   libra/data/processors/libra_processor.py:65-111  LibraEvalImageProcessor / LibraImageProcessor = CLIPImageProcessor
//   libra/models/clip/image_processing_clip.py:124-150 resize (shortest edge -> size, other edge int(size * long / short)),
//     :152-174 center crop, :176-217 rescale 1/255 + normalise, :296-337 order of the steps
//   and below them transformers.image_transforms.resize -> Pillow Image.resize(resample=BICUBIC) = ImagingResample
//   (libImaging/Resample.c): separable, horizontal pass first, double-precision coefficients (support 2 * max(scale, 1):
//   antialiased when shrinking) normalised to sum 1, 22-bit fixed point, int32 accumulation from 1 << 21, >> 22 and clamp to
//   uint8 after EACH pass.
//
// Layout / plan.  Images arrive packed in ONE device buffer (HWC uint8, image i at offsets[i]); sizes live on the host.  Per
// call the host computes only scalar geometry per image (canvas, output size, crop window, filter support) and uploads the
// descriptors in one small copy; everything else runs on the device:
//   pass 0 (tables): thread = (image, axis, crop index) computes its resampling window and weights in DOUBLE precision with
//     explicitly unfused IEEE operations (__dmul_rn / __dadd_rn / __ddiv_rn: the compiler must not contract a*b+c into an
//     FMA, Pillow's C code on x86-64 does not), normalises them to sum 1 and quantises to 22-bit fixed point -- the same bits
//     as Pillow's precompute_coeffs + normalize_coeffs_8bpc (host restatement below: lb_clip_resample_coeffs, tested against
//     the oracle on the CPU); only the crop window's `crop` columns and `crop` rows get tables;
//   pass 1 (horizontal): thread = crop column over 16 needed source rows, 3 channels; taps read from the source row (the
//     canvas colour outside the pasted image enters as colour x sum of those taps); uint8 result into the workspace
//     [rows_needed][crop][3];
//   pass 2 (vertical):   thread = crop column over 8 crop rows; taps walk down the workspace column (coalesced across the warp);
//     the uint8 result goes through a 3 x 256 float table (rescale in double -> float, (x - mean) / std in float: the
//     reference's rounding sequence) and is stored planar [3][crop][crop] as fp32 or bf16.
// HBM-bound byte work: algorithmic bytes per image = the source rows/columns the crop window touches (once) + 3 * crop^2 *
// sizeof(out); no tensor cores, no shared-memory staging needed (neighbouring threads share their taps through L1).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace lb {
namespace pp {

constexpr int PRECISION_BITS = 32 - 8 - 2;

struct ImgDesc {
    long long src_off;          // bytes into the packed image buffer
    long long tmp_off;          // bytes into the workspace's intermediate region
    int h, w;                   // source image
    int pad_top, pad_left;      // position of the image on the (virtual) canvas
    int y_first, rows_needed;   // canvas rows pass 1 produces
    int ksize_h, ksize_v;
    int kh_off, kv_off;         // int32 offsets into the table: horizontal [ksize][crop] (tap-major: a warp reads one line),
                                // vertical [crop][ksize] (pass 2 reads one row per block)
    int bh_off, bv_off;         // int32 offsets into the table: [crop][2] = (first input index, taps)
    int in_w, in_h;             // canvas size (resampling input)
    int out_w, out_h;           // resized size
    int left, top;              // crop window in the resized image
};

struct Params {
    const uint8_t* images;
    const ImgDesc* desc;
    const int32_t* tab;         // coefficients and bounds
    uint8_t* tmp;
    const float* lut;           // [3][256]
    void* out;
    uint8_t* out_u8;
    int crop, out_bf16;
    uint8_t bg[3];
};

static double bicubic_filter(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
    if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
    return 0.0;
}

// Resample.c precompute_coeffs + normalize_coeffs_8bpc for output indices [o0, o0 + n) of an axis resized in_size -> out_size.
// Appends n * ksize coefficients and n * 2 bounds; returns ksize.
static int coeffs(int in_size, int out_size, int o0, int n, std::vector<int32_t>& kk, std::vector<int32_t>& bounds) {
    double filterscale, scale;
    filterscale = scale = (double)((float)in_size - 0.0f) / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = 2.0 * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    std::vector<double> k((size_t)ksize);
    for (int xx = o0; xx < o0 + n; ++xx) {
        const double center = 0.0 + (xx + 0.5) * scale;
        double ww = 0.0;
        const double ss = 1.0 / filterscale;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; ++x) {
            const double w = bicubic_filter((x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        for (int x = 0; x < xmax; ++x)
            if (ww != 0.0) k[x] /= ww;
        for (int x = 0; x < ksize; ++x) {
            int32_t q = 0;
            if (x < xmax) q = k[x] < 0 ? (int)(-0.5 + k[x] * (1 << PRECISION_BITS)) : (int)(0.5 + k[x] * (1 << PRECISION_BITS));
            kk.push_back(q);
        }
        bounds.push_back(xmin);
        bounds.push_back(xmax);
    }
    return ksize;
}

// scalar geometry of one axis (host): ksize, and the window of one output index -- the same expressions as coeffs()
struct Axis {
    double scale, filterscale, support;
    int ksize;
};
static Axis axis_of(int in_size, int out_size) {
    Axis a;
    a.filterscale = a.scale = (double)((float)in_size - 0.0f) / out_size;
    if (a.filterscale < 1.0) a.filterscale = 1.0;
    a.support = 2.0 * a.filterscale;
    a.ksize = (int)ceil(a.support) * 2 + 1;
    return a;
}
static void window_of(const Axis& a, int in_size, int xx, int* xmin_out, int* n_out) {
    const double center = 0.0 + (xx + 0.5) * a.scale;
    int xmin = (int)(center - a.support + 0.5);
    if (xmin < 0) xmin = 0;
    int xmax = (int)(center + a.support + 0.5);
    if (xmax > in_size) xmax = in_size;
    *xmin_out = xmin;
    *n_out = xmax - xmin;
}

// ---- pass 0 on the device: every operation spelled out so that nothing is contracted into an FMA
__device__ __forceinline__ double bicubic_dev(double x) {
    const double a = -0.5;
    if (x < 0.0) x = -x;
    if (x < 1.0)      // ((a + 2.0) * x - (a + 3.0)) * x * x + 1
        return __dadd_rn(__dmul_rn(__dmul_rn(__dadd_rn(__dmul_rn(a + 2.0, x), -(a + 3.0)), x), x), 1.0);
    if (x < 2.0)      // (((x - 5) * x + 8) * x - 4) * a
        return __dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(__dmul_rn(__dadd_rn(x, -5.0), x), 8.0), x), -4.0), a);
    return 0.0;
}

// grid (ceil(crop / 128), 2, n_images), block 128: axis 0 = horizontal, 1 = vertical
__global__ void __launch_bounds__(128) tables_kernel(const ImgDesc* __restrict__ desc, int32_t* __restrict__ tab, int crop) {
    const ImgDesc d = desc[blockIdx.z];
    const int i = blockIdx.x * 128 + threadIdx.x;
    if (i >= crop) return;
    const bool vert = blockIdx.y != 0;
    const int in_size = vert ? d.in_h : d.in_w, out_size = vert ? d.out_h : d.out_w;
    const int xx = (vert ? d.top : d.left) + i;
    const int ksize = vert ? d.ksize_v : d.ksize_h;
    double filterscale, scale;
    filterscale = scale = __ddiv_rn((double)((float)in_size - 0.0f), (double)out_size);
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = __dmul_rn(2.0, filterscale);
    const double center = __dadd_rn(0.0, __dmul_rn(__dadd_rn((double)xx, 0.5), scale));
    const double ss = __ddiv_rn(1.0, filterscale);
    int xmin = __double2int_rz(__dadd_rn(__dadd_rn(center, -support), 0.5));
    if (xmin < 0) xmin = 0;
    int xmax = __double2int_rz(__dadd_rn(__dadd_rn(center, support), 0.5));
    if (xmax > in_size) xmax = in_size;
    xmax -= xmin;
    double ww = 0.0;
    for (int x = 0; x < xmax; ++x)
        ww = __dadd_rn(ww, bicubic_dev(__dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss)));
    int32_t* k = tab + (vert ? d.kv_off + (long long)i * ksize : d.kh_off + i);
    const long long step = vert ? 1 : crop;
    for (int x = 0; x < ksize; ++x) {
        int32_t q = 0;
        if (x < xmax) {
            double w = bicubic_dev(__dmul_rn(__dadd_rn(__dadd_rn((double)(x + xmin), -center), 0.5), ss));     // same bits as above
            if (ww != 0.0) w = __ddiv_rn(w, ww);
            const double sc = __dmul_rn(w, (double)(1 << PRECISION_BITS));
            q = w < 0 ? __double2int_rz(__dadd_rn(-0.5, sc)) : __double2int_rz(__dadd_rn(0.5, sc));
        }
        k[x * step] = q;
    }
    int32_t* b = tab + (vert ? d.bv_off : d.bh_off) + 2 * i;
    b[0] = xmin;
    b[1] = xmax;
}

struct Plan {
    std::vector<ImgDesc> desc;
    long long tab_ints = 0;
    long long tmp_bytes = 0;
    int max_rows = 0;
};

static int make_plan(const int64_t* offsets, const int32_t* heights, const int32_t* widths, int n, int size, int crop, int pad_sq, Plan& pl) {
    LB_REQUIRE(heights && widths && n > 0 && size > 0 && crop > 0 && crop <= size, LB_EINVAL,
               "clip_preprocess: need n > 0 images and 0 < crop <= size (got n=%d size=%d crop=%d)", n, size, crop);
    pl.desc.resize((size_t)n);
    long long ti = 0;
    for (int i = 0; i < n; ++i) {
        const int h = heights[i], w = widths[i];
        LB_REQUIRE(h > 0 && w > 0, LB_EINVAL, "clip_preprocess: image %d has size %d x %d", i, h, w);
        ImgDesc& d = pl.desc[(size_t)i];
        d.src_off = offsets ? offsets[i] : 0;
        d.h = h; d.w = w;
        int ch = h, cw = w;                                       // canvas
        d.pad_top = d.pad_left = 0;
        if (pad_sq && h != w) {
            const int s = h > w ? h : w;
            if (w > h) d.pad_top = (w - h) / 2; else d.pad_left = (h - w) / 2;
            ch = cw = s;
        }
        // get_resize_output_image_size(size=int, default_to_square=False)
        const int shrt = cw <= ch ? cw : ch, lng = cw <= ch ? ch : cw;
        const int new_long = (int)((double)size * (double)lng / (double)shrt);
        const int oh = cw <= ch ? new_long : size, ow = cw <= ch ? size : new_long;
        const int top = (oh - crop) / 2, left = (ow - crop) / 2;
        LB_REQUIRE(top >= 0 && left >= 0, LB_EINVAL, "clip_preprocess: crop %d exceeds the resized image %d x %d", crop, oh, ow);
        d.in_w = cw; d.in_h = ch; d.out_w = ow; d.out_h = oh; d.left = left; d.top = top;
        const Axis ah = axis_of(cw, ow), av = axis_of(ch, oh);
        d.ksize_h = ah.ksize; d.ksize_v = av.ksize;
        d.kh_off = (int)ti; ti += (long long)crop * ah.ksize;
        d.kv_off = (int)ti; ti += (long long)crop * av.ksize;
        d.bh_off = (int)ti; ti += 2 * crop;
        d.bv_off = (int)ti; ti += 2 * crop;
        LB_REQUIRE(ti < (1ll << 31), LB_EINVAL, "clip_preprocess: coefficient tables exceed 2^31 entries");
        int y0, n0, y1, n1;
        window_of(av, ch, top, &y0, &n0);
        window_of(av, ch, top + crop - 1, &y1, &n1);
        d.y_first = y0;
        d.rows_needed = y1 + n1 - y0;
        d.tmp_off = pl.tmp_bytes;
        pl.tmp_bytes += ((long long)d.rows_needed * crop * 3 + 255) / 256 * 256;
        if (d.rows_needed > pl.max_rows) pl.max_rows = d.rows_needed;
    }
    LB_REQUIRE(pl.max_rows <= 65535 * 16, LB_EINVAL, "clip_preprocess: %d source rows per image exceed the grid limit", pl.max_rows);
    pl.tab_ints = ti;
    return LB_OK;
}

static inline long long align256(long long x) { return (x + 255) / 256 * 256; }

// workspace: [descriptors | lut | table | intermediate rows] (descriptors and lut are uploaded together)
struct Layout {
    long long desc_off, tab_off, lut_off, tmp_off, total;
};
static Layout layout(const Plan& pl) {
    Layout L;
    L.desc_off = 0;
    L.lut_off = align256((long long)pl.desc.size() * (long long)sizeof(ImgDesc));
    L.tab_off = L.lut_off + align256(3 * 256 * 4);
    L.tmp_off = L.tab_off + align256(pl.tab_ints * 4);
    L.total = L.tmp_off + pl.tmp_bytes;
    return L;
}

__device__ __forceinline__ uint8_t clip8(int v) {
    v >>= PRECISION_BITS;
    return (uint8_t)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

constexpr int RPB_H = 16, RPB_V = 8;      // rows per block: a block's set-up (descriptor, bounds) is paid once per RPB rows

// pass 1: grid (ceil(crop / 128), ceil(max rows_needed / RPB_H), n_images), block 128; thread = crop column, loops over rows
__global__ void __launch_bounds__(128) resize_h_kernel(const Params p) {
    const ImgDesc d = p.desc[blockIdx.z];
    const int r0 = blockIdx.y * RPB_H, xx = blockIdx.x * 128 + threadIdx.x;
    if (r0 >= d.rows_needed || xx >= p.crop) return;
    const int32_t* b = p.tab + d.bh_off + 2 * xx;
    const int x0 = b[0] - d.pad_left, n = b[1];                   // first source column of the window (may lie on the canvas border)
    const int32_t* k = p.tab + d.kh_off + xx;                     // tap t at k[t * crop]
    const int r1 = min(r0 + RPB_H, d.rows_needed);
    const int bg0 = p.bg[0], bg1 = p.bg[1], bg2 = p.bg[2];
    // taps [t_in0, t_in1) fall on the pasted image, the others on the canvas colour
    const int t_in0 = min(n, max(0, -x0)), t_in1 = max(t_in0, min(n, d.w - x0));
    int kb = 0;                                                    // sum of the coefficients that see the canvas colour
    for (int t = 0; t < t_in0; ++t) kb += __ldg(k + (long long)t * p.crop);
    for (int t = t_in1; t < n; ++t) kb += __ldg(k + (long long)t * p.crop);
    int kall = kb;
    for (int t = t_in0; t < t_in1; ++t) kall += __ldg(k + (long long)t * p.crop);
    for (int r = r0; r < r1; ++r) {
        const int cy = d.y_first + r - d.pad_top;                 // source row of this canvas row (may lie outside)
        int s0, s1, s2;
        s0 = s1 = s2 = 1 << (PRECISION_BITS - 1);
        if (cy >= 0 && cy < d.h) {
            const uint8_t* px = p.images + d.src_off + ((long long)cy * d.w + (x0 + t_in0)) * 3;
#pragma unroll 4
            for (int t = t_in0; t < t_in1; ++t, px += 3) {
                const int kv = __ldg(k + (long long)t * p.crop);
                s0 += (int)__ldg(px) * kv; s1 += (int)__ldg(px + 1) * kv; s2 += (int)__ldg(px + 2) * kv;
            }
            s0 += bg0 * kb; s1 += bg1 * kb; s2 += bg2 * kb;
        } else {
            s0 += bg0 * kall; s1 += bg1 * kall; s2 += bg2 * kall;
        }
        uint8_t* o = p.tmp + d.tmp_off + ((long long)r * p.crop + xx) * 3;
        o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
    }
}

// pass 2: grid (ceil(crop / 128), ceil(crop / RPB_V), n_images), block 128; thread = crop column, loops over output rows
__global__ void __launch_bounds__(128) resize_v_kernel(const Params p) {
    const ImgDesc d = p.desc[blockIdx.z];
    const int xx = blockIdx.x * 128 + threadIdx.x;
    if (xx >= p.crop) return;
    const long long plane = (long long)p.crop * p.crop;
    const int y1 = min((int)(blockIdx.y + 1) * RPB_V, p.crop);
    for (int yy = blockIdx.y * RPB_V; yy < y1; ++yy) {
        const int32_t* b = p.tab + d.bv_off + 2 * yy;
        const int y0 = b[0] - d.y_first, n = b[1];
        const int32_t* k = p.tab + d.kv_off + (long long)yy * d.ksize_v;
        int s0, s1, s2;
        s0 = s1 = s2 = 1 << (PRECISION_BITS - 1);
        const uint8_t* px = p.tmp + d.tmp_off + ((long long)y0 * p.crop + xx) * 3;
#pragma unroll 4
        for (int t = 0; t < n; ++t, px += (long long)p.crop * 3) {
            const int kv = __ldg(k + t);
            s0 += (int)px[0] * kv; s1 += (int)px[1] * kv; s2 += (int)px[2] * kv;
        }
        const uint8_t u0 = clip8(s0), u1 = clip8(s1), u2 = clip8(s2);
        const long long pix = (long long)yy * p.crop + xx;
        const float f0 = __ldg(p.lut + u0), f1 = __ldg(p.lut + 256 + u1), f2 = __ldg(p.lut + 512 + u2);
        const long long base = (long long)blockIdx.z * 3 * plane + pix;
        if (p.out_bf16) {
            __nv_bfloat16* o = (__nv_bfloat16*)p.out + base;
            o[0] = __float2bfloat16_rn(f0); o[plane] = __float2bfloat16_rn(f1); o[2 * plane] = __float2bfloat16_rn(f2);
        } else {
            float* o = (float*)p.out + base;
            o[0] = f0; o[plane] = f1; o[2 * plane] = f2;
        }
        if (p.out_u8) {
            uint8_t* o = p.out_u8 + ((long long)blockIdx.z * plane + pix) * 3;
            o[0] = u0; o[1] = u1; o[2] = u2;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Word-load variants (the default when crop % 4 == 0).  The byte-per-load kernels above are bound by the load/store unit (20
// LSU instructions per output pixel at 5 taps); here a thread fetches its taps as aligned 32-bit words:
//   pass 1: 4 taps = 12 source bytes = 4 aligned words funnel-shifted to the window's first byte (the packed buffer must
//     be readable 16 bytes past its last image: the last chunk of the last row over-reads, multiplied by zero coefficients);
//   pass 2: the filter is the same for every byte of a row, so a thread owns 4 consecutive BYTES of the output row
//     (crop * 3 bytes, a multiple of 4) and reads one word per tap; byte f maps to pixel f / 3, channel f % 3 on output.
// Same integer sums in the same 32-bit arithmetic: bit-identical results (tests/test_gpu_preprocess.py runs both).
__global__ void __launch_bounds__(128) resize_h_words_kernel(const Params p) {
    const ImgDesc d = p.desc[blockIdx.z];
    const int r0 = blockIdx.y * RPB_H, xx = blockIdx.x * 128 + threadIdx.x;
    if (r0 >= d.rows_needed || xx >= p.crop) return;
    const int32_t* b = p.tab + d.bh_off + 2 * xx;
    const int x0 = b[0] - d.pad_left, n = b[1];
    const int32_t* k = p.tab + d.kh_off + xx;                     // tap t at k[t * crop]
    const int r1 = min(r0 + RPB_H, d.rows_needed);
    const int bg0 = p.bg[0], bg1 = p.bg[1], bg2 = p.bg[2];
    const int t_in0 = min(n, max(0, -x0)), t_in1 = max(t_in0, min(n, d.w - x0));
    int kb = 0;
    for (int t = 0; t < t_in0; ++t) kb += __ldg(k + (long long)t * p.crop);
    for (int t = t_in1; t < n; ++t) kb += __ldg(k + (long long)t * p.crop);
    int kall = kb;
    for (int t = t_in0; t < t_in1; ++t) kall += __ldg(k + (long long)t * p.crop);
    const uint8_t* const base = p.images + d.src_off;
    for (int r = r0; r < r1; ++r) {
        const int cy = d.y_first + r - d.pad_top;
        int s0, s1, s2;
        s0 = s1 = s2 = 1 << (PRECISION_BITS - 1);
        if (cy >= 0 && cy < d.h) {
            for (int t = t_in0; t < t_in1; t += 4) {
                const uintptr_t A = (uintptr_t)(base + ((long long)cy * d.w + (x0 + t)) * 3);
                const uint32_t* wp = reinterpret_cast<const uint32_t*>(A & ~(uintptr_t)3);
                const uint32_t sh = (uint32_t)(A & 3) * 8;
                const uint32_t w0 = __ldg(wp), w1 = __ldg(wp + 1), w2 = __ldg(wp + 2), w3 = __ldg(wp + 3);
                const uint32_t a0 = __funnelshift_r(w0, w1, sh), a1 = __funnelshift_r(w1, w2, sh), a2 = __funnelshift_r(w2, w3, sh);
                const int k0 = __ldg(k + (long long)t * p.crop);
                const int k1 = t + 1 < t_in1 ? __ldg(k + (long long)(t + 1) * p.crop) : 0;
                const int k2 = t + 2 < t_in1 ? __ldg(k + (long long)(t + 2) * p.crop) : 0;
                const int k3 = t + 3 < t_in1 ? __ldg(k + (long long)(t + 3) * p.crop) : 0;
                // bytes 0..11 of the window: tap q, channel c = byte 3 q + c
                s0 += (int)(a0 & 0xff) * k0;         s1 += (int)((a0 >> 8) & 0xff) * k0;  s2 += (int)((a0 >> 16) & 0xff) * k0;
                s0 += (int)(a0 >> 24) * k1;          s1 += (int)(a1 & 0xff) * k1;         s2 += (int)((a1 >> 8) & 0xff) * k1;
                s0 += (int)((a1 >> 16) & 0xff) * k2; s1 += (int)(a1 >> 24) * k2;          s2 += (int)(a2 & 0xff) * k2;
                s0 += (int)((a2 >> 8) & 0xff) * k3;  s1 += (int)((a2 >> 16) & 0xff) * k3; s2 += (int)(a2 >> 24) * k3;
            }
            s0 += bg0 * kb; s1 += bg1 * kb; s2 += bg2 * kb;
        } else {
            s0 += bg0 * kall; s1 += bg1 * kall; s2 += bg2 * kall;
        }
        uint8_t* o = p.tmp + d.tmp_off + ((long long)r * p.crop + xx) * 3;
        o[0] = clip8(s0); o[1] = clip8(s1); o[2] = clip8(s2);
    }
}

// grid (ceil(crop * 3 / 4 / 128), ceil(crop / RPB_V), n_images), block 128; thread = 4 consecutive bytes of the output row
__global__ void __launch_bounds__(128) resize_v_words_kernel(const Params p) {
    const ImgDesc d = p.desc[blockIdx.z];
    const int P = p.crop * 3;                                      // bytes per row (multiple of 4)
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (4 * j >= P) return;
    const long long plane = (long long)p.crop * p.crop;
    const int y1 = min((int)(blockIdx.y + 1) * RPB_V, p.crop);
    // byte f = 4 j + i of the row is channel f % 3 of pixel f / 3
    int px[4], chn[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { px[i] = (4 * j + i) / 3; chn[i] = (4 * j + i) - 3 * px[i]; }
    for (int yy = blockIdx.y * RPB_V; yy < y1; ++yy) {
        const int32_t* b = p.tab + d.bv_off + 2 * yy;
        const int y0 = b[0] - d.y_first, n = b[1];
        const int32_t* k = p.tab + d.kv_off + (long long)yy * d.ksize_v;
        int s[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) s[i] = 1 << (PRECISION_BITS - 1);
        const uint8_t* src = p.tmp + d.tmp_off + (long long)y0 * P + 4 * j;
#pragma unroll 4
        for (int t = 0; t < n; ++t, src += P) {
            const int kv = __ldg(k + t);
            const uint32_t w = *reinterpret_cast<const uint32_t*>(src);
            s[0] += (int)(w & 0xff) * kv; s[1] += (int)((w >> 8) & 0xff) * kv; s[2] += (int)((w >> 16) & 0xff) * kv; s[3] += (int)(w >> 24) * kv;
        }
        uint32_t packed = 0;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t u = clip8(s[i]);
            packed |= u << (8 * i);
            const float f = __ldg(p.lut + chn[i] * 256 + u);
            const long long o = ((long long)blockIdx.z * 3 + chn[i]) * plane + (long long)yy * p.crop + px[i];
            if (p.out_bf16) ((__nv_bfloat16*)p.out)[o] = __float2bfloat16_rn(f); else ((float*)p.out)[o] = f;
        }
        if (p.out_u8) *reinterpret_cast<uint32_t*>(p.out_u8 + ((long long)blockIdx.z * plane + (long long)yy * p.crop) * 3 + 4 * j) = packed;
    }
}

}  // namespace pp
}  // namespace lb

using namespace lb;

extern "C" {

int lb_clip_resample_coeffs(int in_size, int out_size, int first_out, int n_out, int32_t* coeffs, int coeffs_capacity, int32_t* bounds) {
    LB_REQUIRE(in_size > 0 && out_size > 0 && first_out >= 0 && n_out > 0 && first_out + n_out <= out_size && bounds, LB_EINVAL,
               "clip_resample_coeffs: bad range");
    std::vector<int32_t> kk, bb;
    const int ksize = pp::coeffs(in_size, out_size, first_out, n_out, kk, bb);
    if (coeffs) {
        LB_REQUIRE((int64_t)coeffs_capacity >= (int64_t)kk.size(), LB_EINVAL, "clip_resample_coeffs: %zu coefficients, capacity %d",
                   kk.size(), coeffs_capacity);
        memcpy(coeffs, kk.data(), kk.size() * 4);
    }
    memcpy(bounds, bb.data(), bb.size() * 4);
    return ksize;
}

int64_t lb_clip_preprocess_workspace(const int32_t* heights, const int32_t* widths, int n_images, int size, int crop, int pad_to_square) {
    pp::Plan pl;
    if (pp::make_plan(nullptr, heights, widths, n_images, size, crop, pad_to_square, pl)) return -1;
    return pp::layout(pl).total;
}

int lb_clip_preprocess(const uint8_t* images, const int64_t* offsets, const int32_t* heights, const int32_t* widths, int n_images,
                       int size, int crop, int pad_to_square, const uint8_t* pad_rgb, const float* mean, const float* std,
                       double rescale_factor, void* out, int out_dtype, uint8_t* out_u8, void* workspace, int64_t workspace_bytes,
                       void* stream) {
    int rc = require_sm100();
    if (rc) return rc;
    LB_REQUIRE(images && offsets && mean && std && out && workspace, LB_EINVAL, "clip_preprocess: null argument");
    LB_REQUIRE(out_dtype == LB_DT_BF16 || out_dtype == LB_DT_F32, LB_EDTYPE, "clip_preprocess: out_dtype %d", out_dtype);
    LB_REQUIRE(!pad_to_square || pad_rgb, LB_EINVAL, "clip_preprocess: pad_to_square needs pad_rgb");
    pp::Plan pl;
    rc = pp::make_plan(offsets, heights, widths, n_images, size, crop, pad_to_square, pl);
    if (rc) return rc;
    const pp::Layout L = pp::layout(pl);
    LB_REQUIRE(workspace_bytes >= L.total, LB_EINVAL, "clip_preprocess: workspace of %lld bytes needed, %lld given", L.total,
               (long long)workspace_bytes);
    // the reference's rounding sequence: uint8 -> float64 * scale -> float32; (x - mean) / std in float32
    float lut[3 * 256];
    for (int c = 0; c < 3; ++c)
        for (int v = 0; v < 256; ++v) {
            const float x = (float)((double)v * rescale_factor);
            lut[c * 256 + v] = (x - mean[c]) / std[c];
        }
    // descriptors + value table: one small host block -> one copy (pageable memory: staged by the runtime before the call returns)
    std::vector<uint8_t> host((size_t)L.tab_off, 0);
    memcpy(host.data() + L.desc_off, pl.desc.data(), pl.desc.size() * sizeof(pp::ImgDesc));
    memcpy(host.data() + L.lut_off, lut, sizeof(lut));
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaMemcpyAsync(workspace, host.data(), (size_t)L.tab_off, cudaMemcpyHostToDevice, st);
    if (e != cudaSuccess) return fail(LB_ELAUNCH, "clip_preprocess: table upload: %s", cudaGetErrorString(e));
    pp::Params p;
    p.images = images;
    p.desc = (const pp::ImgDesc*)((char*)workspace + L.desc_off);
    p.tab = (const int32_t*)((char*)workspace + L.tab_off);
    p.lut = (const float*)((char*)workspace + L.lut_off);
    p.tmp = (uint8_t*)workspace + L.tmp_off;
    p.out = out;
    p.out_u8 = out_u8;
    p.crop = crop;
    p.out_bf16 = out_dtype == LB_DT_BF16;
    for (int c = 0; c < 3; ++c) p.bg[c] = pad_rgb ? pad_rgb[c] : 0;
    const unsigned gx = (unsigned)ceil_div(crop, 128);
    pp::tables_kernel<<<dim3(gx, 2, (unsigned)n_images), 128, 0, st>>>(p.desc, (int32_t*)((char*)workspace + L.tab_off), crop);
    static int force_bytes = -1;                                  // LB_PREPROC_BYTES=1: the byte-per-load kernels (A/B, cross-check)
    if (force_bytes < 0) { const char* e = getenv("LB_PREPROC_BYTES"); force_bytes = (e && atoi(e) == 1) ? 1 : 0; }
    if (crop % 4 == 0 && !force_bytes) {
        pp::resize_h_words_kernel<<<dim3(gx, (unsigned)ceil_div(pl.max_rows, pp::RPB_H), (unsigned)n_images), 128, 0, st>>>(p);
        pp::resize_v_words_kernel<<<dim3((unsigned)ceil_div(crop * 3 / 4, 128), (unsigned)ceil_div(crop, pp::RPB_V), (unsigned)n_images), 128, 0, st>>>(p);
    } else {
        pp::resize_h_kernel<<<dim3(gx, (unsigned)ceil_div(pl.max_rows, pp::RPB_H), (unsigned)n_images), 128, 0, st>>>(p);
        pp::resize_v_kernel<<<dim3(gx, (unsigned)ceil_div(crop, pp::RPB_V), (unsigned)n_images), 128, 0, st>>>(p);
    }
    return check_launch("clip_preprocess");
}

}

// Host-side plumbing of the C ABI: thread-local error text, device checks and
// TMA tensor-map encoding through the driver entry point.
#include <stdlib.h>
#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <mutex>

#include "common.cuh"

namespace lb {

static thread_local char g_err[512] = {0};

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}

int attn_head_group() {
    static int g = -1;
    if (g < 0) {
        const char* e = getenv("LB_ATTN_HEAD_GROUP");
        g = e ? atoi(e) : 8;
        if (g < 1) g = 1;
    }
    return g;
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(LB_ELAUNCH, "%s: %s", what, cudaGetErrorString(e));
    return LB_OK;
}

int g_pdl = 0;
bool pdl_on() { return g_pdl != 0; }

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

int require_sm100() {
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return fail(LB_ELAUNCH, "cudaGetDevice: %s", cudaGetErrorString(e));
    int major = 0;
    e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
    if (e != cudaSuccess) return fail(LB_ELAUNCH, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
    if (major != 10) return fail(LB_EARCH, "libra_b200 needs an sm_100 device (compute capability 10.x), got %d.x", major);
    return LB_OK;
}

typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
    static PFN_encodeTiled fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_encodeTiled>(p);
    });
    return fn;
}

int make_tmap_bf16_nd(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                      const uint32_t* box, int swizzle128) {
    PFN_encodeTiled enc = get_encode();
    if (!enc) return fail(LB_EDRIVER, "cuTensorMapEncodeTiled entry point unavailable");
    if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(LB_EALIGN, "TMA base pointer must be 16-byte aligned");
    cuuint64_t gdim[5];
    cuuint64_t gstr[4];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; ++i) {
        gdim[i] = dims[i];
        bx[i] = box[i];
        es[i] = 1;
        if (i > 0) {
            gstr[i - 1] = strides_bytes[i - 1];
            if (gstr[i - 1] % 16) return fail(LB_EALIGN, "TMA stride %d (%llu B) must be a multiple of 16", i,
                                              (unsigned long long)gstr[i - 1]);
        }
    }
    CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_NONE,
                     CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail(LB_EDRIVER, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
    return LB_OK;
}

// Tensor-map cache.  cuTensorMapEncodeTiled costs ~1 us and the same (pointer, shape, pitch, box) tuples recur every step
// (weights always, activations through the caching allocator), so encoded 2-D maps are kept in a direct-mapped table: a
// steady-state training step performs no encode at all.  A map is a pure function of its key, so a stale entry cannot exist.
struct TmapKey {
    const void* base;
    uint64_t d0, d1, ld;
    uint32_t b0, b1;
};
struct TmapEntry {
    TmapKey key;
    CUtensorMap map;
    bool valid;
};
static constexpr int TMAP_CACHE = 16384;
static TmapEntry* g_tmap_cache = nullptr;
static std::mutex g_tmap_mu;
static uint64_t g_tmap_hits = 0, g_tmap_misses = 0;

int tmap_cache_stats(int64_t* hits, int64_t* misses) {
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    if (hits) *hits = (int64_t)g_tmap_hits;
    if (misses) *misses = (int64_t)g_tmap_misses;
    return LB_OK;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols) {
    TmapKey k;
    memset(&k, 0, sizeof(k));
    k.base = base; k.d0 = cols; k.d1 = rows; k.ld = ld_elems; k.b0 = box_cols; k.b1 = box_rows;
    uint64_t h = (uint64_t)(uintptr_t)base * 0x9E3779B97F4A7C15ull;
    h ^= (rows * 0xC2B2AE3D27D4EB4Full) ^ (cols * 0x165667B19E3779F9ull) ^ (ld_elems << 17) ^ ((uint64_t)box_rows << 40) ^ box_cols;
    h ^= h >> 29;
    const int slot = (int)(h % TMAP_CACHE);
    {
        std::lock_guard<std::mutex> lk(g_tmap_mu);
        if (!g_tmap_cache) g_tmap_cache = (TmapEntry*)calloc(TMAP_CACHE, sizeof(TmapEntry));
        TmapEntry& e = g_tmap_cache[slot];
        if (e.valid && memcmp(&e.key, &k, sizeof(k)) == 0) {
            *out = e.map;
            ++g_tmap_hits;
            return LB_OK;
        }
        ++g_tmap_misses;
    }
    uint64_t dims[2] = {cols, rows};
    uint64_t strides[1] = {ld_elems * 2};
    uint32_t box[2] = {box_cols, box_rows};
    int rc = make_tmap_bf16_nd(out, base, 2, dims, strides, box, 1);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    TmapEntry& e = g_tmap_cache[slot];
    e.key = k;
    e.map = *out;
    e.valid = true;
    return LB_OK;
}

// 3-D view [batch][rows][cols] of a row-major bf16 matrix [batch*rows, cols] (box = [box_rows, box_cols] of one sample,
// SWIZZLE_128B): TMA stores clip at the END OF THE SAMPLE, which a 2-D [batch*rows, cols] map cannot do.  Cached like the 2-D maps.
int make_tmap_bf16_3d(CUtensorMap* out, const void* base, uint64_t batch, uint64_t rows, uint64_t cols, uint32_t box_rows,
                      uint32_t box_cols) {
    TmapKey k;
    memset(&k, 0, sizeof(k));
    k.base = base; k.d0 = cols; k.d1 = rows; k.ld = (batch << 20) | 0x3D3D3ull | (1ull << 63); k.b0 = box_cols; k.b1 = box_rows;
    uint64_t h = (uint64_t)(uintptr_t)base * 0x9E3779B97F4A7C15ull;
    h ^= (rows * 0xC2B2AE3D27D4EB4Full) ^ (cols * 0x165667B19E3779F9ull) ^ (k.ld << 17) ^ ((uint64_t)box_rows << 40) ^ box_cols;
    h ^= h >> 29;
    const int slot = (int)(h % TMAP_CACHE);
    {
        std::lock_guard<std::mutex> lk(g_tmap_mu);
        if (!g_tmap_cache) g_tmap_cache = (TmapEntry*)calloc(TMAP_CACHE, sizeof(TmapEntry));
        TmapEntry& e = g_tmap_cache[slot];
        if (e.valid && memcmp(&e.key, &k, sizeof(k)) == 0) {
            *out = e.map;
            ++g_tmap_hits;
            return LB_OK;
        }
        ++g_tmap_misses;
    }
    uint64_t dims[3] = {cols, rows, batch};
    uint64_t strides[2] = {cols * 2, rows * cols * 2};
    uint32_t box[3] = {box_cols, box_rows, 1};
    int rc = make_tmap_bf16_nd(out, base, 3, dims, strides, box, 1);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(g_tmap_mu);
    TmapEntry& e = g_tmap_cache[slot];
    e.key = k;
    e.map = *out;
    e.valid = true;
    return LB_OK;
}

}  // namespace lb

extern "C" {

int lb_version(void) { return 100; }

int lb_set_pdl(int on) {
    const int prev = lb::g_pdl;
    lb::g_pdl = on ? 1 : 0;
    return prev;
}

int lb_last_error(char* buf, int n) {
    if (!buf || n <= 0) return LB_EINVAL;
    strncpy(buf, lb::g_err, (size_t)n - 1);
    buf[n - 1] = 0;
    return LB_OK;
}

int lb_device_check(void) { return lb::require_sm100(); }

int lb_sm_count(void) { return lb::sm_count(); }

}

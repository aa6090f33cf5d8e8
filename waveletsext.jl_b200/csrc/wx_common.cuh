// wx_common.cuh -- shared host/device helpers for libwx_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdarg>
#include <cstdint>
#include <cstring>
#include <atomic>
#include "../../include/wx_b200.h"

// ----------------------------------------------------------------------------------------------
// error plumbing: no exceptions across the C ABI
// ----------------------------------------------------------------------------------------------
extern thread_local char wx_errbuf[512];
extern std::atomic<unsigned long long> wx_launches;

static inline int wx_fail(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(wx_errbuf, sizeof(wx_errbuf), fmt, ap);
    va_end(ap);
    return code;
}

#define WX_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess)                                                                 \
            return wx_fail(WX_ECUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),   \
                           __FILE__, __LINE__);                                                \
    } while (0)

#define WX_REQUIRE(cond, ...)                          \
    do {                                               \
        if (!(cond)) return wx_fail(WX_EINVAL, __VA_ARGS__); \
    } while (0)

#define WX_LAUNCHED()                                   \
    do {                                                \
        wx_launches.fetch_add(1, std::memory_order_relaxed); \
        WX_CUDA(cudaGetLastError());                    \
    } while (0)

// ----------------------------------------------------------------------------------------------
// taps travel by value in kernel parameter space (constant bank): compile-time tap indices in the
// unrolled kernels become c[bank][offset] operands of the FMAs.
// ----------------------------------------------------------------------------------------------
template <typename T>
struct Taps {
    T h[WX_MAX_TAPS];   // detail  (mirror qmf)      / Q for the autocorrelation transforms
    T g[WX_MAX_TAPS];   // scaling (reversed qmf)    / P
    int F;
};

template <typename T>
static inline int wx_make_taps(Taps<T> &t, const double *h, const double *g, int F)
{
    if (!h || !g) return wx_fail(WX_EINVAL, "null filter taps");
    if (F < 1 || F > WX_MAX_TAPS) return wx_fail(WX_EUNSUPPORTED, "filter length %d outside 1..%d", F, WX_MAX_TAPS);
    memset(&t, 0, sizeof(t));
    for (int i = 0; i < F; ++i) { t.h[i] = (T)h[i]; t.g[i] = (T)g[i]; }
    t.F = F;
    return WX_OK;
}

static inline int wx_ilog2l(long i) { int d = 0; while (i > 1) { i >>= 1; ++d; } return d; }
static inline bool wx_ispow2(long v) { return v > 0 && (v & (v - 1)) == 0; }
static inline int wx_maxlevels(long n) { int k = 0; if (n <= 0) return 0; while ((n & 1) == 0) { n >>= 1; ++k; } return k; }
// getdepth(i,:quad)  utils/utils_tree.jl:257-259, integer version
static inline int wx_quaddepthl(long i) { int d = 0; long last = 1, w = 1; while (i > last) { w *= 4; last += w; ++d; } return d; }
// getrowrange/getcolrange  Utils.jl:465-542 (0-based start, extents)
static inline void wx_quadrangel(long m, long n, long idx, long *r0, long *c0, long *nr, long *nc)
{
    if (idx == 1) { *r0 = 0; *c0 = 0; *nr = m; *nc = n; return; }
    long parent = (idx + 2) / 4, pr0, pc0, pnr, pnc;
    wx_quadrangel(m, n, parent, &pr0, &pc0, &pnr, &pnc);
    *nr = pnr / 2; *nc = pnc / 2;
    *r0 = (idx < 4 * parent) ? pr0 : pr0 + pnr / 2;
    *c0 = (idx % 2 == 0) ? pc0 : pc0 + pnc / 2;
}

struct WxDev { int sms; size_t smem_optin; int dev; cudaMemPool_t pool; };
int wx_devinfo(WxDev &d);   // cached per device (thread safe); creates the library's private scratch pool on first use
int wx_pool_alloc(void **p, size_t bytes, cudaStream_t s);   // stream-ordered allocation from that pool (free: cudaFreeAsync)

// ----------------------------------------------------------------------------------------------
// Residency tuning of the persistent kernels (wx_runtime.cu).  The write-dominated kernels of this library are sensitive to
// the number of CTAs resident per SM (concurrent store streams against warps in flight); the optimum moves with filter length,
// element type and node size (profiles/r2_wpd1d_residency_sweep.jsonl), so the first LARGE launch of a shape measures the
// candidates on the caller's own data (every candidate launch is a complete, correct run of the kernel) and the choice is
// cached per (kernel, shape).  Small launches, captured streams and WX_B200_AUTOTUNE=0 use the caller's rule instead.
// ----------------------------------------------------------------------------------------------
#include <functional>
struct WxTuneKey { const void *kernel; long a, b, c, d; };
int wx_tuned_choice(const WxTuneKey &key, int ncand, const int *cand, int fallback, bool big_enough, cudaStream_t s,
                    const std::function<int(int)> &launch, int *choice);

// ----------------------------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------------------------
// periodic wrap of an index that may be negative or exceed the period several times
__device__ __forceinline__ long wx_wrapl(long a, long p) { long r = a % p; return r < 0 ? r + p : r; }
__device__ __forceinline__ int wx_wrapi(int a, int p) { int r = a % p; return r < 0 ? r + p : r; }

// 16-byte chunk swizzle used by every kernel that keeps a signal in shared memory:
// chunk c -> c ^ ((c >> 3) & 7)  (permutes the eight 16 B chunks inside each 128 B row; identical to the
// TMA SWIZZLE_128B pattern).  It makes "thread t reads chunks 4t..4t+7" and "thread t writes chunks
// 2t, 2t+1" both bank-conflict free.
__device__ __forceinline__ int wx_swz_chunk(int c) { return c ^ ((c >> 3) & 7); }

template <typename T> struct WxVec;
template <> struct WxVec<double> { static constexpr int N = 2; using type = double2; };
template <> struct WxVec<float> { static constexpr int N = 4; using type = float4; };

template <typename T>
__device__ __forceinline__ int wx_swz_elem(int e)
{
    constexpr int V = WxVec<T>::N;
    return wx_swz_chunk(e / V) * V + (e % V);
}

// wx_denoise.cu -- the step on the far side of the path in the reference's pipeline: threshold
// determination and thresholding of the expansion coefficients between getbasiscoefall and the inverse transform.
//   noisest            Denoising.jl:214-232  (Wavelets.Threshold.mad!: median of |y - median(y)|, / 0.6745)
//   surethreshold      Denoising.jl:142-166
//   relerrorthreshold  Denoising.jl:285-328 with orth2relerror :344-349 and findelbow :366-381
//   threshold!         Wavelets.jl Threshold.jl (HardTH / SoftTH / SemiSoftTH / SteinTH), as called by denoise :483-600
// One CTA per signal for the order statistics: the signal's coefficients are sorted by a bitonic network in shared
// memory (or, when they do not fit, in a global scratch slab), everything after the sort is a block scan / block argmin.
// The thresholding itself is one streaming pass (2 s bytes per coefficient).
#include "wx_steps.cuh"
#include "wx_2d.cuh"
#include <vector>
#include <limits>

namespace {

constexpr int kTS = 1024;                      // threads of the per-signal kernels

template <typename T> __device__ __forceinline__ T wx_inf();
template <> __device__ __forceinline__ double wx_inf<double>() { return __longlong_as_double(0x7ff0000000000000LL); }
template <> __device__ __forceinline__ float wx_inf<float>() { return __int_as_float(0x7f800000); }

// ascending bitonic sort of a[0..P), P a power of two, by all threads of the CTA (a in shared or global memory)
template <typename T>
__device__ void bitonic_sort(T *a, int P)
{
    const int half = P >> 1;
    for (int k = 2; k <= P; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < half; t += blockDim.x) {
                const int i = 2 * t - (t & (j - 1)), l = i + j;
                const bool up = (i & k) == 0;
                const T x = a[i], y = a[l];
                if ((x > y) == up) { a[i] = y; a[l] = x; }
            }
            __syncthreads();
        }
}

// median of the sorted a[0..M): Statistics.median! -> middle(a, b) = a/2 + b/2 for even M
template <typename T>
__device__ __forceinline__ T sorted_median(const T *a, int M)
{
    return (M & 1) ? a[M >> 1] : a[(M >> 1) - 1] / (T)2 + a[M >> 1] / (T)2;
}

// sigma[k] = mad(x[off .. off+len) of signal k) / 0.6745
template <typename T>
__global__ void __launch_bounds__(kTS) mad_k(double *__restrict__ sigma, const T *__restrict__ x, long stride, long off, int len, int P, T *gbuf)
{
    extern __shared__ __align__(16) unsigned char wx_dn_smem[];
    T *a = gbuf ? gbuf + (long)blockIdx.x * P : reinterpret_cast<T *>(wx_dn_smem);
    const T *src = x + (long)blockIdx.x * stride + off;
    for (int i = threadIdx.x; i < P; i += blockDim.x) a[i] = i < len ? src[i] : wx_inf<T>();
    __syncthreads();
    bitonic_sort(a, P);
    const T m = sorted_median(a, len);
    __syncthreads();
    for (int i = threadIdx.x; i < len; i += blockDim.x) a[i] = fabs(a[i] - m);
    __syncthreads();
    bitonic_sort(a, P);
    if (threadIdx.x == 0) sigma[blockIdx.x] = (double)sorted_median(a, len) / 0.6745;
}

// srt (N, P): |coefficients| of the selected columns of each signal's (n, K) slab, ascending, padded with +inf
template <typename T>
__global__ void __launch_bounds__(kTS) sort_abs_k(T *__restrict__ srt, const T *__restrict__ x, long slab, int n, const int *__restrict__ cols, int M, int P,
                                                 int in_smem)
{
    extern __shared__ __align__(16) unsigned char wx_dn_smem[];
    T *out = srt + (long)blockIdx.x * P;
    T *a = in_smem ? reinterpret_cast<T *>(wx_dn_smem) : out;
    const T *src = x + (long)blockIdx.x * slab;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        T v = wx_inf<T>();
        if (i < M) { const int c = i / n, r = i - c * n; v = fabs(src[(long)(cols ? cols[c] : c) * n + r]); }
        a[i] = v;
    }
    __syncthreads();
    bitonic_sort(a, P);
    if (in_smem) for (int i = threadIdx.x; i < M; i += blockDim.x) out[i] = a[i];
}

// ---- block collectives (blockDim.x = kTS) ------------------------------------------------------------------
__device__ __forceinline__ double warp_incl_scan(double v)
{
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_up_sync(0xffffffffu, v, o); if ((threadIdx.x & 31) >= o) v += u; }
    return v;
}
// exclusive prefix of one value per thread (thread order) and the block total; sh holds >= 33 doubles
__device__ double block_excl_scan(double v, double *sh, double *total)
{
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    const double inc = warp_incl_scan(v);
    __syncthreads();
    if (lane == 31) sh[w] = inc;
    __syncthreads();
    if (w == 0) {
        const double s = lane < nw ? sh[lane] : 0.0;
        const double si = warp_incl_scan(s);
        sh[lane] = si - s;
        if (lane == 31) sh[32] = si;
    }
    __syncthreads();
    const double r = sh[w] + inc - v;
    *total = sh[32];
    return r;
}
// arg-extremum with the smallest index among equal values (findmax / argmin return the first)
template <bool MAX>
__device__ void block_argext(double &v, int &idx, double *shv, int *shi)
{
    auto better = [](double a, int ia, double b, int ib) {
        if (ia < 0) return false;
        if (ib < 0) return true;
        return MAX ? (a > b || (a == b && ia < ib)) : (a < b || (a == b && ia < ib));
    };
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_down_sync(0xffffffffu, v, o);
        const int oi = __shfl_down_sync(0xffffffffu, idx, o);
        if (better(ov, oi, v, idx)) { v = ov; idx = oi; }
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane == 0) { shv[w] = v; shi[w] = idx; }
    __syncthreads();
    if (w == 0) {
        v = lane < nw ? shv[lane] : 0.0; idx = lane < nw ? shi[lane] : -1;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            const double ov = __shfl_down_sync(0xffffffffu, v, o);
            const int oi = __shfl_down_sync(0xffffffffu, idx, o);
            if (better(ov, oi, v, idx)) { v = ov; idx = oi; }
        }
        if (lane == 0) { shv[0] = v; shi[0] = idx; }
    }
    __syncthreads();
    v = shv[0]; idx = shi[0];
    __syncthreads();
}

// surethreshold :158-165 on the sorted magnitudes: a = sorted.^2, b = cumsum(a), risk_i = (M - 2i + b_i + (M-i) a_i)/M, sqrt(a[argmin])
template <typename T>
__global__ void __launch_bounds__(kTS) sure_k(double *__restrict__ t, const T *__restrict__ srt, int M, int P)
{
    __shared__ double shv[40];
    __shared__ int shi[32];
    const T *a = srt + (long)blockIdx.x * P;
    const int per = (M + blockDim.x - 1) / blockDim.x;
    const int lo = min((int)threadIdx.x * per, M), hi = min(lo + per, M);
    double s = 0;
    for (int i = lo; i < hi; ++i) { const double v = (double)a[i]; s += v * v; }
    double total;
    double b = block_excl_scan(s, shv, &total);
    double best = 0; int bi = -1;
    for (int i = lo; i < hi; ++i) {
        const double v = (double)a[i], av = v * v;
        b += av;
        const double risk = ((double)(M - 2 * (i + 1)) + (b + (double)(M - 1 - i) * av)) / (double)M;
        if (bi < 0 || risk < best) { best = risk; bi = i; }
    }
    block_argext<false>(best, bi, shv, shi);
    if (threadIdx.x == 0) { const double v = (double)a[bi]; t[blockIdx.x] = sqrt(v * v); }
}

// relerrorthreshold :301-327.  Points p = 0..M of the curve: X_p = (p ? asc[p-1] : 0)/xmax, Y_p = r[M-p]/ymax (Y_M = r[1]/ymax),
// r_j = sqrt|S - cum_j| / sqrt(S) with cum the running sum of the squares in DESCENDING order (orth2relerror).  Yw (N, M+1) scratch.
template <typename T>
__global__ void __launch_bounds__(kTS) relerr_k(double *__restrict__ t, const T *__restrict__ srt, double *__restrict__ Yw, int M, int P, int elbows)
{
    __shared__ double shv[40];
    __shared__ int shi[32];
    const T *asc = srt + (long)blockIdx.x * P;
    double *Y = Yw + (long)blockIdx.x * (M + 1);
    const int per = (M + blockDim.x - 1) / blockDim.x;
    const int lo = min((int)threadIdx.x * per, M), hi = min(lo + per, M);       // descending positions j-1 = lo..hi-1  <->  asc[M-1-(j-1)]
    double s = 0;
    for (int q = lo; q < hi; ++q) { const double v = (double)asc[M - 1 - q]; s += v * v; }
    double S;
    double cum = block_excl_scan(s, shv, &S);
    const double rS = sqrt(S);
    double ymax = 0; int yi = -1;
    for (int q = lo; q < hi; ++q) {
        const double v = (double)asc[M - 1 - q];
        cum += v * v;
        const double r = sqrt(fabs(S - cum)) / rS;           // r_{q+1}
        Y[M - 1 - q] = r;                                    // point p = M - j
        if (q == 0) Y[M] = r;
        if (yi < 0 || r > ymax) { ymax = r; yi = q; }
    }
    block_argext<true>(ymax, yi, shv, shi);                  // also orders the Y writes before the reads below
    const double xmax = (double)asc[M - 1];
    const double x0 = 0.0 / xmax, y0 = Y[0] / ymax;
    int end = M;                                             // last point of the current curve
    for (int e = 0; e < elbows; ++e) {
        const double xe = (double)asc[end > 0 ? end - 1 : 0] / xmax;
        double vx = (end > 0 ? xe : x0) - x0, vy = Y[end] / ymax - y0;
        const double nv = sqrt(vx * vx + vy * vy);
        vx /= nv; vy /= nv;
        double best = 0; int bi = -1;
        for (int p = threadIdx.x; p <= end; p += blockDim.x) {
            const double dx = (p ? (double)asc[p - 1] / xmax : x0) - x0, dy = Y[p] / ymax - y0;
            const double H = sqrt(dx * dx + dy * dy), A = dx * vx + dy * vy;
            const double O = sqrt(fabs(H * H - A * A));
            // findmax skips nothing: a NaN compares false everywhere, the first element wins then
            if (bi < 0 || O > best) { best = O; bi = p; }
        }
        block_argext<true>(best, bi, shv, shi);
        end = bi < 0 ? 0 : bi;
    }
    if (threadIdx.x == 0) t[blockIdx.x] = ((end ? (double)asc[end - 1] : 0.0) / xmax) * xmax;
}

// ---- thresholding ----------------------------------------------------------------------------------------
// th: 0 hard, 1 soft, 2 semisoft, 3 stein -- arithmetic in Float64 (the threshold is a Float64 in the reference), stored as T
template <typename T>
__device__ __forceinline__ T apply_th(T xv, double t, int th)
{
    const double x = (double)xv;
    switch (th) {
    case 0: return fabs(x) <= t ? (T)0 : xv;
    case 1: { const double sh = fabs(x) - t; return sh < 0 ? (T)0 : (T)((x > 0 ? 1.0 : (x < 0 ? -1.0 : x)) * sh); }
    case 2: {
        if (x <= 2 * t) {
            const double sh = fabs(x) - t;
            if (sh < 0) return (T)0;
            if (sh - t < 0) return (T)((x > 0 ? 1.0 : (x < 0 ? -1.0 : x)) * sh * 2);
        }
        return xv;
    }
    default: { const double sh = 1.0 - t * t / (x * x); return sh < 0 ? (T)0 : (T)(x * sh); }
    }
}

constexpr int kTT = 256, kTE = 8;              // threads, elements per thread of the thresholding pass
template <typename T>
__global__ void __launch_bounds__(kTT) threshold_k(T *__restrict__ y, const T *__restrict__ x, long slab, Div32 dn, const unsigned char *__restrict__ colmask,
                                                  long keep_lo, long keep_hi, int th, const double *__restrict__ sigma, double tmul, Div32 dchunks)
{
    const unsigned k = div32(blockIdx.x, dchunks), chunk = blockIdx.x - k * dchunks.d;
    const double t = sigma ? sigma[k] * tmul : tmul;
    const long base = (long)chunk * (kTT * kTE);
    const T *xs = x + (long)k * slab;
    T *ys = y + (long)k * slab;
    T v[kTE];
#pragma unroll
    for (int u = 0; u < kTE; ++u) { const long e = base + u * kTT + threadIdx.x; if (e < slab) v[u] = xs[e]; }
#pragma unroll
    for (int u = 0; u < kTE; ++u) {
        const long e = base + u * kTT + threadIdx.x;
        if (e < slab) {
            bool on = e < keep_lo || e >= keep_hi;
            if (colmask) on = on && colmask[div32((unsigned)e, dn)];
            ys[e] = on ? apply_th<T>(v[u], t, th) : v[u];
        }
    }
}

struct DevBuf {                        // stream-ordered scratch released on scope exit
    void *p = nullptr; cudaStream_t s;
    explicit DevBuf(cudaStream_t st) : s(st) {}
    ~DevBuf() { if (p) cudaFreeAsync(p, s); }
    int alloc(size_t bytes)
    {
        unsigned char *d; int rc = wx_scratch(&d, bytes, s); if (rc) return rc;
        p = d;
        return WX_OK;
    }
    int upload(const void *h, size_t bytes)
    {
        int rc = alloc(bytes); if (rc) return rc;
        WX_CUDA(cudaMemcpyAsync(p, h, bytes, cudaMemcpyHostToDevice, s));
        return WX_OK;
    }
};

static inline int next_pow2(long v) { int p = 1; while (p < v) p <<= 1; return p; }
static inline int sort_threads(int P) { int t = P / 2; if (t > kTS) t = kTS; if (t < 32) t = 32; return t; }

template <typename T>
int noisest_impl(double *sigma, const T *x, long stride, long off, long len, long N, cudaStream_t s)
{
    WX_REQUIRE(sigma && x && stride >= 1 && off >= 0 && len >= 1 && off + len <= stride && N >= 0, "bad arguments");
    WX_REQUIRE(len <= (1L << 28) && N < (1L << 31), "range too long");
    if (N == 0) return WX_OK;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const int P = next_pow2(len);
    const size_t bytes = (size_t)P * sizeof(T);
    DevBuf gb(s);
    auto kern = mad_k<T>;
    if (bytes <= dv.smem_optin) WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    else { rc = gb.alloc(bytes * N); if (rc) return rc; }
    kern<<<(unsigned)N, sort_threads(P), gb.p ? 0 : bytes, s>>>(sigma, x, stride, off, (int)len, P, (T *)gb.p);
    WX_LAUNCHED();
    return WX_OK;
}

// sorted magnitudes of the selected columns into srt (N, P)
template <typename T>
int sort_selected(DevBuf &srt, int *Mo, int *Po, const T *x, long n, long K, const unsigned char *colmask, long N, cudaStream_t s)
{
    std::vector<int> sel;
    for (long c = 0; c < K; ++c) if (!colmask || colmask[c]) sel.push_back((int)c);
    WX_REQUIRE(!sel.empty(), "no column selected");
    const long M = (long)sel.size() * n;
    WX_REQUIRE(M <= (1L << 28) && n < (1L << 31) && N < (1L << 31), "too many coefficients per signal");
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const bool all = (long)sel.size() == K;
    DevBuf cols(s);
    if (!all) { rc = cols.upload(sel.data(), sel.size() * sizeof(int)); if (rc) return rc; }
    const int P = next_pow2(M);
    const size_t bytes = (size_t)P * sizeof(T);
    rc = srt.alloc(bytes * N); if (rc) return rc;
    const int in_smem = bytes <= dv.smem_optin;
    auto kern = sort_abs_k<T>;
    if (in_smem) WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    kern<<<(unsigned)N, sort_threads(P), in_smem ? bytes : 0, s>>>((T *)srt.p, x, n * K, (int)n, all ? nullptr : (const int *)cols.p, (int)M, P, in_smem);
    WX_LAUNCHED();
    *Mo = (int)M; *Po = P;
    return WX_OK;
}

template <typename T>
int sure_impl(double *t, const T *x, long n, long K, const unsigned char *colmask, long N, cudaStream_t s)
{
    WX_REQUIRE(t && x && n >= 1 && K >= 1 && N >= 0, "bad arguments");
    if (N == 0) return WX_OK;
    DevBuf srt(s); int M, P;
    int rc = sort_selected(srt, &M, &P, x, n, K, colmask, N, s); if (rc) return rc;
    sure_k<T><<<(unsigned)N, kTS, 0, s>>>(t, (const T *)srt.p, M, P);
    WX_LAUNCHED();
    return WX_OK;
}

template <typename T>
int relerr_impl(double *t, const T *x, long n, long K, const unsigned char *colmask, int elbows, long N, cudaStream_t s)
{
    WX_REQUIRE(t && x && n >= 1 && K >= 1 && N >= 0, "bad arguments");
    WX_REQUIRE(elbows >= 1, "AssertionError: elbows >= 1");
    if (N == 0) return WX_OK;
    DevBuf srt(s), Yw(s); int M, P;
    int rc = sort_selected(srt, &M, &P, x, n, K, colmask, N, s); if (rc) return rc;
    rc = Yw.alloc((size_t)(M + 1) * N * sizeof(double)); if (rc) return rc;
    relerr_k<T><<<(unsigned)N, kTS, 0, s>>>(t, (const T *)srt.p, (double *)Yw.p, M, P, elbows);
    WX_LAUNCHED();
    return WX_OK;
}

template <typename T>
int threshold_impl(T *y, const T *x, long n, long K, const unsigned char *colmask, long keep_lo, long keep_hi, int th, const double *sigma, double tmul,
                   long N, cudaStream_t s)
{
    WX_REQUIRE(y && x && n >= 1 && K >= 1 && N >= 0, "bad arguments");
    WX_REQUIRE(th >= 0 && th <= 3, "unknown threshold type %d", th);
    WX_REQUIRE(tmul >= 0, "AssertionError: t >= 0");
    WX_REQUIRE(keep_lo >= 0 && keep_hi <= n * K, "keep range outside the signal");
    if (N == 0) return WX_OK;
    const long slab = n * K;
    WX_REQUIRE(slab < (1L << 31) && n < (1L << 31), "signal slab too large");
    const long chunks = (slab + kTT * kTE - 1) / (kTT * kTE);
    WX_REQUIRE(chunks * N < (1L << 31), "too many coefficients for one launch");
    DevBuf mask(s);
    if (colmask) { int rc = mask.upload(colmask, (size_t)K); if (rc) return rc; }
    if (keep_hi < keep_lo) keep_hi = keep_lo;
    threshold_k<T><<<(unsigned)(chunks * N), kTT, 0, s>>>(y, x, slab, make_div32((unsigned)n), (const unsigned char *)mask.p, keep_lo, keep_hi, th, sigma, tmul,
                                                        make_div32((unsigned)chunks));
    WX_LAUNCHED();
    return WX_OK;
}

}  // namespace

extern "C" {
int wx_noisest_f64(double *sigma, const double *x, long stride, long off, long len, long N, void *s) { return noisest_impl<double>(sigma, x, stride, off, len, N, (cudaStream_t)s); }
int wx_noisest_f32(double *sigma, const float *x, long stride, long off, long len, long N, void *s) { return noisest_impl<float>(sigma, x, stride, off, len, N, (cudaStream_t)s); }
int wx_surethreshold_f64(double *t, const double *x, long n, long K, const unsigned char *colmask, long N, void *s) { return sure_impl<double>(t, x, n, K, colmask, N, (cudaStream_t)s); }
int wx_surethreshold_f32(double *t, const float *x, long n, long K, const unsigned char *colmask, long N, void *s) { return sure_impl<float>(t, x, n, K, colmask, N, (cudaStream_t)s); }
int wx_relerrorthreshold_f64(double *t, const double *x, long n, long K, const unsigned char *colmask, int elbows, long N, void *s) { return relerr_impl<double>(t, x, n, K, colmask, elbows, N, (cudaStream_t)s); }
int wx_relerrorthreshold_f32(double *t, const float *x, long n, long K, const unsigned char *colmask, int elbows, long N, void *s) { return relerr_impl<float>(t, x, n, K, colmask, elbows, N, (cudaStream_t)s); }
int wx_threshold_f64(double *y, const double *x, long n, long K, const unsigned char *colmask, long keep_lo, long keep_hi, int th, const double *sigma, double tmul, long N, void *s)
{ return threshold_impl<double>(y, x, n, K, colmask, keep_lo, keep_hi, th, sigma, tmul, N, (cudaStream_t)s); }
int wx_threshold_f32(float *y, const float *x, long n, long K, const unsigned char *colmask, long keep_lo, long keep_hi, int th, const double *sigma, double tmul, long N, void *s)
{ return threshold_impl<float>(y, x, n, K, colmask, keep_lo, keep_hi, th, sigma, tmul, N, (cudaStream_t)s); }
}

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
#include <cstdlib>

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

// ---- warp-resident sort: one signal per warp, E elements per lane in registers (P = 32 E <= 1024) ----------------
// Element g of the sorted sequence lives in lane g / E, register g % E.  Bitonic network in its "flip" form: stage k first
// compares g with g ^ (k-1), then g with g ^ j for j = k/4 .. 1; every compare-exchange is ascending, so there are no
// direction flags.  Partners at distance < E are registers of the same lane (compile-time indices), the others come
// through shfl_xor.
template <typename T> __device__ __forceinline__ T shfl_xor_t(T v, int m) { return __shfl_xor_sync(0xffffffffu, v, m); }
template <typename T> __device__ __forceinline__ T shfl_idx_t(T v, int l) { return __shfl_sync(0xffffffffu, v, l); }

template <typename T> __device__ __forceinline__ void cex(T &a, T &b) { const bool sw = a > b; const T lo = sw ? b : a, hi = sw ? a : b; a = lo; b = hi; }

template <typename T, int E>
__device__ __forceinline__ void warp_bitonic(T (&v)[E])
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int k = 2; k <= 32 * E; k <<= 1) {
        if (k <= E) {
#pragma unroll
            for (int e = 0; e < E; ++e) { const int p = e ^ (k - 1); if (p > e) cex(v[e], v[p]); }
        } else {
            const int lm = k / E - 1;
            const bool keepmin = (lane & ((k / E) >> 1)) == 0;
#pragma unroll
            for (int e = 0; e < (E + 1) / 2; ++e) {
                const int f = E - 1 - e;
                const T o1 = shfl_xor_t(v[f], lm), o2 = shfl_xor_t(v[e], lm);
                const T a = v[e], b = v[f];
                v[e] = ((o1 < a) == keepmin) ? o1 : a;           // equal values: either copy will do
                if (f != e) v[f] = ((o2 < b) == keepmin) ? o2 : b;
            }
        }
#pragma unroll
        for (int j = k >> 2; j > 0; j >>= 1) {
            if (j < E) {
#pragma unroll
                for (int e = 0; e < E; ++e) if ((e & j) == 0) cex(v[e], v[e | j]);
            } else {
                const int lm = j / E;
                const bool keepmin = (lane & lm) == 0;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const T o = shfl_xor_t(v[e], lm), a = v[e];
                    v[e] = ((o < a) == keepmin) ? o : a;
                }
            }
        }
    }
}

// element g of the warp-sorted sequence, broadcast to all lanes
template <typename T, int E>
__device__ __forceinline__ T warp_elem(const T (&v)[E], int g)
{
    const int ge = g % E;
    T x = v[0];
#pragma unroll
    for (int e = 1; e < E; ++e) if (e == ge) x = v[e];
    return shfl_idx_t(x, g / E);
}
template <typename T, int E>
__device__ __forceinline__ T warp_median(const T (&v)[E], int M)
{
    return (M & 1) ? warp_elem<T, E>(v, M >> 1) : warp_elem<T, E>(v, (M >> 1) - 1) / (T)2 + warp_elem<T, E>(v, M >> 1) / (T)2;
}

constexpr int kWS = 4;                          // signals (warps) per CTA of the warp-resident kernels
template <typename T, int E>
__global__ void __launch_bounds__(32 * kWS) mad_warp_k(double *__restrict__ sigma, const T *__restrict__ x, long stride, long off, int len, long N)
{
    const long k = (long)blockIdx.x * kWS + (threadIdx.x >> 5);
    if (k >= N) return;
    const int lane = threadIdx.x & 31;
    const T *src = x + k * stride + off;
    T v[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { const int i = e * 32 + lane; v[e] = i < len ? src[i] : wx_inf<T>(); }     // any order will do
    T md = 0;
#pragma unroll 1
    for (int pass = 0; pass < 2; ++pass) {          // one copy of the (fully unrolled) network in the instruction stream
        warp_bitonic<T, E>(v);
        md = warp_median<T, E>(v, len);
        if (pass == 0) {
#pragma unroll
            for (int e = 0; e < E; ++e) v[e] = (lane * E + e) < len ? (T)fabs(v[e] - md) : wx_inf<T>();
        }
    }
    if (lane == 0) sigma[k] = (double)md / 0.6745;
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

// ---- warp-resident surethreshold / relerrorthreshold (M <= 1024 selected coefficients per signal) ----------------------
template <typename T, int E>
__device__ __forceinline__ void warp_load_abs_sorted(T (&v)[E], const T *__restrict__ src, int n, const int *__restrict__ cols, int M)
{
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int i = e * 32 + lane;
        T a = wx_inf<T>();
        if (i < M) { const int c = i / n, r = i - c * n; a = fabs(src[(long)(cols ? cols[c] : c) * n + r]); }
        v[e] = a;
    }
    warp_bitonic<T, E>(v);
}
template <bool MAX>
__device__ __forceinline__ void warp_argext(double &v, int &idx)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, v, o);
        const int oi = __shfl_xor_sync(0xffffffffu, idx, o);
        const bool better = oi >= 0 && (idx < 0 || (MAX ? (ov > v || (ov == v && oi < idx)) : (ov < v || (ov == v && oi < idx))));
        if (better) { v = ov; idx = oi; }
    }
}

template <typename T, int E>
__global__ void __launch_bounds__(32 * kWS) sure_warp_k(double *__restrict__ t, const T *__restrict__ x, long slab, int n, const int *__restrict__ cols, int M, long N)
{
    const long k = (long)blockIdx.x * kWS + (threadIdx.x >> 5);
    if (k >= N) return;
    const int lane = threadIdx.x & 31;
    T v[E];
    warp_load_abs_sorted<T, E>(v, x + k * slab, n, cols, M);
    double loc = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) { const double a = (double)v[e]; if (lane * E + e < M) loc += a * a; }
    double b = warp_incl_scan(loc) - loc;
    double best = 0; int bi = -1;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        const int g = lane * E + e;
        if (g < M) {
            const double a = (double)v[e], av = a * a;
            b += av;
            const double risk = ((double)(M - 2 * (g + 1)) + (b + (double)(M - 1 - g) * av)) / (double)M;
            if (bi < 0 || risk < best) { best = risk; bi = g; }
        }
    }
    warp_argext<false>(best, bi);
    const double a = (double)warp_elem<T, E>(v, bi);
    if (lane == 0) t[k] = sqrt(a * a);
}

template <typename T, int E>
__global__ void __launch_bounds__(32 * kWS) relerr_warp_k(double *__restrict__ t, const T *__restrict__ x, long slab, int n, const int *__restrict__ cols, int M, int elbows,
                                                         long N)
{
    const long k = (long)blockIdx.x * kWS + (threadIdx.x >> 5);
    if (k >= N) return;
    const int lane = threadIdx.x & 31;
    T v[E];
    warp_load_abs_sorted<T, E>(v, x + k * slab, n, cols, M);
    // sum of the squares above each element (descending running sum of orth2relerror), S = all of them
    double Y[E];
    double loc = 0;
#pragma unroll
    for (int e = E - 1; e >= 0; --e) { Y[e] = loc; const double a = (double)v[e]; if (lane * E + e < M) loc += a * a; }
    double inc = loc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double u = __shfl_down_sync(0xffffffffu, inc, o); if (lane + o < 32) inc += u; }
    const double S = __shfl_sync(0xffffffffu, inc, 0), above = inc - loc, rS = sqrt(S);
    const double xmax = (double)warp_elem<T, E>(v, M - 1);
#pragma unroll
    for (int e = 0; e < E; ++e) Y[e] = lane * E + e < M ? sqrt(fabs(S - (above + Y[e]))) / rS : 0.0;      // point g+1 (g = M-1 is patched next)
    const double r1 = M > 1 ? warp_elem<double, E>(Y, M - 2) : sqrt(fabs(S - S)) / rS;
    double ymax = 0;
#pragma unroll
    for (int e = 0; e < E; ++e) {
        if (lane * E + e == M - 1) Y[e] = r1;
        ymax = fmax(ymax, Y[e]);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) ymax = fmax(ymax, __shfl_xor_sync(0xffffffffu, ymax, o));
    const double x0 = 0.0 / xmax, y0 = (sqrt(fabs(S - S)) / rS) / ymax;          // point 0: (0, r_M)
    // the normalised curve, once: X_p = x / xmax (kept in Float64 in place of Y's neighbour), Y_p = r / ymax
    double X[E];
#pragma unroll
    for (int e = 0; e < E; ++e) { X[e] = (double)v[e] / xmax; Y[e] = Y[e] / ymax; }
    int end = M;
    for (int el = 0; el < elbows; ++el) {
        const double xe = end > 0 ? warp_elem<double, E>(X, end - 1) : x0;
        const double ye = end > 0 ? warp_elem<double, E>(Y, end - 1) : y0;
        double vx = xe - x0, vy = ye - y0;
        const double nv = sqrt(vx * vx + vy * vy);
        vx /= nv; vy /= nv;
        double best = 0; int bi = -1;
        if (lane == 0) { const double H = sqrt(0.0), A = 0.0 * vx + 0.0 * vy; best = sqrt(fabs(H * H - A * A)); bi = 0; }
#pragma unroll
        for (int e = 0; e < E; ++e) {
            const int p = lane * E + e + 1;
            if (p <= end) {
                const double dx = X[e] - x0, dy = Y[e] - y0;
                const double H = sqrt(dx * dx + dy * dy), A = dx * vx + dy * vy;
                const double O = sqrt(fabs(H * H - A * A));
                if (bi < 0 || O > best) { best = O; bi = p; }
            }
        }
        warp_argext<true>(best, bi);
        end = bi < 0 ? 0 : bi;
    }
    const double xn = end ? warp_elem<double, E>(X, end - 1) : 0.0 / xmax;
    if (lane == 0) t[k] = xn * xmax;
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
// the batch is one flat run of N slabs: a CTA takes kTT*kTE consecutive coefficients wherever the slab boundaries fall
template <typename T>
__global__ void __launch_bounds__(kTT) threshold_k(T *__restrict__ y, const T *__restrict__ x, Div32 dslab, Div32 dn, const unsigned char *__restrict__ colmask,
                                                  unsigned keep_lo, unsigned keep_hi, int th, const double *__restrict__ sigma, double tmul, long total)
{
    const long base = (long)blockIdx.x * (kTT * kTE);
    const long k0 = base / dslab.d;
    const unsigned e0 = (unsigned)(base - k0 * dslab.d);
    T v[kTE];
#pragma unroll
    for (int u = 0; u < kTE; ++u) { const long i = base + u * kTT + threadIdx.x; if (i < total) v[u] = x[i]; }
#pragma unroll
    for (int u = 0; u < kTE; ++u) {
        const unsigned idx = u * kTT + threadIdx.x;
        const long i = base + idx;
        if (i < total) {
            const unsigned kk = div32(e0 + idx, dslab), e = e0 + idx - kk * dslab.d;
            const double t = sigma ? sigma[k0 + kk] * tmul : tmul;
            bool on = e < keep_lo || e >= keep_hi;
            if (colmask) on = on && colmask[div32(e, dn)];
            y[i] = on ? apply_th<T>(v[u], t, th) : v[u];
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
    static const bool nowarp = getenv("WX_B200_NO_WARP_SORT") != nullptr;
    if (P <= 1024 && !nowarp) {
        const unsigned grid = (unsigned)((N + kWS - 1) / kWS);
#define WX_MW(EE) case EE: mad_warp_k<T, EE><<<grid, 32 * kWS, 0, s>>>(sigma, x, stride, off, (int)len, N); break;
        switch (P <= 32 ? 1 : P / 32) { WX_MW(1) WX_MW(2) WX_MW(4) WX_MW(8) WX_MW(16) WX_MW(32) }
#undef WX_MW
        WX_LAUNCHED();
        return WX_OK;
    }
    const size_t bytes = (size_t)P * sizeof(T);
    DevBuf gb(s);
    auto kern = mad_k<T>;
    if (bytes <= dv.smem_optin) WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    else { rc = gb.alloc(bytes * N); if (rc) return rc; }
    kern<<<(unsigned)N, sort_threads(P), gb.p ? 0 : bytes, s>>>(sigma, x, stride, off, (int)len, P, (T *)gb.p);
    WX_LAUNCHED();
    return WX_OK;
}

// the selected columns: cols.p = device list (null = all K columns), *Mo = coefficients per signal
static int select_columns(DevBuf &cols, long *Mo, long n, long K, const unsigned char *colmask, long N)
{
    std::vector<int> sel;
    for (long c = 0; c < K; ++c) if (!colmask || colmask[c]) sel.push_back((int)c);
    WX_REQUIRE(!sel.empty(), "no column selected");
    *Mo = (long)sel.size() * n;
    WX_REQUIRE(*Mo <= (1L << 28) && n < (1L << 31) && N < (1L << 31), "too many coefficients per signal");
    if ((long)sel.size() != K) return cols.upload(sel.data(), sel.size() * sizeof(int));
    return WX_OK;
}
static inline bool warp_sortable(long M)
{
    static const bool nowarp = getenv("WX_B200_NO_WARP_SORT") != nullptr;
    return M <= 1024 && !nowarp;
}

// sorted magnitudes of the selected columns into srt (N, P)
template <typename T>
int sort_selected(DevBuf &srt, int *Mo, int *Po, const T *x, long n, long K, long M, const DevBuf &cols, long N, cudaStream_t s)
{
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const int P = next_pow2(M);
    const size_t bytes = (size_t)P * sizeof(T);
    rc = srt.alloc(bytes * N); if (rc) return rc;
    const int in_smem = bytes <= dv.smem_optin;
    auto kern = sort_abs_k<T>;
    if (in_smem) WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    kern<<<(unsigned)N, sort_threads(P), in_smem ? bytes : 0, s>>>((T *)srt.p, x, n * K, (int)n, (const int *)cols.p, (int)M, P, in_smem);
    WX_LAUNCHED();
    *Mo = (int)M; *Po = P;
    return WX_OK;
}

template <typename T>
int sure_impl(double *t, const T *x, long n, long K, const unsigned char *colmask, long N, cudaStream_t s)
{
    WX_REQUIRE(t && x && n >= 1 && K >= 1 && N >= 0, "bad arguments");
    if (N == 0) return WX_OK;
    DevBuf srt(s), cols(s); int M, P; long Ml;
    int rc = select_columns(cols, &Ml, n, K, colmask, N); if (rc) return rc;
    if (warp_sortable(Ml)) {
        const unsigned grid = (unsigned)((N + kWS - 1) / kWS);
#define WX_SW(EE) case EE: sure_warp_k<T, EE><<<grid, 32 * kWS, 0, s>>>(t, x, n * K, (int)n, (const int *)cols.p, (int)Ml, N); break;
        switch (Ml <= 32 ? 1 : next_pow2(Ml) / 32) { WX_SW(1) WX_SW(2) WX_SW(4) WX_SW(8) WX_SW(16) WX_SW(32) }
#undef WX_SW
        WX_LAUNCHED();
        return WX_OK;
    }
    rc = sort_selected(srt, &M, &P, x, n, K, Ml, cols, N, s); if (rc) return rc;
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
    DevBuf srt(s), Yw(s), cols(s); int M, P; long Ml;
    int rc = select_columns(cols, &Ml, n, K, colmask, N); if (rc) return rc;
    if (warp_sortable(Ml)) {
        const unsigned grid = (unsigned)((N + kWS - 1) / kWS);
#define WX_RW(EE) case EE: relerr_warp_k<T, EE><<<grid, 32 * kWS, 0, s>>>(t, x, n * K, (int)n, (const int *)cols.p, (int)Ml, elbows, N); break;
        switch (Ml <= 32 ? 1 : next_pow2(Ml) / 32) { WX_RW(1) WX_RW(2) WX_RW(4) WX_RW(8) WX_RW(16) WX_RW(32) }
#undef WX_RW
        WX_LAUNCHED();
        return WX_OK;
    }
    rc = sort_selected(srt, &M, &P, x, n, K, Ml, cols, N, s); if (rc) return rc;
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
    WX_REQUIRE(slab < (1L << 31) - kTT * kTE && n < (1L << 31), "signal slab too large");
    const long total = slab * N, ctas = (total + kTT * kTE - 1) / (kTT * kTE);
    WX_REQUIRE(ctas < (1L << 31), "too many coefficients for one launch");
    DevBuf mask(s);
    if (colmask) { int rc = mask.upload(colmask, (size_t)K); if (rc) return rc; }
    if (keep_hi < keep_lo) keep_hi = keep_lo;
    threshold_k<T><<<(unsigned)ctas, kTT, 0, s>>>(y, x, make_div32(slab), make_div32(n), (const unsigned char *)mask.p, (unsigned)keep_lo, (unsigned)keep_hi, th, sigma,
                                                tmul, total);
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

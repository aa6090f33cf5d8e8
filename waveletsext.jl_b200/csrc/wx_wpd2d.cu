// wx_wpd2d.cu -- batched 2-D wavelet packet decomposition with the separable step fused in shared memory.
//
// Reference: wpdall on images dwt/dwt_all.jl:260-282 -> wpd! 2-D DWT.jl:164-209 -> dwt_step! 2-D
// dwt/dwt_one_level.jl:319-354 (columns into temp, rows into the four quadrants w1 = LL top-left, w2 top-right,
// w3 bottom-left, w4 bottom-right).   x(m,n,N) -> y(m,n,L+1,N), level d+1 node (jr,jc) = quadrants of level d.
//
// Two kernels, both HBM-bound (2F FMAs per pixel per level):
//  * wpd2d_tile_k  (nodes larger than shared memory): a CTA produces a tr x tc tile of each of the four children of
//    one node from the (2tr+F-2) x (2tc+F-2) parent patch (periodic halo); column pass and row pass both run in shared
//    memory, so a level costs one read of the parent (+halo) and one write of the children instead of two round trips
//    through a temp image.  At depth 0 the patch core is also written to level 0 (y[:,:,1] = x, DWT.jl:176).
//  * wpd2d_block_k (nodes that fit): a CTA keeps a whole node in shared memory and runs every remaining level there
//    (node -> 4 children -> 16 grandchildren ...), writing each level slice once; the parent is never re-read.
// The detail outputs are taken S = (F-2)/2 positions ahead of the scaling outputs so that both filters read the same
// window v[2i .. 2i+F-1]; accumulation order per output is the reference's tap order.
#include "wx_steps.cuh"
#include <cstdlib>

namespace {

constexpr int kT2 = 256;

template <typename T> struct Pair;
template <> struct Pair<double> { using type = double2; };
template <> struct Pair<float> { using type = float2; };

// w[0..F-1] = v[2i .. 2i+F-1]  ->  lo = w1[i],  hi = w2[i + (F-2)/2]        dwt/dwt_one_level.jl:97-104
template <typename T, int F>
__device__ __forceinline__ void dwt_dots(const T *w, const Taps<T> &tp, T &lo, T &hi)
{
    T a = tp.g[F - 1] * w[0];
    T b = tp.h[0] * w[F - 1];
#pragma unroll
    for (int j = 1; j < F; ++j) {
        a = fma(tp.g[F - 1 - j], w[j], a);
        b = fma(tp.h[j], w[F - 1 - j], b);
    }
    lo = a;
    hi = b;
}

struct FastDiv {                    // division by a runtime constant that is usually a power of two
    int d, lg;
    __host__ __device__ FastDiv() : d(1), lg(0) {}
    __host__ __device__ explicit FastDiv(int dd) : d(dd), lg(-1)
    {
        if (dd > 0 && (dd & (dd - 1)) == 0) { lg = 0; while ((1 << lg) < dd) ++lg; }
    }
    __device__ __forceinline__ int div(int x) const { return lg >= 0 ? (x >> lg) : (x / d); }
    __device__ __forceinline__ int mod(int x) const { return lg >= 0 ? (x & (d - 1)) : (x % d); }
};

// ---------------------------------------------------------------------------------------------------------
// one level, tiles with halo
// ---------------------------------------------------------------------------------------------------------
template <typename T, int F>
__global__ void __launch_bounds__(kT2) wpd2d_tile_k(T *__restrict__ y, const T *__restrict__ x, long m, long n, int L, int d, int tr, int tc,
                                                   Taps<T> tp)
{
    using P2 = typename Pair<T>::type;
    constexpr int S = (F - 2) / 2;
    extern __shared__ __align__(16) unsigned char wx_2d_smem[];
    const int PR = 2 * tr + F - 2, PC = 2 * tc + F - 2;
    T *P = reinterpret_cast<T *>(wx_2d_smem);           // (PR, PC) parent patch, column-major
    T *Tm = P + (size_t)PR * PC;                         // (2tr, PC) after the column pass: rows [0,tr) scaling, [tr,2tr) detail
    const int tid = threadIdx.x;
    const int mp = (int)(m >> d), np = (int)(n >> d), hr = mp / 2, hc = np / 2;
    const int tiles_r = hr / tr, tiles_c = hc / tc, nodes = 1 << d;
    // block -> (tile row, node row, tile col, node col, image); neighbouring CTAs walk down the rows of the image
    long bid = blockIdx.x;
    const int ti = (int)(bid % tiles_r); bid /= tiles_r;
    const int jr = (int)(bid % nodes); bid /= nodes;
    const int tk = (int)(bid % tiles_c); bid /= tiles_c;
    const int jc = (int)(bid % nodes); bid /= nodes;
    const long k = bid;
    const long img = m * n;
    T *yk = y + k * img * (L + 1);
    const bool from_x = (d == 0 && x != nullptr);
    const T *par = from_x ? (x + k * img) : (yk + (long)d * img);
    const int nr0 = jr * mp, nc0 = jc * np;              // node origin in the image
    const int i0 = ti * tr, k0 = tk * tc;                // tile origin in child coordinates

    // ---- parent patch (periodic inside the node), two rows per thread ----
    const int PR2 = PR / 2;
    for (int idx = tid; idx < PR2 * PC; idx += kT2) {
        const int b = idx / PR2, a = 2 * (idx - b * PR2);
        int rr = 2 * i0 + a; rr %= mp;
        int cc = 2 * k0 + b; cc %= np;
        const P2 v = *reinterpret_cast<const P2 *>(par + (long)(nc0 + cc) * m + nr0 + rr);
        *reinterpret_cast<P2 *>(P + (size_t)b * PR + a) = v;
        if (from_x && a < 2 * tr && b < 2 * tc)          // y[:,:,1] = x   DWT.jl:176
            *reinterpret_cast<P2 *>(yk + (long)(nc0 + 2 * k0 + b) * m + nr0 + 2 * i0 + a) = v;
    }
    __syncthreads();
    // ---- column pass: every column of the patch, tr output pairs ----
    for (int idx = tid; idx < tr * PC; idx += kT2) {
        const int b = idx / tr, il = idx - b * tr;
        T w[F];
        const T *src = P + (size_t)b * PR + 2 * il;
#pragma unroll
        for (int q = 0; q < F / 2; ++q) {
            const P2 v = *reinterpret_cast<const P2 *>(src + 2 * q);
            w[2 * q] = v.x; w[2 * q + 1] = v.y;
        }
        T lo, hi;
        dwt_dots<T, F>(w, tp, lo, hi);
        Tm[(size_t)b * (2 * tr) + il] = lo;
        Tm[(size_t)b * (2 * tr) + tr + il] = hi;
    }
    __syncthreads();
    // ---- row pass + store: (2tr rows) x (tc output pairs) ----
    T *ynext = yk + (long)(d + 1) * img;
    const int R2 = 2 * tr;
    for (int idx = tid; idx < R2 * tc; idx += kT2) {
        const int kl = idx / R2, r = idx - kl * R2;
        T w[F];
#pragma unroll
        for (int j = 0; j < F; ++j) w[j] = Tm[(size_t)(2 * kl + j) * R2 + r];
        T lo, hi;
        dwt_dots<T, F>(w, tp, lo, hi);
        int ci, qr;
        if (r < tr) { ci = i0 + r; qr = 0; }
        else { ci = (i0 + (r - tr) + S) % hr; qr = hr; }
        const int ck_lo = k0 + kl, ck_hi = (k0 + kl + S) % hc;
        T *o = ynext + nr0 + qr + ci;
        o[(long)(nc0 + ck_lo) * m] = lo;                 // w1 / w3
        o[(long)(nc0 + hc + ck_hi) * m] = hi;            // w2 / w4
    }
}

// ---------------------------------------------------------------------------------------------------------
// all remaining levels of a node that fits shared memory
// ---------------------------------------------------------------------------------------------------------
template <typename T, int F>
__global__ void __launch_bounds__(kT2) wpd2d_block_k(T *__restrict__ y, const T *__restrict__ x, long m, long n, int L, int db, int dend,
                                                    Taps<T> tp)
{
    using P2 = typename Pair<T>::type;
    constexpr int S = (F - 2) / 2;
    extern __shared__ __align__(16) unsigned char wx_2d_smem[];
    const int BR = (int)(m >> db), BC = (int)(n >> db);
    T *A = reinterpret_cast<T *>(wx_2d_smem);            // (BR, BC) column-major: the block at the current level
    T *Tm = A + (size_t)BR * BC;                          // after the column pass
    const int tid = threadIdx.x;
    const int nodes = 1 << db;
    long bid = blockIdx.x;
    const int jr = (int)(bid % nodes); bid /= nodes;
    const int jc = (int)(bid % nodes); bid /= nodes;
    const long k = bid;
    const long img = m * n;
    T *yk = y + k * img * (L + 1);
    const bool from_x = (db == 0 && x != nullptr);
    const T *par = from_x ? (x + k * img) : (yk + (long)db * img);
    const long org = (long)(jc * BC) * m + jr * BR;       // block origin in the image
    const int BR2 = BR / 2;
    const FastDiv dBR2(BR2), dBR(BR);

    for (int idx = tid; idx < BR2 * BC; idx += kT2) {
        const int b = dBR2.div(idx), a = 2 * (idx - b * BR2);
        const P2 v = *reinterpret_cast<const P2 *>(par + org + (long)b * m + a);
        *reinterpret_cast<P2 *>(A + (size_t)b * BR + a) = v;
        if (from_x) *reinterpret_cast<P2 *>(yk + org + (long)b * m + a) = v;
    }
    __syncthreads();
    for (int l = db; l < dend; ++l) {
        const int mpl = (int)(m >> l), npl = (int)(n >> l), hr = mpl / 2, hc = npl / 2;
        const FastDiv dhr(hr), dhc(hc);
        // column pass A -> Tm, per node: scaling rows on top, detail rows below (shift resolved here)
        for (int idx = tid; idx < BR2 * BC; idx += kT2) {
            const int b = dBR2.div(idx), ig = idx - b * BR2;
            const int jn = dhr.div(ig), il = ig - jn * hr, r0 = jn * mpl;
            const T *src = A + (size_t)b * BR + r0;
            T w[F];
            if (mpl >= F) {                                // at most one wrap, pairs stay together
#pragma unroll
                for (int q = 0; q < F / 2; ++q) {
                    int rr = 2 * il + 2 * q; if (rr >= mpl) rr -= mpl;
                    const P2 v = *reinterpret_cast<const P2 *>(src + rr);
                    w[2 * q] = v.x; w[2 * q + 1] = v.y;
                }
            } else {
#pragma unroll
                for (int j = 0; j < F; ++j) w[j] = src[(2 * il + j) % mpl];
            }
            T lo, hi;
            dwt_dots<T, F>(w, tp, lo, hi);
            T *dst = Tm + (size_t)b * BR + r0;
            dst[il] = lo;
            int ih = il + S; if (ih >= hr) ih = dhr.mod(ih);
            dst[hr + ih] = hi;
        }
        __syncthreads();
        // row pass Tm -> A, per node: scaling columns left, detail columns right
        for (int idx = tid; idx < BR * (BC / 2); idx += kT2) {
            const int kg = dBR.div(idx), r = idx - kg * BR;
            const int jn = dhc.div(kg), kl = kg - jn * hc, c0 = jn * npl;
            const T *src = Tm + (size_t)c0 * BR + r;
            T w[F];
            if (npl >= F) {
#pragma unroll
                for (int j = 0; j < F; ++j) { int cc = 2 * kl + j; if (cc >= npl) cc -= npl; w[j] = src[(size_t)cc * BR]; }
            } else {
#pragma unroll
                for (int j = 0; j < F; ++j) w[j] = src[(size_t)((2 * kl + j) % npl) * BR];
            }
            T lo, hi;
            dwt_dots<T, F>(w, tp, lo, hi);
            T *dst = A + (size_t)c0 * BR + r;
            dst[(size_t)kl * BR] = lo;
            int kh = kl + S; if (kh >= hc) kh = dhc.mod(kh);
            dst[(size_t)(hc + kh) * BR] = hi;
        }
        __syncthreads();
        // level l+1 slice
        T *ynext = yk + (long)(l + 1) * img + org;
        for (int idx = tid; idx < BR2 * BC; idx += kT2) {
            const int b = dBR2.div(idx), a = 2 * (idx - b * BR2);
            *reinterpret_cast<P2 *>(ynext + (long)b * m + a) = *reinterpret_cast<const P2 *>(A + (size_t)b * BR + a);
        }
        // the next column pass only reads A (complete after the barrier above) and writes Tm (free): no barrier needed here
    }
}

static int largest_divisor_le(long v, int cap)
{
    int best = 1;
    for (int t = 1; t <= cap && t <= v; ++t) if (v % t == 0) best = t;
    return best;
}

template <typename T, int F>
int wpd2d_run(T *y, const T *x, long m, long n, int L, long N, const Taps<T> &t, cudaStream_t s)
{
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    // depth from which a whole node fits the block kernel's two buffers (64 KB keeps three CTAs per SM)
    const size_t budget = 65536;
    int db = 0;
    while (db < L && (size_t)2 * (m >> db) * (n >> db) * sizeof(T) > budget) ++db;
    static const char *env = getenv("WX_B200_WPD2D_TILE");      // measurement knob: tile edge of the halo kernel
    const int cap = env ? atoi(env) : 32;
    for (int d = 0; d < db; ++d) {
        const long hr = (m >> d) / 2, hc = (n >> d) / 2;
        const int tr = largest_divisor_le(hr, cap), tc = largest_divisor_le(hc, cap);
        const int PR = 2 * tr + F - 2, PC = 2 * tc + F - 2;
        const size_t smem = ((size_t)PR * PC + (size_t)2 * tr * PC) * sizeof(T);
        if (smem > dv.smem_optin) return wx_fail(WX_EUNSUPPORTED, "wpd 2-D tile does not fit shared memory");
        const long blocks = (hr / tr) * (hc / tc) * (1L << (2 * d)) * N;
        if (blocks >= (1L << 31)) return wx_fail(WX_EUNSUPPORTED, "wpd 2-D: too many tiles for one launch");
        auto kern = wpd2d_tile_k<T, F>;
        WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, kT2, smem, s>>>(y, d == 0 ? x : nullptr, m, n, L, d, tr, tc, t);
        WX_LAUNCHED();
    }
    if (db < L) {
        const size_t smem = (size_t)2 * (m >> db) * (n >> db) * sizeof(T);
        const long blocks = (1L << (2 * db)) * N;
        if (blocks >= (1L << 31)) return wx_fail(WX_EUNSUPPORTED, "wpd 2-D: too many blocks for one launch");
        auto kern = wpd2d_block_k<T, F>;
        WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<(unsigned)blocks, kT2, smem, s>>>(y, db == 0 ? x : nullptr, m, n, L, db, L, t);
        WX_LAUNCHED();
    }
    return WX_OK;
}

}  // namespace

// x(m,n,N) -> y(m,n,L+1,N) including the level-0 copy.  *handled = false: shape not covered (caller runs the per-level path).
template <typename T>
int wx_wpd2d_fused(T *y, const T *x, long m, long n, int L, long N, const Taps<T> &t, cudaStream_t s, bool *handled)
{
    *handled = false;
    static const bool off = getenv("WX_B200_NO_FUSED_WPD2D") != nullptr;   // debugging / A-B measurements only
    if (off || L < 1 || N < 1 || m >= (1L << 20) || n >= (1L << 20)) return WX_OK;
    if (((((uintptr_t)y) | ((uintptr_t)x)) & 15) != 0) return WX_OK;
    if ((m >> (L - 1)) % 2 != 0 || (n >> (L - 1)) % 2 != 0) return WX_OK;
    int rc;
#define WX_2D_CASE(FF) case FF: rc = wpd2d_run<T, FF>(y, x, m, n, L, N, t, s); break;
    switch (t.F) {
        WX_2D_CASE(2) WX_2D_CASE(4) WX_2D_CASE(6) WX_2D_CASE(8) WX_2D_CASE(10) WX_2D_CASE(12) WX_2D_CASE(16) WX_2D_CASE(20)
        default: return WX_OK;
    }
#undef WX_2D_CASE
    if (rc == WX_EUNSUPPORTED) return WX_OK;          // the per-level path recomputes everything
    if (rc == WX_OK) *handled = true;
    return rc;
}
template int wx_wpd2d_fused<double>(double *, const double *, long, long, int, long, const Taps<double> &, cudaStream_t, bool *);
template int wx_wpd2d_fused<float>(float *, const float *, long, long, int, long, const Taps<float> &, cudaStream_t, bool *);

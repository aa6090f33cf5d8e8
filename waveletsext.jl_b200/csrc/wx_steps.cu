// wx_steps.cu -- generic one-level step kernels (strided, batched, any n / F) and the single-step C ABI.
// One thread per output (pair); taps are read from kernel-parameter constant memory with a runtime
// index.  These kernels favour generality; the bandwidth-critical batched trees have fused kernels of
// their own (wx_wpd1d.cu, ...).
#include "wx_steps.cuh"
#include <cstdlib>

namespace {

constexpr int kThreads = 256;

// Division of a 32-bit index by a launch constant without the integer divide sequence (round-up multiply-high, exact for
// every 32-bit dividend); powers of two are shifts.
struct Div32 {
    unsigned d, m, s;
    bool p2;
};
static inline Div32 make_div32(long dd)
{
    Div32 r;
    r.d = (unsigned)dd; r.p2 = (dd & (dd - 1)) == 0; r.m = 0; r.s = 0;
    if (r.p2) { while ((1L << r.s) < dd) ++r.s; return r; }
    unsigned l = 0;
    while ((1UL << l) < (unsigned long)dd) ++l;                       // l = ceil(log2 d)
    r.m = (unsigned)((((1UL << l) - (unsigned long)dd) << 32) / (unsigned long)dd + 1);
    r.s = l - 1;
    return r;
}
__device__ __forceinline__ unsigned div32(unsigned x, const Div32 &v)
{
    if (v.p2) return x >> v.s;
    const unsigned t = __umulhi(v.m, x);
    return (t + ((x - t) >> 1)) >> v.s;
}

// flattened launch index -> (element i, batch coordinates); 32-bit fast path when the launch has < 2^32 threads
struct Geo {
    Batch b;
    long nout;
    bool small;
    Div32 dn, d0, d1;
};
static inline Geo make_geo(long nout, const Batch &b)
{
    Geo g; g.b = b; g.nout = nout;
    const long tot = nout * b.B0 * b.B1 * b.B2;
    g.small = tot < (1L << 32) && nout < (1L << 31) && b.B0 < (1L << 31) && b.B1 < (1L << 31);
    g.dn = make_div32(nout > 0 ? nout : 1); g.d0 = make_div32(b.B0 > 0 ? b.B0 : 1); g.d1 = make_div32(b.B1 > 0 ? b.B1 : 1);
    return g;
}

template <typename T>
__device__ __forceinline__ long voff(const View<T> &v, long b0, long b1, long b2)
{
    return b0 * v.s0 + b1 * v.s1 + b2 * v.s2;
}

__device__ __forceinline__ bool decomp(const Geo &g, long &i, long &b0, long &b1, long &b2)
{
    if (g.small) {
        const unsigned long tot = (unsigned long)g.nout * g.b.B0 * g.b.B1 * g.b.B2;
        const unsigned long gid = (unsigned long)blockIdx.x * kThreads + threadIdx.x;
        if (gid >= tot) return false;
        unsigned idx = (unsigned)gid, q;
        if (g.b.batch_fast) { q = div32(idx, g.d0); b0 = idx - q * g.d0.d; idx = q; q = div32(idx, g.dn); i = idx - q * g.dn.d; idx = q; }
        else                { q = div32(idx, g.dn); i = idx - q * g.dn.d; idx = q; q = div32(idx, g.d0); b0 = idx - q * g.d0.d; idx = q; }
        q = div32(idx, g.d1); b1 = idx - q * g.d1.d; b2 = q;
        return true;
    }
    long idx = (long)blockIdx.x * kThreads + threadIdx.x;
    const long tot = g.nout * g.b.B0 * g.b.B1 * g.b.B2;
    if (idx >= tot) return false;
    if (g.b.batch_fast) { b0 = idx % g.b.B0; idx /= g.b.B0; i = idx % g.nout; idx /= g.nout; }
    else                { i = idx % g.nout; idx /= g.nout; b0 = idx % g.b.B0; idx /= g.b.B0; }
    b1 = idx % g.b.B1; b2 = idx / g.b.B1;
    return true;
}

static inline unsigned grid_for(long total) { return (unsigned)((total + kThreads - 1) / kThreads); }

// FF > 0: filter length known at compile time (taps become immediate constant-bank operands, loops unroll); FF = 0: tp.F
#define WX_F (FF > 0 ? FF : tp.F)

// a1 dwt_step!  dwt/dwt_one_level.jl:94-105
template <typename T, int FF>
__global__ void __launch_bounds__(kThreads) dwt_step_k(View<T> w1, View<T> w2, View<const T> v, long n, Geo geo, Taps<T> tp, int sh)
{
    long i, b0, b1, b2;
    if (!decomp(geo, i, b0, b1, b2)) return;
    const T *pv = v.p + voff(v, b0, b1, b2);
    const int F = WX_F;
    long k1 = 2 * i - sh, k2 = 2 * i + 1 - sh;     // sh = 1: sidwt_step!(..., s = true) siwt/siwt_one_level.jl:88-89
    if (k1 < 0) k1 += n;
    if (k2 >= n) k2 -= n;
    T a1 = tp.g[F - 1] * pv[k1 * v.es];
    T a2 = tp.h[0] * pv[k2 * v.es];
#pragma unroll
    for (int j = 1; j < F; ++j) {
        k1 += 1; if (k1 >= n) k1 -= n;
        k2 -= 1; if (k2 < 0) k2 += n;
        a1 = fma(tp.g[F - 1 - j], pv[k1 * v.es], a1);
        a2 = fma(tp.h[j], pv[k2 * v.es], a2);
    }
    w1.p[voff(w1, b0, b1, b2) + i * w1.es] = a1;
    w2.p[voff(w2, b0, b1, b2) + i * w2.es] = a2;
}

// a2 idwt_step! dwt/dwt_one_level.jl:207-221
template <typename T, int FF>
__global__ void __launch_bounds__(kThreads) idwt_step_k(View<T> v, View<const T> w1, View<const T> w2, long n, Geo geo, Taps<T> tp, int sh)
{
    long i0b, b0, b1, b2;
    if (!decomp(geo, i0b, b0, b1, b2)) return;
    const T *p1 = w1.p + voff(w1, b0, b1, b2);
    const T *p2 = w2.p + voff(w2, b0, b1, b2);
    const int F = WX_F;
    const long n1 = n / 2;
    long i = i0b + 1;                              // Julia's 1-based i
    int j0 = (i & 1) ? 1 : 2;
    int j1 = F - j0 + 1;
    int j2 = ((i + 1) & 1) ? 1 : 2;
    long k1 = (i + 1) >> 1, k2 = k1;
    T acc = fma(tp.g[j1 - 1], p1[(k1 - 1) * w1.es], tp.h[j2 - 1] * p2[(k2 - 1) * w2.es]);
#pragma unroll
    for (int jj = 1; jj < (F + 1) / 2; ++jj) {
        const int j = j0 + 2 * jj;
        if (j > F) break;
        j1 = F - j + 1;
        j2 = j + ((j & 1) ? 1 : -1);
        k1 -= 1; if (k1 <= 0) k1 += n1;
        k2 += 1; if (k2 > n1) k2 -= n1;
        acc += fma(tp.g[j1 - 1], p1[(k1 - 1) * w1.es], tp.h[j2 - 1] * p2[(k2 - 1) * w2.es]);
    }
    long l = i0b - sh;                             // sh = 1: isidwt_step!(..., s = true) siwt/siwt_one_level.jl:168
    if (l < 0) l += n;
    v.p[voff(v, b0, b1, b2) + l * v.es] = acc;
}

// a9 sdwt_step! swt/swt_one_level.jl:114-125  /  a16 acdwt_step! acwt/acwt_one_level.jl:115-126
template <typename T, int AC, int FF>
__global__ void __launch_bounds__(kThreads) rdwt_step_k(View<T> w1, View<T> w2, View<const T> v, long n, long D, Geo geo, Taps<T> tp)
{
    long i, b0, b1, b2;
    if (!decomp(geo, i, b0, b1, b2)) return;
    const T *pv = v.p + voff(v, b0, b1, b2);
    const int F = WX_F;
    const long Dm = D % n;
    if (AC == 0) {
        long k1 = wx_wrapl(i - D, n), k2 = i;
        T a1 = tp.g[F - 1] * pv[k1 * v.es];
        T a2 = tp.h[0] * pv[k2 * v.es];
#pragma unroll
        for (int j = 1; j < F; ++j) {
            k1 += Dm; if (k1 >= n) k1 -= n;
            k2 -= Dm; if (k2 < 0) k2 += n;
            a1 = fma(tp.g[F - 1 - j], pv[k1 * v.es], a1);
            a2 = fma(tp.h[j], pv[k2 * v.es], a2);
        }
        w1.p[voff(w1, b0, b1, b2) + i * w1.es] = a1;
        w2.p[voff(w2, b0, b1, b2) + i * w2.es] = a2;
    } else {
        long t = wx_wrapl(i + D, n);                                   // 0-based image of t = i+2^d
        long io = wx_wrapl(i + (long)(F / 2 + 1) * D, n);              // output position
        T xv = pv[t * v.es];
        T a1 = tp.g[0] * xv, a2 = tp.h[0] * xv;
#pragma unroll
        for (int k = 1; k < F; ++k) {
            t += Dm; if (t >= n) t -= n;
            xv = pv[t * v.es];
            a1 = fma(tp.g[k], xv, a1);
            a2 = fma(tp.h[k], xv, a2);
        }
        w1.p[voff(w1, b0, b1, b2) + io * w1.es] = a1;
        w2.p[voff(w2, b0, b1, b2) + io * w2.es] = a2;
    }
}

// one output of the shift-based inverse stationary step, swt/swt_one_level.jl:301-315 ; t is 1-based
template <typename T, int FF>
__device__ __forceinline__ T isdwt_elem(const T *p1, long e1, const T *p2, long e2, long n, int d, long sw, long t,
                                        const Taps<T> &tp, bool has_init, T init)
{
    const int F = WX_F;
    const long sc = 1L << (d + 1), ic = sw + 1;
    int i0 = (t & 1) ? 1 : 2;
    int i1 = F - i0 + 1;
    int i2 = ((t + 1) & 1) ? 1 : 2;
    long k1 = ((t - 1) >> 1) * sc + ic, k2 = k1;
    T acc;
    if (has_init) acc = fma(tp.h[i2 - 1], p2[(k2 - 1) * e2], fma(tp.g[i1 - 1], p1[(k1 - 1) * e1], init));
    else          acc = fma(tp.g[i1 - 1], p1[(k1 - 1) * e1], tp.h[i2 - 1] * p2[(k2 - 1) * e2]);
#pragma unroll
    for (int ii = 1; ii < (F + 1) / 2; ++ii) {
        const int i = i0 + 2 * ii;
        if (i > F) break;
        i1 = F - i + 1;
        i2 = i + ((i & 1) ? 1 : -1);
        k1 -= sc; if (k1 <= 0) k1 = wx_wrapl(k1 - 1, n) + 1;
        k2 += sc; if (k2 > n) k2 = wx_wrapl(k2 - 1, n) + 1;
        acc += fma(tp.g[i1 - 1], p1[(k1 - 1) * e1], tp.h[i2 - 1] * p2[(k2 - 1) * e2]);
    }
    return acc;
}

// a10 isdwt_step! shift based: one thread per coset element
template <typename T, int FF>
__global__ void __launch_bounds__(kThreads) isdwt_shift_k(View<T> v, View<const T> w1, View<const T> w2, long n, int d, long sv, long sw,
                                                           int add2out, Geo geo, Taps<T> tp)
{
    long t0, b0, b1, b2;
    if (!decomp(geo, t0, b0, b1, b2)) return;
    const T *p1 = w1.p + voff(w1, b0, b1, b2);
    const T *p2 = w2.p + voff(w2, b0, b1, b2);
    T *pv = v.p + voff(v, b0, b1, b2);
    const long D = 1L << d;
    long t = t0 + 1, m = sv + 1 + t0 * D;                              // 1-based
    long j = (sw == sv) ? wx_wrapl(m - D - 1, n) : wx_wrapl(m - 1, n); // 0-based write position
    T init = add2out ? pv[j * v.es] : (T)0;
    pv[j * v.es] = isdwt_elem<T, FF>(p1, w1.es, p2, w2.es, n, d, sw, t, tp, add2out != 0, init);
}

// the same step with the filter length known at compile time (even F), 32-bit indices and the parity of t resolved once: the element
// routine above spends ~280 instructions per output on run-time tap indices and 64-bit wraps.  Same terms in the same order.
template <typename T, int FF>
__global__ void __launch_bounds__(kThreads) isdwt_shift_fast_k(View<T> v, View<const T> w1, View<const T> w2, int n, int D, int sv, int sw,
                                                                int add2out, Geo geo, Taps<T> tp)
{
    constexpr int F = FF > 0 ? FF : 2, R = F / 2;
    long t0, b0, b1, b2;
    if (!decomp(geo, t0, b0, b1, b2)) return;
    const T *p1 = w1.p + voff(w1, b0, b1, b2);
    const T *p2 = w2.p + voff(w2, b0, b1, b2);
    T *pv = v.p + voff(v, b0, b1, b2);
    const int t = (int)t0 + 1, m = sv + 1 + (int)t0 * D;               // 1-based
    int j = ((sw == sv) ? m - D - 1 : m - 1) % n; if (j < 0) j += n;   // 0-based write position
    const bool odd = (t & 1) != 0;
    const int sc = 2 * D;
    int k1 = ((t - 1) >> 1) * sc + sw, k2 = k1;                        // 0-based
    // term r: t odd -> g[F-1-2r], h[2r+1]; t even -> g[F-2-2r], h[2r]
    T gr = odd ? tp.g[F - 1] : tp.g[F - 2], hr = odd ? tp.h[1] : tp.h[0];
    T acc;
    if (add2out) acc = fma(hr, p2[(long)k2 * w2.es], fma(gr, p1[(long)k1 * w1.es], pv[(long)j * v.es]));
    else         acc = fma(gr, p1[(long)k1 * w1.es], hr * p2[(long)k2 * w2.es]);
#pragma unroll
    for (int r = 1; r < R; ++r) {
        k1 -= sc; if (k1 < 0) k1 += n;
        k2 += sc; if (k2 >= n) k2 -= n;
        gr = odd ? tp.g[F - 1 - 2 * r] : tp.g[F - 2 - 2 * r];
        hr = odd ? tp.h[2 * r + 1] : tp.h[2 * r];
        acc += fma(gr, p1[(long)k1 * w1.es], hr * p2[(long)k2 * w2.es]);
    }
    pv[(long)j * v.es] = acc;
}

// isdwt!(xw, wt, sm) SWT.jl:270-282 as ONE launch: a CTA keeps the running reconstruction of one signal in shared memory (ping-pong) and
// walks the levels d = L-1 .. 0; level d writes the sv(d) coset from the sw(d) = sv(d+1) coset of the previous level and of detail column
// L-d (read straight from the table), so nothing off those cosets is ever needed.  Replaces L copies of x plus L step launches.
struct WxShiftList { int s[34]; };
template <typename T, int FF>
__global__ void __launch_bounds__(kThreads) isdwt_shift_chain_k(T *__restrict__ x, const T *__restrict__ xw, int n, int L, WxShiftList sl, Taps<T> tp)
{
    constexpr int F = FF > 0 ? FF : 2, R = F / 2;
    extern __shared__ __align__(16) unsigned char wx_chain_smem[];
    T *cur = reinterpret_cast<T *>(wx_chain_smem), *nxt = cur + n;
    const long k = blockIdx.x;
    const T *col = xw + k * (long)(L + 1) * n;
    const int tid = threadIdx.x;
    for (int i = tid; i < n; i += kThreads) cur[i] = col[i];      // (cp.async element copies measured slower here: 0.15 -> 0.19 ms)
    __syncthreads();
    for (int d = L - 1; d >= 0; --d) {
        const int D = 1 << d, sc = 2 * D, sv = sl.s[d], sw = sl.s[d + 1];
        const int cnt = (n - 1 - sv) / D + 1;
        const T *det = col + (long)(L - d) * n;
        for (int t0 = tid; t0 < cnt; t0 += kThreads) {
            const int t = t0 + 1, m = sv + 1 + t0 * D;
            int j = ((sw == sv) ? m - D - 1 : m - 1) % n; if (j < 0) j += n;
            const bool odd = (t & 1) != 0;
            int k1 = ((t - 1) >> 1) * sc + sw, k2 = k1;
            T gr = odd ? tp.g[F - 1] : tp.g[F - 2], hr = odd ? tp.h[1] : tp.h[0];
            T acc = fma(gr, cur[k1], hr * det[k2]);
#pragma unroll
            for (int r = 1; r < R; ++r) {
                k1 -= sc; if (k1 < 0) k1 += n;
                k2 += sc; if (k2 >= n) k2 -= n;
                gr = odd ? tp.g[F - 1 - 2 * r] : tp.g[F - 2 - 2 * r];
                hr = odd ? tp.h[2 * r + 1] : tp.h[2 * r];
                acc += fma(gr, cur[k1], hr * det[k2]);
            }
            nxt[j] = acc;
        }
        __syncthreads();
        T *tsw = cur; cur = nxt; nxt = tsw;
    }
    for (int i = tid; i < n; i += kThreads) x[k * n + i] = cur[i];
}

// a10 isdwt_step! average based: one thread per output position (requires n % 2^(d+1) == 0)
template <typename T, int FF>
__global__ void __launch_bounds__(kThreads) isdwt_avg_k(View<T> v, View<const T> w1, View<const T> w2, long n, int d, Geo geo, Taps<T> tp)
{
    long pos, b0, b1, b2;
    if (!decomp(geo, pos, b0, b1, b2)) return;
    const T *p1 = w1.p + voff(w1, b0, b1, b2);
    const T *p2 = w2.p + voff(w2, b0, b1, b2);
    const long D = 1L << d;
    const long sv = pos & (D - 1);
    long m1 = pos + 1 + D; if (m1 > n) m1 -= n;          // first call (sw = sv) writes position m-D
    long t1 = ((m1 - 1 - sv) >> d) + 1;
    long t2 = ((pos - sv) >> d) + 1;                     // second call (sw = sv+D) adds at position m
    T a = isdwt_elem<T, FF>(p1, w1.es, p2, w2.es, n, d, sv, t1, tp, false, (T)0);
    a = isdwt_elem<T, FF>(p1, w1.es, p2, w2.es, n, d, sv + D, t2, tp, true, a);
    v.p[voff(v, b0, b1, b2) + pos * v.es] = a / (T)2;
}

// the same step in closed form: the two shifted reconstructions the reference averages (swt/swt_one_level.jl:257-277) add up to half
// the adjoint of the a-trous analysis step,
//     v[p] = 1/2 ( sum_j g[j] w1[p + (j+2-F) D] + sum_j h[j] w2[p + j D] ),   D = 2^d, indices mod n
// (the fused tree kernels of wx_irwpd_fused.cu accumulate in exactly this order).  2F loads and 2F FMAs per output with
// compile-time taps and 32-bit indices instead of two calls of the literal element routine (ncu: 560 instructions per output).
template <typename T, int FF>
__global__ void __launch_bounds__(kThreads) isdwt_avg_fast_k(View<T> v, View<const T> w1, View<const T> w2, int n, int D, Geo geo, Taps<T> tp)
{
    constexpr int F = FF > 0 ? FF : 2;
    long pos, b0, b1, b2;
    if (!decomp(geo, pos, b0, b1, b2)) return;
    const T *p1 = w1.p + voff(w1, b0, b1, b2);
    const T *p2 = w2.p + voff(w2, b0, b1, b2);
    int i1 = ((int)pos + (2 - F) * D) % n; if (i1 < 0) i1 += n;
    int i2 = (int)pos;
    T s = tp.g[0] * p1[(long)i1 * w1.es];
#pragma unroll
    for (int j = 1; j < F; ++j) { i1 += D; if (i1 >= n) i1 -= n; s = fma(tp.g[j], p1[(long)i1 * w1.es], s); }
#pragma unroll
    for (int j = 0; j < F; ++j) { s = fma(tp.h[j], p2[(long)i2 * w2.es], s); i2 += D; if (i2 >= n) i2 -= n; }
    v.p[voff(v, b0, b1, b2) + pos * v.es] = s * (T)0.5;
}

// a17 iacdwt_step! acwt/acwt_one_level.jl:221-223
template <typename T>
__global__ void __launch_bounds__(kThreads) iacdwt_step_k(View<T> v, View<const T> w1, View<const T> w2, long n, Geo geo)
{
    long i, b0, b1, b2;
    if (!decomp(geo, i, b0, b1, b2)) return;
    T a = w1.p[voff(w1, b0, b1, b2) + i * w1.es] + w2.p[voff(w2, b0, b1, b2) + i * w2.es];
    v.p[voff(v, b0, b1, b2) + i * v.es] = a / (T)1.4142135623730951;
}

template <typename T>
__global__ void __launch_bounds__(kThreads) copy_k(View<T> dst, View<const T> src, long n, Geo geo)
{
    long i, b0, b1, b2;
    if (!decomp(geo, i, b0, b1, b2)) return;
    dst.p[voff(dst, b0, b1, b2) + i * dst.es] = src.p[voff(src, b0, b1, b2) + i * src.es];
}

// instantiate KERNEL<..., FF> for the filter lengths of the shipped wavelets (orthogonal F and autocorrelation 2F-1), else FF = 0
#define WX_DISPATCH_F(Fv, CALL)                                                                                           \
    switch (Fv) {                                                                                                         \
        case 2: { constexpr int FF = 2; CALL; } break;   case 3: { constexpr int FF = 3; CALL; } break;                   \
        case 4: { constexpr int FF = 4; CALL; } break;   case 6: { constexpr int FF = 6; CALL; } break;                   \
        case 7: { constexpr int FF = 7; CALL; } break;   case 8: { constexpr int FF = 8; CALL; } break;                   \
        case 10: { constexpr int FF = 10; CALL; } break; case 11: { constexpr int FF = 11; CALL; } break;                 \
        case 12: { constexpr int FF = 12; CALL; } break; case 15: { constexpr int FF = 15; CALL; } break;                 \
        case 16: { constexpr int FF = 16; CALL; } break; case 23: { constexpr int FF = 23; CALL; } break;                 \
        case 31: { constexpr int FF = 31; CALL; } break; case 14: { constexpr int FF = 14; CALL; } break;                 \
        case 18: { constexpr int FF = 18; CALL; } break; case 20: { constexpr int FF = 20; CALL; } break;                 \
        case 24: { constexpr int FF = 24; CALL; } break;                                                                  \
        default: { constexpr int FF = 0; CALL; } break;                                                                   \
    }

static inline long btot(const Batch &b) { return b.B0 * b.B1 * b.B2; }

}  // namespace

template <typename T>
int wx_launch_dwt_step(View<T> w1, View<T> w2, View<const T> v, long n, Batch b, const Taps<T> &t, cudaStream_t s, int shift)
{
    WX_REQUIRE(n >= 2 && n % 2 == 0, "dwt_step: parent length %ld must be even and >= 2", n);
    long total = (n / 2) * btot(b);
    if (total == 0) return WX_OK;
    const Geo geo = make_geo(n / 2, b);
    WX_DISPATCH_F(t.F, (dwt_step_k<T, FF><<<grid_for(total), kThreads, 0, s>>>(w1, w2, v, n, geo, t, shift)))
    WX_LAUNCHED();
    return WX_OK;
}

template <typename T>
int wx_launch_idwt_step(View<T> v, View<const T> w1, View<const T> w2, long n, Batch b, const Taps<T> &t, cudaStream_t s, int shift)
{
    WX_REQUIRE(n >= 2 && n % 2 == 0, "idwt_step: parent length %ld must be even and >= 2", n);
    long total = n * btot(b);
    if (total == 0) return WX_OK;
    const Geo geo = make_geo(n, b);
    WX_DISPATCH_F(t.F, (idwt_step_k<T, FF><<<grid_for(total), kThreads, 0, s>>>(v, w1, w2, n, geo, t, shift)))
    WX_LAUNCHED();
    return WX_OK;
}

template <typename T>
int wx_launch_rdwt_step(int ac, View<T> w1, View<T> w2, View<const T> v, long n, int d, Batch b, const Taps<T> &t, cudaStream_t s)
{
    WX_REQUIRE(n >= 1 && d >= 0 && d < 62, "rdwt_step: bad n=%ld or d=%d", n, d);
    long total = n * btot(b);
    if (total == 0) return WX_OK;
    const Geo geo = make_geo(n, b);
    if (ac) { WX_DISPATCH_F(t.F, (rdwt_step_k<T, 1, FF><<<grid_for(total), kThreads, 0, s>>>(w1, w2, v, n, 1L << d, geo, t))) }
    else    { WX_DISPATCH_F(t.F, (rdwt_step_k<T, 0, FF><<<grid_for(total), kThreads, 0, s>>>(w1, w2, v, n, 1L << d, geo, t))) }
    WX_LAUNCHED();
    return WX_OK;
}

template <typename T>
int wx_launch_isdwt_shift(View<T> v, View<const T> w1, View<const T> w2, long n, int d, long sv, long sw, int add2out, Batch b,
                          const Taps<T> &t, cudaStream_t s)
{
    // the reference's @assert (swt/swt_one_level.jl:289-290)
    WX_REQUIRE(d >= 0 && d < 61, "isdwt_step: bad depth %d", d);
    WX_REQUIRE(0 <= sv && sv < (1L << d), "AssertionError: 0 <= sv < 1<<d (sv=%ld, d=%d)", sv, d);
    WX_REQUIRE(sv <= sw && sw < (1L << (d + 1)), "AssertionError: sv <= sw < 1<<(d+1) (sv=%ld, sw=%ld, d=%d)", sv, sw, d);
    WX_REQUIRE(n % (1L << (d + 1)) == 0, "isdwt_step: n=%ld must be a multiple of 2^(d+1)", n);
    long cnt = (n - 1 - sv) / (1L << d) + 1;
    long total = cnt * btot(b);
    if (total <= 0) return WX_OK;
    const Geo geo = make_geo(cnt, b);
    static const bool literal = getenv("WX_B200_ISDWT_SHIFT_LITERAL") != nullptr;     // A-B measurements
    bool known = false;
    WX_DISPATCH_F(t.F, (known = FF > 0 && FF % 2 == 0))
    if (!literal && known && n < (1L << 30) && (1L << d) < n) {                      // one conditional wrap per step needs 2^(d+1) <= n (checked above)
        WX_DISPATCH_F(t.F, (isdwt_shift_fast_k<T, FF><<<grid_for(total), kThreads, 0, s>>>(v, w1, w2, (int)n, (int)(1L << d), (int)sv, (int)sw, add2out, geo, t)))
    } else {
        WX_DISPATCH_F(t.F, (isdwt_shift_k<T, FF><<<grid_for(total), kThreads, 0, s>>>(v, w1, w2, n, d, sv, sw, add2out, geo, t)))
    }
    WX_LAUNCHED();
    return WX_OK;
}

template <typename T>
int wx_launch_isdwt_avg(View<T> v, View<const T> w1, View<const T> w2, long n, int d, Batch b, const Taps<T> &t, cudaStream_t s)
{
    WX_REQUIRE(d >= 0 && d < 61, "isdwt_step: bad depth %d", d);
    WX_REQUIRE(n % (1L << (d + 1)) == 0, "isdwt_step: n=%ld must be a multiple of 2^(d+1)", n);
    long total = n * btot(b);
    if (total == 0) return WX_OK;
    const Geo geo = make_geo(n, b);
    static const bool literal = getenv("WX_B200_ISDWT_AVG_LITERAL") != nullptr;       // A-B measurements: the two-call form of the reference
    const long D = 1L << d;
    bool known = false;                                                              // compile-time filter length available?
    WX_DISPATCH_F(t.F, (known = FF > 0))
    if (!literal && known && n < (1L << 30) && (long)t.F * D < (1L << 30)) {
        WX_DISPATCH_F(t.F, (isdwt_avg_fast_k<T, FF><<<grid_for(total), kThreads, 0, s>>>(v, w1, w2, (int)n, (int)D, geo, t)))
    } else {
        WX_DISPATCH_F(t.F, (isdwt_avg_k<T, FF><<<grid_for(total), kThreads, 0, s>>>(v, w1, w2, n, d, geo, t)))
    }
    WX_LAUNCHED();
    return WX_OK;
}

namespace {
template <typename T, int FF>
int shift_chain_launch(T *x, const T *xw, long n, int L, long N, const WxShiftList &sl, size_t smem, const Taps<T> &t, cudaStream_t s)
{
    auto kern = isdwt_shift_chain_k<T, FF>;
    WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<(unsigned)N, kThreads, smem, s>>>(x, xw, (int)n, L, sl, t);
    return WX_OK;
}
}  // namespace

// the whole shift-based isdwt! chain in one launch (see isdwt_shift_chain_k); *handled = false: shape not covered, nothing launched
template <typename T>
int wx_isdwt_shift_chain(T *x, const T *xw, long n, int L, long N, const long *sd, const Taps<T> &t, cudaStream_t s, bool *handled)
{
    *handled = false;
    static const bool off = getenv("WX_B200_NO_SHIFT_CHAIN") != nullptr;            // A-B measurements
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    bool known = false;
    WX_DISPATCH_F(t.F, (known = FF > 0 && FF % 2 == 0))
    const size_t smem = (size_t)2 * n * sizeof(T);
    if (off || !known || L < 1 || L > 32 || N < 1 || N >= (1L << 31) || n >= (1L << 28) || smem > dv.smem_optin || n % (1L << L) != 0) return WX_OK;
    WxShiftList sl;
    for (int d = 0; d <= L; ++d) sl.s[d] = (int)sd[d];
    WX_DISPATCH_F(t.F, (rc = shift_chain_launch<T, FF>(x, xw, n, L, N, sl, smem, t, s)))
    if (rc) return rc;
    WX_LAUNCHED();
    *handled = true;
    return WX_OK;
}
template int wx_isdwt_shift_chain<double>(double *, const double *, long, int, long, const long *, const Taps<double> &, cudaStream_t, bool *);
template int wx_isdwt_shift_chain<float>(float *, const float *, long, int, long, const long *, const Taps<float> &, cudaStream_t, bool *);

template <typename T>
int wx_launch_iacdwt_step(View<T> v, View<const T> w1, View<const T> w2, long n, Batch b, cudaStream_t s)
{
    long total = n * btot(b);
    if (total == 0) return WX_OK;
    iacdwt_step_k<T><<<grid_for(total), kThreads, 0, s>>>(v, w1, w2, n, make_geo(n, b));
    WX_LAUNCHED();
    return WX_OK;
}

template <typename T>
int wx_launch_copy(View<T> dst, View<const T> src, long n, Batch b, cudaStream_t s)
{
    long total = n * btot(b);
    if (total == 0) return WX_OK;
    copy_k<T><<<grid_for(total), kThreads, 0, s>>>(dst, src, n, make_geo(n, b));
    WX_LAUNCHED();
    return WX_OK;
}

#define WX_INST(T)                                                                                                               \
    template int wx_launch_dwt_step<T>(View<T>, View<T>, View<const T>, long, Batch, const Taps<T> &, cudaStream_t, int);             \
    template int wx_launch_idwt_step<T>(View<T>, View<const T>, View<const T>, long, Batch, const Taps<T> &, cudaStream_t, int);      \
    template int wx_launch_rdwt_step<T>(int, View<T>, View<T>, View<const T>, long, int, Batch, const Taps<T> &, cudaStream_t);  \
    template int wx_launch_isdwt_shift<T>(View<T>, View<const T>, View<const T>, long, int, long, long, int, Batch,              \
                                          const Taps<T> &, cudaStream_t);                                                        \
    template int wx_launch_isdwt_avg<T>(View<T>, View<const T>, View<const T>, long, int, Batch, const Taps<T> &, cudaStream_t); \
    template int wx_launch_iacdwt_step<T>(View<T>, View<const T>, View<const T>, long, Batch, cudaStream_t);                     \
    template int wx_launch_copy<T>(View<T>, View<const T>, long, Batch, cudaStream_t);
WX_INST(double)
WX_INST(float)

// ================================================================================================
// C ABI: single steps
// ================================================================================================
namespace {

template <typename T> static View<const T> cview1(const T *p) { return View<const T>{p, 1, 0, 0, 0}; }

template <typename T>
int dwt_step_1d(T *w1, T *w2, const T *v, long n, const double *h, const double *g, int F, void *stream, int shift = 0)
{
    WX_REQUIRE(w1 && w2 && v, "null signal pointer");
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    return wx_launch_dwt_step<T>(view1(w1), view1(w2), cview1(v), n, batch1(), t, (cudaStream_t)stream, shift ? 1 : 0);
}
template <typename T>
int idwt_step_1d(T *v, const T *w1, const T *w2, long n, const double *h, const double *g, int F, void *stream, int shift = 0)
{
    WX_REQUIRE(w1 && w2 && v, "null signal pointer");
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    return wx_launch_idwt_step<T>(view1(v), cview1(w1), cview1(w2), n, batch1(), t, (cudaStream_t)stream, shift ? 1 : 0);
}

// ---- nonstandard-form transform (row f-4): ns_dwt / ns_idwt  wavemult/transforms.jl:52-74, 120-139 ----
// ndyad(l, Lmax, gender) wavemult/utils.jl:146-155 -> 0-based start of the range, length 1 << (Lmax - l)
static inline long ndyad0(int l, int Lmax, bool female) { const long k = Lmax - l; return female ? (1L << (k + 1)) + (1L << k) : (1L << (k + 1)); }

template <typename T>
__global__ void __launch_bounds__(kThreads) ns_add_k(T *__restrict__ out, const T *__restrict__ a, long as0, const T *__restrict__ b, long bs0, long len, long N)
{
    const long idx = (long)blockIdx.x * kThreads + threadIdx.x;
    if (idx >= len * N) return;
    const long k = idx / len, i = idx - k * len;
    out[idx] = a[k * as0 + i] + b[k * bs0 + i];
}

template <typename T>
int ns_dwt_impl(T *nxw, const T *x, long n, int L, long N, const double *h, const double *g, int F, cudaStream_t s)
{
    WX_REQUIRE(nxw && x && N >= 0, "bad arguments");
    WX_REQUIRE(wx_ispow2(n), "AssertionError: ispow2(n)");
    const int Lmax = wx_maxlevels(n);
    WX_REQUIRE(1 <= L && L <= Lmax, "AssertionError: 1 <= L <= Lmax");
    if (N == 0) return WX_OK;
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    WX_CUDA(cudaMemsetAsync(nxw, 0, (size_t)2 * n * N * sizeof(T), s));
    const Batch b{N, 1, 1, false};
    for (int l = 1; l <= L; ++l) {
        View<T> w1{nxw + ndyad0(l, Lmax, false), 1, 2 * n, 0, 0}, w2{nxw + ndyad0(l, Lmax, true), 1, 2 * n, 0, 0};
        View<const T> v = l == 1 ? View<const T>{x, 1, n, 0, 0} : View<const T>{nxw + ndyad0(l - 1, Lmax, false), 1, 2 * n, 0, 0};
        rc = wx_launch_dwt_step<T>(w1, w2, v, n >> (l - 1), b, t, s); if (rc) return rc;
    }
    // nxw[1 : 1<<(Lmax-L)] = nxw[ndyad(L, Lmax, false)]
    return wx_launch_copy<T>(View<T>{nxw, 1, 2 * n, 0, 0}, View<const T>{nxw + ndyad0(L, Lmax, false), 1, 2 * n, 0, 0}, 1L << (Lmax - L), b, s);
}

template <typename T>
int ns_idwt_impl(T *x, const T *nxw, long n2, int L, long N, const double *h, const double *g, int F, cudaStream_t s)
{
    WX_REQUIRE(nxw && x && N >= 0 && n2 >= 2, "bad arguments");
    const long n = n2 / 2;
    WX_REQUIRE(wx_ispow2(n) && n2 == 2 * n, "AssertionError: ispow2(n)");
    const int Lmax = wx_maxlevels(n2) - 1;
    WX_REQUIRE(1 <= L && L <= Lmax, "AssertionError: 1 <= L <= Lmax");
    if (N == 0) return WX_OK;
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    WX_CUDA(cudaMemsetAsync(x, 0, (size_t)n * N * sizeof(T), s));
    const Batch b{N, 1, 1, false};
    rc = wx_launch_copy<T>(View<T>{x, 1, n, 0, 0}, View<const T>{nxw, 1, n2, 0, 0}, 1L << (Lmax - L), b, s); if (rc) return rc;
    T *tmp; rc = wx_scratch(&tmp, (size_t)(n / 2) * N, s); if (rc) return rc;
    for (int l = L; l >= 1 && !rc; --l) {
        const long len = 1L << (Lmax - l);
        // w1 = nxw[ndyad(l, Lmax, false)] + x[1 : len]  (the reference materialises it before idwt_step! overwrites x)
        ns_add_k<T><<<grid_for(len * N), kThreads, 0, s>>>(tmp, nxw + ndyad0(l, Lmax, false), n2, x, n, len, N);
        wx_launches.fetch_add(1, std::memory_order_relaxed);
        rc = wx_launch_idwt_step<T>(View<T>{x, 1, n, 0, 0}, View<const T>{tmp, 1, len, 0, 0}, View<const T>{nxw + ndyad0(l, Lmax, true), 1, n2, 0, 0}, 2 * len, b, t, s);
    }
    int rc2 = wx_scratch_free(tmp, s);
    if (!rc) WX_CUDA(cudaGetLastError());
    return rc ? rc : rc2;
}
template <typename T>
int rdwt_step_1d(int ac, T *w1, T *w2, const T *v, long n, int d, const double *h, const double *g, int F, void *stream)
{
    WX_REQUIRE(w1 && w2 && v, "null signal pointer");
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    return wx_launch_rdwt_step<T>(ac, view1(w1), view1(w2), cview1(v), n, d, batch1(), t, (cudaStream_t)stream);
}
template <typename T>
int isdwt_shift_1d(T *v, const T *w1, const T *w2, long n, int d, long sv, long sw, const double *h, const double *g, int F,
                   int add2out, void *stream)
{
    WX_REQUIRE(w1 && w2 && v, "null signal pointer");
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    return wx_launch_isdwt_shift<T>(view1(v), cview1(w1), cview1(w2), n, d, sv, sw, add2out, batch1(), t, (cudaStream_t)stream);
}
template <typename T>
int isdwt_avg_1d(T *v, const T *w1, const T *w2, long n, int d, const double *h, const double *g, int F, void *stream)
{
    WX_REQUIRE(w1 && w2 && v, "null signal pointer");
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    return wx_launch_isdwt_avg<T>(view1(v), cview1(w1), cview1(w2), n, d, batch1(), t, (cudaStream_t)stream);
}

// 2-D steps on contiguous column-major matrices -------------------------------------------------
// dwt_step! 2-D  dwt/dwt_one_level.jl:335-352 : columns into temp, then rows
template <typename T>
int dwt_step_2d(T *w1, T *w2, T *w3, T *w4, const T *v, long nr, long nc, const double *h, const double *g, int F, void *stream)
{
    WX_REQUIRE(w1 && w2 && w3 && w4 && v, "null signal pointer");
    WX_REQUIRE(nr >= 1 && nc >= 1, "dwt_step 2-D: empty child");
    cudaStream_t s = (cudaStream_t)stream;
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    T *temp; rc = wx_scratch(&temp, (size_t)4 * nr * nc, s); if (rc) return rc;
    const long ld = 2 * nr;
    // columns: problem b0 = column j (stride ld), contiguous elements
    rc = wx_launch_dwt_step<T>(View<T>{temp, 1, ld, 0, 0}, View<T>{temp + nr, 1, ld, 0, 0}, View<const T>{v, 1, ld, 0, 0}, 2 * nr,
                               Batch{2 * nc, 1, 1, false}, t, s);
    // rows: problem b0 = row i (stride 1), element stride = leading dimension
    if (!rc) rc = wx_launch_dwt_step<T>(View<T>{w1, nr, 1, 0, 0}, View<T>{w2, nr, 1, 0, 0}, View<const T>{temp, ld, 1, 0, 0}, 2 * nc,
                                        Batch{nr, 1, 1, true}, t, s);
    if (!rc) rc = wx_launch_dwt_step<T>(View<T>{w3, nr, 1, 0, 0}, View<T>{w4, nr, 1, 0, 0}, View<const T>{temp + nr, ld, 1, 0, 0}, 2 * nc,
                                        Batch{nr, 1, 1, true}, t, s);
    int rc2 = wx_scratch_free(temp, s);
    return rc ? rc : rc2;
}

// idwt_step! 2-D  dwt/dwt_one_level.jl:417-434 : rows into temp, then columns
template <typename T>
int idwt_step_2d(T *v, const T *w1, const T *w2, const T *w3, const T *w4, long nr, long nc, const double *h, const double *g, int F,
                 void *stream)
{
    WX_REQUIRE(w1 && w2 && w3 && w4 && v, "null signal pointer");
    WX_REQUIRE(nr >= 1 && nc >= 1, "idwt_step 2-D: empty child");
    cudaStream_t s = (cudaStream_t)stream;
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    T *temp; rc = wx_scratch(&temp, (size_t)4 * nr * nc, s); if (rc) return rc;
    const long ld = 2 * nr;
    rc = wx_launch_idwt_step<T>(View<T>{temp, ld, 1, 0, 0}, View<const T>{w1, nr, 1, 0, 0}, View<const T>{w2, nr, 1, 0, 0}, 2 * nc,
                                Batch{nr, 1, 1, true}, t, s);
    if (!rc) rc = wx_launch_idwt_step<T>(View<T>{temp + nr, ld, 1, 0, 0}, View<const T>{w3, nr, 1, 0, 0}, View<const T>{w4, nr, 1, 0, 0},
                                         2 * nc, Batch{nr, 1, 1, true}, t, s);
    if (!rc) rc = wx_launch_idwt_step<T>(View<T>{v, 1, ld, 0, 0}, View<const T>{temp, 1, ld, 0, 0}, View<const T>{temp + nr, 1, ld, 0, 0},
                                         2 * nr, Batch{2 * nc, 1, 1, false}, t, s);
    int rc2 = wx_scratch_free(temp, s);
    return rc ? rc : rc2;
}

// sdwt_step!/acdwt_step! 2-D  swt/swt_one_level.jl:352-368, acwt/acwt_one_level.jl:258-274 ; all (nr x nc)
template <typename T>
int rdwt_step_2d(int ac, T *w1, T *w2, T *w3, T *w4, const T *v, long nr, long nc, int d, const double *h, const double *g, int F,
                 void *stream)
{
    WX_REQUIRE(w1 && w2 && w3 && w4 && v, "null signal pointer");
    cudaStream_t s = (cudaStream_t)stream;
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    T *temp; rc = wx_scratch(&temp, (size_t)2 * nr * nc, s); if (rc) return rc;
    T *t1 = temp, *t2 = temp + nr * nc;
    rc = wx_launch_rdwt_step<T>(ac, View<T>{t1, 1, nr, 0, 0}, View<T>{t2, 1, nr, 0, 0}, View<const T>{v, 1, nr, 0, 0}, nr, d,
                                Batch{nc, 1, 1, false}, t, s);
    if (!rc) rc = wx_launch_rdwt_step<T>(ac, View<T>{w1, nr, 1, 0, 0}, View<T>{w2, nr, 1, 0, 0}, View<const T>{t1, nr, 1, 0, 0}, nc, d,
                                         Batch{nr, 1, 1, true}, t, s);
    if (!rc) rc = wx_launch_rdwt_step<T>(ac, View<T>{w3, nr, 1, 0, 0}, View<T>{w4, nr, 1, 0, 0}, View<const T>{t2, nr, 1, 0, 0}, nc, d,
                                         Batch{nr, 1, 1, true}, t, s);
    int rc2 = wx_scratch_free(temp, s);
    return rc ? rc : rc2;
}

// isdwt_step! 2-D swt/swt_one_level.jl:395-469 (mode 0 average, 1 shift) ; iacdwt_step! 2-D acwt/acwt_one_level.jl:288-322 (mode 2)
template <typename T>
int irdwt_step_2d(int mode, T *v, const T *w1, const T *w2, const T *w3, const T *w4, long nr, long nc, int d, long sv, long sw,
                  const double *h, const double *g, int F, void *stream)
{
    WX_REQUIRE(w1 && w2 && w3 && w4 && v, "null signal pointer");
    WX_REQUIRE(mode >= 0 && mode <= 2, "bad inverse mode %d", mode);
    cudaStream_t s = (cudaStream_t)stream;
    Taps<T> t;
    int rc;
    if (mode == 2) { memset(&t, 0, sizeof(t)); t.F = 1; }
    else { rc = wx_make_taps(t, h, g, F); if (rc) return rc; }
    T *temp; rc = wx_scratch(&temp, (size_t)2 * nr * nc, s); if (rc) return rc;
    T *t1 = temp, *t2 = temp + nr * nc;
    if (mode == 1) WX_CUDA(cudaMemsetAsync(temp, 0, (size_t)2 * nr * nc * sizeof(T), s));   // positions off the coset stay defined
    auto step = [&](View<T> vo, View<const T> a, View<const T> b, long n, Batch bt) -> int {
        if (mode == 2) return wx_launch_iacdwt_step<T>(vo, a, b, n, bt, s);
        if (mode == 1) return wx_launch_isdwt_shift<T>(vo, a, b, n, d, sv, sw, 0, bt, t, s);
        return wx_launch_isdwt_avg<T>(vo, a, b, n, d, bt, t, s);
    };
    rc = step(View<T>{t1, nr, 1, 0, 0}, View<const T>{w1, nr, 1, 0, 0}, View<const T>{w2, nr, 1, 0, 0}, nc, Batch{nr, 1, 1, true});
    if (!rc) rc = step(View<T>{t2, nr, 1, 0, 0}, View<const T>{w3, nr, 1, 0, 0}, View<const T>{w4, nr, 1, 0, 0}, nc, Batch{nr, 1, 1, true});
    if (!rc) rc = step(View<T>{v, 1, nr, 0, 0}, View<const T>{t1, 1, nr, 0, 0}, View<const T>{t2, 1, nr, 0, 0}, nr, Batch{nc, 1, 1, false});
    int rc2 = wx_scratch_free(temp, s);
    return rc ? rc : rc2;
}

}  // namespace

extern "C" {

int wx_dwt_step_f64(double *w1, double *w2, const double *v, long n, const double *h, const double *g, int F, void *s) { return dwt_step_1d<double>(w1, w2, v, n, h, g, F, s); }
int wx_dwt_step_f32(float *w1, float *w2, const float *v, long n, const double *h, const double *g, int F, void *s) { return dwt_step_1d<float>(w1, w2, v, n, h, g, F, s); }
int wx_idwt_step_f64(double *v, const double *w1, const double *w2, long n, const double *h, const double *g, int F, void *s) { return idwt_step_1d<double>(v, w1, w2, n, h, g, F, s); }
int wx_idwt_step_f32(float *v, const float *w1, const float *w2, long n, const double *h, const double *g, int F, void *s) { return idwt_step_1d<float>(v, w1, w2, n, h, g, F, s); }
int wx_sidwt_step_f64(double *w1, double *w2, const double *v, long n, const double *h, const double *g, int F, int shifted, void *s) { return dwt_step_1d<double>(w1, w2, v, n, h, g, F, s, shifted); }
int wx_sidwt_step_f32(float *w1, float *w2, const float *v, long n, const double *h, const double *g, int F, int shifted, void *s) { return dwt_step_1d<float>(w1, w2, v, n, h, g, F, s, shifted); }
int wx_isidwt_step_f64(double *v, const double *w1, const double *w2, long n, const double *h, const double *g, int F, int shifted, void *s) { return idwt_step_1d<double>(v, w1, w2, n, h, g, F, s, shifted); }
int wx_isidwt_step_f32(float *v, const float *w1, const float *w2, long n, const double *h, const double *g, int F, int shifted, void *s) { return idwt_step_1d<float>(v, w1, w2, n, h, g, F, s, shifted); }
int wx_ns_dwt_f64(double *nxw, const double *x, long n, int L, long N, const double *h, const double *g, int F, void *s) { return ns_dwt_impl<double>(nxw, x, n, L, N, h, g, F, (cudaStream_t)s); }
int wx_ns_dwt_f32(float *nxw, const float *x, long n, int L, long N, const double *h, const double *g, int F, void *s) { return ns_dwt_impl<float>(nxw, x, n, L, N, h, g, F, (cudaStream_t)s); }
int wx_ns_idwt_f64(double *x, const double *nxw, long n2, int L, long N, const double *h, const double *g, int F, void *s) { return ns_idwt_impl<double>(x, nxw, n2, L, N, h, g, F, (cudaStream_t)s); }
int wx_ns_idwt_f32(float *x, const float *nxw, long n2, int L, long N, const double *h, const double *g, int F, void *s) { return ns_idwt_impl<float>(x, nxw, n2, L, N, h, g, F, (cudaStream_t)s); }
int wx_sdwt_step_f64(double *w1, double *w2, const double *v, long n, int d, const double *h, const double *g, int F, void *s) { return rdwt_step_1d<double>(0, w1, w2, v, n, d, h, g, F, s); }
int wx_sdwt_step_f32(float *w1, float *w2, const float *v, long n, int d, const double *h, const double *g, int F, void *s) { return rdwt_step_1d<float>(0, w1, w2, v, n, d, h, g, F, s); }
int wx_acdwt_step_f64(double *w1, double *w2, const double *v, long n, int d, const double *h, const double *g, int Lf, void *s) { return rdwt_step_1d<double>(1, w1, w2, v, n, d, h, g, Lf, s); }
int wx_acdwt_step_f32(float *w1, float *w2, const float *v, long n, int d, const double *h, const double *g, int Lf, void *s) { return rdwt_step_1d<float>(1, w1, w2, v, n, d, h, g, Lf, s); }
int wx_isdwt_step_shift_f64(double *v, const double *w1, const double *w2, long n, int d, long sv, long sw, const double *h, const double *g, int F, int add2out, void *s) { return isdwt_shift_1d<double>(v, w1, w2, n, d, sv, sw, h, g, F, add2out, s); }
int wx_isdwt_step_shift_f32(float *v, const float *w1, const float *w2, long n, int d, long sv, long sw, const double *h, const double *g, int F, int add2out, void *s) { return isdwt_shift_1d<float>(v, w1, w2, n, d, sv, sw, h, g, F, add2out, s); }
int wx_isdwt_step_avg_f64(double *v, const double *w1, const double *w2, long n, int d, const double *h, const double *g, int F, void *s) { return isdwt_avg_1d<double>(v, w1, w2, n, d, h, g, F, s); }
int wx_isdwt_step_avg_f32(float *v, const float *w1, const float *w2, long n, int d, const double *h, const double *g, int F, void *s) { return isdwt_avg_1d<float>(v, w1, w2, n, d, h, g, F, s); }
int wx_iacdwt_step_f64(double *v, const double *w1, const double *w2, long n, void *s)
{
    WX_REQUIRE(w1 && w2 && v, "null signal pointer");
    return wx_launch_iacdwt_step<double>(view1(v), cview1(w1), cview1(w2), n, batch1(), (cudaStream_t)s);
}
int wx_iacdwt_step_f32(float *v, const float *w1, const float *w2, long n, void *s)
{
    WX_REQUIRE(w1 && w2 && v, "null signal pointer");
    return wx_launch_iacdwt_step<float>(view1(v), cview1(w1), cview1(w2), n, batch1(), (cudaStream_t)s);
}

int wx_dwt_step2_f64(double *w1, double *w2, double *w3, double *w4, const double *v, long nr, long nc, const double *h, const double *g, int F, void *s) { return dwt_step_2d<double>(w1, w2, w3, w4, v, nr, nc, h, g, F, s); }
int wx_dwt_step2_f32(float *w1, float *w2, float *w3, float *w4, const float *v, long nr, long nc, const double *h, const double *g, int F, void *s) { return dwt_step_2d<float>(w1, w2, w3, w4, v, nr, nc, h, g, F, s); }
int wx_idwt_step2_f64(double *v, const double *w1, const double *w2, const double *w3, const double *w4, long nr, long nc, const double *h, const double *g, int F, void *s) { return idwt_step_2d<double>(v, w1, w2, w3, w4, nr, nc, h, g, F, s); }
int wx_idwt_step2_f32(float *v, const float *w1, const float *w2, const float *w3, const float *w4, long nr, long nc, const double *h, const double *g, int F, void *s) { return idwt_step_2d<float>(v, w1, w2, w3, w4, nr, nc, h, g, F, s); }
int wx_rdwt_step2_f64(int ac, double *w1, double *w2, double *w3, double *w4, const double *v, long nr, long nc, int d, const double *h, const double *g, int F, void *s) { return rdwt_step_2d<double>(ac, w1, w2, w3, w4, v, nr, nc, d, h, g, F, s); }
int wx_rdwt_step2_f32(int ac, float *w1, float *w2, float *w3, float *w4, const float *v, long nr, long nc, int d, const double *h, const double *g, int F, void *s) { return rdwt_step_2d<float>(ac, w1, w2, w3, w4, v, nr, nc, d, h, g, F, s); }
int wx_irdwt_step2_f64(int mode, double *v, const double *w1, const double *w2, const double *w3, const double *w4, long nr, long nc, int d, long sv, long sw, const double *h, const double *g, int F, void *s) { return irdwt_step_2d<double>(mode, v, w1, w2, w3, w4, nr, nc, d, sv, sw, h, g, F, s); }
int wx_irdwt_step2_f32(int mode, float *v, const float *w1, const float *w2, const float *w3, const float *w4, long nr, long nc, int d, long sv, long sw, const double *h, const double *g, int F, void *s) { return irdwt_step_2d<float>(mode, v, w1, w2, w3, w4, nr, nc, d, sv, sw, h, g, F, s); }

}  // extern "C"

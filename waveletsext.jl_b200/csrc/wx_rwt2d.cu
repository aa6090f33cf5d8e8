// wx_rwt2d.cu -- fused 2-D redundant (a-trous) step for the stationary and autocorrelation families:
// sdwt_step! 2-D swt/swt_one_level.jl:334-370 and acdwt_step! 2-D acwt/acwt_one_level.jl:240-276 (columns into temp, rows
// into w1..w4), batched over the nodes of one depth and the images of a chunk (swpd!/swpt!/sdwt! 2-D SWT.jl:133-158,
// 475-513, 870-902; ACWT.jl:133-157, 463-501, 761-793).
//
// One launch per depth instead of three: a CTA produces a tr x tc tile of each of the four children of one node from the
// (tr + (F-1)D) x (tc + (F-1)D) parent patch (periodic halo, D = 2^d); the column pass and the row pass both run in shared
// memory, so a node costs one read of the parent (+halo, served by L2) and one write of the four children -- the
// algorithmic traffic -- instead of 9 image sizes through a temp array.
//   stationary : w_L[i] = sum_j g[F-1-j] v[i-D+jD],  w_H[i] = sum_j h[j] v[i-jD]; the detail outputs are taken (F-2)D positions
//                ahead so that both filters read the same samples v[i-D .. i+(F-2)D]
//   autocorr.  : w[k] = sum_j f[j] v[k+(j-Lf/2)D]  (centred, both filters share the window)
#include "wx_steps.cuh"
#include "wx_2d.cuh"
#include <cstdlib>

namespace {

// both outputs of one position from the window w[q] = v[i + (q - q0) D], q = 0..F-1
template <typename T, int F, int AC>
__device__ __forceinline__ void rdots(const T *w, const Taps<T> &tp, T &lo, T &hi)
{
    T a, b;
    if (AC) {
        a = tp.g[0] * w[0]; b = tp.h[0] * w[0];
#pragma unroll
        for (int j = 1; j < F; ++j) { a = fma(tp.g[j], w[j], a); b = fma(tp.h[j], w[j], b); }
    } else {
        a = tp.g[F - 1] * w[0]; b = tp.h[0] * w[F - 1];
#pragma unroll
        for (int j = 1; j < F; ++j) { a = fma(tp.g[F - 1 - j], w[j], a); b = fma(tp.h[j], w[F - 1 - j], b); }
    }
    lo = a; hi = b;
}

// parents: v + node*vns + image*vis ; children: w1 + node*wns + image*wis + c*wq (c = 0..3).  All slices m x n column-major.
template <typename T, int F, int AC>
__global__ void __launch_bounds__(kT2) rdwt2d_tile_k(T *__restrict__ w1, long wns, long wis, long wq, const T *__restrict__ v, long vns, long vis,
                                                    int m, int n, int D, int tr, int tc, int PR, int PC, int tiles_r, int tiles_c, long nodes,
                                                    Taps<T> tp)
{
    extern __shared__ __align__(16) unsigned char wx_r2_smem[];
    T *P = reinterpret_cast<T *>(wx_r2_smem);            // (PR, PC) parent patch, column-major
    T *Tm = P + PR * PC;                                  // (2 tr, PC) column-pass output: rows [0,tr) scaling, [tr,2tr) detail
    const int tid = threadIdx.x;
    const int halo = (F - 1) * D;
    const int back = AC ? (F / 2) * D : D;                // the window of output i starts at i - back
    const int SH = AC ? 0 : (F - 2) * D;                  // detail outputs are produced SH positions ahead
    const bool fullr = PR == m && tr + halo > m, fullc = PC == n && tc + halo > n;      // patch = whole extent: index modulo
    const int ti = blockIdx.y % tiles_r, tk = blockIdx.y / tiles_r;          // grid.y = tile, grid.x = node + nodes * image
    const long node = blockIdx.x % nodes, k = blockIdx.x / nodes;
    const int r0 = ti * tr, c0 = tk * tc;
    const T *par = v + node * vns + k * vis;
    T *ch = w1 + node * wns + k * wis;

    // ---- parent patch ----
    {
        int rs = (r0 - back) % m; if (rs < 0) rs += m;
        int cs = (c0 - back) % n; if (cs < 0) cs += n;
        if (fullr) rs = 0;
        if (fullc) cs = 0;
        for (Walk2 w(tid, PR); w.hi < PC; w.next()) {
            int rr = rs + w.lo; while (rr >= m) rr -= m;
            int cc = cs + w.hi; while (cc >= n) cc -= n;
            P[w.hi * PR + w.lo] = par[(long)cc * m + rr];
        }
    }
    __syncthreads();
    // ---- column pass (along the rows): every patch column, tr output rows ----
    {
        int base = fullr ? ((r0 - back) % m + m) % m : 0;
        for (Walk2 w(tid, tr); w.hi < PC; w.next()) {
            const int il = w.lo, b = w.hi;
            const T *src = P + b * PR;
            T win[F];
            if (fullr) {
                int e = base + il; while (e >= m) e -= m;
                const int Dm = D % m;
#pragma unroll
                for (int q = 0; q < F; ++q) { win[q] = src[e]; e += Dm; if (e >= m) e -= m; }
            } else {
#pragma unroll
                for (int q = 0; q < F; ++q) win[q] = src[il + q * D];
            }
            T lo, hi;
            rdots<T, F, AC>(win, tp, lo, hi);
            Tm[b * (2 * tr) + il] = lo;
            Tm[b * (2 * tr) + tr + il] = hi;
        }
    }
    __syncthreads();
    // ---- row pass (along the columns) + store: 2 tr rows x tc output columns ----
    {
        const int R2 = 2 * tr;
        int base = fullc ? ((c0 - back) % n + n) % n : 0;
        for (Walk2 w(tid, R2); w.hi < tc; w.next()) {
            const int r = w.lo, kl = w.hi;
            const T *src = Tm + r;
            T win[F];
            if (fullc) {
                int e = base + kl; while (e >= n) e -= n;
                const int Dn = D % n;
#pragma unroll
                for (int q = 0; q < F; ++q) { win[q] = src[(long)e * R2]; e += Dn; if (e >= n) e -= n; }
            } else {
#pragma unroll
                for (int q = 0; q < F; ++q) win[q] = src[(kl + q * D) * R2];
            }
            T lo, hi;
            rdots<T, F, AC>(win, tp, lo, hi);
            const bool hpart = r >= tr;
            int row = r0 + (hpart ? r - tr + SH : r); row %= m;
            const int clo = c0 + kl;
            int chi = (c0 + kl + SH) % n;
            T *o = ch + (hpart ? 2 * wq : 0) + row;
            o[(long)clo * m] = lo;                         // w1 / w3
            o[wq + (long)chi * m] = hi;                    // w2 / w4
        }
    }
}

template <typename T, int F, int AC>
int launch_rdwt2d(T *w1, long wns, long wis, long wq, const T *v, long vns, long vis, long m, long n, long nodes, long Nc, int d,
                  const Taps<T> &t, cudaStream_t s, bool *handled)
{
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const long D = 1L << d, halo = (long)(F - 1) * D;
    // largest tile (<= 32, dividing the image) whose patch and column-pass buffer fit shared memory
    int tr = 0, tc = 0; long PR = 0, PC = 0; size_t smem = 0;
    for (int cap = 32; cap >= 4 && tr == 0; cap /= 2) {
        int a = 1, b = 1;
        for (int q = 1; q <= cap; ++q) { if (m % q == 0) a = q; if (n % q == 0) b = q; }
        const long pr = (a + halo > m) ? m : a + halo, pc = (b + halo > n) ? n : b + halo;
        const size_t need = ((size_t)pr * pc + (size_t)2 * a * pc) * sizeof(T);
        if (need <= dv.smem_optin) { tr = a; tc = b; PR = pr; PC = pc; smem = need; }
    }
    if (tr == 0 || m >= (1L << 30) || n >= (1L << 30) || D >= (1L << 30)) return WX_OK;
    // a halo much wider than the tile makes the column pass recompute (1 + halo/tile) times: beyond this the three-pass path
    // (9 image sizes of traffic, no redundant arithmetic) is faster -- measured on 256 x 256 images, see DESIGN.md
    static const char *env = getenv("WX_B200_RWT2D_MAXHALO");
    const long maxhalo = env ? atol(env) : 32;
    if (halo > maxhalo && (tr + halo <= m || tc + halo <= n)) return WX_OK;
    const long gy = (m / tr) * (n / tc), gx = nodes * Nc;
    if (gx >= (1L << 31) || gy > 65535) return WX_OK;
    auto kern = rdwt2d_tile_k<T, F, AC>;
    WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)gx, (unsigned)gy), kT2, smem, s>>>(w1, wns, wis, wq, v, vns, vis, (int)m, (int)n, (int)D, tr, tc, (int)PR, (int)PC,
                                                             (int)(m / tr), (int)(n / tc), nodes, t);
    WX_LAUNCHED();
    *handled = true;
    return WX_OK;
}

}  // namespace

// one depth of a 2-D redundant tree for `nodes` parents x Nc images.  The children must not alias the parents (the in-place
// swpt!/sdwt! layouts go through a copy of the parent, like the reference's).  *handled = false: not covered, nothing launched.
template <typename T>
int wx_rdwt2d_fused(int ac, T *w1, long wns, long wis, long wq, const T *v, long vns, long vis, long m, long n, long nodes, long Nc, int d,
                    const Taps<T> &t, cudaStream_t s, bool *handled)
{
    *handled = false;
    static const bool off = getenv("WX_B200_NO_FUSED_RWT2D") != nullptr;
    if (off || nodes < 1 || Nc < 1 || d > 28) return WX_OK;
#define WX_R2_CASE(FF, AA) case FF: return launch_rdwt2d<T, FF, AA>(w1, wns, wis, wq, v, vns, vis, m, n, nodes, Nc, d, t, s, handled);
    if (ac) {
        switch (t.F) { WX_R2_CASE(3, 1) WX_R2_CASE(7, 1) WX_R2_CASE(11, 1) WX_R2_CASE(15, 1) WX_R2_CASE(19, 1) WX_R2_CASE(23, 1) WX_R2_CASE(31, 1) WX_R2_CASE(39, 1) }
    } else {
        switch (t.F) { WX_R2_CASE(2, 0) WX_R2_CASE(4, 0) WX_R2_CASE(6, 0) WX_R2_CASE(8, 0) WX_R2_CASE(10, 0) WX_R2_CASE(12, 0) WX_R2_CASE(16, 0) WX_R2_CASE(20, 0) }
    }
#undef WX_R2_CASE
    return WX_OK;
}
template int wx_rdwt2d_fused<double>(int, double *, long, long, long, const double *, long, long, long, long, long, long, int, const Taps<double> &, cudaStream_t, bool *);
template int wx_rdwt2d_fused<float>(int, float *, long, long, long, const float *, long, long, long, long, long, long, int, const Taps<float> &, cudaStream_t, bool *);

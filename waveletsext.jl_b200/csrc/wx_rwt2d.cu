// wx_rwt2d.cu -- fused 2-D redundant (a-trous) step for the stationary and autocorrelation families:
// sdwt_step! 2-D swt/swt_one_level.jl:334-370 and acdwt_step! 2-D acwt/acwt_one_level.jl:240-276 (columns into temp, rows
// into w1..w4), batched over the nodes of one depth and the images of a chunk (swpd!/swpt!/sdwt! 2-D SWT.jl:133-158,
// 475-513, 870-902; ACWT.jl:133-157, 463-501, 761-793).
//
// One launch per depth instead of three: a CTA produces a tr x tc tile of each of the four children of one node from the
// (tr + (F-1)D) x (tc + F-1) parent patch (periodic halo, D = 2^d, columns of one a-trous coset); the column pass and the row pass both run in shared
// memory, so a node costs one read of the parent (+halo, served by L2) and one write of the four children -- the
// algorithmic traffic -- instead of 9 image sizes through a temp array.
//   stationary : w_L[i] = sum_j g[F-1-j] v[i-D+jD],  w_H[i] = sum_j h[j] v[i-jD]; the detail outputs are taken (F-2)D positions
//                ahead so that both filters read the same samples v[i-D .. i+(F-2)D]
//   autocorr.  : w[k] = sum_j f[j] v[k+(j-Lf/2)D]  (centred, both filters share the window)
#include "wx_steps.cuh"
#include "wx_2d.cuh"
#include <cstdlib>
#include <mutex>
#include <vector>

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

// outputs per thread of the sliding-window passes: windows of K + F - 1 samples; long filters take 2 (110 registers with 4 for the
// 15-tap autocorrelation filters in Float64: two resident CTAs and a slower kernel than one output per thread)
__host__ __device__ constexpr int wx_kr2(int F) { return F >= 12 ? 2 : 4; }

// parents: v + node*vns + image*vis ; children: w1 + node*wns + image*wis + c*wq (c = 0..3).  All slices m x n column-major.
// Rows (the contiguous dimension) are tiled as tr consecutive rows with a halo of (F-1) D rows; the COLUMNS of a tile belong to
// one coset (gamma mod D) of the a-trous lattice -- image column gamma + D j is coset column j, nc = n / D of them -- because the
// dilated filter never mixes cosets: in coset coordinates it is an ordinary stride-1 filter and the column halo is F-1 coset columns
// whatever the depth ((tc + (F-1) D)^2 patches made level 2 of a 256^2 swpd read 3.5 x its input).
template <typename T, int F, int AC>
__global__ void __launch_bounds__(kT2, 3) rdwt2d_tile_k(T *__restrict__ w1, long wns, long wis, long wq, const T *__restrict__ v, long vns, long vis,
                                                    int m, int n, int D, int tr, int tc, int PR, int PC, int tiles_r, int tiles_c, long nodes,
                                                    int LDP, int LDT, Taps<T> tp)
{
    constexpr int KR2 = wx_kr2(F);
    extern __shared__ __align__(16) unsigned char wx_r2_smem[];
    T *P = reinterpret_cast<T *>(wx_r2_smem);            // (PR, PC) parent patch, column-major, leading dimension LDP
    T *Tm = P + LDP * PC;                                 // (2 tr, PC) column-pass output: rows [0,tr) scaling, [tr,2tr) detail; ld LDT
    const int tid = threadIdx.x;
    const int nc = n / D;                                 // coset columns
    const int halo = (F - 1) * D;
    const int back = AC ? (F / 2) * D : D;                // rows: the window of output i starts at i - back
    const int SH = AC ? 0 : (F - 2) * D;                  // rows: detail outputs are produced SH positions ahead
    constexpr int backc = AC ? F / 2 : 1, SHc = AC ? 0 : F - 2;      // the same in coset columns
    const bool fullr = PR == m && tr + halo > m, fullc = PC == nc && tc + F - 1 > nc;      // patch = whole extent: index modulo
    // grid.y = (row tile, column tile, coset), grid.x = node + nodes * image
    const int ti = blockIdx.y % tiles_r, tq = blockIdx.y / tiles_r, tk = tq % tiles_c, gam = tq / tiles_c;
    const long node = blockIdx.x % nodes, k = blockIdx.x / nodes;
    const int r0 = ti * tr, c0 = tk * tc;
    const T *par = v + node * vns + k * vis;
    T *ch = w1 + node * wns + k * wis;

    // ---- parent patch ----
    {
        int rs = (r0 - back) % m; if (rs < 0) rs += m;
        int cs = (c0 - backc) % nc; if (cs < 0) cs += nc;
        if (fullr) rs = 0;
        if (fullc) cs = 0;
        for (Walk2 w(tid, PR); w.hi < PC; w.next()) {
            int rr = rs + w.lo; while (rr >= m) rr -= m;
            int cc = cs + w.hi; while (cc >= nc) cc -= nc;
            cp_async_elem<T>(P + w.hi * LDP + w.lo, par + (long)(gam + cc * D) * m + rr);      // asynchronous: every copy of the thread in flight
        }
        cp_async_wait_all();
    }
    __syncthreads();
    // ---- column pass (along the rows): every patch column, tr output rows ----
    if (!fullr && tr % (KR2 * D) == 0) {
        // a thread owns one patch column, one row coset rho and KR2 consecutive coset rows: a window of KR2 + F - 1 samples (stride D)
        // slides down them; lanes run across the columns (LDP, LDT odd: conflict free)
        const int per_col = tr / KR2;                     // (rho, group) items per column
        for (int t = tid; t < per_col * PC; t += kT2) {
            const int b = t % PC, u = t / PC, rho = u % D, g = u / D;
            const int il0 = rho + D * KR2 * g;
            const T *src = P + b * LDP + il0;
            T win[KR2 + F - 1];
#pragma unroll
            for (int q = 0; q < KR2 + F - 1; ++q) win[q] = src[q * D];
            T *dst = Tm + b * LDT + il0;
#pragma unroll
            for (int j = 0; j < KR2; ++j) {
                T lo, hi;
                rdots<T, F, AC>(&win[j], tp, lo, hi);
                dst[j * D] = lo;
                dst[tr + j * D] = hi;
            }
        }
    } else {
        int base = fullr ? ((r0 - back) % m + m) % m : 0;
        for (Walk2 w(tid, tr); w.hi < PC; w.next()) {
            const int il = w.lo, b = w.hi;
            const T *src = P + b * LDP;
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
            Tm[b * LDT + il] = lo;
            Tm[b * LDT + tr + il] = hi;
        }
    }
    __syncthreads();
    // ---- row pass (along the coset columns) + store: 2 tr rows x tc output columns ----
    if (!fullc && tc % KR2 == 0) {
        // a thread owns one row and KR2 consecutive coset columns: one window of KR2 + F - 1 samples slides along the row
        const int R2 = 2 * tr;
        for (Walk2 w(tid, R2); w.hi < tc / KR2; w.next()) {
            const int r = w.lo, kl0 = KR2 * w.hi;
            const T *src = Tm + r + kl0 * LDT;
            T win[KR2 + F - 1];
#pragma unroll
            for (int q = 0; q < KR2 + F - 1; ++q) win[q] = src[q * LDT];
            const bool hpart = r >= tr;
            int row = r0 + (hpart ? r - tr + SH : r); row %= m;
            T *o = ch + (hpart ? 2 * wq : 0) + row;
            int chc = (c0 + kl0 + SHc) % nc;
#pragma unroll
            for (int j = 0; j < KR2; ++j) {
                T lo, hi;
                rdots<T, F, AC>(&win[j], tp, lo, hi);
                o[(long)(gam + (c0 + kl0 + j) * D) * m] = lo;            // w1 / w3
                o[wq + (long)(gam + chc * D) * m] = hi;                   // w2 / w4
                if (++chc >= nc) chc -= nc;
            }
        }
    } else {
        const int R2 = 2 * tr;
        int base = fullc ? ((c0 - backc) % nc + nc) % nc : 0;
        for (Walk2 w(tid, R2); w.hi < tc; w.next()) {
            const int r = w.lo, kl = w.hi;
            const T *src = Tm + r;
            T win[F];
            if (fullc) {
                int e = (base + kl) % nc;
#pragma unroll
                for (int q = 0; q < F; ++q) { win[q] = src[(long)e * LDT]; if (++e >= nc) e -= nc; }
            } else {
#pragma unroll
                for (int q = 0; q < F; ++q) win[q] = src[(kl + q) * LDT];
            }
            T lo, hi;
            rdots<T, F, AC>(win, tp, lo, hi);
            const bool hpart = r >= tr;
            int row = r0 + (hpart ? r - tr + SH : r); row %= m;
            const int clo = gam + (c0 + kl) * D;
            const int chi = gam + ((c0 + kl + SHc) % nc) * D;
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
    constexpr int KR2 = wx_kr2(F);
    const long D = 1L << d, halo = (long)(F - 1) * D;
    if (m >= (1L << 30) || n >= (1L << 30) || D >= (1L << 30) || n % D != 0) return WX_OK;
    const long nc = n / D;
    // tile (tr rows dividing m, tc coset columns dividing nc) with the least shared-memory traffic per output -- patch loads plus the
    // column pass over the column halo -- among those that keep three CTAs per SM; larger shared-memory budgets only if none fits
    int tr = 0, tc = 0; long PR = 0, PC = 0; size_t smem = 0;
    // the search walks up to 16 K candidates: remember the answer per (shape, depth) -- one entry per template instantiation
    struct Choice { long m, n, D; int tr, tc; long PR, PC; size_t smem; };
    static std::mutex mtx;
    static std::vector<Choice> cache;
    {
        std::lock_guard<std::mutex> lk(mtx);
        for (const Choice &c : cache) if (c.m == m && c.n == n && c.D == D) { tr = c.tr; tc = c.tc; PR = c.PR; PC = c.PC; smem = c.smem; break; }
    }
    if (tr == 0)
    for (size_t budget : {(size_t)72 << 10, (size_t)110 << 10, dv.smem_optin}) {
        double best = 0;
        for (long a = 1; a <= 256 && a <= m; ++a) {
            if (m % a) continue;
            for (long b = 1; b <= 64 && b <= nc; ++b) {
                if (nc % b) continue;
                const long pr = (a + halo > m) ? m : a + halo, pc = (b + F - 1 > nc) ? nc : b + F - 1;
                const size_t need = ((size_t)(pr | 1) * pc + (size_t)(2 * a + 1) * pc) * sizeof(T);      // odd leading dimensions
                if (need > budget || need > dv.smem_optin) continue;
                const bool fr = pr == m && a + halo > m, fc = pc == nc && b + F - 1 > nc;
                const double lc = (!fr && a % (KR2 * D) == 0) ? (double)(KR2 + F - 1) / KR2 : (double)F;      // loads per column-pass output pair
                const double lr = (!fc && b % KR2 == 0) ? (double)(KR2 + F - 1) / KR2 : (double)F;            // loads per row-pass output pair
                const double cost = ((double)pr * pc + (lc + 2.0) * a * pc + 2.0 * lr * a * b) / ((double)a * b);
                if (tr == 0 || cost < best * 0.999 || (cost <= best * 1.001 && a * b > (long)tr * tc)) { best = cost; tr = (int)a; tc = (int)b; PR = pr; PC = pc; smem = need; }
            }
        }
        if (tr) break;
    }
    if (tr == 0) return WX_OK;
    {
        std::lock_guard<std::mutex> lk(mtx);
        bool have = false;
        for (const Choice &c : cache) if (c.m == m && c.n == n && c.D == D) have = true;
        if (!have && cache.size() < 256) cache.push_back(Choice{m, n, D, tr, tc, PR, PC, smem});
    }
    const long gy = (m / tr) * (nc / tc) * D, gx = nodes * Nc;
    if (gx >= (1L << 31) || gy > 65535) return WX_OK;
    auto kern = rdwt2d_tile_k<T, F, AC>;
    WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<dim3((unsigned)gx, (unsigned)gy), kT2, smem, s>>>(w1, wns, wis, wq, v, vns, vis, (int)m, (int)n, (int)D, tr, tc, (int)PR, (int)PC,
                                                             (int)(m / tr), (int)(nc / tc), nodes, (int)(PR | 1), 2 * tr + 1, t);
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

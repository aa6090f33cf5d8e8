// wx_rwt.cu -- batched redundant (undecimated) transforms: stationary (SWT.jl) and autocorrelation (ACWT.jl)
// families, forward and inverse, 1-D and 2-D, all three output shapes (dwt / wpt / wpd).
// General path: one launch of the batched step kernel per tree depth (all nodes of the depth x all signals).
// The bandwidth-critical 1-D swpd/acwpd shapes are overridden by the fused kernel in wx_rwpd_fused.cu.
//
// Reference: SWT.jl:109-158 (sdwt!), :259-358 (isdwt!), :439-513 (swpt!), :613-758 (iswpt!), :840-902 (swpd!),
//            :1035-1199 (iswpd!); ACWT.jl:109-157, 287-329, 427-501, 581-648, 733-793, 917-1000;
//            batch loops swt/swt_all.jl, acwt/acwt_all.jl.
#include "wx_steps.cuh"
#include <vector>
#include <cstdlib>

template <typename T>
int wx_rwpd1d_fused(int ac, int wpt, T *xw, const T *x, long n, int L, long N, const Taps<T> &t, cudaStream_t s, int *done);
template <typename T>
int wx_rdwt1d_fused(int ac, T *xw, const T *x, long n, int L, long N, const Taps<T> &t, cudaStream_t s, int *done);
template <typename T> int wx_iac_tree_sum(T *x, const T *xw, long n, long ncols, long c0, int L, long N, cudaStream_t s);
template <typename T> int wx_iac_chain_sum(T *x, const T *xw, long n, int L, long N, cudaStream_t s);
template <typename T> int wx_isdwt_shift_chain(T *x, const T *xw, long n, int L, long N, const long *sd, const Taps<T> &t, cudaStream_t s, bool *handled);
// fused 2-D a-trous step (wx_rwt2d.cu)
template <typename T>
int wx_rdwt2d_fused(int ac, T *w1, long wns, long wis, long wq, const T *v, long vns, long vis, long m, long n, long nodes, long Nc, int d,
                    const Taps<T> &t, cudaStream_t s, bool *handled);
// fused average-based stationary inverses (wx_irwpd_fused.cu)
template <typename T> int wx_irwpd_avg_fused(T *x, const T *xw, long in_sig, long in_col0, long n, int Lt, long N, const Taps<T> &t, cudaStream_t s, bool *handled);
template <typename T> int wx_irdwt_chain_depth(const T *x, const T *xw, long n, int L, long N, const Taps<T> &t, int *dhi);
template <typename T> int wx_irdwt_chain_run(T *x, const T *xw, long n, int L, int dhi, long N, const Taps<T> &t, cudaStream_t s);

namespace {

static long pow4(int d) { return 1L << (2 * d); }
static long quad_first(int d) { return (pow4(d) - 1) / 3 + 1; }      // first heap index of depth d
static long scratch_budget_elems(size_t elt) { return (long)(((size_t)3 << 29) / elt); }   // 1.5 GiB

// ---- 2-D batched step: parents v, children w1..w4 given as (node stride, image stride) views -----------
// forward: columns into temp, rows into the children  (swt/swt_one_level.jl:352-368, acwt/acwt_one_level.jl:258-274)
template <typename T>
int fwd2d(int ac, T *w1, long wns, long wis, long wq /*slice distance between w1..w4*/, const T *v, long vns, long vis, T *temp, long m, long n,
          long nodes, long Nc, int d, const Taps<T> &t, cudaStream_t s)
{
    const long img = m * n;
    {   // fused tile kernel: one launch per depth.  When the children overwrite their parents (swpt! / sdwt! layouts) the
        // parents are first copied to the scratch, like the reference's `copy(xw[:,:,j])` (SWT.jl:150,498).
        const T *v_hi = v + (nodes - 1) * vns + img, *w_hi = w1 + (nodes - 1) * wns + 3 * wq + img;
        const bool disjoint = (vis == wis || Nc == 1) && (v_hi <= w1 || w_hi <= v);
        bool handled = false;
        int rc0 = WX_OK;
        static const bool off = getenv("WX_B200_NO_FUSED_RWT2D") != nullptr;
        if (off) { /* per-pass path below */ }
        else if (disjoint) rc0 = wx_rdwt2d_fused<T>(ac, w1, wns, wis, wq, v, vns, vis, m, n, nodes, Nc, d, t, s, &handled);
        else {
            rc0 = wx_launch_copy<T>(View<T>{temp, 1, img, img * nodes, 0}, View<const T>{v, 1, vns, vis, 0}, img, Batch{nodes, Nc, 1, false}, s);
            if (!rc0) rc0 = wx_rdwt2d_fused<T>(ac, w1, wns, wis, wq, temp, img, img * nodes, m, n, nodes, Nc, d, t, s, &handled);
        }
        if (rc0 || handled) return rc0;
    }
    // temp (m, n, 2, nodes, Nc)
    View<T> t1{temp, 1, m, 2 * img, 2 * img * nodes}, t2{temp + img, 1, m, 2 * img, 2 * img * nodes};
    int rc = wx_launch_rdwt_step<T>(ac, t1, t2, View<const T>{v, 1, m, vns, vis}, m, d, Batch{n, nodes, Nc, false}, t, s);
    if (rc) return rc;
    rc = wx_launch_rdwt_step<T>(ac, View<T>{w1, m, 1, wns, wis}, View<T>{w1 + wq, m, 1, wns, wis},
                                View<const T>{temp, m, 1, 2 * img, 2 * img * nodes}, n, d, Batch{m, nodes, Nc, true}, t, s);
    if (rc) return rc;
    return wx_launch_rdwt_step<T>(ac, View<T>{w1 + 2 * wq, m, 1, wns, wis}, View<T>{w1 + 3 * wq, m, 1, wns, wis},
                                  View<const T>{temp + img, m, 1, 2 * img, 2 * img * nodes}, n, d, Batch{m, nodes, Nc, true}, t, s);
}

// a child operand of the batched 2-D inverse: base pointer, node stride, image stride
template <typename T>
struct Child {
    const T *p;
    long ns, is;
};

// inverse: rows into temp, columns into the parent (swt/swt_one_level.jl:395-469, acwt/acwt_one_level.jl:288-322)
// imode 0 average, 1 shift, 2 autocorrelation
template <typename T>
int inv2d(int imode, T *v, long vns, long vis, Child<T> c1, Child<T> c2, Child<T> c3, Child<T> c4, T *temp, long m, long n,
          long nodes, long Nc, int d, long sv, long sw, const Taps<T> &t, cudaStream_t s)
{
    const long img = m * n;
    auto step = [&](View<T> vo, View<const T> a, View<const T> b, long len, Batch bt) -> int {
        if (imode == 2) return wx_launch_iacdwt_step<T>(vo, a, b, len, bt, s);
        if (imode == 1) return wx_launch_isdwt_shift<T>(vo, a, b, len, d, sv, sw, 0, bt, t, s);
        return wx_launch_isdwt_avg<T>(vo, a, b, len, d, bt, t, s);
    };
    if (imode == 1) WX_CUDA(cudaMemsetAsync(temp, 0, (size_t)2 * img * nodes * Nc * sizeof(T), s));
    int rc = step(View<T>{temp, m, 1, 2 * img, 2 * img * nodes}, View<const T>{c1.p, m, 1, c1.ns, c1.is}, View<const T>{c2.p, m, 1, c2.ns, c2.is}, n,
                  Batch{m, nodes, Nc, true});
    if (rc) return rc;
    rc = step(View<T>{temp + img, m, 1, 2 * img, 2 * img * nodes}, View<const T>{c3.p, m, 1, c3.ns, c3.is}, View<const T>{c4.p, m, 1, c4.ns, c4.is}, n,
              Batch{m, nodes, Nc, true});
    if (rc) return rc;
    return step(View<T>{v, 1, m, vns, vis}, View<const T>{temp, 1, m, 2 * img, 2 * img * nodes},
                View<const T>{temp + img, 1, m, 2 * img, 2 * img * nodes}, m, Batch{n, nodes, Nc, false});
}

// =====================================================================================================
// forward
// =====================================================================================================
template <typename T>
int rwt_fwd_1d(int ac, int mode, T *xw, const T *x, long n, int L, long N, const Taps<T> &t, cudaStream_t s)
{
    int rc;
    if (mode == WX_MODE_WPD) {
        int done = 0;                                        // leading levels already produced by the fused kernel
        rc = wx_rwpd1d_fused<T>(ac, 0, xw, x, n, L, N, t, s, &done);
        if (rc || done == L) return rc;
        const long ncols = (1L << (L + 1)) - 1, str = n * ncols;
        if (done == 0) rc = wx_launch_copy<T>(View<T>{xw, 1, str, 0, 0}, View<const T>{x, 1, n, 0, 0}, n, Batch{N, 1, 1, false}, s);
        for (int d = done; d < L && !rc; ++d) {
            const long nd = 1L << d;
            T *w1 = xw + ((1L << (d + 1)) - 1) * n;
            rc = wx_launch_rdwt_step<T>(ac, View<T>{w1, 1, 2 * n, str, 0}, View<T>{w1 + n, 1, 2 * n, str, 0},
                                        View<const T>{xw + (nd - 1) * n, 1, n, str, 0}, n, d, Batch{nd, N, 1, false}, t, s);
        }
        return rc;
    }
    if (mode == WX_MODE_WPT) {
        // swpt! SWT.jl:454-470 : in place, parent column is copied before its children overwrite it
        int done = 0;
        rc = wx_rwpd1d_fused<T>(ac, 1, xw, x, n, L, N, t, s, &done);
        if (rc || done == L) return rc;
        const long ncols = 1L << L, str = n * ncols;
        rc = wx_launch_copy<T>(View<T>{xw, 1, str, 0, 0}, View<const T>{x, 1, n, 0, 0}, n, Batch{N, 1, 1, false}, s);
        if (rc) return rc;
        T *tmp; rc = wx_scratch(&tmp, (size_t)n * (ncols / 2) * N, s); if (rc) return rc;
        for (int d = 0; d < L && !rc; ++d) {
            const long nd = 1L << d, np = ncols / nd;
            rc = wx_launch_copy<T>(View<T>{tmp, 1, n, n * nd, 0}, View<const T>{xw, 1, np * n, str, 0}, n, Batch{nd, N, 1, false}, s);
            if (!rc) rc = wx_launch_rdwt_step<T>(ac, View<T>{xw, 1, np * n, str, 0}, View<T>{xw + (np / 2) * n, 1, np * n, str, 0},
                                                 View<const T>{tmp, 1, n, n * nd, 0}, n, d, Batch{nd, N, 1, false}, t, s);
        }
        int rc2 = wx_scratch_free(tmp, s);
        return rc ? rc : rc2;
    }
    // sdwt! SWT.jl:120-129 : xw(n, L+1, N); column L = x; depth d: parent = copy(col L-d) -> cols L-d-1 (scaling), L-d (detail)
    const long str = n * (L + 1);
    int done = 0;                                            // leading levels produced by the fused chain kernel
    rc = wx_rdwt1d_fused<T>(ac, xw, x, n, L, N, t, s, &done);
    if (rc || done == L) return rc;
    if (done == 0) {
        rc = wx_launch_copy<T>(View<T>{xw + (long)L * n, 1, str, 0, 0}, View<const T>{x, 1, n, 0, 0}, n, Batch{N, 1, 1, false}, s);
        if (rc) return rc;
    }
    T *tmp; rc = wx_scratch(&tmp, (size_t)n * N, s); if (rc) return rc;
    for (int d = done; d < L && !rc; ++d) {
        T *par = xw + (long)(L - d) * n;
        rc = wx_launch_copy<T>(View<T>{tmp, 1, n, 0, 0}, View<const T>{par, 1, str, 0, 0}, n, Batch{N, 1, 1, false}, s);
        if (!rc) rc = wx_launch_rdwt_step<T>(ac, View<T>{par - n, 1, str, 0, 0}, View<T>{par, 1, str, 0, 0}, View<const T>{tmp, 1, n, 0, 0}, n, d,
                                             Batch{N, 1, 1, false}, t, s);
    }
    int rc2 = wx_scratch_free(tmp, s);
    return rc ? rc : rc2;
}

template <typename T>
int rwt_fwd_2d(int ac, int mode, T *xw, const T *x, long m, long n, int L, long N, const Taps<T> &t, cudaStream_t s)
{
    const long img = m * n;
    const long nsl = mode == WX_MODE_WPD ? (pow4(L + 1) - 1) / 3 : (mode == WX_MODE_WPT ? pow4(L) : 3L * L + 1);
    const long xs = img * nsl;
    const long root = (mode == WX_MODE_DWT) ? 3L * L : 0;
    int rc = wx_launch_copy<T>(View<T>{xw + root * img, 1, xs, 0, 0}, View<const T>{x, 1, img, 0, 0}, img, Batch{N, 1, 1, false}, s);
    if (rc) return rc;
    const long maxnodes = (mode == WX_MODE_DWT) ? 1 : pow4(L - 1);
    long Nc = scratch_budget_elems(sizeof(T)) / (2 * img * maxnodes);
    if (Nc < 1) Nc = 1;
    if (Nc > N) Nc = N;
    T *temp; rc = wx_scratch(&temp, (size_t)2 * img * maxnodes * Nc, s); if (rc) return rc;
    for (long k0 = 0; k0 < N && !rc; k0 += Nc) {
        const long nk = (N - k0 < Nc) ? N - k0 : Nc;
        T *xk = xw + k0 * xs;
        for (int d = 0; d < L && !rc; ++d) {
            if (mode == WX_MODE_WPD) {
                // swpd! 2-D SWT.jl:886-901 : node i -> children 4i-2..4i+1 (slices 4i-3..4i 0-based)
                const long f = quad_first(d), nd = pow4(d);
                rc = fwd2d<T>(ac, xk + (4 * f - 3) * img, 4 * img, xs, img, xk + (f - 1) * img, img, xs, temp, m, n, nd, nk, d, t, s);
            } else if (mode == WX_MODE_WPT) {
                // swpt! 2-D SWT.jl:493-511 : parent slice j1 = 4b*nc is also the first child; children nc slices apart
                const long nd = pow4(d), np = pow4(L) / nd, ncl = np / 4;
                rc = fwd2d<T>(ac, xk, np * img, xs, ncl * img, xk, np * img, xs, temp, m, n, nd, nk, d, t, s);
            } else {
                // sdwt! 2-D SWT.jl:149-156 : parent slice 3(L-d) (0-based) -> slices 3(L-d)-3 .. 3(L-d)
                const long b = 3L * (L - d);
                rc = fwd2d<T>(ac, xk + (b - 3) * img, 0, xs, img, xk + b * img, 0, xs, temp, m, n, 1, nk, d, t, s);
            }
        }
    }
    int rc2 = wx_scratch_free(temp, s);
    return rc ? rc : rc2;
}

template <typename T>
int rwt_impl(int ac, int mode, T *xw, const T *x, long m, long n, int L, long N, const double *h, const double *g, int F, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    WX_REQUIRE(mode >= 0 && mode <= 2, "bad mode %d", mode);
    WX_REQUIRE(n >= 1 && m >= 0 && N >= 0, "bad sizes");
    const int Lmax = m > 0 ? (wx_maxlevels(m) < wx_maxlevels(n) ? wx_maxlevels(m) : wx_maxlevels(n)) : wx_maxlevels(n);
    // reference: ArgumentError("Too many transform levels") / ("L must be >= 1")   SWT.jl:114-116
    WX_REQUIRE(L <= Lmax, "ArgumentError: Too many transform levels (length(x) < 2^L)");
    WX_REQUIRE(L >= 1, "ArgumentError: L must be >= 1");
    if (N == 0) return WX_OK;
    WX_REQUIRE(xw && x, "null signal pointer");
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    return m > 0 ? rwt_fwd_2d<T>(ac, mode, xw, x, m, n, L, N, t, s) : rwt_fwd_1d<T>(ac, mode, xw, x, n, L, N, t, s);
}

// =====================================================================================================
// inverse
// =====================================================================================================
static void depth_shifts(std::vector<long> &sd, long sm, int L)
{
    // main2depthshift Utils.jl:297-305
    sd.assign((size_t)L + 1, 0);
    long acc = 0;
    for (int d = 0; d < L; ++d) { acc += ((sm >> d) & 1L) << d; sd[(size_t)d + 1] = acc; }
}

template <typename T>
int istep1d(int imode, View<T> v, View<const T> w1, View<const T> w2, long n, int d, long sv, long sw, Batch b, const Taps<T> &t, cudaStream_t s)
{
    if (imode == 2) return wx_launch_iacdwt_step<T>(v, w1, w2, n, b, s);
    if (imode == 1) return wx_launch_isdwt_shift<T>(v, w1, w2, n, d, sv, sw, 0, b, t, s);
    return wx_launch_isdwt_avg<T>(v, w1, w2, n, d, b, t, s);
}

template <typename T>
int irwt_1d(int imode, int mode, T *x, const T *xw, long n, long ncols, int L, long N, const unsigned char *tree, long ntree,
            const std::vector<long> &sd, const Taps<T> &t, cudaStream_t s)
{
    int rc = WX_OK;
    const long str = n * ncols;
    auto SV = [&](int d) { return imode == 1 ? sd[(size_t)d] : 0L; };
    auto SW = [&](int d) { return imode == 1 ? sd[(size_t)d + 1] : 0L; };
    if (imode == 2 && mode == WX_MODE_DWT) return wx_iac_chain_sum<T>(x, xw, n, L, N, s);        // plain sums: one pass over the table
    if (imode == 2 && mode == WX_MODE_WPT && L < 32) return wx_iac_tree_sum<T>(x, xw, n, ncols, 0, L, N, s);
    if (mode == WX_MODE_DWT) {
        // isdwt! SWT.jl:270-282, 311-328 ; iacdwt! ACWT.jl:292-303 : x = col 0; for d = L-1..0: x = step(copy(x), col L-d)
        int dhi = 0;                                         // levels d < dhi run in the fused chain kernel (average based only)
        if (imode == 0) { rc = wx_irdwt_chain_depth<T>(x, xw, n, L, N, t, &dhi); if (rc) return rc; }
        if (dhi == L) return wx_irdwt_chain_run<T>(x, xw, n, L, dhi, N, t, s);
        if (imode == 1) {                                    // shift based: the whole chain in one launch when a signal fits shared memory
            bool handled = false;
            rc = wx_isdwt_shift_chain<T>(x, xw, n, L, N, sd.data(), t, s, &handled);
            if (rc || handled) return rc;
        }
        T *tmp; rc = wx_scratch(&tmp, (size_t)n * N, s); if (rc) return rc;
        rc = wx_launch_copy<T>(View<T>{x, 1, n, 0, 0}, View<const T>{xw, 1, str, 0, 0}, n, Batch{N, 1, 1, false}, s);
        for (int d = L - 1; d >= dhi && !rc; --d) {
            WX_CUDA(cudaMemcpyAsync(tmp, x, (size_t)n * N * sizeof(T), cudaMemcpyDeviceToDevice, s));
            rc = istep1d<T>(imode, View<T>{x, 1, n, 0, 0}, View<const T>{tmp, 1, n, 0, 0}, View<const T>{xw + (long)(L - d) * n, 1, str, 0, 0}, n, d,
                            SV(d), SW(d), Batch{N, 1, 1, false}, t, s);
        }
        if (!rc && dhi > 0) rc = wx_irdwt_chain_run<T>(x, xw, n, L, dhi, N, t, s);
        int rc2 = wx_scratch_free(tmp, s);
        return rc ? rc : rc2;
    }
    if (mode == WX_MODE_WPT && imode == 0) {
        bool handled = false;
        rc = wx_irwpd_avg_fused<T>(x, xw, str, 0, n, L, N, t, s, &handled);
        if (rc || handled) return rc;
    }
    if (mode == WX_MODE_WPT) {
        // iswpt! SWT.jl:628-645, 700-715 ; iacwpt! ACWT.jl:594-607.  Depth-d nodes are kept compacted in a workspace
        // (n, 2^d, N): W_d[b] = step(W_{d+1}[2b], W_{d+1}[2b+1]); the deepest level reads xw itself.
        T *wa = nullptr, *wb = nullptr;
        if (L >= 2) { rc = wx_scratch(&wa, (size_t)n * (ncols / 2) * N, s); if (rc) return rc; }
        if (L >= 3) { rc = wx_scratch(&wb, (size_t)n * (ncols / 4) * N, s); if (rc) return rc; }
        // shift based: a step writes only the sv coset of its parent and the next step reads exactly that coset of its children
        // (sw of depth d = sv of depth d+1, main2depthshift), and depth 0 writes every position of x, so nothing off the cosets is
        // ever read.  (Round 1 zero-filled both workspaces first: 6.4 GB of memset for 2048 x 2048, L = 8 -- half the run time.)
        const T *src = xw; long srcstr = str;
        T *bufs[2] = {wa, wb};
        int which = 0;
        for (int d = L - 1; d >= 0 && !rc; --d) {
            const long nd = 1L << d;
            T *dst = (d == 0) ? x : bufs[which];
            const long dststr = n * nd;
            rc = istep1d<T>(imode, View<T>{dst, 1, n, dststr, 0}, View<const T>{src, 1, 2 * n, srcstr, 0}, View<const T>{src + n, 1, 2 * n, srcstr, 0}, n, d,
                            SV(d), SW(d), Batch{nd, N, 1, false}, t, s);
            src = dst; srcstr = dststr; which ^= 1;
        }
        int rc2 = wx_scratch_free(wa, s), rc3 = wx_scratch_free(wb, s);
        return rc ? rc : (rc2 ? rc2 : rc3);
    }
    // iswpd! by tree SWT.jl:1078-1093, 1143-1156 ; iacwpd! ACWT.jl:954-969.
    // workspace W holds nodes 1..2^Lx-1 (initialised from xw); children deeper than that are read from xw.
    const int Lx = wx_ilog2l(ncols + 1) - 1;                 // ncols = 2^(Lx+1)-1
    long lastsplit = 0;
    for (long i = ntree; i >= 1; --i) if (tree[i - 1]) { lastsplit = i; break; }
    if (imode == 2 && lastsplit > 0) {                       // autocorrelation + complete tree of depth Lt: pairwise sum of the depth-Lt columns
        const int Lt = wx_ilog2l(lastsplit) + 1;
        bool full = lastsplit == (1L << Lt) - 1 && Lt < 32 && (1L << (Lt + 1)) - 1 <= ncols;
        for (long i = 1; full && i <= lastsplit; ++i) full = tree[i - 1] != 0;
        if (full) return wx_iac_tree_sum<T>(x, xw, n, ncols, (1L << Lt) - 1, Lt, N, s);
    }
    if (imode == 0 && lastsplit > 0) {                       // average based + complete tree of depth Lt: fused tree reduction
        const int Lt = wx_ilog2l(lastsplit) + 1;
        bool full = lastsplit == (1L << Lt) - 1 && Lt < 31 && (1L << (Lt + 1)) - 1 <= ncols;
        for (long i = 1; full && i <= lastsplit; ++i) full = tree[i - 1] != 0;
        if (full) {
            bool handled = false;
            rc = wx_irwpd_avg_fused<T>(x, xw, str, (1L << Lt) - 1, n, Lt, N, t, s, &handled);
            if (rc || handled) return rc;
        }
    }
    if (lastsplit == 0) {                                    // root is a leaf: x = column 0
        return wx_launch_copy<T>(View<T>{x, 1, n, 0, 0}, View<const T>{xw, 1, str, 0, 0}, n, Batch{N, 1, 1, false}, s);
    }
    WX_REQUIRE(2 * lastsplit + 1 <= ncols, "tree is deeper than the packet table (node %ld has no children in xw)", lastsplit);
    const long wcols = (1L << Lx) - 1;                       // internal nodes
    T *W = nullptr;
    const long wstr = n * wcols;
    if (wcols > 1) {
        rc = wx_scratch(&W, (size_t)wstr * N, s); if (rc) return rc;
        rc = wx_launch_copy<T>(View<T>{W, 1, wstr, 0, 0}, View<const T>{xw, 1, str, 0, 0}, wstr, Batch{N, 1, 1, false}, s);
    }
    if (imode == 1) WX_CUDA(cudaMemsetAsync(x, 0, (size_t)n * N * sizeof(T), s));
    for (int d = wx_ilog2l(lastsplit); d >= 0 && !rc; --d) {
        const long first = 1L << d, last = (1L << (d + 1)) - 1;
        long i = first;
        while (i <= last && !rc) {
            if (!(i <= ntree && tree[i - 1])) { ++i; continue; }
            long j = i;
            while (j + 1 <= last && j + 1 <= ntree && tree[j]) ++j;      // run of split nodes i..j
            const long run = j - i + 1;
            const bool deep = (2 * i > wcols);                          // children live only in xw
            const T *cb = deep ? xw : W;
            const long cstr = deep ? str : wstr;
            T *vb = (i == 1) ? x : W + (i - 1) * n;
            const long vstr = (i == 1) ? n : wstr;
            rc = istep1d<T>(imode, View<T>{vb, 1, n, vstr, 0}, View<const T>{cb + (2 * i - 1) * n, 1, 2 * n, cstr, 0},
                            View<const T>{cb + (2 * i) * n, 1, 2 * n, cstr, 0}, n, d, SV(d), SW(d), Batch{run, N, 1, false}, t, s);
            i = j + 1;
        }
    }
    int rc2 = wx_scratch_free(W, s);
    return rc ? rc : rc2;
}

template <typename T>
int irwt_2d(int imode, int mode, T *x, const T *xw, long m, long n, long nsl, int L, long N, const unsigned char *tree, long ntree,
            const std::vector<long> &sd, const Taps<T> &t, cudaStream_t s)
{
    int rc = WX_OK;
    const long img = m * n, xs = img * nsl;
    auto SV = [&](int d) { return imode == 1 ? sd[(size_t)d] : 0L; };
    auto SW = [&](int d) { return imode == 1 ? sd[(size_t)d + 1] : 0L; };
    if (imode == 1) WX_CUDA(cudaMemsetAsync(x, 0, (size_t)img * N * sizeof(T), s));
    // Autocorrelation inverses are position-wise sums (iacdwt_step! 2-D acwt/acwt_one_level.jl:288-322: rows (w1+w2)/sqrt2, (w3+w4)/sqrt2, then
    // columns): the four children of a quad node are four consecutive quarter-ranges of slices, so the quad-tree reduction of depth L IS the
    // binary pairwise reduction of depth 2L over the slices in their natural order -- one pass over the table instead of three launches
    // and 9 image sizes per depth.
    if (imode == 2 && mode == WX_MODE_WPT && 2 * L < 32) return wx_iac_tree_sum<T>(x, xw, img, nsl, 0, 2 * L, N, s);
    if (imode == 2 && mode == WX_MODE_WPD) {
        long last = 0;
        for (long i = ntree; i >= 1; --i) if (tree[i - 1]) { last = i; break; }
        if (last > 0) {
            const int Lt = wx_quaddepthl(last) + 1;                                  // depth of the leaves if the tree is complete
            const long inner = (pow4(Lt) - 1) / 3;                                   // nodes above depth Lt
            bool full = last == inner && 2 * Lt < 32 && inner + pow4(Lt) <= nsl;
            for (long i = 1; full && i <= last; ++i) full = tree[i - 1] != 0;
            if (full) return wx_iac_tree_sum<T>(x, xw, img, nsl, inner, 2 * Lt, N, s);
        }
    }
    if (mode == WX_MODE_DWT) {
        // isdwt! 2-D SWT.jl:296-308, 345-356 ; iacdwt! 2-D ACWT.jl:317-327
        T *tmp, *temp;
        rc = wx_scratch(&tmp, (size_t)img * N, s); if (rc) return rc;
        rc = wx_scratch(&temp, (size_t)2 * img * N, s); if (rc) return rc;
        rc = wx_launch_copy<T>(View<T>{x, 1, img, 0, 0}, View<const T>{xw, 1, xs, 0, 0}, img, Batch{N, 1, 1, false}, s);
        for (int d = L - 1; d >= 0 && !rc; --d) {
            const long b = 3L * (L - d);
            WX_CUDA(cudaMemcpyAsync(tmp, x, (size_t)img * N * sizeof(T), cudaMemcpyDeviceToDevice, s));
            rc = inv2d<T>(imode, x, 0, img, Child<T>{tmp, 0, img}, Child<T>{xw + (b - 2) * img, 0, xs}, Child<T>{xw + (b - 1) * img, 0, xs},
                          Child<T>{xw + b * img, 0, xs}, temp, m, n, 1, N, d, SV(d), SW(d), t, s);
        }
        int rc2 = wx_scratch_free(tmp, s), rc3 = wx_scratch_free(temp, s);
        return rc ? rc : (rc2 ? rc2 : rc3);
    }
    if (mode == WX_MODE_WPT) {
        // iswpt! 2-D SWT.jl:663-684, 733-757 ; iacwpt! 2-D ACWT.jl:627-647.  compacted workspaces per depth.
        const long tot = pow4(L);
        long Nc = scratch_budget_elems(sizeof(T)) / (img * tot);
        if (Nc < 1) Nc = 1;
        if (Nc > N) Nc = N;
        T *wa = nullptr, *wb = nullptr, *temp = nullptr;
        rc = wx_scratch(&temp, (size_t)2 * img * (tot / 4) * Nc, s); if (rc) return rc;
        if (L >= 2) { rc = wx_scratch(&wa, (size_t)img * (tot / 4) * Nc, s); if (rc) return rc; }
        if (L >= 3) { rc = wx_scratch(&wb, (size_t)img * (tot / 16) * Nc, s); if (rc) return rc; }
        if (imode == 1) {
            if (wa) WX_CUDA(cudaMemsetAsync(wa, 0, (size_t)img * (tot / 4) * Nc * sizeof(T), s));
            if (wb) WX_CUDA(cudaMemsetAsync(wb, 0, (size_t)img * (tot / 16) * Nc * sizeof(T), s));
        }
        for (long k0 = 0; k0 < N && !rc; k0 += Nc) {
            const long nk = (N - k0 < Nc) ? N - k0 : Nc;
            const T *src = xw + k0 * xs; long sis = xs;
            T *bufs[2] = {wa, wb};
            int which = 0;
            for (int d = L - 1; d >= 0 && !rc; --d) {
                const long nd = pow4(d);
                T *dst = (d == 0) ? x + k0 * img : bufs[which];
                const long dis = img * nd;
                rc = inv2d<T>(imode, dst, img, dis, Child<T>{src, 4 * img, sis}, Child<T>{src + img, 4 * img, sis}, Child<T>{src + 2 * img, 4 * img, sis},
                              Child<T>{src + 3 * img, 4 * img, sis}, temp, m, n, nd, nk, d, SV(d), SW(d), t, s);
                src = dst; sis = dis; which ^= 1;
            }
        }
        int rc2 = wx_scratch_free(wa, s), rc3 = wx_scratch_free(wb, s), rc4 = wx_scratch_free(temp, s);
        return rc ? rc : (rc2 ? rc2 : (rc3 ? rc3 : rc4));
    }
    // iswpd! 2-D by tree SWT.jl:1112-1128, 1178-1198 ; iacwpd! 2-D ACWT.jl:982-999
    const int Lx = wx_quaddepthl(nsl);
    long lastsplit = 0;
    for (long i = ntree; i >= 1; --i) if (tree[i - 1]) { lastsplit = i; break; }
    if (lastsplit == 0)
        return wx_launch_copy<T>(View<T>{x, 1, img, 0, 0}, View<const T>{xw, 1, xs, 0, 0}, img, Batch{N, 1, 1, false}, s);
    WX_REQUIRE(4 * lastsplit + 1 <= nsl, "tree is deeper than the packet table (node %ld has no children in xw)", lastsplit);
    const long wsl = (pow4(Lx) - 1) / 3;                     // internal nodes
    long Nc = scratch_budget_elems(sizeof(T)) / (img * (wsl > 0 ? wsl : 1));
    if (Nc < 1) Nc = 1;
    if (Nc > N) Nc = N;
    T *W = nullptr, *temp = nullptr;
    const long wis = img * wsl;
    rc = wx_scratch(&temp, (size_t)2 * img * pow4(Lx > 0 ? Lx - 1 : 0) * Nc, s); if (rc) return rc;
    if (wsl > 1) { rc = wx_scratch(&W, (size_t)wis * Nc, s); if (rc) return rc; }
    for (long k0 = 0; k0 < N && !rc; k0 += Nc) {
        const long nk = (N - k0 < Nc) ? N - k0 : Nc;
        const T *xk = xw + k0 * xs;
        if (W) rc = wx_launch_copy<T>(View<T>{W, 1, wis, 0, 0}, View<const T>{xk, 1, xs, 0, 0}, wis, Batch{nk, 1, 1, false}, s);
        for (int d = wx_quaddepthl(lastsplit); d >= 0 && !rc; --d) {
            const long first = quad_first(d), last = quad_first(d + 1) - 1;
            long i = first;
            while (i <= last && !rc) {
                if (!(i <= ntree && tree[i - 1])) { ++i; continue; }
                long j = i;
                while (j + 1 <= last && j + 1 <= ntree && tree[j]) ++j;
                const long run = j - i + 1;
                const bool deep = (4 * i - 2 > wsl);
                const T *cb = deep ? xk : W;
                const long cis = deep ? xs : wis;
                T *vb = (i == 1) ? x + k0 * img : W + (i - 1) * img;
                const long vis = (i == 1) ? img : wis;
                const T *c1 = cb + (4 * i - 3) * img;
                rc = inv2d<T>(imode, vb, img, vis, Child<T>{c1, 4 * img, cis}, Child<T>{c1 + img, 4 * img, cis}, Child<T>{c1 + 2 * img, 4 * img, cis},
                              Child<T>{c1 + 3 * img, 4 * img, cis}, temp, m, n, run, nk, d, SV(d), SW(d), t, s);
                i = j + 1;
            }
        }
    }
    int rc2 = wx_scratch_free(W, s), rc3 = wx_scratch_free(temp, s);
    return rc ? rc : (rc2 ? rc2 : rc3);
}

template <typename T>
int irwt_impl(int ac, int mode, T *x, const T *xw, long m, long n, long ncols, int L, long N, const unsigned char *tree, long ntree, long sm,
              const double *h, const double *g, int F, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    WX_REQUIRE(mode >= 0 && mode <= 2, "bad mode %d", mode);
    WX_REQUIRE(n >= 1 && m >= 0 && N >= 0 && ncols >= 1, "bad sizes");
    const int imode = ac ? 2 : (sm < 0 ? 0 : 1);
    const int Lmax = m > 0 ? (wx_maxlevels(m) < wx_maxlevels(n) ? wx_maxlevels(m) : wx_maxlevels(n)) : wx_maxlevels(n);
    int Leff = L;
    if (mode == WX_MODE_DWT) {
        Leff = (int)(m > 0 ? (ncols - 1) / 3 : ncols - 1);
        WX_REQUIRE(m > 0 ? (ncols == 3L * Leff + 1) : true, "sdwt table must have 3L+1 slices");
    } else if (mode == WX_MODE_WPT) {
        Leff = m > 0 ? wx_ilog2l(ncols) / 2 : wx_ilog2l(ncols);
        WX_REQUIRE(m > 0 ? pow4(Leff) == ncols : (1L << Leff) == ncols, "ArgumentError: number of nodes in xw is not a power of %d", m > 0 ? 4 : 2);
    } else {
        Leff = m > 0 ? wx_quaddepthl(ncols) : wx_ilog2l(ncols + 1) - 1;
        WX_REQUIRE(m > 0 ? (pow4(Leff + 1) - 1) / 3 == ncols : (1L << (Leff + 1)) - 1 == ncols, "xw does not hold a full node table");
        WX_REQUIRE(tree || ntree == 0, "null tree");
    }
    WX_REQUIRE(Leff >= 0 && Leff <= Lmax, "ArgumentError: more nodes in xw than possible for this signal size");
    std::vector<long> sd;
    if (imode == 1) {
        // main2depthshift: @assert sm < 1<<L    Utils.jl:298
        WX_REQUIRE(Leff < 62 && sm < (1L << Leff), "AssertionError: sm < 1<<L (sm=%ld, L=%d)", sm, Leff);
        depth_shifts(sd, sm, Leff);
    }
    if (N == 0) return WX_OK;
    WX_REQUIRE(x && xw, "null signal pointer");
    Taps<T> t;
    if (imode == 2) { memset(&t, 0, sizeof(t)); t.F = 1; }
    else { int rc = wx_make_taps(t, h, g, F); if (rc) return rc; }
    if (Leff == 0) {
        const long sz = (m > 0 ? m : 1) * n;
        return wx_launch_copy<T>(View<T>{x, 1, sz, 0, 0}, View<const T>{xw, 1, sz * ncols, 0, 0}, sz, Batch{N, 1, 1, false}, s);
    }
    return m > 0 ? irwt_2d<T>(imode, mode, x, xw, m, n, ncols, Leff, N, tree, ntree, sd, t, s)
                 : irwt_1d<T>(imode, mode, x, xw, n, ncols, Leff, N, tree, ntree, sd, t, s);
}

}  // namespace

extern "C" {
int wx_rwt_f64(int ac, int mode, double *xw, const double *x, long m, long n, int L, long N, const double *h, const double *g, int F, void *s) { return rwt_impl<double>(ac, mode, xw, x, m, n, L, N, h, g, F, s); }
int wx_rwt_f32(int ac, int mode, float *xw, const float *x, long m, long n, int L, long N, const double *h, const double *g, int F, void *s) { return rwt_impl<float>(ac, mode, xw, x, m, n, L, N, h, g, F, s); }
int wx_irwt_f64(int ac, int mode, double *x, const double *xw, long m, long n, long ncols, int L, long N, const unsigned char *tree, long ntree, long sm, const double *h, const double *g, int F, void *s) { return irwt_impl<double>(ac, mode, x, xw, m, n, ncols, L, N, tree, ntree, sm, h, g, F, s); }
int wx_irwt_f32(int ac, int mode, float *x, const float *xw, long m, long n, long ncols, int L, long N, const unsigned char *tree, long ntree, long sm, const double *h, const double *g, int F, void *s) { return irwt_impl<float>(ac, mode, x, xw, m, n, ncols, L, N, tree, ntree, sm, h, g, F, s); }
}

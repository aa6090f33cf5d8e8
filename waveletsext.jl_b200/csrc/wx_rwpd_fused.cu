// wx_rwpd_fused.cu -- fused 1-D redundant packet trees: swpd / swpt (SWT.jl:840-868, 439-471) and acwpd / acwpt
// (ACWT.jl:733-759, 427-460) over sdwt_step! (swt/swt_one_level.jl:99-127) / acdwt_step! (acwt/acwt_one_level.jl:101-128).
//
// The table xw(n, 2^(L+1)-1, N) is pure write traffic (read x once, write 2^(L+1)-1 columns), so the kernel is built
// around keeping every parent on chip:
//  * a CTA walks the subtree of one node depth-first; the ancestors of the node being computed sit in a stack of
//    shared-memory buffers (one per depth), so a parent is never re-read from HBM;
//  * every finished node leaves through ONE 1-D bulk async copy (cp.async.bulk.global.shared::cta) issued by a single
//    thread and overlapped with the computation of the next node; the SM's LSUs never touch global memory;
//  * a-trous filtering with dilation D = 2^d splits a node into D independent cosets.  A thread owns K consecutive
//    elements of one coset (positions base + k*D), loads a register window of K+F-1 coset samples and produces its K
//    outputs with fully unrolled FMAs (taps from the kernel-parameter constant bank, accumulated in the reference's tap
//    order).  Lanes of a warp walk consecutive positions, so shared-memory loads and stores are conflict free for D >= 16.
//  * the two children of the deepest parent are computed in one pass from one window (the detail outputs are taken F-2
//    coset positions ahead so both filters read the same samples);
//  * long trees are split in two launches (top levels d < d0, then one item per depth-d0 node) so that the stack is short
//    enough for two resident CTAs per SM.
#include "wx_steps.cuh"
#include "wx_tma.cuh"
#include <cstdlib>

namespace {

template <typename T, int F>
struct RwCfg {
    static constexpr int LGK = (F <= 16) ? 4 : 3;       // (K = 8 for Float32 was measured: slower, 1.94 vs 1.71 ms -- more loads per output)
    static constexpr int K = 1 << LGK;                 // outputs per thread (consecutive elements of one coset)
};

// window start (in coset steps relative to the first output) of each filter
//   stationary w1: v[i - D + jD]     -> -1          swt/swt_one_level.jl:117-118
//   stationary w2: v[i - jD]         -> -(F-1)      swt/swt_one_level.jl:121-122
//   autocorr     : v[i + (t - c) D]  -> 1 - c, c = F/2 + 1   acwt/acwt_one_level.jl:118-124
template <int F, int AC, int WHICH>
struct RwWin { static constexpr int M0 = AC ? -(F / 2) : (WHICH == 0 ? -1 : -(F - 1)); };

// AC = 2: autocorrelation filters with EXACTLY symmetric taps (checked on the host): P = [rev(b), c, b], Q = [-rev(b), c, -b]
// (acwt/acwt_utils.jl:27-48), so  w1 = c v0 + s,  w2 = c v0 - s  with  s = sum_j b_j (v[+j] + v[-j]):  F + 2 flops per output
// pair instead of 4F - 2.  Every tree kernel of this file uses the same form for a given filter, so acwpt == acwpd[:, leaves]
// (test/transforms.jl:153-154) stays bitwise; the result differs from the tap-by-tap order of acdwt_step! by rounding only.
template <typename T, int F>
__device__ __forceinline__ T rw_ac_sym_s(const T *win, const Taps<T> &tp)
{
    constexpr int M = (F - 1) / 2;
    T s = tp.g[M + 1] * (win[M + 1] + win[M - 1]);
#pragma unroll
    for (int j = 2; j <= M; ++j) s = fma(tp.g[M + j], win[M + j] + win[M - j], s);
    return s;
}

template <typename T, int F, int AC, int WHICH>
__device__ __forceinline__ T rw_dot(const T *win, const Taps<T> &tp)
{
    T a;
    if (AC == 2) {
        constexpr int M = (F - 1) / 2;
        const T s = rw_ac_sym_s<T, F>(win, tp);
        a = fma(tp.g[M], win[M], WHICH == 0 ? s : -s);
    } else if (AC) {
        a = (WHICH == 0 ? tp.g[0] : tp.h[0]) * win[0];
#pragma unroll
        for (int j = 1; j < F; ++j) a = fma(WHICH == 0 ? tp.g[j] : tp.h[j], win[j], a);
    } else if (WHICH == 0) {
        a = tp.g[F - 1] * win[0];
#pragma unroll
        for (int j = 1; j < F; ++j) a = fma(tp.g[F - 1 - j], win[j], a);
    } else {
        a = tp.h[0] * win[F - 1];
#pragma unroll
        for (int j = 1; j < F; ++j) a = fma(tp.h[j], win[F - 1 - j], a);
    }
    return a;
}

// one child (WHICH = 0: w1 / scaling / P, 1: w2 / detail / Q) of the node in src; parent depth d
template <typename T, int F, int AC, int WHICH>
__device__ __forceinline__ void rw_pass1(const T *__restrict__ src, T *__restrict__ dst, int n, int d, const Taps<T> &tp, int tid, int nthr)
{
    using C = RwCfg<T, F>;
    constexpr int K = C::K, W = K + F - 1, M0 = RwWin<F, AC, WHICH>::M0;
    const int D = 1 << d, mask = n - 1;
    for (int u = tid; u < (n >> C::LGK); u += nthr) {
        const int base = ((u >> d) << (d + C::LGK)) | (u & (D - 1));
        T win[W];
        int idx = (base + M0 * D) & mask;
#pragma unroll
        for (int m = 0; m < W; ++m) { win[m] = src[idx]; idx = (idx + D) & mask; }
#pragma unroll
        for (int k = 0; k < K; ++k) dst[base + k * D] = rw_dot<T, F, AC, WHICH>(&win[k], tp);
    }
}

// both children in one pass
template <typename T, int F, int AC>
__device__ __forceinline__ void rw_pass2(const T *__restrict__ src, T *__restrict__ dst0, T *__restrict__ dst1, int n, int d, const Taps<T> &tp,
                                         int tid, int nthr)
{
    using C = RwCfg<T, F>;
    constexpr int K = C::K, W = K + F - 1, M0 = RwWin<F, AC, 0>::M0;
    constexpr int SH = AC ? 0 : F - 2;                 // detail outputs are taken SH coset positions ahead
    const int D = 1 << d, mask = n - 1;
    for (int u = tid; u < (n >> C::LGK); u += nthr) {
        const int base = ((u >> d) << (d + C::LGK)) | (u & (D - 1));
        T win[W];
        int idx = (base + M0 * D) & mask;
#pragma unroll
        for (int m = 0; m < W; ++m) { win[m] = src[idx]; idx = (idx + D) & mask; }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            if (AC == 2) {                               // both children from one symmetric sum (same operations as rw_dot<.., 2, ..>)
                constexpr int M = (F - 1) / 2;
                const T sm = rw_ac_sym_s<T, F>(&win[k], tp);
                dst0[base + k * D] = fma(tp.g[M], win[k + M], sm);
                dst1[base + k * D] = fma(tp.g[M], win[k + M], -sm);
            } else {
                dst0[base + k * D] = rw_dot<T, F, AC, 0>(&win[k], tp);
                dst1[(base + (k + SH) * D) & mask] = rw_dot<T, F, AC, 1>(&win[k], tp);
            }
        }
    }
}

// column of node (depth d, index idx) in the output table
__device__ __forceinline__ long rw_col(int wpt, int L, int d, long idx) { return wpt ? (idx << (L - d)) : ((1L << d) - 1 + idx); }

// item = (signal k, node j0 of depth dstart); the CTA computes every descendant of that node down to depth dend.
// wpt = 0: heap-ordered packet table, every node is stored.  wpt = 1: only depth-dend nodes are stored, at column
// idx << (L - dend) (swpt!'s in-place order, SWT.jl:454-470).
template <typename T, int F, int AC>
__global__ void __launch_bounds__(256) rwpd_dfs_k(T *__restrict__ xw, const T *__restrict__ x, int n, long ncols, int dstart, int dend, int L,
                                                 int wpt, int chain, long items, Taps<T> tp)
{
    extern __shared__ __align__(128) unsigned char wx_rw_smem[];
    __shared__ __align__(8) unsigned long long bar;
    T *bufs = reinterpret_cast<T *>(wx_rw_smem);
    const int E = dend - dstart;
    T *lb0 = bufs + (size_t)E * n, *lb1 = lb0 + n;
    const unsigned nbytes = (unsigned)n * (unsigned)sizeof(T);
    const int tid = threadIdx.x, nthr = blockDim.x;

    if (tid == 0) {
        wx_mbar_init(&bar, 1);
        wx_fence_mbar_init();
    }
    __syncthreads();
    unsigned parity = 0;

    for (long item = blockIdx.x; item < items; item += gridDim.x) {
        const long k = item >> dstart;
        const long j0 = item & ((1L << dstart) - 1);
        T *xk = xw + k * ncols * n;
        if (chain) {
            // One launch for the whole tree: the item rebuilds its depth-dstart node from x by walking the dstart ancestors on
            // its path (dstart single-child passes, ping-pong between the two leaf buffers, last one into bufs[0]).  An ancestor
            // is stored by the item that is its leftmost descendant, column 0 (= x) by item 0 of the signal.
            T *cur = lb0;
            if (tid == 0) {
                wx_bulk_wait_read0();
                wx_mbar_expect_tx(&bar, nbytes);
                wx_bulk_load_1d(cur, x + k * n, nbytes, &bar);
            }
            wx_mbar_wait(&bar, parity);
            parity ^= 1;
            if (j0 == 0 && !wpt && tid == 0) {                         // xw[:,1] = x    SWT.jl:857
                wx_bulk_store_1d(xk, cur, nbytes);
                wx_bulk_commit();
            }
            for (int d = 0; d < dstart; ++d) {
                T *dst = (d == dstart - 1) ? bufs : (cur == lb0 ? lb1 : lb0);
                const long anc = j0 >> (dstart - 1 - d);               // index of the path node of depth d+1
                if (tid == 0) wx_bulk_wait_read0();                    // dst may still be feeding the store issued two passes ago
                __syncthreads();
                if (anc & 1) rw_pass1<T, F, AC, 1>(cur, dst, n, d, tp, tid, nthr);
                else         rw_pass1<T, F, AC, 0>(cur, dst, n, d, tp, tid, nthr);
                wx_fence_proxy_async();
                __syncthreads();
                const bool mine = (j0 & ((1L << (dstart - 1 - d)) - 1)) == 0;      // leftmost descendant of that node
                if (tid == 0 && !wpt && mine) {
                    wx_bulk_store_1d(xk + rw_col(0, L, d + 1, anc) * n, dst, nbytes);
                    wx_bulk_commit();
                }
                cur = dst;
            }
        } else {
            const T *src = (dstart == 0) ? (x + k * n) : (xk + rw_col(wpt, L, dstart, j0) * n);
            if (tid == 0) {
                wx_bulk_wait_read0();                                  // stores of the previous item have released smem
                wx_mbar_expect_tx(&bar, nbytes);
                wx_bulk_load_1d(bufs, src, nbytes, &bar);
            }
            wx_mbar_wait(&bar, parity);
            parity ^= 1;
            if (dstart == 0 && !wpt && tid == 0) {                     // xw[:,1] = x    SWT.jl:857
                wx_bulk_store_1d(xk, bufs, nbytes);
                wx_bulk_commit();
            }
        }
        const int npairs = 1 << (E - 1);
        for (int c = 0; c < npairs; ++c) {
            // internal nodes (relative depths 1..E-1) whose path prefix changed with c, top down
            int es = 1;
            if (c != 0) { es = E - 1 - (__ffs(c) - 1); if (es < 1) es = 1; }
            for (int e = es; e < E; ++e) {
                const int pe = c >> (E - 1 - e);
                const T *par = bufs + (size_t)(e - 1) * n;
                T *cur = bufs + (size_t)e * n;
                if (pe & 1) rw_pass1<T, F, AC, 1>(par, cur, n, dstart + e - 1, tp, tid, nthr);
                else        rw_pass1<T, F, AC, 0>(par, cur, n, dstart + e - 1, tp, tid, nthr);
                wx_fence_proxy_async();
                if (tid == 0) wx_bulk_wait_read0();                    // the store issued one node ago has drained its buffer
                __syncthreads();
                if (tid == 0 && !wpt) {
                    wx_bulk_store_1d(xk + rw_col(0, L, dstart + e, (j0 << e) | pe) * n, cur, nbytes);
                    wx_bulk_commit();
                }
            }
            rw_pass2<T, F, AC>(bufs + (size_t)(E - 1) * n, lb0, lb1, n, dend - 1, tp, tid, nthr);
            wx_fence_proxy_async();
            if (tid == 0) wx_bulk_wait_read0();
            __syncthreads();
            if (tid == 0) {
                const long idx = (j0 << E) | (2L * c);
                wx_bulk_store_1d(xk + rw_col(wpt, L, dend, idx) * n, lb0, nbytes);
                wx_bulk_store_1d(xk + rw_col(wpt, L, dend, idx + 1) * n, lb1, nbytes);
                wx_bulk_commit();
            }
        }
    }
    if (tid == 0) wx_bulk_wait_all();
}

// sdwt! / acdwt! (SWT.jl:120-129, ACWT.jl:120-131): the scaling branch only.  xw(n, L+1, N): column L-d = detail of depth d+1,
// column 0 = scaling of depth L.  One signal per CTA iteration: the scaling node ping-pongs between two buffers, each
// detail leaves through a bulk store from one of two detail buffers; after `dend` levels the scaling node is stored in
// column L-dend (= column 0 when dend == L; otherwise the per-depth path continues from there).
template <typename T, int F, int AC>
__global__ void __launch_bounds__(256) rdwt_chain_k(T *__restrict__ xw, const T *__restrict__ x, int n, int L, int dend, long N, Taps<T> tp)
{
    extern __shared__ __align__(128) unsigned char wx_rw_smem[];
    __shared__ __align__(8) unsigned long long bar;
    T *sc[2] = {reinterpret_cast<T *>(wx_rw_smem), reinterpret_cast<T *>(wx_rw_smem) + n};
    T *dt[2] = {sc[1] + n, sc[1] + 2 * (size_t)n};
    const unsigned nbytes = (unsigned)n * (unsigned)sizeof(T);
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) {
        wx_mbar_init(&bar, 1);
        wx_fence_mbar_init();
    }
    __syncthreads();
    unsigned parity = 0;
    for (long k = blockIdx.x; k < N; k += gridDim.x) {
        T *xk = xw + k * (long)(L + 1) * n;
        if (tid == 0) {
            wx_bulk_wait_read0();
            wx_mbar_expect_tx(&bar, nbytes);
            wx_bulk_load_1d(sc[0], x + k * n, nbytes, &bar);
        }
        wx_mbar_wait(&bar, parity);
        parity ^= 1;
        int cur = 0;
        for (int d = 0; d < dend; ++d) {
            rw_pass2<T, F, AC>(sc[cur], sc[cur ^ 1], dt[d & 1], n, d, tp, tid, nthr);
            wx_fence_proxy_async();
            if (tid == 0) wx_bulk_wait_read0();
            __syncthreads();
            if (tid == 0) {
                wx_bulk_store_1d(xk + (long)(L - d) * n, dt[d & 1], nbytes);
                wx_bulk_commit();
            }
            cur ^= 1;
        }
        if (tid == 0) {
            wx_bulk_store_1d(xk + (long)(L - dend) * n, sc[cur], nbytes);
            wx_bulk_commit();
        }
    }
    if (tid == 0) wx_bulk_wait_all();
}

template <typename T, int F, int AC>
int rdwt_chain_plan(T *xw, const T *x, long n, int L, long N, const Taps<T> &t, cudaStream_t s, int *done)
{
    using C = RwCfg<T, F>;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const int lgn = wx_ilog2l(n);
    if (lgn < C::LGK) return WX_OK;
    int dend = lgn - C::LGK + 1;
    if (dend > L) dend = L;
    const size_t smem = (size_t)4 * n * sizeof(T);
    if (smem > dv.smem_optin) return WX_OK;
    int threads = (int)(((n >> C::LGK) + 31) / 32 * 32);
    if (threads > 256) threads = 256;
    auto kern = rdwt_chain_k<T, F, AC>;
    WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    WX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) return WX_OK;
    long blocks = (long)dv.sms * occ;
    if (blocks > N) blocks = N;
    kern<<<(unsigned)blocks, threads, smem, s>>>(xw, x, (int)n, L, dend, N, t);
    WX_LAUNCHED();
    *done = dend;
    return WX_OK;
}

template <typename T, int F, int AC>
int rwpd_launch(T *xw, const T *x, long n, long ncols, int dstart, int dend, int L, int wpt, int chain, long N, const Taps<T> &t, cudaStream_t s)
{
    using C = RwCfg<T, F>;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const size_t smem = (size_t)(dend - dstart + 2) * n * sizeof(T);
    int threads = (int)(((n >> C::LGK) + 31) / 32 * 32);
    if (threads > 256) threads = 256;
    auto kern = rwpd_dfs_k<T, F, AC>;
    WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occmax = 0;
    WX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occmax, kern, threads, smem));
    if (occmax < 1) return wx_fail(WX_EUNSUPPORTED, "rwpd fused kernel does not fit (smem %zu)", smem);
    const long items = N << dstart;
    auto launch = [&](int occ) -> int {
        long blocks = (long)dv.sms * occ;
        if (blocks > items) blocks = items;
        kern<<<(unsigned)blocks, threads, smem, s>>>(xw, x, (int)n, ncols, dstart, dend, L, wpt, chain, items, t);
        WX_LAUNCHED();
        return WX_OK;
    };
    // resident CTAs per SM: a pure store stream like wpd1d_tma_k (see wx_wpd1d.cu) -- the first large launch of a shape measures
    // the candidates, small launches take the maximum
    int cand[8], nc = 0, occ = occmax;
    for (int c = 1; c <= occmax && c <= 6; ++c) cand[nc++] = c;
    const char *oenv = getenv("WX_B200_RWPD_OCC");
    if (oenv && atoi(oenv) >= 1) occ = atoi(oenv) < occmax ? atoi(oenv) : occmax;
    else {
        const bool big = items >= 4L * dv.sms * occmax && (double)N * (double)n * (double)(1L << (dend + (wpt ? 0 : 1))) * sizeof(T) >= 256e6;
        rc = wx_tuned_choice(WxTuneKey{(const void *)kern, n, (long)dstart * 64 + dend, (long)wpt * 2 + chain, (long)threads}, nc, cand, occmax, big, s, launch, &occ);
        if (rc) return rc;
    }
    return launch(occ);
}

template <typename T, int F, int AC>
int rwpd_plan(T *xw, const T *x, long n, int L, int wpt, long N, const Taps<T> &t, cudaStream_t s, int *done)
{
    using C = RwCfg<T, F>;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const int lgn = wx_ilog2l(n);
    if (lgn < C::LGK) return WX_OK;
    int dend = lgn - C::LGK + 1;                       // deepest parent has K*2^d <= n
    if (dend > L) dend = L;
    if (wpt && dend < L) return WX_OK;                 // in-place leaf order is only produced for complete trees
    const size_t buf = (size_t)n * sizeof(T);
    const long ncols = wpt ? (1L << L) : ((1L << (L + 1)) - 1);
    // stages: the last one (most of the table) gets a stack short enough for two CTAs per SM when that still fuses
    // >= 3 levels; the levels above it run in stages of as many levels as fit one CTA per SM.
    const long efull = (long)(dv.smem_optin / buf) - 2;
    const long ehalf = (long)((dv.smem_optin / 2 - 1024) / buf) - 2;
    if (efull < 1) return WX_OK;
    long elast = (ehalf >= 3 || ehalf >= dend) ? ehalf : efull;
    if (elast > 6) elast = 6;                          // measured: deeper stacks lose more to occupancy than they save in re-reads
    static const char *env = getenv("WX_B200_RWPD_ELAST");        // measurement knob
    if (env && atol(env) >= 1 && atol(env) <= efull) elast = atol(env);
    if (elast > dend) elast = dend;
    int d = 0;
    const int top = dend - (int)elast;
    static const bool nochain = getenv("WX_B200_RWPD_NO_CHAIN") != nullptr;      // measurement knob
    if (top > 0 && top <= 4 && !nochain && !AC && sizeof(T) == 8) {
        // short top: every item rebuilds its depth-`top` node from x itself (top extra single-child passes per item, ~5 % more
        // arithmetic) -- one launch, and the depth-`top` nodes are never re-read from HBM.  Measured: stationary Float64 3.08 ->
        // 2.91 ms; the autocorrelation filters (2F-1 taps) and Float32 (issue bound) lose to the extra passes, so they keep two launches
        rc = rwpd_launch<T, F, AC>(xw, x, n, ncols, top, dend, L, wpt, 1, N, t, s);
        if (rc) return rc;
        *done = dend;
        return WX_OK;
    }
    while (d < top) {
        const int e = (top - d > efull) ? (int)efull : top - d;
        rc = rwpd_launch<T, F, AC>(xw, x, n, ncols, d, d + e, L, wpt, 0, N, t, s);
        if (rc) return rc;
        d += e;
    }
    rc = rwpd_launch<T, F, AC>(xw, x, n, ncols, top, dend, L, wpt, 0, N, t, s);
    if (rc) return rc;
    *done = dend;
    return WX_OK;
}

// the autocorrelation taps (g = P, h = Q) have the exact symmetric structure the AC = 2 kernels assume
// ... and the fused kernel covers the WHOLE tree (n >= K * 2^(L-1)): a partially fused tree finishes through the tap-by-tap step
// kernels, and acwpt (leaves only, not fused in that case) must stay bitwise equal to the leaves of acwpd
template <typename T>
bool ac_taps_symmetric(const Taps<T> &t, long n, int L)
{
    if (t.F < 3 || (t.F & 1) == 0) return false;
    if (wx_ilog2l(n) - ((t.F <= 16) ? 4 : 3) + 1 < L) return false;
    static const bool off = getenv("WX_B200_NO_AC_SYM") != nullptr;          // A-B measurements only
    if (off) return false;
    const int M = (t.F - 1) / 2;
    if (t.h[M] != t.g[M]) return false;
    for (int j = 1; j <= M; ++j)
        if (t.g[M + j] != t.g[M - j] || t.h[M + j] != -t.g[M + j] || t.h[M - j] != -t.g[M - j]) return false;
    return true;
}

}  // namespace

// Runs the leading `*done` levels of the tree (0 = nothing handled; the caller continues with the per-depth path).
// wpt = 0: swpd/acwpd table (column 0 = x included), wpt = 1: swpt/acwpt leaves.
template <typename T>
int wx_rwpd1d_fused(int ac, int wpt, T *xw, const T *x, long n, int L, long N, const Taps<T> &t, cudaStream_t s, int *done)
{
    *done = 0;
    if (L < 1 || N < 1 || !wx_ispow2(n) || n >= (1L << 30) || L > 30) return WX_OK;
    if ((n * sizeof(T)) % 16 != 0 || ((((uintptr_t)xw) | ((uintptr_t)x)) & 15) != 0) return WX_OK;
    static const bool off = getenv("WX_B200_NO_FUSED_RWPD") != nullptr;    // debugging / A-B measurements only
    if (off) return WX_OK;
#define WX_RW_CASE(FF, AA) case FF: return rwpd_plan<T, FF, AA>(xw, x, n, L, wpt, N, t, s, done);
    if (ac && ac_taps_symmetric(t, n, L)) {
        switch (t.F) { WX_RW_CASE(3, 2) WX_RW_CASE(7, 2) WX_RW_CASE(11, 2) WX_RW_CASE(15, 2) WX_RW_CASE(19, 2) WX_RW_CASE(23, 2) WX_RW_CASE(27, 2) WX_RW_CASE(31, 2) WX_RW_CASE(35, 2) WX_RW_CASE(39, 2) }
    } else if (ac) {
        switch (t.F) { WX_RW_CASE(3, 1) WX_RW_CASE(7, 1) WX_RW_CASE(11, 1) WX_RW_CASE(15, 1) WX_RW_CASE(19, 1) WX_RW_CASE(23, 1) WX_RW_CASE(27, 1) WX_RW_CASE(31, 1) WX_RW_CASE(35, 1) WX_RW_CASE(39, 1) }
    } else {
        switch (t.F) { WX_RW_CASE(2, 0) WX_RW_CASE(4, 0) WX_RW_CASE(6, 0) WX_RW_CASE(8, 0) WX_RW_CASE(10, 0) WX_RW_CASE(12, 0) WX_RW_CASE(14, 0) WX_RW_CASE(16, 0) WX_RW_CASE(18, 0) WX_RW_CASE(20, 0) WX_RW_CASE(24, 0) }
    }
#undef WX_RW_CASE
    return WX_OK;
}
template int wx_rwpd1d_fused<double>(int, int, double *, const double *, long, int, long, const Taps<double> &, cudaStream_t, int *);
template int wx_rwpd1d_fused<float>(int, int, float *, const float *, long, int, long, const Taps<float> &, cudaStream_t, int *);

// sdwt / acdwt (scaling chain): runs the leading `*done` levels; the scaling node of depth *done is left in column L-*done.
template <typename T>
int wx_rdwt1d_fused(int ac, T *xw, const T *x, long n, int L, long N, const Taps<T> &t, cudaStream_t s, int *done)
{
    *done = 0;
    if (L < 1 || N < 1 || !wx_ispow2(n) || n >= (1L << 30) || L > 30) return WX_OK;
    if ((n * sizeof(T)) % 16 != 0 || ((((uintptr_t)xw) | ((uintptr_t)x)) & 15) != 0) return WX_OK;
    static const bool off = getenv("WX_B200_NO_FUSED_RWPD") != nullptr;
    if (off) return WX_OK;
#define WX_RC_CASE(FF, AA) case FF: return rdwt_chain_plan<T, FF, AA>(xw, x, n, L, N, t, s, done);
    if (ac && ac_taps_symmetric(t, n, L)) {
        switch (t.F) { WX_RC_CASE(3, 2) WX_RC_CASE(7, 2) WX_RC_CASE(11, 2) WX_RC_CASE(15, 2) WX_RC_CASE(19, 2) WX_RC_CASE(23, 2) WX_RC_CASE(27, 2) WX_RC_CASE(31, 2) WX_RC_CASE(35, 2) WX_RC_CASE(39, 2) }
    } else if (ac) {
        switch (t.F) { WX_RC_CASE(3, 1) WX_RC_CASE(7, 1) WX_RC_CASE(11, 1) WX_RC_CASE(15, 1) WX_RC_CASE(19, 1) WX_RC_CASE(23, 1) WX_RC_CASE(27, 1) WX_RC_CASE(31, 1) WX_RC_CASE(35, 1) WX_RC_CASE(39, 1) }
    } else {
        switch (t.F) { WX_RC_CASE(2, 0) WX_RC_CASE(4, 0) WX_RC_CASE(6, 0) WX_RC_CASE(8, 0) WX_RC_CASE(10, 0) WX_RC_CASE(12, 0) WX_RC_CASE(14, 0) WX_RC_CASE(16, 0) WX_RC_CASE(18, 0) WX_RC_CASE(20, 0) WX_RC_CASE(24, 0) }
    }
#undef WX_RC_CASE
    return WX_OK;
}
template int wx_rdwt1d_fused<double>(int, double *, const double *, long, int, long, const Taps<double> &, cudaStream_t, int *);
template int wx_rdwt1d_fused<float>(int, float *, const float *, long, int, long, const Taps<float> &, cudaStream_t, int *);

// ---- autocorrelation inverses are plain sums: iacdwt_step!(v, w1, w2) = (w1 + w2) / sqrt(2)  acwt/acwt_one_level.jl:217-224 ----
namespace {

// iacwpt! / iacwpd!(xw, L) (ACWT.jl:594-607, 954-969): x = pairwise tree sum of the 2^L depth-L columns starting at column c0,
// in the reference's order ((w_{2b} + w_{2b+1}) / sqrt2 at every level).  One thread per (position, signal): a binary-counter
// reduction, eight columns at a time in registers.
template <typename T>
__global__ void __launch_bounds__(256) iac_tree_sum_k(T *__restrict__ x, const T *__restrict__ xw, long n, long ncols, long c0, int L, long N)
{
    const long gid = (long)blockIdx.x * 256 + threadIdx.x;
    if (gid >= n * N) return;
    const long k = gid / n, i = gid - k * n;
    const T *p = xw + (k * ncols + c0) * n + i;
    const T r2 = (T)1.4142135623730951;
    T stack[32];
    const long nleaf = 1L << L;
    if (L < 3) {
        for (long c = 0; c < nleaf; ++c) {
            T v = p[c * n];
            int lvl = 0;
            while ((c >> lvl) & 1) { v = (stack[lvl] + v) / r2; ++lvl; }
            stack[lvl] = v;
        }
        x[gid] = stack[L];
        return;
    }
    for (long c8 = 0; c8 < (nleaf >> 3); ++c8) {
        T v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) v[j] = __ldcs(p + (c8 * 8 + j) * n);
#pragma unroll
        for (int j = 0; j < 4; ++j) v[j] = (v[2 * j] + v[2 * j + 1]) / r2;
        v[0] = (v[0] + v[1]) / r2; v[1] = (v[2] + v[3]) / r2;
        T a = (v[0] + v[1]) / r2;
        int lvl = 0;
        while ((c8 >> lvl) & 1) { a = (stack[lvl] + a) / r2; ++lvl; }
        stack[lvl] = a;
    }
    x[gid] = stack[L - 3];
}

// iacdwt! (ACWT.jl:292-303): x = col 0; for d = L-1..0: x = (x + col L-d) / sqrt2
template <typename T>
__global__ void __launch_bounds__(256) iac_chain_sum_k(T *__restrict__ x, const T *__restrict__ xw, long n, int L, long N)
{
    const long gid = (long)blockIdx.x * 256 + threadIdx.x;
    if (gid >= n * N) return;
    const long k = gid / n, i = gid - k * n;
    const T *p = xw + k * (long)(L + 1) * n + i;
    const T r2 = (T)1.4142135623730951;
    T a = p[0];
    for (int c = 1; c <= L; ++c) a = (a + __ldcs(p + (long)c * n)) / r2;
    x[gid] = a;
}

}  // namespace

template <typename T>
int wx_iac_tree_sum(T *x, const T *xw, long n, long ncols, long c0, int L, long N, cudaStream_t s)
{
    if (n * N == 0) return WX_OK;
    iac_tree_sum_k<T><<<(unsigned)((n * N + 255) / 256), 256, 0, s>>>(x, xw, n, ncols, c0, L, N);
    WX_LAUNCHED();
    return WX_OK;
}
template <typename T>
int wx_iac_chain_sum(T *x, const T *xw, long n, int L, long N, cudaStream_t s)
{
    if (n * N == 0) return WX_OK;
    iac_chain_sum_k<T><<<(unsigned)((n * N + 255) / 256), 256, 0, s>>>(x, xw, n, L, N);
    WX_LAUNCHED();
    return WX_OK;
}
template int wx_iac_tree_sum<double>(double *, const double *, long, long, long, int, long, cudaStream_t);
template int wx_iac_tree_sum<float>(float *, const float *, long, long, long, int, long, cudaStream_t);
template int wx_iac_chain_sum<double>(double *, const double *, long, int, long, cudaStream_t);
template int wx_iac_chain_sum<float>(float *, const float *, long, int, long, cudaStream_t);

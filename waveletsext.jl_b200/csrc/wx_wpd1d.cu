// wx_wpd1d.cu -- batched 1-D wavelet packet decomposition, all levels fused in one launch.
//
// Reference: wpdall dwt/dwt_all.jl:260-282 -> wpd! DWT.jl:131-161 -> dwt_step! dwt/dwt_one_level.jl:79-107.
//   x(n,N) -> y(n,L+1,N);  level 0 = x;  level d+1 node 2j / 2j+1 = low / high of level d node j.
//
// Design (B200, HBM-bound: algorithmic traffic = (L+2)*n*sizeof(T) per signal, 2*F*L flops per sample):
//  * persistent CTAs, one signal (or one depth-d0 node of a long signal) per CTA iteration;
//  * the node lives in shared memory (ping-pong, 16-byte-chunk XOR swizzle so both the strided window reads
//    and the packed output writes are bank-conflict free); intermediate levels never round-trip HBM;
//  * "wide" levels (node half-length multiple of K): each thread loads a register window of 2S+2K samples with
//    128-bit LDS and produces K low + K high outputs with fully unrolled FMAs (taps come straight from the
//    kernel-parameter constant bank).  The high-pass outputs are taken S positions ahead of the low-pass ones so
//    both filters read the same window;  smem traffic is (2S+2K)/(2K*F) loads per FMA instead of 1;
//  * "small" levels (node length 2,4,8): one thread owns whole nodes in registers, periodic wrap resolved at
//    compile time;
//  * every level row is written to HBM straight from registers with 128-bit streaming stores (32 B / thread,
//    contiguous across the warp), level 0 is copied while the signal is staged.
//  * accumulation order per output is the reference's tap order (j ascending); products use FMA.
#include "wx_steps.cuh"
#include "wx_tma.cuh"
#include "wx_levels.cuh"
#include <cstdlib>

namespace {

// ---- the fused kernel ----------------------------------------------------------------------------------
// item = (signal k, node j0 of depth d0); the CTA computes levels d0+1..L of that node.
template <typename T, int F>
__global__ void __launch_bounds__(256) wpd1d_fused_k(T *__restrict__ y, const T *__restrict__ x, long n, int L, int d0, long items,
                                                    int bufelems, Taps<T> tp)
{
    using C = WpdCfg<T, F>;
    using VT = typename WxVec<T>::type;
    constexpr int V = C::V, K = C::K;
    extern __shared__ __align__(128) unsigned char wx_smem[];
    T *buf0 = reinterpret_cast<T *>(wx_smem);
    T *buf1 = buf0 + bufelems;
    const int n0 = (int)(n >> d0);
    const int tid = threadIdx.x, nthreads = blockDim.x;

    for (long item = blockIdx.x; item < items; item += gridDim.x) {
        const long k = item >> d0;
        const long j0 = item & ((1L << d0) - 1);
        T *ybase = y + k * n * (L + 1) + j0 * n0;
        const T *src = (d0 == 0) ? (x + k * n) : (ybase + (long)d0 * n);

        // stage the node (and emit level 0 when we start from the signal itself)
        for (int c = tid; c < n0 / V; c += nthreads) {
            VT v = wx_ldg_stream<T>(src + c * V);
            *reinterpret_cast<VT *>(buf0 + wx_swz_chunk(c) * V) = v;
            if (d0 == 0) wx_stg_stream(ybase + c * V, v);
        }
        __syncthreads();

        T *a = buf0, *b = buf1;
        const int nlev = L - d0;
        for (int l = 0; l < nlev; ++l) {
            const int p = n0 >> l;
            const bool last = (l == nlev - 1);
            T *grow = ybase + (long)(d0 + l + 1) * n;
            wpd_level<T, F, true>(a, b, grow, n0, p, last, tp, tid, nthreads);
            __syncthreads();
            T *t = a; a = b; b = t;
        }
    }
}


// ---- the fused kernel, TMA variant ----------------------------------------------------------------------
// Same level code, but the SM's load/store units only ever touch shared memory: the node is staged by a TMA
// tensor load (SWIZZLE_128B == wx_swz_chunk) and every level row -- including the level-0 copy -- leaves through
// a TMA bulk tensor store of the smem buffer, issued by one thread and overlapped with the next level's math.
// Buffer reuse: store S_l reads buffer b_l during level l+1; thread 0 waits for it (bulk wait_group.read) right
// before the barrier that lets level l+2 overwrite that buffer, so the wait is normally already satisfied.
// Tensor maps view x and y as 2-D arrays of 128-byte rows: coordinates {0, row}.
// PF (three buffers): the next item's node is prefetched into the buffer the current item does not use, so the load latency of an
// item is hidden behind the previous item's levels instead of being exposed at the top of every iteration.
template <typename T, int F, int KM, bool PF>
__global__ void __launch_bounds__(512) wpd1d_tma_k(const __grid_constant__ CUtensorMap mapx, const __grid_constant__ CUtensorMap mapy, long n, int L,
                                                  int d0, long items, int bufelems, int boxrows, int l2hint, Taps<T> tp)
{
    constexpr int RE = 128 / (int)sizeof(T);          // elements per 128-byte row
    const unsigned long long pol = wx_policy_evict_first();
    extern __shared__ unsigned char wx_smem_raw[];
    __shared__ __align__(8) unsigned long long bar;
    // 1024-byte alignment for SWIZZLE_128B, as an offset from the array so that the accesses stay LDS/STS (not generic LD/ST)
    T *X = reinterpret_cast<T *>(wx_smem_raw + ((1024u - (wx_smem_u32(wx_smem_raw) & 1023u)) & 1023u));   // holds the item's node
    T *P = X + bufelems;                                                                                  // first level's output
    T *N = PF ? P + bufelems : nullptr;                                                                   // landing buffer of the NEXT item
    const int n0 = (int)(n >> d0);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int noderows = n0 / RE, nbox = noderows / boxrows;
    const long sigrows = n / RE;
    const int nlev = L - d0;

    if (tid == 0) {
        wx_mbar_init(&bar, 1);
        wx_fence_mbar_init();
    }
    __syncthreads();
    unsigned parity = 0;

    // node of `item` -> dst (one thread)
    auto load_item = [&](long item, T *dst) {
        const long k = item >> d0;
        const long j0 = item & ((1L << d0) - 1);
        const long yrow0 = k * sigrows * (L + 1) + j0 * noderows;
        wx_mbar_expect_tx(&bar, (unsigned)(n0 * sizeof(T)));
        for (int bx = 0; bx < nbox; ++bx) {
            if (d0 == 0 && l2hint) wx_tma_load_2d_hint(dst + (long)bx * boxrows * RE, &mapx, 0, (int)(k * sigrows + (long)bx * boxrows), &bar, pol);
            else if (d0 == 0) wx_tma_load_2d(dst + (long)bx * boxrows * RE, &mapx, 0, (int)(k * sigrows + (long)bx * boxrows), &bar);
            else         wx_tma_load_2d(dst + (long)bx * boxrows * RE, &mapy, 0, (int)(yrow0 + (long)d0 * sigrows + (long)bx * boxrows), &bar);
        }
    };
    auto store_row = [&](long row, const T *src) {
        for (int bx = 0; bx < nbox; ++bx) {
            if (l2hint) wx_tma_store_2d_hint(&mapy, 0, (int)(row + (long)bx * boxrows), src + (long)bx * boxrows * RE, pol);
            else wx_tma_store_2d(&mapy, 0, (int)(row + (long)bx * boxrows), src + (long)bx * boxrows * RE);
        }
        wx_bulk_commit();
    };

    if (PF && tid == 0 && (long)blockIdx.x < items) load_item(blockIdx.x, X);
    for (long item = blockIdx.x; item < items; item += gridDim.x) {
        const long k = item >> d0;
        const long j0 = item & ((1L << d0) - 1);
        const long yrow0 = k * sigrows * (L + 1) + j0 * noderows;      // row of level 0, this node
        if (PF) {
            // level 1 writes P, which fed the SECOND most recent store of the previous item (the most recent one reads the buffer that
            // is now N and is waited for before the prefetch below): leave one group pending
            if (tid == 0) wx_bulk_wait_read1();
            wx_mbar_wait(&bar, parity);                                // this item's node, prefetched during the previous item
            parity ^= 1;
            __syncthreads();                                           // thread 0's wait holds for everybody
        } else {
            if (tid == 0) {
                wx_bulk_wait_read0();                                  // stores of the previous item have released smem
                load_item(item, X);
            }
            wx_mbar_wait(&bar, parity);
            parity ^= 1;
        }
        if (d0 == 0 && tid == 0) store_row(yrow0, X);                  // level 0 = x   (DWT.jl:142)

        T *a = X, *b = P;
        for (int l = 0; l < nlev; ++l) {
            const int p = n0 >> l;
            wpd_level<T, F, false, false, KM>(a, b, nullptr, n0, p, false, tp, tid, nthreads);
            wx_fence_proxy_async();                                    // my smem writes -> visible to the TMA engine
            if (tid == 0) wx_bulk_wait_read0();                        // buffer `a` is no longer being read by an older store
            __syncthreads();
            if (tid == 0) {
                store_row(yrow0 + (long)(d0 + l + 1) * sigrows, b);
                // every older store has finished reading shared memory (wait above): N is free, fetch the next item's node
                if (PF && l == 0 && item + gridDim.x < items) load_item(item + gridDim.x, N);
            }
            T *t = a; a = b; b = t;
        }
        if (PF) {
            // rotate: the prefetched buffer becomes X; the buffer of the last level's output (its store is the most recent group)
            // becomes the next landing buffer, the other one the next P
            T *Fb = (nlev & 1) ? P : X, *Gb = (nlev & 1) ? X : P;
            X = N; P = Gb; N = Fb;
        }
    }
    if (tid == 0) wx_bulk_wait_all();
}

// one level for the whole batch through the generic step kernel:  level i of y -> level i+1 of y
template <typename T>
int wpd_level_generic(T *y, long n, int L, long N, int i, const Taps<T> &t, cudaStream_t s)
{
    const long np = n >> i, nr = np / 2, str = n * (L + 1);
    const T *v = y + (long)i * n;
    T *w = y + (long)(i + 1) * n;
    return wx_launch_dwt_step<T>(View<T>{w, 1, np, str, 0}, View<T>{w + nr, 1, np, str, 0}, View<const T>{v, 1, np, str, 0}, np,
                                 Batch{1L << i, N, 1, false}, t, s);
}

template <typename T, int F>
int wpd1d_launch_fused(T *y, const T *x, long n, int L, long N, int d0, const Taps<T> &t, cudaStream_t s)
{
    using C = WpdCfg<T, F>;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const long n0 = n >> d0;
    const long bufbytes = ((n0 * (long)sizeof(T) + 127) / 128) * 128;
    const size_t smem = (size_t)2 * bufbytes;
    long units = n0 / (2 * C::K);
    int threads = (int)((units + 31) / 32 * 32);
    if (threads < 64) threads = 64;
    if (threads > 256) threads = 256;
    auto kern = wpd1d_fused_k<T, F>;
    WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    WX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) return wx_fail(WX_EUNSUPPORTED, "wpd1d fused kernel does not fit (smem %zu)", smem);
    const long items = N << d0;
    long blocks = (long)dv.sms * occ;
    if (blocks > items) blocks = items;
    kern<<<(unsigned)blocks, threads, smem, s>>>(y, x, n, L, d0, items, (int)(bufbytes / sizeof(T)), t);
    WX_LAUNCHED();
    return WX_OK;
}


template <typename T, int F, int KM>
int wpd1d_launch_tma(T *y, const T *x, long n, int L, long N, int d0, const Taps<T> &t, cudaStream_t s, bool *handled)
{
    using C = WpdCfg<T, F, KM>;
    *handled = false;
    constexpr long RE = 128 / (long)sizeof(T);
    const long n0 = n >> d0;
    if (n0 % RE != 0 || n % RE != 0) return WX_OK;
    const long xrows = N * (n / RE), yrows = xrows * (L + 1);
    if (yrows >= (1L << 31)) return WX_OK;
    const long noderows = n0 / RE;
    long boxrows = noderows < 256 ? noderows : 256;
    while (noderows % boxrows) --boxrows;
    if (noderows / boxrows > 64) return WX_OK;
    CUtensorMap mx, my;
    if (wx_make_rowmap(&mx, (const void *)x, sizeof(T), xrows, boxrows) != WX_OK) return WX_OK;   // no TMA -> LSU variant
    if (wx_make_rowmap(&my, (const void *)y, sizeof(T), yrows, boxrows) != WX_OK) return WX_OK;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const long bufbytes = ((n0 * (long)sizeof(T) + 1023) / 1024) * 1024;
    const size_t smem2 = (size_t)2 * bufbytes + 1024, smem3 = (size_t)3 * bufbytes + 1024;
    if (smem2 > dv.smem_optin) return WX_OK;
    long units = n0 / (2 * C::K);
    int threads = (int)((units + 31) / 32 * 32);
    if (threads < 64) threads = 64;
    // Resident CTAs per SM and the prefetch buffer.  The kernel is write-dominated (L+1 rows out per row in) and the best residency
    // is NOT the maximum: fewer concurrent store streams suit the DRAM better as long as the resident warps still cover the
    // arithmetic.  Measured (profiles/r2_wpd1d_residency_sweep.jsonl, 8 filters x 2 element types x n in {1024, 4096}): haar F64
    // n = 4096 runs 4.63 ms with ONE CTA per SM against 5.25 ms with three; db4 F64 wants 2, coif4 / sym8 3; n = 1024 wants 3-6,
    // Float32 2-8.  The three-buffer variant (next node prefetched during the current item) helps the long filters (sym8 F64
    // 5.79 -> 5.36 ms) and hurts db4 / db5 F64 (4.81 -> 5.24 ms, profiles/r2_wpd1d_prefetch_ab.jsonl).  The first large launch of a
    // shape therefore measures the (prefetch, residency) candidates (wx_tuned_choice); small launches use the warp-count rule.
    // Knobs: WX_B200_WPD1D_THREADS (256 / 512), WX_B200_WPD1D_OCC, WX_B200_WPD1D_NO_PREFETCH, WX_B200_WPD1D_L2HINT, WX_B200_AUTOTUNE=0.
    const char *tenv = getenv("WX_B200_WPD1D_THREADS");             // read per call so that one process can sweep them
    const char *oenv = getenv("WX_B200_WPD1D_OCC");
    const char *henv = getenv("WX_B200_WPD1D_L2HINT");
    const char *penv = getenv("WX_B200_WPD1D_NO_PREFETCH");
    const int l2hint = henv ? atoi(henv) : 0;       // evict_first hints on the bulk tensor copies: measured, no effect (profiles/r2_wpd1d_ab.jsonl)
    if (threads > 256) threads = (tenv && atoi(tenv) == 512 && threads >= 512) ? 512 : 256;
    auto kern2 = wpd1d_tma_k<T, F, KM, false>;
    auto kern3 = wpd1d_tma_k<T, F, KM, true>;
    WX_CUDA(cudaFuncSetAttribute(kern2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
    int occmax2 = 0, occmax3 = 0;
    WX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occmax2, kern2, threads, smem2));
    if (occmax2 < 1) return WX_OK;
    if (!penv && smem3 <= dv.smem_optin) {
        WX_CUDA(cudaFuncSetAttribute(kern3, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3));
        WX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occmax3, kern3, threads, smem3));
    }
    const long items = N << d0;
    // candidate code = residency + 100 * prefetch
    auto launch = [&](int code) -> int {
        const int occ = code % 100;
        long blocks = (long)dv.sms * occ;
        if (blocks > items) blocks = items;
        if (code >= 100) kern3<<<(unsigned)blocks, threads, smem3, s>>>(mx, my, n, L, d0, items, (int)(bufbytes / sizeof(T)), (int)boxrows, l2hint, t);
        else             kern2<<<(unsigned)blocks, threads, smem2, s>>>(mx, my, n, L, d0, items, (int)(bufbytes / sizeof(T)), (int)boxrows, l2hint, t);
        WX_LAUNCHED();
        return WX_OK;
    };
    int code;
    if (oenv && atoi(oenv) >= 1) {
        code = atoi(oenv) % 100 < occmax2 ? atoi(oenv) % 100 : occmax2;
        if (atoi(oenv) >= 100 && occmax3 >= 1) code = 100 + (atoi(oenv) % 100 < occmax3 ? atoi(oenv) % 100 : occmax3);
    } else {
        // rule: two buffers, resident warps that cover the arithmetic of an F-tap filter (8 + F in Float64, 10 + 1.5 F in Float32)
        const double want = sizeof(T) == 8 ? 8.0 + F : 10.0 + 1.5 * F;
        int rule = (int)(want / (threads / 32) + 0.5);
        if (rule < 1) rule = 1;
        if (rule > occmax2) rule = occmax2;
        if (items <= (long)dv.sms * occmax2) rule = occmax2;    // a launch of at most one wave is latency bound: everything resident
        int cand[40], nc = 0;
        for (int pf = 0; pf < 2; ++pf) {
            const int om = pf ? occmax3 : occmax2;
            for (int c = 1; c <= om && c <= 6; ++c) cand[nc++] = c + 100 * pf;
            if (om >= 8) cand[nc++] = 8 + 100 * pf;
            if (om >= 12) cand[nc++] = 12 + 100 * pf;
            if (om > 6 && om != 8 && om != 12) cand[nc++] = om + 100 * pf;
        }
        const bool big = items >= 4L * dv.sms * occmax2 && (double)N * (double)n * (L + 2) * sizeof(T) >= 256e6;
        rc = wx_tuned_choice(WxTuneKey{(const void *)kern2, n0, (long)(L - d0), (long)threads, (long)l2hint + 2 * (penv ? 1 : 0)}, nc, cand, rule, big, s,
                             launch, &code);
        if (rc) return rc;
    }
    rc = launch(code);
    if (rc) return rc;
    *handled = true;
    return WX_OK;
}

template <typename T>
int wpd1d_impl(T *y, const T *x, long n, int L, long N, const double *h, const double *g, int F, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    WX_REQUIRE(n >= 1 && N >= 0, "wpd: bad sizes n=%ld N=%ld", n, N);
    // reference: @assert 0 <= L <= maxtransformlevels(x)   DWT.jl:137
    WX_REQUIRE(L >= 0 && L <= wx_maxlevels(n), "AssertionError: 0 <= L <= maxtransformlevels(x) (n=%ld, L=%d)", n, L);
    if (N == 0) return WX_OK;
    WX_REQUIRE(y && x, "null signal pointer");
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    WxDev dv; rc = wx_devinfo(dv); if (rc) return rc;

    constexpr int V = WxVec<T>::N;
    const bool aligned = (((uintptr_t)y | (uintptr_t)x) & 15) == 0;
    const bool fusedF = (F >= 2 && F <= 20 && F % 2 == 0) || F == 24;      // every even length of Wavelets.jl's tables (dbN, symN, coifN, beyl, vaid)
    // smallest start depth whose node fits the ping-pong buffers
    int d0 = 0;
    while (d0 < L && (size_t)2 * (((n >> d0) * sizeof(T) + 127) / 128 * 128) > dv.smem_optin) ++d0;
    const long n0 = n >> d0;
    const bool fits = (size_t)2 * ((n0 * sizeof(T) + 127) / 128 * 128) <= dv.smem_optin;
    const bool fused = L > 0 && aligned && fusedF && fits && n0 % (2 * V) == 0 && n < (1L << 30) && d0 < L;

    if (!fused || d0 > 0) {
        // level 0 = x                                                                DWT.jl:142
        rc = wx_launch_copy<T>(View<T>{y, 1, n * (L + 1), 0, 0}, View<const T>{x, 1, n, 0, 0}, n, Batch{N, 1, 1, false}, s);
        if (rc) return rc;
    }
    const int pre = fused ? d0 : L;
    for (int i = 0; i < pre; ++i) { rc = wpd_level_generic<T>(y, n, L, N, i, t, s); if (rc) return rc; }
    if (!fused) return WX_OK;
    static const bool no_tma = getenv("WX_B200_NO_TMA") != nullptr;       // debugging / A-B measurements only
#define WX_WPD_CASE(FF)                                                                             \
    case FF: {                                                                                      \
        bool handled = false;                                                                       \
        if (!no_tma) {                                                                              \
            /* 14+ taps on nodes of at least 128 wide units: the wider window (WpdCfg KM = 4) */        \
            if (FF >= 14 && (n >> d0) / (8 * V) >= 128) rc = wpd1d_launch_tma<T, FF, (FF >= 14 ? 4 : 2)>(y, x, n, L, N, d0, t, s, &handled); \
            else rc = wpd1d_launch_tma<T, FF, 2>(y, x, n, L, N, d0, t, s, &handled);                  \
            if (rc) return rc;                                                                      \
        }                                                                                           \
        if (handled) return WX_OK;                                                                  \
        return wpd1d_launch_fused<T, FF>(y, x, n, L, N, d0, t, s);                                  \
    }
    switch (F) {
        WX_WPD_CASE(2) WX_WPD_CASE(4) WX_WPD_CASE(6) WX_WPD_CASE(8) WX_WPD_CASE(10) WX_WPD_CASE(12) WX_WPD_CASE(14) WX_WPD_CASE(16) WX_WPD_CASE(18) WX_WPD_CASE(20) WX_WPD_CASE(24)
    }
#undef WX_WPD_CASE
    return wx_fail(WX_EUNSUPPORTED, "unreachable");
}

}  // namespace

extern "C" {
int wx_wpd1d_f64(double *y, const double *x, long n, int L, long N, const double *h, const double *g, int F, void *stream)
{
    return wpd1d_impl<double>(y, x, n, L, N, h, g, F, stream);
}
int wx_wpd1d_f32(float *y, const float *x, long n, int L, long N, const double *h, const double *g, int F, void *stream)
{
    return wpd1d_impl<float>(y, x, n, L, N, h, g, F, stream);
}
}

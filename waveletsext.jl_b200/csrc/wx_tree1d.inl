// wx_tree1d.inl (compiled once per element type by wx_tree1d_f64.cu / wx_tree1d_f32.cu: the translation unit is the longest of the
// build) -- fused 1-D wavelet packet transforms by tree: wpt / iwpt (Wavelets.jl wpt!/iwpt!, call sites
// dwt/dwt_all.jl:162,221 -> wptall / iwptall) and iwpd (DWT.jl:337-351 = getbasiscoef + iwpt!), every level in ONE launch.
//
// A persistent CTA stages one signal (or one depth-d0 node of a signal too long for shared memory) in shared memory,
// runs all levels between two swizzled ping-pong buffers (wx_levels.cuh) and writes the result once: HBM traffic is the
// algorithmic 2*n*sizeof(T) per signal instead of 2*n*sizeof(T) per level.  Nodes the tree does not split are copied
// through (children occupy exactly the parent's range).  For iwpd the getbasiscoef gather (Utils.jl:101-134) is fused
// into the staging loads through a per-position leaf-depth map.
#include "wx_steps.cuh"
#include "wx_levels.cuh"

#ifndef WX_TREE_MAXT
#define WX_TREE_MAXT 256          // threads per CTA of the fused tree kernels (A-B builds: 512 with WX_IWPT_KM=2)
#endif

namespace {

__device__ __forceinline__ void wx_cp_async16(void *smem_dst, const void *gsrc)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void wx_cp_async_wait() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// item = (signal k, node j0 of depth d0).  Levels d0..nlev-1 of that node are processed (forward: top down, inverse:
// bottom up).  depth != nullptr: x is a packet table (n, Kx, N) and position e is read from level depth[e].
template <typename T, int F, bool INV, bool TREE, int KM>
__global__ void __launch_bounds__(WX_TREE_MAXT) tree1d_fused_k(T *__restrict__ y, const T *__restrict__ x, long n, int d0, int nlev, long items, int bufelems,
                                                     const unsigned char *__restrict__ tree, long ntree, const unsigned char *__restrict__ depth,
                                                     int Kx, int vecgather, Taps<T> tp)
{
    using VT = typename WxVec<T>::type;
    constexpr int V = WxVec<T>::N;
    extern __shared__ __align__(128) unsigned char wx_tr_smem[];
    T *buf0 = reinterpret_cast<T *>(wx_tr_smem);
    T *buf1 = buf0 + bufelems;
    const int n0 = (int)(n >> d0);
    const int tid = threadIdx.x, nthreads = blockDim.x;
    const int nl = nlev - d0;

    for (long item = blockIdx.x; item < items; item += gridDim.x) {
        const long k = item >> d0;
        const long j0 = item & ((1L << d0) - 1);
        const long pos0 = j0 * n0;
        // ---- stage in -------------------------------------------------------------------------------
        if (depth == nullptr) {                       // 16-byte asynchronous copies: every chunk of the thread in flight, no register staging
            const T *src = x + k * n + pos0;
            for (int c = tid; c < n0 / V; c += nthreads) wx_cp_async16(buf0 + wx_swz_chunk(c) * V, src + c * V);
            wx_cp_async_wait();
        } else if (vecgather) {                       // every leaf is at least one 16-byte chunk long
            const T *src = x + k * (long)Kx * n + pos0;
            for (int c = tid; c < n0 / V; c += nthreads) {
                const long lv = depth[pos0 + (long)c * V];
                wx_cp_async16(buf0 + wx_swz_chunk(c) * V, src + lv * n + c * V);
            }
            wx_cp_async_wait();
        } else {
            const T *src = x + k * (long)Kx * n + pos0;
            for (int e = tid; e < n0; e += nthreads) buf0[wx_swz_elem<T>(e)] = src[(long)depth[pos0 + e] * n + e];
        }
        __syncthreads();
        // ---- levels ---------------------------------------------------------------------------------
        T *a = buf0, *b = buf1;
        int q0 = 0;
        if (INV && nl >= 4 && (n0 >> (nl - 1)) == 2 && n0 % 16 == 0) {
            // down to nodes of length 2: the four deepest levels in registers (along the tree: nodes that are not split pass through)
            if (TREE) {
                auto mask = [&](int l) {                   // split flags of relative level l of this item
                    const long first = ((1L << (d0 + l)) - 1) + (j0 << l);
                    return TreeMask{tree + first, ntree - first};
                };
                iwpt_small_levels4_tree<T, F>(a, b, n0, tp, tid, nthreads, mask(nl - 1), mask(nl - 2), mask(nl - 3), mask(nl - 4));
            } else {
                iwpt_small_levels4<T, F>(a, b, n0, tp, tid, nthreads);
            }
            __syncthreads();
            T *t = a; a = b; b = t;
            q0 = 4;
        }
        // forward: the four deepest levels (nodes of length 16, 8, 4, 2) run in registers after the loop
        const bool fwd4 = !INV && nl >= 4 && (n0 >> (nl - 1)) == 2 && n0 % 16 == 0;
        const int qend = fwd4 ? nl - 4 : nl;
        for (int q = q0; q < qend; ++q) {
            const int l = INV ? nl - 1 - q : q;       // level relative to the staged node
            const int d = d0 + l;
            const long first = ((1L << d) - 1) + (j0 << l);               // 0-based heap position of the node's first depth-d descendant
            TreeMask tm{TREE ? tree + first : nullptr, ntree - first};
            if (INV) iwpt_level<T, F, TREE, KM>(a, b, n0, n0 >> l, tp, tid, nthreads, tm);
            else     wpd_level<T, F, false, TREE, KM>(a, b, nullptr, n0, n0 >> l, false, tp, tid, nthreads, tm);
            __syncthreads();
            T *t = a; a = b; b = t;
        }
        if (fwd4) {
            auto mask = [&](int l) {
                const long first = ((1L << (d0 + l)) - 1) + (j0 << l);
                return TreeMask{TREE ? tree + first : nullptr, ntree - first};
            };
            wpd_small_levels4<T, F, TREE>(a, b, n0, tp, tid, nthreads, mask(nl - 4), mask(nl - 3), mask(nl - 2), mask(nl - 1));
            __syncthreads();
            T *t = a; a = b; b = t;
        }
        // ---- stage out ------------------------------------------------------------------------------
        T *dst = y + k * n + pos0;
        for (int c = tid; c < n0 / V; c += nthreads)
            wx_stg_stream(dst + c * V, *reinterpret_cast<const VT *>(a + wx_swz_chunk(c) * V));
        __syncthreads();
    }
}

template <typename T, int F, bool INV, bool TREE, int KM>
int launch_km(T *y, const T *x, long n, long N, int d0, int nlev, const unsigned char *dtree, long ntree, const unsigned char *ddepth, int Kx,
              int vecgather, const Taps<T> &t, cudaStream_t s)
{
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    constexpr int V = WxVec<T>::N;
    const long n0 = n >> d0;
    const long bufbytes = ((n0 * (long)sizeof(T) + 127) / 128) * 128;
    const size_t smem = (size_t)2 * bufbytes;
    long units = n0 / (2 * (INV ? IwptCfg<T, F, KM>::K : WpdCfg<T, F, KM>::K));
    int threads = (int)((units + 31) / 32 * 32);
    if (threads < 64) threads = 64;
    if (threads > WX_TREE_MAXT) threads = WX_TREE_MAXT;
    auto kern = tree1d_fused_k<T, F, INV, TREE, KM>;
    WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    WX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) return wx_fail(WX_EUNSUPPORTED, "tree1d fused kernel does not fit (smem %zu)", smem);
    const long items = N << d0;
    long blocks = (long)dv.sms * occ;
    if (blocks > items) blocks = items;
    kern<<<(unsigned)blocks, threads, smem, s>>>(y, x, n, d0, nlev, items, (int)(bufbytes / sizeof(T)), dtree, ntree, ddepth, Kx, vecgather, t);
    WX_LAUNCHED();
    return WX_OK;
}

// 4V output pairs per thread unless the node is too short to give a CTA 64 such units (then 2V).  (The by-tree kernels are bound by
// the shared-memory / FP pipes, not by HBM like wpdall, so the wider window pays for every filter length.)
template <typename T, int F, bool INV, bool TREE>
int launch(T *y, const T *x, long n, long N, int d0, int nlev, const unsigned char *dtree, long ntree, const unsigned char *ddepth, int Kx,
           int vecgather, const Taps<T> &t, cudaStream_t s)
{
    constexpr int V = WxVec<T>::N;
    static const bool fwd2 = getenv("WX_B200_WPT_KM2") != nullptr;          // A-B measurements only
    if ((n >> d0) / (8 * V) < 64 || (!INV && fwd2)) return launch_km<T, F, INV, TREE, 2>(y, x, n, N, d0, nlev, dtree, ntree, ddepth, Kx, vecgather, t, s);
    return launch_km<T, F, INV, TREE, 4>(y, x, n, N, d0, nlev, dtree, ntree, ddepth, Kx, vecgather, t, s);
}

template <typename T, int F>
int launch_f(bool inverse, bool full, T *y, const T *x, long n, long N, int d0, int nlev, const unsigned char *dtree, long ntree,
             const unsigned char *ddepth, int Kx, int vecgather, const Taps<T> &t, cudaStream_t s)
{
    if (inverse) {
        if (full) return launch<T, F, true, false>(y, x, n, N, d0, nlev, dtree, ntree, ddepth, Kx, vecgather, t, s);
        return launch<T, F, true, true>(y, x, n, N, d0, nlev, dtree, ntree, ddepth, Kx, vecgather, t, s);
    }
    if (full) return launch<T, F, false, false>(y, x, n, N, d0, nlev, dtree, ntree, ddepth, Kx, vecgather, t, s);
    return launch<T, F, false, true>(y, x, n, N, d0, nlev, dtree, ntree, ddepth, Kx, vecgather, t, s);
}

}  // namespace

// smallest start depth whose node fits the ping-pong buffers, or -1 when the fused kernel does not cover the shape
template <typename T>
int wx_tree1d_fused_depth(const T *y, const T *x, long n, int nlev, int F)
{
    WxDev dv;
    if (wx_devinfo(dv)) return -1;
    constexpr int V = WxVec<T>::N;
    const bool fusedF = (F >= 2 && F <= 20 && F % 2 == 0) || F == 24;
    if (!fusedF || nlev < 1 || n >= (1L << 30) || ((((uintptr_t)y) | ((uintptr_t)x)) & 15) != 0) return -1;
    int d0 = 0;
    while (d0 < nlev && (size_t)2 * (((n >> d0) * sizeof(T) + 127) / 128 * 128) > dv.smem_optin) ++d0;
    if (d0 >= nlev) return -1;
    if ((n >> d0) % (2 * V) != 0) return -1;
    return d0;
}

// levels d0..nlev-1 of the tree in one launch.  full: every node above depth nlev is split (dtree unused).
// ddepth != nullptr: x is a packet table (n, Kx, N) gathered through the per-position leaf depth map (d0 must be 0).
template <typename T>
int wx_tree1d_fused(bool inverse, bool full, T *y, const T *x, long n, long N, int d0, int nlev, const unsigned char *dtree, long ntree,
                    const unsigned char *ddepth, int Kx, int vecgather, const Taps<T> &t, cudaStream_t s)
{
#define WX_TR_CASE(FF) case FF: return launch_f<T, FF>(inverse, full, y, x, n, N, d0, nlev, dtree, ntree, ddepth, Kx, vecgather, t, s);
    switch (t.F) { WX_TR_CASE(2) WX_TR_CASE(4) WX_TR_CASE(6) WX_TR_CASE(8) WX_TR_CASE(10) WX_TR_CASE(12) WX_TR_CASE(14) WX_TR_CASE(16) WX_TR_CASE(18) WX_TR_CASE(20) WX_TR_CASE(24) }
#undef WX_TR_CASE
    return wx_fail(WX_EUNSUPPORTED, "tree1d fused: filter length %d", t.F);
}

#define WX_TR_INST(T)                                                                                                                       \
    template int wx_tree1d_fused_depth<T>(const T *, const T *, long, int, int);                                                            \
    template int wx_tree1d_fused<T>(bool, bool, T *, const T *, long, long, int, int, const unsigned char *, long, const unsigned char *, \
                                    int, int, const Taps<T> &, cudaStream_t);
WX_TR_INST(WX_TR_TYPE)

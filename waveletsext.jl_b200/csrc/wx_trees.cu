// wx_trees.cu -- decimated tree drivers other than the fused 1-D WPD: 2-D WPD, WPT / iWPT by tree (1-D, 2-D),
// getbasiscoefall gather and iWPD.  One launch per level for the whole batch; a node that the tree does not
// split is passed through unchanged (the children of a node occupy exactly the parent's range / block).
//
// Reference: DWT.jl:164-209 (wpd! 2-D), :500-548 (wpt! 2-D), :662-710 (iwpt! 2-D), :322-401 (iwpd!),
//            Utils.jl:101-197 (getbasiscoef / getbasiscoefall); the 1-D wpt!/iwpt! are Wavelets.jl's
//            (call sites dwt/dwt_all.jl:162,221), restated as dwt_step!/idwt_step! applied over the tree.
#include "wx_steps.cuh"
#include <vector>
#include <cstdlib>

// fused all-levels-in-one-launch kernels (wx_tree1d.inl, one translation unit per element type)
template <typename T> int wx_tree1d_fused_depth(const T *y, const T *x, long n, int nlev, int F);
template <typename T>
int wx_tree1d_fused(bool inverse, bool full, T *y, const T *x, long n, long N, int d0, int nlev, const unsigned char *dtree, long ntree,
                    const unsigned char *ddepth, int Kx, int vecgather, const Taps<T> &t, cudaStream_t s);

// fused 2-D packet decomposition (wx_wpd2d.cu) and inverse by quad tree (wx_iwpt2d.cu)
template <typename T>
int wx_iwpt2d_fused(T *y, const T *xw, T *scratch, long m, long n, int nlev, long N, const unsigned char *dtree, long ntree, const Taps<T> &t,
                    cudaStream_t s, bool *handled);
template <typename T> int wx_wpd2d_fused(T *y, const T *x, long m, long n, int L, long N, const Taps<T> &t, cudaStream_t s, bool *handled);
template <typename T>
int wx_wpt2d_fused(T *y, const T *x, T *scratch, long m, long n, int nlev, long N, const unsigned char *dtree, long ntree, const Taps<T> &t,
                   cudaStream_t s, bool *handled);

template <typename T>
int wx_gather_multi(T *out, const T *Xw, long m, long n, int K, long N, const unsigned char *trees, long ntree, cudaStream_t s);

namespace {

constexpr int kT = 256;
static inline unsigned gridf(long total) { return (unsigned)((total + kT - 1) / kT); }

// heap index of quad node (depth d, block row jr, block col jc): children 4i-2 (TL) 4i-1 (TR) 4i (BL) 4i+1 (BR)
__device__ __forceinline__ long quad_index(int d, int jr, int jc)
{
    long idx = 1;
    for (int b = d - 1; b >= 0; --b) idx = 4 * idx - 2 + 2 * ((jr >> b) & 1) + ((jc >> b) & 1);
    return idx;
}
__device__ __forceinline__ bool node_on(const unsigned char *tree, long ntree, long idx)
{
    return tree == nullptr || (idx <= ntree && tree[idx - 1]);
}

// one analysis output pair from a strided periodic vector (a1 dwt_step!)
template <typename T>
__device__ __forceinline__ void dwt_pair(const T *pv, long es, long n, long i, const Taps<T> &tp, T &lo, T &hi)
{
    const int F = tp.F;
    long k1 = 2 * i, k2 = 2 * i + 1;
    T a1 = tp.g[F - 1] * pv[k1 * es];
    T a2 = tp.h[0] * pv[k2 * es];
    for (int j = 1; j < F; ++j) {
        k1 += 1; if (k1 >= n) k1 -= n;
        k2 -= 1; if (k2 < 0) k2 += n;
        a1 = fma(tp.g[F - 1 - j], pv[k1 * es], a1);
        a2 = fma(tp.h[j], pv[k2 * es], a2);
    }
    lo = a1; hi = a2;
}

// one synthesis output (a2 idwt_step!), i0 0-based
template <typename T>
__device__ __forceinline__ T idwt_elem(const T *p1, const T *p2, long es, long n, long i0, const Taps<T> &tp)
{
    const int F = tp.F;
    const long n1 = n / 2;
    long i = i0 + 1;
    int j0 = (i & 1) ? 1 : 2;
    int j1 = F - j0 + 1;
    int j2 = ((i + 1) & 1) ? 1 : 2;
    long k1 = (i + 1) >> 1, k2 = k1;
    T acc = fma(tp.g[j1 - 1], p1[(k1 - 1) * es], tp.h[j2 - 1] * p2[(k2 - 1) * es]);
    for (int j = j0 + 2; j <= F; j += 2) {
        j1 = F - j + 1;
        j2 = j + ((j & 1) ? 1 : -1);
        k1 -= 1; if (k1 <= 0) k1 += n1;
        k2 += 1; if (k2 > n1) k2 -= n1;
        acc += fma(tp.g[j1 - 1], p1[(k1 - 1) * es], tp.h[j2 - 1] * p2[(k2 - 1) * es]);
    }
    return acc;
}

// ---- 1-D, one depth of a tree -----------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kT) wpt1_level_k(T *dst, const T *src, long n, long N, int d, const unsigned char *tree, long ntree, Taps<T> tp)
{
    long idx = (long)blockIdx.x * kT + threadIdx.x;
    if (idx >= (n / 2) * N) return;
    const long gi = idx % (n / 2), k = idx / (n / 2);
    const long p = n >> d, half = p / 2;
    const long j = gi / half, i = gi - j * half;
    const T *s = src + k * n + j * p;
    T *o = dst + k * n + j * p;
    if (node_on(tree, ntree, (1L << d) + j)) {
        T lo, hi;
        dwt_pair(s, 1, p, i, tp, lo, hi);
        o[i] = lo; o[half + i] = hi;
    } else {
        o[2 * i] = s[2 * i]; o[2 * i + 1] = s[2 * i + 1];
    }
}

template <typename T>
__global__ void __launch_bounds__(kT) iwpt1_level_k(T *dst, const T *src, long n, long N, int d, const unsigned char *tree, long ntree, Taps<T> tp)
{
    long idx = (long)blockIdx.x * kT + threadIdx.x;
    if (idx >= n * N) return;
    const long e = idx % n, k = idx / n;
    const long p = n >> d;
    const long j = e / p, i = e - j * p;
    const T *s = src + k * n + j * p;
    T v;
    if (node_on(tree, ntree, (1L << d) + j)) v = idwt_elem(s, s + p / 2, 1, p, i, tp);
    else v = s[i];
    dst[k * n + e] = v;
}

// ---- 2-D, one depth: images (m rows x n cols), column-major; ss/ts/ds = image strides of src/temp/dst ----
template <typename T>
__global__ void __launch_bounds__(kT) dwt2_cols_k(T *temp, long ts, const T *src, long ss, long m, long n, long N, int d,
                                                  const unsigned char *tree, long ntree, Taps<T> tp)
{
    long idx = (long)blockIdx.x * kT + threadIdx.x;
    const long hm = m / 2;
    if (idx >= hm * n * N) return;
    const long r2 = idx % hm, c = (idx / hm) % n, k = idx / (hm * n);
    const long mp = m >> d, np = n >> d, hmp = mp / 2;
    const long jr = r2 / hmp, i = r2 - jr * hmp, jc = c / np;
    if (!node_on(tree, ntree, quad_index(d, (int)jr, (int)jc))) return;
    T lo, hi;
    dwt_pair(src + k * ss + c * m + jr * mp, 1, mp, i, tp, lo, hi);
    T *t = temp + k * ts + c * m + jr * mp;
    t[i] = lo; t[hmp + i] = hi;
}

template <typename T>
__global__ void __launch_bounds__(kT) dwt2_rows_k(T *dst, long ds, const T *temp, long ts, const T *src, long ss, long m, long n, long N, int d,
                                                  const unsigned char *tree, long ntree, Taps<T> tp)
{
    long idx = (long)blockIdx.x * kT + threadIdx.x;
    const long hn = n / 2;
    if (idx >= m * hn * N) return;
    const long r = idx % m, c2 = (idx / m) % hn, k = idx / (m * hn);
    const long mp = m >> d, np = n >> d, hnp = np / 2;
    const long jc = c2 / hnp, i = c2 - jc * hnp, jr = r / mp;
    T *o = dst + k * ds + r;
    if (node_on(tree, ntree, quad_index(d, (int)jr, (int)jc))) {
        T lo, hi;
        dwt_pair(temp + k * ts + (jc * np) * m + r, m, np, i, tp, lo, hi);
        o[(jc * np + i) * m] = lo; o[(jc * np + hnp + i) * m] = hi;
    } else if (dst != src) {
        const T *s = src + k * ss + r;
        o[(2 * c2) * m] = s[(2 * c2) * m]; o[(2 * c2 + 1) * m] = s[(2 * c2 + 1) * m];
    }
}

template <typename T>
__global__ void __launch_bounds__(kT) idwt2_rows_k(T *temp, long ts, const T *src, long ss, long m, long n, long N, int d,
                                                   const unsigned char *tree, long ntree, Taps<T> tp)
{
    long idx = (long)blockIdx.x * kT + threadIdx.x;
    if (idx >= m * n * N) return;
    const long r = idx % m, c = (idx / m) % n, k = idx / (m * n);
    const long mp = m >> d, np = n >> d;
    const long jc = c / np, i = c - jc * np, jr = r / mp;
    if (!node_on(tree, ntree, quad_index(d, (int)jr, (int)jc))) return;
    const T *s = src + k * ss + (jc * np) * m + r;
    temp[k * ts + c * m + r] = idwt_elem(s, s + (np / 2) * m, m, np, i, tp);
}

template <typename T>
__global__ void __launch_bounds__(kT) idwt2_cols_k(T *dst, long ds, const T *temp, long ts, const T *src, long ss, long m, long n, long N, int d,
                                                   const unsigned char *tree, long ntree, Taps<T> tp)
{
    long idx = (long)blockIdx.x * kT + threadIdx.x;
    if (idx >= m * n * N) return;
    const long r = idx % m, c = (idx / m) % n, k = idx / (m * n);
    const long mp = m >> d, np = n >> d;
    const long jr = r / mp, i = r - jr * mp, jc = c / np;
    T v;
    if (node_on(tree, ntree, quad_index(d, (int)jr, (int)jc))) {
        const T *t = temp + k * ts + c * m + jr * mp;
        v = idwt_elem(t, t + mp / 2, 1, mp, i, tp);
    } else {
        v = src[k * ss + c * m + r];
    }
    dst[k * ds + c * m + r] = v;
}

// gather: out[k, e] = Xw[k, depth[e], e]   (getbasiscoefall, Utils.jl:169-197)
template <typename T>
__global__ void __launch_bounds__(kT) gather_k(T *out, const T *Xw, long sz, long K, long N, const unsigned char *depth)
{
    long idx = (long)blockIdx.x * kT + threadIdx.x;
    if (idx >= sz * N) return;
    const long e = idx % sz, k = idx / sz;
    out[idx] = Xw[(k * K + depth[e]) * sz + e];
}
// grid.x = signal, grid.y * blockDim walks the positions: no 64-bit division per element, four gathers in flight per thread
template <typename T>
__global__ void __launch_bounds__(kT) gather_sig_k(T *__restrict__ out, const T *__restrict__ Xw, long sz, long K, const unsigned char *__restrict__ depth)
{
    const long k = blockIdx.x;
    const T *Xk = Xw + k * K * sz;
    T *ok = out + k * sz;
    const long step = (long)gridDim.y * kT;
    long e = (long)blockIdx.y * kT + threadIdx.x;
    for (; e + 3 * step < sz; e += 4 * step) {
        T v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = Xk[(long)depth[e + q * step] * sz + e + q * step];
#pragma unroll
        for (int q = 0; q < 4; ++q) ok[e + q * step] = v[q];
    }
    for (; e < sz; e += step) ok[e] = Xk[(long)depth[e] * sz + e];
}

// the same with V consecutive positions per thread (one 16-byte load / store) when no leaf boundary falls inside an aligned group of V
// positions; grid.x = signal, grid.y * blockDim walks the groups -- no 64-bit division per element
template <typename T>
__global__ void __launch_bounds__(kT) gather_vec_k(T *__restrict__ out, const T *__restrict__ Xw, long sz, long K, const unsigned char *__restrict__ depth)
{
    using VT = typename WxVec<T>::type;
    constexpr int V = WxVec<T>::N;
    const long k = blockIdx.x;
    const T *Xk = Xw + k * K * sz;
    T *ok = out + k * sz;
    const long groups = sz / V;
    for (long c = (long)blockIdx.y * kT + threadIdx.x; c < groups; c += (long)gridDim.y * kT) {
        const long e = c * V;
        *reinterpret_cast<VT *>(ok + e) = __ldcs(reinterpret_cast<const VT *>(Xk + (long)depth[e] * sz + e));
    }
}

// ---- host helpers -------------------------------------------------------------------------------------
struct DevTree {
    unsigned char *d = nullptr;
    cudaStream_t s;
    int upload(const unsigned char *h, long n, cudaStream_t st)
    {
        s = st;
        int rc = wx_scratch(&d, (size_t)n, st);
        if (rc) return rc;
        WX_CUDA(cudaMemcpyAsync(d, h, (size_t)n, cudaMemcpyHostToDevice, st));
        return WX_OK;
    }
    ~DevTree() { if (d) cudaFreeAsync(d, s); }
};

static int tree_maxdepth1(const unsigned char *tree, long ntree)
{
    long last = 0;
    for (long i = ntree; i >= 1; --i) if (tree[i - 1]) { last = i; break; }
    return last ? wx_ilog2l(last) + 1 : 0;       // number of levels to run
}
static int tree_maxdepth2(const unsigned char *tree, long ntree)
{
    long last = 0;
    for (long i = ntree; i >= 1; --i) if (tree[i - 1]) { last = i; break; }
    return last ? wx_quaddepthl(last) + 1 : 0;
}

// wpt / iwpt 1-D by tree
template <typename T>
int tree1d(bool inverse, T *y, const T *x, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    WX_REQUIRE(n >= 1 && N >= 0 && ntree >= 0, "wpt: bad sizes");
    if (N == 0) return WX_OK;
    WX_REQUIRE(y && x && (tree || ntree == 0), "null pointer");
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    const int nlev = tree_maxdepth1(tree, ntree);
    WX_REQUIRE(nlev <= wx_maxlevels(n), "tree is deeper than maxtransformlevels(n)");
    if (nlev == 0) {
        if (y != x) WX_CUDA(cudaMemcpyAsync(y, x, (size_t)n * N * sizeof(T), cudaMemcpyDeviceToDevice, s));
        return WX_OK;
    }
    DevTree dt; rc = dt.upload(tree, ntree, s); if (rc) return rc;
    // every node above depth nlev split?  then the kernels need no per-node test
    bool full = ntree >= (1L << nlev) - 1;
    for (long i = 0; full && i < (1L << nlev) - 1; ++i) full = tree[i] != 0;
    static const bool nofuse = getenv("WX_B200_NO_FUSED_TREE") != nullptr;       // debugging / A-B measurements only
    const int d0 = nofuse ? -1 : wx_tree1d_fused_depth<T>(y, x, n, nlev, F);
    if (d0 == 0) return wx_tree1d_fused<T>(inverse, full, y, x, n, N, 0, nlev, dt.d, ntree, nullptr, 0, 0, t, s);
    T *tmp = nullptr;
    rc = wx_scratch(&tmp, (size_t)n * N, s); if (rc) return rc;
    // levels the per-level kernel runs: all of them, or the d0 coarsest ones around the fused launch
    const int ngen = d0 > 0 ? d0 : nlev;
    const T *cur = x;
    if (d0 > 0 && inverse) {                         // deep levels first, fused; lands where the ping-pong below expects it
        T *fo = (ngen % 2 == 0) ? y : tmp;
        rc = wx_tree1d_fused<T>(true, full, fo, x, n, N, d0, nlev, dt.d, ntree, nullptr, 0, 0, t, s);
        if (rc) { wx_scratch_free(tmp, s); return rc; }
        cur = fo;
    }
    // ping-pong so that the last level lands in y
    for (int q = 0; q < ngen; ++q) {
        const int d = inverse ? ngen - 1 - q : q;
        T *nxt = ((ngen - 1 - q) % 2 == 0) ? y : tmp;
        if (nxt == cur) nxt = (nxt == y) ? tmp : y;     // only when x aliases y
        if (inverse) iwpt1_level_k<T><<<gridf(n * N), kT, 0, s>>>(nxt, cur, n, N, d, dt.d, ntree, t);
        else         wpt1_level_k<T><<<gridf((n / 2) * N), kT, 0, s>>>(nxt, cur, n, N, d, dt.d, ntree, t);
        WX_LAUNCHED();
        cur = nxt;
    }
    if (d0 > 0 && !inverse) {                        // remaining deep levels, fused (in place per node when cur == y)
        rc = wx_tree1d_fused<T>(false, full, y, cur, n, N, d0, nlev, dt.d, ntree, nullptr, 0, 0, t, s);
        if (rc) { wx_scratch_free(tmp, s); return rc; }
        cur = y;
    }
    if (cur != y) WX_CUDA(cudaMemcpyAsync(y, cur, (size_t)n * N * sizeof(T), cudaMemcpyDeviceToDevice, s));
    return wx_scratch_free(tmp, s);
}

// one 2-D level, forward: src -> (temp) -> dst ; inverse likewise.  Processes the batch in chunks to bound scratch.
template <typename T>
int level2d(bool inverse, T *dst, long ds, const T *src, long ss, T *temp, long m, long n, long N, int d,
            const unsigned char *dtree, long ntree, const Taps<T> &t, cudaStream_t s)
{
    const long ts = m * n;
    if (!inverse) {
        dwt2_cols_k<T><<<gridf((m / 2) * n * N), kT, 0, s>>>(temp, ts, src, ss, m, n, N, d, dtree, ntree, t);
        WX_LAUNCHED();
        dwt2_rows_k<T><<<gridf(m * (n / 2) * N), kT, 0, s>>>(dst, ds, temp, ts, src, ss, m, n, N, d, dtree, ntree, t);
        WX_LAUNCHED();
    } else {
        idwt2_rows_k<T><<<gridf(m * n * N), kT, 0, s>>>(temp, ts, src, ss, m, n, N, d, dtree, ntree, t);
        WX_LAUNCHED();
        idwt2_cols_k<T><<<gridf(m * n * N), kT, 0, s>>>(dst, ds, temp, ts, src, ss, m, n, N, d, dtree, ntree, t);
        WX_LAUNCHED();
    }
    return WX_OK;
}

static long chunk_images(long m, long n, size_t elt, long N)
{
    // scratch budget ~1 GiB per buffer
    long per = (long)((size_t)1 << 30) / (long)((size_t)m * n * elt);
    if (per < 1) per = 1;
    return per < N ? per : N;
}

// wpd 2-D : x(m,n,N) -> y(m,n,L+1,N)    DWT.jl:164-209
template <typename T>
int wpd2d_impl(T *y, const T *x, long m, long n, int L, long N, const double *h, const double *g, int F, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    WX_REQUIRE(m >= 1 && n >= 1 && N >= 0, "wpd 2-D: bad sizes");
    const int Lmax = wx_maxlevels(m) < wx_maxlevels(n) ? wx_maxlevels(m) : wx_maxlevels(n);
    WX_REQUIRE(L >= 0 && L <= Lmax, "AssertionError: 0 <= L <= maxtransformlevels(x) (m=%ld, n=%ld, L=%d)", m, n, L);
    if (N == 0) return WX_OK;
    WX_REQUIRE(y && x, "null signal pointer");
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    const long img = m * n, ys = img * (L + 1);
    bool handled = false;
    rc = wx_wpd2d_fused<T>(y, x, m, n, L, N, t, s, &handled);
    if (rc || handled) return rc;
    rc = wx_launch_copy<T>(View<T>{y, 1, ys, 0, 0}, View<const T>{x, 1, img, 0, 0}, img, Batch{N, 1, 1, false}, s);
    if (rc || L == 0) return rc;
    const long Nc = chunk_images(m, n, sizeof(T), N);
    T *temp; rc = wx_scratch(&temp, (size_t)img * Nc, s); if (rc) return rc;
    for (long k0 = 0; k0 < N && !rc; k0 += Nc) {
        const long nk = (N - k0 < Nc) ? N - k0 : Nc;
        T *yk = y + k0 * ys;
        for (int d = 0; d < L && !rc; ++d)
            rc = level2d<T>(false, yk + (long)(d + 1) * img, ys, yk + (long)d * img, ys, temp, m, n, nk, d, nullptr, 0, t, s);
    }
    int rc2 = wx_scratch_free(temp, s);
    return rc ? rc : rc2;
}

template <typename T>
int gather_impl(T *out, const T *Xw, long m, long n, int K, long N, const unsigned char *tree, long ntree, void *stream);

// wpt / iwpt 2-D by quad tree : x(m,n,N) -> y(m,n,N)   DWT.jl:500-548, 662-710
template <typename T>
int tree2d(bool inverse, T *y, const T *x, long m, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F,
           void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    WX_REQUIRE(m >= 1 && n >= 1 && N >= 0 && ntree >= 0, "wpt 2-D: bad sizes");
    if (N == 0) return WX_OK;
    WX_REQUIRE(y && x && (tree || ntree == 0), "null pointer");
    Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
    const int nlev = tree_maxdepth2(tree, ntree);
    const int Lmax = wx_maxlevels(m) < wx_maxlevels(n) ? wx_maxlevels(m) : wx_maxlevels(n);
    WX_REQUIRE(nlev <= Lmax, "tree is deeper than maxtransformlevels(x)");
    const long img = m * n;
    if (nlev == 0) {
        if (y != x) WX_CUDA(cudaMemcpyAsync(y, x, (size_t)img * N * sizeof(T), cudaMemcpyDeviceToDevice, s));
        return WX_OK;
    }
    DevTree dt; rc = dt.upload(tree, ntree, s); if (rc) return rc;
    const long Nc = chunk_images(m, n, sizeof(T), N);
    if (y != x) {
        // fused kernels (inverse: wx_iwpt2d.cu, forward: wx_wpd2d.cu in tree mode): one halo-tile launch per level whose nodes exceed
        // shared memory, the whole-node kernel for all deeper levels
        bool full = true;
        { const long nfull = ((1L << (2 * nlev)) - 1) / 3; full = ntree >= nfull; for (long i = 0; full && i < nfull; ++i) full = tree[i] != 0; }
        T *scr; rc = wx_scratch(&scr, (size_t)img * Nc, s); if (rc) return rc;
        bool all = true;
        for (long k0 = 0; k0 < N && !rc && all; k0 += Nc) {
            const long nk = (N - k0 < Nc) ? N - k0 : Nc;
            bool handled = false;
            if (inverse) rc = wx_iwpt2d_fused<T>(y + k0 * img, x + k0 * img, scr, m, n, nlev, nk, full ? nullptr : dt.d, ntree, t, s, &handled);
            else rc = wx_wpt2d_fused<T>(y + k0 * img, x + k0 * img, scr, m, n, nlev, nk, full ? nullptr : dt.d, ntree, t, s, &handled);
            if (!handled) all = false;                   // decided from the shape: the same for every chunk; y is recomputed below
        }
        int rcf = wx_scratch_free(scr, s);
        if (rc || rcf) return rc ? rc : rcf;
        if (all) return WX_OK;
    }
    if (!inverse && y != x && m * n * (long)sizeof(T) * (nlev + 1) <= (1L << 31)) {
        // forward by tree = packet table of the chunk (fused wpd kernels) + leaf gather (getbasiscoef)
        long Nt = (long)(((size_t)3 << 30) / ((size_t)img * (nlev + 1) * sizeof(T)));
        if (Nt < 1) Nt = 1;
        if (Nt > N) Nt = N;
        T *tab; rc = wx_scratch(&tab, (size_t)img * (nlev + 1) * Nt, s); if (rc) return rc;
        bool all = true;
        for (long k0 = 0; k0 < N && !rc && all; k0 += Nt) {
            const long nk = (N - k0 < Nt) ? N - k0 : Nt;
            bool handled = false;
            rc = wx_wpd2d_fused<T>(tab, x + k0 * img, m, n, nlev, nk, t, s, &handled);
            if (!rc && !handled) { all = false; break; }
            if (!rc) rc = gather_impl<T>(y + k0 * img, tab, m, n, nlev + 1, nk, tree, ntree, stream);
        }
        int rcf = wx_scratch_free(tab, s);
        if (rc || rcf) return rc ? rc : rcf;
        if (all) return WX_OK;
    }
    T *temp, *pp;
    rc = wx_scratch(&temp, (size_t)img * Nc, s); if (rc) return rc;
    rc = wx_scratch(&pp, (size_t)img * Nc, s); if (rc) return rc;
    for (long k0 = 0; k0 < N && !rc; k0 += Nc) {
        const long nk = (N - k0 < Nc) ? N - k0 : Nc;
        const T *cur = x + k0 * img;
        T *yk = y + k0 * img;
        for (int q = 0; q < nlev && !rc; ++q) {
            const int d = inverse ? nlev - 1 - q : q;
            T *nxt = ((nlev - 1 - q) % 2 == 0) ? yk : pp;
            if (nxt == cur) nxt = (nxt == yk) ? pp : yk;
            rc = level2d<T>(inverse, nxt, img, cur, img, temp, m, n, nk, d, dt.d, ntree, t, s);
            cur = nxt;
        }
        if (!rc && cur != yk) WX_CUDA(cudaMemcpyAsync(yk, cur, (size_t)img * nk * sizeof(T), cudaMemcpyDeviceToDevice, s));
    }
    int rc2 = wx_scratch_free(temp, s), rc3 = wx_scratch_free(pp, s);
    return rc ? rc : (rc2 ? rc2 : rc3);
}

// depth of the leaf covering every position (getleaf utils/utils_tree.jl:122-157 + Utils.jl:116-131)
static int leaf_depth_map(std::vector<unsigned char> &depth, long m, long n, int K, const unsigned char *tree, long ntree)
{
    const bool two = m > 0;
    const long sz = two ? m * n : n;
    depth.assign((size_t)sz, 0);
    // walk the tree from the root; a node that is not split is a leaf
    std::vector<long> stack{1};
    while (!stack.empty()) {
        long i = stack.back(); stack.pop_back();
        const bool split = i <= ntree && tree[i - 1];
        if (split) {
            if (two) for (int c = 0; c < 4; ++c) stack.push_back(4 * i - 2 + c);
            else { stack.push_back(2 * i); stack.push_back(2 * i + 1); }
            continue;
        }
        const int d = two ? wx_quaddepthl(i) : wx_ilog2l(i);
        if (d >= K) return wx_fail(WX_EINVAL, "ArgumentError: Not enough decomposition levels in Xw.");
        if (two) {
            long r0, c0, nr, nc;
            wx_quadrangel(m, n, i, &r0, &c0, &nr, &nc);
            for (long c = 0; c < nc; ++c)
                for (long r = 0; r < nr; ++r) depth[(size_t)((c0 + c) * m + r0 + r)] = (unsigned char)d;
        } else {
            const long n0 = n >> d, nn = i - (1L << d);
            for (long e = 0; e < n0; ++e) depth[(size_t)(nn * n0 + e)] = (unsigned char)d;
        }
    }
    return WX_OK;
}

template <typename T>
int gather_impl(T *out, const T *Xw, long m, long n, int K, long N, const unsigned char *tree, long ntree, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    WX_REQUIRE(n >= 1 && m >= 0 && K >= 1 && N >= 0 && ntree >= 0, "getbasiscoefall: bad sizes");
    if (N == 0) return WX_OK;
    WX_REQUIRE(out && Xw && (tree || ntree == 0), "null pointer");
    std::vector<unsigned char> depth;
    int rc = leaf_depth_map(depth, m, n, K, tree, ntree); if (rc) return rc;
    const long sz = (long)depth.size();
    DevTree dd; rc = dd.upload(depth.data(), sz, s); if (rc) return rc;
    constexpr int V = WxVec<T>::N;
    bool vec = sz % V == 0 && N < (1L << 31) && ((((uintptr_t)out) | ((uintptr_t)Xw)) & 15) == 0;
    for (long e = 0; vec && e < sz; e += V)
        for (int q = 1; q < V; ++q) if (depth[(size_t)(e + q)] != depth[(size_t)e]) { vec = false; break; }
    if (vec) {
        long gy = (sz / V + kT - 1) / kT;
        if (gy > 64) gy = 64;
        gather_vec_k<T><<<dim3((unsigned)N, (unsigned)gy), kT, 0, s>>>(out, Xw, sz, K, dd.d);
    } else if (N < (1L << 31)) {
        long gy = (sz + 4 * kT - 1) / (4 * kT);
        if (gy > 64) gy = 64;
        gather_sig_k<T><<<dim3((unsigned)N, (unsigned)gy), kT, 0, s>>>(out, Xw, sz, K, dd.d);
    } else {
        gather_k<T><<<gridf(sz * N), kT, 0, s>>>(out, Xw, sz, K, N, dd.d);
    }
    WX_LAUNCHED();
    return WX_OK;
}

// iwpd by tree = getbasiscoef + iwpt   (DWT.jl:337-351 ; the 2-D loop :354-401 computes the same values)
template <typename T>
int iwpd_impl(T *x, const T *Xw, long m, long n, int K, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F,
              void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    if (N == 0) return WX_OK;
    const long sz = (m > 0 ? m : 1) * n;
    if (m == 0 && x && Xw && (tree || ntree == 0) && n >= 1 && K >= 1) {
        // 1-D: gather fused into the staging loads of the all-levels kernel
        const int nlev = tree_maxdepth1(tree, ntree);
        static const bool nofuse = getenv("WX_B200_NO_FUSED_TREE") != nullptr;
        if (!nofuse && nlev >= 1 && nlev <= wx_maxlevels(n) && wx_tree1d_fused_depth<T>(x, Xw, n, nlev, F) == 0) {
            Taps<T> t; int rc = wx_make_taps(t, h, g, F); if (rc) return rc;
            std::vector<unsigned char> depth;
            rc = leaf_depth_map(depth, 0, n, K, tree, ntree); if (rc) return rc;
            long minleaf = n;
            for (long e = 0; e < n;) { const long len = n >> depth[(size_t)e]; if (len < minleaf) minleaf = len; e += len; }
            bool full = ntree >= (1L << nlev) - 1;
            for (long i = 0; full && i < (1L << nlev) - 1; ++i) full = tree[i] != 0;
            DevTree dt, dd;
            rc = dt.upload(tree, ntree, s); if (rc) return rc;
            rc = dd.upload(depth.data(), n, s); if (rc) return rc;
            return wx_tree1d_fused<T>(true, full, x, Xw, n, N, 0, nlev, dt.d, ntree, dd.d, K, minleaf >= WxVec<T>::N ? 1 : 0, t, s);
        }
    }
    T *w; int rc = wx_scratch(&w, (size_t)sz * N, s); if (rc) return rc;
    rc = gather_impl<T>(w, Xw, m, n, K, N, tree, ntree, stream);
    if (!rc) rc = (m > 0) ? tree2d<T>(true, x, w, m, n, N, tree, ntree, h, g, F, stream)
                          : tree1d<T>(true, x, w, n, N, tree, ntree, h, g, F, stream);
    int rc2 = wx_scratch_free(w, s);
    return rc ? rc : rc2;
}

}  // namespace

// internal entry used by the fused iwpt path when a shape is not covered
template <typename T>
int wx_tree1d_generic(bool inverse, T *y, const T *x, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F,
                      void *stream)
{
    return tree1d<T>(inverse, y, x, n, N, tree, ntree, h, g, F, stream);
}
template int wx_tree1d_generic<double>(bool, double *, const double *, long, long, const unsigned char *, long, const double *, const double *, int, void *);
template int wx_tree1d_generic<float>(bool, float *, const float *, long, long, const unsigned char *, long, const double *, const double *, int, void *);

extern "C" {

int wx_wpd2d_f64(double *y, const double *x, long m, long n, int L, long N, const double *h, const double *g, int F, void *s) { return wpd2d_impl<double>(y, x, m, n, L, N, h, g, F, s); }
int wx_wpd2d_f32(float *y, const float *x, long m, long n, int L, long N, const double *h, const double *g, int F, void *s) { return wpd2d_impl<float>(y, x, m, n, L, N, h, g, F, s); }
int wx_wpt2d_f64(double *y, const double *x, long m, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *s) { return tree2d<double>(false, y, x, m, n, N, tree, ntree, h, g, F, s); }
int wx_wpt2d_f32(float *y, const float *x, long m, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *s) { return tree2d<float>(false, y, x, m, n, N, tree, ntree, h, g, F, s); }
int wx_iwpt2d_f64(double *y, const double *xw, long m, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *s) { return tree2d<double>(true, y, xw, m, n, N, tree, ntree, h, g, F, s); }
int wx_iwpt2d_f32(float *y, const float *xw, long m, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *s) { return tree2d<float>(true, y, xw, m, n, N, tree, ntree, h, g, F, s); }
int wx_gather_basis_f64(double *out, const double *Xw, long m, long n, int K, long N, const unsigned char *tree, long ntree, void *s) { return gather_impl<double>(out, Xw, m, n, K, N, tree, ntree, s); }
int wx_gather_basis_f32(float *out, const float *Xw, long m, long n, int K, long N, const unsigned char *tree, long ntree, void *s) { return gather_impl<float>(out, Xw, m, n, K, N, tree, ntree, s); }
int wx_gather_basis_multi_f64(double *out, const double *Xw, long m, long n, int K, long N, const unsigned char *trees_dev, long ntree, void *s)
{
    WX_REQUIRE(out && Xw && trees_dev && n >= 1 && m >= 0 && K >= 1 && N >= 0 && ntree >= 0, "getbasiscoefall: bad arguments");
    return wx_gather_multi<double>(out, Xw, m, n, K, N, trees_dev, ntree, (cudaStream_t)s);
}
int wx_gather_basis_multi_f32(float *out, const float *Xw, long m, long n, int K, long N, const unsigned char *trees_dev, long ntree, void *s)
{
    WX_REQUIRE(out && Xw && trees_dev && n >= 1 && m >= 0 && K >= 1 && N >= 0 && ntree >= 0, "getbasiscoefall: bad arguments");
    return wx_gather_multi<float>(out, Xw, m, n, K, N, trees_dev, ntree, (cudaStream_t)s);
}
int wx_iwpd_f64(double *x, const double *Xw, long m, long n, int K, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *s) { return iwpd_impl<double>(x, Xw, m, n, K, N, tree, ntree, h, g, F, s); }
int wx_iwpd_f32(float *x, const float *Xw, long m, long n, int K, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *s) { return iwpd_impl<float>(x, Xw, m, n, K, N, tree, ntree, h, g, F, s); }
int wx_wpt1d_f64(double *y, const double *x, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *s) { return tree1d<double>(false, y, x, n, N, tree, ntree, h, g, F, s); }
int wx_wpt1d_f32(float *y, const float *x, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *s) { return tree1d<float>(false, y, x, n, N, tree, ntree, h, g, F, s); }
int wx_iwpt1d_f64(double *y, const double *xw, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *s) { return tree1d<double>(true, y, xw, n, N, tree, ntree, h, g, F, s); }
int wx_iwpt1d_f32(float *y, const float *xw, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *s) { return tree1d<float>(true, y, xw, n, N, tree, ntree, h, g, F, s); }

}  // extern "C"

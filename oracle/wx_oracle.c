/*
 * wx_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU oracle for the filter-bank hot path of WaveletsExt.jl v0.2.3: a plain-C restatement of the
 * reference's Julia loops (the reference is pure Julia and Julia is not available in this image,
 * so the reference itself cannot be executed here; see DESIGN.md "Oracle").
 *
 * PARITY PINNING: checked in tests/test_oracle_golden.py against every known-answer vector of
 * the reference's own tests for this path (test/transforms.jl:3-22,54-87,124-145,
 * test/wavemult.jl:26-30, test/utils.jl:6-121) and its structural identities (round trips,
 * swpt == swpd[:,leaves], wpd == [x wpt1 wpt2 ...]).  The LSDB cost (AverageShiftedHistograms.jl)
 * and the exact JBB cost values are NOT pinned by any reference test: "parity unpinned" for those.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library.  The product (libwx_b200.so) never links or calls it.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fopenmp -shared -fPIC).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define WXO_API __attribute__((visibility("default")))

/* ---------------- integer helpers (host index algebra) ---------------- */

/* Julia mod(a,n) for n>0 (result in [0,n)) */
static inline long wx_pmod(long a, long n) { long r = a % n; return r < 0 ? r + n : r; }

/* getdepth(i,:binary) utils/utils_tree.jl:252-256 = floor(log2(i)), integer version */
WXO_API int wx_ilog2(long i) { int d = 0; while (i > 1) { i >>= 1; ++d; } return d; }

/* getdepth(i,:quad) utils/utils_tree.jl:257-259 = floor(log(4, 3i-2)), integer version */
WXO_API int wx_quaddepth(long i)
{
    int d = 0; long last = 1, w = 1;      /* last index at depth d */
    while (i > last) { w *= 4; last += w; ++d; }
    return d;
}

/* getrowrange / getcolrange Utils.jl:465-542, 0-based start + extent */
WXO_API void wx_quadrange(long m, long n, long idx, long *r0, long *c0, long *nr, long *nc)
{
    if (idx == 1) { *r0 = 0; *c0 = 0; *nr = m; *nc = n; return; }
    long parent = (idx + 2) / 4;
    long pr0, pc0, pnr, pnc;
    wx_quadrange(m, n, parent, &pr0, &pc0, &pnr, &pnc);
    *nr = pnr / 2; *nc = pnc / 2;
    *r0 = (idx < 4 * parent) ? pr0 : pr0 + pnr / 2;      /* 4i-2, 4i-1 upper ; 4i, 4i+1 lower */
    *c0 = (idx % 2 == 0) ? pc0 : pc0 + pnc / 2;          /* even -> left ; odd -> right */
}

/* maxtransformlevels (Wavelets.jl, recalled): largest k with 2^k | n */
WXO_API int wx_maxtransformlevels(long n) { int k = 0; if (n <= 0) return 0; while (n % 2 == 0) { n /= 2; ++k; } return k; }

/* gettreelength(n,m) utils/utils_tree.jl:290-293 */
WXO_API long wx_treelength2(long nr, long nc)
{
    int L = wx_maxtransformlevels(nr < nc ? nr : nc);
    return ((1L << (2 * L)) - 1) / 3;
}

/* 1-D maketree (Wavelets.jl, recalled; outputs pinned by test/utils.jl:7-8,104): n-1 entries;
 * :full -> first 2^L-1 true ; :dwt -> nodes 1,2,4,...,2^(L-1).  dwt!=0 selects :dwt */
WXO_API void wx_maketree1(unsigned char *tree, long n, int L, int dwt)
{
    long nt = n - 1;
    memset(tree, 0, (size_t)(nt > 0 ? nt : 0));
    if (!dwt) { for (long i = 1; i <= (1L << L) - 1 && i <= nt; ++i) tree[i - 1] = 1; }
    else      { for (int i = 0; i < L; ++i) if ((1L << i) <= nt) tree[(1L << i) - 1] = 1; }
}

/* 2-D maketree utils/utils_tree.jl:197-221 */
WXO_API void wx_maketree2(unsigned char *tree, long nr, long nc, int L, int dwt)
{
    long nq = wx_treelength2(nr, nc);
    memset(tree, 0, (size_t)nq);
    if (!dwt) {
        long cnt = ((1L << (2 * L)) - 1) / 3;
        for (long i = 1; i <= cnt; ++i) tree[i - 1] = 1;
    } else {
        if (nq > 0 && L > 0) tree[0] = 1;
        for (int i = 0; i <= L - 2; ++i) tree[((1L << (2 * i + 2)) + 2) / 3 - 1] = 1;
    }
}

/* delete_subtree! BestBasis.jl:128-140 ; arity 2 or 4 */
WXO_API void wx_delete_subtree(unsigned char *bt, long nt, long i, int arity)
{
    bt[i - 1] = 0;
    for (int c = 0; c < arity; ++c) {
        long ch = (arity == 2) ? 2 * i + c : 4 * i - 2 + c;
        if (ch <= nt && bt[ch - 1]) wx_delete_subtree(bt, nt, ch, arity);
    }
}

/* getleaf(:binary) utils/utils_tree.jl:122-157 ; leaf has 2*nt+1 entries */
WXO_API void wx_getleaf_binary(const unsigned char *tree, long nt, unsigned char *leaf)
{
    memset(leaf, 0, (size_t)(2 * nt + 1));
    leaf[0] = 1;
    for (long i = 1; i <= nt; ++i) {
        if (!tree[i - 1]) continue;
        leaf[i - 1] = 0; leaf[2 * i - 1] = 1; leaf[2 * i] = 1;
    }
}

/* getleaf(:quad) ; leaf has 4*nt+1 entries */
WXO_API void wx_getleaf_quad(const unsigned char *tree, long nt, unsigned char *leaf)
{
    memset(leaf, 0, (size_t)(4 * nt + 1));
    leaf[0] = 1;
    for (long i = 1; i <= nt; ++i) {
        if (!tree[i - 1]) continue;
        leaf[i - 1] = 0;
        leaf[4 * i - 3] = 1; leaf[4 * i - 2] = 1; leaf[4 * i - 1] = 1; leaf[4 * i] = 1;
    }
}

/* 2-D isvalidtree utils/utils_tree.jl:13-29 ; 1-D version is Wavelets.jl (recalled: parent must
 * exist for a node to exist) */
WXO_API int wx_isvalidtree(const unsigned char *tree, long nt, int arity)
{
    for (long i = 1; i <= nt; ++i) {
        if (tree[i - 1]) continue;
        for (int c = 0; c < arity; ++c) {
            long ch = (arity == 2) ? 2 * i + c : 4 * i - 2 + c;
            if (ch <= nt && tree[ch - 1]) return 0;
        }
    }
    return 1;
}

/* main2depthshift Utils.jl:297-305 ; sd has L+1 entries. returns -1 when the @assert fires */
WXO_API int wx_main2depthshift(long *sd, long sm, int L)
{
    if (!(sm < (1L << L)) || sm < 0) return -1;
    long acc = 0;
    sd[0] = 0;
    for (int d = 0; d < L; ++d) { acc += ((sm >> d) & 1L) << d; sd[d + 1] = acc; }
    return 0;
}

/* filters: WT.makereverseqmfpair(wt, true) (Wavelets.jl, recalled, confirmed by the golden vectors):
 * g = reverse(qmf) (scaling), h[i] = qmf[i]*(-1)^i (detail) */
WXO_API void wx_makereverseqmfpair(const double *q, int F, double *g, double *h)
{
    for (int i = 0; i < F; ++i) { g[i] = q[F - 1 - i]; h[i] = (i & 1) ? -q[i] : q[i]; }
}

/* a15 autocorr / pfilter / qfilter / make_acreverseqmfpair  acwt/acwt_utils.jl:7-72
 * P,Q get 2F-1 taps (already "reversed", which is a no-op for the exactly symmetric filters) */
WXO_API void wx_make_acreverseqmfpair(const double *q, int F, double *P, double *Q)
{
    int l = F;
    double *a = (double *)calloc((size_t)(l > 1 ? l - 1 : 1), sizeof(double));
    for (int k = 1; k <= l - 1; ++k) {
        for (int i = 1; i <= l - k; ++i) a[k - 1] += q[i - 1] * q[i + k - 1];
        a[k - 1] *= 2;
    }
    double c1 = 1 / sqrt(2.0), c2 = c1 / 2;
    int Lf = 2 * l - 1;
    double *p = (double *)malloc((size_t)Lf * sizeof(double)), *qq = (double *)malloc((size_t)Lf * sizeof(double));
    for (int k = 0; k < l - 1; ++k) {
        double b = c2 * a[k], bn = -c2 * a[k];
        p[l - 2 - k] = b;  p[l + k] = b;
        qq[l - 2 - k] = bn; qq[l + k] = bn;
    }
    p[l - 1] = c1; qq[l - 1] = c1;
    for (int i = 0; i < Lf; ++i) { P[i] = p[Lf - 1 - i]; Q[i] = qq[Lf - 1 - i]; }
    free(a); free(p); free(qq);
}

/* ---------------- type-templated transforms ---------------- */
#define T double
#define SUF(name) wxo_##name##_f64
#include "wx_oracle_impl.h"
#undef T
#undef SUF

#define T float
#define SUF(name) wxo_##name##_f32
#include "wx_oracle_impl.h"
#undef T
#undef SUF

/* ---------------- batch drivers (the reference's serial `*all` loops) ----------------
 * a6 wpdall dwt/dwt_all.jl:260-282.  Mirrors the reference's per-signal work: a copy of the slice
 * (`Array(x_i)`, :278) and the filter pair rebuilt per signal (DWT.jl:141).  nthreads = 1 is the
 * reference's own (serial) behaviour; nthreads > 1 parallelises over signals with OpenMP and is what
 * a `Threads.@threads` loop over eachslice would give. */
#define WX_BATCH_WPD(SFX, TY)                                                                            \
    WXO_API void wxo_wpdall1_##SFX(TY *y, const TY *x, long n, int L, long N, const double *q, int F,    \
                                   int nthreads)                                                         \
    {                                                                                                    \
        _Pragma("omp parallel for schedule(static) num_threads(nthreads)")                               \
        for (long k = 0; k < N; ++k) {                                                                   \
            double g[64], h[64];                                                                         \
            TY *xi = (TY *)malloc((size_t)n * sizeof(TY));                                               \
            memcpy(xi, x + k * n, (size_t)n * sizeof(TY));                                               \
            wx_makereverseqmfpair(q, F, g, h);                                                           \
            wxo_wpd1_##SFX(y + k * n * (L + 1), xi, n, L, h, g, F);                                      \
            free(xi);                                                                                    \
        }                                                                                                \
    }                                                                                                    \
    WXO_API void wxo_wpdall2_##SFX(TY *y, const TY *x, long m, long n, int L, long N, const double *q,   \
                                   int F, int nthreads)                                                  \
    {                                                                                                    \
        _Pragma("omp parallel for schedule(static) num_threads(nthreads)")                               \
        for (long k = 0; k < N; ++k) {                                                                   \
            double g[64], h[64];                                                                         \
            wx_makereverseqmfpair(q, F, g, h);                                                           \
            wxo_wpd2_##SFX(y + k * m * n * (L + 1), x + k * m * n, m, n, L, h, g, F);                    \
        }                                                                                                \
    }                                                                                                    \
    WXO_API void wxo_iwptall1_##SFX(TY *y, const TY *xw, long n, long N, const unsigned char *tree,      \
                                    long ntree, const double *q, int F, int nthreads)                    \
    {                                                                                                    \
        _Pragma("omp parallel for schedule(static) num_threads(nthreads)")                               \
        for (long k = 0; k < N; ++k) {                                                                   \
            double g[64], h[64];                                                                         \
            wx_makereverseqmfpair(q, F, g, h);                                                           \
            wxo_iwpt1_##SFX(y + k * n, xw + k * n, n, tree, ntree, h, g, F);                             \
        }                                                                                                \
    }                                                                                                    \
    WXO_API void wxo_rwpdall1_##SFX(int ac, TY *xw, const TY *x, long n, int L, long N,                  \
                                    const double *h, const double *g, int F, int nthreads)               \
    {                                                                                                    \
        long cols = (1L << (L + 1)) - 1;                                                                 \
        _Pragma("omp parallel for schedule(static) num_threads(nthreads)")                               \
        for (long k = 0; k < N; ++k) wxo_rwpd1_##SFX(ac, xw + k * n * cols, x + k * n, n, L, h, g, F);   \
    }

WX_BATCH_WPD(f64, double)
WX_BATCH_WPD(f32, float)

WXO_API int wxo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/*
 * wx_oracle_impl.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (plain C) of the filter-bank hot path of WaveletsExt.jl v0.2.3.
 * Included twice by wx_oracle.c, once with T=double (suffix _f64) and once with
 * T=float (suffix _f32).  Every function cites the reference file:line it follows
 * (paths relative to the reference checkout, src/mod/...).  Loop order, tap order and
 * accumulation order follow the reference statement by statement; indices are the
 * 0-based image of the 1-based Julia indices.
 *
 * Element-type semantics: the reference's taps are always Float64
 * (WT.makereverseqmfpair default eltype), so for T=Float32 every `w[i] += g*v` is
 * evaluated in Float64 and rounded to Float32 on the store.  The casts below do
 * exactly that; compile with -ffp-contract=off so no FMA contraction happens
 * (Julia does not contract `a*b + c`).
 *
 * Memory layout is Julia's column-major layout: first index fastest, batch last.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may call into this file.
 */

#ifndef T
#error "define T and SUF before including"
#endif

/* ------------------------------------------------------------------------------------------
 * a1  dwt_step!  1-D   dwt/dwt_one_level.jl:79-107
 * w1,w2: children (length n/2, strides s1,s2); v: parent (length n, stride sv)
 * h = detail (mirror qmf), g = scaling (reversed qmf); F taps each.
 * ---------------------------------------------------------------------------------------- */
WXO_API void SUF(dwt_step)(T *w1, long s1, T *w2, long s2, const T *v, long sv, long n,
                          const double *h, const double *g, int F)
{
    long n1 = n / 2;
    for (long i = 0; i < n1; ++i) {
        long k1 = 2 * i;     /* Julia k1 = 2i-1 (1-based) */
        long k2 = 2 * i + 1; /* Julia k2 = 2i   (1-based) */
        T a1 = (T)(g[F - 1] * (double)v[k1 * sv]);
        T a2 = (T)(h[0] * (double)v[k2 * sv]);
        for (int j = 1; j < F; ++j) {
            k1 = k1 + 1; if (k1 >= n) k1 = wx_pmod(k1, n);
            k2 = k2 - 1; if (k2 < 0)  k2 = wx_pmod(k2, n);
            a1 = (T)((double)a1 + g[F - 1 - j] * (double)v[k1 * sv]);
            a2 = (T)((double)a2 + h[j] * (double)v[k2 * sv]);
        }
        w1[i * s1] = a1;
        w2[i * s2] = a2;
    }
}

/* a2  idwt_step!  1-D   dwt/dwt_one_level.jl:192-223 */
WXO_API void SUF(idwt_step)(T *v, long sv, const T *w1, long s1, const T *w2, long s2, long n,
                           const double *h, const double *g, int F)
{
    long n1 = n / 2;
    for (long i = 1; i <= n; ++i) {              /* keep Julia's 1-based i for the parity logic */
        int j0 = (i % 2 == 1) ? 1 : 2;           /* mod1(i,2) */
        int j1 = F - j0 + 1;                     /* 1-based index into g */
        int j2 = ((i + 1) % 2 == 1) ? 1 : 2;     /* mod1(i+1,2), 1-based index into h */
        long k1 = (i + 1) >> 1;                  /* 1-based */
        long k2 = (i + 1) >> 1;
        T acc = (T)(g[j1 - 1] * (double)w1[(k1 - 1) * s1] + h[j2 - 1] * (double)w2[(k2 - 1) * s2]);
        for (int j = j0 + 2; j <= F; j += 2) {
            j1 = F - j + 1;
            j2 = j + ((j & 1) ? 1 : -1);         /* j + isodd(j) - iseven(j) */
            k1 = k1 - 1; if (k1 <= 0) k1 = wx_pmod(k1 - 1, n1) + 1;   /* mod1 */
            k2 = k2 + 1; if (k2 > n1) k2 = wx_pmod(k2 - 1, n1) + 1;
            acc = (T)((double)acc +
                      (g[j1 - 1] * (double)w1[(k1 - 1) * s1] + h[j2 - 1] * (double)w2[(k2 - 1) * s2]));
        }
        v[(i - 1) * sv] = acc;
    }
}

/* a3  dwt_step! 2-D   dwt/dwt_one_level.jl:319-354
 * children are (nr x nc) blocks, parent/temp are (2nr x 2nc); ld* = leading dimension (rows)
 * of the array each view lives in.  Columns first into temp, then rows. */
WXO_API void SUF(dwt_step2)(T *w1, T *w2, T *w3, T *w4, long ldw,
                           const T *v, long ldv, T *temp, long ldt,
                           long nr, long nc, const double *h, const double *g, int F)
{
    for (long j = 0; j < 2 * nc; ++j)
        SUF(dwt_step)(temp + j * ldt, 1, temp + nr + j * ldt, 1, v + j * ldv, 1, 2 * nr, h, g, F);
    for (long i = 0; i < nr; ++i) {
        SUF(dwt_step)(w1 + i, ldw, w2 + i, ldw, temp + i, ldt, 2 * nc, h, g, F);
        SUF(dwt_step)(w3 + i, ldw, w4 + i, ldw, temp + nr + i, ldt, 2 * nc, h, g, F);
    }
}

/* a3  idwt_step! 2-D   dwt/dwt_one_level.jl:401-436 (rows first, then columns) */
WXO_API void SUF(idwt_step2)(T *v, long ldv, const T *w1, const T *w2, const T *w3, const T *w4,
                            long ldw, T *temp, long ldt, long nr, long nc,
                            const double *h, const double *g, int F)
{
    for (long i = 0; i < nr; ++i) {
        SUF(idwt_step)(temp + i, ldt, w1 + i, ldw, w2 + i, ldw, 2 * nc, h, g, F);
        SUF(idwt_step)(temp + nr + i, ldt, w3 + i, ldw, w4 + i, ldw, 2 * nc, h, g, F);
    }
    for (long j = 0; j < 2 * nc; ++j)
        SUF(idwt_step)(v + j * ldv, 1, temp + j * ldt, 1, temp + nr + j * ldt, 1, 2 * nr, h, g, F);
}

/* a9  sdwt_step! 1-D   swt/swt_one_level.jl:99-127 */
WXO_API void SUF(sdwt_step)(T *w1, long s1, T *w2, long s2, const T *v, long sv, long n, int d,
                           const double *h, const double *g, int F)
{
    long D = 1L << d;
    for (long i = 0; i < n; ++i) {
        long k1 = wx_pmod(i - D, n);   /* mod1(i-(1<<d), n) */
        long k2 = i;
        T a1 = (T)(g[F - 1] * (double)v[k1 * sv]);
        T a2 = (T)(h[0] * (double)v[k2 * sv]);
        for (int j = 1; j < F; ++j) {
            k1 = k1 + D; if (k1 >= n) k1 = wx_pmod(k1, n);
            k2 = k2 - D; if (k2 < 0)  k2 = wx_pmod(k2, n);
            a1 = (T)((double)a1 + g[F - 1 - j] * (double)v[k1 * sv]);
            a2 = (T)((double)a2 + h[j] * (double)v[k2 * sv]);
        }
        w1[i * s1] = a1;
        w2[i * s2] = a2;
    }
}

/* a10 isdwt_step! shift based, 1-D   swt/swt_one_level.jl:279-318
 * Julia precedence: `m-1<<d` == m-(1<<d) ; `m+sp-1<<d` == m+sp-(1<<d) == m.
 * returns 0, or -1 when the reference's @assert on (sv,sw) would fire. */
WXO_API int SUF(isdwt_step_shift)(T *v, long svs, const T *w1, long s1, const T *w2, long s2, long n,
                                 int d, long sv, long sw, const double *h, const double *g, int F,
                                 int add2out)
{
    if (!(0 <= sv && sv < (1L << d))) return -1;
    if (!(sv <= sw && sw < (1L << (d + 1)))) return -1;
    long ip = sv + 1, sp = 1L << d, ic = sw + 1, sc = 1L << (d + 1);
    long t = 0;
    for (long m = ip; m <= n; m += sp) {
        ++t;                                         /* enumerate: t = 1,2,... */
        int i0 = (t % 2 == 1) ? 1 : 2;
        int i1 = F - i0 + 1;
        int i2 = ((t + 1) % 2 == 1) ? 1 : 2;
        long j = (sw == sv) ? wx_pmod(m - sp - 1, n) + 1 : wx_pmod(m - 1, n) + 1;   /* 1-based */
        long k1 = ((t - 1) >> 1) * sc + ic;
        long k2 = k1;
        T acc;
        if (add2out)
            acc = (T)(((double)v[(j - 1) * svs] + g[i1 - 1] * (double)w1[(k1 - 1) * s1]) +
                      h[i2 - 1] * (double)w2[(k2 - 1) * s2]);
        else
            acc = (T)(g[i1 - 1] * (double)w1[(k1 - 1) * s1] + h[i2 - 1] * (double)w2[(k2 - 1) * s2]);
        for (int i = i0 + 2; i <= F; i += 2) {
            i1 = F - i + 1;
            i2 = i + ((i & 1) ? 1 : -1);
            k1 = k1 - sc; if (k1 <= 0) k1 = wx_pmod(k1 - 1, n) + 1;
            k2 = k2 + sc; if (k2 > n)  k2 = wx_pmod(k2 - 1, n) + 1;
            acc = (T)((double)acc +
                      (g[i1 - 1] * (double)w1[(k1 - 1) * s1] + h[i2 - 1] * (double)w2[(k2 - 1) * s2]));
        }
        v[(j - 1) * svs] = acc;
    }
    return 0;
}

/* a10 isdwt_step! average based, 1-D   swt/swt_one_level.jl:257-277 */
WXO_API void SUF(isdwt_step_avg)(T *v, long svs, const T *w1, long s1, const T *w2, long s2, long n,
                                int d, const double *h, const double *g, int F)
{
    long nd = 1L << d;
    for (long sv = 0; sv < nd; ++sv) {
        long sw1 = sv, sw2 = sv + (1L << d);
        SUF(isdwt_step_shift)(v, svs, w1, s1, w2, s2, n, d, sv, sw1, h, g, F, 0);
        SUF(isdwt_step_shift)(v, svs, w1, s1, w2, s2, n, d, sv, sw2, h, g, F, 1);
    }
    for (long i = 0; i < n; ++i) v[i * svs] = (T)(v[i * svs] / (T)2);
}

/* a11 sdwt_step! 2-D   swt/swt_one_level.jl:334-370 ; all arrays (nr x nc), temp (nr x nc x 2) */
WXO_API void SUF(sdwt_step2)(T *w1, T *w2, T *w3, T *w4, const T *v, T *temp, long nr, long nc,
                            int d, const double *h, const double *g, int F)
{
    T *t1 = temp, *t2 = temp + nr * nc;
    for (long j = 0; j < nc; ++j)
        SUF(sdwt_step)(t1 + j * nr, 1, t2 + j * nr, 1, v + j * nr, 1, nr, d, h, g, F);
    for (long i = 0; i < nr; ++i) {
        SUF(sdwt_step)(w1 + i, nr, w2 + i, nr, t1 + i, nr, nc, d, h, g, F);
        SUF(sdwt_step)(w3 + i, nr, w4 + i, nr, t2 + i, nr, nc, d, h, g, F);
    }
}

/* a11 isdwt_step! 2-D average based   swt/swt_one_level.jl:395-431 */
WXO_API void SUF(isdwt_step2_avg)(T *v, const T *w1, const T *w2, const T *w3, const T *w4, T *temp,
                                 long nr, long nc, int d, const double *h, const double *g, int F)
{
    T *t1 = temp, *t2 = temp + nr * nc;
    for (long i = 0; i < nr; ++i) {
        SUF(isdwt_step_avg)(t1 + i, nr, w1 + i, nr, w2 + i, nr, nc, d, h, g, F);
        SUF(isdwt_step_avg)(t2 + i, nr, w3 + i, nr, w4 + i, nr, nc, d, h, g, F);
    }
    for (long j = 0; j < nc; ++j)
        SUF(isdwt_step_avg)(v + j * nr, 1, t1 + j * nr, 1, t2 + j * nr, 1, nr, d, h, g, F);
}

/* a11 isdwt_step! 2-D shift based   swt/swt_one_level.jl:433-469 */
WXO_API int SUF(isdwt_step2_shift)(T *v, const T *w1, const T *w2, const T *w3, const T *w4, T *temp,
                                  long nr, long nc, int d, long sv, long sw,
                                  const double *h, const double *g, int F)
{
    T *t1 = temp, *t2 = temp + nr * nc;
    int rc = 0;
    for (long i = 0; i < nr; ++i) {
        rc |= SUF(isdwt_step_shift)(t1 + i, nr, w1 + i, nr, w2 + i, nr, nc, d, sv, sw, h, g, F, 0);
        rc |= SUF(isdwt_step_shift)(t2 + i, nr, w3 + i, nr, w4 + i, nr, nc, d, sv, sw, h, g, F, 0);
    }
    for (long j = 0; j < nc; ++j)
        rc |= SUF(isdwt_step_shift)(v + j * nr, 1, t1 + j * nr, 1, t2 + j * nr, 1, nr, d, sv, sw, h, g, F, 0);
    return rc;
}

/* a16 acdwt_step! 1-D   acwt/acwt_one_level.jl:101-128 ; w1 uses g, w2 uses h, Lf taps each.
 * `2^d` in the reference is an integer power. */
WXO_API void SUF(acdwt_step)(T *w1, long s1, T *w2, long s2, const T *v, long sv, long n, int d,
                            const double *h, const double *g, int Lf)
{
    long D = 1L << d;
    for (long i1b = 1; i1b <= n; ++i1b) {
        long t = i1b + D; if (t > n) t = wx_pmod(t - 1, n) + 1;
        long io = wx_pmod(i1b + (long)(Lf / 2 + 1) * D - 1, n) + 1;
        T a1 = (T)(g[0] * (double)v[(t - 1) * sv]);
        T a2 = (T)(h[0] * (double)v[(t - 1) * sv]);
        for (int k = 2; k <= Lf; ++k) {
            t = t + D; if (t > n) t = wx_pmod(t - 1, n) + 1;
            a1 = (T)((double)a1 + g[k - 1] * (double)v[(t - 1) * sv]);
            a2 = (T)((double)a2 + h[k - 1] * (double)v[(t - 1) * sv]);
        }
        w1[(io - 1) * s1] = a1;
        w2[(io - 1) * s2] = a2;
    }
}

/* a17 iacdwt_step! 1-D   acwt/acwt_one_level.jl:217-224 ; Julia: (w1+w2)/sqrt(2) in T (irrational
 * constant rounds to T). */
WXO_API void SUF(iacdwt_step)(T *v, long sv, const T *w1, long s1, const T *w2, long s2, long n)
{
    const T r2 = (T)1.4142135623730951;
    for (long i = 0; i < n; ++i) v[i * sv] = (T)((T)(w1[i * s1] + w2[i * s2]) / r2);
}

/* acdwt_step! 2-D   acwt/acwt_one_level.jl:240-276 */
WXO_API void SUF(acdwt_step2)(T *w1, T *w2, T *w3, T *w4, const T *v, T *temp, long nr, long nc,
                             int d, const double *h, const double *g, int Lf)
{
    T *t1 = temp, *t2 = temp + nr * nc;
    for (long j = 0; j < nc; ++j)
        SUF(acdwt_step)(t1 + j * nr, 1, t2 + j * nr, 1, v + j * nr, 1, nr, d, h, g, Lf);
    for (long i = 0; i < nr; ++i) {
        SUF(acdwt_step)(w1 + i, nr, w2 + i, nr, t1 + i, nr, nc, d, h, g, Lf);
        SUF(acdwt_step)(w3 + i, nr, w4 + i, nr, t2 + i, nr, nc, d, h, g, Lf);
    }
}

/* iacdwt_step! 2-D   acwt/acwt_one_level.jl:288-322 */
WXO_API void SUF(iacdwt_step2)(T *v, const T *w1, const T *w2, const T *w3, const T *w4, T *temp,
                              long nr, long nc)
{
    T *t1 = temp, *t2 = temp + nr * nc;
    for (long i = 0; i < nr; ++i) {
        SUF(iacdwt_step)(t1 + i, nr, w1 + i, nr, w2 + i, nr, nc);
        SUF(iacdwt_step)(t2 + i, nr, w3 + i, nr, w4 + i, nr, nc);
    }
    for (long j = 0; j < nc; ++j)
        SUF(iacdwt_step)(v + j * nr, 1, t1 + j * nr, 1, t2 + j * nr, 1, nr);
}

/* ==========================================================================================
 * Tree drivers, decimated
 * ======================================================================================== */

/* a4  wpd! 1-D   DWT.jl:131-161 ; y is (n, L+1) */
WXO_API void SUF(wpd1)(T *y, const T *x, long n, int L, const double *h, const double *g, int F)
{
    memcpy(y, x, (size_t)n * sizeof(T));
    for (int i = 0; i < L; ++i) {
        long np = n >> i;
        for (long j = 0; j < (1L << i); ++j) {
            const T *v = y + (long)i * n + j * np;
            long nr = np / 2;
            T *w1 = y + (long)(i + 1) * n + 2 * j * nr;
            T *w2 = y + (long)(i + 1) * n + (2 * j + 1) * nr;
            SUF(dwt_step)(w1, 1, w2, 1, v, 1, np, h, g, F);
        }
    }
}

/* a5  wpd! 2-D   DWT.jl:164-209 ; x (m,n), y (m,n,L+1) */
WXO_API void SUF(wpd2)(T *y, const T *x, long m, long n, int L, const double *h, const double *g, int F)
{
    T *temp = (T *)malloc((size_t)m * n * sizeof(T));
    memcpy(y, x, (size_t)m * n * sizeof(T));
    for (int i = 0; i < L; ++i) {
        long mp = m >> i, np = n >> i, mr = mp / 2, nr = np / 2;
        for (long j = 0; j < (1L << i); ++j)
            for (long k = 0; k < (1L << i); ++k) {
                const T *sp = y + (long)i * m * n;
                T *sr = y + (long)(i + 1) * m * n;
                const T *v = sp + j * mp + k * np * m;
                T *w1 = sr + (2 * j) * mr + (2 * k) * nr * m;
                T *w2 = sr + (2 * j) * mr + (2 * k + 1) * nr * m;
                T *w3 = sr + (2 * j + 1) * mr + (2 * k) * nr * m;
                T *w4 = sr + (2 * j + 1) * mr + (2 * k + 1) * nr * m;
                T *tk = temp + j * mp + k * np * m;
                SUF(dwt_step2)(w1, w2, w3, w4, m, v, m, tk, m, mr, nr, h, g, F);
            }
    }
    free(temp);
}

/* wpt 1-D by tree: Wavelets.jl `wpt!` (third party, not in the reference tree; call sites
 * dwt/dwt_all.jl:162, pinned by test/transforms.jl:25-30 `wpd(x) ~ [x wpt1 wpt2 wpt3]`).
 * Restated as: apply a1 top-down to every node flagged in `tree`, children replace the
 * parent's range (natural order). tree has n-1 entries (1-D heap order). */
WXO_API void SUF(wpt1)(T *y, const T *x, long n, const unsigned char *tree, long ntree,
                      const double *h, const double *g, int F)
{
    T *tmp = (T *)malloc((size_t)n * sizeof(T));
    memcpy(y, x, (size_t)n * sizeof(T));
    for (long i = 1; i <= ntree; ++i) {
        if (!tree[i - 1]) continue;
        int d = wx_ilog2(i);
        long np = n >> d, j = i - (1L << d);
        if (np < 2) continue;
        memcpy(tmp, y + j * np, (size_t)np * sizeof(T));
        SUF(dwt_step)(y + j * np, 1, y + j * np + np / 2, 1, tmp, 1, np, h, g, F);
    }
    free(tmp);
}

/* iwpt 1-D by tree: Wavelets.jl `iwpt!` (third party; call sites dwt/dwt_all.jl:221, DWT.jl:349;
 * pinned by test/transforms.jl:31-33,296-299 round trips).  Restated as a2 applied bottom-up. */
WXO_API void SUF(iwpt1)(T *xh, const T *xw, long n, const unsigned char *tree, long ntree,
                       const double *h, const double *g, int F)
{
    T *tmp = (T *)malloc((size_t)n * sizeof(T));
    memcpy(xh, xw, (size_t)n * sizeof(T));
    for (long i = ntree; i >= 1; --i) {
        if (!tree[i - 1]) continue;
        int d = wx_ilog2(i);
        long np = n >> d, j = i - (1L << d);
        if (np < 2) continue;
        memcpy(tmp, xh + j * np, (size_t)np * sizeof(T));
        SUF(idwt_step)(xh + j * np, 1, tmp, 1, tmp + np / 2, 1, np, h, g, F);
    }
    free(tmp);
}

/* quad-tree node -> (row0,col0,rows,cols), 0-based; Utils.jl:465-542 (getrowrange/getcolrange):
 * children 4i-2 (TL) 4i-1 (TR) 4i (BL) 4i+1 (BR); rows: idx<4*parent -> upper half; cols: even -> left */

/* 2-D wpt! by tree   DWT.jl:500-548 */
WXO_API void SUF(wpt2)(T *y, const T *x, long m, long n, const unsigned char *tree, long ntree,
                      const double *h, const double *g, int F)
{
    size_t bytes = (size_t)m * n * sizeof(T);
    T *yt = (T *)malloc(bytes), *temp = (T *)malloc(bytes);
    memcpy(y, x, bytes); memcpy(yt, x, bytes);
    for (long i = 1; i <= ntree; ++i) {
        if (!tree[i - 1]) continue;
        long r0, c0, nr, nc;
        wx_quadrange(m, n, i, &r0, &c0, &nr, &nc);
        long hr = nr / 2, hc = nc / 2;
        const T *v = yt + r0 + c0 * m;
        T *w1 = y + r0 + c0 * m, *w2 = y + r0 + (c0 + hc) * m;
        T *w3 = y + r0 + hr + c0 * m, *w4 = y + r0 + hr + (c0 + hc) * m;
        SUF(dwt_step2)(w1, w2, w3, w4, m, v, m, temp + r0 + c0 * m, m, hr, hc, h, g, F);
        if (4 * i < ntree)
            for (long c = 0; c < nc; ++c)
                memcpy(yt + r0 + (c0 + c) * m, y + r0 + (c0 + c) * m, (size_t)nr * sizeof(T));
    }
    free(yt); free(temp);
}

/* 2-D iwpt! by tree   DWT.jl:662-710 */
WXO_API void SUF(iwpt2)(T *xh, const T *xw, long m, long n, const unsigned char *tree, long ntree,
                       const double *h, const double *g, int F)
{
    size_t bytes = (size_t)m * n * sizeof(T);
    T *xt = (T *)malloc(bytes), *temp = (T *)malloc(bytes);
    memcpy(xh, xw, bytes); memcpy(xt, xw, bytes);
    for (long i = ntree; i >= 1; --i) {
        if (!tree[i - 1]) continue;
        long r0, c0, nr, nc;
        wx_quadrange(m, n, i, &r0, &c0, &nr, &nc);
        long hr = nr / 2, hc = nc / 2;
        T *v = xh + r0 + c0 * m;
        const T *w1 = xt + r0 + c0 * m, *w2 = xt + r0 + (c0 + hc) * m;
        const T *w3 = xt + r0 + hr + c0 * m, *w4 = xt + r0 + hr + (c0 + hc) * m;
        SUF(idwt_step2)(v, m, w1, w2, w3, w4, m, temp + r0 + c0 * m, m, hr, hc, h, g, F);
        if (i > 1)
            for (long c = 0; c < nc; ++c)
                memcpy(xt + r0 + (c0 + c) * m, xh + r0 + (c0 + c) * m, (size_t)nr * sizeof(T));
    }
    free(xt); free(temp);
}

/* a24 getbasiscoef 1-D   Utils.jl:101-134 with getleaf utils_tree.jl:122-157.
 * Xw (n,K); leaf i at depth d -> rows [nn*n0, (nn+1)*n0) of column d. returns -2 if a leaf is deeper
 * than the table (ArgumentError in the reference). */
WXO_API int SUF(getbasiscoef1)(T *xw, const T *Xw, long n, int K, const unsigned char *tree, long ntree)
{
    long nleaf = 2 * ntree + 1;
    unsigned char *leaf = (unsigned char *)calloc((size_t)nleaf + 1, 1);
    wx_getleaf_binary(tree, ntree, leaf);
    int rc = 0;
    for (long i = 1; i <= nleaf; ++i) {
        if (!leaf[i - 1]) continue;
        int d = wx_ilog2(i);
        if (d >= K) { rc = -2; break; }
        long nn = i - (1L << d), n0 = n >> d;
        memcpy(xw + nn * n0, Xw + (long)d * n + nn * n0, (size_t)n0 * sizeof(T));
    }
    free(leaf);
    return rc;
}

/* getbasiscoef 2-D  Utils.jl:101-134 (N==3 branch); Xw (m,n,K) */
WXO_API int SUF(getbasiscoef2)(T *xw, const T *Xw, long m, long n, int K, const unsigned char *tree, long ntree)
{
    long nleaf = 4 * ntree + 1;
    unsigned char *leaf = (unsigned char *)calloc((size_t)nleaf + 1, 1);
    wx_getleaf_quad(tree, ntree, leaf);
    int rc = 0;
    for (long i = 1; i <= nleaf; ++i) {
        if (!leaf[i - 1]) continue;
        int d = wx_quaddepth(i);
        if (d >= K) { rc = -2; break; }
        long r0, c0, nr, nc;
        wx_quadrange(m, n, i, &r0, &c0, &nr, &nc);
        for (long c = 0; c < nc; ++c)
            memcpy(xw + r0 + (c0 + c) * m, Xw + (long)d * m * n + r0 + (c0 + c) * m, (size_t)nr * sizeof(T));
    }
    free(leaf);
    return rc;
}

/* a8  iwpd! 1-D by tree   DWT.jl:337-351 : getbasiscoef + iwpt */
WXO_API int SUF(iwpd1)(T *xh, const T *Xw, long n, int K, const unsigned char *tree, long ntree,
                      const double *h, const double *g, int F)
{
    T *w = (T *)malloc((size_t)n * sizeof(T));
    int rc = SUF(getbasiscoef1)(w, Xw, n, K, tree, ntree);
    if (rc == 0) SUF(iwpt1)(xh, w, n, tree, ntree, h, g, F);
    free(w);
    return rc;
}

/* a8  iwpd! 2-D by tree   DWT.jl:354-401 ; Xw (m,n,K) */
WXO_API void SUF(iwpd2)(T *xh, const T *Xw, long m, long n, int K, const unsigned char *tree, long ntree,
                       const double *h, const double *g, int F)
{
    size_t bytes = (size_t)m * n * sizeof(T);
    T *xt = (T *)malloc(bytes * (size_t)K), *temp = (T *)malloc(bytes);
    memcpy(xt, Xw, bytes * (size_t)K);
    /* a root that is not split: the reference never writes x-hat in that case (DWT.jl:366-399 only stores inside the loop);
     * defined here, and in the CUDA path, as the level-0 slice (what getbasiscoef returns for that tree) */
    if (ntree < 1 || !tree[0]) memcpy(xh, Xw, bytes);
    for (long i = ntree; i >= 1; --i) {
        if (!tree[i - 1]) continue;
        int d = wx_quaddepth(i);
        long r0, c0, nr, nc;
        wx_quadrange(m, n, i, &r0, &c0, &nr, &nc);
        long hr = nr / 2, hc = nc / 2;
        T *v = (d == 0) ? xh : xt + (long)d * m * n + r0 + c0 * m;
        const T *sl = xt + (long)(d + 1) * m * n;
        const T *w1 = sl + r0 + c0 * m, *w2 = sl + r0 + (c0 + hc) * m;
        const T *w3 = sl + r0 + hr + c0 * m, *w4 = sl + r0 + hr + (c0 + hc) * m;
        SUF(idwt_step2)(v, m, w1, w2, w3, w4, m, temp + r0 + c0 * m, m, hr, hc, h, g, F);
    }
    free(xt); free(temp);
}

/* ==========================================================================================
 * Tree drivers, stationary (SWT.jl) and autocorrelation (ACWT.jl); ac != 0 selects the AC step.
 * For ac the pair (h,g) is (Q,P) with Lf taps, exactly how ACWT.jl:120-131,756 passes them.
 * ======================================================================================== */
WXO_API void SUF(rstep)(int ac, T *w1, T *w2, const T *v, long n, int d, const double *h, const double *g, int F)
{
    if (ac) SUF(acdwt_step)(w1, 1, w2, 1, v, 1, n, d, h, g, F);
    else    SUF(sdwt_step)(w1, 1, w2, 1, v, 1, n, d, h, g, F);
}
WXO_API void SUF(rstep2)(int ac, T *w1, T *w2, T *w3, T *w4, const T *v, T *temp, long nr, long nc, int d,
                        const double *h, const double *g, int F)
{
    if (ac) SUF(acdwt_step2)(w1, w2, w3, w4, v, temp, nr, nc, d, h, g, F);
    else    SUF(sdwt_step2)(w1, w2, w3, w4, v, temp, nr, nc, d, h, g, F);
}

/* a14 sdwt! 1-D SWT.jl:109-131 / acdwt! ACWT.jl:109-131 ; xw (n, L+1) */
WXO_API void SUF(rdwt1)(int ac, T *xw, const T *x, long n, int L, const double *h, const double *g, int F)
{
    T *v = (T *)malloc((size_t)n * sizeof(T));
    memcpy(xw + (long)L * n, x, (size_t)n * sizeof(T));
    for (int d = 0; d < L; ++d) {
        memcpy(v, xw + (long)(L - d) * n, (size_t)n * sizeof(T));
        SUF(rstep)(ac, xw + (long)(L - d - 1) * n, xw + (long)(L - d) * n, v, n, d, h, g, F);
    }
    free(v);
}

/* sdwt! 2-D SWT.jl:133-158 / acdwt! 2-D ACWT.jl:133-157 ; x (nr,nc), xw (nr,nc,3L+1) */
WXO_API void SUF(rdwt2)(int ac, T *xw, const T *x, long nr, long nc, int L, const double *h, const double *g, int F)
{
    long sz = nr * nc;
    T *v = (T *)malloc((size_t)sz * sizeof(T)), *temp = (T *)malloc((size_t)2 * sz * sizeof(T));
    memcpy(xw + (long)(3 * L) * sz, x, (size_t)sz * sizeof(T));
    for (int d = 0; d < L; ++d) {
        long b = 3L * (L - d);      /* Julia slice 3(L-d)+1 -> 0-based b */
        memcpy(v, xw + b * sz, (size_t)sz * sizeof(T));
        SUF(rstep2)(ac, xw + (b - 3) * sz, xw + (b - 2) * sz, xw + (b - 1) * sz, xw + b * sz, v, temp, nr, nc, d, h, g, F);
    }
    free(v); free(temp);
}

/* swpt! 1-D SWT.jl:439-471 / acwpt! ACWT.jl:427-460 ; xw (n, 2^L) */
WXO_API void SUF(rwpt1)(int ac, T *xw, const T *x, long n, int L, const double *h, const double *g, int F)
{
    T *v = (T *)malloc((size_t)n * sizeof(T));
    memcpy(xw, x, (size_t)n * sizeof(T));
    for (int d = 0; d < L; ++d) {
        long nn = 1L << d;
        for (long b = 0; b < nn; ++b) {
            long np = (1L << L) / nn, ncld = np / 2;
            long j1 = (2 * b) * ncld, j2 = (2 * b + 1) * ncld;
            memcpy(v, xw + j1 * n, (size_t)n * sizeof(T));
            SUF(rstep)(ac, xw + j1 * n, xw + j2 * n, v, n, d, h, g, F);
        }
    }
    free(v);
}

/* swpt! 2-D SWT.jl:472-513 / acwpt! 2-D ACWT.jl:462-501 ; xw (nr,nc,4^L) */
WXO_API void SUF(rwpt2)(int ac, T *xw, const T *x, long nr, long nc, int L, const double *h, const double *g, int F)
{
    long sz = nr * nc;
    T *v = (T *)malloc((size_t)sz * sizeof(T)), *temp = (T *)malloc((size_t)2 * sz * sizeof(T));
    memcpy(xw, x, (size_t)sz * sizeof(T));
    long tot = 1L << (2 * L);
    for (int d = 0; d < L; ++d) {
        long nn = 1L << (2 * d);
        for (long b = 0; b < nn; ++b) {
            long np = tot / nn, ncld = np / 4;
            long j1 = (4 * b) * ncld, j2 = (4 * b + 1) * ncld, j3 = (4 * b + 2) * ncld, j4 = (4 * b + 3) * ncld;
            memcpy(v, xw + j1 * sz, (size_t)sz * sizeof(T));
            SUF(rstep2)(ac, xw + j1 * sz, xw + j2 * sz, xw + j3 * sz, xw + j4 * sz, v, temp, nr, nc, d, h, g, F);
        }
    }
    free(v); free(temp);
}

/* a12 swpd! 1-D SWT.jl:840-868 / a18 acwpd! ACWT.jl:733-759 ; xw (n, 2^(L+1)-1) heap order */
WXO_API void SUF(rwpd1)(int ac, T *xw, const T *x, long n, int L, const double *h, const double *g, int F)
{
    long n0 = (1L << (L + 1)) - 1, n1 = n0 - (1L << L);
    memcpy(xw, x, (size_t)n * sizeof(T));
    for (long i = 1; i <= n1; ++i) {
        int d = wx_ilog2(i);
        SUF(rstep)(ac, xw + (2 * i - 1) * n, xw + (2 * i) * n, xw + (i - 1) * n, n, d, h, g, F);
    }
}

/* swpd! 2-D SWT.jl:870-902 / acwpd! 2-D ACWT.jl:761-793 ; xw (nr,nc,sum 4^(0:L)) heap order */
WXO_API void SUF(rwpd2)(int ac, T *xw, const T *x, long nr, long nc, int L, const double *h, const double *g, int F)
{
    long sz = nr * nc;
    long k = ((1L << (2 * L + 2)) - 1) / 3, n1 = k - (1L << (2 * L));
    T *temp = (T *)malloc((size_t)2 * sz * sizeof(T));
    memcpy(xw, x, (size_t)sz * sizeof(T));
    for (long i = 1; i <= n1; ++i) {
        int d = wx_quaddepth(i);
        SUF(rstep2)(ac, xw + (4 * i - 3) * sz, xw + (4 * i - 2) * sz, xw + (4 * i - 1) * sz, xw + (4 * i) * sz,
                    xw + (i - 1) * sz, temp, nr, nc, d, h, g, F);
    }
    free(temp);
}

/* ---- inverses.  mode: 0 = average based (SWT) ; 1 = shift based with sm (SWT) ; 2 = AC ---- */
WXO_API int SUF(ristep)(int mode, T *v, const T *w1, const T *w2, long n, int d, long sv, long sw,
                       const double *h, const double *g, int F)
{
    if (mode == 2) { SUF(iacdwt_step)(v, 1, w1, 1, w2, 1, n); return 0; }
    if (mode == 1) return SUF(isdwt_step_shift)(v, 1, w1, 1, w2, 1, n, d, sv, sw, h, g, F, 0);
    SUF(isdwt_step_avg)(v, 1, w1, 1, w2, 1, n, d, h, g, F);
    return 0;
}
WXO_API int SUF(ristep2)(int mode, T *v, const T *w1, const T *w2, const T *w3, const T *w4, T *temp,
                        long nr, long nc, int d, long sv, long sw, const double *h, const double *g, int F)
{
    if (mode == 2) { SUF(iacdwt_step2)(v, w1, w2, w3, w4, temp, nr, nc); return 0; }
    if (mode == 1) return SUF(isdwt_step2_shift)(v, w1, w2, w3, w4, temp, nr, nc, d, sv, sw, h, g, F);
    SUF(isdwt_step2_avg)(v, w1, w2, w3, w4, temp, nr, nc, d, h, g, F);
    return 0;
}

/* isdwt! 1-D SWT.jl:259-283 (shift), :311-328 (average) ; iacdwt! ACWT.jl:287-304. sd = main2depthshift */
WXO_API int SUF(irdwt1)(int mode, T *x, const T *xw, long n, int L, const long *sd,
                       const double *h, const double *g, int F)
{
    T *w1 = (T *)malloc((size_t)n * sizeof(T));
    int rc = 0;
    memcpy(x, xw, (size_t)n * sizeof(T));
    for (int d = L - 1; d >= 0; --d) {
        memcpy(w1, x, (size_t)n * sizeof(T));
        rc |= SUF(ristep)(mode, x, w1, xw + (long)(L - d) * n, n, d, sd ? sd[d] : 0, sd ? sd[d + 1] : 0, h, g, F);
    }
    free(w1);
    return rc;
}

/* isdwt! 2-D SWT.jl:284-309 (shift), :329-358 (average) ; iacdwt! 2-D ACWT.jl:305-329 */
WXO_API int SUF(irdwt2)(int mode, T *x, const T *xw, long nr, long nc, int L, const long *sd,
                       const double *h, const double *g, int F)
{
    long sz = nr * nc;
    T *w1 = (T *)malloc((size_t)sz * sizeof(T)), *temp = (T *)malloc((size_t)2 * sz * sizeof(T));
    int rc = 0;
    memcpy(x, xw, (size_t)sz * sizeof(T));
    for (int d = L - 1; d >= 0; --d) {
        long b = 3L * (L - d);
        memcpy(w1, x, (size_t)sz * sizeof(T));
        rc |= SUF(ristep2)(mode, x, w1, xw + (b - 2) * sz, xw + (b - 1) * sz, xw + b * sz, temp, nr, nc, d,
                           sd ? sd[d] : 0, sd ? sd[d + 1] : 0, h, g, F);
    }
    free(w1); free(temp);
    return rc;
}

/* iswpt! 1-D SWT.jl:613-646 (shift), :687-716 (average) ; iacwpt! ACWT.jl:581-608 ; xw (n,2^L) */
WXO_API int SUF(irwpt1)(int mode, T *x, const T *xw, long n, int L, const long *sd,
                       const double *h, const double *g, int F)
{
    long cols = 1L << L;
    T *tmp = (T *)malloc((size_t)n * cols * sizeof(T)), *w1 = (T *)malloc((size_t)n * sizeof(T));
    int rc = 0;
    memcpy(tmp, xw, (size_t)n * cols * sizeof(T));
    for (int d = L - 1; d >= 0; --d) {
        long nn = 1L << d;
        for (long b = 0; b < nn; ++b) {
            long np = cols / nn, ncld = np / 2;
            long j1 = (2 * b) * ncld, j2 = (2 * b + 1) * ncld;
            T *v = (d == 0) ? x : tmp + j1 * n;
            memcpy(w1, tmp + j1 * n, (size_t)n * sizeof(T));
            rc |= SUF(ristep)(mode, v, w1, tmp + j2 * n, n, d, sd ? sd[d] : 0, sd ? sd[d + 1] : 0, h, g, F);
        }
    }
    free(tmp); free(w1);
    return rc;
}

/* iswpt! 2-D SWT.jl:648-685 (shift), :718-758 (average) ; iacwpt! 2-D ACWT.jl:610-648 */
WXO_API int SUF(irwpt2)(int mode, T *x, const T *xw, long nr, long nc, int L, const long *sd,
                       const double *h, const double *g, int F)
{
    long sz = nr * nc, tot = 1L << (2 * L);
    T *xt = (T *)malloc((size_t)sz * tot * sizeof(T)), *w1 = (T *)malloc((size_t)sz * sizeof(T));
    T *temp = (T *)malloc((size_t)2 * sz * sizeof(T));
    int rc = 0;
    memcpy(xt, xw, (size_t)sz * tot * sizeof(T));
    for (int d = L - 1; d >= 0; --d) {
        long nn = 1L << (2 * d);
        for (long b = 0; b < nn; ++b) {
            long np = tot / nn, ncld = np / 4;
            long j1 = (4 * b) * ncld, j2 = (4 * b + 1) * ncld, j3 = (4 * b + 2) * ncld, j4 = (4 * b + 3) * ncld;
            T *v = (d == 0) ? x : xt + j1 * sz;
            memcpy(w1, xt + j1 * sz, (size_t)sz * sizeof(T));
            rc |= SUF(ristep2)(mode, v, w1, xt + j2 * sz, xt + j3 * sz, xt + j4 * sz, temp, nr, nc, d,
                               sd ? sd[d] : 0, sd ? sd[d + 1] : 0, h, g, F);
        }
    }
    free(xt); free(w1); free(temp);
    return rc;
}

/* iswpd! 1-D by tree SWT.jl:1064-1095 (shift), :1131-1157 (average) ; iacwpd! ACWT.jl:946-970.
 * xw (n, ncols) heap order; tree has ntree entries; every flagged node needs 2i+1 <= ncols. */
WXO_API int SUF(irwpd1)(int mode, T *x, const T *xw, long n, long ncols, const unsigned char *tree, long ntree,
                       const long *sd, const double *h, const double *g, int F)
{
    T *tmp = (T *)malloc((size_t)n * ncols * sizeof(T));
    int rc = 0;
    memcpy(tmp, xw, (size_t)n * ncols * sizeof(T));
    if (ntree == 0 || !tree[0]) memcpy(x, xw, (size_t)n * sizeof(T));   /* root is a leaf: identity */
    for (long i = ntree; i >= 1; --i) {
        if (!tree[i - 1]) continue;
        if (2 * i + 1 > ncols) { rc = -2; break; }
        int d = wx_ilog2(i);
        T *v = (i == 1) ? x : tmp + (i - 1) * n;
        rc |= SUF(ristep)(mode, v, tmp + (2 * i - 1) * n, tmp + (2 * i) * n, n, d,
                          sd ? sd[d] : 0, sd ? sd[d + 1] : 0, h, g, F);
    }
    free(tmp);
    return rc;
}

/* iswpd! 2-D by tree SWT.jl:1096-1129 (shift), :1158-1199 (average) ; iacwpd! 2-D ACWT.jl:972-1000 */
WXO_API int SUF(irwpd2)(int mode, T *x, const T *xw, long nr, long nc, long nsl, const unsigned char *tree,
                       long ntree, const long *sd, const double *h, const double *g, int F)
{
    long sz = nr * nc;
    T *xt = (T *)malloc((size_t)sz * nsl * sizeof(T)), *temp = (T *)malloc((size_t)2 * sz * sizeof(T));
    int rc = 0;
    memcpy(xt, xw, (size_t)sz * nsl * sizeof(T));
    if (ntree == 0 || !tree[0]) memcpy(x, xw, (size_t)sz * sizeof(T));
    for (long i = ntree; i >= 1; --i) {
        if (!tree[i - 1]) continue;
        if (4 * i + 1 > nsl) { rc = -2; break; }
        int d = wx_quaddepth(i);
        T *v = (i == 1) ? x : xt + (i - 1) * sz;
        rc |= SUF(ristep2)(mode, v, xt + (4 * i - 3) * sz, xt + (4 * i - 2) * sz, xt + (4 * i - 1) * sz,
                           xt + (4 * i) * sz, temp, nr, nc, d, sd ? sd[d] : 0, sd ? sd[d + 1] : 0, h, g, F);
    }
    free(xt); free(temp);
    return rc;
}

/* ==========================================================================================
 * Best basis:  a19 tree_costs(JBB)  bestbasis/bestbasis_tree.jl:150-207 with
 *              a20 coefcost LoglpCost / NormCost  bestbasis/bestbasis_costs.jl:127-132
 * X is (sz, K, N) with sz = n (1-D) or nr*nc (2-D).  Moments are accumulated sequentially over the
 * batch (Julia sum(dims=3) order) in T, then sigma = (EX2 - EX^2)^0.5 in T.
 * cost_kind: 0 = Loglp (p * sum log|s|), 1 = Norm (norm(s,p)^p).
 * returns -1 if any sigma is NaN/negative (the reference's DomainError / @assert).
 * ======================================================================================== */
WXO_API T SUF(coefcost_jbb)(const T *s, long r0, long c0, long nr, long nc, long ld, int kind, double p)
{
    /* the reference sums with Julia's pairwise `sum`; the order used here is plain sequential in
     * Float64 -- costs are compared at 1e-12 relative, trees on tie-free data. */
    double acc = 0.0;
    for (long c = 0; c < nc; ++c)
        for (long r = 0; r < nr; ++r) {
            double a = fabs((double)s[r0 + r + (c0 + c) * ld]);
            if (kind == 0) acc += (double)(T)log((double)(T)a);
            else acc += pow(a, p);
        }
    if (kind == 0) return (T)(p * acc);
    return (T)acc;      /* norm(x,p)^p == sum |x|^p */
}

WXO_API int SUF(jbb_sigma)(T *sigma, const T *X, long sz, long K, long N)
{
    long tot = sz * K;
    T *ex = (T *)calloc((size_t)tot, sizeof(T)), *ex2 = (T *)calloc((size_t)tot, sizeof(T));
    for (long k = 0; k < N; ++k)
        for (long e = 0; e < tot; ++e) {
            T xv = X[k * tot + e];
            ex[e] = (T)(ex[e] + xv);
            ex2[e] = (T)(ex2[e] + (T)(xv * xv));
        }
    int rc = 0;
    for (long e = 0; e < tot; ++e) {
        T m = (T)(ex[e] / (T)N), m2 = (T)(ex2[e] / (T)N);
        T var = (T)(m2 - (T)(m * m));
        if (!(var >= 0)) rc = -1;
        sigma[e] = (T)sqrt((double)var);
    }
    free(ex); free(ex2);
    return rc;
}

WXO_API int SUF(tree_costs_jbb1)(T *costs, const T *X, long n, long K, long N, int redundant, int kind, double p)
{
    T *sigma = (T *)malloc((size_t)n * K * sizeof(T));
    int rc = SUF(jbb_sigma)(sigma, X, n, K, N);
    if (redundant) {
        for (long i = 1; i <= K; ++i) {
            int j = wx_ilog2(i);
            costs[i - 1] = (T)(SUF(coefcost_jbb)(sigma + (i - 1) * n, 0, 0, n, 1, n, kind, p) / (T)(1L << j));
        }
    } else {
        long i = 0;
        for (long lvl = 0; lvl < K; ++lvl) {
            long n0 = n >> lvl;
            for (long node = 0; node < (1L << lvl); ++node)
                costs[i++] = SUF(coefcost_jbb)(sigma + lvl * n, node * n0, 0, n0, 1, n, kind, p);
        }
    }
    free(sigma);
    return rc;
}

WXO_API int SUF(tree_costs_jbb2)(T *costs, const T *X, long nr, long nc, long K, long N, int redundant, int kind, double p)
{
    long sz = nr * nc;
    T *sigma = (T *)malloc((size_t)sz * K * sizeof(T));
    int rc = SUF(jbb_sigma)(sigma, X, sz, K, N);
    if (redundant) {
        for (long i = 1; i <= K; ++i) {
            int d = wx_quaddepth(i);
            costs[i - 1] = (T)(SUF(coefcost_jbb)(sigma + (i - 1) * sz, 0, 0, nr, nc, nr, kind, p) / (T)(1L << (2 * d)));
        }
    } else {
        long ncost = ((1L << (2 * K)) - 1) / 3;
        for (long i = 1; i <= ncost; ++i) {
            int d = wx_quaddepth(i);
            long r0, c0, rr, cc;
            wx_quadrange(nr, nc, i, &r0, &c0, &rr, &cc);
            costs[i - 1] = SUF(coefcost_jbb)(sigma + (long)d * sz, r0, c0, rr, cc, nr, kind, p);
        }
    }
    free(sigma);
    return rc;
}

/* a21 coefcost(::DifferentialEntropyCost) for one coefficient position  bestbasis/bestbasis_costs.jl:135-155
 * PARITY UNPINNED: relies on AverageShiftedHistograms.jl (0.8/0.9, not vendored) `ash`/`pdf` with the
 * triangular kernel, restated from its published algorithm: bin k = floor((x-a)/delta + 1.5) (1-based),
 * density y[i] = sum_k c[k]*K((i-k)/m) for |i-k| < m, normalised by 1/(sum(y)*delta), pdf = linear
 * interpolation between grid points (0 outside).  x has stride `st`, N samples.  Always Float64 (`ent`). */
WXO_API double SUF(diffentropy)(const T *x, long st, long N)
{
    long nbins = (long)ceil(pow(30.0 * (double)N, 0.2));
    long mbins = (long)ceil(50.0 / (double)nbins);
    double mean = 0.0, mn = (double)x[0], mx = (double)x[0];
    for (long k = 0; k < N; ++k) { double v = (double)x[k * st]; mean += v; if (v < mn) mn = v; if (v > mx) mx = v; }
    mean /= (double)N;
    double ss = 0.0;
    for (long k = 0; k < N; ++k) { double dv = (double)x[k * st] - mean; ss += dv * dv; }
    double sigma = (N > 1) ? sqrt(ss / (double)(N - 1)) : NAN;    /* Statistics.std: corrected */
    sigma = (double)(T)sigma;
    double s = 0.5;
    long npts = (nbins + 1) * mbins;
    double delta = ((double)(T)(mx - mn) + sigma) / (double)(npts - 1);
    delta = (double)(T)delta;
    double a = (double)(T)(mn - s * sigma);
    /* a constant sample gives a zero step: the reference's range construction `a:0.0:b` throws ArgumentError
     * (bestbasis_costs.jl:146); the restatement reports NaN and its Python wrapper raises */
    if (!(delta > 0.0) || !isfinite(delta)) return NAN;
    double *cnt = (double *)calloc((size_t)npts, sizeof(double));
    double *y = (double *)calloc((size_t)npts, sizeof(double));
    double dinv = 1.0 / delta;
    for (long k = 0; k < N; ++k) {
        long ki = (long)floor(((double)x[k * st] - a) * dinv + 1.5);
        if (ki >= 1 && ki <= npts) cnt[ki - 1] += 1.0;
    }
    for (long k = 1; k <= npts; ++k) {
        if (cnt[k - 1] == 0.0) continue;
        long lo = k - mbins + 1 < 1 ? 1 : k - mbins + 1, hi = k + mbins - 1 > npts ? npts : k + mbins - 1;
        for (long i = lo; i <= hi; ++i) {
            double u = fabs((double)(i - k) / (double)mbins);
            y[i - 1] += cnt[k - 1] * (u <= 1.0 ? 1.0 - u : 0.0);
        }
    }
    double tot = 0.0;
    for (long i = 0; i < npts; ++i) tot += y[i];
    double den = 1.0 / (tot * delta);
    for (long i = 0; i < npts; ++i) y[i] *= den;
    double ent = 0.0;
    for (long k = 0; k < N; ++k) {
        double xv = (double)x[k * st];
        long i = (long)floor((xv - a) * dinv) + 1;       /* searchsortedlast on the grid, 1-based */
        while (i >= 1 && i <= npts && a + (double)(i - 1) * delta > xv) --i;
        while (i + 1 <= npts && a + (double)i * delta <= xv) ++i;
        double pdf = 0.0;
        if (i >= 1 && i < npts) {
            double g0 = a + (double)(i - 1) * delta, g1 = a + (double)i * delta;
            pdf = y[i - 1] + (y[i] - y[i - 1]) * (xv - g0) / (g1 - g0);
        }
        ent -= (1.0 / (double)N) * log(pdf);
    }
    free(cnt); free(y);
    return ent;
}

/* a21 tree_costs(LSDB) 1-D  bestbasis/bestbasis_tree.jl:104-124 ; X (n,K,N) */
WXO_API void SUF(tree_costs_lsdb1)(T *costs, const T *X, long n, long K, long N, int redundant)
{
    long tot = n * K;
    if (redundant) {
        for (long i = 1; i <= K; ++i) {
            int j = wx_ilog2(i);
            double c = 0.0;
            for (long r = 0; r < n; ++r) c += SUF(diffentropy)(X + (i - 1) * n + r, tot, N);
            costs[i - 1] = (T)(c / (double)(1L << j));
        }
    } else {
        long i = 0;
        for (long d = 0; d < K; ++d) {
            long n0 = n >> d;
            for (long node = 0; node < (1L << d); ++node) {
                double c = 0.0;
                for (long r = 0; r < n0; ++r) c += SUF(diffentropy)(X + d * n + node * n0 + r, tot, N);
                costs[i++] = (T)c;
            }
        }
    }
}

/* a21 tree_costs(LSDB) 2-D  bestbasis/bestbasis_tree.jl:126-147 ; X (nr,nc,K,N) */
WXO_API void SUF(tree_costs_lsdb2)(T *costs, const T *X, long nr, long nc, long K, long N, int redundant)
{
    long sz = nr * nc, tot = sz * K;
    if (redundant) {
        for (long i = 1; i <= K; ++i) {
            int d = wx_quaddepth(i);
            double c = 0.0;
            for (long e = 0; e < sz; ++e) c += SUF(diffentropy)(X + (i - 1) * sz + e, tot, N);
            costs[i - 1] = (T)(c / (double)(1L << (2 * d)));
        }
    } else {
        long ncost = ((1L << (2 * K)) - 1) / 3;
        for (long i = 1; i <= ncost; ++i) {
            int d = wx_quaddepth(i);
            long r0, c0, rr, cc;
            wx_quadrange(nr, nc, i, &r0, &c0, &rr, &cc);
            double c = 0.0;
            for (long q = 0; q < cc; ++q)
                for (long r = 0; r < rr; ++r)
                    c += SUF(diffentropy)(X + (long)d * sz + r0 + r + (c0 + q) * nr, tot, N);
            costs[i - 1] = (T)c;
        }
    }
}

/* f-1 coefcost(x, ::ShannonEntropyCost / ::LogEnergyEntropyCost, nrm)  bestbasis/bestbasis_costs.jl:103-125
 * kind 0: s = (x/nrm)^2, -s*log(s) (0 when s == 0); kind 1: -log(s) (0 when s == 0); nrm == 0 -> 0.  Summed in index
 * order in the element type.  PARITY UNPINNED by the reference's tests (test/bestbasis.jl:13-21 only check isvalidtree);
 * the arithmetic is fully visible in the reference and restated 1:1. */
static T SUF(coefcost_bb)(const T *x, long r0, long c0, long rr, long cc, long ld, int kind, T nrm)
{
    T sum = (T)0;
    if (nrm == sum) return sum;
    for (long c = 0; c < cc; ++c)
        for (long r = 0; r < rr; ++r) {
            T q = x[(c0 + c) * ld + r0 + r] / nrm;
            T s = q * q;
            T v = (s == (T)0) ? (T)(-0.0) : (kind == 0 ? (T)(-s * (T)log((double)s)) : (T)(-(T)log((double)s)));
            sum += v;
        }
    return sum;
}
static T SUF(norm2)(const T *x, long n)
{
    double a = 0.0;
    for (long i = 0; i < n; ++i) a += (double)x[i] * (double)x[i];
    return (T)sqrt(a);
}

/* f-1 tree_costs(X, ::BB) 1-D  bestbasis/bestbasis_tree.jl:210-233 ; X (n, K) one signal */
WXO_API void SUF(tree_costs_bb1)(T *costs, const T *X, long n, long K, int redundant, int kind)
{
    T nrm = SUF(norm2)(X, n);                         /* norm(X[:,1]) */
    if (redundant) {
        for (long i = 1; i <= K; ++i) {
            int j = wx_ilog2(i);
            costs[i - 1] = (T)(SUF(coefcost_bb)(X + (i - 1) * n, 0, 0, n, 1, n, kind, nrm) / (T)(1L << j));
        }
    } else {
        long i = 0;
        for (long d = 0; d < K; ++d) {
            long n0 = n >> d;
            for (long node = 0; node < (1L << d); ++node)
                costs[i++] = SUF(coefcost_bb)(X + d * n, node * n0, 0, n0, 1, n, kind, nrm);
        }
    }
}

/* f-1 tree_costs(X, ::BB) 2-D  bestbasis/bestbasis_tree.jl:234-256 ; X (nr, nc, K) one image.  The non-redundant branch
 * calls coefcost WITHOUT nrm (:252), i.e. each node is normalised by its own norm -- restated as written. */
WXO_API void SUF(tree_costs_bb2)(T *costs, const T *X, long nr, long nc, long K, int redundant, int kind)
{
    long sz = nr * nc;
    T nrm = SUF(norm2)(X, sz);                        /* norm(X[:,:,1]) */
    if (redundant) {
        for (long i = 1; i <= K; ++i) {
            int d = wx_quaddepth(i);
            costs[i - 1] = (T)(SUF(coefcost_bb)(X + (i - 1) * sz, 0, 0, nr, nc, nr, kind, nrm) / (T)(1L << (2 * d)));
        }
    } else {
        long ncost = ((1L << (2 * K)) - 1) / 3;
        for (long i = 1; i <= ncost; ++i) {
            int d = wx_quaddepth(i);
            long r0, c0, rr, cc;
            wx_quadrange(nr, nc, i, &r0, &c0, &rr, &cc);
            double a = 0.0;
            for (long c = 0; c < cc; ++c)
                for (long r = 0; r < rr; ++r) { double v = (double)X[(long)d * sz + (c0 + c) * nr + r0 + r]; a += v * v; }
            costs[i - 1] = SUF(coefcost_bb)(X + (long)d * sz, r0, c0, rr, cc, nr, kind, (T)sqrt(a));
        }
    }
}

/* a22 bestbasis_treeselection 1-D BestBasis.jl:59-83 (+ delete_subtree! :128-140).
 * costs (ncost) is modified in place like the reference; tree gets n-1 entries. minmax: 0 = :min, 1 = :max */
WXO_API void SUF(tree_select1)(unsigned char *tree, T *costs, long ncost, long n, int minmax)
{
    long ntree = n - 1;
    int L = wx_ilog2(ncost);
    wx_maketree1(tree, n, L, 0);
    for (long i = ntree; i >= 1; --i) {
        if (!tree[i - 1]) continue;
        T pc = costs[i - 1];
        T cc = (T)(costs[2 * i - 1] + costs[2 * i]);
        if (minmax == 0 && cc < pc) costs[i - 1] = cc;
        else if (minmax == 1 && cc > pc) costs[i - 1] = cc;
        else wx_delete_subtree(tree, ntree, i, 2);
    }
}

/* a22 bestbasis_treeselection 2-D BestBasis.jl:85-110 */
WXO_API void SUF(tree_select2)(unsigned char *tree, T *costs, long ncost, long nr, long nc, int minmax)
{
    long ntree = wx_treelength2(nr, nc);
    int L = wx_quaddepth(ncost);
    wx_maketree2(tree, nr, nc, L, 0);
    for (long i = ntree; i >= 1; --i) {
        if (!tree[i - 1]) continue;
        T pc = costs[i - 1];
        T cc = (T)((T)((T)(costs[4 * i - 3] + costs[4 * i - 2]) + costs[4 * i - 1]) + costs[4 * i]);
        if (minmax == 0 && cc < pc) costs[i - 1] = cc;
        else if (minmax == 1 && cc > pc) costs[i - 1] = cc;
        else wx_delete_subtree(tree, ntree, i, 4);
    }
}

// wx_iwpt2d.cu -- batched 2-D inverse wavelet packet transform by quad tree with the separable step fused in shared memory.
//
// Reference: iwptall on images dwt/dwt_all.jl:210-225 -> iwpt! 2-D DWT.jl:662-710 (and iwpd! 2-D DWT.jl:354-401 after the
// getbasiscoef gather) -> idwt_step! 2-D dwt/dwt_one_level.jl:401-436: rows first (left | right quadrants -> temp), then
// columns (top | bottom halves of temp -> v).  1-D synthesis (dwt/dwt_one_level.jl:207-221), R = F/2:
//     v[2t]   = sum_r g[F-1-2r] w1[t-r] + h[2r+1] w2[t+r],    v[2t+1] = sum_r g[F-2-2r] w1[t-r] + h[2r] w2[t+r]   (mod half)
//
// Mirror of wx_wpd2d.cu:
//  * iwpt2d_block_k (nodes that fit shared memory): a CTA keeps a block of the coefficient image in shared memory and runs
//    every level from the deepest one up to the block's own depth in place (row pass A -> T, column pass T -> A); nodes the
//    tree does not split are simply left alone.  One read and one write of the image for all those levels.
//  * iwpt2d_tile_k (larger nodes, one level per launch): a CTA rebuilds a 2tr x 2tc tile of a node from four
//    (tr+hh) x (tc+hh) patches of its children (periodic halo hh >= F/2-1); both passes run in shared memory.
#include "wx_steps.cuh"
#include "wx_2d.cuh"
#include "wx_tma.cuh"
#include <cstdlib>

namespace {

// a[0..R-1] = w1[t-(R-1) .. t],  b[0..R-1] = w2[t .. t+R-1]  ->  v[2t], v[2t+1]
template <typename T, int F>
__device__ __forceinline__ void idwt_pair(const T *a, const T *b, const Taps<T> &tp, T &ev, T &od)
{
    constexpr int R = F / 2;
    T e = tp.g[F - 1] * a[R - 1];
    T o = tp.g[F - 2] * a[R - 1];
    e = fma(tp.h[1], b[0], e);
    o = fma(tp.h[0], b[0], o);
#pragma unroll
    for (int r = 1; r < R; ++r) {
        e = fma(tp.g[F - 1 - 2 * r], a[R - 1 - r], e);
        o = fma(tp.g[F - 2 - 2 * r], a[R - 1 - r], o);
        e = fma(tp.h[2 * r + 1], b[r], e);
        o = fma(tp.h[2 * r], b[r], o);
    }
    ev = e; od = o;
}

// ---------------------------------------------------------------------------------------------------------
// one level, tiles with halo: node (jr, jc) of depth d in `dst` from its four quadrants in `src`
// ---------------------------------------------------------------------------------------------------------
template <typename T, int F>
__global__ void __launch_bounds__(kT2) iwpt2d_tile_k(T *__restrict__ dst, const T *__restrict__ src, int m, int n, int d, int tr, int tc,
                                                    const unsigned char *__restrict__ tree, long ntree, Taps<T> tp)
{
    using P2 = typename Pair<T>::type;
    constexpr int R = F / 2, H = R - 1, HH = (H + 1) & ~1;          // halo, rounded up to even so that element pairs stay aligned
    extern __shared__ __align__(16) unsigned char wx_2d_smem[];
    const int PRh = tr + HH, PCh = tc + HH, LDP = 2 * PRh;
    T *P = reinterpret_cast<T *>(wx_2d_smem);                        // (2 PRh, 2 PCh): [lo rows | hi rows] x [left cols | right cols]
    T *Tm = P + LDP * 2 * PCh;                                        // (2 PRh, 2 tc) after the row pass
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int mp = m >> d, np = n >> d, hr = mp / 2, hc = np / 2;
    const int tiles_r = hr / tr, tiles_c = hc / tc, nodes = 1 << d;
    const unsigned rowtiles = (unsigned)(tiles_r * nodes);
    const unsigned k = blockIdx.x / rowtiles, rt = blockIdx.x - k * rowtiles;
    const int jr = (int)(rt / (unsigned)tiles_r), ti = (int)(rt - (unsigned)jr * tiles_r);
    const int jc = (int)(blockIdx.y / (unsigned)tiles_c), tk = (int)(blockIdx.y - (unsigned)jc * tiles_c);
    const long img = (long)m * n;
    const int nr0 = jr * mp, nc0 = jc * np, i0 = ti * tr, k0 = tk * tc;
    const T *sn = src + (long)k * img + (long)nc0 * m + nr0;          // node origin
    T *dn = dst + (long)k * img + (long)nc0 * m + nr0;

    if (!split2(tree, ntree, d, jr, jc)) {                            // leaf of the tree: the region passes through
        if (dst != src && lane < tr) {
            for (int b = warp; b < 2 * tc; b += kT2 / 32) {
                const long off = (long)(2 * k0 + b) * m + 2 * i0 + 2 * lane;
                *reinterpret_cast<P2 *>(dn + off) = *reinterpret_cast<const P2 *>(sn + off);
            }
        }
        return;
    }
    // ---- four child patches, two rows per asynchronous copy; a warp per patch column ----
    {
        const int PR2 = PRh;                                          // row pairs per patch column (2 PRh rows)
        for (int b = warp; b < 2 * PCh; b += kT2 / 32) {
            int cc;
            if (b < PCh) { cc = k0 - HH + b; while (cc < 0) cc += hc; while (cc >= hc) cc -= hc; }
            else { cc = k0 + (b - PCh); while (cc >= hc) cc -= hc; cc += hc; }
            for (int a2 = lane; a2 < PR2; a2 += 32) {
                const int a = 2 * a2;
                int rr;
                if (a < PRh) { rr = i0 - HH + a; while (rr < 0) rr += hr; while (rr >= hr) rr -= hr; }
                else { rr = i0 + (a - PRh); while (rr >= hr) rr -= hr; rr += hr; }
                cp_async_pair<T>(P + b * LDP + a, sn + cc * m + rr);
            }
        }
        cp_async_wait_all();
    }
    __syncthreads();
    // ---- row pass (along the columns): every patch row, tc output pairs; a thread slides over KR consecutive pairs ----
    {
        constexpr int KR = 8;
        const int ngrp = (tc + KR - 1) / KR;
        for (Walk2 w(tid, LDP); w.hi < ngrp; w.next()) {
            const int a = w.lo, kl0 = KR * w.hi;
            T wl[KR + H], wrt[KR + H];
            const T *pl = P + (kl0 + HH - H) * LDP + a;               // w1[k - r]: left cols  kl + HH - r
            const T *pr = P + (PCh + kl0) * LDP + a;                  // w2[k + r]: right cols kl + r
#pragma unroll
            for (int j = 0; j < KR + H; ++j) {
                const bool ok = kl0 + j < tc + H;                     // stay inside the patch for a partial last group
                wl[j] = ok ? pl[j * LDP] : (T)0;
                wrt[j] = ok ? pr[j * LDP] : (T)0;
            }
            T *o = Tm + (2 * kl0) * LDP + a;
#pragma unroll
            for (int p = 0; p < KR; ++p) {
                if (kl0 + p < tc) {
                    T ev, od;
                    idwt_pair<T, F>(&wl[p], &wrt[p], tp, ev, od);
                    o[(2 * p) * LDP] = ev;
                    o[(2 * p + 1) * LDP] = od;
                }
            }
        }
    }
    __syncthreads();
    // ---- column pass + store: lanes walk the output pairs of a column (consecutive 16-byte stores), two columns in flight ----
    if (tr <= 32 && (32 % tr) == 0) {
        const int cpw = 32 / tr, tl = lane % tr, bstep = (kT2 / 32) * cpw;
        T *o = dn + (long)(2 * k0) * m + 2 * i0 + 2 * tl;
        for (int b = warp * cpw + lane / tr; b < 2 * tc; b += 2 * bstep) {
            const bool two = b + bstep < 2 * tc;
            const T *c0 = Tm + b * LDP + tl;
            const T *c1 = two ? c0 + bstep * LDP : c0;
            T a0[R], b0[R], a1[R], b1[R];
#pragma unroll
            for (int r = 0; r < R; ++r) {
                a0[r] = c0[HH - H + r]; b0[r] = c0[PRh + r];
                a1[r] = c1[HH - H + r]; b1[r] = c1[PRh + r];
            }
            T e0, o0, e1, o1;
            idwt_pair<T, F>(a0, b0, tp, e0, o0);
            idwt_pair<T, F>(a1, b1, tp, e1, o1);
            P2 v; v.x = e0; v.y = o0;
            *reinterpret_cast<P2 *>(o + (long)b * m) = v;
            if (two) { v.x = e1; v.y = o1; *reinterpret_cast<P2 *>(o + (long)(b + bstep) * m) = v; }
        }
    } else {
        T *o = dn + (long)(2 * k0) * m + 2 * i0;
        for (Walk2 w(tid, tr); w.hi < 2 * tc; w.next()) {
            const int tl = w.lo, b = w.hi;
            const T *c0 = Tm + b * LDP + tl;
            T a0[R], b0[R];
#pragma unroll
            for (int r = 0; r < R; ++r) { a0[r] = c0[HH - H + r]; b0[r] = c0[PRh + r]; }
            T e0, o0;
            idwt_pair<T, F>(a0, b0, tp, e0, o0);
            P2 v; v.x = e0; v.y = o0;
            *reinterpret_cast<P2 *>(o + (long)b * m + 2 * tl) = v;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// one level, 64 x 64 output tiles (32 x 32 pairs) with every extent known at compile time.  One shared array P of
// ROWS = 2 (32 + HH) rows [lo part | hi part] x COLS = 2 (32 + H) columns [left | right]; the row pass runs IN PLACE
// (every thread first loads its window of 2 (8 + H) values into registers, barrier, then overwrites columns 0..63 of its
// row), so the tile needs 41 KB instead of 78 KB and three to four CTAs stay resident.  The column pass does the same down
// KSEG output pairs of one column with lanes across columns (pair loads / stores, conflict free because LDP/2 is odd); the
// 64 finished columns of the tile leave shared memory as 64 bulk copies (UBLKCP) of one contiguous run each.
// ---------------------------------------------------------------------------------------------------------
__host__ __device__ constexpr int wx_ld_pairs_i(int rows) { return ((rows / 2) % 2 == 0) ? rows + 2 : rows; }

template <typename T, int F>
struct ITile {
    static constexpr int R = F / 2, H = R - 1, HH = (H + 1) & ~1, TR = 32, PRh = TR + HH, ROWS = 2 * PRh, LDP = wx_ld_pairs_i(ROWS);
    static constexpr int PCh = TR + H, COLS = 2 * PCh, KR = 8, NG = TR / KR, NT = ROWS * NG, KSEG = 8;
    static constexpr int MINB = NT <= 288 ? 3 : 2;
    static constexpr size_t SMEM = (size_t)LDP * COLS * sizeof(T);
};

template <typename T, int F>
__global__ void __launch_bounds__(ITile<T, F>::NT, ITile<T, F>::MINB)
iwpt2d_tile32_k(T *__restrict__ dst, const T *__restrict__ src, int m, int n, int d, const unsigned char *__restrict__ tree, long ntree, Taps<T> tp)
{
    using P2 = typename Pair<T>::type;
    using C = ITile<T, F>;
    constexpr int H = C::H, HH = C::HH, TR = C::TR, PRh = C::PRh, LDP = C::LDP, PCh = C::PCh, COLS = C::COLS, KR = C::KR, NT = C::NT, KSEG = C::KSEG;
    extern __shared__ __align__(16) unsigned char wx_2d_smem[];
    T *P = reinterpret_cast<T *>(wx_2d_smem);
    const int tid = threadIdx.x;
    const int mp = m >> d, np = n >> d, hr = mp / 2, hc = np / 2;
    const int ltr = 31 - __clz(hr / TR), ltc = 31 - __clz(hc / TR);      // tiles per node edge are powers of two here (host checks)
    // grid.x = image * node rows * row tiles (fastest), grid.y = node cols * col tiles
    const unsigned rt = blockIdx.x & ((1u << (d + ltr)) - 1), k = blockIdx.x >> (d + ltr);
    const int jr = (int)(rt >> ltr), ti = (int)(rt & ((1u << ltr) - 1));
    const int jc = (int)(blockIdx.y >> ltc), tk = (int)(blockIdx.y & ((1u << ltc) - 1));
    const long img = (long)m * n;
    const int nr0 = jr * mp, nc0 = jc * np, i0 = ti * TR, k0 = tk * TR;
    const T *sn = src + (long)k * img + (long)nc0 * m + nr0;          // node origin
    T *dn = dst + (long)k * img + (long)nc0 * m + nr0;

    if (!split2(tree, ntree, d, jr, jc)) {                            // leaf of the tree: the region passes through
        if (dst != src) {
            for (int idx = tid; idx < 2 * TR * TR; idx += NT) {
                const int b = idx / TR, a2 = idx % TR;
                const long off = (long)(2 * k0 + b) * m + 2 * i0 + 2 * a2;
                *reinterpret_cast<P2 *>(dn + off) = *reinterpret_cast<const P2 *>(sn + off);
            }
        }
        return;
    }
    // ---- four child patches (periodic halo), one element pair per asynchronous copy ----
    for (int idx = tid; idx < COLS * PRh; idx += NT) {
        const int b = idx / PRh, a = 2 * (idx - b * PRh);
        int cc, rr;
        if (b < PCh) { cc = k0 - H + b; if (cc < 0) cc += hc; }
        else { cc = k0 + (b - PCh); if (cc >= hc) cc -= hc; cc += hc; }
        if (a < PRh) { rr = i0 - HH + a; if (rr < 0) rr += hr; }
        else { rr = i0 + (a - PRh); if (rr >= hr) rr -= hr; rr += hr; }
        cp_async_pair<T>(P + b * LDP + a, sn + (long)cc * m + rr);
    }
    cp_async_wait_all();
    __syncthreads();
    // ---- row pass (along the columns), in place: thread = (patch row a, group of KR output pairs) ----
    {
        const int a = tid % C::ROWS, kl0 = KR * (tid / C::ROWS);
        T wa[KR + H], wb[KR + H];
        const T *pl = P + kl0 * LDP + a;                              // w1[k - H + j]: left column kl0 + j
        const T *pr = P + (PCh + kl0) * LDP + a;                      // w2[k + j]:     right column kl0 + j
#pragma unroll
        for (int j = 0; j < KR + H; ++j) { wa[j] = pl[j * LDP]; wb[j] = pr[j * LDP]; }
        __syncthreads();                                              // every window is in registers: the row may be overwritten
        T *o = P + (2 * kl0) * LDP + a;
#pragma unroll
        for (int p = 0; p < KR; ++p) {
            T ev, od;
            idwt_pair<T, F>(&wa[p], &wb[p], tp, ev, od);
            o[(2 * p) * LDP] = ev;
            o[(2 * p + 1) * LDP] = od;
        }
    }
    __syncthreads();
    // ---- column pass, in place like the row pass: thread = (column b, KSEG consecutive output pairs), lanes across the columns ----
    {
        constexpr int NW = KSEG + HH;
        const int b = tid % (2 * TR), tl0 = KSEG * (tid / (2 * TR));
        const bool act = tid < 2 * TR * (TR / KSEG);
        T wa[NW], wb[NW];
        if (act) {
            const T *q = P + b * LDP + tl0;
#pragma unroll
            for (int j = 0; j < NW / 2; ++j) {
                const P2 v = *reinterpret_cast<const P2 *>(q + 2 * j);
                const P2 u = *reinterpret_cast<const P2 *>(q + PRh + 2 * j);
                wa[2 * j] = v.x; wa[2 * j + 1] = v.y;
                wb[2 * j] = u.x; wb[2 * j + 1] = u.y;
            }
        }
        __syncthreads();
        if (act) {
            T *o = P + b * LDP + 2 * tl0;
#pragma unroll
            for (int p = 0; p < KSEG; ++p) {
                P2 v;
                idwt_pair<T, F>(&wa[p + HH - H], &wb[p], tp, v.x, v.y);
                *reinterpret_cast<P2 *>(o + 2 * p) = v;
            }
        }
    }
    // ---- the 64 finished columns leave as bulk copies (one contiguous run of 64 elements each) ----
    if (sizeof(T) == 8) {
        wx_fence_proxy_async();
        __syncthreads();
        if (tid < 2 * TR) {
            wx_bulk_store_1d(dn + (long)(2 * k0 + tid) * m + 2 * i0, P + tid * LDP, (unsigned)(2 * TR * sizeof(T)));
            wx_bulk_commit();
            wx_bulk_wait_read0();
        }
    } else {                                                          // Float32 columns start 8-byte aligned only (LDP = 2 mod 4): pair copies
        __syncthreads();
        T *o0 = dn + (long)(2 * k0) * m + 2 * i0;
        for (int idx = tid; idx < 2 * TR * TR; idx += NT) {
            const int b = idx / TR, a2 = idx % TR;
            *reinterpret_cast<P2 *>(o0 + (long)b * m + 2 * a2) = *reinterpret_cast<const P2 *>(P + b * LDP + 2 * a2);
        }
    }
}

// One level of the whole-block kernel with everything known at compile time: block edge BE, node edge MPL (powers of two,
// MPL/2 >= 8).  Node / offset splits are shifts, periodic wraps are masks; both passes slide a register window
// (row pass: KROW output pairs along a row, lanes down the rows; column pass: KSEG output pairs down a column, lanes across
// the columns with pair loads / stores, conflict free because LD/2 is odd).  jrb, jcb: depth-l node index of the block's first node.
constexpr int KSEGI = 4;
template <typename T, int F, int BE, int MPL>
__device__ __forceinline__ void iwpt2d_block_level_ct(T *__restrict__ A, T *__restrict__ Tm, const Taps<T> &tp, int tid,
                                                      const unsigned char *__restrict__ tree, long ntree, int l, int jrb, int jcb)
{
    using P2 = typename Pair<T>::type;
    constexpr int R = F / 2, H = R - 1, HH = (H + 1) & ~1, LD = wx_ld_pairs_i(BE), HR = MPL / 2, KROW = (F >= 12 ? 4 : 8);
    // ---- row pass A -> Tm ----
    {
        const int r = tid % BE, ir = r / MPL;
        for (int g = tid / BE; g < BE / (2 * KROW); g += kT2 / BE) {
            const int kg0 = KROW * g, jn = kg0 / HR, kl0 = kg0 % HR, c0 = jn * MPL;
            if (tree != nullptr && !split2(tree, ntree, l, jrb + ir, jcb + jn)) continue;
            const T *q = A + c0 * LD + r;
            T wa[KROW + H], wb[KROW + H];
#pragma unroll
            for (int j = 0; j < KROW + H; ++j) {
                wa[j] = q[((kl0 - H + j) & (HR - 1)) * LD];
                wb[j] = q[(HR + ((kl0 + j) & (HR - 1))) * LD];
            }
            T *o = Tm + (c0 + 2 * kl0) * LD + r;
#pragma unroll
            for (int p = 0; p < KROW; ++p) {
                T ev, od;
                idwt_pair<T, F>(&wa[p], &wb[p], tp, ev, od);
                o[(2 * p) * LD] = ev;
                o[(2 * p + 1) * LD] = od;
            }
        }
    }
    __syncthreads();
    // ---- column pass Tm -> A ----
    for (int t = tid; t < (BE / 2 / KSEGI) * BE; t += kT2) {
        const int sg = t / BE, c = t % BE, ig0 = sg * KSEGI;
        const int jn = ig0 / HR, tl0 = ig0 % HR, r0 = jn * MPL;
        if (tree != nullptr && !split2(tree, ntree, l, jrb + jn, jcb + c / MPL)) continue;
        const T *q = Tm + c * LD + r0;
        T wa[KSEGI + HH], wb[KSEGI + HH];
#pragma unroll
        for (int j = 0; j < (KSEGI + HH) / 2; ++j) {
            const P2 v = *reinterpret_cast<const P2 *>(q + ((tl0 - HH + 2 * j) & (HR - 1)));
            const P2 u = *reinterpret_cast<const P2 *>(q + HR + ((tl0 + 2 * j) & (HR - 1)));
            wa[2 * j] = v.x; wa[2 * j + 1] = v.y;
            wb[2 * j] = u.x; wb[2 * j + 1] = u.y;
        }
        T *o = A + c * LD + r0 + 2 * tl0;
#pragma unroll
        for (int p = 0; p < KSEGI; ++p) {
            P2 v;
            idwt_pair<T, F>(&wa[p + HH - H], &wb[p], tp, v.x, v.y);
            *reinterpret_cast<P2 *>(o + 2 * p) = v;
        }
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------------------
// levels dend-1 .. db of every node inside a block (= node of depth db) that fits shared memory, in place
// ---------------------------------------------------------------------------------------------------------
// BE > 0: square block of edge BE known at compile time (padded leading dimension, compile-time-shaped levels)
template <typename T, int F, int BE>
__global__ void __launch_bounds__(kT2, (BE > 0 ? 3 : 1)) iwpt2d_block_k(T *__restrict__ dst, const T *__restrict__ src, int m, int n, int db, int dend,
                                                     const unsigned char *__restrict__ tree, long ntree, Taps<T> tp)
{
    using P2 = typename Pair<T>::type;
    constexpr int R = F / 2;
    extern __shared__ __align__(16) unsigned char wx_2d_smem[];
    const int BR = BE > 0 ? BE : (m >> db), BC = BE > 0 ? BE : (n >> db);
    const int LD = BE > 0 ? wx_ld_pairs_i(BE) : BR;
    T *A = reinterpret_cast<T *>(wx_2d_smem);
    T *Tm = A + LD * BC;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const unsigned k = blockIdx.x >> db;
    const int jr = (int)(blockIdx.x & ((1u << db) - 1)), jc = (int)blockIdx.y;
    const long img = (long)m * n;
    const long org = (long)(jc * BC) * m + jr * BR;
    const T *sp = src + (long)k * img + org;
    T *dp = dst + (long)k * img + org;
    const int BR2 = BR / 2;
    const bool colwarp = BR2 <= 32 && (32 % BR2) == 0;
    const int cpw = colwarp ? 32 / BR2 : 1, ca = 2 * (lane % BR2), cb0 = warp * cpw + lane / BR2, cbs = (kT2 / 32) * cpw;

    if (colwarp) { for (int b = cb0; b < BC; b += cbs) cp_async_pair<T>(A + b * LD + ca, sp + b * m + ca); }
    else { for (Walk2 w(tid, BR2); w.hi < BC; w.next()) cp_async_pair<T>(A + w.hi * LD + 2 * w.lo, sp + w.hi * m + 2 * w.lo); }
    cp_async_wait_all();
    __syncthreads();
    for (int l = dend - 1; l >= db; --l) {
        const int mpl = m >> l, npl = n >> l, hr = mpl / 2, hc = npl / 2;
        const int sh = l - db;                                        // nodes of depth l per block edge = 1 << sh
        if (BE == 64 && mpl == npl && mpl >= 16) {
            constexpr int BEc = BE > 0 ? BE : 64;
            if (mpl == 64) iwpt2d_block_level_ct<T, F, BEc, 64>(A, Tm, tp, tid, tree, ntree, l, jr << sh, jc << sh);
            else if (mpl == 32) iwpt2d_block_level_ct<T, F, BEc, 32>(A, Tm, tp, tid, tree, ntree, l, jr << sh, jc << sh);
            else iwpt2d_block_level_ct<T, F, BEc, 16>(A, Tm, tp, tid, tree, ntree, l, jr << sh, jc << sh);
            continue;
        }
        const FastDiv dhr(hr), dhc(hc);
        // ---- row pass A -> Tm: every row of a split node, hc output pairs (left | right quadrant columns) ----
        for (Walk2 w(tid, BR); w.hi < BC / 2; w.next()) {
            const int r = w.lo, jn = dhc.div(w.hi), kl = w.hi - jn * hc, c0 = jn * npl;
            const int ir = r / mpl;                                   // node row inside the block
            if (!split2(tree, ntree, l, (jr << sh) + ir, (jc << sh) + jn)) continue;
            const T *q = A + c0 * LD + r;
            T a[R], b[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                int cl = kl - (R - 1) + rr; if (cl < 0) cl = (hc >= R) ? cl + hc : ((cl % hc) + hc) % hc;
                int cr = kl + rr; if (cr >= hc) cr = (hc >= R) ? cr - hc : cr % hc;
                a[rr] = q[cl * LD];
                b[rr] = q[(hc + cr) * LD];
            }
            T ev, od;
            idwt_pair<T, F>(a, b, tp, ev, od);
            T *t = Tm + c0 * LD + r;
            t[(2 * kl) * LD] = ev;
            t[(2 * kl + 1) * LD] = od;
        }
        __syncthreads();
        // ---- column pass Tm -> A: every column of a split node, hr output pairs (top | bottom halves) ----
        for (Walk2 w(tid, BR2); w.hi < BC; w.next()) {
            const int c = w.hi, jn = dhr.div(w.lo), tl = w.lo - jn * hr, r0 = jn * mpl;
            const int ic = c / npl;
            if (!split2(tree, ntree, l, (jr << sh) + jn, (jc << sh) + ic)) continue;
            const T *q = Tm + c * LD + r0;
            T a[R], b[R];
#pragma unroll
            for (int rr = 0; rr < R; ++rr) {
                int rl = tl - (R - 1) + rr; if (rl < 0) rl = (hr >= R) ? rl + hr : ((rl % hr) + hr) % hr;
                int rh = tl + rr; if (rh >= hr) rh = (hr >= R) ? rh - hr : rh % hr;
                a[rr] = q[rl];
                b[rr] = q[hr + rh];
            }
            T ev, od;
            idwt_pair<T, F>(a, b, tp, ev, od);
            P2 v; v.x = ev; v.y = od;
            *reinterpret_cast<P2 *>(A + c * LD + r0 + 2 * tl) = v;
        }
        __syncthreads();
    }
    if (colwarp) { for (int b = cb0; b < BC; b += cbs) *reinterpret_cast<P2 *>(dp + b * m + ca) = *reinterpret_cast<const P2 *>(A + b * LD + ca); }
    else { for (Walk2 w(tid, BR2); w.hi < BC; w.next()) *reinterpret_cast<P2 *>(dp + w.hi * m + 2 * w.lo) = *reinterpret_cast<const P2 *>(A + w.hi * LD + 2 * w.lo); }
}

static int largest_even_divisor_le(long v, int cap)
{
    int best = 0;
    for (int t = 2; t <= cap && t <= v; t += 2) if (v % t == 0) best = t;
    return best;
}

template <typename T, int F>
int iwpt2d_run(T *y, const T *xw, T *scratch, long m, long n, int nlev, long N, const unsigned char *dtree, long ntree, const Taps<T> &t,
               cudaStream_t s)
{
    constexpr int H = F / 2 - 1, HH = (H + 1) & ~1;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const size_t budget = 65536;
    int db = 0;
    while (db < nlev && (size_t)2 * (m >> db) * (n >> db) * sizeof(T) > budget) ++db;
    // plan the tile levels first: every one must be coverable, else the caller's per-level path runs everything
    int trs[64], tcs[64];
    for (int d = 0; d < db; ++d) {
        const long hr = (m >> d) / 2, hc = (n >> d) / 2;
        if (hr % 2 != 0 || hc % 2 != 0) return wx_fail(WX_EUNSUPPORTED, "odd half node");
        trs[d] = largest_even_divisor_le(hr, 32); tcs[d] = largest_even_divisor_le(hc, 32);
        if (trs[d] < 2 || tcs[d] < 2) return wx_fail(WX_EUNSUPPORTED, "no tile");
        const size_t smem = ((size_t)2 * (trs[d] + HH) * 2 * (tcs[d] + HH) + (size_t)2 * (trs[d] + HH) * 2 * tcs[d]) * sizeof(T);
        if (smem > dv.smem_optin) return wx_fail(WX_EUNSUPPORTED, "tile does not fit");
        if ((hr / trs[d]) * (1L << d) * N >= (1L << 31) || (hc / tcs[d]) * (1L << d) > 65535) return wx_fail(WX_EUNSUPPORTED, "grid");
    }
    if ((1L << db) * N >= (1L << 31) || (1L << db) > 65535) return wx_fail(WX_EUNSUPPORTED, "grid");
    // ping-pong between y and scratch so that the last launch writes y
    const int nlaunch = (db < nlev ? 1 : 0) + (db < nlev ? db : nlev);
    const T *cur = xw;
    int left = nlaunch;
    auto next_dst = [&]() { T *d2 = (left % 2 == 1) ? y : scratch; --left; return d2; };
    static const bool no_ct = getenv("WX_B200_IWPT2D_GENERIC") != nullptr;      // A-B knob: the run-time-shaped kernels only
    if (db < nlev) {
        T *dst = next_dst();
        const dim3 grid((unsigned)((1L << db) * N), (unsigned)(1L << db));
        if (!no_ct && (m >> db) == 64 && (n >> db) == 64) {
            const size_t smem = (size_t)2 * wx_ld_pairs_i(64) * 64 * sizeof(T);
            auto kern = iwpt2d_block_k<T, F, 64>;
            WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, kT2, smem, s>>>(dst, cur, (int)m, (int)n, db, nlev, dtree, ntree, t);
        } else {
            const size_t smem = (size_t)2 * (m >> db) * (n >> db) * sizeof(T);
            auto kern = iwpt2d_block_k<T, F, 0>;
            WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<grid, kT2, smem, s>>>(dst, cur, (int)m, (int)n, db, nlev, dtree, ntree, t);
        }
        WX_LAUNCHED();
        cur = dst;
    }
    auto pow2 = [](long v) { return v > 0 && (v & (v - 1)) == 0; };
    for (int d = (db < nlev ? db : nlev) - 1; d >= 0; --d) {
        T *dst = next_dst();
        const long hr = (m >> d) / 2, hc = (n >> d) / 2;
        if (!no_ct && hr >= 32 && hc >= 32 && pow2(hr) && pow2(hc)) {               // 64 x 64 output tiles, compile-time shaped
            using C = ITile<T, F>;
            auto kern = iwpt2d_tile32_k<T, F>;
            WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)C::SMEM));
            kern<<<dim3((unsigned)((hr / 32) * (1L << d) * N), (unsigned)((hc / 32) * (1L << d))), C::NT, C::SMEM, s>>>(dst, cur, (int)m, (int)n, d,
                                                                                                                     dtree, ntree, t);
            WX_LAUNCHED();
            cur = dst;
            continue;
        }
        const int tr = trs[d], tc = tcs[d];
        const size_t smem = ((size_t)2 * (tr + HH) * 2 * (tc + HH) + (size_t)2 * (tr + HH) * 2 * tc) * sizeof(T);
        auto kern = iwpt2d_tile_k<T, F>;
        WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<dim3((unsigned)((hr / tr) * (1L << d) * N), (unsigned)((hc / tc) * (1L << d))), kT2, smem, s>>>(dst, cur, (int)m, (int)n, d, tr, tc,
                                                                                                     dtree, ntree, t);
        WX_LAUNCHED();
        cur = dst;
    }
    return WX_OK;
}

}  // namespace

// xw(m,n,N) -> y(m,n,N) along the quad tree (nlev = number of levels to undo); dtree = device copy of the tree or nullptr
// for a complete tree.  scratch: N images (needed when more than one launch runs; may alias nothing else).
// *handled = false: shape not covered, nothing was launched.
template <typename T>
int wx_iwpt2d_fused(T *y, const T *xw, T *scratch, long m, long n, int nlev, long N, const unsigned char *dtree, long ntree, const Taps<T> &t,
                    cudaStream_t s, bool *handled)
{
    *handled = false;
    static const bool off = getenv("WX_B200_NO_FUSED_WPD2D") != nullptr;
    if (off || nlev < 1 || nlev > 30 || N < 1 || m * n >= (1L << 31) || y == xw) return WX_OK;
    if (((((uintptr_t)y) | ((uintptr_t)xw) | ((uintptr_t)scratch)) & 15) != 0) return WX_OK;
    if ((m >> (nlev - 1)) % 2 != 0 || (n >> (nlev - 1)) % 2 != 0) return WX_OK;
    int rc;
#define WX_I2_CASE(FF) case FF: rc = iwpt2d_run<T, FF>(y, xw, scratch, m, n, nlev, N, dtree, ntree, t, s); break;
    switch (t.F) {
        WX_I2_CASE(2) WX_I2_CASE(4) WX_I2_CASE(6) WX_I2_CASE(8) WX_I2_CASE(10) WX_I2_CASE(12) WX_I2_CASE(16) WX_I2_CASE(20)
        default: return WX_OK;
    }
#undef WX_I2_CASE
    if (rc == WX_EUNSUPPORTED) return WX_OK;         // planned before any launch: nothing ran
    if (rc == WX_OK) *handled = true;
    return rc;
}
template int wx_iwpt2d_fused<double>(double *, const double *, double *, long, long, int, long, const unsigned char *, long, const Taps<double> &, cudaStream_t, bool *);
template int wx_iwpt2d_fused<float>(float *, const float *, float *, long, long, int, long, const unsigned char *, long, const Taps<float> &, cudaStream_t, bool *);

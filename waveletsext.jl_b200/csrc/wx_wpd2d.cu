// wx_wpd2d.cu -- batched 2-D wavelet packet decomposition with the separable step fused in shared memory.
//
// Reference: wpdall on images dwt/dwt_all.jl:260-282 -> wpd! 2-D DWT.jl:164-209 -> dwt_step! 2-D
// dwt/dwt_one_level.jl:319-354 (columns into temp, rows into the four quadrants w1 = LL top-left, w2 top-right,
// w3 bottom-left, w4 bottom-right).   x(m,n,N) -> y(m,n,L+1,N), level d+1 node (jr,jc) = quadrants of level d.
//
// Two kernels, both HBM-bound (2F FMAs per pixel per level):
//  * wpd2d_tile_k  (nodes larger than shared memory): a CTA produces a tr x tc tile of each of the four children of
//    one node from the (2tr+F-2) x (2tc+F-2) parent patch (periodic halo); column pass and row pass both run in shared
//    memory, so a level costs one read of the parent (+halo) and one write of the children instead of two round trips
//    through a temp image.  At depth 0 the patch core is also written to level 0 (y[:,:,1] = x, DWT.jl:176).
//  * wpd2d_block_k (nodes that fit): a CTA keeps a whole node in shared memory and runs every remaining level there
//    (node -> 4 children -> 16 grandchildren ...), writing each level slice once; the parent is never re-read.
// The detail outputs are taken S = (F-2)/2 positions ahead of the scaling outputs so that both filters read the same
// window v[2i .. 2i+F-1]; accumulation order per output is the reference's tap order.
#include "wx_steps.cuh"
#include "wx_2d.cuh"
#include <cstdlib>

namespace {

// K consecutive output pairs from one window: win[0 .. 2K+F-3] = v[2i .. 2i+2K+F-3]; pair p uses win[2p .. 2p+F-1]
// row pass: output columns per thread (a window of 2*KROW+F-2 values slides along the row); halved for long filters so that the
// window stays in registers (F = 16 / 20 spilled 1.7 - 2.3 KB per thread with 8)
__host__ __device__ constexpr int wx_krow(int F) { return F >= 12 ? 4 : 8; }
constexpr int KSEG = 4;       // column pass of the compile-time-shaped kernels: output pairs per thread (window slides down the column)
// Whole-node kernel, WIDE variant (long filters in Float64): windows of 8 output pairs in both passes -- (16 + F - 2) / 16 loads per input
// instead of (8 + F - 2) / 8 -- under a two-CTA launch bound (128 registers), because the 80 registers of three resident CTAs spill them
// (ptxas: 256 - 784 B for 10 - 20 taps).
template <int F, bool WIDE> struct BlkCfg {
    static constexpr int KS = WIDE ? 8 : KSEG, KR = WIDE ? 8 : wx_krow(F), MINB = WIDE ? 2 : 3;
};

// leading dimensions of the shared-memory arrays.  Column pass: lanes walk COLUMNS and read element pairs, conflict free when
// (ld / 2) is odd; its outputs are single elements written by lanes that walk columns, conflict free when ld is odd.
__host__ __device__ constexpr int wx_ld_pairs(int rows) { return ((rows / 2) % 2 == 0) ? rows + 2 : rows; }
__host__ __device__ constexpr int wx_ld_odd(int rows) { return rows | 1; }

// Where a launch reads its parents and writes its children.  Packet table (wpd): par = x or a level slice of y, out = the next level
// slice, copy = level 0 of y on the first launch.  By quad tree (wpt, TREE = true): par / out are whole images that ping-pong between the
// launches, nodes the tree does not split pass through unchanged, and the whole-node kernel stores only what it holds after its last level.
template <typename T>
struct Io2D {
    const T *par; long pstride;              // parents of image k: par + k * pstride
    T *out; long ostride;                    // children of image k: out + k * ostride (whole-node kernel, table mode: + (l + 1) * img per level)
    T *copy; long cstride;                   // y[:,:,1] = x target (DWT.jl:176) or nullptr
    const unsigned char *tree; long ntree;   // TREE: quad tree on the device, nullptr = every node of these levels splits
};

// ---------------------------------------------------------------------------------------------------------
// one level, tiles with halo.  Shared memory: P (parent patch, PR rows x PC cols, column-major) and Tm (column-pass output,
// 2tr rows x PC cols: rows [0,tr) scaling, [tr,2tr) detail).
// TRC > 0: tile edge known at compile time (tr = tc = TRC): index arithmetic folds to shifts / immediates and the column pass
// slides a register window down each column (KSEG pairs per thread, lanes across columns).
// ---------------------------------------------------------------------------------------------------------
template <typename T, int F, int TRC, int TCC = TRC, int MINB = 3, bool TREE = false, int KRW = 0>
__global__ void __launch_bounds__(kT2, MINB) wpd2d_tile_k(Io2D<T> io, int m, int n, int d, int tr_, int tc_,
                                                      Div32 drowtiles, Div32 dtiles_r, Div32 dtiles_c, Div32 dgx, long ntiles, Taps<T> tp)
{
    const int tr = TRC > 0 ? TRC : tr_, tc = TRC > 0 ? TCC : tc_;
    using P2 = typename Pair<T>::type;
    // narrow tiles: 4 column groups keep all 256 threads busy; KRW = 8: long filters with the 8-pair row window (two-CTA launch bound)
    constexpr int S = (F - 2) / 2, KROW = KRW > 0 ? KRW : ((TRC > 0 && TCC == 16) ? 4 : wx_krow(F));
    extern __shared__ __align__(16) unsigned char wx_2d_smem[];
    const int PR = 2 * tr + F - 2, PC = 2 * tc + F - 2, R2 = 2 * tr;
    const int LDP = TRC > 0 ? wx_ld_pairs(PR) : PR + 2 * (tr & 1);      // generic shape: odd tiles padded (see host)
    const int LDT = TRC > 0 ? wx_ld_odd(R2) : R2;
    T *P = reinterpret_cast<T *>(wx_2d_smem);
    T *Tm = P + LDP * PC;
    const int tid = threadIdx.x;
    const int mp = m >> d, np = n >> d, hr = mp / 2, hc = np / 2;
    const bool from_x = io.copy != nullptr;
    const int lane = tid & 31, warp = tid >> 5;

    // tile id -> (bx, by) in the order the hardware would schedule a (gx, gy) grid: bx = image * (row tiles of all nodes) fastest,
    // by = column tiles of all nodes; neighbouring CTAs walk down the rows (their halos meet in L2)
    struct Tile { long k; int nr0, nc0, i0, k0; bool split; };
    auto decode = [&](long id) {
        const unsigned by = div32((unsigned)id, dgx), bx = (unsigned)id - by * dgx.d;
        const unsigned k = div32(bx, drowtiles), rt = bx - k * drowtiles.d;
        const int jr = (int)div32(rt, dtiles_r), ti = (int)(rt - (unsigned)jr * dtiles_r.d);
        const int jc = (int)div32(by, dtiles_c), tk = (int)(by - (unsigned)jc * dtiles_c.d);
        Tile t; t.k = k; t.nr0 = jr * mp; t.nc0 = jc * np; t.i0 = ti * tr; t.k0 = tk * tc;
        t.split = !TREE || split2(io.tree, io.ntree, d, jr, jc);
        return t;
    };
    // parent patch of a tile (periodic inside the node) into P, two rows per asynchronous copy; a warp per column
    auto prefetch = [&](const Tile &t) {
        if (TREE && !t.split) return;
        const T *par = io.par + t.k * io.pstride + (long)t.nc0 * m + t.nr0;
        const int PR2 = PR / 2;
        int rr = 2 * t.i0 + 2 * lane; while (rr >= mp) rr -= mp;
        for (int b = warp; b < PC; b += kT2 / 32) {
            int cc = 2 * t.k0 + b; while (cc >= np) cc -= np;
            if (lane < PR2) cp_async_pair<T>(P + b * LDP + 2 * lane, par + cc * m + rr);
        }
        const int rem = PR2 - 32;                         // pairs 32.. of every column (halo rows of a 32-row tile)
        if (rem > 0) {
            for (int idx = tid; idx < rem * PC; idx += kT2) {
                const int b = idx / rem, a = 2 * (32 + idx - b * rem);
                int r2 = 2 * t.i0 + a; while (r2 >= mp) r2 -= mp;
                int cc = 2 * t.k0 + b; while (cc >= np) cc -= np;
                cp_async_pair<T>(P + b * LDP + a, par + cc * m + r2);
            }
        }
    };

    long id = blockIdx.x;
    if (id >= ntiles) return;
    Tile cur = decode(id);
    prefetch(cur);
    for (; id < ntiles; id += gridDim.x) {
    const int nr0 = cur.nr0, nc0 = cur.nc0, i0 = cur.i0, k0 = cur.k0;
    const long cur_k = cur.k;
    cp_async_wait_all();
    __syncthreads();                                      // patch complete; every thread is past the previous tile's row pass (Tm free)
    if (TREE && !cur.split) {                             // leaf of the tree: this tile's share of the node passes through
        const long org = (long)(nc0 + 2 * k0) * m + nr0 + 2 * i0;
        const T *sp = io.par + cur.k * io.pstride + org;
        T *dp = io.out + cur.k * io.ostride + org;
        if (dp != sp)
            for (Walk2 w(tid, tr); w.hi < 2 * tc; w.next())
                *reinterpret_cast<P2 *>(dp + (long)w.hi * m + 2 * w.lo) = *reinterpret_cast<const P2 *>(sp + (long)w.hi * m + 2 * w.lo);
        if (id + gridDim.x < ntiles) { cur = decode(id + gridDim.x); prefetch(cur); }
        continue;
    }
    if (from_x && lane < tr) {                            // y[:,:,1] = x   DWT.jl:176 : the core of the patch (tr <= 32 pairs per column)
        T *y0 = io.copy + cur.k * io.cstride + (long)(nc0 + 2 * k0) * m + nr0 + 2 * i0 + 2 * lane;
        for (int b = warp; b < 2 * tc; b += kT2 / 32)
            *reinterpret_cast<P2 *>(y0 + b * m) = *reinterpret_cast<const P2 *>(P + b * LDP + 2 * lane);
    }
    // ---- column pass ----
    if (TRC > 0 && TRC % KSEG == 0) {
        // lanes walk the columns (pair loads, conflict free by the choice of LDP); a thread slides one window of
        // 2*KSEG+F-2 samples down KSEG output pairs of its column
        constexpr int PCc = 2 * TCC + F - 2, NSEG = (TRC > 0 ? TRC : KSEG) / KSEG, W = 2 * KSEG + F - 2;
        for (int t = tid; t < PCc * NSEG; t += kT2) {
            const int sg = t / PCc, b = t - sg * PCc, il0 = sg * KSEG;
            const T *src = P + b * LDP + 2 * il0;
            T win[W];
#pragma unroll
            for (int q = 0; q < W / 2; ++q) {
                const P2 v = *reinterpret_cast<const P2 *>(src + 2 * q);
                win[2 * q] = v.x; win[2 * q + 1] = v.y;
            }
            T *d0 = Tm + b * LDT + il0;
#pragma unroll
            for (int p = 0; p < KSEG; ++p) {
                T lo, hi;
                dwt_dots<T, F>(&win[2 * p], tp, lo, hi);
                d0[p] = lo; d0[tr + p] = hi;
            }
        }
    } else if (tr <= 32 && (32 % tr) == 0) {
        // lanes walk il (conflict-free 16-byte loads), 32/tr columns per warp, warps stride over the columns, two columns in flight
        const int cpw = 32 / tr, il = lane % tr, bstep = (kT2 / 32) * cpw;
        for (int b = warp * cpw + lane / tr; b < PC; b += 2 * bstep) {
            const bool two = b + bstep < PC;
            const T *s0 = P + b * LDP + 2 * il;
            const T *s1 = two ? s0 + bstep * LDP : s0;
            T w0[F], w1[F];
#pragma unroll
            for (int q = 0; q < F / 2; ++q) {
                const P2 v = *reinterpret_cast<const P2 *>(s0 + 2 * q);
                const P2 u = *reinterpret_cast<const P2 *>(s1 + 2 * q);
                w0[2 * q] = v.x; w0[2 * q + 1] = v.y;
                w1[2 * q] = u.x; w1[2 * q + 1] = u.y;
            }
            T lo0, hi0, lo1, hi1;
            dwt_dots<T, F>(w0, tp, lo0, hi0);
            dwt_dots<T, F>(w1, tp, lo1, hi1);
            T *d0 = Tm + b * LDT + il;
            d0[0] = lo0; d0[tr] = hi0;
            if (two) { d0[bstep * LDT] = lo1; d0[bstep * LDT + tr] = hi1; }
        }
    } else {
        const int trh = (tr + 1) / 2;
        for (Walk2 w(tid, trh); w.hi < PC; w.next()) {
            const int b = w.hi, il = w.lo;
            const bool two = il + trh < tr;
            T w0[F], w1[F];
            const T *src = P + b * LDP + 2 * il;
#pragma unroll
            for (int q = 0; q < F / 2; ++q) {
                const P2 v = *reinterpret_cast<const P2 *>(src + 2 * q);
                w0[2 * q] = v.x; w0[2 * q + 1] = v.y;
                const P2 u = *reinterpret_cast<const P2 *>(src + (two ? 2 * trh : 0) + 2 * q);
                w1[2 * q] = u.x; w1[2 * q + 1] = u.y;
            }
            T lo0, hi0, lo1, hi1;
            dwt_dots<T, F>(w0, tp, lo0, hi0);
            dwt_dots<T, F>(w1, tp, lo1, hi1);
            T *dst = Tm + b * LDT + il;
            dst[0] = lo0; dst[tr] = hi0;
            if (two) { dst[trh] = lo1; dst[tr + trh] = hi1; }
        }
    }
    __syncthreads();
    // P is dead from here on: the next tile's patch streams in while this tile's row pass runs
    if (id + gridDim.x < ntiles) { cur = decode(id + gridDim.x); prefetch(cur); }
    // ---- row pass + store: (2tr rows) x (tc output pairs) ----
    T *ynext = io.out + cur_k * io.ostride + (long)nc0 * m + nr0;
    if ((kT2 % R2) == 0 && (tc % KROW) == 0) {
        // a thread owns one row r and KROW consecutive output columns: one window of 2*KROW+F-2 samples slides along the row
        const int r = tid % R2;
        int ci, qr;
        if (r < tr) { ci = i0 + r; qr = 0; }
        else { ci = i0 + (r - tr) + S; while (ci >= hr) ci -= hr; qr = hr; }
        T *o = ynext + qr + ci;
        for (int g = tid / R2; g < tc / KROW; g += kT2 / R2) {
            const int kl0 = KROW * g;
            T win[2 * KROW + F - 2];
            const T *src = Tm + (2 * kl0) * LDT + r;
#pragma unroll
            for (int j = 0; j < 2 * KROW + F - 2; ++j) win[j] = src[j * LDT];
            int ch = k0 + kl0 + S; while (ch >= hc) ch -= hc;
            T *olo = o + (k0 + kl0) * m, *ohi = o + (hc + ch) * m;
            const long back = (long)hc * m;
#pragma unroll
            for (int p = 0; p < KROW; ++p) {
                T lo, hi;
                dwt_dots<T, F>(&win[2 * p], tp, lo, hi);
                *olo = lo;                                // w1 / w3
                *ohi = hi;                                // w2 / w4
                olo += m; ohi += m;
                if (++ch >= hc) { ch -= hc; ohi -= back; }
            }
        }
    } else {
        const int tcq = (tc + 1) / 2;
        for (Walk2 w(tid, R2); w.hi < tcq; w.next()) {
            const int kl = 2 * w.hi, r = w.lo;
            T win[F + 2];
            const T *src = Tm + (2 * kl) * LDT + r;
#pragma unroll
            for (int j = 0; j < F + 2; ++j) win[j] = src[j * LDT];
            T lo0, hi0, lo1, hi1;
            dwt_dots2<T, F>(win, tp, lo0, hi0, lo1, hi1);
            int ci, qr;
            if (r < tr) { ci = i0 + r; qr = 0; }
            else { ci = i0 + (r - tr) + S; while (ci >= hr) ci -= hr; qr = hr; }
            int ch0 = k0 + kl + S; while (ch0 >= hc) ch0 -= hc;
            T *o = ynext + qr + ci;
            o[(k0 + kl) * m] = lo0;                       // w1 / w3
            o[(hc + ch0) * m] = hi0;                      // w2 / w4
            if (kl + 1 < tc) {
                int ch1 = ch0 + 1; if (ch1 >= hc) ch1 -= hc;
                o[(k0 + kl + 1) * m] = lo1;
                o[(hc + ch1) * m] = hi1;
            }
        }
    }
    }   // tiles of this CTA
}

// ---------------------------------------------------------------------------------------------------------
// all remaining levels of a node that fits shared memory
// ---------------------------------------------------------------------------------------------------------
// one window of NW values starting at element `start` of a periodic vector of length len (stride st)
template <typename T, int NW>
__device__ __forceinline__ void load_window(T *win, const T *src, int start, int len, int st)
{
    if (len >= NW) {                                       // at most one wrap
#pragma unroll
        for (int j = 0; j < NW; ++j) { int e = start + j; if (e >= len) e -= len; win[j] = src[e * st]; }
    } else {
#pragma unroll
        for (int j = 0; j < NW; ++j) win[j] = src[((start + j) % len) * st];
    }
}

// same for a contiguous vector, even start and even length: 16-byte (8-byte for Float32) loads of element pairs
template <typename T, int NW>
__device__ __forceinline__ void load_window_pairs(T *win, const T *src, int start, int len)
{
    using P2 = typename Pair<T>::type;
    if (len >= NW) {
#pragma unroll
        for (int q = 0; q < NW / 2; ++q) {
            int e = start + 2 * q; if (e >= len) e -= len;
            const P2 v = *reinterpret_cast<const P2 *>(src + e);
            win[2 * q] = v.x; win[2 * q + 1] = v.y;
        }
    } else {
#pragma unroll
        for (int j = 0; j < NW; ++j) win[j] = src[(start + j) % len];
    }
}

// One level of the whole-node kernel with EVERYTHING known at compile time: block edge BE, current node edge MPL (both powers
// of two, MPL/2 a multiple of KSEG and KROW).  Node / offset splits are shifts, periodic wraps are masks (any number of wraps,
// so filters longer than the node need no special case), and the integer work per output drops to a few instructions.
// node (ir, ic) of depth l inside a block whose first node of that depth is (jrb, jcb): does the tree split it?
struct Split2 {
    const unsigned char *tree; long ntree; int l, jrb, jcb;
    __device__ __forceinline__ bool operator()(int ir, int ic) const { return tree == nullptr || split2(tree, ntree, l, jrb + ir, jcb + ic); }
};

template <typename T, int F, int BE, int MPL, bool TREE, bool WIDE>
__device__ __forceinline__ void wpd2d_block_level_ct(T *__restrict__ A, T *__restrict__ Tm, const Taps<T> &tp, int tid, const Split2 &sp)
{
    using P2 = typename Pair<T>::type;
    constexpr int S = (F - 2) / 2, LDA = wx_ld_pairs(BE), LDT = wx_ld_odd(BE), KROW = BlkCfg<F, WIDE>::KR, KSEG = BlkCfg<F, WIDE>::KS;
    constexpr int HR = MPL / 2, WS = 2 * KSEG + F - 2, WR = 2 * KROW + F - 2;
    // ---- column pass A -> Tm ----
    for (int t = tid; t < (BE / 2 / KSEG) * BE; t += kT2) {
        const int sg = t / BE, c = t % BE, ig0 = sg * KSEG;
        const int r0 = (ig0 / HR) * MPL, il0 = ig0 % HR;
        if (TREE && !sp(ig0 / HR, c / MPL)) continue;
        const T *src = A + c * LDA + r0;
        T win[WS];
#pragma unroll
        for (int q = 0; q < WS / 2; ++q) {
            const P2 v = *reinterpret_cast<const P2 *>(src + ((2 * il0 + 2 * q) & (MPL - 1)));
            win[2 * q] = v.x; win[2 * q + 1] = v.y;
        }
        T *dst = Tm + c * LDT + r0;
#pragma unroll
        for (int p = 0; p < KSEG; ++p) {
            T lo, hi;
            dwt_dots<T, F>(&win[2 * p], tp, lo, hi);
            dst[il0 + p] = lo;
            dst[HR + ((il0 + S + p) & (HR - 1))] = hi;
        }
    }
    __syncthreads();
    // ---- row pass Tm -> A ----
    {
        const int r = tid % BE;
        for (int g = tid / BE; g < BE / (2 * KROW); g += kT2 / BE) {
            const int kg0 = KROW * g, c0 = (kg0 / HR) * MPL, kl0 = kg0 % HR;
            if (TREE && !sp(r / MPL, kg0 / HR)) continue;
            const T *src = Tm + c0 * LDT + r;
            T win[WR];
#pragma unroll
            for (int j = 0; j < WR; ++j) win[j] = src[((2 * kl0 + j) & (MPL - 1)) * LDT];
            T *dst = A + c0 * LDA + r;
#pragma unroll
            for (int p = 0; p < KROW; ++p) {
                T lo, hi;
                dwt_dots<T, F>(&win[2 * p], tp, lo, hi);
                dst[(kl0 + p) * LDA] = lo;
                dst[(HR + ((kl0 + S + p) & (HR - 1))) * LDA] = hi;
            }
        }
    }
    __syncthreads();
}

// BE > 0: square block of edge BE known at compile time (padded leading dimensions, sliding-window column pass)
template <typename T, int F, int BE, bool TREE = false, bool WIDE = false>
__global__ void __launch_bounds__(kT2, (BlkCfg<F, WIDE>::MINB)) wpd2d_block_k(Io2D<T> io, int m, int n, int db, int dend, Taps<T> tp)
{
    using P2 = typename Pair<T>::type;
    constexpr int S = (F - 2) / 2, KROW = wx_krow(F);
    extern __shared__ __align__(16) unsigned char wx_2d_smem[];
    const int BR = BE > 0 ? BE : (m >> db), BC = BE > 0 ? BE : (n >> db);
    const int LDA = BE > 0 ? wx_ld_pairs(BR) : BR, LDT = BE > 0 ? wx_ld_odd(BR) : BR;
    T *A = reinterpret_cast<T *>(wx_2d_smem);            // (BR, BC) column-major: the block at the current level
    T *Tm = A + LDA * BC;                                 // after the column pass
    const int tid = threadIdx.x;
    // grid.x = image * nodes + node row, grid.y = node column
    const unsigned k = blockIdx.x >> db;
    const int jr = (int)(blockIdx.x & ((1u << db) - 1)), jc = (int)blockIdx.y;
    const long img = (long)m * n;
    T *yk = io.out + (long)k * io.ostride;
    const bool from_x = io.copy != nullptr;
    const long org = (long)(jc * BC) * m + jr * BR;       // block origin in the image
    const T *par = io.par + (long)k * io.pstride + org;
    const int BR2 = BR / 2;
    const int lane = tid & 31, warp = tid >> 5;
    // global <-> shared copies of the block: lanes walk the row pairs of a column when a column is at most one warp wide
    const bool colwarp = BR2 <= 32 && (32 % BR2) == 0;
    const int cpw = colwarp ? 32 / BR2 : 1, ca = 2 * (lane % BR2), cb0 = warp * cpw + lane / BR2, cbs = (kT2 / 32) * cpw;

    if (colwarp) { for (int b = cb0; b < BC; b += cbs) cp_async_pair<T>(A + b * LDA + ca, par + b * m + ca); }
    else { for (Walk2 w(tid, BR2); w.hi < BC; w.next()) cp_async_pair<T>(A + w.hi * LDA + 2 * w.lo, par + w.hi * m + 2 * w.lo); }
    cp_async_wait_all();
    __syncthreads();
    if (from_x) {                                         // y[:,:,1] = x   DWT.jl:176
        T *y0 = io.copy + (long)k * io.cstride + org;
        if (colwarp) { for (int b = cb0; b < BC; b += cbs) *reinterpret_cast<P2 *>(y0 + b * m + ca) = *reinterpret_cast<const P2 *>(A + b * LDA + ca); }
        else { for (Walk2 w(tid, BR2); w.hi < BC; w.next()) *reinterpret_cast<P2 *>(y0 + w.hi * m + 2 * w.lo) = *reinterpret_cast<const P2 *>(A + w.hi * LDA + 2 * w.lo); }
    }
    for (int l = db; l < dend; ++l) {
        const int mpl = m >> l, npl = n >> l, hr = mpl / 2, hc = npl / 2;
        const Split2 sp{TREE ? io.tree : nullptr, io.ntree, l, jr << (l - db), jc << (l - db)};
        if (BE == 64 && kT2 % 64 == 0 && mpl == npl && mpl >= 16) {
            // compile-time-shaped levels (node edge 64, 32, 16): both passes with constant geometry
            if (mpl == 64) wpd2d_block_level_ct<T, F, (BE > 0 ? BE : 64), 64, TREE, WIDE>(A, Tm, tp, tid, sp);
            else if (mpl == 32) wpd2d_block_level_ct<T, F, (BE > 0 ? BE : 64), 32, TREE, WIDE>(A, Tm, tp, tid, sp);
            else wpd2d_block_level_ct<T, F, (BE > 0 ? BE : 64), 16, TREE, WIDE>(A, Tm, tp, tid, sp);
            if (!TREE) {
                T *ynext = yk + (long)(l + 1) * img + org;
                for (int b = cb0; b < BC; b += cbs) *reinterpret_cast<P2 *>(ynext + b * m + ca) = *reinterpret_cast<const P2 *>(A + b * LDA + ca);
            }
            continue;
        }
        // ---- column pass A -> Tm, per node: scaling rows on top, detail rows below (shift resolved here) ----
        constexpr int WS = 2 * KSEG + F - 2;
        if (BE > 0 && hr % KSEG == 0 && mpl >= WS) {
            // lanes walk the columns; a thread slides one window down KSEG output pairs of one node
            const FastDiv dq(hr);
            for (int t = tid; t < (BR2 / KSEG) * BC; t += kT2) {
                const int sg = t / BC, c = t - sg * BC, ig0 = sg * KSEG;
                const int jn = dq.div(ig0), il0 = ig0 - jn * hr, r0 = jn * mpl;
                if (TREE && !sp(jn, c / npl)) continue;
                T win[WS];
                load_window_pairs<T, WS>(win, A + c * LDA + r0, 2 * il0, mpl);
                T *dst = Tm + c * LDT + r0;
                int ih = il0 + S; if (ih >= hr) ih %= hr;
#pragma unroll
                for (int p = 0; p < KSEG; ++p) {
                    T lo, hi;
                    dwt_dots<T, F>(&win[2 * p], tp, lo, hi);
                    dst[il0 + p] = lo;
                    dst[hr + ih] = hi;
                    if (++ih >= hr) ih -= hr;
                }
            }
        } else if (BR2 % 2 == 0) {                         // two pairs per thread: ig and ig + BR2/2 (consecutive lanes, consecutive pairs)
            const FastDiv dq(hr);
            for (Walk2 w(tid, BR2 / 2); w.hi < BC; w.next()) {
                const int b = w.hi;
                T lo[2], hi[2];
                int il[2], r0[2];
                bool on[2];
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    const int ig = w.lo + t * (BR2 / 2), jn = dq.div(ig);
                    il[t] = ig - jn * hr; r0[t] = jn * mpl;
                    on[t] = !TREE || sp(jn, b / npl);
                    T win[F];
                    load_window_pairs<T, F>(win, A + b * LDA + r0[t], 2 * il[t], mpl);
                    dwt_dots<T, F>(win, tp, lo[t], hi[t]);
                }
#pragma unroll
                for (int t = 0; t < 2; ++t) {
                    if (!on[t]) continue;
                    T *dst = Tm + b * LDT + r0[t];
                    dst[il[t]] = lo[t];
                    int ih = il[t] + S; if (ih >= hr) ih %= hr;
                    dst[hr + ih] = hi[t];
                }
            }
        } else {
            const FastDiv dq(hr);
            for (Walk2 w(tid, BR2); w.hi < BC; w.next()) {
                const int b = w.hi, jn = dq.div(w.lo), il = w.lo - jn * hr, r0 = jn * mpl;
                if (TREE && !sp(jn, b / npl)) continue;
                T win[F];
                load_window_pairs<T, F>(win, A + b * LDA + r0, 2 * il, mpl);
                T lo, hi;
                dwt_dots<T, F>(win, tp, lo, hi);
                T *dst = Tm + b * LDT + r0;
                dst[il] = lo;
                dst[hr + (il + S) % hr] = hi;
            }
        }
        __syncthreads();
        // ---- row pass Tm -> A, per node: scaling columns left, detail columns right ----
        if ((kT2 % BR) == 0 && (hc % KROW) == 0 && npl >= F) {
            // a thread owns one row r and KROW consecutive output columns of one node: sliding window along the row
            const int r = tid % BR;
            for (int g = tid / BR; g < BC / (2 * KROW); g += kT2 / BR) {
                const int kg0 = KROW * g, jn = kg0 / hc, kl0 = kg0 - jn * hc, c0 = jn * npl;
                if (TREE && !sp(r / mpl, jn)) continue;
                T win[2 * KROW + F - 2];
                const T *src = Tm + c0 * LDT + r;
#pragma unroll
                for (int j = 0; j < 2 * KROW + F - 2; ++j) { int cc = 2 * kl0 + j; if (cc >= npl) cc -= npl; win[j] = src[cc * LDT]; }
                int kh = kl0 + S; if (kh >= hc) kh %= hc;
                T *dst = A + c0 * LDA + r;
#pragma unroll
                for (int p = 0; p < KROW; ++p) {
                    T lo, hi;
                    dwt_dots<T, F>(&win[2 * p], tp, lo, hi);
                    dst[(kl0 + p) * LDA] = lo;
                    dst[(hc + kh) * LDA] = hi;
                    ++kh; if (kh >= hc) kh -= hc;
                }
            }
        } else if (hc % 2 == 0) {
            const FastDiv dq(hc / 2);
            for (Walk2 w(tid, BR); w.hi < BC / 4; w.next()) {
                const int r = w.lo, jn = dq.div(w.hi), kl = 2 * (w.hi - jn * (hc / 2)), c0 = jn * npl;
                if (TREE && !sp(r / mpl, jn)) continue;
                T win[F + 2];
                load_window<T, F + 2>(win, Tm + c0 * LDT + r, 2 * kl, npl, LDT);
                T lo0, hi0, lo1, hi1;
                dwt_dots2<T, F>(win, tp, lo0, hi0, lo1, hi1);
                T *dst = A + c0 * LDA + r;
                dst[kl * LDA] = lo0; dst[(kl + 1) * LDA] = lo1;
                int kh = kl + S; if (kh >= hc) kh %= hc;
                dst[(hc + kh) * LDA] = hi0;
                ++kh; if (kh >= hc) kh -= hc;
                dst[(hc + kh) * LDA] = hi1;
            }
        } else {
            const FastDiv dq(hc);
            for (Walk2 w(tid, BR); w.hi < BC / 2; w.next()) {
                const int r = w.lo, jn = dq.div(w.hi), kl = w.hi - jn * hc, c0 = jn * npl;
                if (TREE && !sp(r / mpl, jn)) continue;
                T win[F];
                load_window<T, F>(win, Tm + c0 * LDT + r, 2 * kl, npl, LDT);
                T lo, hi;
                dwt_dots<T, F>(win, tp, lo, hi);
                T *dst = A + c0 * LDA + r;
                dst[kl * LDA] = lo;
                dst[(hc + (kl + S) % hc) * LDA] = hi;
            }
        }
        __syncthreads();
        // ---- level l+1 slice ----
        if (TREE) continue;
        T *ynext = yk + (long)(l + 1) * img + org;
        if (colwarp) { for (int b = cb0; b < BC; b += cbs) *reinterpret_cast<P2 *>(ynext + b * m + ca) = *reinterpret_cast<const P2 *>(A + b * LDA + ca); }
        else { for (Walk2 w(tid, BR2); w.hi < BC; w.next()) *reinterpret_cast<P2 *>(ynext + w.hi * m + 2 * w.lo) = *reinterpret_cast<const P2 *>(A + w.hi * LDA + 2 * w.lo); }
        // the next column pass only reads A (complete after the barrier above) and writes Tm (free): no barrier needed here
    }
    if (TREE) {                                           // by tree: only the coefficients of the last level leave
        T *yo = yk + org;
        if (colwarp) { for (int b = cb0; b < BC; b += cbs) *reinterpret_cast<P2 *>(yo + b * m + ca) = *reinterpret_cast<const P2 *>(A + b * LDA + ca); }
        else { for (Walk2 w(tid, BR2); w.hi < BC; w.next()) *reinterpret_cast<P2 *>(yo + w.hi * m + 2 * w.lo) = *reinterpret_cast<const P2 *>(A + w.hi * LDA + 2 * w.lo); }
    }
}

// ---------------------------------------------------------------------------------------------------------
// Two-tap filters (haar): no halo, so a TR x TC tile of the IMAGE determines a (TR>>l) x (TC>>l) patch of every node of every
// level.  One CTA keeps the tile in shared memory, runs all L levels there (one 2 x 2 butterfly per thread per step: column pass
// and row pass fused in registers, same tap order as the separate passes) and scatters each level's patches to its slice: the
// whole transform moves exactly the algorithmic (L+2) image sizes.  TR, TC are powers of two, TR >> L >= 2.
// ---------------------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kT2) wpd2d_haar_k(T *__restrict__ y, const T *__restrict__ x, int m, int n, int L, int lr, int lc, Div32 dtiles, Div32 dtiles_r,
                                                   Taps<T> tp)
{
    using P2 = typename Pair<T>::type;
    extern __shared__ __align__(16) unsigned char wx_2d_smem[];
    const int TR = 1 << lr, TC = 1 << lc, TR2 = TR >> 1;
    T *A = reinterpret_cast<T *>(wx_2d_smem);
    T *B = A + TR * TC;
    const int tid = threadIdx.x;
    const unsigned k = div32(blockIdx.x, dtiles), tl = blockIdx.x - k * dtiles.d;
    const unsigned tcx = div32(tl, dtiles_r), trx = tl - tcx * dtiles_r.d;          // neighbouring CTAs walk down the rows
    const int R0 = (int)trx << lr, C0 = (int)tcx << lc;
    const long img = (long)m * n;
    T *yk = y + (long)k * img * (L + 1);
    const T *xk = x + (long)k * img + (long)C0 * m + R0;
    const int npairs = TR2 * TC;
    for (int q = tid; q < npairs; q += kT2) {
        const int c = q >> (lr - 1), r2 = q & (TR2 - 1);
        cp_async_pair<T>(A + c * TR + 2 * r2, xk + (long)c * m + 2 * r2);
    }
    cp_async_wait_all();
    __syncthreads();
    {                                                     // y[:,:,1] = x   DWT.jl:176
        T *y0 = yk + (long)C0 * m + R0;
        for (int q = tid; q < npairs; q += kT2) {
            const int c = q >> (lr - 1), r2 = q & (TR2 - 1);
            *reinterpret_cast<P2 *>(y0 + (long)c * m + 2 * r2) = *reinterpret_cast<const P2 *>(A + c * TR + 2 * r2);
        }
    }
    T *src = A, *dst = B;
    for (int l = 0; l < L; ++l) {
        // sub-block (a, b) of the tile at this level: rows a*mpl .., cols b*npl ..; its four children land in its quadrants
        const int lm = lr - l, ln = lc - l;               // log2 of the sub-block extents
        const int hr = 1 << (lm - 1), hc = 1 << (ln - 1);
        const int nblk = TR2 * (TC >> 1);
        for (int q = tid; q < nblk; q += kT2) {
            const int ig = q & (TR2 - 1), jg = q >> (lr - 1);
            const int a = ig >> (lm - 1), il = ig & (hr - 1), b = jg >> (ln - 1), jl = jg & (hc - 1);
            const P2 v0 = *reinterpret_cast<const P2 *>(src + (2 * jg) * TR + 2 * ig);
            const P2 v1 = *reinterpret_cast<const P2 *>(src + (2 * jg + 1) * TR + 2 * ig);
            T w[2], lo0, hi0, lo1, hi1, ll, lh, hl, hh;
            w[0] = v0.x; w[1] = v0.y; dwt_dots<T, 2>(w, tp, lo0, hi0);          // column pass, column 2jg
            w[0] = v1.x; w[1] = v1.y; dwt_dots<T, 2>(w, tp, lo1, hi1);          // column 2jg+1
            w[0] = lo0; w[1] = lo1; dwt_dots<T, 2>(w, tp, ll, lh);              // row pass over the scaling rows
            w[0] = hi0; w[1] = hi1; dwt_dots<T, 2>(w, tp, hl, hh);              // ... and the detail rows
            T *o = dst + ((b << ln) + jl) * TR + (a << lm) + il;
            o[0] = ll; o[hr] = hl; o[hc * TR] = lh; o[hc * TR + hr] = hh;
        }
        __syncthreads();
        // ---- scatter the level l+1 patches: sub-block (a', b') of extent (TR >> (l+1)) x (TC >> (l+1)) -> node (a', b') ----
        const int lv = l + 1, er = lr - lv, ec = lc - lv; // log2 extents
        const int mn = m >> lv, nn = n >> lv, r0 = R0 >> lv, c0 = C0 >> lv;
        T *yl = yk + (long)lv * img;
        for (int q = tid; q < npairs; q += kT2) {
            const int c = q >> (lr - 1), r = 2 * (q & (TR2 - 1));
            const int a = r >> er, ri = r & ((1 << er) - 1), b = c >> ec, ci = c & ((1 << ec) - 1);
            *reinterpret_cast<P2 *>(yl + (long)(b * nn + c0 + ci) * m + a * mn + r0 + ri) = *reinterpret_cast<const P2 *>(dst + c * TR + r);
        }
        T *t2 = src; src = dst; dst = t2;                 // the next step reads what this one wrote (complete after the barrier above);
        // its writes go to the buffer the scatter above does not read -- no barrier needed here
    }
}

template <typename T>
int wpd2d_haar_run(T *y, const T *x, long m, long n, int L, long N, const Taps<T> &t, cudaStream_t s, bool *done)
{
    *done = false;
    static const bool off = getenv("WX_B200_NO_HAAR2D") != nullptr;
    if (off) return WX_OK;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    int lr = sizeof(T) == 8 ? 7 : 8, lc = 5;              // 128 (Float64) / 256 (Float32) x 32 tile: 64 KB for the two buffers, three CTAs
                                                          // per SM; the deepest level of L = 5 still stores 32-byte runs
    while ((1 << lc) < (1 << L)) ++lc;                    // every level needs at least one column ...
    while ((1 << lr) < (2 << L)) ++lr;                    // ... and one row PAIR per node patch (16-byte stores)
    while (lr > 1 && (m % (1L << lr)) != 0 && (1 << (lr - 1)) >= (2 << L)) --lr;
    while (lc > 0 && (n % (1L << lc)) != 0 && (1 << (lc - 1)) >= (1 << L)) --lc;
    if (m % (1L << lr) != 0 || n % (1L << lc) != 0) return WX_OK;
    const size_t smem = (size_t)2 * sizeof(T) << (lr + lc);
    if (smem > dv.smem_optin) return WX_OK;
    const long tiles_r = m >> lr, tiles = tiles_r * (n >> lc);
    if (tiles * N >= (1L << 31)) return WX_OK;
    auto kern = wpd2d_haar_k<T>;
    // Residency: one CTA per tile, hardware scheduled, three per SM by shared memory.  Capping it (by asking for more dynamic shared
    // memory, knob WX_B200_HAAR2D_OCC) was measured in round 2: 9.39 ms with three, 9.97 with two, 13.3 with one (Float32 5.98 / 6.41 /
    // 8.60) -- unlike the persistent 1-D kernels this one wants everything resident, so it is not tuned at run time.  (A first
    // version let wx_tuned_choice try the candidates: the launches with inflated shared memory left every later launch ~7 % slower.)
    size_t ask = smem;
    if (const char *oenv = getenv("WX_B200_HAAR2D_OCC")) {
        const int occ = atoi(oenv), occfit = (int)(dv.smem_optin / (smem + 1024));
        if (occ >= 1 && occ < occfit) { ask = (dv.smem_optin / (size_t)occ - 1024) & ~(size_t)127; if (ask < smem) ask = smem; }
    }
    WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ask));
    kern<<<(unsigned)(tiles * N), kT2, ask, s>>>(y, x, (int)m, (int)n, L, lr, lc, make_div32(tiles), make_div32(tiles_r), t);
    WX_LAUNCHED();
    *done = true;
    return WX_OK;
}

static int largest_divisor_le(long v, int cap)
{
    int best = 1;
    for (int t = 1; t <= cap && t <= v; ++t) if (v % t == 0) best = t;
    return best;
}

// TREE = false: y = packet table (m,n,L+1,N).  TREE = true: y = coefficient images (m,n,N) along the quad tree of depth L (dtree on the
// device, nullptr = complete), scratch = N more images; the launches ping-pong between y and scratch so that the last one writes y.
template <typename T, int F, bool TREE>
int wpd2d_run_chunk(T *y, const T *x, T *scratch, long m, long n, int L, long N, const unsigned char *dtree, long ntree, const Taps<T> &t,
                    cudaStream_t s)
{
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    // depth from which a whole node fits the block kernel's two buffers (64 KB keeps three CTAs per SM)
    const size_t budget = 65536;
    const long img = m * n;
    int db = 0;
    while (db < L && (size_t)2 * (m >> db) * (n >> db) * sizeof(T) > budget) ++db;
    int left = db + (db < L ? 1 : 0);                          // launches to go (TREE: ping-pong parity)
    const T *cur = x;
    auto level_io = [&](int d) {
        Io2D<T> io;
        if (TREE) {
            T *dst = (left % 2 == 1) ? y : scratch;
            io.par = cur; io.pstride = img; io.out = dst; io.ostride = img; io.copy = nullptr; io.cstride = 0; io.tree = dtree; io.ntree = ntree;
            cur = dst;
        } else {
            io.par = d == 0 ? x : y + (long)d * img; io.pstride = d == 0 ? img : img * (L + 1);
            io.out = y + (long)(d + 1) * img; io.ostride = img * (L + 1);
            io.copy = d == 0 ? y : nullptr; io.cstride = img * (L + 1); io.tree = nullptr; io.ntree = 0;
        }
        --left;
        return io;
    };
    static const char *env = getenv("WX_B200_WPD2D_TILE");      // measurement knob: tile edge of the halo kernel
    const int cap = (env && atoi(env) >= 1 && atoi(env) <= 32) ? atoi(env) : 32;     // the kernel maps one lane per row pair: tr <= 32
    for (int d = 0; d < db; ++d) {
        const long hr = (m >> d) / 2, hc = (n >> d) / 2;
        // Narrow tiles (32 x 16, knob WX_B200_WPD2D_NARROW): 41 KB instead of 76 KB of shared memory and a 64-register build, i.e. up
        // to five resident CTAs instead of three for the latency-bound tile pass, at the price of a wider relative halo
        static const char *nenv = getenv("WX_B200_WPD2D_NARROW");
        const int narrow = (nenv && !TREE) ? atoi(nenv) : 0;
        const int tr = largest_divisor_le(hr, cap);
        int tc = largest_divisor_le(hc, cap);
        if (narrow && tr == 32 && tc == 32) tc = 16;
        const int PR = 2 * tr + F - 2, PC = 2 * tc + F - 2;
        const bool shaped = (tr == 32 && (tc == 32 || tc == 16));          // the compile-time-shaped instantiations (padded leading dimensions)
        const size_t smem = shaped ? ((size_t)wx_ld_pairs(PR) * PC + (size_t)wx_ld_odd(2 * tr) * PC) * sizeof(T)
                                   : ((size_t)(PR + 2 * (tr & 1)) * PC + (size_t)2 * tr * (PC + 2 * (tc & 1))) * sizeof(T);
        if (smem > dv.smem_optin) return wx_fail(WX_EUNSUPPORTED, "wpd 2-D tile does not fit shared memory");
        const long gx = (hr / tr) * (1L << d) * N, gy = (hc / tc) * (1L << d);
        if (gx * gy >= (1L << 31)) return wx_fail(WX_EUNSUPPORTED, "wpd 2-D: too many tiles for one launch");
        const Div32 drt = make_div32((hr / tr) * (1L << d)), dtr = make_div32(hr / tr), dtc = make_div32(hc / tc), dgx = make_div32(gx);
        const long ntiles = gx * gy;
        // Measurement knob: CTAs per SM of a persistent tile loop that prefetches the next tile's patch under the current row pass.
        // Measured in round 1 (4096 x 512^2, db4 F64): 3 per SM 18.7 ms, 6 per SM 18.3 ms, one CTA per tile 16.7 ms -- the hardware
        // scheduler refilling three CTA slots overlaps loads and passes better than the in-CTA pipeline, so the default is 0.
        static const char *penv = getenv("WX_B200_WPD2D_PERSIST");
        const int per_sm = penv ? atoi(penv) : 0;
        const long ctas = per_sm > 0 && ntiles > (long)dv.sms * per_sm ? (long)dv.sms * per_sm : ntiles;
        const Io2D<T> io = level_io(d);
        static const char *twenv = getenv("WX_B200_WPD2D_TILEWIDE");      // A-B knob
        // 8-pair row window under a two-CTA bound: only 20 taps in Float64 gain (7.44 -> 6.77 ms per 1024 images; Float32 loses 10-17 %)
        const bool wide_tile = F >= 12 && (twenv ? atoi(twenv) != 0 : (sizeof(T) == 8 && F >= 20));
        if (tr == 32 && tc == 16 && !TREE) {
            auto kern = wpd2d_tile_k<T, F, 32, 16, 4, false>;
            WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)ctas, kT2, smem, s>>>(io, (int)m, (int)n, d, tr, tc, drt, dtr, dtc, dgx, ntiles, t);
        } else if (tr == 32 && tc == 32 && wide_tile) {
            auto kern = wpd2d_tile_k<T, F, 32, 32, 2, TREE, (F >= 12 ? 8 : 0)>;
            WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)ctas, kT2, smem, s>>>(io, (int)m, (int)n, d, tr, tc, drt, dtr, dtc, dgx, ntiles, t);
        } else if (tr == 32 && tc == 32) {
            auto kern = wpd2d_tile_k<T, F, 32, 32, 3, TREE>;
            WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)ctas, kT2, smem, s>>>(io, (int)m, (int)n, d, tr, tc, drt, dtr, dtc, dgx, ntiles, t);
        } else {
            auto kern = wpd2d_tile_k<T, F, 0, 0, 3, TREE>;
            WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<(unsigned)ctas, kT2, smem, s>>>(io, (int)m, (int)n, d, tr, tc, drt, dtr, dtc, dgx, ntiles, t);
        }
        WX_LAUNCHED();
    }
    if (db < L) {
        const bool shaped = (m >> db) == 64 && (n >> db) == 64;
        const size_t smem = shaped ? (size_t)(wx_ld_pairs(64) + wx_ld_odd(64)) * 64 * sizeof(T) : (size_t)2 * (m >> db) * (n >> db) * sizeof(T);
        const long gx = (1L << db) * N, gy = 1L << db;
        if (gx >= (1L << 31) || gy > 65535) return wx_fail(WX_EUNSUPPORTED, "wpd 2-D: too many blocks for one launch");
        Io2D<T> io = level_io(db);
        if (!TREE) { io.out = y; }                            // the whole-node kernel adds (l + 1) * img per level itself
        // measured (profiles/r2_wpd2d_blkwide_ab.jsonl, 1024 x 512^2): Float64 10 / 12 / 16 / 20 taps 5.12 / 5.74 / 7.13 / 8.91 -> 4.89 / 4.96 / 5.83 / 7.40 ms,
        // Float32 12 / 16 / 20 taps 3.11 / 3.69 / 4.32 -> 3.06 / 3.41 / 4.12 ms, Float32 10 taps loses (2.86 -> 2.94 ms)
        static const char *wenv = getenv("WX_B200_WPD2D_BLKWIDE");        // A-B knob: 0 = never, 1 = every filter from 10 taps, unset = the rule
        const bool wide = F >= 10 && (wenv ? atoi(wenv) != 0 : (sizeof(T) == 8 || F >= 12));
        if (shaped && wide) {
            auto kern = wpd2d_block_k<T, F, 64, TREE, (F >= 10)>;
            WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<dim3((unsigned)gx, (unsigned)gy), kT2, smem, s>>>(io, (int)m, (int)n, db, L, t);
        } else if (shaped) {
            auto kern = wpd2d_block_k<T, F, 64, TREE>;
            WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<dim3((unsigned)gx, (unsigned)gy), kT2, smem, s>>>(io, (int)m, (int)n, db, L, t);
        } else {
            auto kern = wpd2d_block_k<T, F, 0, TREE>;
            WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            kern<<<dim3((unsigned)gx, (unsigned)gy), kT2, smem, s>>>(io, (int)m, (int)n, db, L, t);
        }
        WX_LAUNCHED();
    }
    return WX_OK;
}

// Optional chunking of the batch (measurement knob WX_B200_WPD2D_CHUNK = images per chunk).  The idea -- keep the level a
// launch writes resident in the 126 MB L2 for the launch that reads it -- was measured in round 1 and LOSES: with 8..64 images
// per chunk the extra launches and partial waves cost more than the saved DRAM reads (19.1 ms -> 21.2..32.5 ms, db4 F64,
// 4096 x 512^2), so the default is one launch per level for the whole batch.
template <typename T, int F>
int wpd2d_run(T *y, const T *x, long m, long n, int L, long N, const Taps<T> &t, cudaStream_t s)
{
    static const char *env = getenv("WX_B200_WPD2D_CHUNK");
    long chunk = (env && atol(env) > 0) ? atol(env) : N;
    if (chunk < 1) chunk = 1;
    if (chunk > N) chunk = N;
    const long img = m * n;
    for (long k0 = 0; k0 < N; k0 += chunk) {
        const long nk = (N - k0 < chunk) ? N - k0 : chunk;
        int rc = wpd2d_run_chunk<T, F, false>(y + k0 * img * (L + 1), x + k0 * img, nullptr, m, n, L, nk, nullptr, 0, t, s);
        if (rc) return rc;
    }
    return WX_OK;
}

}  // namespace

// x(m,n,N) -> y(m,n,L+1,N) including the level-0 copy.  *handled = false: shape not covered (caller runs the per-level path).
template <typename T>
int wx_wpd2d_fused(T *y, const T *x, long m, long n, int L, long N, const Taps<T> &t, cudaStream_t s, bool *handled)
{
    *handled = false;
    static const bool off = getenv("WX_B200_NO_FUSED_WPD2D") != nullptr;   // debugging / A-B measurements only
    if (off || L < 1 || N < 1 || m * n >= (1L << 31)) return WX_OK;
    if (((((uintptr_t)y) | ((uintptr_t)x)) & 15) != 0) return WX_OK;
    if ((m >> (L - 1)) % 2 != 0 || (n >> (L - 1)) % 2 != 0) return WX_OK;
    int rc;
    if (t.F == 2) {
        bool done = false;
        rc = wpd2d_haar_run<T>(y, x, m, n, L, N, t, s, &done);
        if (rc) return rc;
        if (done) { *handled = true; return WX_OK; }
    }
#define WX_2D_CASE(FF) case FF: rc = wpd2d_run<T, FF>(y, x, m, n, L, N, t, s); break;
    switch (t.F) {
        WX_2D_CASE(2) WX_2D_CASE(4) WX_2D_CASE(6) WX_2D_CASE(8) WX_2D_CASE(10) WX_2D_CASE(12) WX_2D_CASE(16) WX_2D_CASE(20)
        default: return WX_OK;
    }
#undef WX_2D_CASE
    if (rc == WX_EUNSUPPORTED) return WX_OK;          // the per-level path recomputes everything
    if (rc == WX_OK) *handled = true;
    return rc;
}
template int wx_wpd2d_fused<double>(double *, const double *, long, long, int, long, const Taps<double> &, cudaStream_t, bool *);
template int wx_wpd2d_fused<float>(float *, const float *, long, long, int, long, const Taps<float> &, cudaStream_t, bool *);

// x(m,n,N) -> y(m,n,N) along the quad tree of depth nlev (wpt! 2-D DWT.jl:500-548); dtree = device copy of the tree or nullptr for a
// complete one; scratch: N images.  One halo-tile launch per level whose nodes exceed shared memory, then the whole-node kernel for all
// deeper levels: min(nlev, db) + 1 round trips through HBM instead of the packet table + leaf gather.
template <typename T>
int wx_wpt2d_fused(T *y, const T *x, T *scratch, long m, long n, int nlev, long N, const unsigned char *dtree, long ntree, const Taps<T> &t,
                   cudaStream_t s, bool *handled)
{
    *handled = false;
    static const bool off = getenv("WX_B200_NO_FUSED_WPD2D") != nullptr || getenv("WX_B200_WPT2D_TABLE") != nullptr;
    if (off || nlev < 1 || nlev > 30 || N < 1 || m * n >= (1L << 31) || y == x) return WX_OK;
    if (((((uintptr_t)y) | ((uintptr_t)x) | ((uintptr_t)scratch)) & 15) != 0) return WX_OK;
    if ((m >> (nlev - 1)) % 2 != 0 || (n >> (nlev - 1)) % 2 != 0) return WX_OK;
    int rc;
#define WX_2D_CASE(FF) case FF: rc = wpd2d_run_chunk<T, FF, true>(y, x, scratch, m, n, nlev, N, dtree, ntree, t, s); break;
    switch (t.F) {
        WX_2D_CASE(2) WX_2D_CASE(4) WX_2D_CASE(6) WX_2D_CASE(8) WX_2D_CASE(10) WX_2D_CASE(12) WX_2D_CASE(16) WX_2D_CASE(20)
        default: return WX_OK;
    }
#undef WX_2D_CASE
    if (rc == WX_EUNSUPPORTED) return WX_OK;          // y is recomputed from x by the caller's other paths
    if (rc == WX_OK) *handled = true;
    return rc;
}
template int wx_wpt2d_fused<double>(double *, const double *, double *, long, long, int, long, const unsigned char *, long, const Taps<double> &, cudaStream_t, bool *);
template int wx_wpt2d_fused<float>(float *, const float *, float *, long, long, int, long, const unsigned char *, long, const Taps<float> &, cudaStream_t, bool *);

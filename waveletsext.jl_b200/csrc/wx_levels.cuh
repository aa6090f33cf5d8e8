// wx_levels.cuh -- one decomposition / reconstruction level of every node of a signal staged in shared memory
// (16-byte-chunk XOR swizzle, see wx_common.cuh).  Shared by the fused wpd kernel (wx_wpd1d.cu) and the fused
// tree kernels (wx_tree1d.inl).
//   forward: dwt_step!  dwt/dwt_one_level.jl:79-107      inverse: idwt_step!  dwt/dwt_one_level.jl:192-223
// TREE = true: `tm[j]` (j < ntm) says whether node j of this depth is split; other nodes are copied through.
#pragma once
#include "wx_common.cuh"

template <typename T, int F, int KM = 2>
struct WpdCfg {
    static constexpr int V = WxVec<T>::N;                              // elements per 16 B chunk
    // output pairs per window.  K = 4V (thread stride of one 128-byte row, fully conflict-free window loads) was measured in
    // round 1: same speed within noise (the kernel is HBM-bound, 5.21 vs 5.13 ms) at 78 instead of 54 registers -- kept at 2V.
    // KM = 4 (chosen by the wpdall launcher for 14+ taps on nodes long enough to keep the CTA busy): the window re-reads
    // (2S + 2K) / K samples per output pair, 3.0 with K = 2V against 2.0 with K = 4V for 16 taps -- long filters are bound by the
    // shared-memory and FP64 pipes together, not by HBM alone.  Measured (profiles/r2_k4v_threshold.jsonl): sym8 F64 n = 4096
    // 5.9 -> 5.4 ms, F32 2.67 -> 2.50 ms; 10-12 taps lose, and so do short nodes (F32 n = 1024: 32 units for a 64-thread CTA).
    static constexpr int K = KM * V;
    static constexpr int S = (((F - 2) / 2) + V - 1) / V * V;          // high-pass look-ahead (multiple of V)
    static constexpr int W = 2 * S + 2 * K;                            // window length (elements)
};

// split flags of the nodes of one depth: node j is split iff j < ntm && tm[j]  (tm == nullptr: every node is split)
struct TreeMask {
    const unsigned char *tm;
    long ntm;
    __device__ __forceinline__ bool on(long j) const { return tm == nullptr || (j < ntm && tm[j]); }
};

template <typename T> __device__ __forceinline__ typename WxVec<T>::type wx_ldg_stream(const T *p);
template <> __device__ __forceinline__ double2 wx_ldg_stream<double>(const double *p) { return __ldcs(reinterpret_cast<const double2 *>(p)); }
template <> __device__ __forceinline__ float4 wx_ldg_stream<float>(const float *p) { return __ldcs(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ void wx_stg_stream(double *p, double2 v) { __stcs(reinterpret_cast<double2 *>(p), v); }
__device__ __forceinline__ void wx_stg_stream(float *p, float4 v) { __stcs(reinterpret_cast<float4 *>(p), v); }

__device__ __forceinline__ void wx_unpack(double *d, double2 v) { d[0] = v.x; d[1] = v.y; }
__device__ __forceinline__ void wx_unpack(float *d, float4 v) { d[0] = v.x; d[1] = v.y; d[2] = v.z; d[3] = v.w; }
__device__ __forceinline__ double2 wx_pack(const double *d) { return make_double2(d[0], d[1]); }
__device__ __forceinline__ float4 wx_pack(const float *d) { return make_float4(d[0], d[1], d[2], d[3]); }

// store one 16 B chunk of outputs (element index e, chunk aligned) to the next-level smem buffer and to HBM
// GST = true : outputs go to HBM straight from registers (and to smem unless this is the last level)
// GST = false: outputs go to smem only; the level row is written to HBM by a TMA bulk store of the smem buffer
template <typename T, bool GST>
__device__ __forceinline__ void wx_put_chunk(T *dst, int soff, T *grow, int e, const T *vals, bool last)   // soff = swizzled element offset of e
{
    using VT = typename WxVec<T>::type;
    VT v = wx_pack(vals);
    if (!GST || !last) *reinterpret_cast<VT *>(dst + soff) = v;
    if (GST) wx_stg_stream(grow + e, v);
}
template <typename T, bool GST>
__device__ __forceinline__ void wx_put_chunk(T *dst, T *grow, int e, const T *vals, bool last)
{
    constexpr int V = WxVec<T>::N;
    using VT = typename WxVec<T>::type;
    VT v = wx_pack(vals);
    if (!GST || !last) *reinterpret_cast<VT *>(dst + wx_swz_chunk(e / V) * V) = v;
    if (GST) wx_stg_stream(grow + e, v);
}

// element offset of the swizzled position of element e, for e a multiple of 4 chunks: the other chunks of that aligned
// group of four (eight when e is a multiple of 8 chunks) are at  wx_swz_group(e) ^ (i * V)  -- one XOR per chunk.
template <typename T>
__device__ __forceinline__ int wx_swz_group(int e)
{
    constexpr int V = WxVec<T>::N;
    return wx_swz_chunk(e / V) * V;
}

// ---- wide level: node half-length is a multiple of K -------------------------------------------------
template <typename T, int F, bool POW2, bool GST, bool TREE = false, int KM = 2>
__device__ __forceinline__ void wpd_wide_level(const T *__restrict__ src, T *__restrict__ dst, T *__restrict__ grow, int n0, int p,
                                               bool last, const Taps<T> &tp, int tid, int nthreads, TreeMask tmk = TreeMask{nullptr, 0})
{
    using C = WpdCfg<T, F, KM>;
    using VT = typename WxVec<T>::type;
    constexpr int V = C::V, K = C::K, S = C::S, W = C::W;
    const int half = p >> 1;
    const int units = n0 / (2 * K);
    const int lgh = 31 - __clz(half);
    // Lanes -> units.  A unit reads the 4 chunks it owns plus the first chunks of the NEXT unit (the filter's look-ahead) and
    // stores its detail outputs S/V chunks ahead.  With the 128-byte XOR swizzle, 8 consecutive units are conflict free on
    // their own chunks but not on the neighbour's (units u+1 .. u+8 straddle two swizzle periods): ncu showed 6 instead of 4
    // wavefronts on half of the window loads and 7 on the detail stores.  Letting a quarter warp take the EVEN units of a
    // group of 16 and the next quarter the ODD ones makes both the own and the neighbour's chunks hit 8 distinct bank groups.
    const bool perm16 = (2 * K / V == 4) && (units & 15) == 0;          // units of 8 chunks are conflict free in lane order
    for (int u0 = tid; u0 < units; u0 += nthreads) {
        const int u = perm16 ? ((u0 & ~15) | ((u0 & 7) << 1) | ((u0 >> 3) & 1)) : u0;
        const int gi = u * K;
        const int j = POW2 ? (gi >> lgh) : (gi / half);
        const int i = gi - j * half;
        const int base = j * p;
        if (TREE && !tmk.on(j)) {                                      // leaf of the tree: pass the 2K samples through
#pragma unroll
            for (int c = 0; c < 2 * K / V; ++c) {
                const int ch = wx_swz_group<T>(base + 2 * i) ^ (c * V);      // 2K elements = one aligned group of 4 chunks
                *reinterpret_cast<VT *>(dst + ch) = *reinterpret_cast<const VT *>(src + ch);
            }
            continue;
        }
        // the window starts on a 4-chunk boundary (2i is a multiple of 2K); one swizzle per aligned group of 4 chunks
        constexpr int NG = (W / V + 3) / 4;
        int gw[NG];
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            int off = 2 * i + g * 4 * V;
            off = POW2 ? (off & (p - 1)) : (off % p);
            gw[g] = wx_swz_group<T>(base + off);
        }
        T win[W];
#pragma unroll
        for (int c = 0; c < W / V; ++c) wx_unpack(&win[c * V], *reinterpret_cast<const VT *>(src + (gw[c / 4] ^ ((c % 4) * V))));
        T lo[K], hi[K];
#pragma unroll
        for (int k = 0; k < K; ++k) {
            T a = tp.g[F - 1] * win[2 * k];
            T b = tp.h[0] * win[2 * (S + k) + 1];
#pragma unroll
            for (int jj = 1; jj < F; ++jj) {
                a = fma(tp.g[F - 1 - jj], win[2 * k + jj], a);
                b = fma(tp.h[jj], win[2 * (S + k) + 1 - jj], b);
            }
            lo[k] = a;
            hi[k] = b;
        }
#pragma unroll
        for (int c = 0; c < K / V; ++c) {                              // K / V = 2 chunks, the low-pass pair is aligned
            wx_put_chunk<T, GST>(dst, wx_swz_group<T>(base + i) ^ (c * V), grow, base + i + c * V, &lo[c * V], last);
            int io = i + S + c * V;
            io = POW2 ? (io & (half - 1)) : (io % half);
            wx_put_chunk<T, GST>(dst, wx_swz_group<T>(base + half + io), grow, base + half + io, &hi[c * V], last);
        }
    }
}

// ---- small level: node length P in {2,4,8}; a thread owns max(P,V) consecutive elements = whole nodes ----
template <typename T, int F, int P, bool GST, bool TREE = false>
__device__ __forceinline__ void wpd_small_level(const T *__restrict__ src, T *__restrict__ dst, T *__restrict__ grow, int n0, bool last,
                                                const Taps<T> &tp, int tid, int nthreads, TreeMask tmk = TreeMask{nullptr, 0})
{
    using VT = typename WxVec<T>::type;
    constexpr int V = WxVec<T>::N;
    constexpr int G = P > V ? P : V;
    const int groups = n0 / G;
    for (int u = tid; u < groups; u += nthreads) {
        const int e0 = u * G;
        T v[G], o[G];
#pragma unroll
        const int g0 = wx_swz_chunk(e0 / V) * V;         // G/V <= 4 chunks, aligned: chunk c sits at g0 ^ (c * V)
#pragma unroll
        for (int c = 0; c < G / V; ++c) wx_unpack(&v[c * V], *reinterpret_cast<const VT *>(src + (g0 ^ (c * V))));
#pragma unroll
        for (int nd = 0; nd < G / P; ++nd) {
            if (TREE && !tmk.on(e0 / P + nd)) {
#pragma unroll
                for (int i = 0; i < P; ++i) o[nd * P + i] = v[nd * P + i];
                continue;
            }
#pragma unroll
            for (int i = 0; i < P / 2; ++i) {
                T a = tp.g[F - 1] * v[nd * P + ((2 * i) & (P - 1))];
                T b = tp.h[0] * v[nd * P + ((2 * i + 1) & (P - 1))];
#pragma unroll
                for (int jj = 1; jj < F; ++jj) {
                    a = fma(tp.g[F - 1 - jj], v[nd * P + ((2 * i + jj) & (P - 1))], a);
                    b = fma(tp.h[jj], v[nd * P + ((2 * i + 1 - jj) & (P - 1))], b);
                }
                o[nd * P + i] = a;
                o[nd * P + P / 2 + i] = b;
            }
        }
#pragma unroll
        for (int c = 0; c < G / V; ++c) wx_put_chunk<T, GST>(dst, g0 ^ (c * V), grow, e0 + c * V, &o[c * V], last);
    }
}

// one forward level of the nodes of length P inside a register-resident run of G elements (compile-time wraps); along a tree the nodes
// that are not split pass through
template <typename T, int F, int P, int G, bool TREE>
__device__ __forceinline__ void wpd_regs_level(const T *v, T *o, const Taps<T> &tp, const TreeMask &tm, long node0)
{
#pragma unroll
    for (int nd = 0; nd < G / P; ++nd) {
        if (!TREE || tm.on(node0 + nd)) {
#pragma unroll
            for (int i = 0; i < P / 2; ++i) {
                T a = tp.g[F - 1] * v[nd * P + ((2 * i) & (P - 1))];
                T b = tp.h[0] * v[nd * P + ((2 * i + 1) & (P - 1))];
#pragma unroll
                for (int jj = 1; jj < F; ++jj) {
                    a = fma(tp.g[F - 1 - jj], v[nd * P + ((2 * i + jj) & (P - 1))], a);
                    b = fma(tp.h[jj], v[nd * P + ((2 * i + 1 - jj) & (P - 1))], b);
                }
                o[nd * P + i] = a;
                o[nd * P + P / 2 + i] = b;
            }
        } else {
#pragma unroll
            for (int i = 0; i < P; ++i) o[nd * P + i] = v[nd * P + i];
        }
    }
}
// the four deepest levels of a by-tree forward transform that goes down to nodes of length 2: a thread owns a node of length 16 and
// everything below it, splits 16 -> 8 -> 4 -> 2 in registers and writes the run once (three shared-memory round trips and barriers fewer).
// tm16 .. tm2: split flags of the depths whose nodes have length 16 .. 2.
template <typename T, int F, bool TREE>
__device__ __forceinline__ void wpd_small_levels4(const T *__restrict__ src, T *__restrict__ dst, int n0, const Taps<T> &tp, int tid, int nthreads,
                                                  TreeMask tm16, TreeMask tm8, TreeMask tm4, TreeMask tm2)
{
    using VT = typename WxVec<T>::type;
    constexpr int V = WxVec<T>::N, G = 16;
    for (int u = tid; u < n0 / G; u += nthreads) {
        const int g0 = wx_swz_chunk((u * G) / V) * V;
        T v[G], o[G];
#pragma unroll
        for (int c = 0; c < G / V; ++c) wx_unpack(&v[c * V], *reinterpret_cast<const VT *>(src + (g0 ^ (c * V))));
        wpd_regs_level<T, F, 16, G, TREE>(v, o, tp, tm16, (long)u);
        wpd_regs_level<T, F, 8, G, TREE>(o, v, tp, tm8, (long)u * 2);
        wpd_regs_level<T, F, 4, G, TREE>(v, o, tp, tm4, (long)u * 4);
        wpd_regs_level<T, F, 2, G, TREE>(o, v, tp, tm2, (long)u * 8);
#pragma unroll
        for (int c = 0; c < G / V; ++c) *reinterpret_cast<VT *>(dst + (g0 ^ (c * V))) = wx_pack(&v[c * V]);
    }
}

// ---- generic level: any even node length, one output pair per thread ----------------------------------
template <typename T, int F, bool GST, bool TREE = false>
__device__ __forceinline__ void wpd_generic_level(const T *__restrict__ src, T *__restrict__ dst, T *__restrict__ grow, int n0, int p,
                                                  bool last, const Taps<T> &tp, int tid, int nthreads, TreeMask tmk = TreeMask{nullptr, 0})
{
    const int half = p >> 1;
    for (int gi = tid; gi < n0 / 2; gi += nthreads) {
        const int j = gi / half;
        const int i = gi - j * half;
        const int base = j * p;
        if (TREE && !tmk.on(j)) {
            dst[wx_swz_elem<T>(base + 2 * i)] = src[wx_swz_elem<T>(base + 2 * i)];
            dst[wx_swz_elem<T>(base + 2 * i + 1)] = src[wx_swz_elem<T>(base + 2 * i + 1)];
            continue;
        }
        int k1 = (2 * i) % p, k2 = (2 * i + 1) % p;
        T a = tp.g[F - 1] * src[wx_swz_elem<T>(base + k1)];
        T b = tp.h[0] * src[wx_swz_elem<T>(base + k2)];
#pragma unroll 4
        for (int jj = 1; jj < F; ++jj) {
            k1 += 1; if (k1 >= p) k1 -= p;
            k2 -= 1; if (k2 < 0) k2 += p;
            a = fma(tp.g[F - 1 - jj], src[wx_swz_elem<T>(base + k1)], a);
            b = fma(tp.h[jj], src[wx_swz_elem<T>(base + k2)], b);
        }
        const int elo = base + i, ehi = base + half + i;
        if (!GST || !last) { dst[wx_swz_elem<T>(elo)] = a; dst[wx_swz_elem<T>(ehi)] = b; }
        if (GST) { grow[elo] = a; grow[ehi] = b; }
    }
}

// one decomposition level of every node of the staged signal: picks the wide / small / generic path
template <typename T, int F, bool GST, bool TREE = false, int KM = 2>
__device__ __forceinline__ void wpd_level(const T *__restrict__ a, T *__restrict__ b, T *__restrict__ grow, int n0, int p, bool last,
                                          const Taps<T> &tp, int tid, int nthreads, TreeMask tmk = TreeMask{nullptr, 0})
{
    using C = WpdCfg<T, F, KM>;
    constexpr int V = C::V, K = C::K;
    const int half = p >> 1;
    const bool pow2 = (p & (p - 1)) == 0;
    if (half % K == 0) {
        if (pow2) wpd_wide_level<T, F, true, GST, TREE, KM>(a, b, grow, n0, p, last, tp, tid, nthreads, tmk);
        else      wpd_wide_level<T, F, false, GST, TREE, KM>(a, b, grow, n0, p, last, tp, tid, nthreads, tmk);
    } else if (p == 2 && n0 % (V > 2 ? V : 2) == 0) {
        wpd_small_level<T, F, 2, GST, TREE>(a, b, grow, n0, last, tp, tid, nthreads, tmk);
    } else if (p == 4) {
        wpd_small_level<T, F, 4, GST, TREE>(a, b, grow, n0, last, tp, tid, nthreads, tmk);
    } else if (p == 8) {
        wpd_small_level<T, F, 8, GST, TREE>(a, b, grow, n0, last, tp, tid, nthreads, tmk);
    } else if (p == 16 && V == 4) {
        wpd_small_level<T, F, 16, GST, TREE>(a, b, grow, n0, last, tp, tid, nthreads, tmk);
    } else {
        wpd_generic_level<T, F, GST, TREE>(a, b, grow, n0, p, last, tp, tid, nthreads, tmk);
    }
}


// =====================================================================================================
// inverse levels: children (w1 = first half, w2 = second half of the node) -> parent, idwt_step!
// dwt/dwt_one_level.jl:207-221.  With R = F/2 and 0-based pair index t (outputs 2t, 2t+1):
//   v[2t]   = sum_r g[F-1-2r] w1[t-r] + h[2r+1] w2[t+r]
//   v[2t+1] = sum_r g[F-2-2r] w1[t-r] + h[2r]   w2[t+r]          (indices mod p/2, r = 0..R-1 in the reference's order)
// =====================================================================================================
template <typename T, int F, int KM = 4>
struct IwptCfg {
    static constexpr int V = WxVec<T>::N;
    // output pairs per thread.  KM = 4 (reads at a 4-chunk stride) is the default; the launcher takes KM = 2 for nodes too short to
    // give a CTA 64 units of 4V pairs (Float32 n = 1024: iwptall 0.78 -> 0.57 ms, profiles/r2_iwpt_km_ab.jsonl)
    static constexpr int K = KM * V;
    static constexpr int R = F / 2;
    static constexpr int S = ((R - 1) + V - 1) / V * V;                // w1 look-behind (multiple of V)
    static constexpr int W = K + S;                                    // window length of each child
};

// one output pair from the two windows: a[s + k - r] = w1[t-r], b[k + r] = w2[t+r]
template <typename T, int F>
__device__ __forceinline__ void iwpt_pair(const T *a, const T *b, const Taps<T> &tp, T &ev, T &od)
{
    constexpr int R = F / 2;
    T e = tp.g[F - 1] * a[0];
    T o = tp.g[F - 2] * a[0];
    e = fma(tp.h[1], b[0], e);
    o = fma(tp.h[0], b[0], o);
#pragma unroll
    for (int r = 1; r < R; ++r) {
        e = fma(tp.g[F - 1 - 2 * r], a[-r], e);
        o = fma(tp.g[F - 2 - 2 * r], a[-r], o);
        e = fma(tp.h[2 * r + 1], b[r], e);
        o = fma(tp.h[2 * r], b[r], o);
    }
    ev = e;
    od = o;
}

template <typename T, int F, bool POW2, bool TREE, int KM = 4>
__device__ __forceinline__ void iwpt_wide_level(const T *__restrict__ src, T *__restrict__ dst, int n0, int p, const Taps<T> &tp, int tid,
                                                int nthreads, TreeMask tmk)
{
    using C = IwptCfg<T, F, KM>;
    using VT = typename WxVec<T>::type;
    constexpr int V = C::V, K = C::K, S = C::S, W = C::W;
    constexpr int GA = (S + K - 1) / K;                 // aligned groups of K elements behind t0 touched by the w1 window
    constexpr int NB = (W + K - 1) / K;                 // groups touched by the w2 window
    const int half = p >> 1;
    const int units = n0 / (2 * K);
    const int lgh = 31 - __clz(half);
    for (int u = tid; u < units; u += nthreads) {
        const int gi = u * K;
        const int j = POW2 ? (gi >> lgh) : (gi / half);
        const int t0 = gi - j * half;
        const int base = j * p;
        const int go = wx_swz_group<T>(base + 2 * t0);  // the 2K outputs fill one aligned run of 8 chunks
        if (TREE && !tmk.on(j)) {
#pragma unroll
            for (int c = 0; c < 2 * K / V; ++c) *reinterpret_cast<VT *>(dst + (go ^ (c * V))) = *reinterpret_cast<const VT *>(src + (go ^ (c * V)));
            continue;
        }
        int ga[GA + 1], gb[NB];
#pragma unroll
        for (int g = 0; g <= GA; ++g) {
            int o = t0 - (GA - g) * K;
            if (POW2) o &= half - 1; else { o %= half; if (o < 0) o += half; }
            ga[g] = wx_swz_group<T>(base + o);
        }
#pragma unroll
        for (int g = 0; g < NB; ++g) {
            int o = t0 + g * K;
            if (POW2) o &= half - 1; else o %= half;
            gb[g] = wx_swz_group<T>(base + half + o);
        }
        T a[W], b[W];
#pragma unroll
        for (int c = 0; c < W / V; ++c) {
            const int pa = GA * K - S + c * V;           // element position of this chunk relative to the first w1 group
            const int pb = c * V;
            wx_unpack(&a[c * V], *reinterpret_cast<const VT *>(src + (ga[pa / K] ^ (pa % K))));
            wx_unpack(&b[c * V], *reinterpret_cast<const VT *>(src + (gb[pb / K] ^ (pb % K))));
        }
        T out[2 * K];
#pragma unroll
        for (int k = 0; k < K; ++k) iwpt_pair<T, F>(&a[S + k], &b[k], tp, out[2 * k], out[2 * k + 1]);
#pragma unroll
        for (int c = 0; c < 2 * K / V; ++c) *reinterpret_cast<VT *>(dst + (go ^ (c * V))) = wx_pack(&out[c * V]);
    }
}

// node length P in {2,4,8}: a thread owns max(P,V) consecutive elements = whole nodes, wraps resolved at compile time
template <typename T, int F, int P, bool TREE>
__device__ __forceinline__ void iwpt_small_level(const T *__restrict__ src, T *__restrict__ dst, int n0, const Taps<T> &tp, int tid, int nthreads,
                                                 TreeMask tmk)
{
    using VT = typename WxVec<T>::type;
    constexpr int V = WxVec<T>::N;
    constexpr int G = P > V ? P : V;
    constexpr int H = P / 2, R = F / 2;
    const int groups = n0 / G;
    for (int u = tid; u < groups; u += nthreads) {
        const int e0 = u * G;
        T v[G], o[G];
        const int g0 = wx_swz_chunk(e0 / V) * V;         // G/V <= 4 chunks, aligned: chunk c sits at g0 ^ (c * V)
#pragma unroll
        for (int c = 0; c < G / V; ++c) wx_unpack(&v[c * V], *reinterpret_cast<const VT *>(src + (g0 ^ (c * V))));
#pragma unroll
        for (int nd = 0; nd < G / P; ++nd) {
            if (TREE && !tmk.on(e0 / P + nd)) {
#pragma unroll
                for (int i = 0; i < P; ++i) o[nd * P + i] = v[nd * P + i];
                continue;
            }
#pragma unroll
            for (int t = 0; t < H; ++t) {
                const T *w1 = &v[nd * P], *w2 = &v[nd * P + H];
                T e = tp.g[F - 1] * w1[t];
                T od = tp.g[F - 2] * w1[t];
                e = fma(tp.h[1], w2[t], e);
                od = fma(tp.h[0], w2[t], od);
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    e = fma(tp.g[F - 1 - 2 * r], w1[(t - r) & (H - 1)], e);
                    od = fma(tp.g[F - 2 - 2 * r], w1[(t - r) & (H - 1)], od);
                    e = fma(tp.h[2 * r + 1], w2[(t + r) & (H - 1)], e);
                    od = fma(tp.h[2 * r], w2[(t + r) & (H - 1)], od);
                }
                o[nd * P + 2 * t] = e;
                o[nd * P + 2 * t + 1] = od;
            }
        }
#pragma unroll
        for (int c = 0; c < G / V; ++c) *reinterpret_cast<VT *>(dst + (g0 ^ (c * V))) = wx_pack(&o[c * V]);
    }
}

// one inverse level of the nodes of length P inside a register-resident run of G elements (compile-time wraps)
template <typename T, int F, int P, int G>
__device__ __forceinline__ void iwpt_regs_level(const T *v, T *o, const Taps<T> &tp)
{
    constexpr int H = P / 2, R = F / 2;
#pragma unroll
    for (int nd = 0; nd < G / P; ++nd) {
        const T *w1 = &v[nd * P], *w2 = &v[nd * P + H];
#pragma unroll
        for (int t = 0; t < H; ++t) {
            T e = tp.g[F - 1] * w1[t];
            T od = tp.g[F - 2] * w1[t];
            e = fma(tp.h[1], w2[t], e);
            od = fma(tp.h[0], w2[t], od);
#pragma unroll
            for (int r = 1; r < R; ++r) {
                e = fma(tp.g[F - 1 - 2 * r], w1[(t - r) & (H - 1)], e);
                od = fma(tp.g[F - 2 - 2 * r], w1[(t - r) & (H - 1)], od);
                e = fma(tp.h[2 * r + 1], w2[(t + r) & (H - 1)], e);
                od = fma(tp.h[2 * r], w2[(t + r) & (H - 1)], od);
            }
            o[nd * P + 2 * t] = e;
            o[nd * P + 2 * t + 1] = od;
        }
    }
}

// the four deepest levels of a complete tree in one pass: a thread owns 16 consecutive elements (= one node of length 16 and
// everything below it), rebuilds the nodes of length 2, 4, 8 and 16 in registers and writes the run once -- three
// shared-memory round trips and three barriers fewer than level by level.  n0 must be a multiple of 16.
template <typename T, int F>
__device__ __forceinline__ void iwpt_small_levels4(const T *__restrict__ src, T *__restrict__ dst, int n0, const Taps<T> &tp, int tid, int nthreads)
{
    using VT = typename WxVec<T>::type;
    constexpr int V = WxVec<T>::N, G = 16;
    for (int u = tid; u < n0 / G; u += nthreads) {
        const int g0 = wx_swz_chunk((u * G) / V) * V;     // 16 elements = an aligned run of 8 (F64) / 4 (F32) chunks
        T v[G], o[G];
#pragma unroll
        for (int c = 0; c < G / V; ++c) wx_unpack(&v[c * V], *reinterpret_cast<const VT *>(src + (g0 ^ (c * V))));
        iwpt_regs_level<T, F, 2, G>(v, o, tp);
        iwpt_regs_level<T, F, 4, G>(o, v, tp);
        iwpt_regs_level<T, F, 8, G>(v, o, tp);
        iwpt_regs_level<T, F, 16, G>(o, v, tp);
#pragma unroll
        for (int c = 0; c < G / V; ++c) *reinterpret_cast<VT *>(dst + (g0 ^ (c * V))) = wx_pack(&v[c * V]);
    }
}

// the same along a tree: a node the tree does not split passes through (its children occupy exactly its range).  tm[q] = split flags of
// the depth whose nodes have 2 << q ... i.e. q = 0: nodes of length 2, 1: 4, 2: 8, 3: 16; node indices are relative to the staged node.
template <typename T, int F, int P, int G>
__device__ __forceinline__ void iwpt_regs_level_tree(const T *v, T *o, const Taps<T> &tp, const TreeMask &tm, long node0)
{
    constexpr int H = P / 2;
#pragma unroll
    for (int nd = 0; nd < G / P; ++nd) {
        if (tm.on(node0 + nd)) {
            const T *w1 = &v[nd * P], *w2 = &v[nd * P + H];
#pragma unroll
            for (int t = 0; t < H; ++t) {
                constexpr int R = F / 2;
                T e = tp.g[F - 1] * w1[t];
                T od = tp.g[F - 2] * w1[t];
                e = fma(tp.h[1], w2[t], e);
                od = fma(tp.h[0], w2[t], od);
#pragma unroll
                for (int r = 1; r < R; ++r) {
                    e = fma(tp.g[F - 1 - 2 * r], w1[(t - r) & (H - 1)], e);
                    od = fma(tp.g[F - 2 - 2 * r], w1[(t - r) & (H - 1)], od);
                    e = fma(tp.h[2 * r + 1], w2[(t + r) & (H - 1)], e);
                    od = fma(tp.h[2 * r], w2[(t + r) & (H - 1)], od);
                }
                o[nd * P + 2 * t] = e;
                o[nd * P + 2 * t + 1] = od;
            }
        } else {
#pragma unroll
            for (int i = 0; i < P; ++i) o[nd * P + i] = v[nd * P + i];
        }
    }
}
template <typename T, int F>
__device__ __forceinline__ void iwpt_small_levels4_tree(const T *__restrict__ src, T *__restrict__ dst, int n0, const Taps<T> &tp, int tid, int nthreads,
                                                        TreeMask tm2, TreeMask tm4, TreeMask tm8, TreeMask tm16)
{
    using VT = typename WxVec<T>::type;
    constexpr int V = WxVec<T>::N, G = 16;
    for (int u = tid; u < n0 / G; u += nthreads) {
        const int g0 = wx_swz_chunk((u * G) / V) * V;
        T v[G], o[G];
#pragma unroll
        for (int c = 0; c < G / V; ++c) wx_unpack(&v[c * V], *reinterpret_cast<const VT *>(src + (g0 ^ (c * V))));
        iwpt_regs_level_tree<T, F, 2, G>(v, o, tp, tm2, (long)u * 8);
        iwpt_regs_level_tree<T, F, 4, G>(o, v, tp, tm4, (long)u * 4);
        iwpt_regs_level_tree<T, F, 8, G>(v, o, tp, tm8, (long)u * 2);
        iwpt_regs_level_tree<T, F, 16, G>(o, v, tp, tm16, (long)u);
#pragma unroll
        for (int c = 0; c < G / V; ++c) *reinterpret_cast<VT *>(dst + (g0 ^ (c * V))) = wx_pack(&v[c * V]);
    }
}

// any even node length, one output pair per thread
template <typename T, int F, bool TREE>
__device__ __forceinline__ void iwpt_generic_level(const T *__restrict__ src, T *__restrict__ dst, int n0, int p, const Taps<T> &tp, int tid,
                                                   int nthreads, TreeMask tmk)
{
    constexpr int R = F / 2;
    const int half = p >> 1;
    for (int gi = tid; gi < n0 / 2; gi += nthreads) {
        const int j = gi / half;
        const int t = gi - j * half;
        const int base = j * p;
        if (TREE && !tmk.on(j)) {
            dst[wx_swz_elem<T>(base + 2 * t)] = src[wx_swz_elem<T>(base + 2 * t)];
            dst[wx_swz_elem<T>(base + 2 * t + 1)] = src[wx_swz_elem<T>(base + 2 * t + 1)];
            continue;
        }
        int k1 = t, k2 = t;
        T x1 = src[wx_swz_elem<T>(base + k1)], x2 = src[wx_swz_elem<T>(base + half + k2)];
        T e = tp.g[F - 1] * x1;
        T od = tp.g[F - 2] * x1;
        e = fma(tp.h[1], x2, e);
        od = fma(tp.h[0], x2, od);
#pragma unroll 4
        for (int r = 1; r < R; ++r) {
            k1 -= 1; if (k1 < 0) k1 += half;
            k2 += 1; if (k2 >= half) k2 -= half;
            x1 = src[wx_swz_elem<T>(base + k1)];
            x2 = src[wx_swz_elem<T>(base + half + k2)];
            e = fma(tp.g[F - 1 - 2 * r], x1, e);
            od = fma(tp.g[F - 2 - 2 * r], x1, od);
            e = fma(tp.h[2 * r + 1], x2, e);
            od = fma(tp.h[2 * r], x2, od);
        }
        dst[wx_swz_elem<T>(base + 2 * t)] = e;
        dst[wx_swz_elem<T>(base + 2 * t + 1)] = od;
    }
}

template <typename T, int F, bool TREE, int KM = 4>
__device__ __forceinline__ void iwpt_level(const T *__restrict__ a, T *__restrict__ b, int n0, int p, const Taps<T> &tp, int tid, int nthreads,
                                           TreeMask tmk)
{
    using C = IwptCfg<T, F, KM>;
    constexpr int V = C::V, K = C::K;
    const int half = p >> 1;
    const bool pow2 = (p & (p - 1)) == 0;
    if (half % K == 0) {
        if (pow2) iwpt_wide_level<T, F, true, TREE, KM>(a, b, n0, p, tp, tid, nthreads, tmk);
        else      iwpt_wide_level<T, F, false, TREE, KM>(a, b, n0, p, tp, tid, nthreads, tmk);
    } else if (p == 2 && n0 % (V > 2 ? V : 2) == 0) {
        iwpt_small_level<T, F, 2, TREE>(a, b, n0, tp, tid, nthreads, tmk);
    } else if (p == 4) {
        iwpt_small_level<T, F, 4, TREE>(a, b, n0, tp, tid, nthreads, tmk);
    } else if (p == 8) {
        iwpt_small_level<T, F, 8, TREE>(a, b, n0, tp, tid, nthreads, tmk);
    } else if (p == 16 && V == 4) {
        iwpt_small_level<T, F, 16, TREE>(a, b, n0, tp, tid, nthreads, tmk);
    } else {
        iwpt_generic_level<T, F, TREE>(a, b, n0, p, tp, tid, nthreads, tmk);
    }
}

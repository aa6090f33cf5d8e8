// wx_2d.cuh -- helpers shared by the fused 2-D kernels (wx_wpd2d.cu forward, wx_iwpt2d.cu inverse)
#pragma once
#include "wx_steps.cuh"

constexpr int kT2 = 256;

template <typename T> struct Pair;
template <> struct Pair<double> { using type = double2; };
template <> struct Pair<float> { using type = float2; };

// w[0..F-1] = v[2i .. 2i+F-1]  ->  lo = w1[i],  hi = w2[i + (F-2)/2]        dwt/dwt_one_level.jl:97-104
template <typename T, int F>
__device__ __forceinline__ void dwt_dots(const T *w, const Taps<T> &tp, T &lo, T &hi)
{
    T a = tp.g[F - 1] * w[0];
    T b = tp.h[0] * w[F - 1];
#pragma unroll
    for (int j = 1; j < F; ++j) {
        a = fma(tp.g[F - 1 - j], w[j], a);
        b = fma(tp.h[j], w[F - 1 - j], b);
    }
    lo = a;
    hi = b;
}

// asynchronous global -> shared copy of one pair of elements (LDGSTS): many copies in flight per thread, no register staging
template <typename T>
__device__ __forceinline__ void cp_async_pair(T *smem_dst, const T *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (sizeof(T) == 8) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
    else                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
}
// one element per asynchronous copy (8 / 4 bytes): for patches whose rows are not pair aligned
template <typename T>
__device__ __forceinline__ void cp_async_elem(T *smem_dst, const T *gsrc)
{
    const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
    if (sizeof(T) == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(sa), "l"(gsrc) : "memory");
    else                asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory"); }

// division of a 32-bit index by a launch constant (host-built): shift for powers of two, else round-up multiply-high
struct Div32 {
    unsigned d, m, s;
    int p2;
};
static inline Div32 make_div32(long dd)
{
    Div32 r;
    r.d = (unsigned)dd; r.p2 = (dd & (dd - 1)) == 0; r.m = 0; r.s = 0;
    if (r.p2) { while ((1L << r.s) < dd) ++r.s; return r; }
    unsigned l = 0;
    while ((1UL << l) < (unsigned long)dd) ++l;
    r.m = (unsigned)((((1UL << l) - (unsigned long)dd) << 32) / (unsigned long)dd + 1);
    r.s = l - 1;
    return r;
}
__device__ __forceinline__ unsigned div32(unsigned x, const Div32 &v)
{
    if (v.p2) return x >> v.s;
    const unsigned t = __umulhi(v.m, x);
    return (t + ((x - t) >> 1)) >> v.s;
}

struct FastDiv {                    // division by a runtime constant that is usually a power of two
    int d, lg;
    __host__ __device__ FastDiv() : d(1), lg(0) {}
    __host__ __device__ explicit FastDiv(int dd) : d(dd), lg(-1)
    {
        if (dd > 0 && (dd & (dd - 1)) == 0) { lg = 0; while ((1 << lg) < dd) ++lg; }
    }
    __device__ __forceinline__ int div(int x) const { return lg >= 0 ? (x >> lg) : (x / d); }
    __device__ __forceinline__ int mod(int x) const { return lg >= 0 ? (x & (d - 1)) : (x % d); }
};

// walks the index pairs (hi, lo), lo < len, of idx = tid, tid + kT2, ... without a division per step
struct Walk2 {
    int lo, hi, qlo, qhi, len;
    __device__ __forceinline__ Walk2(int tid, int l) : len(l)
    {
        hi = tid / l; lo = tid - hi * l;
        qhi = kT2 / l; qlo = kT2 - qhi * l;
    }
    __device__ __forceinline__ void next()
    {
        lo += qlo; hi += qhi;
        if (lo >= len) { lo -= len; ++hi; }
    }
};

// Two adjacent output pairs from one window: w[0..F+1] = v[2i .. 2i+F+1]
template <typename T, int F>
__device__ __forceinline__ void dwt_dots2(const T *w, const Taps<T> &tp, T &lo0, T &hi0, T &lo1, T &hi1)
{
    T a0 = tp.g[F - 1] * w[0], a1 = tp.g[F - 1] * w[2];
    T b0 = tp.h[0] * w[F - 1], b1 = tp.h[0] * w[F + 1];
#pragma unroll
    for (int j = 1; j < F; ++j) {
        a0 = fma(tp.g[F - 1 - j], w[j], a0);
        a1 = fma(tp.g[F - 1 - j], w[j + 2], a1);
        b0 = fma(tp.h[j], w[F - 1 - j], b0);
        b1 = fma(tp.h[j], w[F + 1 - j], b1);
    }
    lo0 = a0; lo1 = a1; hi0 = b0; hi1 = b1;
}

// heap index of quad node (depth d, block row jr, block col jc): children 4i-2 (TL) 4i-1 (TR) 4i (BL) 4i+1 (BR)
__device__ __forceinline__ long quad_index2(int d, int jr, int jc)
{
    long idx = 1;
    for (int b = d - 1; b >= 0; --b) idx = 4 * idx - 2 + 2 * ((jr >> b) & 1) + ((jc >> b) & 1);
    return idx;
}
__device__ __forceinline__ bool split2(const unsigned char *tree, long ntree, int d, int jr, int jc)
{
    if (tree == nullptr) return true;
    const long i = quad_index2(d, jr, jc);
    return i <= ntree && tree[i - 1];
}

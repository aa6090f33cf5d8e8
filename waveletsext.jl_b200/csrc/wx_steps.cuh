// wx_steps.cuh -- generic (any n, any F, strided, batched) one-level step launchers.
// These are the general-purpose kernels behind the single-step entry points, the 2-D drivers and
// every shape the fused kernels do not cover.  One thread computes one output (pair).
#pragma once
#include "wx_common.cuh"

// A batch of strided 1-D vectors: element i of problem (b0,b1,b2) lives at
//   p + i*es + b0*s0 + b1*s1 + b2*s2
template <typename T>
struct View {
    T *p;
    long es, s0, s1, s2;
};
struct Batch {
    long B0, B1, B2;
    bool batch_fast;    // true: consecutive threads walk b0 (use when s0 == 1), else they walk i
};

template <typename T> static inline View<T> view1(T *p) { return View<T>{p, 1, 0, 0, 0}; }
static inline Batch batch1() { return Batch{1, 1, 1, false}; }

// forward steps ---------------------------------------------------------------------------------
// dwt_step!  dwt/dwt_one_level.jl:79-107 ; v length n -> w1,w2 length n/2
template <typename T> int wx_launch_dwt_step(View<T> w1, View<T> w2, View<const T> v, long n, Batch b, const Taps<T> &t, cudaStream_t s, int shift = 0);
// idwt_step! dwt/dwt_one_level.jl:192-223
template <typename T> int wx_launch_idwt_step(View<T> v, View<const T> w1, View<const T> w2, long n, Batch b, const Taps<T> &t, cudaStream_t s, int shift = 0);
// sdwt_step! swt/swt_one_level.jl:99-127 (ac = 0) / acdwt_step! acwt/acwt_one_level.jl:101-128 (ac = 1)
template <typename T> int wx_launch_rdwt_step(int ac, View<T> w1, View<T> w2, View<const T> v, long n, int d, Batch b, const Taps<T> &t, cudaStream_t s);
// isdwt_step! shift based swt/swt_one_level.jl:279-318 (writes only the sv coset of v)
template <typename T> int wx_launch_isdwt_shift(View<T> v, View<const T> w1, View<const T> w2, long n, int d, long sv, long sw, int add2out, Batch b, const Taps<T> &t, cudaStream_t s);
// isdwt_step! average based swt/swt_one_level.jl:257-277
template <typename T> int wx_launch_isdwt_avg(View<T> v, View<const T> w1, View<const T> w2, long n, int d, Batch b, const Taps<T> &t, cudaStream_t s);
// iacdwt_step! acwt/acwt_one_level.jl:217-224
template <typename T> int wx_launch_iacdwt_step(View<T> v, View<const T> w1, View<const T> w2, long n, Batch b, cudaStream_t s);
// strided batched copy dst <- src (n elements per problem)
template <typename T> int wx_launch_copy(View<T> dst, View<const T> src, long n, Batch b, cudaStream_t s);

// stream-ordered scratch
template <typename T>
static inline int wx_scratch(T **p, size_t elems, cudaStream_t s)
{
    return wx_pool_alloc((void **)p, (elems ? elems : 1) * sizeof(T), s);
}
static inline int wx_scratch_free(void *p, cudaStream_t s)
{
    if (p) WX_CUDA(cudaFreeAsync(p, s));
    return WX_OK;
}

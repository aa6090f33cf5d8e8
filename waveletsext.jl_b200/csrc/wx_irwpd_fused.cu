// wx_irwpd_fused.cu -- fused average-based inverse stationary transforms: iswpt!, iswpd! on a complete tree and isdwt!
// without a shift (SWT.jl:311-328, 700-715, 1143-1156) over the average-based isdwt_step! (swt/swt_one_level.jl:257-277).
//
// The average-based step runs the shift-based reconstruction for both child cosets of every parent coset and halves the
// sum.  Written out that is half the adjoint of the a-trous analysis step:
//     v[p] = 1/2 * ( sum_j g[j] * w1[p + (j+2-F) D]  +  sum_j h[j] * w2[p + j D] ),   D = 2^d, indices mod n
// so the same coset register-window scheme as the forward kernel applies (a thread owns K consecutive elements of one coset
// and reads one window of K+F-1 coset samples from each child).
//
//  * irwpd_tree_k: a CTA reduces a subtree of E levels in shared memory in post order (two child buffers per level): the
//    2^E input columns are read once with bulk async copies, the subtree root is written once.  Trees deeper than E run
//    as a chain of launches over compacted workspaces: the table is read once, every later stage moves 2^-E of the data.
//  * irdwt_chain_k: the isdwt! chain x <- step(x, detail of depth d), all levels per signal in shared memory.
#include "wx_steps.cuh"
#include "wx_tma.cuh"
#include <cstdlib>

namespace {

template <typename T, int F>
struct IrCfg {
#ifndef WX_IR_LGK
#define WX_IR_LGK 4
#endif
    static constexpr int LGK = (F <= 16) ? WX_IR_LGK : 3;
    static constexpr int K = 1 << LGK;
};

// parent (depth d) from its two children, all of length n (power of two), linear shared-memory buffers
template <typename T, int F>
__device__ __forceinline__ void ir_combine(const T *__restrict__ w1, const T *__restrict__ w2, T *__restrict__ dst, int n, int d, const Taps<T> &tp,
                                           int tid, int nthr)
{
    using C = IrCfg<T, F>;
    constexpr int K = C::K, W = K + F - 1;
    const int D = 1 << d, mask = n - 1;
    for (int u = tid; u < (n >> C::LGK); u += nthr) {
        const int base = ((u >> d) << (d + C::LGK)) | (u & (D - 1));
        T a[W], b[W];
        int i1 = (base + (2 - F) * D) & mask, i2 = base;
#pragma unroll
        for (int m = 0; m < W; ++m) {
            a[m] = w1[i1]; b[m] = w2[i2];
            i1 = (i1 + D) & mask; i2 = (i2 + D) & mask;
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            T s = tp.g[0] * a[k];
#pragma unroll
            for (int j = 1; j < F; ++j) s = fma(tp.g[j], a[k + j], s);
#pragma unroll
            for (int j = 0; j < F; ++j) s = fma(tp.h[j], b[k + j], s);
            dst[base + k * D] = s * (T)0.5;
        }
    }
}

// item = (signal k, node j0 of depth dr): reduce its 2^E descendants of depth dr+E (columns in_col0 + index of `in`) to the
// node itself (column out_col0 + j0 of `out`).  Buffers: R (result) + two per level 1..E.
template <typename T, int F>
__global__ void __launch_bounds__(256) irwpd_tree_k(T *__restrict__ out, long out_sig, long out_col0, const T *__restrict__ in, long in_sig,
                                                   long in_col0, int n, int dr, int E, long items, Taps<T> tp)
{
    extern __shared__ __align__(128) unsigned char wx_ir_smem[];
    __shared__ __align__(8) unsigned long long bar;
    T *R = reinterpret_cast<T *>(wx_ir_smem);
    auto B = [&](int e, int side) { return R + (size_t)(1 + 2 * (e - 1) + side) * n; };
    const unsigned nbytes = (unsigned)n * (unsigned)sizeof(T);
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) {
        wx_mbar_init(&bar, 1);
        wx_fence_mbar_init();
    }
    __syncthreads();
    unsigned parity = 0;
    // leaf pair c of an item -> the two landing buffers B(E, .) (one thread)
    auto load_pair = [&](long item, int c) {
        const long k = item >> dr, j0 = item & ((1L << dr) - 1);
        const T *src = in + k * in_sig + (in_col0 + (j0 << E)) * n;
        wx_mbar_expect_tx(&bar, 2 * nbytes);
        wx_bulk_load_1d(B(E, 0), src + (long)c * n, nbytes, &bar);
        wx_bulk_load_1d(B(E, 1), src + (long)(c + 1) * n, nbytes, &bar);
    };
    if (tid == 0 && (long)blockIdx.x < items) load_pair(blockIdx.x, 0);
    for (long item = blockIdx.x; item < items; item += gridDim.x) {
        const long k = item >> dr, j0 = item & ((1L << dr) - 1);
        for (int c = 0; c < (1 << E); c += 2) {
            wx_mbar_wait(&bar, parity);
            parity ^= 1;
            int e = E, idx = c >> 1;
            while (true) {
                T *dst = (e == 1) ? R : B(e - 1, idx & 1);
                ir_combine<T, F>(B(e, 0), B(e, 1), dst, n, dr + e - 1, tp, tid, nthr);
                __syncthreads();
                if (e == E && tid == 0) {
                    // the landing buffers are free again: the next pair (of this item or the next one) streams in under the
                    // remaining combines of this pair and the root store (ncu before: 35 % of the samples on the mbarrier spin)
                    if (c + 2 < (1 << E)) load_pair(item, c + 2);
                    else if (item + gridDim.x < items) load_pair(item + gridDim.x, 0);
                }
                --e;
                if (e == 0 || (idx & 1) == 0) break;
                idx >>= 1;
            }
        }
        // the subtree root
        wx_fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            wx_bulk_store_1d(out + k * out_sig + (out_col0 + j0) * n, R, nbytes);
            wx_bulk_commit();
            wx_bulk_wait_read0();                         // R is rewritten only after 2^E - 1 combines, but keep it simple and safe
        }
        __syncthreads();
    }
    if (tid == 0) wx_bulk_wait_all();
}

// isdwt! average based (SWT.jl:311-328): x = col 0; for d = L-1..0: x = step_d(w1 = x, w2 = col L-d).  Levels [dlo, dhi) of the
// chain run here, top index first: the scaling input is column `c_in` of xin (signal stride in_sig), details come from xw.
template <typename T, int F>
__global__ void __launch_bounds__(256) irdwt_chain_k(T *__restrict__ x, const T *__restrict__ xw, int n, int L, int dhi, long N, Taps<T> tp)
{
    extern __shared__ __align__(128) unsigned char wx_ir_smem[];
    __shared__ __align__(8) unsigned long long bar;
    T *sc[2] = {reinterpret_cast<T *>(wx_ir_smem), reinterpret_cast<T *>(wx_ir_smem) + n};
    T *dt = sc[1] + n;
    const unsigned nbytes = (unsigned)n * (unsigned)sizeof(T);
    const int tid = threadIdx.x, nthr = blockDim.x;
    if (tid == 0) {
        wx_mbar_init(&bar, 1);
        wx_fence_mbar_init();
    }
    __syncthreads();
    unsigned parity = 0;
    for (long k = blockIdx.x; k < N; k += gridDim.x) {
        const T *xk = xw + k * (long)(L + 1) * n;
        int cur = 0;
        for (int d = dhi - 1; d >= 0; --d) {
            if (tid == 0) {
                const bool first = (d == dhi - 1);
                wx_mbar_expect_tx(&bar, first ? 2 * nbytes : nbytes);
                // the scaling node of depth dhi sits in x when the deeper levels were done by the per-depth path, else in column 0
                if (first) wx_bulk_load_1d(sc[0], dhi == L ? xk : x + k * n, nbytes, &bar);
                wx_bulk_load_1d(dt, xk + (long)(L - d) * n, nbytes, &bar);
            }
            wx_mbar_wait(&bar, parity);
            parity ^= 1;
            ir_combine<T, F>(sc[cur], dt, sc[cur ^ 1], n, d, tp, tid, nthr);
            __syncthreads();
            cur ^= 1;
        }
        wx_fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            wx_bulk_store_1d(x + k * n, sc[cur], nbytes);
            wx_bulk_commit();
            wx_bulk_wait_read0();
        }
        __syncthreads();
    }
    if (tid == 0) wx_bulk_wait_all();
}

template <typename T, int F>
int tree_launch(T *out, long out_sig, long out_col0, const T *in, long in_sig, long in_col0, long n, int dr, int E, long N, const Taps<T> &t,
                cudaStream_t s)
{
    using C = IrCfg<T, F>;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const size_t smem = (size_t)(2 * E + 1) * n * sizeof(T);
    int threads = (int)(((n >> C::LGK) + 31) / 32 * 32);
    if (threads > 256) threads = 256;
    auto kern = irwpd_tree_k<T, F>;
    WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    WX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) return wx_fail(WX_EUNSUPPORTED, "irwpd fused kernel does not fit (smem %zu)", smem);
    const long items = N << dr;
    long blocks = (long)dv.sms * occ;
    if (blocks > items) blocks = items;
    kern<<<(unsigned)blocks, threads, smem, s>>>(out, out_sig, out_col0, in, in_sig, in_col0, (int)n, dr, E, items, t);
    WX_LAUNCHED();
    return WX_OK;
}

// columns in_col0 .. in_col0 + 2^Lt - 1 of xw (signal stride in_sig) are the depth-Lt nodes; x(n, N) receives the root
template <typename T, int F>
int tree_plan(T *x, const T *xw, long in_sig, long in_col0, long n, int Lt, long N, const Taps<T> &t, cudaStream_t s, bool *handled)
{
    using C = IrCfg<T, F>;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const int lgn = wx_ilog2l(n);
    if (lgn < C::LGK || Lt - 1 > lgn - C::LGK) return WX_OK;        // deepest parent needs K * 2^d <= n
    const size_t buf = (size_t)n * sizeof(T);
    long ehalf = ((long)((dv.smem_optin / 2 - 1024) / buf) - 1) / 2, efull = ((long)(dv.smem_optin / buf) - 1) / 2;
    long E = ehalf >= 2 ? ehalf : efull;
    if (E < 1) return WX_OK;
    if (E > 6) E = 6;
    // stages bottom-up; intermediate levels live in two compact workspaces (n, 2^d, N)
    T *wsp[2] = {nullptr, nullptr};
    int d = Lt, which = 0;
    const T *in = xw; long isig = in_sig, icol0 = in_col0;
    while (d > 0 && !rc) {
        const int e = (int)(d < E ? d : E), dr = d - e;
        T *out; long osig;
        if (dr == 0) { out = x; osig = n; }
        else {
            if (!wsp[which]) { rc = wx_scratch(&wsp[which], (size_t)n * (1L << dr) * N, s); if (rc) break; }
            out = wsp[which]; osig = n * (1L << dr);
        }
        rc = tree_launch<T, F>(out, osig, 0, in, isig, icol0, n, dr, e, N, t, s);
        in = out; isig = osig; icol0 = 0; d = dr; which ^= 1;
    }
    int rc2 = wx_scratch_free(wsp[0], s), rc3 = wx_scratch_free(wsp[1], s);
    if (rc) return rc;
    if (rc2 || rc3) return rc2 ? rc2 : rc3;
    *handled = true;
    return WX_OK;
}

template <typename T, int F>
int chain_plan(T *x, const T *xw, long n, int L, long N, const Taps<T> &t, cudaStream_t s, int *dhi)
{
    using C = IrCfg<T, F>;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const int lgn = wx_ilog2l(n);
    if (lgn < C::LGK) return WX_OK;
    int top = lgn - C::LGK + 1;                        // levels d < top can run fused
    if (top > L) top = L;
    const size_t smem = (size_t)3 * n * sizeof(T);
    if (smem > dv.smem_optin) return WX_OK;
    *dhi = top;
    return WX_OK;
}

template <typename T, int F>
int chain_launch(T *x, const T *xw, long n, int L, int dhi, long N, const Taps<T> &t, cudaStream_t s)
{
    using C = IrCfg<T, F>;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const size_t smem = (size_t)3 * n * sizeof(T);
    int threads = (int)(((n >> C::LGK) + 31) / 32 * 32);
    if (threads > 256) threads = 256;
    auto kern = irdwt_chain_k<T, F>;
    WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int occ = 0;
    WX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, threads, smem));
    if (occ < 1) return wx_fail(WX_EUNSUPPORTED, "irdwt chain kernel does not fit");
    long blocks = (long)dv.sms * occ;
    if (blocks > N) blocks = N;
    kern<<<(unsigned)blocks, threads, smem, s>>>(x, xw, (int)n, L, dhi, N, t);
    WX_LAUNCHED();
    return WX_OK;
}

static bool shape_ok(const void *a, const void *b, long n, size_t elt, long N, int L)
{
    static const bool off = getenv("WX_B200_NO_FUSED_RWPD") != nullptr;
    if (off || L < 1 || N < 1 || !wx_ispow2(n) || n >= (1L << 30) || L > 30) return false;
    return (n * elt) % 16 == 0 && ((((uintptr_t)a) | ((uintptr_t)b)) & 15) == 0;
}

}  // namespace

#define WX_IR_SWITCH(CALL)                                                                                      \
    switch (t.F) {                                                                                              \
        case 2: { constexpr int FF = 2; CALL; } break;   case 4: { constexpr int FF = 4; CALL; } break;         \
        case 6: { constexpr int FF = 6; CALL; } break;   case 8: { constexpr int FF = 8; CALL; } break;         \
        case 10: { constexpr int FF = 10; CALL; } break; case 12: { constexpr int FF = 12; CALL; } break;       \
        case 16: { constexpr int FF = 16; CALL; } break; case 20: { constexpr int FF = 20; CALL; } break;       \
        case 14: { constexpr int FF = 14; CALL; } break; case 18: { constexpr int FF = 18; CALL; } break;       \
        case 24: { constexpr int FF = 24; CALL; } break;                                                        \
        default: break;                                                                                         \
    }

// average-based inverse over a complete tree of depth Lt whose nodes are columns in_col0.. of xw (signal stride in_sig)
template <typename T>
int wx_irwpd_avg_fused(T *x, const T *xw, long in_sig, long in_col0, long n, int Lt, long N, const Taps<T> &t, cudaStream_t s, bool *handled)
{
    *handled = false;
    if (!shape_ok(x, xw, n, sizeof(T), N, Lt) || ((in_sig * sizeof(T)) % 16) != 0) return WX_OK;
    int rc = WX_OK;
    WX_IR_SWITCH(rc = (tree_plan<T, FF>(x, xw, in_sig, in_col0, n, Lt, N, t, s, handled)))
    return rc;
}

// isdwt! average based: levels d < *dhi can run fused (0 = not covered).  Call wx_irdwt_chain_run after the per-depth path has
// produced the scaling node of depth dhi in x (dhi == L: it is column 0 of xw).
template <typename T>
int wx_irdwt_chain_depth(const T *x, const T *xw, long n, int L, long N, const Taps<T> &t, int *dhi)
{
    *dhi = 0;
    if (!shape_ok(x, xw, n, sizeof(T), N, L)) return WX_OK;
    int rc = WX_OK;
    WX_IR_SWITCH(rc = (chain_plan<T, FF>((T *)nullptr, xw, n, L, N, t, (cudaStream_t)0, dhi)))
    return rc;
}
template <typename T>
int wx_irdwt_chain_run(T *x, const T *xw, long n, int L, int dhi, long N, const Taps<T> &t, cudaStream_t s)
{
    int rc = wx_fail(WX_EUNSUPPORTED, "irdwt chain: filter length %d", t.F);
    WX_IR_SWITCH(rc = (chain_launch<T, FF>(x, xw, n, L, dhi, N, t, s)))
    return rc;
}

#define WX_IR_INST(T)                                                                                                                      \
    template int wx_irwpd_avg_fused<T>(T *, const T *, long, long, long, int, long, const Taps<T> &, cudaStream_t, bool *);               \
    template int wx_irdwt_chain_depth<T>(const T *, const T *, long, int, long, const Taps<T> &, int *);                                  \
    template int wx_irdwt_chain_run<T>(T *, const T *, long, int, int, long, const Taps<T> &, cudaStream_t);
WX_IR_INST(double)
WX_IR_INST(float)

// wx_ldb.cu -- the feature side of the Local Discriminant Basis object (LDB.jl:186-330, ldb/ldb_measures.jl:427-479): gathering the
// n_features most discriminant coefficients of every signal in `order` (transform / fit_transform), scattering them back
// (inverse_transform) and the per-class mean / variance of every coefficient (FishersClassSeparability).  The energy maps,
// discriminant measures and node costs live in wx_bestbasis.cu.
#include "wx_steps.cuh"

namespace {

constexpr int kTL = 256;

// out(nf, N)[i, k] = X(nelem, N)[order[i], k]      transform LDB.jl:291-294
template <typename T>
__global__ void __launch_bounds__(kTL) select_features_k(T *__restrict__ out, const T *__restrict__ X, const int *__restrict__ order, long nf, long nelem, long total)
{
    const long idx = (long)blockIdx.x * kTL + threadIdx.x;
    if (idx >= total) return;
    const long k = idx / nf, i = idx - k * nf;
    out[idx] = X[k * nelem + order[i]];
}
// Xc(nelem, N)[order[i], k] = F(nf, N)[i, k] on a zero-filled Xc      inverse_transform LDB.jl:372-378
template <typename T>
__global__ void __launch_bounds__(kTL) scatter_features_k(T *__restrict__ Xc, const T *__restrict__ F, const int *__restrict__ order, long nf, long nelem, long total)
{
    const long idx = (long)blockIdx.x * kTL + threadIdx.x;
    if (idx >= total) return;
    const long k = idx / nf, i = idx - k * nf;
    Xc[k * nelem + order[i]] = F[idx];
}

// E, V (nc, nelem): mean and corrected variance over the signals of each class (two passes, like Statistics.var).
// grid (ceil(nelem / 32), nc), block (32, 8): lanes walk the coefficients, the 8 rows stride over the class's signals.
template <typename T>
__global__ void __launch_bounds__(256) class_moments_k(double *__restrict__ E, double *__restrict__ V, const T *__restrict__ X, const int *__restrict__ sig,
                                                      const int *__restrict__ off, long nelem)
{
    __shared__ double sh[8][33];
    const int tx = threadIdx.x, ty = threadIdx.y, c = blockIdx.y;
    const long f = (long)blockIdx.x * 32 + tx;
    const int s0 = off[c], s1 = off[c + 1], cnt = s1 - s0;
    double a = 0;
    if (f < nelem) for (int s = s0 + ty; s < s1; s += 8) a += (double)X[(long)sig[s] * nelem + f];
    sh[ty][tx] = a;
    __syncthreads();
    double mean = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) mean += sh[q][tx];
    mean /= (double)cnt;
    __syncthreads();
    double b = 0;
    if (f < nelem) for (int s = s0 + ty; s < s1; s += 8) { const double d = (double)X[(long)sig[s] * nelem + f] - mean; b = fma(d, d, b); }
    sh[ty][tx] = b;
    __syncthreads();
    if (ty == 0 && f < nelem) {
        double v = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) v += sh[q][tx];
        E[(long)c * nelem + f] = mean;
        V[(long)c * nelem + f] = v / (double)(cnt - 1);
    }
}

template <typename T>
int features_impl(bool scatter, T *dst, const T *src, const int *order, long nf, long nelem, long N, cudaStream_t s)
{
    WX_REQUIRE(dst && src && order && nf >= 1 && nf <= nelem && N >= 0, "bad arguments");
    if (N == 0) return WX_OK;
    const long total = nf * N;
    WX_REQUIRE((total + kTL - 1) / kTL < (1L << 31), "too many features for one launch");
    if (scatter) {
        WX_CUDA(cudaMemsetAsync(dst, 0, (size_t)nelem * N * sizeof(T), s));
        scatter_features_k<T><<<(unsigned)((total + kTL - 1) / kTL), kTL, 0, s>>>(dst, src, order, nf, nelem, total);
    } else {
        select_features_k<T><<<(unsigned)((total + kTL - 1) / kTL), kTL, 0, s>>>(dst, src, order, nf, nelem, total);
    }
    WX_LAUNCHED();
    return WX_OK;
}

template <typename T>
int moments_impl(double *E, double *V, const T *X, const int *sig, const int *off, int nc, long nelem, cudaStream_t s)
{
    WX_REQUIRE(E && V && X && sig && off && nc >= 1 && nelem >= 1, "bad arguments");
    WX_REQUIRE(nc <= 65535, "too many classes");
    class_moments_k<T><<<dim3((unsigned)((nelem + 31) / 32), (unsigned)nc), dim3(32, 8), 0, s>>>(E, V, X, sig, off, nelem);
    WX_LAUNCHED();
    return WX_OK;
}

}  // namespace

extern "C" {
int wx_select_features_f64(double *out, const double *X, const int *order_dev, long nf, long nelem, long N, void *s) { return features_impl<double>(false, out, X, order_dev, nf, nelem, N, (cudaStream_t)s); }
int wx_select_features_f32(float *out, const float *X, const int *order_dev, long nf, long nelem, long N, void *s) { return features_impl<float>(false, out, X, order_dev, nf, nelem, N, (cudaStream_t)s); }
int wx_scatter_features_f64(double *Xc, const double *F, const int *order_dev, long nf, long nelem, long N, void *s) { return features_impl<double>(true, Xc, F, order_dev, nf, nelem, N, (cudaStream_t)s); }
int wx_scatter_features_f32(float *Xc, const float *F, const int *order_dev, long nf, long nelem, long N, void *s) { return features_impl<float>(true, Xc, F, order_dev, nf, nelem, N, (cudaStream_t)s); }
int wx_class_moments_f64(double *E, double *V, const double *X, const int *sig_dev, const int *off_dev, int nc, long nelem, void *s) { return moments_impl<double>(E, V, X, sig_dev, off_dev, nc, nelem, (cudaStream_t)s); }
int wx_class_moments_f32(double *E, double *V, const float *X, const int *sig_dev, const int *off_dev, int nc, long nelem, void *s) { return moments_impl<float>(E, V, X, sig_dev, off_dev, nc, nelem, (cudaStream_t)s); }
}

// wx_host.cu -- host-buffer entry points: what a drop-in caller with ordinary (host) arrays invokes.
// wpdall: the batch is cut into chunks that cycle through three device slots, each with its own stream
// (H2D -> fused kernel -> D2H), so copies in both directions overlap the kernel of the neighbouring chunks.
// Reference: wpdall dwt/dwt_all.jl:260-282 (x and y are host arrays there).
#include "wx_common.cuh"

namespace {

constexpr int kSlots = 3;

template <typename T>
int wpdall_host(T *y, const T *x, long n, int L, long N, const double *h, const double *g, int F, long chunk,
                int (*kern)(T *, const T *, long, int, long, const double *, const double *, int, void *))
{
    WX_REQUIRE(n >= 1 && N >= 0, "wpdall: bad sizes n=%ld N=%ld", n, N);
    WX_REQUIRE(L >= 0 && L <= wx_maxlevels(n), "AssertionError: 0 <= L <= maxtransformlevels(x) (n=%ld, L=%d)", n, L);
    if (N == 0) return WX_OK;
    WX_REQUIRE(y && x, "null host pointer");
    const size_t in_b = (size_t)n * sizeof(T), out_b = in_b * (size_t)(L + 1);
    if (chunk <= 0) {
        chunk = (long)(((size_t)768 << 20) / (in_b + out_b));       // ~768 MiB of device memory per slot
        if (chunk < 1) chunk = 1;
    }
    if (chunk > N) chunk = N;
    { WxDev dv; int rc0 = wx_devinfo(dv); if (rc0) return rc0; }      // also keeps freed pool blocks cached across calls
    cudaStream_t st[kSlots] = {nullptr, nullptr, nullptr};
    T *dx[kSlots] = {nullptr, nullptr, nullptr}, *dy[kSlots] = {nullptr, nullptr, nullptr};
    int rc = WX_OK;
    auto cleanup = [&]() {
        for (int i = 0; i < kSlots; ++i) {
            if (st[i]) cudaStreamSynchronize(st[i]);
            if (dx[i]) cudaFreeAsync(dx[i], st[i]);         // back to the default pool (release threshold = max, wx_devinfo):
            if (dy[i]) cudaFreeAsync(dy[i], st[i]);         // the next call reuses the slots instead of paying cudaMalloc again
            if (st[i]) { cudaStreamSynchronize(st[i]); cudaStreamDestroy(st[i]); }
        }
    };
    const long nchunks = (N + chunk - 1) / chunk;
    const int slots = nchunks < kSlots ? (int)nchunks : kSlots;
    for (int i = 0; i < slots; ++i) {
        cudaError_t e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&dx[i], in_b * (size_t)chunk, st[i]);
        if (e == cudaSuccess) e = cudaMallocAsync((void **)&dy[i], out_b * (size_t)chunk, st[i]);
        if (e != cudaSuccess) {
            cleanup();
            cudaGetLastError();
            return wx_fail(e == cudaErrorMemoryAllocation ? WX_ENOMEM : WX_ECUDA, "wpdall_host setup: %s", cudaGetErrorString(e));
        }
    }
    for (long c = 0; c < nchunks && rc == WX_OK; ++c) {
        const int sl = (int)(c % slots);
        const long k0 = c * chunk, nk = (N - k0 < chunk) ? N - k0 : chunk;
        cudaError_t e = cudaMemcpyAsync(dx[sl], x + k0 * n, in_b * (size_t)nk, cudaMemcpyHostToDevice, st[sl]);
        if (e != cudaSuccess) { rc = wx_fail(WX_ECUDA, "H2D: %s", cudaGetErrorString(e)); break; }
        rc = kern(dy[sl], dx[sl], n, L, nk, h, g, F, (void *)st[sl]);
        if (rc) break;
        e = cudaMemcpyAsync(y + k0 * n * (L + 1), dy[sl], out_b * (size_t)nk, cudaMemcpyDeviceToHost, st[sl]);
        if (e != cudaSuccess) { rc = wx_fail(WX_ECUDA, "D2H: %s", cudaGetErrorString(e)); break; }
    }
    for (int i = 0; i < slots && rc == WX_OK; ++i) {
        cudaError_t e = cudaStreamSynchronize(st[i]);
        if (e != cudaSuccess) rc = wx_fail(WX_ECUDA, "sync: %s", cudaGetErrorString(e));
    }
    cleanup();
    return rc;
}

}  // namespace

extern "C" {
int wx_wpdall_host_f64(double *y, const double *x, long n, int L, long N, const double *h, const double *g, int F, long chunk)
{
    return wpdall_host<double>(y, x, n, L, N, h, g, F, chunk, wx_wpd1d_f64);
}
int wx_wpdall_host_f32(float *y, const float *x, long n, int L, long N, const double *h, const double *g, int F, long chunk)
{
    return wpdall_host<float>(y, x, n, L, N, h, g, F, chunk, wx_wpd1d_f32);
}
}

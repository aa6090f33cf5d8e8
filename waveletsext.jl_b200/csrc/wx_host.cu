// wx_host.cu -- host-buffer entry points: what a drop-in caller with ordinary (host) arrays invokes.
// wpdall: the batch is cut into chunks that cycle through three device slots, each with its own stream
// (H2D -> fused kernel -> D2H), so copies in both directions overlap the kernel of the neighbouring chunks.
// Reference: wpdall dwt/dwt_all.jl:260-282 (x and y are host arrays there).
#include "wx_common.cuh"
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

namespace {

constexpr int kSlots = 3;

// Level 0 of the packet table is a bit copy of x (y[:,1,k] = x[:,k], DWT.jl:141) and x already lives in host memory: a few host
// threads fill those rows straight from x while the DMA engine brings back levels 1..L only, so the PCIe-bound call moves
// L/(L+1) of the table instead of all of it.
template <typename T>
void level0_rows(T *y, const T *x, long n, int L, long k0, long k1)
{
    for (long k = k0; k < k1; ++k) std::memcpy(y + (size_t)k * (L + 1) * n, x + (size_t)k * n, (size_t)n * sizeof(T));
}

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
    WxDev dv; { int rc0 = wx_devinfo(dv); if (rc0) return rc0; }      // the library's scratch pool keeps freed slots cached across calls
    cudaStream_t st[kSlots] = {nullptr, nullptr, nullptr};
    T *dx[kSlots] = {nullptr, nullptr, nullptr}, *dy[kSlots] = {nullptr, nullptr, nullptr};
    int rc = WX_OK;
    auto cleanup = [&]() {
        for (int i = 0; i < kSlots; ++i) {
            if (st[i]) cudaStreamSynchronize(st[i]);
            if (dx[i]) cudaFreeAsync(dx[i], st[i]);         // back to the library's scratch pool (wx_devinfo):
            if (dy[i]) cudaFreeAsync(dy[i], st[i]);         // the next call reuses the slots instead of paying cudaMalloc again
            if (st[i]) { cudaStreamSynchronize(st[i]); cudaStreamDestroy(st[i]); }
        }
    };
    const long nchunks = (N + chunk - 1) / chunk;
    const int slots = nchunks < kSlots ? (int)nchunks : kSlots;
    for (int i = 0; i < slots; ++i) {
        cudaError_t e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaMallocFromPoolAsync((void **)&dx[i], in_b * (size_t)chunk, dv.pool, st[i]);
        if (e == cudaSuccess) e = cudaMallocFromPoolAsync((void **)&dy[i], out_b * (size_t)chunk, dv.pool, st[i]);
        if (e != cudaSuccess) {
            cleanup();
            cudaGetLastError();
            return wx_fail(e == cudaErrorMemoryAllocation ? WX_ENOMEM : WX_ECUDA, "wpdall_host setup: %s", cudaGetErrorString(e));
        }
    }
    // level 0 on the host (see level0_rows); WX_B200_HOST_LEVEL0=0 sends the whole table over PCIe instead (A-B measurements)
    static const char *l0env = getenv("WX_B200_HOST_LEVEL0");
    const bool host_l0 = !(l0env && l0env[0] == '0') && (const void *)y != (const void *)x;
    std::vector<std::thread> workers;
    if (host_l0) {
        unsigned hc = std::thread::hardware_concurrency();
        long nw = ((size_t)N * in_b >= ((size_t)64 << 20) && hc >= 4) ? 2 : 1;      // 2.1 GB of rows take ~0.15 s on two threads, the D2H ~0.45 s
        if (nw > N) nw = N;
        try {
            for (long w = 0; w < nw; ++w) workers.emplace_back(level0_rows<T>, y, x, n, L, N * w / nw, N * (w + 1) / nw);
        } catch (...) {                                       // no thread could be started: the rest is copied here at the end
        }
        if (workers.empty()) level0_rows<T>(y, x, n, L, 0, N);
        else if ((long)workers.size() < nw) level0_rows<T>(y, x, n, L, N * (long)workers.size() / nw, N);
    }
    for (long c = 0; c < nchunks && rc == WX_OK; ++c) {
        const int sl = (int)(c % slots);
        const long k0 = c * chunk, nk = (N - k0 < chunk) ? N - k0 : chunk;
        cudaError_t e = cudaMemcpyAsync(dx[sl], x + k0 * n, in_b * (size_t)nk, cudaMemcpyHostToDevice, st[sl]);
        if (e != cudaSuccess) { rc = wx_fail(WX_ECUDA, "H2D: %s", cudaGetErrorString(e)); break; }
        rc = kern(dy[sl], dx[sl], n, L, nk, h, g, F, (void *)st[sl]);
        if (rc) break;
        if (!host_l0) e = cudaMemcpyAsync(y + k0 * n * (L + 1), dy[sl], out_b * (size_t)nk, cudaMemcpyDeviceToHost, st[sl]);
        else if (L > 0) e = cudaMemcpy2DAsync(y + k0 * n * (L + 1) + n, out_b, dy[sl] + n, out_b, in_b * (size_t)L, (size_t)nk, cudaMemcpyDeviceToHost, st[sl]);
        if (e != cudaSuccess) { rc = wx_fail(WX_ECUDA, "D2H: %s", cudaGetErrorString(e)); break; }
    }
    for (int i = 0; i < slots && rc == WX_OK; ++i) {
        cudaError_t e = cudaStreamSynchronize(st[i]);
        if (e != cudaSuccess) rc = wx_fail(WX_ECUDA, "sync: %s", cudaGetErrorString(e));
    }
    for (auto &w : workers) w.join();
    cleanup();
    return rc;
}

// wpdall -> bestbasistree(JBB | LSDB) -> getbasiscoefall for HOST arrays: the packet table (N, L+1, n) is built chunk by chunk
// while the next chunk of x is still crossing PCIe, stays in HBM for the cost-tree reduction (NCCL exchange when comm spans
// several ranks) and only the best-basis coefficients (as many bytes as x) travel back.
// Reference: the pipeline of paper/paper.md:60-118 = wpdall dwt/dwt_all.jl:260-282, bestbasistree BestBasis.jl:185-217,
// getbasiscoefall Utils.jl:169-197.
template <typename T>
int wpd_bestbasis_host(wx_comm_t *comm, T *coef, unsigned char *tree, long ntree, const T *x, long n, int L, long N, const double *h,
                       const double *g, int F, int method, int cost_kind, double p, long chunk,
                       int (*wpd)(T *, const T *, long, int, long, const double *, const double *, int, void *),
                       int (*bbt)(wx_comm_t *, int, unsigned char *, long, double *, const T *, long, long, int, long, int, int, double, void *),
                       int (*gat)(T *, const T *, long, long, int, long, const unsigned char *, long, void *))
{
    WX_REQUIRE(n >= 2 && N >= 0, "bad sizes n=%ld N=%ld", n, N);
    WX_REQUIRE(L >= 0 && L <= wx_maxlevels(n), "AssertionError: 0 <= L <= maxtransformlevels(x) (n=%ld, L=%d)", n, L);
    WX_REQUIRE(tree && ntree == n - 1, "tree buffer must hold n-1 = %ld entries", n - 1);
    WX_REQUIRE(N == 0 || (coef && x), "null host pointer");
    const size_t in_b = (size_t)n * sizeof(T), row_b = in_b * (size_t)(L + 1);
    if (chunk <= 0) { chunk = (long)(((size_t)256 << 20) / in_b); if (chunk < 1) chunk = 1; }
    if (chunk > N && N > 0) chunk = N;
    WxDev dv; { int rc0 = wx_devinfo(dv); if (rc0) return rc0; }
    cudaStream_t st[2] = {nullptr, nullptr};
    T *dx[2] = {nullptr, nullptr}, *y = nullptr;
    cudaEvent_t ev = nullptr;
    int rc = WX_OK;
    auto fail = [&](cudaError_t e, const char *what) { rc = wx_fail(e == cudaErrorMemoryAllocation ? WX_ENOMEM : WX_ECUDA, "%s: %s", what, cudaGetErrorString(e)); cudaGetLastError(); };
    cudaError_t e = cudaSuccess;
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaStreamCreateWithFlags(&st[i], cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ev, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaMallocFromPoolAsync((void **)&y, row_b * (size_t)(N > 0 ? N : 1), dv.pool, st[0]);
    for (int i = 0; i < 2 && e == cudaSuccess; ++i) e = cudaMallocFromPoolAsync((void **)&dx[i], in_b * (size_t)(chunk > 0 ? chunk : 1), dv.pool, st[i]);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st[0]);          // y is used from both streams
    if (e != cudaSuccess) fail(e, "wpd_bestbasis_host setup");
    const long nchunks = N > 0 ? (N + chunk - 1) / chunk : 0;
    for (long c = 0; c < nchunks && rc == WX_OK; ++c) {
        const int sl = (int)(c & 1);
        const long k0 = c * chunk, nk = (N - k0 < chunk) ? N - k0 : chunk;
        e = cudaMemcpyAsync(dx[sl], x + k0 * n, in_b * (size_t)nk, cudaMemcpyHostToDevice, st[sl]);
        if (e != cudaSuccess) { fail(e, "H2D"); break; }
        rc = wpd(y + (size_t)k0 * (L + 1) * n, dx[sl], n, L, nk, h, g, F, (void *)st[sl]);
    }
    if (rc == WX_OK && st[1]) {                                       // the reduction on st[0] sees every chunk
        e = cudaEventRecord(ev, st[1]);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(st[0], ev, 0);
        if (e != cudaSuccess) fail(e, "join");
    }
    if (rc == WX_OK) rc = bbt(comm, method, tree, ntree, nullptr, y, 0, n, L + 1, N, 0, cost_kind, p, (void *)st[0]);   // synchronises st[0]
    for (long c = 0; c < nchunks && rc == WX_OK; ++c) {
        const int sl = (int)(c & 1);
        const long k0 = c * chunk, nk = (N - k0 < chunk) ? N - k0 : chunk;
        rc = gat(dx[sl], y + (size_t)k0 * (L + 1) * n, 0, n, L + 1, nk, tree, ntree, (void *)st[sl]);
        if (rc) break;
        e = cudaMemcpyAsync(coef + k0 * n, dx[sl], in_b * (size_t)nk, cudaMemcpyDeviceToHost, st[sl]);
        if (e != cudaSuccess) { fail(e, "D2H"); break; }
    }
    for (int i = 0; i < 2; ++i) {
        if (!st[i]) continue;
        e = cudaStreamSynchronize(st[i]);
        if (e != cudaSuccess && rc == WX_OK) fail(e, "sync");
        if (dx[i]) cudaFreeAsync(dx[i], st[i]);
        if (i == 0 && y) cudaFreeAsync(y, st[0]);
        cudaStreamSynchronize(st[i]);
        cudaStreamDestroy(st[i]);
    }
    if (ev) cudaEventDestroy(ev);
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
int wx_wpd_bestbasis_host_f64(wx_comm_t *comm, double *coef_host, unsigned char *tree_out, long ntree, const double *x_host, long n, int L, long N,
                               const double *h, const double *g, int F, int method, int cost_kind, double p, long chunk)
{
    return wpd_bestbasis_host<double>(comm, coef_host, tree_out, ntree, x_host, n, L, N, h, g, F, method, cost_kind, p, chunk, wx_wpd1d_f64,
                                      wx_bestbasistree_f64, wx_gather_basis_f64);
}
int wx_wpd_bestbasis_host_f32(wx_comm_t *comm, float *coef_host, unsigned char *tree_out, long ntree, const float *x_host, long n, int L, long N,
                               const double *h, const double *g, int F, int method, int cost_kind, double p, long chunk)
{
    return wpd_bestbasis_host<float>(comm, coef_host, tree_out, ntree, x_host, n, L, N, h, g, F, method, cost_kind, p, chunk, wx_wpd1d_f32,
                                     wx_bestbasistree_f32, wx_gather_basis_f32);
}
}

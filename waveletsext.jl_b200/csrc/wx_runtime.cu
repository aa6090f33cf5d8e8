// wx_runtime.cu -- runtime part of the C ABI (device memory, copies, sync, error text).
#include "wx_common.cuh"
#include <cstdlib>

thread_local char wx_errbuf[512] = "";
std::atomic<unsigned long long> wx_launches{0};

// Per-device context, created once under a mutex (two host threads may make their first call at the same time).
// Library scratch is stream-ordered memory of a PRIVATE pool per device: the process-wide default pool (and whoever else uses
// it) keeps its own release threshold.  Freed blocks stay cached in the pool up to a quarter of the device memory (so a
// repeated call does not re-map its workspace after every synchronisation); WX_B200_SCRATCH_KEEP_MB overrides the bound and
// wx_trim_scratch hands the cache back to the driver.
#include <mutex>
static std::mutex wx_dev_mutex;
static WxDev wx_dev_cache[64];
static bool wx_dev_have[64] = {false};

int wx_devinfo(WxDev &d)
{
    int dev = 0;
    WX_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return wx_fail(WX_EUNSUPPORTED, "device index %d outside 0..63", dev);
    std::lock_guard<std::mutex> lock(wx_dev_mutex);
    if (wx_dev_have[dev]) { d = wx_dev_cache[dev]; return WX_OK; }
    int sms = 0, smem = 0;
    WX_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    WX_CUDA(cudaDeviceGetAttribute(&smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    d.sms = sms; d.smem_optin = (size_t)smem; d.dev = dev; d.pool = nullptr;
    cudaMemPoolProps props;
    memset(&props, 0, sizeof(props));
    props.allocType = cudaMemAllocationTypePinned;
    props.handleTypes = cudaMemHandleTypeNone;
    props.location.type = cudaMemLocationTypeDevice;
    props.location.id = dev;
    WX_CUDA(cudaMemPoolCreate(&d.pool, &props));
    size_t freeb = 0, totalb = 0;
    unsigned long long keep = 0;
    if (cudaMemGetInfo(&freeb, &totalb) == cudaSuccess) keep = (unsigned long long)(totalb / 4);
    if (const char *env = getenv("WX_B200_SCRATCH_KEEP_MB")) keep = (unsigned long long)atoll(env) << 20;
    WX_CUDA(cudaMemPoolSetAttribute(d.pool, cudaMemPoolAttrReleaseThreshold, &keep));
    wx_dev_cache[dev] = d; wx_dev_have[dev] = true;
    return WX_OK;
}

int wx_pool_alloc(void **p, size_t bytes, cudaStream_t s)
{
    *p = nullptr;
    WxDev d; int rc = wx_devinfo(d); if (rc) return rc;
    cudaError_t e = cudaMallocFromPoolAsync(p, bytes ? bytes : 1, d.pool, s);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return wx_fail(WX_ENOMEM, "scratch of %zu bytes: out of device memory", bytes); }
    WX_CUDA(e);
    return WX_OK;
}

// ---- residency tuning (see wx_common.cuh) ----------------------------------------------------------------
#include <map>
#include <tuple>
#include <vector>
#include <algorithm>
static std::mutex wx_tune_mutex;
static std::map<std::tuple<const void *, int, long, long, long, long>, int> wx_tune_cache;

int wx_tuned_choice(const WxTuneKey &key, int ncand, const int *cand, int fallback, bool big_enough, cudaStream_t s,
                    const std::function<int(int)> &launch, int *choice)
{
    int dev = 0;
    WX_CUDA(cudaGetDevice(&dev));
    const auto k = std::make_tuple(key.kernel, dev, key.a, key.b, key.c, key.d);
    {
        std::lock_guard<std::mutex> lock(wx_tune_mutex);
        auto it = wx_tune_cache.find(k);
        if (it != wx_tune_cache.end()) { *choice = it->second; return WX_OK; }
    }
    *choice = fallback;
    const char *env = getenv("WX_B200_AUTOTUNE");
    if (!big_enough || ncand < 2 || (env && atoi(env) == 0)) return WX_OK;
    cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(s, &cap) != cudaSuccess || cap != cudaStreamCaptureStatusNone) { cudaGetLastError(); return WX_OK; }
    cudaEvent_t ev[4];
    for (auto &e : ev) WX_CUDA(cudaEventCreate(&e));
    int rc = WX_OK, bestc = fallback;
    float best = 0.f;
    // time `reps` back-to-back launches of one candidate after a warm-up launch; the minimum counts
    auto measure = [&](int c, int reps, float *tmin) -> int {
        int r = launch(c);                                                // warm-up of this candidate (also pages the code in)
        if (r) return r;
        cudaEventRecord(ev[0], s);
        for (int i = 0; i < reps; ++i) {
            r = launch(c);
            if (r) return r;
            cudaEventRecord(ev[i + 1], s);
        }
        if (cudaEventSynchronize(ev[reps]) != cudaSuccess) return wx_fail(WX_ECUDA, "autotune: %s", cudaGetErrorString(cudaGetLastError()));
        *tmin = 1e30f;
        for (int i = 0; i < reps; ++i) { float t = 0.f; cudaEventElapsedTime(&t, ev[i], ev[i + 1]); if (t < *tmin) *tmin = t; }
        return WX_OK;
    };
    // pass 1: every candidate, two timed launches; pass 2: the three fastest and the caller's rule again, three timed launches each
    // (launch-to-launch noise is a few per cent and the candidates are often that close)
    std::vector<std::pair<float, int>> first;
    for (int i = 0; i < ncand && rc == WX_OK; ++i) {
        float t = 0.f;
        rc = measure(cand[i], 2, &t);
        if (!rc) first.push_back({t, cand[i]});
    }
    float tfall = -1.f;
    if (!rc) {
        std::sort(first.begin(), first.end());
        std::vector<int> finalists;
        for (size_t i = 0; i < first.size() && finalists.size() < 3; ++i) finalists.push_back(first[i].second);
        bool has_fall = false;
        for (int c : finalists) has_fall |= (c == fallback);
        for (int i = 0; i < ncand && !has_fall; ++i) if (cand[i] == fallback) { finalists.push_back(fallback); has_fall = true; }
        for (size_t i = 0; i < finalists.size() && rc == WX_OK; ++i) {
            float t = 0.f;
            rc = measure(finalists[i], 3, &t);
            if (rc) break;
            if (i == 0 || t < best) { best = t; bestc = finalists[i]; }
            if (finalists[i] == fallback) tfall = t;
        }
    }
    // leave the caller's rule unless a candidate beats it clearly
    if (!rc && tfall > 0.f && best > 0.97f * tfall) { bestc = fallback; best = tfall; }
    for (auto &e : ev) cudaEventDestroy(e);
    if (rc) return rc;
    {
        std::lock_guard<std::mutex> lock(wx_tune_mutex);
        wx_tune_cache[k] = bestc;
    }
    if (const char *v = getenv("WX_B200_AUTOTUNE_VERBOSE")) if (atoi(v)) fprintf(stderr, "[wx_b200] tuned kernel %p shape (%ld,%ld,%ld,%ld): %d (%.3f ms)\n", key.kernel, key.a, key.b, key.c, key.d, bestc, best);
    *choice = bestc;
    return WX_OK;
}

extern "C" {

int wx_version(void) { return 100; }

const char *wx_last_error(void) { return wx_errbuf; }

unsigned long long wx_launch_count(void) { return wx_launches.load(); }

int wx_device_count(int *count)
{
    WX_REQUIRE(count, "null count");
    *count = 0;
    WX_CUDA(cudaGetDeviceCount(count));
    return WX_OK;
}

int wx_set_device(int dev)
{
    WX_CUDA(cudaSetDevice(dev));
    return WX_OK;
}

int wx_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *smem_optin, size_t *total_mem)
{
    int dev = 0;
    WX_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp p;
    WX_CUDA(cudaGetDeviceProperties(&p, dev));
    if (sm_count) *sm_count = p.multiProcessorCount;
    if (cc_major) *cc_major = p.major;
    if (cc_minor) *cc_minor = p.minor;
    if (smem_optin) *smem_optin = p.sharedMemPerBlockOptin;
    if (total_mem) *total_mem = p.totalGlobalMem;
    return WX_OK;
}

int wx_malloc(void **dptr, size_t bytes)
{
    WX_REQUIRE(dptr, "null dptr");
    *dptr = nullptr;
    cudaError_t e = cudaMalloc(dptr, bytes ? bytes : 1);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return wx_fail(WX_ENOMEM, "cudaMalloc(%zu) out of memory", bytes); }
    WX_CUDA(e);
    return WX_OK;
}

int wx_trim_scratch(size_t keep_bytes)
{
    WxDev d; int rc = wx_devinfo(d); if (rc) return rc;
    WX_CUDA(cudaDeviceSynchronize());
    WX_CUDA(cudaMemPoolTrimTo(d.pool, keep_bytes));
    return WX_OK;
}

int wx_free(void *dptr)
{
    if (dptr) WX_CUDA(cudaFree(dptr));
    return WX_OK;
}

int wx_malloc_host(void **hptr, size_t bytes)
{
    WX_REQUIRE(hptr, "null hptr");
    *hptr = nullptr;
    cudaError_t e = cudaHostAlloc(hptr, bytes ? bytes : 1, cudaHostAllocDefault);
    if (e == cudaErrorMemoryAllocation) { cudaGetLastError(); return wx_fail(WX_ENOMEM, "cudaHostAlloc(%zu) out of memory", bytes); }
    WX_CUDA(e);
    return WX_OK;
}

int wx_free_host(void *hptr)
{
    if (hptr) WX_CUDA(cudaFreeHost(hptr));
    return WX_OK;
}

int wx_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream)
{
    WX_CUDA(cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream));
    return WX_OK;
}

int wx_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream)
{
    WX_CUDA(cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return WX_OK;
}

int wx_stream_sync(void *stream)
{
    WX_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return WX_OK;
}

int wx_device_sync(void)
{
    WX_CUDA(cudaDeviceSynchronize());
    return WX_OK;
}

}  // extern "C"

// ---- TMA row maps (see wx_tma.cuh) ----------------------------------------------------------------------
#include "wx_tma.cuh"
typedef CUresult (*wx_encode_tiled_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                       const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                       CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int wx_make_rowmap(CUtensorMap *map, const void *base, size_t elt, long rows, long boxrows)
{
    static wx_encode_tiled_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (wx_encode_tiled_fn)p;
        else
            cudaGetLastError();
    }
    if (!fn) return wx_fail(WX_EUNSUPPORTED, "cuTensorMapEncodeTiled unavailable");
    if (((uintptr_t)base & 15) != 0 || rows < 1 || boxrows < 1 || boxrows > 256) return wx_fail(WX_EUNSUPPORTED, "row map: bad geometry");
    const cuuint64_t gdim[2] = {(cuuint64_t)(128 / elt), (cuuint64_t)rows};
    const cuuint64_t gstr[1] = {128};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / elt), (cuuint32_t)boxrows};
    const cuuint32_t estr[2] = {1, 1};
    const CUtensorMapDataType dt = elt == 8 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT64 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
    CUresult r = fn(map, dt, 2, const_cast<void *>(base), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return wx_fail(WX_EUNSUPPORTED, "cuTensorMapEncodeTiled failed (%d)", (int)r);
    return WX_OK;
}

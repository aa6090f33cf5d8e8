// wx_comm.cu -- the one collective of the path, inside the library: NCCL communicators behind the C ABI and the fused
// best-basis drivers (local reduction kernels -> NCCL exchange of the small per-position state -> per-node costs ->
// bestbasis_treeselection) with no host language in between.
//
// Reference: tree_costs(X, ::JBB) bestbasis/bestbasis_tree.jl:150-207 (sum(X, dims=3) spans the whole batch),
// tree_costs(X, ::LSDB) :104-147 + DifferentialEntropyCost bestbasis/bestbasis_costs.jl:135-164 (per-position statistics
// over the whole batch), bestbasistree BestBasis.jl:185-217.
//
// NCCL is resolved at run time (dlopen of the libnccl.so.2 already mapped by the host process, else WX_B200_NCCL, else the
// loader path): libwx_b200.so carries no link-time dependency on it, single-GPU hosts never touch it, and a multi-GPU host
// that cannot provide it gets WX_EUNSUPPORTED with a message -- never a silent fallback.
#include "wx_common.cuh"
#include "wx_steps.cuh"
#include <dlfcn.h>
#include <mutex>
#include <vector>
#include <cstdlib>

// minimal declarations of the stable NCCL 2.x ABI used here (nccl.h is not required to build the library)
typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { wxNcclUint8 = 1, wxNcclInt64 = 4, wxNcclFloat32 = 7, wxNcclFloat64 = 8 };
enum { wxNcclSum = 0, wxNcclMax = 2, wxNcclMin = 3 };

namespace {

struct Nccl {
    void *handle = nullptr;
    ncclResult_t (*GetVersion)(int *) = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(ncclResult_t) = nullptr;
    char why[256] = "";
};

Nccl g_nccl;
std::once_flag g_nccl_once;

void nccl_load()
{
    Nccl &N = g_nccl;
    const char *env = getenv("WX_B200_NCCL");
    void *h = nullptr;
    if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);      // the copy the host process (e.g. torch) already mapped
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { snprintf(N.why, sizeof(N.why), "libnccl.so.2 not found (set WX_B200_NCCL): %s", dlerror()); return; }
    bool ok = true;
#define WX_SYM(name) do { *(void **)(&N.name) = dlsym(h, "nccl" #name); if (!N.name) { ok = false; snprintf(N.why, sizeof(N.why), "nccl" #name " missing"); } } while (0)
    WX_SYM(GetVersion); WX_SYM(GetUniqueId); WX_SYM(CommInitRank); WX_SYM(CommInitAll); WX_SYM(CommDestroy); WX_SYM(AllReduce);
    WX_SYM(AllGather); WX_SYM(Broadcast); WX_SYM(GroupStart); WX_SYM(GroupEnd); WX_SYM(GetErrorString);
#undef WX_SYM
    if (ok) N.handle = h;
}

int nccl_get(Nccl **out)
{
    std::call_once(g_nccl_once, nccl_load);
    if (!g_nccl.handle) return wx_fail(WX_EUNSUPPORTED, "NCCL unavailable: %s", g_nccl.why);
    *out = &g_nccl;
    return WX_OK;
}

#define WX_NCCL(N, expr)                                                                                          \
    do {                                                                                                          \
        ncclResult_t _r = (expr);                                                                                 \
        if (_r != 0) return wx_fail(WX_ECUDA, "%s failed: %s (%s:%d)", #expr, (N)->GetErrorString(_r), __FILE__, __LINE__); \
    } while (0)

}  // namespace

struct wx_comm {
    ncclComm_t nccl;
    int rank, world, dev;
    cudaStream_t stream;      // the communicator's own non-blocking stream (what the *_multi drivers launch on; wx_comm_info returns it)
};

namespace {

int dtype_of(int dt, int *nd, size_t *sz)
{
    switch (dt) {
        case WX_DT_F64: *nd = wxNcclFloat64; *sz = 8; return WX_OK;
        case WX_DT_F32: *nd = wxNcclFloat32; *sz = 4; return WX_OK;
        case WX_DT_I64: *nd = wxNcclInt64; *sz = 8; return WX_OK;
        case WX_DT_U8: *nd = wxNcclUint8; *sz = 1; return WX_OK;
    }
    return wx_fail(WX_EINVAL, "unknown dtype code %d", dt);
}

int op_of(int op, int *no)
{
    switch (op) {
        case WX_OP_SUM: *no = wxNcclSum; return WX_OK;
        case WX_OP_MIN: *no = wxNcclMin; return WX_OK;
        case WX_OP_MAX: *no = wxNcclMax; return WX_OK;
    }
    return wx_fail(WX_EINVAL, "unknown reduction op %d", op);
}

struct DevGuard {            // the multi-device drivers hop between devices; restore the caller's on exit
    int saved = -1;
    DevGuard() { cudaGetDevice(&saved); }
    ~DevGuard() { if (saved >= 0) cudaSetDevice(saved); }
};

template <typename T> __global__ void to_f64_k(double *out, const T *in, long n)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = (double)in[i];
}
__global__ void set_f64_k(double *p, double v) { *p = v; }
__global__ void set_i64_k(long *p, long v) { *p = v; }

// one rank of a best-basis reduction: a shard of the packet table on one device
template <typename T>
struct Shard {
    wx_comm *c;               // NULL: single GPU, no exchange
    cudaStream_t s;
    const T *X;
    long Nloc;
    // state
    double *buf = nullptr, *parts = nullptr, *counts = nullptr, *logsum = nullptr;
    long *cnt = nullptr;
};

template <typename T> int on(const Shard<T> &sh) { if (sh.c) WX_CUDA(cudaSetDevice(sh.c->dev)); return WX_OK; }

template <typename T> struct Fn;
template <> struct Fn<double> {
    static constexpr int elt = 8;
    static int moments(double *a, double *b, const double *X, long szK, long N, void *s) { return wx_jbb_moments_f64(a, b, X, szK, N, s); }
    static int p1(double *st, const double *X, long szK, long N, void *s) { return wx_lsdb_pass1_f64(st, X, szK, N, s); }
    static int p2(double *c, const double *st, const double *X, long szK, long N, long Nt, void *s) { return wx_lsdb_pass2_f64(c, st, X, szK, N, Nt, s); }
    static int p3(double *l, const double *c, const double *st, const double *X, long szK, long N, long Nt, void *s) { return wx_lsdb_pass3_f64(l, c, st, X, szK, N, Nt, s); }
};
template <> struct Fn<float> {
    static constexpr int elt = 4;
    static int moments(double *a, double *b, const float *X, long szK, long N, void *s) { return wx_jbb_moments_f32(a, b, X, szK, N, s); }
    static int p1(double *st, const float *X, long szK, long N, void *s) { return wx_lsdb_pass1_f32(st, X, szK, N, s); }
    static int p2(double *c, const double *st, const float *X, long szK, long N, long Nt, void *s) { return wx_lsdb_pass2_f32(c, st, X, szK, N, Nt, s); }
    static int p3(double *l, const double *c, const double *st, const float *X, long szK, long N, long Nt, void *s) { return wx_lsdb_pass3_f32(l, c, st, X, szK, N, Nt, s); }
};

template <typename T>
void release(std::vector<Shard<T>> &sh)
{
    for (auto &q : sh) {
        if (q.c) cudaSetDevice(q.c->dev);
        if (q.buf) cudaFreeAsync(q.buf, q.s);
        if (q.parts) cudaFreeAsync(q.parts, q.s);
        if (q.counts) cudaFreeAsync(q.counts, q.s);
        if (q.logsum) cudaFreeAsync(q.logsum, q.s);
        if (q.cnt) cudaFreeAsync(q.cnt, q.s);
        q.buf = q.parts = q.counts = q.logsum = nullptr; q.cnt = nullptr;
    }
}

// tree_costs(X, ::JBB) over all shards; costs (host) are written from the first shard (identical on all by construction)
template <typename T>
int jbb_costs_shards(std::vector<Shard<T>> &sh, double *costs_host, long m, long n, int K, int redundant, int cost_kind, double p, long *Ntot_out)
{
    const long szK = (m > 0 ? m : 1) * n * K;
    const int world = sh[0].c ? sh[0].c->world : 1;
    Nccl *N = nullptr;
    if (world > 1) { int rc = nccl_get(&N); if (rc) return rc; }
    for (auto &q : sh) {
        int rc = on(q); if (rc) return rc;
        rc = wx_scratch(&q.buf, (size_t)2 * szK + 1, q.s); if (rc) return rc;
        rc = Fn<T>::moments(q.buf, q.buf + szK, q.X, szK, q.Nloc, q.s); if (rc) return rc;
        set_f64_k<<<1, 1, 0, q.s>>>(q.buf + 2 * szK, (double)q.Nloc);      // the signal count rides along (exact below 2^53)
        WX_LAUNCHED();
    }
    if (world > 1) {
        // the only collective of the JBB path: 2*n*K+1 doubles (bestbasis_tree.jl:153-154 over the whole batch)
        WX_NCCL(N, N->GroupStart());
        for (auto &q : sh) {
            int rc = on(q); if (rc) return rc;
            WX_NCCL(N, N->AllReduce(q.buf, q.buf, (size_t)2 * szK + 1, wxNcclFloat64, wxNcclSum, q.c->nccl, q.s));
        }
        WX_NCCL(N, N->GroupEnd());
    }
    Shard<T> &q0 = sh[0];
    { int rc = on(q0); if (rc) return rc; }
    double nt = 0;
    WX_CUDA(cudaMemcpyAsync(&nt, q0.buf + 2 * szK, sizeof(double), cudaMemcpyDeviceToHost, q0.s));
    WX_CUDA(cudaStreamSynchronize(q0.s));
    const long Ntot = (long)nt;
    WX_REQUIRE(Ntot >= 1, "tree_costs: empty batch");
    if (Ntot_out) *Ntot_out = Ntot;
    return wx_jbb_costs(costs_host, q0.buf, q0.buf + szK, Ntot, m, n, K, redundant, cost_kind, p, Fn<T>::elt, q0.s);
}

// tree_costs(X, ::LSDB) over all shards
template <typename T>
int lsdb_costs_shards(std::vector<Shard<T>> &sh, double *costs_host, long m, long n, int K, int redundant, long *Ntot_out)
{
    const long szK = (m > 0 ? m : 1) * n * K;
    const int world = sh[0].c ? sh[0].c->world : 1;
    Nccl *N = nullptr;
    if (world > 1) { int rc = nccl_get(&N); if (rc) return rc; }
    // 1. signal counts of every rank -> total, owner of the first signal of the global batch
    std::vector<long> counts(world, 0);
    if (world > 1) {
        for (auto &q : sh) {
            int rc = on(q); if (rc) return rc;
            rc = wx_scratch(&q.cnt, (size_t)world + 1, q.s); if (rc) return rc;
            set_i64_k<<<1, 1, 0, q.s>>>(q.cnt + world, q.Nloc);
            WX_LAUNCHED();
        }
        WX_NCCL(N, N->GroupStart());
        for (auto &q : sh) {
            int rc = on(q); if (rc) return rc;
            WX_NCCL(N, N->AllGather(q.cnt + world, q.cnt, 1, wxNcclInt64, q.c->nccl, q.s));
        }
        WX_NCCL(N, N->GroupEnd());
        int rc = on(sh[0]); if (rc) return rc;
        WX_CUDA(cudaMemcpyAsync(counts.data(), sh[0].cnt, sizeof(long) * world, cudaMemcpyDeviceToHost, sh[0].s));
        WX_CUDA(cudaStreamSynchronize(sh[0].s));
    } else {
        counts[0] = sh[0].Nloc;
    }
    long Ntot = 0; int root = -1;
    for (int r = 0; r < world; ++r) { Ntot += counts[r]; if (root < 0 && counts[r] > 0) root = r; }
    WX_REQUIRE(Ntot >= 2, "LSDB needs at least two signals");
    if (Ntot_out) *Ntot_out = Ntot;
    long nb, mb, npts;
    { int rc = wx_lsdb_grid(Ntot, &nb, &mb, &npts); if (rc) return rc; }
    // 2. common shift (first signal of the global batch, bestbasis_costs.jl:143 is shift invariant; see wx_b200.h) + pass 1
    for (auto &q : sh) {
        int rc = on(q); if (rc) return rc;
        rc = wx_scratch(&q.buf, (size_t)7 * szK, q.s); if (rc) return rc;
        if (world > 1) { rc = wx_scratch(&q.parts, (size_t)world * 2 * szK, q.s); if (rc) return rc; }
        rc = wx_scratch(&q.counts, (size_t)npts * szK, q.s); if (rc) return rc;
        rc = wx_scratch(&q.logsum, (size_t)2 * szK, q.s); if (rc) return rc;
        const int myrank = q.c ? q.c->rank : 0;
        if (myrank == root) {
            to_f64_k<T><<<(unsigned)((szK + 255) / 256), 256, 0, q.s>>>(q.buf, q.X, szK);
            WX_LAUNCHED();
        }
    }
    if (world > 1) {
        WX_NCCL(N, N->GroupStart());
        for (auto &q : sh) {
            int rc = on(q); if (rc) return rc;
            WX_NCCL(N, N->Broadcast(q.buf, q.buf, (size_t)szK, wxNcclFloat64, root, q.c->nccl, q.s));
        }
        WX_NCCL(N, N->GroupEnd());
    }
    for (auto &q : sh) {
        int rc = on(q); if (rc) return rc;
        rc = Fn<T>::p1(q.buf, q.X, szK, q.Nloc, q.s); if (rc) return rc;
    }
    // 3. sums as double-double pairs combined in rank order (the grid every sample is binned on must not depend on the
    //    sharding), min / max all-reduced
    if (world > 1) {
        for (int pair = 0; pair < 2; ++pair) {
            WX_NCCL(N, N->GroupStart());
            for (auto &q : sh) {
                int rc = on(q); if (rc) return rc;
                WX_NCCL(N, N->AllGather(q.buf + (1 + 2 * pair) * szK, q.parts, (size_t)2 * szK, wxNcclFloat64, q.c->nccl, q.s));
                if (pair == 0) {
                    WX_NCCL(N, N->AllReduce(q.buf + 5 * szK, q.buf + 5 * szK, (size_t)szK, wxNcclFloat64, wxNcclMin, q.c->nccl, q.s));
                    WX_NCCL(N, N->AllReduce(q.buf + 6 * szK, q.buf + 6 * szK, (size_t)szK, wxNcclFloat64, wxNcclMax, q.c->nccl, q.s));
                }
            }
            WX_NCCL(N, N->GroupEnd());
            for (auto &q : sh) {
                int rc = on(q); if (rc) return rc;
                rc = wx_dd_sum(q.buf + (1 + 2 * pair) * szK, q.parts, szK, world, q.s); if (rc) return rc;
            }
        }
    }
    // 4. histogram counts on the common grid (integer valued: exact in any order)
    for (auto &q : sh) {
        int rc = on(q); if (rc) return rc;
        rc = Fn<T>::p2(q.counts, q.buf, q.X, szK, q.Nloc, Ntot, q.s); if (rc) return rc;
    }
    if (world > 1) {
        WX_NCCL(N, N->GroupStart());
        for (auto &q : sh) {
            int rc = on(q); if (rc) return rc;
            WX_NCCL(N, N->AllReduce(q.counts, q.counts, (size_t)npts * szK, wxNcclFloat64, wxNcclSum, q.c->nccl, q.s));
        }
        WX_NCCL(N, N->GroupEnd());
    }
    // 5. sum of log pdf per position, again as double-double pairs in rank order
    for (auto &q : sh) {
        int rc = on(q); if (rc) return rc;
        rc = Fn<T>::p3(q.logsum, q.counts, q.buf, q.X, szK, q.Nloc, Ntot, q.s); if (rc) return rc;
    }
    if (world > 1) {
        WX_NCCL(N, N->GroupStart());
        for (auto &q : sh) {
            int rc = on(q); if (rc) return rc;
            WX_NCCL(N, N->AllGather(q.logsum, q.parts, (size_t)2 * szK, wxNcclFloat64, q.c->nccl, q.s));
        }
        WX_NCCL(N, N->GroupEnd());
        for (auto &q : sh) {
            int rc = on(q); if (rc) return rc;
            rc = wx_dd_sum(q.logsum, q.parts, szK, world, q.s); if (rc) return rc;
        }
    }
    int rc = on(sh[0]); if (rc) return rc;
    return wx_lsdb_costs(costs_host, sh[0].logsum, Ntot, m, n, K, redundant, sh[0].s);
}

long ncosts_of(long m, int K, int redundant)
{
    if (redundant) return K;
    return m > 0 ? ((1L << (2 * K)) - 1) / 3 : (1L << K) - 1;
}

long ntree_of(long m, long n)
{
    if (m == 0) return n - 1;
    const int Lm = wx_maxlevels(m < n ? m : n);
    return ((1L << (2 * Lm)) - 1) / 3;
}

// method 0 = JBB, 1 = LSDB.  costs_host (may be NULL) receives the node costs BEFORE the selection pass modifies them.
template <typename T>
int bestbasis_shards(std::vector<Shard<T>> &sh, int method, unsigned char *tree_out, long ntree, double *costs_host, long m, long n, int K,
                     int redundant, int cost_kind, double p)
{
    WX_REQUIRE(method == 0 || method == 1, "method must be 0 (JBB) or 1 (LSDB)");
    WX_REQUIRE(n >= 1 && m >= 0 && K >= 1, "bad sizes");
    if (!redundant) WX_REQUIRE(m > 0 ? 2 * K < 62 : K < 62, "too many levels");
    for (auto &q : sh) WX_REQUIRE(q.Nloc >= 0 && (q.Nloc == 0 || q.X), "null shard");
    const long nc = ncosts_of(m, K, redundant);
    if (tree_out) WX_REQUIRE(ntree == ntree_of(m, n), "tree buffer must hold %ld entries, got %ld", ntree_of(m, n), ntree);
    std::vector<double> costs((size_t)nc);
    DevGuard guard;
    int rc = method == 0 ? jbb_costs_shards(sh, costs.data(), m, n, K, redundant, cost_kind, p, nullptr)
                         : lsdb_costs_shards(sh, costs.data(), m, n, K, redundant, nullptr);
    release(sh);
    if (rc) return rc;
    if (costs_host) memcpy(costs_host, costs.data(), sizeof(double) * (size_t)nc);
    if (!tree_out) return WX_OK;
    // bestbasis_treeselection(costs, n[, m])  BestBasis.jl:59-110 (redundant tables carry one cost per node column, same heap order)
    return wx_tree_select(tree_out, costs.data(), nc, m, n, 0);
}

template <typename T>
int bestbasis_one(wx_comm *c, int method, unsigned char *tree_out, long ntree, double *costs_host, const T *X, long m, long n, int K, long Nlocal,
                  int redundant, int cost_kind, double p, void *stream)
{
    std::vector<Shard<T>> sh(1);
    sh[0].c = (c && c->world > 1) ? c : nullptr;
    if (c && c->world > 1) {
        int dev = -1; WX_CUDA(cudaGetDevice(&dev));
        WX_REQUIRE(dev == c->dev, "communicator lives on device %d, current device is %d", c->dev, dev);
    }
    sh[0].s = (cudaStream_t)stream; sh[0].X = X; sh[0].Nloc = Nlocal;
    return bestbasis_shards(sh, method, tree_out, ntree, costs_host, m, n, K, redundant, cost_kind, p);
}

template <typename T>
int bestbasis_multi(wx_comm *const *comms, int ndev, int method, unsigned char *tree_out, long ntree, double *costs_host, const T *const *X,
                    const long *Nlocal, long m, long n, int K, int redundant, int cost_kind, double p)
{
    WX_REQUIRE(comms && ndev >= 1 && X && Nlocal, "null argument");
    std::vector<Shard<T>> sh((size_t)ndev);
    for (int i = 0; i < ndev; ++i) {
        WX_REQUIRE(comms[i] && comms[i]->world == ndev && comms[i]->rank == i, "comms must be the %d communicators of wx_comm_init_all, in rank order", ndev);
        sh[i].c = ndev > 1 ? comms[i] : nullptr;
        sh[i].s = comms[i]->stream; sh[i].X = X[i]; sh[i].Nloc = Nlocal[i];
    }
    if (ndev == 1) {      // no exchange, but the shard still lives on the communicator's device
        DevGuard guard;
        WX_CUDA(cudaSetDevice(comms[0]->dev));
        return bestbasis_shards(sh, method, tree_out, ntree, costs_host, m, n, K, redundant, cost_kind, p);
    }
    return bestbasis_shards(sh, method, tree_out, ntree, costs_host, m, n, K, redundant, cost_kind, p);
}

}  // namespace

extern "C" {

int wx_nccl_version(int *version)
{
    WX_REQUIRE(version, "null version");
    Nccl *N; int rc = nccl_get(&N); if (rc) return rc;
    WX_NCCL(N, N->GetVersion(version));
    return WX_OK;
}

int wx_comm_unique_id(unsigned char *id128)
{
    WX_REQUIRE(id128, "null id");
    Nccl *N; int rc = nccl_get(&N); if (rc) return rc;
    ncclUniqueId id;
    WX_NCCL(N, N->GetUniqueId(&id));
    memcpy(id128, id.internal, WX_COMM_ID_BYTES);
    return WX_OK;
}

int wx_comm_init_rank(wx_comm_t **comm, const unsigned char *id128, int rank, int world)
{
    WX_REQUIRE(comm && id128, "null argument");
    WX_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d of %d", rank, world);
    *comm = nullptr;
    Nccl *N; int rc = nccl_get(&N); if (rc) return rc;
    int dev = 0;
    WX_CUDA(cudaGetDevice(&dev));
    ncclUniqueId id;
    memcpy(id.internal, id128, WX_COMM_ID_BYTES);
    wx_comm *c = new wx_comm{nullptr, rank, world, dev, nullptr};
    ncclResult_t r = N->CommInitRank(&c->nccl, world, id, rank);
    if (r != 0) { delete c; return wx_fail(WX_ECUDA, "ncclCommInitRank failed: %s", N->GetErrorString(r)); }
    cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { N->CommDestroy(c->nccl); delete c; return wx_fail(WX_ECUDA, "cudaStreamCreate: %s", cudaGetErrorString(e)); }
    *comm = c;
    return WX_OK;
}

int wx_comm_init_all(wx_comm_t **comms, int ndev, const int *devlist)
{
    WX_REQUIRE(comms && ndev >= 1 && ndev <= 64, "bad arguments");
    for (int i = 0; i < ndev; ++i) comms[i] = nullptr;
    int have = 0;
    WX_CUDA(cudaGetDeviceCount(&have));
    std::vector<int> devs((size_t)ndev);
    for (int i = 0; i < ndev; ++i) {
        devs[i] = devlist ? devlist[i] : i;
        WX_REQUIRE(devs[i] >= 0 && devs[i] < have, "device %d not present (%d devices)", devs[i], have);
    }
    DevGuard guard;
    std::vector<ncclComm_t> nc((size_t)ndev, nullptr);
    if (ndev > 1) {
        Nccl *N; int rc = nccl_get(&N); if (rc) return rc;
        WX_NCCL(N, N->CommInitAll(nc.data(), ndev, devs.data()));
    }
    for (int i = 0; i < ndev; ++i) {
        wx_comm *c = new wx_comm{nc[i], i, ndev, devs[i], nullptr};
        cudaError_t e = cudaSetDevice(devs[i]);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        comms[i] = c;
        if (e != cudaSuccess) {
            for (int j = 0; j <= i; ++j) { wx_comm_destroy(comms[j]); comms[j] = nullptr; }
            for (int j = i + 1; j < ndev; ++j) if (nc[j]) g_nccl.CommDestroy(nc[j]);
            return wx_fail(WX_ECUDA, "wx_comm_init_all: %s", cudaGetErrorString(e));
        }
    }
    return WX_OK;
}

int wx_comm_info(const wx_comm_t *comm, int *rank, int *world, int *dev, void **stream)
{
    WX_REQUIRE(comm, "null communicator");
    if (rank) *rank = comm->rank;
    if (world) *world = comm->world;
    if (dev) *dev = comm->dev;
    if (stream) *stream = (void *)comm->stream;
    return WX_OK;
}

int wx_comm_destroy(wx_comm_t *comm)
{
    if (!comm) return WX_OK;
    DevGuard guard;
    cudaSetDevice(comm->dev);
    if (comm->stream) { cudaStreamSynchronize(comm->stream); cudaStreamDestroy(comm->stream); }
    if (comm->nccl && g_nccl.handle) g_nccl.CommDestroy(comm->nccl);
    cudaGetLastError();
    delete comm;
    return WX_OK;
}

int wx_group_start(void)
{
    Nccl *N; int rc = nccl_get(&N); if (rc) return rc;
    WX_NCCL(N, N->GroupStart());
    return WX_OK;
}

int wx_group_end(void)
{
    Nccl *N; int rc = nccl_get(&N); if (rc) return rc;
    WX_NCCL(N, N->GroupEnd());
    return WX_OK;
}

int wx_allreduce(wx_comm_t *comm, void *buf, long count, int dtype, int op, void *stream)
{
    WX_REQUIRE(comm && count >= 0 && (count == 0 || buf), "bad arguments");
    int nd, no; size_t sz;
    int rc = dtype_of(dtype, &nd, &sz); if (rc) return rc;
    rc = op_of(op, &no); if (rc) return rc;
    if (comm->world == 1 || count == 0) return WX_OK;
    Nccl *N; rc = nccl_get(&N); if (rc) return rc;
    WX_NCCL(N, N->AllReduce(buf, buf, (size_t)count, nd, no, comm->nccl, (cudaStream_t)stream));
    return WX_OK;
}

int wx_allgather(wx_comm_t *comm, void *recv, const void *send, long count, int dtype, void *stream)
{
    WX_REQUIRE(comm && count >= 0 && (count == 0 || (recv && send)), "bad arguments");
    int nd; size_t sz;
    int rc = dtype_of(dtype, &nd, &sz); if (rc) return rc;
    if (count == 0) return WX_OK;
    if (comm->world == 1) {
        if (recv != send) WX_CUDA(cudaMemcpyAsync(recv, send, (size_t)count * sz, cudaMemcpyDeviceToDevice, (cudaStream_t)stream));
        return WX_OK;
    }
    Nccl *N; rc = nccl_get(&N); if (rc) return rc;
    WX_NCCL(N, N->AllGather(send, recv, (size_t)count, nd, comm->nccl, (cudaStream_t)stream));
    return WX_OK;
}

int wx_broadcast(wx_comm_t *comm, void *buf, long count, int dtype, int root, void *stream)
{
    WX_REQUIRE(comm && count >= 0 && (count == 0 || buf), "bad arguments");
    WX_REQUIRE(root >= 0 && root < comm->world, "bad root %d", root);
    int nd; size_t sz;
    int rc = dtype_of(dtype, &nd, &sz); if (rc) return rc;
    if (comm->world == 1 || count == 0) return WX_OK;
    Nccl *N; rc = nccl_get(&N); if (rc) return rc;
    WX_NCCL(N, N->Broadcast(buf, buf, (size_t)count, nd, root, comm->nccl, (cudaStream_t)stream));
    return WX_OK;
}

// ---- fused best-basis drivers ------------------------------------------------------------------------------
int wx_tree_costs_jbb_f64(wx_comm_t *comm, double *costs_host, const double *X, long m, long n, int K, long Nlocal, int redundant, int cost_kind, double p, void *stream)
{
    WX_REQUIRE(costs_host, "null costs");
    return bestbasis_one<double>(comm, 0, nullptr, 0, costs_host, X, m, n, K, Nlocal, redundant, cost_kind, p, stream);
}
int wx_tree_costs_jbb_f32(wx_comm_t *comm, double *costs_host, const float *X, long m, long n, int K, long Nlocal, int redundant, int cost_kind, double p, void *stream)
{
    WX_REQUIRE(costs_host, "null costs");
    return bestbasis_one<float>(comm, 0, nullptr, 0, costs_host, X, m, n, K, Nlocal, redundant, cost_kind, p, stream);
}
int wx_tree_costs_lsdb_f64(wx_comm_t *comm, double *costs_host, const double *X, long m, long n, int K, long Nlocal, int redundant, void *stream)
{
    WX_REQUIRE(costs_host, "null costs");
    return bestbasis_one<double>(comm, 1, nullptr, 0, costs_host, X, m, n, K, Nlocal, redundant, 0, 0.0, stream);
}
int wx_tree_costs_lsdb_f32(wx_comm_t *comm, double *costs_host, const float *X, long m, long n, int K, long Nlocal, int redundant, void *stream)
{
    WX_REQUIRE(costs_host, "null costs");
    return bestbasis_one<float>(comm, 1, nullptr, 0, costs_host, X, m, n, K, Nlocal, redundant, 0, 0.0, stream);
}
int wx_bestbasistree_f64(wx_comm_t *comm, int method, unsigned char *tree_out, long ntree, double *costs_host, const double *X, long m, long n, int K,
                         long Nlocal, int redundant, int cost_kind, double p, void *stream)
{
    WX_REQUIRE(tree_out, "null tree");
    return bestbasis_one<double>(comm, method, tree_out, ntree, costs_host, X, m, n, K, Nlocal, redundant, cost_kind, p, stream);
}
int wx_bestbasistree_f32(wx_comm_t *comm, int method, unsigned char *tree_out, long ntree, double *costs_host, const float *X, long m, long n, int K,
                         long Nlocal, int redundant, int cost_kind, double p, void *stream)
{
    WX_REQUIRE(tree_out, "null tree");
    return bestbasis_one<float>(comm, method, tree_out, ntree, costs_host, X, m, n, K, Nlocal, redundant, cost_kind, p, stream);
}
int wx_bestbasistree_multi_f64(wx_comm_t *const *comms, int ndev, int method, unsigned char *tree_out, long ntree, double *costs_host,
                               const double *const *X, const long *Nlocal, long m, long n, int K, int redundant, int cost_kind, double p)
{
    return bestbasis_multi<double>(comms, ndev, method, tree_out, ntree, costs_host, X, Nlocal, m, n, K, redundant, cost_kind, p);
}
int wx_bestbasistree_multi_f32(wx_comm_t *const *comms, int ndev, int method, unsigned char *tree_out, long ntree, double *costs_host,
                               const float *const *X, const long *Nlocal, long m, long n, int K, int redundant, int cost_kind, double p)
{
    return bestbasis_multi<float>(comms, ndev, method, tree_out, ntree, costs_host, X, Nlocal, m, n, K, redundant, cost_kind, p);
}

}  // extern "C"

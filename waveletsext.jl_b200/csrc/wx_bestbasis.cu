// wx_bestbasis.cu -- JBB / LSDB cost-tree accumulation and best-basis tree selection.
//
// Reference: tree_costs(X, ::JBB) bestbasis/bestbasis_tree.jl:150-207, tree_costs(X, ::LSDB) :104-147,
//            coefcost bestbasis/bestbasis_costs.jl:127-164, bestbasis_treeselection BestBasis.jl:59-110.
//
// X is a packet table (sz, K, N) [sz = n or m*n].  Everything that depends on the whole batch is expressed as
// per-position state that can be all-reduced across ranks (sum / min / max), so the batch can be sharded:
//   JBB : sum, sumsq                                  -> sigma -> per-node cost
//   LSDB: shifted sums, min, max -> ASH bin counts -> sum of log pdf -> per-node cost
// All floating-point reductions use a fixed two-stage order (no float atomics; the only atomics add 1.0 to
// bin counters, which is exact and order independent), so results are run-to-run and rank-count reproducible
// for a given shard size.
#include "wx_steps.cuh"
#include <vector>
#include <cmath>
#include <type_traits>

namespace {

constexpr int kT = 256;
static inline unsigned gridf(long total) { return (unsigned)((total + kT - 1) / kT); }

// ---- stage 1: per-position partial reductions over a slice of the batch ----------------------------------
// grid.x covers positions (kT per block), grid.y = ksplit.  Partial results (sum, sum of squares) go to part[ks][q][e].
template <typename T>
__global__ void __launch_bounds__(kT) moments_part_k(double *part, const T *X, long szK, long N, long kchunk)
{
    const long e = (long)blockIdx.x * kT + threadIdx.x;
    if (e >= szK) return;
    const long k0 = (long)blockIdx.y * kchunk;
    long k1 = k0 + kchunk; if (k1 > N) k1 = N;
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0, q0 = 0, q1 = 0, q2 = 0, q3 = 0;
    long k = k0;
    const T *p = X + e;
    for (; k + 4 <= k1; k += 4) {
        const double a = (double)p[k * szK], b = (double)p[(k + 1) * szK], cc = (double)p[(k + 2) * szK], d = (double)p[(k + 3) * szK];
        s0 += a; s1 += b; s2 += cc; s3 += d;
        q0 = fma(a, a, q0); q1 = fma(b, b, q1); q2 = fma(cc, cc, q2); q3 = fma(d, d, q3);
    }
    for (; k < k1; ++k) {
        const double a = (double)p[k * szK];
        s0 += a; q0 = fma(a, a, q0);
    }
    double *o = part + ((long)blockIdx.y * 2) * szK + e;
    o[0] = (s0 + s1) + (s2 + s3);
    o[szK] = (q0 + q1) + (q2 + q3);
}

// vectorised variant: a thread owns V = 16/sizeof(T) consecutive positions (128-bit loads), U signals in flight per thread.
// Per position the batch slice is accumulated in four interleaved partial sums (signal index mod 4), like the scalar kernel.
template <typename T>
__global__ void __launch_bounds__(kT) moments_part_vec_k(double *part, const T *X, long szK, long N, long kchunk)
{
    constexpr int V = 16 / (int)sizeof(T), U = 8;
    using VT = typename std::conditional<sizeof(T) == 8, double2, float4>::type;
    const long e = ((long)blockIdx.x * kT + threadIdx.x) * V;
    if (e >= szK) return;
    const long k0 = (long)blockIdx.y * kchunk;
    long k1 = k0 + kchunk; if (k1 > N) k1 = N;
    double s[4][V], q[4][V];
#pragma unroll
    for (int v = 0; v < V; ++v)
#pragma unroll
        for (int a = 0; a < 4; ++a) { s[a][v] = 0; q[a][v] = 0; }
    const T *p = X + e;
    long k = k0;
    for (; k + U <= k1; k += U) {
        VT r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = __ldcs(reinterpret_cast<const VT *>(p + (k + u) * szK));
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const T *rv = reinterpret_cast<const T *>(&r[u]);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const double a = (double)rv[v];
                s[u & 3][v] += a;
                q[u & 3][v] = fma(a, a, q[u & 3][v]);
            }
        }
    }
    for (; k < k1; ++k) {
        const VT r = __ldcs(reinterpret_cast<const VT *>(p + k * szK));
        const T *rv = reinterpret_cast<const T *>(&r);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const double a = (double)rv[v];
            s[0][v] += a; q[0][v] = fma(a, a, q[0][v]);
        }
    }
    double *o = part + ((long)blockIdx.y * 2) * szK + e;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        o[v] = (s[0][v] + s[1][v]) + (s[2][v] + s[3][v]);
        o[szK + v] = (q[0][v] + q[1][v]) + (q[2][v] + q[3][v]);
    }
}

// stage 2: combine the ksplit partials in index order
__global__ void __launch_bounds__(kT) moments_final_k(double *o0, double *o1, const double *part, long szK, int ksplit)
{
    const long e = (long)blockIdx.x * kT + threadIdx.x;
    if (e >= szK) return;
    double s = 0, q = 0;
    for (int ks = 0; ks < ksplit; ++ks) {
        const double *p = part + ((long)ks * 2) * szK + e;
        s += p[0]; q += p[szK];
    }
    o0[e] = s; o1[e] = q;
}

// ---- table-driven natural logarithm -------------------------------------------------------------------------------------
// The per-coefficient entropy terms take one log each (1.4 G of them for the 131072 x 1024 x 11 table) and libdevice's log is ~60
// instructions.  x = 2^e * m, m in [1, 2): i = the top 7 mantissa bits, c_i ~ 1 / m (table), r = m c_i - 1 with |r| < 2^-8 (one
// FMA), log x = e ln2 + (-log c_i) + log1p(r), log1p by a degree-6 Taylor polynomial (truncation < 2^-58).  Absolute error
// <= ~1.5e-16 + 1 ulp of the result (ln2 split in two parts so that e ln2_hi is exact); zero, subnormal, negative, infinite and
// NaN arguments take the libdevice path.  The same function is used on every rank and for every slice, so the bitwise
// independence of the LSDB sums from the sharding is untouched.
// generated: c_i = double(1 / (1 + (i + 0.5) / 128)), T_i = -log(c_i) (correctly rounded), i = 0..127
__device__ const double2 wx_log_tab[128] = {
    {0x1.fe01fe01fe020p-1, 0x1.ff00aa2b10ba0p-9}, {0x1.fa11caa01fa12p-1, 0x1.7dc475f810a69p-7},
    {0x1.f6310aca0dbb5p-1, 0x1.3cea44346a584p-6}, {0x1.f25f644230ab5p-1, 0x1.b9fc027af919ap-6},
    {0x1.ee9c7f8458e02p-1, 0x1.1b0d98923d97fp-5}, {0x1.eae807aba01ebp-1, 0x1.58a5bafc8e4d3p-5},
    {0x1.e741aa59750e4p-1, 0x1.95c830ec8e3f2p-5}, {0x1.e3a9179dc1a73p-1, 0x1.d276b8adb0b56p-5},
    {0x1.e01e01e01e01ep-1, 0x1.075983598e471p-4}, {0x1.dca01dca01dcap-1, 0x1.253f62f0a1417p-4},
    {0x1.d92f2231e7f8ap-1, 0x1.42edcbea646eep-4}, {0x1.d5cac807572b2p-1, 0x1.60658a93750c4p-4},
    {0x1.d272ca3fc5b1ap-1, 0x1.7da766d7b12d0p-4}, {0x1.cf26e5c44bfc6p-1, 0x1.9ab42462033aep-4},
    {0x1.cbe6d9601cbe7p-1, 0x1.b78c82bb0eda0p-4}, {0x1.c8b265afb8a42p-1, 0x1.d4313d66cb35dp-4},
    {0x1.c5894d10d4986p-1, 0x1.f0a30c01162a4p-4}, {0x1.c26b5392ea01cp-1, 0x1.0671512ca596fp-3},
    {0x1.bf583ee868d8bp-1, 0x1.14785846742acp-3}, {0x1.bc4fd65883e7bp-1, 0x1.2266f190a5acdp-3},
    {0x1.b951e2b18ff23p-1, 0x1.303d718e47fd5p-3}, {0x1.b65e2e3beee05p-1, 0x1.3dfc2b0ecc62ap-3},
    {0x1.b37484ad806cep-1, 0x1.4ba36f39a55e5p-3}, {0x1.b094b31d922a4p-1, 0x1.59338d9982085p-3},
    {0x1.adbe87f94905ep-1, 0x1.66acd4272ad51p-3}, {0x1.aaf1d2f87ebfdp-1, 0x1.740f8f54037a3p-3},
    {0x1.a82e65130e159p-1, 0x1.815c0a14357e9p-3}, {0x1.a574107688a4ap-1, 0x1.8e928de886d41p-3},
    {0x1.a2c2a87c51ca0p-1, 0x1.9bb362e7dfb85p-3}, {0x1.a01a01a01a01ap-1, 0x1.a8becfc882f19p-3},
    {0x1.9d79f176b682dp-1, 0x1.b5b519e8fb5a6p-3}, {0x1.9ae24ea5510dap-1, 0x1.c2968558c18c2p-3},
    {0x1.9852f0d8ec0ffp-1, 0x1.cf6354e09c5ddp-3}, {0x1.95cbb0be377aep-1, 0x1.dc1bca0abec7bp-3},
    {0x1.934c67f9b2ce6p-1, 0x1.e8c0252aa5a60p-3}, {0x1.90d4f120190d5p-1, 0x1.f550a564b7b37p-3},
    {0x1.8e6527af1373fp-1, 0x1.00e6c45ad501dp-2}, {0x1.8bfce8062ff3ap-1, 0x1.071b85fcd590dp-2},
    {0x1.899c0f601899cp-1, 0x1.0d46b579ab74bp-2}, {0x1.87427bcc092b9p-1, 0x1.136870293a8b0p-2},
    {0x1.84f00c2780614p-1, 0x1.1980d2dd4236fp-2}, {0x1.82a4a0182a4a0p-1, 0x1.1f8ff9e48a2f3p-2},
    {0x1.8060180601806p-1, 0x1.2596010df763ap-2}, {0x1.7e225515a4f1dp-1, 0x1.2b9303ab89d25p-2},
    {0x1.7beb3922e017cp-1, 0x1.31871c9544185p-2}, {0x1.79baa6bb6398bp-1, 0x1.3772662bfd85cp-2},
    {0x1.77908119ac60dp-1, 0x1.3d54fa5c1f710p-2}, {0x1.756cac201756dp-1, 0x1.432ef2a04e813p-2},
    {0x1.734f0c541fe8dp-1, 0x1.49006804009d0p-2}, {0x1.713786d9c7c09p-1, 0x1.4ec9732600269p-2},
    {0x1.6f26016f26017p-1, 0x1.548a2c3add263p-2}, {0x1.6d1a62681c861p-1, 0x1.5a42ab0f4cfe2p-2},
    {0x1.6b1490aa31a3dp-1, 0x1.5ff3070a793d4p-2}, {0x1.691473a88d0c0p-1, 0x1.659b57303e1f2p-2},
    {0x1.6719f3601671ap-1, 0x1.6b3bb2235943dp-2}, {0x1.6524f853b4aa3p-1, 0x1.70d42e2789236p-2},
    {0x1.63356b88ac0dep-1, 0x1.7664e1239dbcfp-2}, {0x1.614b36831ae94p-1, 0x1.7bede0a37afbfp-2},
    {0x1.5f66434292dfcp-1, 0x1.816f41da0d495p-2}, {0x1.5d867c3ece2a5p-1, 0x1.86e919a330ba1p-2},
    {0x1.5babcc647fa91p-1, 0x1.8c5b7c858b48bp-2}, {0x1.59d61f123ccaap-1, 0x1.91c67eb45a83ep-2},
    {0x1.5805601580560p-1, 0x1.972a341135159p-2}, {0x1.56397ba7c52e2p-1, 0x1.9c86b02dc0862p-2},
    {0x1.54725e6bb82fep-1, 0x1.a1dc064d5b995p-2}, {0x1.52aff56a8054bp-1, 0x1.a72a4966bd9e9p-2},
    {0x1.50f22e111c4c5p-1, 0x1.ac718c258b0e5p-2}, {0x1.4f38f62dd4c9bp-1, 0x1.b1b1e0ebdfc5ap-2},
    {0x1.4d843bedc2c4cp-1, 0x1.b6eb59d3cf35cp-2}, {0x1.4bd3edda68fe1p-1, 0x1.bc1e08b0dad0ap-2},
    {0x1.4a27fad76014ap-1, 0x1.c149ff115f027p-2}, {0x1.4880522014880p-1, 0x1.c66f4e3ff6ff9p-2},
    {0x1.46dce34596066p-1, 0x1.cb8e0744d7acap-2}, {0x1.453d9e2c776cap-1, 0x1.d0a63ae721e64p-2},
    {0x1.43a2730abee4dp-1, 0x1.d5b7f9ae2c684p-2}, {0x1.420b5265e5951p-1, 0x1.dac353e2c5955p-2},
    {0x1.40782d10e6566p-1, 0x1.dfc859906d5b5p-2}, {0x1.3ee8f42a5af07p-1, 0x1.e4c71a8687704p-2},
    {0x1.3d5d991aa75c6p-1, 0x1.e9bfa659861f5p-2}, {0x1.3bd60d9232955p-1, 0x1.eeb20c640ddf3p-2},
    {0x1.3a524387ac822p-1, 0x1.f39e5bc811e5dp-2}, {0x1.38d22d366088ep-1, 0x1.f884a36fe9ec1p-2},
    {0x1.3755bd1c945eep-1, 0x1.fd64f20f61571p-2}, {0x1.35dce5f9f2af8p-1, 0x1.011fab125ff8ap-1},
    {0x1.34679ace01346p-1, 0x1.0389eefce633cp-1}, {0x1.32f5ced6a1dfap-1, 0x1.05f14bd26459cp-1},
    {0x1.3187758e9ebb6p-1, 0x1.0855c884b450ep-1}, {0x1.301c82ac40260p-1, 0x1.0ab76bece14d2p-1},
    {0x1.2eb4ea1fed14bp-1, 0x1.0d163ccb9d6b8p-1}, {0x1.2d50a012d50a0p-1, 0x1.0f7241c9b497dp-1},
    {0x1.2bef98e5a3711p-1, 0x1.11cb81787ccf8p-1}, {0x1.2a91c92f3c105p-1, 0x1.1422025243d45p-1},
    {0x1.293725bb804a5p-1, 0x1.1675cababa60ep-1}, {0x1.27dfa38a1ce4dp-1, 0x1.18c6e0ff5cf07p-1},
    {0x1.268b37cd60127p-1, 0x1.1b154b57da29ep-1}, {0x1.2539d7e9177b2p-1, 0x1.1d610fe677003p-1},
    {0x1.23eb79717605bp-1, 0x1.1faa34b87094cp-1}, {0x1.22a0122a0122ap-1, 0x1.21f0bfc65beecp-1},
    {0x1.21579804855e6p-1, 0x1.2434b6f483934p-1}, {0x1.2012012012012p-1, 0x1.26762013430e0p-1},
    {0x1.1ecf43c7fb84cp-1, 0x1.28b500df60783p-1}, {0x1.1d8f5672e4abdp-1, 0x1.2af15f02640acp-1},
    {0x1.1c522fc1ce059p-1, 0x1.2d2b4012edc9dp-1}, {0x1.1b17c67f2bae3p-1, 0x1.2f62a99509546p-1},
    {0x1.19e0119e0119ep-1, 0x1.3197a0fa7fe6ap-1}, {0x1.18ab083902bdbp-1, 0x1.33ca2ba328994p-1},
    {0x1.1778a191bd684p-1, 0x1.35fa4edd36ea0p-1}, {0x1.1648d50fc3201p-1, 0x1.38280fe58797fp-1},
    {0x1.151b9a3fdd5c9p-1, 0x1.3a5373e7ebdf9p-1}, {0x1.13f0e8d344724p-1, 0x1.3c7c7fff73206p-1},
    {0x1.12c8b89edc0acp-1, 0x1.3ea33936b2f5bp-1}, {0x1.11a3019a74826p-1, 0x1.40c7a4880dceap-1},
    {0x1.107fbbe011080p-1, 0x1.42e9c6ddf80bfp-1}, {0x1.0f5edfab325a2p-1, 0x1.4509a5133bb0ap-1},
    {0x1.0e40655826011p-1, 0x1.472743f33aaadp-1}, {0x1.0d24456359e3ap-1, 0x1.4942a83a2fc07p-1},
    {0x1.0c0a7868b4171p-1, 0x1.4b5bd6956e273p-1}, {0x1.0af2f722eecb5p-1, 0x1.4d72d3a39fd01p-1},
    {0x1.09ddba6af8360p-1, 0x1.4f87a3f5026e9p-1}, {0x1.08cabb37565e2p-1, 0x1.519a4c0ba3446p-1},
    {0x1.07b9f29b8eae2p-1, 0x1.53aad05b99b7cp-1}, {0x1.06ab59c7912fbp-1, 0x1.55b9354b40bcep-1},
    {0x1.059eea0727586p-1, 0x1.57c57f336f191p-1}, {0x1.04949cc1664c5p-1, 0x1.59cfb25fae87fp-1},
    {0x1.038c6b78247fcp-1, 0x1.5bd7d30e71c73p-1}, {0x1.02864fc7729e9p-1, 0x1.5ddde57149923p-1},
    {0x1.0182436517a37p-1, 0x1.5fe1edad18919p-1}, {0x1.0080402010080p-1, 0x1.61e3efda46467p-1},
};

// zero, subnormal, negative, infinite, NaN: not for the table path
__device__ __forceinline__ bool wx_log_special(double x)
{
    const int ex = (int)(__double_as_longlong(x) >> 52);
    return ex <= 0 || ex >= 0x7ff;
}
// the table path itself, straight-line (callers that batch several arguments test wx_log_special once for all of them)
__device__ __forceinline__ double wx_log_core(double x)
{
    const long long b = __double_as_longlong(x);
    const int ex = (int)(b >> 52);
    const int i = (int)(b >> 45) & 127;
    const double m = __longlong_as_double((b & 0x000fffffffffffffLL) | 0x3ff0000000000000LL);
    const double2 t = __ldg(&wx_log_tab[i]);
    const double r = fma(m, t.x, -1.0);
    double p = fma(r, -1.0 / 6.0, 0.2);
    p = fma(r, p, -0.25);
    p = fma(r, p, 1.0 / 3.0);
    p = fma(r, p, -0.5);
    p = fma(r * r, p, r);                                   // log1p(r)
    const double e = (double)(ex - 1023);
    const double res = fma(e, 0x1.62e42fee00000p-1, (t.y + p) + e * 0x1.a39ef35793c76p-33);
    return x == 1.0 ? 0.0 : res;                              // exactly, like log (the table path would give 1.7e-18)
}
__device__ __noinline__ double wx_log_slow(double x) { return log(x); }
__device__ __forceinline__ double wx_log_fast(double x) { return wx_log_special(x) ? wx_log_slow(x) : wx_log_core(x); }

// ---- double-double accumulation (error-free transformations) ------------------------------------------------
// The LSDB grid (origin, step) of a position is a function of the batch statistics, and every sample is binned on it: a
// 1-ulp change of a sum moves bin edges.  The sums that feed the grid (and the final sum of log pdf) are therefore kept
// as unevaluated pairs hi + lo with |lo| <= ulp(hi)/2 (relative error ~1e-32), so their rounded value does not depend on
// how the batch is cut into slices, CTAs or ranks.
__device__ __forceinline__ void two_sum(double a, double b, double &s, double &e)
{
    s = a + b;
    const double bb = s - a;
    e = (a - (s - bb)) + (b - bb);
}
__device__ __forceinline__ void dd_acc(double &hi, double &lo, double x)
{
    double s, e;
    two_sum(hi, x, s, e);
    hi = s; lo += e;
}
__device__ __forceinline__ void dd_norm(double &hi, double &lo)
{
    const double s = hi + lo;
    lo = lo - (s - hi); hi = s;
}
__device__ __forceinline__ void dd_add(double &hi, double &lo, double bh, double bl)
{
    double s, e;
    two_sum(hi, bh, s, e);
    e += lo + bl;
    hi = s; lo = e;
    dd_norm(hi, lo);
}

// LSDB pass 1: per position, over a slice of the batch: sum(x-c), sum((x-c)^2) as double-double, min, max.
// part (ksplit, 6, szK): s_hi, s_lo, q_hi, q_lo, min, max.  V positions per thread (V > 1: 128-bit loads).
template <typename T, int V>
__global__ void __launch_bounds__(kT) lsdb_stats_part_k(double *part, const T *X, const double *shift, long szK, long N, long kchunk)
{
    constexpr int U = 4;
    using VT = typename std::conditional<V == 1, T, typename std::conditional<sizeof(T) == 8, double2, float4>::type>::type;
    const long e = ((long)blockIdx.x * kT + threadIdx.x) * V;
    if (e >= szK) return;
    const long k0 = (long)blockIdx.y * kchunk;
    long k1 = k0 + kchunk; if (k1 > N) k1 = N;
    double c[V], sh[V], sl[V], qh[V], ql[V], mn[V], mx[V];
#pragma unroll
    for (int v = 0; v < V; ++v) { c[v] = shift[e + v]; sh[v] = sl[v] = qh[v] = ql[v] = 0.0; mn[v] = INFINITY; mx[v] = -INFINITY; }
    auto take = [&](const VT &r) {
        const T *rv = reinterpret_cast<const T *>(&r);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const double a = (double)rv[v] - c[v];
            dd_acc(sh[v], sl[v], a);
            const double p = __dmul_rn(a, a);
            ql[v] += fma(a, a, -p);                        // exact product = p + fma(a,a,-p)
            dd_acc(qh[v], ql[v], p);
            mn[v] = fmin(mn[v], a); mx[v] = fmax(mx[v], a);
        }
    };
    const T *p = X + e;
    long k = k0;
    if (k + U <= k1) {
        // software pipeline: the next U loads are in flight while the current U samples go through the double-double chains
        VT cur[U], nxt[U];
#pragma unroll
        for (int u = 0; u < U; ++u) cur[u] = __ldcs(reinterpret_cast<const VT *>(p + (k + u) * szK));
        for (; k + 2 * U <= k1; k += U) {
#pragma unroll
            for (int u = 0; u < U; ++u) nxt[u] = __ldcs(reinterpret_cast<const VT *>(p + (k + U + u) * szK));
#pragma unroll
            for (int u = 0; u < U; ++u) take(cur[u]);
#pragma unroll
            for (int u = 0; u < U; ++u) cur[u] = nxt[u];
        }
#pragma unroll
        for (int u = 0; u < U; ++u) take(cur[u]);
        k += U;
    }
    for (; k < k1; ++k) take(__ldcs(reinterpret_cast<const VT *>(p + k * szK)));
    double *o = part + ((long)blockIdx.y * 6) * szK + e;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        dd_norm(sh[v], sl[v]); dd_norm(qh[v], ql[v]);
        o[v] = sh[v]; o[szK + v] = sl[v]; o[2 * szK + v] = qh[v]; o[3 * szK + v] = ql[v];
        o[4 * szK + v] = mn[v] + c[v]; o[5 * szK + v] = mx[v] + c[v];
    }
}

// combine the slices in index order (double-double): stats rows 1..6 = s_hi, s_lo, q_hi, q_lo, min, max
__global__ void __launch_bounds__(kT) lsdb_stats_final_k(double *stats, const double *part, long szK, int ksplit)
{
    const long e = (long)blockIdx.x * kT + threadIdx.x;
    if (e >= szK) return;
    double sh = 0, sl = 0, qh = 0, ql = 0, mn = INFINITY, mx = -INFINITY;
    for (int ks = 0; ks < ksplit; ++ks) {
        const double *p = part + ((long)ks * 6) * szK + e;
        dd_add(sh, sl, p[0], p[szK]);
        dd_add(qh, ql, p[2 * szK], p[3 * szK]);
        mn = fmin(mn, p[4 * szK]); mx = fmax(mx, p[5 * szK]);
    }
    stats[szK + e] = sh; stats[2 * szK + e] = sl; stats[3 * szK + e] = qh; stats[4 * szK + e] = ql;
    stats[5 * szK + e] = mn; stats[6 * szK + e] = mx;
}

// parts (nparts, 2, cnt) of double-double numbers -> out (2, cnt), summed in index order
__global__ void __launch_bounds__(kT) dd_sum_parts_k(double *out, const double *part, long cnt, int nparts)
{
    const long e = (long)blockIdx.x * kT + threadIdx.x;
    if (e >= cnt) return;
    double h = 0, l = 0;
    for (int q = 0; q < nparts; ++q) dd_add(h, l, part[((long)q * 2) * cnt + e], part[((long)q * 2 + 1) * cnt + e]);
    out[e] = h; out[cnt + e] = l;
}

static int pick_ksplit(long szK, long N, int sms)
{
    long blocks_x = (szK + kT - 1) / kT;
    long want = (long)sms * 8;                       // ~8 resident CTAs of 256 threads per SM
    long ks = (want + blocks_x - 1) / blocks_x;
    if (ks < 1) ks = 1;
    if (ks > 1024) ks = 1024;
    if (ks > (N + 63) / 64) ks = (N + 63) / 64;      // at least 64 signals per slice
    if (ks < 1) ks = 1;
    return (int)ks;
}

template <typename T>
int moments(double *o0, double *o1, const T *X, long szK, long N, cudaStream_t s)
{
    WX_REQUIRE(szK >= 1 && N >= 0, "bad sizes");
    WX_REQUIRE(o0 && o1 && (N == 0 || X), "null pointer");
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    constexpr long V = 16 / (long)sizeof(T);
    const bool vec = szK % V == 0 && (((uintptr_t)X) & 15) == 0;
    const int ksplit = N > 0 ? pick_ksplit(vec ? szK / V : szK, N, dv.sms) : 1;
    const long kchunk = N > 0 ? (N + ksplit - 1) / ksplit : 1;
    double *part; rc = wx_scratch(&part, (size_t)ksplit * 2 * szK, s); if (rc) return rc;
    if (vec) {
        dim3 grid((unsigned)((szK / V + kT - 1) / kT), (unsigned)ksplit);
        moments_part_vec_k<T><<<grid, kT, 0, s>>>(part, X, szK, N, kchunk);
    } else {
        dim3 grid((unsigned)((szK + kT - 1) / kT), (unsigned)ksplit);
        moments_part_k<T><<<grid, kT, 0, s>>>(part, X, szK, N, kchunk);
    }
    WX_LAUNCHED();
    moments_final_k<<<gridf(szK), kT, 0, s>>>(o0, o1, part, szK, ksplit);
    WX_LAUNCHED();
    return wx_scratch_free(part, s);
}

// ---- node geometry ---------------------------------------------------------------------------------------
struct NodeGeom { long m, n; int K; int redundant; };   // m == 0: 1-D

__device__ __forceinline__ int ilog2d(long i) { return 63 - __clzll((unsigned long long)i); }
__device__ __forceinline__ int quaddepthd(long i) { int d = 0; long last = 1, w = 1; while (i > last) { w *= 4; last += w; ++d; } return d; }

// element t of node q (0-based node, level/heap order) -> linear index into (sz, K); returns node size through cnt
__device__ __forceinline__ long node_elem(const NodeGeom &g, long q, long t, long &cnt, double &scale)
{
    const long i = q + 1;
    if (g.m == 0) {
        const int d = ilog2d(i);
        if (g.redundant) { cnt = g.n; scale = 1.0 / (double)(1L << d); return q * g.n + t; }
        const long n0 = g.n >> d, j = i - (1L << d);
        cnt = n0; scale = 1.0;
        return (long)d * g.n + j * n0 + t;
    }
    const long img = g.m * g.n;
    const int d = quaddepthd(i);
    if (g.redundant) { cnt = img; scale = 1.0 / (double)(1L << (2 * d)); return q * img + t; }
    // walk from the root: child c of parent p is 4p-2+c, c = 2*rowbit + colbit
    long first = 1, w = 1;
    for (int k = 0; k < d; ++k) { first += w; w *= 4; }      // first index of depth d
    long off = i - first;                                     // position within the depth, base-4 digits = path
    long r0 = 0, c0 = 0, nr = g.m, nc = g.n;
    for (int k = d - 1; k >= 0; --k) {
        const int dig = (int)((off >> (2 * k)) & 3);
        nr >>= 1; nc >>= 1;
        if (dig & 2) r0 += nr;
        if (dig & 1) c0 += nc;
    }
    cnt = nr * nc; scale = 1.0;
    const long r = t % nr, c = t / nr;
    return (long)d * img + (c0 + c) * g.m + r0 + r;
}

// per-position JBB term from the (all-reduced) moments: log(sigma) or sigma^p ; flag bad variances
__global__ void __launch_bounds__(kT) jbb_term_k(double *term, const double *sum, const double *sumsq, long szK, double Ntot, int kind, double p,
                                                 int elt, int *bad)
{
    const long e = (long)blockIdx.x * kT + threadIdx.x;
    if (e >= szK) return;
    double ex = sum[e] / Ntot, ex2 = sumsq[e] / Ntot;
    if (elt == 4) { ex = (double)(float)ex; ex2 = (double)(float)ex2; }
    double var = ex2 - ex * ex;
    if (elt == 4) var = (double)(float)var;
    if (!(var >= 0.0)) atomicExch(bad, 1);
    double sg = sqrt(var);
    if (elt == 4) sg = (double)(float)sg;
    term[e] = (kind == 0) ? log(fabs(sg)) : pow(fabs(sg), p);
}

// one warp per node, lane-strided partial sums + xor-shuffle tree (fixed order)
__global__ void __launch_bounds__(kT) node_reduce_k(double *costs, const double *term, NodeGeom g, long nnodes, double mult)
{
    const long q = ((long)blockIdx.x * kT + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= nnodes) return;
    long cnt; double scale;
    node_elem(g, q, 0, cnt, scale);
    double acc = 0.0;
    for (long t = lane; t < cnt; t += 32) {
        long c2; double s2;
        acc += term[node_elem(g, q, t, c2, s2)];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) costs[q] = mult * acc * scale;
}

static long count_nodes(long m, int K, int redundant)
{
    if (redundant) return K;
    return m > 0 ? ((1L << (2 * K)) - 1) / 3 : (1L << K) - 1;
}

static int node_costs_to_host(double *costs_host, const double *term, long m, long n, int K, int redundant, double mult, int elt, cudaStream_t s)
{
    const long nn = count_nodes(m, K, redundant);
    double *dc; int rc = wx_scratch(&dc, (size_t)nn, s); if (rc) return rc;
    NodeGeom g{m, n, K, redundant};
    node_reduce_k<<<gridf(nn * 32), kT, 0, s>>>(dc, term, g, nn, mult);
    WX_LAUNCHED();
    WX_CUDA(cudaMemcpyAsync(costs_host, dc, (size_t)nn * sizeof(double), cudaMemcpyDeviceToHost, s));
    WX_CUDA(cudaStreamSynchronize(s));
    if (elt == 4) for (long i = 0; i < nn; ++i) costs_host[i] = (double)(float)costs_host[i];
    return wx_scratch_free(dc, s);
}

// ---- LSDB -----------------------------------------------------------------------------------------------
struct LsdbGrid { long nbins, mbins, npts; };
static LsdbGrid lsdb_grid(long N)
{
    LsdbGrid g;
    g.nbins = (long)std::ceil(std::pow(30.0 * (double)N, 0.2));     // bestbasis_costs.jl:140
    g.mbins = (long)std::ceil(50.0 / (double)g.nbins);              // :141
    g.npts = (g.nbins + 1) * g.mbins;                               // length of rng (:147)
    return g;
}

// per-position grid origin a and step delta from the reduced statistics  (bestbasis_costs.jl:143-147)
// stats rows: 0 shift c, 1/2 sum(x-c) hi/lo, 3/4 sum((x-c)^2) hi/lo, 5 min, 6 max   (hi = the correctly rounded sum)
template <bool SAFE = true>
__device__ __forceinline__ void lsdb_axis(const double *stats, long szK, long e, double Ntot, long npts, double &a, double &delta)
{
    const double s1 = stats[szK + e], s2 = stats[3 * szK + e], mn = stats[5 * szK + e], mx = stats[6 * szK + e];
    double var = (s2 - s1 * s1 / Ntot) / (Ntot - 1.0);               // Statistics.std (corrected)
    if (var < 0) var = 0;
    const double sg = sqrt(var);
    delta = (mx - mn + sg) / (double)(npts - 1);
    a = mn - 0.5 * sg;
    // A position that is constant over the batch has delta == 0 (the reference's range `a:0.0:b` throws, bestbasis_costs.jl:146).
    // The streaming kernels must still terminate on it (0 * inf = NaN converts to LONG_MIN on the device and the grid walk of the
    // log-pdf pass would never end), so they see a unit step there; lsdb_density_k tests the RAW step, poisons the column's density
    // with NaN, the cost of every node that contains the position becomes NaN and wx_lsdb_costs turns that into the error.
    if (SAFE && (!(delta > 0.0) || !isfinite(delta))) delta = 1.0;
}

template <typename T>
__global__ void __launch_bounds__(kT) lsdb_hist_k(double *counts, const double *stats, const T *X, long szK, long N, long kchunk, double Ntot, long npts)
{
    const long e = (long)blockIdx.x * kT + threadIdx.x;
    if (e >= szK) return;
    double a, delta;
    lsdb_axis(stats, szK, e, Ntot, npts, a, delta);
    const double dinv = 1.0 / delta;
    const long k0 = (long)blockIdx.y * kchunk;
    long k1 = k0 + kchunk; if (k1 > N) k1 = N;
    for (long k = k0; k < k1; ++k) {
        const double xv = (double)X[k * szK + e];
        const long ki = (long)floor((xv - a) * dinv + 1.5);          // AverageShiftedHistograms bin rule (1-based)
        if (ki >= 1 && ki <= npts) atomicAdd(&counts[(ki - 1) * szK + e], 1.0);
    }
}

// shared-memory variant: a thread owns one position and keeps its npts bin counters in shared memory (column tid of a
// (npts, kT) table: conflict free, no atomics needed), U signals in flight; the per-CTA counts are then added to the global
// table with one atomicAdd per non-empty bin (integer-valued doubles: exact and order independent).
// CT: counter type.  16-bit counters (a slice of at most 65535 signals per CTA) halve the shared memory, so six CTAs instead of three are
// resident: the pass is latency bound on its loads (ncu: long_scoreboard 8.2 stall cycles per issue at 24 warps).
template <typename T, typename CT>
__global__ void __launch_bounds__(kT) lsdb_hist_smem_k(double *counts, const double *stats, const T *X, long szK, long N, long kchunk, double Ntot,
                                                        int npts)
{
    extern __shared__ unsigned int wx_cnt_raw[];
    CT *wx_cnt = reinterpret_cast<CT *>(wx_cnt_raw);
    constexpr int U = 16;
    const int tid = threadIdx.x;
    const long e = (long)blockIdx.x * kT + tid;
    for (int i = 0; i < npts; ++i) wx_cnt[i * kT + tid] = 0;
    if (e >= szK) return;                                  // a thread only ever touches its own column: no barrier needed
    double a, delta;
    lsdb_axis(stats, szK, e, Ntot, npts, a, delta);
    const double dinv = 1.0 / delta;
    const long k0 = (long)blockIdx.y * kchunk;
    long k1 = k0 + kchunk; if (k1 > N) k1 = N;
    const T *p = X + e;
    long k = k0;
    for (; k + U <= k1; k += U) {
        T r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = __ldcs(p + (k + u) * szK);
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long ki = (long)floor(((double)r[u] - a) * dinv + 1.5);      // AverageShiftedHistograms bin rule (1-based)
            if (ki >= 1 && ki <= npts) wx_cnt[(int)(ki - 1) * kT + tid] += (CT)1;
        }
    }
    for (; k < k1; ++k) {
        const long ki = (long)floor(((double)p[k * szK] - a) * dinv + 1.5);
        if (ki >= 1 && ki <= npts) wx_cnt[(int)(ki - 1) * kT + tid] += (CT)1;
    }
    for (int i = 0; i < npts; ++i) {
        const unsigned c = (unsigned)wx_cnt[i * kT + tid];
        if (c) atomicAdd(&counts[(long)i * szK + e], (double)c);
    }
}

// ASH density on the grid from the (all-reduced) counts, triangular kernel, normalised by 1/(sum(y)*delta)
__global__ void __launch_bounds__(kT) lsdb_density_k(double *dens, const double *counts, const double *stats, long szK, double Ntot, long npts, long mb)
{
    const long e = (long)blockIdx.x * kT + threadIdx.x;
    if (e >= szK) return;
    double a, delta;
    lsdb_axis<false>(stats, szK, e, Ntot, npts, a, delta);
    double tot = 0.0;
    for (long i = 1; i <= npts; ++i) {
        double y = 0.0;
        long lo = i - mb + 1 < 1 ? 1 : i - mb + 1, hi = i + mb - 1 > npts ? npts : i + mb - 1;
        for (long k = lo; k <= hi; ++k) {
            const double u = fabs((double)(i - k) / (double)mb);
            y += counts[(k - 1) * szK + e] * (1.0 - u);
        }
        dens[(i - 1) * szK + e] = y;
        tot += y;
    }
    // a position that is constant over the batch has delta == 0: the reference's range `a:0.0:b` throws (bestbasis_costs.jl:146).
    // Poison the column with NaN; wx_lsdb_costs turns a NaN cost into that error.
    const double den = (delta > 0.0 && isfinite(delta)) ? 1.0 / (tot * delta) : NAN;
    for (long i = 0; i < npts; ++i) dens[i * szK + e] = (den == den) ? dens[i * szK + e] * den : NAN;
}

template <typename T>
__global__ void __launch_bounds__(kT) lsdb_logpdf_part_k(double *part, const double *dens, const double *stats, const T *X, long szK, long N, long kchunk,
                                                          double Ntot, long npts)
{
    const long e = (long)blockIdx.x * kT + threadIdx.x;
    if (e >= szK) return;
    double a, delta;
    lsdb_axis(stats, szK, e, Ntot, npts, a, delta);
    const double dinv = 1.0 / delta;
    const long k0 = (long)blockIdx.y * kchunk;
    long k1 = k0 + kchunk; if (k1 > N) k1 = N;
    double acc = 0.0, accl = 0.0;
    for (long k = k0; k < k1; ++k) {
        const double xv = (double)X[k * szK + e];
        long i = (long)floor((xv - a) * dinv) + 1;                   // searchsortedlast(rng, x), 1-based
        while (i >= 1 && i <= npts && a + (double)(i - 1) * delta > xv) --i;
        while (i + 1 <= npts && a + (double)i * delta <= xv) ++i;
        double pdf = 0.0;
        if (i >= 1 && i < npts) {
            const double g0 = a + (double)(i - 1) * delta, g1 = a + (double)i * delta;
            const double y0 = dens[(i - 1) * szK + e], y1 = dens[i * szK + e];
            pdf = y0 + (y1 - y0) * (xv - g0) / (g1 - g0);
        }
        dd_acc(acc, accl, log(pdf));
    }
    dd_norm(acc, accl);
    double *o = part + ((long)blockIdx.y * 2) * szK + e;
    o[0] = acc; o[szK] = accl;
}

// shared-memory variant: the density columns of the CTA's kL positions are staged in shared memory ((npts, kL) table);
// kH threads share a position (each takes a slice of the CTA's signals).  The interpolation weight uses 1/delta instead of
// 1/(g1-g0) and the logarithm is taken of the product of U consecutive pdf values (log of a product = sum of logs;
// a zero pdf still gives -Inf, and U pdf values of order 1e-3..1e3 cannot over/underflow a double).
template <typename T, int kL, int kH, int U>
__global__ void __launch_bounds__(kL * kH) lsdb_logpdf_smem_k(double *part, const double *dens, const double *stats, const T *X, long szK, long N,
                                                               long kchunk, double Ntot, int npts)
{
    extern __shared__ double wx_dens[];
    const int pos = threadIdx.x % kL, kh = threadIdx.x / kL;
    const long e = (long)blockIdx.x * kL + pos;
    const bool live = e < szK;
    if (live) for (int i = kh; i < npts; i += kH) wx_dens[i * kL + pos] = dens[(long)i * szK + e];
    __syncthreads();
    if (!live) return;
    double a, delta;
    lsdb_axis(stats, szK, e, Ntot, npts, a, delta);
    const double dinv = 1.0 / delta;
    const long c0 = (long)blockIdx.y * kchunk;
    long c1 = c0 + kchunk; if (c1 > N) c1 = N;
    const long sub = (c1 - c0 + kH - 1) / kH;
    const long k0 = c0 + kh * sub;
    long k1 = k0 + sub; if (k1 > c1) k1 = c1;
    const T *p = X + e;
    auto pdf_at = [&](double xv) {
        const double t = (xv - a) * dinv;
        long i = (long)floor(t) + 1;                                   // searchsortedlast(rng, x), 1-based
        double g0 = fma((double)(i - 1), delta, a);
        while (i >= 1 && i <= npts && g0 > xv) { --i; g0 = fma((double)(i - 1), delta, a); }
        while (i + 1 <= npts && fma((double)i, delta, a) <= xv) { ++i; g0 = fma((double)(i - 1), delta, a); }
        double pdf = 0.0;
        if (i >= 1 && i < npts) {
            const double y0 = wx_dens[(int)(i - 1) * kL + pos], y1 = wx_dens[(int)i * kL + pos];
            pdf = fma((y1 - y0) * dinv, xv - g0, y0);
        }
        return pdf;
    };
    double acc = 0.0, accl = 0.0;
    long k = k0;
    for (; k + U <= k1; k += U) {
        T r[U];
#pragma unroll
        for (int u = 0; u < U; ++u) r[u] = __ldcs(p + (k + u) * szK);
        double prod = 1.0;
#pragma unroll
        for (int u = 0; u < U; ++u) prod *= pdf_at((double)r[u]);
        dd_acc(acc, accl, wx_log_fast(prod));
    }
    for (; k < k1; ++k) dd_acc(acc, accl, wx_log_fast(pdf_at((double)p[k * szK])));
    dd_norm(acc, accl);
    double *o = part + ((long)(blockIdx.y * kH + kh) * 2) * szK + e;
    o[0] = acc; o[szK] = accl;
}

template <typename T>
int lsdb_pass1(double *stats, const T *X, long szK, long Nlocal, cudaStream_t s)
{
    WX_REQUIRE(stats && szK >= 1 && Nlocal >= 0 && (Nlocal == 0 || X), "bad arguments");
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    constexpr long V = 16 / (long)sizeof(T);
    const bool vec = szK % V == 0 && (((uintptr_t)X) & 15) == 0;
    const int ksplit = Nlocal > 0 ? pick_ksplit(vec ? szK / V : szK, Nlocal, dv.sms) : 1;
    const long kchunk = Nlocal > 0 ? (Nlocal + ksplit - 1) / ksplit : 1;
    double *part; rc = wx_scratch(&part, (size_t)ksplit * 6 * szK, s); if (rc) return rc;
    if (vec) {
        dim3 grid((unsigned)((szK / V + kT - 1) / kT), (unsigned)ksplit);
        lsdb_stats_part_k<T, (int)V><<<grid, kT, 0, s>>>(part, X, stats, szK, Nlocal, kchunk);
    } else {
        dim3 grid((unsigned)((szK + kT - 1) / kT), (unsigned)ksplit);
        lsdb_stats_part_k<T, 1><<<grid, kT, 0, s>>>(part, X, stats, szK, Nlocal, kchunk);
    }
    WX_LAUNCHED();
    lsdb_stats_final_k<<<gridf(szK), kT, 0, s>>>(stats, part, szK, ksplit);
    WX_LAUNCHED();
    return wx_scratch_free(part, s);
}

template <typename T>
int lsdb_pass2(double *counts, const double *stats, const T *X, long szK, long Nlocal, long Ntotal, cudaStream_t s)
{
    WX_REQUIRE(szK >= 1 && Nlocal >= 0 && Ntotal >= 2, "LSDB needs at least two signals");
    WX_REQUIRE(counts && stats && (Nlocal == 0 || X), "null pointer");
    const LsdbGrid g = lsdb_grid(Ntotal);
    WX_CUDA(cudaMemsetAsync(counts, 0, (size_t)g.npts * szK * sizeof(double), s));
    if (Nlocal == 0) return WX_OK;
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    int ksplit = pick_ksplit(szK, Nlocal, dv.sms);
    long kchunk = (Nlocal + ksplit - 1) / ksplit;
    dim3 grid((unsigned)((szK + kT - 1) / kT), (unsigned)ksplit);
    const size_t smem = (size_t)g.npts * kT * sizeof(unsigned int), smem16 = (size_t)g.npts * kT * sizeof(unsigned short);
    static const bool no16 = getenv("WX_B200_LSDB_HIST32") != nullptr;           // A-B measurements
    if (!no16 && smem16 <= dv.smem_optin) {
        auto kern = lsdb_hist_smem_k<T, unsigned short>;
        WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem16));
        int occ = 0;
        WX_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, kT, smem16));
        // slices: two full waves of resident CTAs, at most 65535 signals (the counter range) and at least 64 signals each
        long ks = occ > 0 ? (2L * dv.sms * occ) / grid.x : ksplit;
        if (ks < 1) ks = 1;
        if (ks > (Nlocal + 63) / 64) ks = (Nlocal + 63) / 64;
        if (ks < (Nlocal + 65534) / 65535) ks = (Nlocal + 65534) / 65535;
        if (ks <= 65535) {
            ksplit = (int)ks; kchunk = (Nlocal + ksplit - 1) / ksplit; grid.y = (unsigned)ksplit;
            kern<<<grid, kT, smem16, s>>>(counts, stats, X, szK, Nlocal, kchunk, (double)Ntotal, (int)g.npts);
            WX_LAUNCHED();
            return WX_OK;
        }
    }
    if (smem <= dv.smem_optin && kchunk < (1L << 32)) {
        auto kern = lsdb_hist_smem_k<T, unsigned int>;
        WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kT, smem, s>>>(counts, stats, X, szK, Nlocal, kchunk, (double)Ntotal, (int)g.npts);
    } else {
        lsdb_hist_k<T><<<grid, kT, 0, s>>>(counts, stats, X, szK, Nlocal, kchunk, (double)Ntotal, g.npts);
    }
    WX_LAUNCHED();
    return WX_OK;
}

template <typename T>
int lsdb_pass3(double *logsum, const double *counts, const double *stats, const T *X, long szK, long Nlocal, long Ntotal, cudaStream_t s)
{
    WX_REQUIRE(szK >= 1 && Nlocal >= 0 && Ntotal >= 2, "LSDB needs at least two signals");
    WX_REQUIRE(logsum && counts && stats && (Nlocal == 0 || X), "null pointer");
    const LsdbGrid g = lsdb_grid(Ntotal);
    if (Nlocal == 0) { WX_CUDA(cudaMemsetAsync(logsum, 0, (size_t)2 * szK * sizeof(double), s)); return WX_OK; }
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    const int ksplit = pick_ksplit(szK, Nlocal, dv.sms);
    const long kchunk = (Nlocal + ksplit - 1) / ksplit;
    static const char *env = getenv("WX_B200_LSDB_P3");       // measurement knob: (positions per CTA, threads per position, logs fused)
    const int variant = env ? atoi(env) : 3;
    const int kL = variant >= 6 ? 32 : (variant >= 2 ? 64 : 128), kH = variant >= 6 ? 8 : (variant >= 2 ? 4 : 2);
    double *dens, *part;
    rc = wx_scratch(&dens, (size_t)g.npts * szK, s); if (rc) return rc;
    rc = wx_scratch(&part, (size_t)ksplit * kH * 2 * szK, s); if (rc) return rc;
    lsdb_density_k<<<gridf(szK), kT, 0, s>>>(dens, counts, stats, szK, (double)Ntotal, g.npts, g.mbins);
    WX_LAUNCHED();
    const size_t smem = (size_t)g.npts * kL * sizeof(double);
    int nparts = ksplit;
    if (smem <= dv.smem_optin) {
        dim3 grid((unsigned)((szK + kL - 1) / kL), (unsigned)ksplit);
#define WX_P3(LL, HH, UU) { auto kern = lsdb_logpdf_smem_k<T, LL, HH, UU>; \
        WX_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        kern<<<grid, LL * HH, smem, s>>>(part, dens, stats, X, szK, Nlocal, kchunk, (double)Ntotal, (int)g.npts); }
        switch (variant) {
            case 0: WX_P3(128, 2, 4) break; case 1: WX_P3(128, 2, 8) break; case 2: WX_P3(64, 4, 4) break; default: WX_P3(64, 4, 8) break;
            case 4: WX_P3(64, 4, 16) break; case 5: WX_P3(64, 4, 12) break; case 6: WX_P3(32, 8, 8) break; case 7: WX_P3(32, 8, 16) break;
        }
#undef WX_P3
        nparts = ksplit * kH;
    } else {
        dim3 grid((unsigned)((szK + kT - 1) / kT), (unsigned)ksplit);
        lsdb_logpdf_part_k<T><<<grid, kT, 0, s>>>(part, dens, stats, X, szK, Nlocal, kchunk, (double)Ntotal, g.npts);
    }
    WX_LAUNCHED();
    dd_sum_parts_k<<<gridf(szK), kT, 0, s>>>(logsum, part, szK, nparts);
    WX_LAUNCHED();
    int rc2 = wx_scratch_free(dens, s), rc3 = wx_scratch_free(part, s);
    return rc2 ? rc2 : rc3;
}

__global__ void __launch_bounds__(kT) scale_k(double *out, const double *in, long n, double f)
{
    const long e = (long)blockIdx.x * kT + threadIdx.x;
    if (e < n) out[e] = in[e] * f;
}

// ---- LDB time-frequency energy map (ldb/ldb_energymap.jl:109-141): per class, per position, sum of squares over the class ----
// part (ksplit, NC, szK).  labels[k] in [0, nc) on the device; this launch accumulates classes c0 .. c0+NC-1.
template <typename T, int NC>
__global__ void __launch_bounds__(kT) energy_tf_part_k(double *part, const T *X, const int *labels, int c0, long szK, long N, long kchunk)
{
    constexpr int V = 16 / (int)sizeof(T), U = 4;
    using VT = typename std::conditional<sizeof(T) == 8, double2, float4>::type;
    const long e = ((long)blockIdx.x * kT + threadIdx.x) * V;
    if (e >= szK) return;
    const long k0 = (long)blockIdx.y * kchunk;
    long k1 = k0 + kchunk; if (k1 > N) k1 = N;
    double q[NC][V];
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int v = 0; v < V; ++v) q[c][v] = 0.0;
    const T *p = X + e;
    long k = k0;
    auto take = [&](const VT &r, int lab) {
        const T *rv = reinterpret_cast<const T *>(&r);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const double a = (double)rv[v], a2 = a * a;
#pragma unroll
            for (int c = 0; c < NC; ++c) q[c][v] += (lab == c0 + c) ? a2 : 0.0;
        }
    };
    for (; k + U <= k1; k += U) {
        VT r[U]; int lab[U];
#pragma unroll
        for (int u = 0; u < U; ++u) { r[u] = __ldcs(reinterpret_cast<const VT *>(p + (k + u) * szK)); lab[u] = labels[k + u]; }
#pragma unroll
        for (int u = 0; u < U; ++u) take(r[u], lab[u]);
    }
    for (; k < k1; ++k) take(__ldcs(reinterpret_cast<const VT *>(p + k * szK)), labels[k]);
    double *o = part + ((long)blockIdx.y * NC) * szK + e;
#pragma unroll
    for (int c = 0; c < NC; ++c)
#pragma unroll
        for (int v = 0; v < V; ++v) o[(long)c * szK + v] = q[c][v];
}

// scalar variant (any szK / alignment)
template <typename T, int NC>
__global__ void __launch_bounds__(kT) energy_tf_part_scalar_k(double *part, const T *X, const int *labels, int c0, long szK, long N, long kchunk)
{
    const long e = (long)blockIdx.x * kT + threadIdx.x;
    if (e >= szK) return;
    const long k0 = (long)blockIdx.y * kchunk;
    long k1 = k0 + kchunk; if (k1 > N) k1 = N;
    double q[NC];
#pragma unroll
    for (int c = 0; c < NC; ++c) q[c] = 0.0;
    for (long k = k0; k < k1; ++k) {
        const double a = (double)X[k * szK + e], a2 = a * a;
        const int lab = labels[k];
#pragma unroll
        for (int c = 0; c < NC; ++c) q[c] += (lab == c0 + c) ? a2 : 0.0;
    }
    for (int c = 0; c < NC; ++c) part[((long)blockIdx.y * NC + c) * szK + e] = q[c];
}

template <int NC>
__global__ void __launch_bounds__(kT) energy_tf_final_k(double *esum, const double *part, long szK, int ksplit, int nvalid)
{
    const long e = (long)blockIdx.x * kT + threadIdx.x;
    if (e >= szK) return;
    for (int c = 0; c < nvalid; ++c) {
        double s = 0.0;
        for (int ks = 0; ks < ksplit; ++ks) s += part[((long)ks * NC + c) * szK + e];
        esum[(long)c * szK + e] = s;
    }
}

// discriminant_measure(Gamma, dm) ldb/ldb_measures.jl:139-183, 302-325: D[e] = sum over class pairs i<j of dm(p_i, p_j) with
// p_c = esum[c][e] * inv_norm[c].  kind 0 asymmetric relative entropy, 1 symmetric, 2 Lp distance (p - q)^pw, 3 Hellinger.
__global__ void __launch_bounds__(kT) ldb_dm_k(double *D, const double *esum, const double *inv_norm, int nc, long szK, int kind, double pw, int elt)
{
    const long e = (long)blockIdx.x * kT + threadIdx.x;
    if (e >= szK) return;
    double acc = 0.0;
    for (int i = 0; i < nc; ++i) {
        double p = esum[(long)i * szK + e] * inv_norm[i];
        if (elt == 4) p = (double)(float)p;
        for (int j = i + 1; j < nc; ++j) {
            double q = esum[(long)j * szK + e] * inv_norm[j];
            if (elt == 4) q = (double)(float)q;
            double v;
            if (kind == 0) v = (p == 0.0 || q == 0.0) ? 0.0 : p * log(p / q);
            else if (kind == 1) v = (p == 0.0 || q == 0.0) ? 0.0 : p * log(p / q) + q * log(q / p);
            else if (kind == 2) v = pow(p - q, pw);
            else { const double t = sqrt(p) - sqrt(q); v = t * t; }
            acc += v;
        }
    }
    D[e] = acc;
}

// ---- BB: per-signal best basis (bestbasis/bestbasis_tree.jl:210-256, bestbasis/bestbasis_costs.jl:103-125) ----------------
// coefcost(x, ::ShannonEntropyCost | ::LogEnergyEntropyCost, nrm): s = (x/nrm)^2 ; -s log s | -log s ; 0 when s == 0
__device__ __forceinline__ double bb_term(double x, double inv_nrm, int kind)
{
    const double q = x * inv_nrm, s = q * q;
    if (s == 0.0) return 0.0;
    const double ls = wx_log_fast(s);
    return kind == 0 ? -s * ls : -ls;
}

// four terms at once: one test for "any argument off the table path" (s = 0 included), then straight-line code -- the per-term
// branches cost more issue slots than the arithmetic (ncu: 78 thread instructions per coefficient, 20 % of them FP64)
template <int KIND>
__device__ __forceinline__ void bb_terms4(const double *v, double inv_nrm, double *t)
{
    double s[4];
    bool sp = false;
#pragma unroll
    for (int i = 0; i < 4; ++i) { const double q = v[i] * inv_nrm; s[i] = q * q; sp |= wx_log_special(s[i]); }
    if (!sp) {
#pragma unroll
        for (int i = 0; i < 4; ++i) { const double ls = wx_log_core(s[i]); t[i] = KIND == 0 ? -s[i] * ls : -ls; }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) t[i] = bb_term(v[i], inv_nrm, KIND);
    }
}

// sum over the G = blockDim-aligned power-of-two group of lanes that share a node (G <= 32: segmented xor shuffles;
// G > 32: whole warps, combined through shared memory by the caller)
__device__ __forceinline__ double group_sum(double v, int G)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) if (o < G) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 1-D tables, one CTA per signal.  Level l of a non-redundant table (n, K) has 2^l nodes of n >> l coefficients; a redundant
// table (n, K nodes) has one node of n coefficients per column, cost divided by 2^depth (:218-222).  Threads split each node
// in G = max(1, 256 >> l) strided parts (coalesced loads), partial sums meet through shuffles / shared memory.
template <typename T>
__global__ void __launch_bounds__(kT) bb_costs_1d_k(double *costs, const T *X, long n, int K, long nn, int redundant, int kind)
{
    __shared__ double wsum[kT / 32];
    __shared__ double s_inv;
    const long k = blockIdx.x;
    const T *Xk = X + k * n * K;
    double *ck = costs + k * nn;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // nrm = norm(X[:,1])
    double a = 0.0;
    for (long e = tid; e < n; e += kT) { const double v = (double)Xk[e]; a = fma(v, v, a); }
    a = group_sum(a, 32);
    if (lane == 0) wsum[warp] = a;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < kT / 32; ++w) t += wsum[w];
        double nrm = sqrt(t);
        if (sizeof(T) == 4) nrm = (double)(float)nrm;
        s_inv = nrm > 0.0 ? 1.0 / nrm : 0.0;                 // nrm == 0: every cost is 0 (bestbasis_costs.jl:119)
    }
    __syncthreads();
    const double inv = s_inv;
    for (int l = 0; l < K; ++l) {
        const long nodes = redundant ? 1 : (1L << l), p = redundant ? n : (n >> l);
        const T *lev = Xk + (long)l * n;
        int lg = 0; while ((1L << lg) < nodes) ++lg;
        const int G = (lg >= 8) ? 1 : (kT >> lg);             // threads per node
        const double scale = redundant ? 1.0 / (double)(1L << ilog2d(l + 1)) : 1.0;
        for (long j0 = 0; j0 < nodes; j0 += kT / G) {          // kT / G nodes per sweep
            const long j = j0 + tid / G;
            const int sub = tid % G;
            double acc = 0.0;
            if (j < nodes) {
                const T *nd = lev + j * p;
                for (long e = sub; e < p; e += G) acc += bb_term((double)nd[e], inv, kind);
            }
            if (G <= 32) {
                acc = group_sum(acc, G);
                if (sub == 0 && j < nodes) ck[redundant ? l : ((1L << l) - 1 + j)] = acc * scale;
            } else {
                acc = group_sum(acc, 32);
                __syncthreads();
                if (lane == 0) wsum[warp] = acc;
                __syncthreads();
                if (sub == 0 && j < nodes) {
                    double t = 0.0;
                    for (int w = 0; w < G / 32; ++w) t += wsum[warp + w];
                    ck[redundant ? l : ((1L << l) - 1 + j)] = t * scale;
                }
            }
        }
    }
}

// The same for the common shape -- decimated table, n a power of two >= 1024 -- with the control flow stripped: the generic kernel
// above spends 112 thread instructions per coefficient (ncu: 78 % of the issue slots, 7 % of them DFMA; 64-bit index arithmetic and
// per-level set-up around four coefficients of work).  Here a thread owns four consecutive coefficients of a 1024-element chunk
// (one 128-bit / two 128-bit loads), all four lie in one node whenever the node has >= 4 coefficients, and the node sums are
// segmented shuffle reductions over the p/4 threads of a node (shared memory only for nodes wider than a warp's 128 coefficients).
template <typename T> struct BbLoad4;
template <> struct BbLoad4<double> {
    __device__ static __forceinline__ void ld(const double *p, double *v)
    {
        const double2 a = __ldcs(reinterpret_cast<const double2 *>(p)), b = __ldcs(reinterpret_cast<const double2 *>(p) + 1);
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    }
};
template <> struct BbLoad4<float> {
    __device__ static __forceinline__ void ld(const float *p, double *v)
    {
        const float4 a = __ldcs(reinterpret_cast<const float4 *>(p));
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    }
};

// one level whose nodes have 2^LGP < 1024 coefficients: a thread's four coefficients share a node when LGP >= 2 and the node sum is
// a segmented shuffle reduction over the 2^(LGP-2) threads of the node (LGP = 8, 9: whole warps, combined through shared memory)
template <typename T, int KIND, int LGP>
__device__ __forceinline__ void bb_level_small(const T *__restrict__ lev, double *__restrict__ cl, int nchunk, int tid, double inv, double *wsum)
{
    const int lane = tid & 31, warp = tid >> 5;
    double v[4], t4[4];
    for (int c = 0; c < nchunk; ++c) {
        const int e0 = c * 1024 + tid * 4;
        BbLoad4<T>::ld(lev + e0, v);
        bb_terms4<KIND>(v, inv, t4);
        if constexpr (LGP >= 2) {
            constexpr int gsz = 1 << (LGP - 2);               // threads per node: 1 .. 128
            double acc = (t4[0] + t4[1]) + (t4[2] + t4[3]);
            if constexpr (gsz <= 32) {
#pragma unroll
                for (int o = gsz / 2; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if ((tid & (gsz - 1)) == 0) cl[e0 >> LGP] = acc;
            } else {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) wsum[warp] = acc;
                __syncthreads();
                if ((tid & (gsz - 1)) == 0) {
                    double t = 0.0;
#pragma unroll
                    for (int w = 0; w < gsz / 32; ++w) t += wsum[warp + w];
                    cl[e0 >> LGP] = t;
                }
                __syncthreads();
            }
        } else if constexpr (LGP == 1) {
            cl[e0 >> 1] = t4[0] + t4[1]; cl[(e0 >> 1) + 1] = t4[2] + t4[3];
        } else {
            cl[e0] = t4[0]; cl[e0 + 1] = t4[1]; cl[e0 + 2] = t4[2]; cl[e0 + 3] = t4[3];
        }
    }
}

template <typename T, int KIND>
__global__ void __launch_bounds__(kT) bb_costs_1d_pow2_k(double *costs, const T *X, int n, int lgn, int K, long nn)
{
    __shared__ double wsum[kT / 32];
    __shared__ double s_inv;
    const long k = blockIdx.x;
    const T *Xk = X + k * (long)n * K;
    double *ck = costs + k * nn;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nchunk = n >> 10;
    double v[4];
    // nrm = norm(X[:,1])
    double a = 0.0;
    for (int c = 0; c < nchunk; ++c) {
        BbLoad4<T>::ld(Xk + c * 1024 + tid * 4, v);
        a = fma(v[0], v[0], a); a = fma(v[1], v[1], a); a = fma(v[2], v[2], a); a = fma(v[3], v[3], a);
    }
    a = group_sum(a, 32);
    if (lane == 0) wsum[warp] = a;
    __syncthreads();
    if (tid == 0) {
        double t = 0.0;
        for (int w = 0; w < kT / 32; ++w) t += wsum[w];
        double nrm = sqrt(t);
        if (sizeof(T) == 4) nrm = (double)(float)nrm;
        s_inv = nrm > 0.0 ? 1.0 / nrm : 0.0;                 // nrm == 0: every cost is 0 (bestbasis_costs.jl:119)
    }
    __syncthreads();
    const double inv = s_inv;
    for (int l = 0; l < K; ++l) {
        const int lgp = lgn - l;                              // node length 2^lgp
        const T *lev = Xk + (long)l * n;
        double *cl = ck + ((1L << l) - 1);
        if (lgp >= 10) {
            // a node is one or more whole chunks: accumulate per thread over the node's chunks, then one block reduction
            const int cpn = 1 << (lgp - 10);
            for (int j = 0; j < (1 << l); ++j) {
                double acc = 0.0;
                for (int c = 0; c < cpn; ++c) {
                    BbLoad4<T>::ld(lev + (j * cpn + c) * 1024 + tid * 4, v);
                    double t4[4];
                    bb_terms4<KIND>(v, inv, t4);
                    acc += (t4[0] + t4[1]) + (t4[2] + t4[3]);
                }
                acc = group_sum(acc, 32);
                if (lane == 0) wsum[warp] = acc;
                __syncthreads();
                if (tid == 0) {
                    double t = 0.0;
                    for (int w = 0; w < kT / 32; ++w) t += wsum[w];
                    cl[j] = t;
                }
                __syncthreads();
            }
            continue;
        }
        switch (lgp) {                                       // node length known at compile time inside each case: no per-step tests
            case 0: bb_level_small<T, KIND, 0>(lev, cl, nchunk, tid, inv, wsum); break;
            case 1: bb_level_small<T, KIND, 1>(lev, cl, nchunk, tid, inv, wsum); break;
            case 2: bb_level_small<T, KIND, 2>(lev, cl, nchunk, tid, inv, wsum); break;
            case 3: bb_level_small<T, KIND, 3>(lev, cl, nchunk, tid, inv, wsum); break;
            case 4: bb_level_small<T, KIND, 4>(lev, cl, nchunk, tid, inv, wsum); break;
            case 5: bb_level_small<T, KIND, 5>(lev, cl, nchunk, tid, inv, wsum); break;
            case 6: bb_level_small<T, KIND, 6>(lev, cl, nchunk, tid, inv, wsum); break;
            case 7: bb_level_small<T, KIND, 7>(lev, cl, nchunk, tid, inv, wsum); break;
            case 8: bb_level_small<T, KIND, 8>(lev, cl, nchunk, tid, inv, wsum); break;
            default: bb_level_small<T, KIND, 9>(lev, cl, nchunk, tid, inv, wsum); break;
        }
    }
}

// any geometry (2-D quad trees, redundant 2-D): one warp per (signal, node) through node_elem.  pernode: the 2-D non-redundant
// branch of the reference normalises every node by its own norm (bestbasis_tree.jl:252 passes no nrm) -- kept as written.
template <typename T>
__global__ void __launch_bounds__(kT) bb_costs_generic_k(double *costs, const T *X, NodeGeom g, long nn, long N, long szK, long sz0, int kind, int pernode)
{
    const long w = ((long)blockIdx.x * kT + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= nn * N) return;
    const long k = w / nn, q = w - k * nn;
    const T *Xk = X + k * szK;
    long cnt; double scale;
    node_elem(g, q, 0, cnt, scale);
    double a = 0.0;
    if (pernode) { for (long t = lane; t < cnt; t += 32) { long c2; double s2; const double v = (double)Xk[node_elem(g, q, t, c2, s2)]; a = fma(v, v, a); } }
    else { for (long e = lane; e < sz0; e += 32) { const double v = (double)Xk[e]; a = fma(v, v, a); } }
    a = group_sum(a, 32);
    double nrm = sqrt(a);
    if (sizeof(T) == 4) nrm = (double)(float)nrm;
    const double inv = nrm > 0.0 ? 1.0 / nrm : 0.0;
    double acc = 0.0;
    for (long t = lane; t < cnt; t += 32) { long c2; double s2; acc += bb_term((double)Xk[node_elem(g, q, t, c2, s2)], inv, kind); }
    acc = group_sum(acc, 32);
    if (lane == 0) costs[k * nn + q] = acc * scale;
}

// bestbasis_treeselection (BestBasis.jl:59-110) for N signals at once, one warp per signal.  Bottom-up: a node keeps its split
// iff the (already updated) children cost less; top-down: a node survives iff all its ancestors kept theirs -- the same tree
// as the reference's delete_subtree! cascade.  costs (N, nn) are updated in place like the reference's.
__global__ void __launch_bounds__(kT) bb_select_k(unsigned char *trees, double *costs, long nn, long ntree, int L, int ar, long N, int elt)
{
    const long k = ((long)blockIdx.x * kT + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (k >= N) return;
    double *c = costs + k * nn;
    unsigned char *t = trees + k * ntree;
    for (long i = lane; i < ntree; i += 32) t[i] = 0;
    __syncwarp();
    for (int l = L - 1; l >= 0; --l) {
        const long first = ar == 2 ? (1L << l) : ((1L << (2 * l)) - 1) / 3 + 1, cntl = ar == 2 ? (1L << l) : (1L << (2 * l));
        for (long j = lane; j < cntl; j += 32) {
            const long i = first + j;                                 // 1-based node
            if (i > ntree) continue;
            const long c0 = ar == 2 ? 2 * i : 4 * i - 2;
            double cc = c[c0 - 1] + c[c0];
            if (elt == 4) cc = (double)(float)cc;
            if (ar == 4) { cc += c[c0 + 1]; if (elt == 4) cc = (double)(float)cc; cc += c[c0 + 2]; if (elt == 4) cc = (double)(float)cc; }
            if (cc < c[i - 1]) { c[i - 1] = cc; t[i - 1] = 1; }
        }
        __syncwarp();
    }
    for (int l = 1; l < L; ++l) {
        const long first = ar == 2 ? (1L << l) : ((1L << (2 * l)) - 1) / 3 + 1, cntl = ar == 2 ? (1L << l) : (1L << (2 * l));
        for (long j = lane; j < cntl; j += 32) {
            const long i = first + j;
            if (i > ntree) continue;
            const long par = ar == 2 ? (i >> 1) : ((i + 2) >> 2);
            if (!t[par - 1]) t[i - 1] = 0;
        }
        __syncwarp();
    }
}

// getbasiscoefall with one tree per signal (Utils.jl:199-225): walk the signal's tree from the root to the leaf that covers
// the position and read that level.  trees (N, ntree) bytes on the device.
template <typename T>
__global__ void __launch_bounds__(kT) gather_multi_k(T *out, const T *Xw, long m, long n, int K, long N, const unsigned char *trees, long ntree)
{
    const long sz = (m > 0 ? m : 1) * n;
    const long gid = (long)blockIdx.x * kT + threadIdx.x;
    if (gid >= sz * N) return;
    const long k = gid / sz, e = gid - k * sz;
    const unsigned char *t = trees + k * ntree;
    int d = 0;
    if (m == 0) {
        long idx = 1;
        while (idx <= ntree && t[idx - 1] && d < K - 1) {
            const long p = n >> (d + 1);                                  // child length
            const long j = idx - (1L << d);                               // node index within the depth
            idx = 2 * idx + ((e - j * 2 * p) >= p ? 1 : 0);
            ++d;
        }
    } else {
        const long r = e % m, c = e / m;
        long idx = 1, r0 = 0, c0 = 0, nr = m, nc = n;
        while (idx <= ntree && t[idx - 1] && d < K - 1) {
            nr >>= 1; nc >>= 1;
            const int rb = (r - r0) >= nr, cb = (c - c0) >= nc;
            if (rb) r0 += nr;
            if (cb) c0 += nc;
            idx = 4 * idx - 2 + 2 * rb + cb;
            ++d;
        }
    }
    out[gid] = Xw[(k * K + d) * sz + e];
}

}  // namespace

extern "C" {

int wx_jbb_moments_f64(double *sum, double *sumsq, const double *X, long szK, long Nlocal, void *stream)
{
    return moments<double>(sum, sumsq, X, szK, Nlocal, (cudaStream_t)stream);
}
int wx_jbb_moments_f32(double *sum, double *sumsq, const float *X, long szK, long Nlocal, void *stream)
{
    return moments<float>(sum, sumsq, X, szK, Nlocal, (cudaStream_t)stream);
}

int wx_jbb_costs(double *costs_host, const double *sum, const double *sumsq, long Ntotal, long m, long n, int K, int redundant, int cost_kind,
                 double p, int elt, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    WX_REQUIRE(costs_host && sum && sumsq, "null pointer");
    WX_REQUIRE(Ntotal >= 1 && n >= 1 && m >= 0 && K >= 1, "bad sizes");
    WX_REQUIRE(cost_kind == 0 || cost_kind == 1, "unknown JBB cost kind %d", cost_kind);
    WX_REQUIRE(elt == 4 || elt == 8, "elt must be 4 or 8");
    if (!redundant) WX_REQUIRE(m > 0 ? 2 * K < 62 : K < 62, "too many levels");
    const long szK = (m > 0 ? m : 1) * n * K;
    double *term; int *bad;
    int rc = wx_scratch(&term, (size_t)szK, s); if (rc) return rc;
    rc = wx_scratch(&bad, 1, s); if (rc) return rc;
    WX_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), s));
    jbb_term_k<<<gridf(szK), kT, 0, s>>>(term, sum, sumsq, szK, (double)Ntotal, cost_kind, p, elt, bad);
    WX_LAUNCHED();
    const double mult = (cost_kind == 0) ? p : 1.0;       // LoglpCost: p * sum(log|x|) ; NormCost: norm(x,p)^p = sum |x|^p
    rc = node_costs_to_host(costs_host, term, m, n, K, redundant, mult, elt, s);
    int hbad = 0;
    if (!rc) { WX_CUDA(cudaMemcpyAsync(&hbad, bad, sizeof(int), cudaMemcpyDeviceToHost, s)); WX_CUDA(cudaStreamSynchronize(s)); }
    int rc2 = wx_scratch_free(term, s), rc3 = wx_scratch_free(bad, s);
    if (rc) return rc;
    if (rc2 || rc3) return rc2 ? rc2 : rc3;
    if (hbad) return wx_fail(WX_EINVAL, "DomainError/AssertionError: negative variance, all(sigma .>= 0) failed");
    return WX_OK;
}

int wx_lsdb_grid(long Ntotal, long *nbins, long *mbins, long *npts)
{
    WX_REQUIRE(Ntotal >= 1, "bad N");
    const LsdbGrid g = lsdb_grid(Ntotal);
    if (nbins) *nbins = g.nbins;
    if (mbins) *mbins = g.mbins;
    if (npts) *npts = g.npts;
    return WX_OK;
}

// stats(7, szK): row 0 = shift (INPUT, caller-filled, e.g. the first signal of the global batch); rows 1..6 = outputs:
// sum(x-c) hi/lo, sum((x-c)^2) hi/lo (double-double), min, max over the LOCAL batch.
int wx_lsdb_pass1_f64(double *stats, const double *X, long szK, long Nlocal, void *stream) { return lsdb_pass1<double>(stats, X, szK, Nlocal, (cudaStream_t)stream); }
int wx_lsdb_pass1_f32(double *stats, const float *X, long szK, long Nlocal, void *stream) { return lsdb_pass1<float>(stats, X, szK, Nlocal, (cudaStream_t)stream); }
// parts (nparts, 2, count) double-double numbers (hi row, lo row) -> out (2, count), summed in index order: the cross-rank
// combination of all-gathered LSDB statistics / log sums (exact to ~1e-32, hence independent of the sharding)
int wx_dd_sum(double *out, const double *parts, long count, int nparts, void *stream)
{
    WX_REQUIRE(out && parts && count >= 0 && nparts >= 1, "bad arguments");
    if (count == 0) return WX_OK;
    dd_sum_parts_k<<<gridf(count), kT, 0, (cudaStream_t)stream>>>(out, parts, count, nparts);
    WX_LAUNCHED();
    return WX_OK;
}
int wx_lsdb_pass2_f64(double *counts, const double *stats, const double *X, long szK, long Nlocal, long Ntotal, void *s) { return lsdb_pass2<double>(counts, stats, X, szK, Nlocal, Ntotal, (cudaStream_t)s); }
int wx_lsdb_pass2_f32(double *counts, const double *stats, const float *X, long szK, long Nlocal, long Ntotal, void *s) { return lsdb_pass2<float>(counts, stats, X, szK, Nlocal, Ntotal, (cudaStream_t)s); }
int wx_lsdb_pass3_f64(double *logsum, const double *counts, const double *stats, const double *X, long szK, long Nlocal, long Ntotal, void *s) { return lsdb_pass3<double>(logsum, counts, stats, X, szK, Nlocal, Ntotal, (cudaStream_t)s); }
int wx_lsdb_pass3_f32(double *logsum, const double *counts, const double *stats, const float *X, long szK, long Nlocal, long Ntotal, void *s) { return lsdb_pass3<float>(logsum, counts, stats, X, szK, Nlocal, Ntotal, (cudaStream_t)s); }

int wx_lsdb_costs(double *costs_host, const double *logsum, long Ntotal, long m, long n, int K, int redundant, void *stream)
{
    cudaStream_t s = (cudaStream_t)stream;
    WX_REQUIRE(costs_host && logsum, "null pointer");
    WX_REQUIRE(Ntotal >= 1 && n >= 1 && m >= 0 && K >= 1, "bad sizes");
    const long szK = (m > 0 ? m : 1) * n * K;
    double *term; int rc = wx_scratch(&term, (size_t)szK, s); if (rc) return rc;
    scale_k<<<gridf(szK), kT, 0, s>>>(term, logsum, szK, -1.0 / (double)Ntotal);     // ent = -(1/N) sum log pdf
    WX_LAUNCHED();
    rc = node_costs_to_host(costs_host, term, m, n, K, redundant, 1.0, 8, s);
    int rc2 = wx_scratch_free(term, s);
    if (rc || rc2) return rc ? rc : rc2;
    const long nn = count_nodes(m, K, redundant);
    for (long i = 0; i < nn; ++i)
        if (costs_host[i] != costs_host[i])
            return wx_fail(WX_EINVAL, "ArgumentError: range step cannot be zero (a coefficient position is constant over the batch, bestbasis_costs.jl:146)");
    return WX_OK;
}

// bestbasis_treeselection  BestBasis.jl:59-110 + delete_subtree! :128-140 (host)
int wx_tree_select(unsigned char *tree_out, double *costs, long ncosts, long m, long n, int minmax)
{
    WX_REQUIRE(tree_out && costs, "null pointer");
    WX_REQUIRE(n >= 1 && m >= 0 && ncosts >= 1, "bad sizes");
    WX_REQUIRE(minmax == 0 || minmax == 1, "ArgumentError: Unsupported type");
    const int ar = m > 0 ? 4 : 2;
    long ntree, nfull;
    int L;
    if (m > 0) {
        const int Lm = wx_maxlevels(m < n ? m : n);
        ntree = ((1L << (2 * Lm)) - 1) / 3;
        WX_REQUIRE(ncosts <= ((1L << (2 * (Lm + 1))) - 1) / 3, "AssertionError: k <= gettreelength(2n,2m)");
        L = wx_quaddepthl(ncosts);
        nfull = ((1L << (2 * L)) - 1) / 3;
    } else {
        ntree = n - 1;
        WX_REQUIRE(ncosts <= (1L << (wx_maxlevels(2 * n))) - 1, "AssertionError: k <= gettreelength(2n)");
        L = wx_ilog2l(ncosts);
        nfull = (1L << L) - 1;
    }
    memset(tree_out, 0, (size_t)(ntree > 0 ? ntree : 0));
    for (long i = 1; i <= nfull && i <= ntree; ++i) tree_out[i - 1] = 1;          // maketree(n, L, :full)
    std::vector<long> stack;
    for (long i = ntree; i >= 1; --i) {
        if (!tree_out[i - 1]) continue;
        const long c0 = (ar == 2) ? 2 * i : 4 * i - 2;
        if (c0 + ar - 1 > ncosts) return wx_fail(WX_EINVAL, "cost vector too short for node %ld", i);
        const double pc = costs[i - 1];
        double cc = costs[c0 - 1] + costs[c0];
        if (ar == 4) cc = (cc + costs[c0 + 1]) + costs[c0 + 2];
        if ((minmax == 0 && cc < pc) || (minmax == 1 && cc > pc)) { costs[i - 1] = cc; continue; }
        stack.assign(1, i);                                                         // delete_subtree!
        while (!stack.empty()) {
            const long q = stack.back(); stack.pop_back();
            tree_out[q - 1] = 0;
            for (int c = 0; c < ar; ++c) {
                const long ch = (ar == 2) ? 2 * q + c : 4 * q - 2 + c;
                if (ch <= ntree && tree_out[ch - 1]) stack.push_back(ch);
            }
        }
    }
    return WX_OK;
}


}  // extern "C"

// tree_costs(X, ::BB) for every signal of a batch: X (sz, K, N) -> costs (nnodes, N) [device, Float64], nnodes as for JBB.
template <typename T>
static int bb_costs_impl(double *costs, const T *X, long m, long n, int K, long N, int redundant, int kind, cudaStream_t s)
{
    WX_REQUIRE(costs && n >= 1 && m >= 0 && K >= 1 && N >= 0 && (N == 0 || X), "bad arguments");
    WX_REQUIRE(kind == 0 || kind == 1, "unknown BB cost kind %d", kind);
    if (!redundant) WX_REQUIRE(m > 0 ? 2 * K < 62 : K < 62, "too many levels");
    if (N == 0) return WX_OK;
    const long nn = count_nodes(m, K, redundant);
    if (m == 0) {
        WX_REQUIRE(N < (1L << 31), "too many signals for one launch");
        static const bool generic = getenv("WX_B200_BB_GENERIC") != nullptr;          // A-B measurements only
        if (!generic && !redundant && wx_ispow2(n) && n >= 1024 && n < (1L << 30) && K <= wx_ilog2l(n) + 1 && (((uintptr_t)X) & 15) == 0) {
            if (kind == 0) bb_costs_1d_pow2_k<T, 0><<<(unsigned)N, kT, 0, s>>>(costs, X, (int)n, wx_ilog2l(n), K, nn);
            else bb_costs_1d_pow2_k<T, 1><<<(unsigned)N, kT, 0, s>>>(costs, X, (int)n, wx_ilog2l(n), K, nn);
        } else {
            bb_costs_1d_k<T><<<(unsigned)N, kT, 0, s>>>(costs, X, n, K, nn, redundant, kind);
        }
    } else {
        NodeGeom g{m, n, K, redundant};
        const long sz = m * n;
        bb_costs_generic_k<T><<<gridf(nn * N * 32), kT, 0, s>>>(costs, X, g, nn, N, sz * K, sz, kind, redundant ? 0 : 1);
    }
    WX_LAUNCHED();
    return WX_OK;
}
extern "C" {

int wx_bb_costs_f64(double *costs, const double *X, long m, long n, int K, long N, int redundant, int kind, void *s) { return bb_costs_impl<double>(costs, X, m, n, K, N, redundant, kind, (cudaStream_t)s); }
int wx_bb_costs_f32(double *costs, const float *X, long m, long n, int K, long N, int redundant, int kind, void *s) { return bb_costs_impl<float>(costs, X, m, n, K, N, redundant, kind, (cudaStream_t)s); }

// energy_map(Xw, y, ::TimeFrequency) numerators: esum (nc, szK) device Float64 = per class, per position sum of squares over the
// LOCAL signals of the class (labels: device int32 in [0, nc)).  The caller all-reduces esum, divides class c by
// norm_sum_c = sum of esum[c] over the level-0 positions (ldb_energymap.jl:130-136).
}  // extern "C"
template <typename T>
static int energy_tf_impl(double *esum, const T *X, const int *labels, int nc, long szK, long N, cudaStream_t s)
{
    WX_REQUIRE(esum && labels && nc >= 1 && szK >= 1 && N >= 0 && (N == 0 || X), "bad arguments");
    if (N == 0) { WX_CUDA(cudaMemsetAsync(esum, 0, (size_t)nc * szK * sizeof(double), s)); return WX_OK; }
    WxDev dv; int rc = wx_devinfo(dv); if (rc) return rc;
    constexpr int NC = 4;
    constexpr long V = 16 / (long)sizeof(T);
    const bool vec = szK % V == 0 && (((uintptr_t)X) & 15) == 0;
    const int ksplit = pick_ksplit(vec ? szK / V : szK, N, dv.sms);
    const long kchunk = (N + ksplit - 1) / ksplit;
    double *part; rc = wx_scratch(&part, (size_t)ksplit * NC * szK, s); if (rc) return rc;
    for (int c0 = 0; c0 < nc; c0 += NC) {
        if (vec) {
            dim3 grid((unsigned)((szK / V + kT - 1) / kT), (unsigned)ksplit);
            energy_tf_part_k<T, NC><<<grid, kT, 0, s>>>(part, X, labels, c0, szK, N, kchunk);
        } else {
            dim3 grid((unsigned)((szK + kT - 1) / kT), (unsigned)ksplit);
            energy_tf_part_scalar_k<T, NC><<<grid, kT, 0, s>>>(part, X, labels, c0, szK, N, kchunk);
        }
        WX_LAUNCHED();
        energy_tf_final_k<NC><<<gridf(szK), kT, 0, s>>>(esum + (long)c0 * szK, part, szK, ksplit, nc - c0 < NC ? nc - c0 : NC);
        WX_LAUNCHED();
    }
    return wx_scratch_free(part, s);
}
extern "C" {
int wx_energy_map_tf_f64(double *esum, const double *X, const int *labels_dev, int nc, long szK, long Nlocal, void *s) { return energy_tf_impl<double>(esum, X, labels_dev, nc, szK, Nlocal, (cudaStream_t)s); }
int wx_energy_map_tf_f32(double *esum, const float *X, const int *labels_dev, int nc, long szK, long Nlocal, void *s) { return energy_tf_impl<float>(esum, X, labels_dev, nc, szK, Nlocal, (cudaStream_t)s); }
// discriminant_measure on the (all-reduced) energy sums: D (szK) device; inv_norm (nc) device = 1 / norm_sum_c
int wx_ldb_discriminant(double *D, const double *esum, const double *inv_norm_dev, int nc, long szK, int kind, double p, int elt, void *stream)
{
    WX_REQUIRE(D && esum && inv_norm_dev && nc >= 2 && szK >= 1, "bad arguments");
    WX_REQUIRE(kind >= 0 && kind <= 3, "unknown discriminant measure %d", kind);
    ldb_dm_k<<<gridf(szK), kT, 0, (cudaStream_t)stream>>>(D, esum, inv_norm_dev, nc, szK, kind, p, elt);
    WX_LAUNCHED();
    return WX_OK;
}
// per-node sums of a per-position term (the LDB node costs with top_k >= node size, LDB.jl:217-237): costs HOST
int wx_node_costs(double *costs_host, const double *term, long m, long n, int K, int redundant, double mult, int elt, void *stream)
{
    WX_REQUIRE(costs_host && term && n >= 1 && m >= 0 && K >= 1, "bad arguments");
    return node_costs_to_host(costs_host, term, m, n, K, redundant, mult, elt, (cudaStream_t)stream);
}

// bestbasis_treeselection for N cost vectors at once: trees (ntree, N) bytes on the device, costs (nnodes, N) updated in place
int wx_bb_select(unsigned char *trees, double *costs, long nnodes, long m, long n, long N, int elt, void *stream)
{
    WX_REQUIRE(trees && costs && nnodes >= 1 && n >= 1 && m >= 0 && N >= 0, "bad arguments");
    WX_REQUIRE(elt == 4 || elt == 8, "elt must be 4 or 8");
    if (N == 0) return WX_OK;
    const int ar = m > 0 ? 4 : 2;
    long ntree; int L;
    if (m > 0) {
        const int Lm = wx_maxlevels(m < n ? m : n);
        ntree = ((1L << (2 * Lm)) - 1) / 3;
        WX_REQUIRE(nnodes <= ((1L << (2 * (Lm + 1))) - 1) / 3, "AssertionError: k <= gettreelength(2n,2m)");
        L = wx_quaddepthl(nnodes);
    } else {
        ntree = n - 1;
        WX_REQUIRE(nnodes <= (1L << (wx_maxlevels(2 * n))) - 1, "AssertionError: k <= gettreelength(2n)");
        L = wx_ilog2l(nnodes);
    }
    if (ntree <= 0) return WX_OK;
    bb_select_k<<<gridf(N * 32), kT, 0, (cudaStream_t)stream>>>(trees, costs, nnodes, ntree, L, ar, N, elt);
    WX_LAUNCHED();
    return WX_OK;
}

}  // extern "C"

// 1-D signals: one CTA per signal.  The signal's tree is staged in shared memory with coalesced loads (the generic kernel walks it
// from global memory, ten dependent byte loads per element: 3.0 ms for 131072 x 1024, 0.11 of the roofline), validated there
// (flag bit 0: a split node whose parent is not split, bit 1: a split node at depth >= K-1) and walked per element.
template <typename T>
__global__ void __launch_bounds__(kT) gather_multi_1d_k(T *out, const T *Xw, long n, int K, const unsigned char *trees, long ntree, int *flag)
{
    extern __shared__ unsigned char wx_gt[];
    const long k = blockIdx.x;
    const unsigned char *t = trees + k * ntree;
    for (long i = threadIdx.x; i < ntree; i += kT) wx_gt[i] = t[i];
    __syncthreads();
    int bad = 0;
    for (long i = threadIdx.x + 1; i <= ntree; i += kT) {                 // 1-based node
        if (!wx_gt[i - 1]) continue;
        if (i > 1 && !wx_gt[i / 2 - 1]) bad |= 1;
        if (ilog2d(i) >= K - 1) bad |= 2;
    }
    if (bad) atomicOr(flag, bad);
    const T *Xk = Xw + k * (long)K * n;
    // leaf depth of position e: walk the staged tree (32-bit indices: ntree = n - 1 < 2^31)
    const int nt = (int)ntree, ni = (int)n;
    auto leaf_depth = [&](int e) {
        int d = 0, idx = 1;
        while (idx <= nt && wx_gt[idx - 1] && d < K - 1) {
            const int p = ni >> (d + 1), j = idx - (1 << d);
            idx = 2 * idx + ((e - j * 2 * p) >= p ? 1 : 0);
            ++d;
        }
        return d;
    };
    // four positions per thread and iteration: the four walks first, then the four gathers in flight together
    int e = threadIdx.x;
    for (; e + 3 * kT < ni; e += 4 * kT) {
        int d[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) d[q] = leaf_depth(e + q * kT);
        T v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = Xk[(long)d[q] * n + e + q * kT];
#pragma unroll
        for (int q = 0; q < 4; ++q) out[k * n + e + q * kT] = v[q];
    }
    for (; e < ni; e += kT) out[k * n + e] = Xk[(long)leaf_depth(e) * n + e];
}

// the checks of getbasiscoefall(Xw, tree::BitArray{2}) Utils.jl:204-218 on the device: every tree valid (a split node's parent is
// split) and no split node at depth >= K-1 ("Not enough decomposition levels in Xw"); flag bit 0 / bit 1
__global__ void __launch_bounds__(kT) trees_check_k(int *flag, const unsigned char *trees, long ntree, long N, int ar, int K)
{
    const long gid = (long)blockIdx.x * kT + threadIdx.x;
    if (gid >= ntree * N) return;
    const long k = gid / ntree, i = gid - k * ntree + 1;                 // 1-based node
    const unsigned char *t = trees + k * ntree;
    if (!t[i - 1]) return;
    int bad = 0;
    if (i > 1) {
        const long p = ar == 2 ? i / 2 : (i + 2) / 4;
        if (!t[p - 1]) bad |= 1;
    }
    const int d = ar == 2 ? ilog2d(i) : quaddepthd(i);
    if (d >= K - 1) bad |= 2;
    if (bad) atomicOr(flag, bad);
}

// per-signal-tree gather (wx_trees.cu exports the C entry points)
template <typename T>
int wx_gather_multi(T *out, const T *Xw, long m, long n, int K, long N, const unsigned char *trees, long ntree, cudaStream_t s)
{
    const long sz = (m > 0 ? m : 1) * n;
    if (sz * N == 0) return WX_OK;
    long expect = n - 1;
    if (m > 0) { const int Lm = wx_maxlevels(m < n ? m : n); expect = ((1L << (2 * Lm)) - 1) / 3; }
    WX_REQUIRE(ntree == expect, "AssertionError: n_t == gettreelength(sz...) (%ld != %ld)", ntree, expect);
    if (m == 0 && ntree > 0 && ntree <= 48 * 1024 && N < (1L << 31)) {
        int *flag; int rc = wx_scratch(&flag, 1, s); if (rc) return rc;
        WX_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
        gather_multi_1d_k<T><<<(unsigned)N, kT, (size_t)ntree, s>>>(out, Xw, n, K, trees, ntree, flag);
        WX_LAUNCHED();
        int h = 0;
        WX_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
        WX_CUDA(cudaStreamSynchronize(s));
        rc = wx_scratch_free(flag, s); if (rc) return rc;
        if (h & 1) return wx_fail(WX_EINVAL, "AssertionError: all trees must be valid (isvalidtree)");
        if (h & 2) return wx_fail(WX_EINVAL, "ArgumentError: Not enough decomposition levels in Xw.");
        return WX_OK;
    }
    if (ntree > 0) {
        int *flag; int rc = wx_scratch(&flag, 1, s); if (rc) return rc;
        WX_CUDA(cudaMemsetAsync(flag, 0, sizeof(int), s));
        trees_check_k<<<gridf(ntree * N), kT, 0, s>>>(flag, trees, ntree, N, m > 0 ? 4 : 2, K);
        WX_LAUNCHED();
        int h = 0;
        WX_CUDA(cudaMemcpyAsync(&h, flag, sizeof(int), cudaMemcpyDeviceToHost, s));
        WX_CUDA(cudaStreamSynchronize(s));
        rc = wx_scratch_free(flag, s); if (rc) return rc;
        if (h & 1) return wx_fail(WX_EINVAL, "AssertionError: all trees must be valid (isvalidtree)");
        if (h & 2) return wx_fail(WX_EINVAL, "ArgumentError: Not enough decomposition levels in Xw.");
    }
    gather_multi_k<T><<<gridf(sz * N), kT, 0, s>>>(out, Xw, m, n, K, N, trees, ntree);
    WX_LAUNCHED();
    return WX_OK;
}
template int wx_gather_multi<double>(double *, const double *, long, long, int, long, const unsigned char *, long, cudaStream_t);
template int wx_gather_multi<float>(float *, const float *, long, long, int, long, const unsigned char *, long, cudaStream_t);

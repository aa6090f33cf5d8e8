/*
 * wx_b200.h -- C ABI of libwx_b200.so, the B200 (sm_100a) implementation of the filter-bank hot path
 * of WaveletsExt.jl v0.2.3 (batched WPD / iWPT, SWT and ACWT families, 2-D WPD, JBB/LSDB cost trees).
 *
 * The reference has no FFI of its own (pure Julia, multiple dispatch); each entry point below is the
 * device counterpart of one reference function and cites it (paths relative to src/mod/ of the
 * reference).  INTEGRATION.md shows the Julia `ccall` binding for each.
 *
 * Conventions
 *  - plain C types only; every function returns 0 on success, a WX_E* code otherwise, and never throws.
 *    wx_last_error() returns a thread-local message for the last failure.
 *  - signal data pointers are DEVICE pointers on the current CUDA device unless the function name ends
 *    in `_host`; taps, trees and cost vectors are HOST pointers.
 *  - memory layout is the reference's (Julia column-major): first index fastest, batch last.
 *    x(n,N)  y(n,L+1,N)  x2(m,n,N)  y2(m,n,L+1,N)  xw(n,nodes,N) ...
 *  - `stream` is a cudaStream_t (0 = default stream).  Calls are asynchronous on that stream unless the
 *    result is returned in host memory (costs, trees), in which case the call synchronises the stream.
 *  - taps are double precision like the reference's (WT.makereverseqmfpair default eltype); for
 *    Float32 signals they are rounded once to Float32 and products are accumulated with FP32 FMAs.
 *  - there is no CPU fallback: without a CUDA device every compute entry point returns WX_ECUDA.
 */
#ifndef WX_B200_H
#define WX_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WX_OK 0
#define WX_EINVAL 1    /* argument the reference would reject with @assert / ArgumentError */
#define WX_ECUDA 2     /* CUDA runtime failure */
#define WX_EUNSUPPORTED 3
#define WX_ENOMEM 4

#define WX_MAX_TAPS 64 /* orthogonal filters up to 32 taps; autocorrelation filters (2F-1) up to 63 */

/* modes of the redundant (stationary / autocorrelation) drivers */
#define WX_MODE_DWT 0 /* sdwt / acdwt   : xw (n, L+1, N)        SWT.jl:109-131, ACWT.jl:109-131 */
#define WX_MODE_WPT 1 /* swpt / acwpt   : xw (n, 2^L, N)        SWT.jl:439-471, ACWT.jl:427-460 */
#define WX_MODE_WPD 2 /* swpd / acwpd   : xw (n, 2^(L+1)-1, N)  SWT.jl:840-868, ACWT.jl:733-759 */

/* ---- runtime ------------------------------------------------------------------------------------ */
int wx_version(void);
const char *wx_last_error(void);
int wx_device_count(int *count);
int wx_set_device(int dev);
int wx_device_info(int *sm_count, int *cc_major, int *cc_minor, size_t *smem_optin, size_t *total_mem);
int wx_malloc(void **dptr, size_t bytes);
int wx_free(void *dptr);
int wx_malloc_host(void **hptr, size_t bytes);      /* pinned host memory */
int wx_free_host(void *hptr);
int wx_h2d(void *dst_dev, const void *src_host, size_t bytes, void *stream);
int wx_d2h(void *dst_host, const void *src_dev, size_t bytes, void *stream);
int wx_stream_sync(void *stream);
int wx_device_sync(void);
/* Scratch of the batched entry points is stream-ordered memory from a PRIVATE pool per device (the process-wide default pool is
 * left alone).  Freed blocks stay cached there up to a quarter of the device memory ($WX_B200_SCRATCH_KEEP_MB overrides the
 * bound) so that repeated calls do not re-map their workspace; wx_trim_scratch hands everything above keep_bytes back to the
 * driver (call it when the embedding application needs the memory: it synchronises the device first). */
int wx_trim_scratch(size_t keep_bytes);
/* number of kernels this library has launched in this process (bench.py's gpu_launches) */
unsigned long long wx_launch_count(void);

/* ---- single steps, 1-D -------------------------------------------------------------------------- */
/* dwt_step!(w1,w2,v,h,g)      dwt/dwt_one_level.jl:79-107 ; v(n) -> w1,w2 (n/2) */
int wx_dwt_step_f64(double *w1, double *w2, const double *v, long n, const double *h, const double *g, int F, void *stream);
int wx_dwt_step_f32(float *w1, float *w2, const float *v, long n, const double *h, const double *g, int F, void *stream);
/* idwt_step!(v,w1,w2,h,g)     dwt/dwt_one_level.jl:192-223 */
int wx_idwt_step_f64(double *v, const double *w1, const double *w2, long n, const double *h, const double *g, int F, void *stream);
int wx_idwt_step_f32(float *v, const float *w1, const float *w2, long n, const double *h, const double *g, int F, void *stream);
/* sdwt_step!(w1,w2,v,d,h,g)   swt/swt_one_level.jl:99-127 */
int wx_sdwt_step_f64(double *w1, double *w2, const double *v, long n, int d, const double *h, const double *g, int F, void *stream);
int wx_sdwt_step_f32(float *w1, float *w2, const float *v, long n, int d, const double *h, const double *g, int F, void *stream);
/* isdwt_step!(v,w1,w2,d,sv,sw,h,g; add2out)  swt/swt_one_level.jl:279-318 (shift based) */
int wx_isdwt_step_shift_f64(double *v, const double *w1, const double *w2, long n, int d, long sv, long sw, const double *h, const double *g, int F, int add2out, void *stream);
int wx_isdwt_step_shift_f32(float *v, const float *w1, const float *w2, long n, int d, long sv, long sw, const double *h, const double *g, int F, int add2out, void *stream);
/* isdwt_step!(v,w1,w2,d,h,g)  swt/swt_one_level.jl:257-277 (average based) */
int wx_isdwt_step_avg_f64(double *v, const double *w1, const double *w2, long n, int d, const double *h, const double *g, int F, void *stream);
int wx_isdwt_step_avg_f32(float *v, const float *w1, const float *w2, long n, int d, const double *h, const double *g, int F, void *stream);
/* acdwt_step!(w1,w2,v,d,h,g)  acwt/acwt_one_level.jl:101-128 ; w1 uses g, w2 uses h, Lf taps */
int wx_acdwt_step_f64(double *w1, double *w2, const double *v, long n, int d, const double *h, const double *g, int Lf, void *stream);
int wx_acdwt_step_f32(float *w1, float *w2, const float *v, long n, int d, const double *h, const double *g, int Lf, void *stream);
/* iacdwt_step!(v,w1,w2)       acwt/acwt_one_level.jl:217-224 */
int wx_iacdwt_step_f64(double *v, const double *w1, const double *w2, long n, void *stream);
int wx_iacdwt_step_f32(float *v, const float *w1, const float *w2, long n, void *stream);

/* ---- single steps, 2-D (matrices (nr x nc) column-major, contiguous) ------------------------------ */
/* dwt_step!(w1..w4,v,h,g,temp)  dwt/dwt_one_level.jl:319-354 ; v (2nr x 2nc) -> four (nr x nc) */
int wx_dwt_step2_f64(double *w1, double *w2, double *w3, double *w4, const double *v, long nr, long nc, const double *h, const double *g, int F, void *stream);
int wx_dwt_step2_f32(float *w1, float *w2, float *w3, float *w4, const float *v, long nr, long nc, const double *h, const double *g, int F, void *stream);
/* idwt_step!(v,w1..w4,h,g,temp) dwt/dwt_one_level.jl:401-436 */
int wx_idwt_step2_f64(double *v, const double *w1, const double *w2, const double *w3, const double *w4, long nr, long nc, const double *h, const double *g, int F, void *stream);
int wx_idwt_step2_f32(float *v, const float *w1, const float *w2, const float *w3, const float *w4, long nr, long nc, const double *h, const double *g, int F, void *stream);
/* sdwt_step! / acdwt_step! 2-D  swt/swt_one_level.jl:334-370, acwt/acwt_one_level.jl:240-276 ; all (nr x nc); ac!=0 selects AC */
int wx_rdwt_step2_f64(int ac, double *w1, double *w2, double *w3, double *w4, const double *v, long nr, long nc, int d, const double *h, const double *g, int F, void *stream);
int wx_rdwt_step2_f32(int ac, float *w1, float *w2, float *w3, float *w4, const float *v, long nr, long nc, int d, const double *h, const double *g, int F, void *stream);
/* isdwt_step! 2-D  swt/swt_one_level.jl:395-469 (sv<0: average based) ; iacdwt_step! 2-D acwt/acwt_one_level.jl:288-322 (mode 2) */
int wx_irdwt_step2_f64(int mode, double *v, const double *w1, const double *w2, const double *w3, const double *w4, long nr, long nc, int d, long sv, long sw, const double *h, const double *g, int F, void *stream);
int wx_irdwt_step2_f32(int mode, float *v, const float *w1, const float *w2, const float *w3, const float *w4, long nr, long nc, int d, long sv, long sw, const double *h, const double *g, int F, void *stream);

/* ---- batched decimated trees -------------------------------------------------------------------- */
/* wpdall / wpd!  dwt/dwt_all.jl:260-282, DWT.jl:131-161 ; x(n,N) -> y(n,L+1,N) */
int wx_wpd1d_f64(double *y, const double *x, long n, int L, long N, const double *h, const double *g, int F, void *stream);
int wx_wpd1d_f32(float *y, const float *x, long n, int L, long N, const double *h, const double *g, int F, void *stream);
/* wpdall on images  DWT.jl:164-209 ; x(m,n,N) -> y(m,n,L+1,N) */
int wx_wpd2d_f64(double *y, const double *x, long m, long n, int L, long N, const double *h, const double *g, int F, void *stream);
int wx_wpd2d_f32(float *y, const float *x, long m, long n, int L, long N, const double *h, const double *g, int F, void *stream);
/* wptall / iwptall by tree (1-D: Wavelets.jl wpt!/iwpt! call sites dwt/dwt_all.jl:162,221) ; x(n,N) -> y(n,N);
 * tree: ntree bytes (0/1), heap order, host memory */
int wx_wpt1d_f64(double *y, const double *x, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *stream);
int wx_wpt1d_f32(float *y, const float *x, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *stream);
int wx_iwpt1d_f64(double *y, const double *xw, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *stream);
int wx_iwpt1d_f32(float *y, const float *xw, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *stream);
/* 2-D wpt!/iwpt! by quad tree  DWT.jl:500-548, 662-710 ; x(m,n,N) -> y(m,n,N) */
int wx_wpt2d_f64(double *y, const double *x, long m, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *stream);
int wx_wpt2d_f32(float *y, const float *x, long m, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *stream);
int wx_iwpt2d_f64(double *y, const double *xw, long m, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *stream);
int wx_iwpt2d_f32(float *y, const float *xw, long m, long n, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *stream);
/* getbasiscoefall  Utils.jl:169-197 ; Xw(n,K,N) -> out(n,N) (m>0: Xw(m,n,K,N) -> out(m,n,N)) */
int wx_gather_basis_f64(double *out, const double *Xw, long m, long n, int K, long N, const unsigned char *tree, long ntree, void *stream);
int wx_gather_basis_f32(float *out, const float *Xw, long m, long n, int K, long N, const unsigned char *tree, long ntree, void *stream);
/* iwpdall by tree  dwt/dwt_all.jl:324-342, DWT.jl:337-401 ; Xw(n,K,N) -> x(n,N)  (m>0: 2-D) */
int wx_iwpd_f64(double *x, const double *Xw, long m, long n, int K, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *stream);
int wx_iwpd_f32(float *x, const float *Xw, long m, long n, int K, long N, const unsigned char *tree, long ntree, const double *h, const double *g, int F, void *stream);

/* ---- batched redundant trees (stationary: ac=0, taps (h,g); autocorrelation: ac=1, taps (Q,P) as (h,g)) ---- */
/* sdwtall/swptall/swpdall swt/swt_all.jl:33,156,279 ; acdwtall/acwptall/acwpdall acwt/acwt_all.jl:33,136,239.
 * 1-D: x(n,N) (m = 0).  2-D: x(m,n,N) with m rows. */
int wx_rwt_f64(int ac, int mode, double *xw, const double *x, long m, long n, int L, long N, const double *h, const double *g, int F, void *stream);
int wx_rwt_f32(int ac, int mode, float *xw, const float *x, long m, long n, int L, long N, const double *h, const double *g, int F, void *stream);
/* inverses swt/swt_all.jl:89-248,343-390 ; acwt/acwt_all.jl:86,189,300.  sm < 0: average based; sm >= 0: shift based.
 * mode WPD takes a tree (ntree bytes); modes DWT/WPT take L. ncols = size of the node dimension of xw. */
int wx_irwt_f64(int ac, int mode, double *x, const double *xw, long m, long n, long ncols, int L, long N, const unsigned char *tree, long ntree, long sm, const double *h, const double *g, int F, void *stream);
int wx_irwt_f32(int ac, int mode, float *x, const float *xw, long m, long n, long ncols, int L, long N, const unsigned char *tree, long ntree, long sm, const double *h, const double *g, int F, void *stream);

/* ---- best basis ----------------------------------------------------------------------------------- */
/* tree_costs(X, ::JBB) step 1  bestbasis/bestbasis_tree.jl:153-154 : per-position sum and sum of squares over the
 * LOCAL batch; X(sz,K,N_local) -> sum(sz*K), sumsq(sz*K) device buffers (always Float64, deterministic order).
 * The caller all-reduces (sum) these two buffers across ranks. */
int wx_jbb_moments_f64(double *sum, double *sumsq, const double *X, long szK, long Nlocal, void *stream);
int wx_jbb_moments_f32(double *sum, double *sumsq, const float *X, long szK, long Nlocal, void *stream);
/* tree_costs(X, ::JBB) step 2  bestbasis/bestbasis_tree.jl:155-207 + coefcost bestbasis/bestbasis_costs.jl:127-132 :
 * sigma and per-node costs from the (all-reduced) moments.  m = 0: 1-D (n,K) ; m > 0: 2-D (m,n,K).
 * cost_kind 0 = LoglpCost(p), 1 = NormCost(p).  costs: HOST, 2^K-1 (1-D) / (4^K-1)/3 (2-D) entries, or K if redundant.
 * elt = 8 or 4 selects the element type the reference would have computed sigma in. returns WX_EINVAL if any
 * variance is negative/NaN (reference: DomainError / @assert all(sigma .>= 0)). */
int wx_jbb_costs(double *costs_host, const double *sum, const double *sumsq, long Ntotal, long m, long n, int K, int redundant, int cost_kind, double p, int elt, void *stream);
/* tree_costs(X, ::LSDB)  bestbasis/bestbasis_tree.jl:104-147 + DifferentialEntropyCost bestbasis/bestbasis_costs.jl:135-164,
 * three passes over the local batch with all-reducible per-position state in between:
 *  pass1: stats(7, szK) device: row 0 = shift c (INPUT: first signal of the global batch), rows 1/2 = sum(x-c) hi/lo,
 *         rows 3/4 = sum((x-c)^2) hi/lo (double-double pairs, exact to ~1e-32), row 5 = min, row 6 = max
 *  pass2: ASH bin counts on the grid derived from the reduced stats -> counts(npts, szK) device
 *  pass3: sum_k log pdf(x_k) per position -> logsum(2, szK) device (hi/lo)
 *  costs: per-node costs from the reduced logsum (row hi).
 * Across ranks: all-gather the hi/lo rows and combine them with wx_dd_sum (fixed order, exact), all-reduce min / max /
 * counts; the costs are then independent of how the batch is sharded. */
int wx_lsdb_pass1_f64(double *stats, const double *X, long szK, long Nlocal, void *stream);
int wx_lsdb_pass1_f32(double *stats, const float *X, long szK, long Nlocal, void *stream);
int wx_lsdb_grid(long Ntotal, long *nbins, long *mbins, long *npts);
int wx_dd_sum(double *out, const double *parts, long count, int nparts, void *stream);
int wx_lsdb_pass2_f64(double *counts, const double *stats, const double *X, long szK, long Nlocal, long Ntotal, void *stream);
int wx_lsdb_pass2_f32(double *counts, const double *stats, const float *X, long szK, long Nlocal, long Ntotal, void *stream);
int wx_lsdb_pass3_f64(double *logsum, const double *counts, const double *stats, const double *X, long szK, long Nlocal, long Ntotal, void *stream);
int wx_lsdb_pass3_f32(double *logsum, const double *counts, const double *stats, const float *X, long szK, long Nlocal, long Ntotal, void *stream);
int wx_lsdb_costs(double *costs_host, const double *logsum, long Ntotal, long m, long n, int K, int redundant, void *stream);
/* tree_costs(X, ::BB)  bestbasis/bestbasis_tree.jl:210-256 + coefcost(::ShannonEntropyCost | ::LogEnergyEntropyCost)
 * bestbasis/bestbasis_costs.jl:103-125, for every signal of a batch (bestbasistreeall BestBasis.jl:253-262):
 * X(sz,K,N) -> costs(nnodes, N) device Float64 (nnodes as for JBB).  cost_kind 0 = Shannon, 1 = log energy.
 * wx_bb_select: bestbasis_treeselection (:min) for the N cost vectors at once, trees(ntree, N) bytes on the DEVICE, costs updated
 * in place.  wx_gather_basis_multi: getbasiscoefall with one tree per signal (Utils.jl:199-225), trees on the device. */
int wx_bb_costs_f64(double *costs_dev, const double *X, long m, long n, int K, long N, int redundant, int cost_kind, void *stream);
int wx_bb_costs_f32(double *costs_dev, const float *X, long m, long n, int K, long N, int redundant, int cost_kind, void *stream);
int wx_bb_select(unsigned char *trees_dev, double *costs_dev, long nnodes, long m, long n, long N, int elt, void *stream);
int wx_gather_basis_multi_f64(double *out, const double *Xw, long m, long n, int K, long N, const unsigned char *trees_dev, long ntree, void *stream);
int wx_gather_basis_multi_f32(float *out, const float *Xw, long m, long n, int K, long N, const unsigned char *trees_dev, long ntree, void *stream);
/* LDB (next row f-2): energy_map(Xw, y, ::TimeFrequency) ldb/ldb_energymap.jl:109-141 numerators (per class sum of squares over
 * the local signals; labels = device int32 in [0, nc)), discriminant_measure ldb/ldb_measures.jl:139-183,302-325
 * (kind 0 AsymmetricRelativeEntropy, 1 SymmetricRelativeEntropy, 2 LpDistance(p), 3 HellingerDistance) and the node sums of
 * LDB.jl:217-237 (top_k >= node size).  Tree: wx_tree_select(..., minmax = 1). */
int wx_energy_map_tf_f64(double *esum_dev, const double *X, const int *labels_dev, int nc, long szK, long Nlocal, void *stream);
int wx_energy_map_tf_f32(double *esum_dev, const float *X, const int *labels_dev, int nc, long szK, long Nlocal, void *stream);
int wx_ldb_discriminant(double *D_dev, const double *esum_dev, const double *inv_norm_dev, int nc, long szK, int kind, double p, int elt, void *stream);
int wx_node_costs(double *costs_host, const double *term_dev, long m, long n, int K, int redundant, double mult, int elt, void *stream);
/* SIWT steps and the nonstandard-form transform (next row f-4).
 * sidwt_step!(w1,w2,v,h,g,s)  siwt/siwt_one_level.jl:71-98 and isidwt_step!(v,w1,w2,h,g,s) :153-184: dwt_step!/idwt_step! with the
 *   input (output) circularly shifted by one sample when shifted != 0.
 * ns_dwt(x, wt, L)  wavemult/transforms.jl:52-74: x (n, N) -> nxw (2n, N), levels stored at ndyad(l, Lmax, gender)
 *   (wavemult/utils.jl:146-155); ns_idwt(nxw, wt, L) :120-139: nxw (n2 = 2n, N) -> x (n, N).  The reference transforms one
 *   vector; N > 1 is the batch of them. */
int wx_sidwt_step_f64(double *w1, double *w2, const double *v, long n, const double *h, const double *g, int F, int shifted, void *stream);
int wx_sidwt_step_f32(float *w1, float *w2, const float *v, long n, const double *h, const double *g, int F, int shifted, void *stream);
int wx_isidwt_step_f64(double *v, const double *w1, const double *w2, long n, const double *h, const double *g, int F, int shifted, void *stream);
int wx_isidwt_step_f32(float *v, const float *w1, const float *w2, long n, const double *h, const double *g, int F, int shifted, void *stream);
int wx_ns_dwt_f64(double *nxw, const double *x, long n, int L, long N, const double *h, const double *g, int F, void *stream);
int wx_ns_dwt_f32(float *nxw, const float *x, long n, int L, long N, const double *h, const double *g, int F, void *stream);
int wx_ns_idwt_f64(double *x, const double *nxw, long n2, int L, long N, const double *h, const double *g, int F, void *stream);
int wx_ns_idwt_f32(float *x, const float *nxw, long n2, int L, long N, const double *h, const double *g, int F, void *stream);
/* Denoising (next row f-3): the threshold determination and thresholding between getbasiscoefall and the inverse transform.
 * All arrays are device pointers except colmask (host bytes).  A signal's coefficients are an (n, K) column-major slab
 * (K = 1 for dwt / wpt vectors, K = L+1 for sdwt / acdwt, K = 2^(L+1)-1 for swpd / acwpd), N slabs back to back.
 * noisest  Denoising.jl:214-232 (Wavelets.jl Threshold.mad!): sigma[k] = median|y - median(y)| / 0.6745 over the range
 *          [off, off+len) of slab k (the caller passes finestdetailrange Utils.jl:416-436, x[:,end] or x[n/2+1:end]).
 * surethreshold      Denoising.jl:142-166 / relerrorthreshold :285-328 (orth2relerror :344-349, findelbow :366-381) over the
 *          columns selected by colmask (NULL = all: dwt, wpt, sdwt; the leaves of the tree for swpd / acwpd); t: N Float64.
 * threshold  Wavelets.jl Threshold.jl threshold! as called by denoise Denoising.jl:483-600: th 0 HardTH, 1 SoftTH,
 *          2 SemiSoftTH, 3 SteinTH; t_k = sigma[k] * tmul (sigma NULL: t = tmul).  Columns with colmask[c] == 0 and the
 *          element range [keep_lo, keep_hi) of every slab are copied unchanged (smooth = :undersmooth).  y may alias x. */
int wx_noisest_f64(double *sigma, const double *x, long stride, long off, long len, long N, void *stream);
int wx_noisest_f32(double *sigma, const float *x, long stride, long off, long len, long N, void *stream);
int wx_surethreshold_f64(double *t, const double *x, long n, long K, const unsigned char *colmask, long N, void *stream);
int wx_surethreshold_f32(double *t, const float *x, long n, long K, const unsigned char *colmask, long N, void *stream);
int wx_relerrorthreshold_f64(double *t, const double *x, long n, long K, const unsigned char *colmask, int elbows, long N, void *stream);
int wx_relerrorthreshold_f32(double *t, const float *x, long n, long K, const unsigned char *colmask, int elbows, long N, void *stream);
int wx_threshold_f64(double *y, const double *x, long n, long K, const unsigned char *colmask, long keep_lo, long keep_hi, int th, const double *sigma, double tmul, long N, void *stream);
int wx_threshold_f32(float *y, const float *x, long n, long K, const unsigned char *colmask, long keep_lo, long keep_hi, int th, const double *sigma, double tmul, long N, void *stream);
/* LDB object, feature side (LDB.jl:246-330): out(nf, N) = X(nelem, N)[order[0..nf), :] (transform / fit_transform :291-294,
 * :349-352), its inverse onto a zero-filled Xc (inverse_transform :372-378), and the per-class mean / corrected variance of every
 * coefficient for FishersClassSeparability (ldb/ldb_measures.jl:441-479): E, V (nc, nelem) Float64; sig_dev = signal indices
 * grouped by class, off_dev = nc+1 group offsets.  order_dev: 0-based int32 on the device. */
int wx_select_features_f64(double *out, const double *X, const int *order_dev, long nf, long nelem, long N, void *stream);
int wx_select_features_f32(float *out, const float *X, const int *order_dev, long nf, long nelem, long N, void *stream);
int wx_scatter_features_f64(double *Xc, const double *F, const int *order_dev, long nf, long nelem, long N, void *stream);
int wx_scatter_features_f32(float *Xc, const float *F, const int *order_dev, long nf, long nelem, long N, void *stream);
int wx_class_moments_f64(double *E, double *V, const double *X, const int *sig_dev, const int *off_dev, int nc, long nelem, void *stream);
int wx_class_moments_f32(double *E, double *V, const float *X, const int *sig_dev, const int *off_dev, int nc, long nelem, void *stream);
/* bestbasis_treeselection  BestBasis.jl:59-110 (host, O(n)); costs are modified in place like the reference.
 * m = 0: binary tree with n-1 entries; m > 0: quad tree. minmax 0 = :min, 1 = :max */
int wx_tree_select(unsigned char *tree_out, double *costs_host, long ncosts, long m, long n, int minmax);

/* ---- the collective of the path: NCCL communicators and the fused best-basis drivers (wx_comm.cu) -------------------
 * The reference reduces over the whole batch inside tree_costs: sum(X, dims=3) bestbasis/bestbasis_tree.jl:153-154 (JBB) and
 * the per-position statistics / ASH counts / log-pdf sums of bestbasis/bestbasis_costs.jl:135-164 (LSDB).  When the batch is
 * sharded over GPUs those reductions become the ONE exchange step of the path, an NCCL all-reduce / all-gather over NVLink of
 * the small per-position state; everything else is shard-local.  NCCL is resolved at run time (the libnccl.so.2 the host
 * process already mapped, else $WX_B200_NCCL, else the loader path); without it these calls return WX_EUNSUPPORTED.
 *  - one process (or thread) per GPU: rank 0 calls wx_comm_unique_id, hands the 128 bytes to the other ranks by whatever means
 *    the host has (a file, MPI, torch.distributed, Distributed.jl), every rank calls wx_comm_init_rank on its current device.
 *  - one host thread driving all GPUs (the reference is single threaded): wx_comm_init_all (ncclCommInitAll) returns one
 *    communicator per device; use the *_multi drivers, or bracket per-device wx_allreduce calls with wx_group_start/_end.
 * dtype codes WX_DT_*, reduction codes WX_OP_*; buffers are device pointers; collectives are in place and asynchronous on
 * `stream`. */
typedef struct wx_comm wx_comm_t;
#define WX_COMM_ID_BYTES 128
#define WX_DT_F64 0
#define WX_DT_F32 1
#define WX_DT_I64 2
#define WX_DT_U8 3
#define WX_OP_SUM 0
#define WX_OP_MIN 1
#define WX_OP_MAX 2
int wx_nccl_version(int *version);
int wx_comm_unique_id(unsigned char *id128);
int wx_comm_init_rank(wx_comm_t **comm, const unsigned char *id128, int rank, int world);
int wx_comm_init_all(wx_comm_t **comms, int ndev, const int *devlist);          /* devlist NULL: devices 0..ndev-1 */
int wx_comm_info(const wx_comm_t *comm, int *rank, int *world, int *dev, void **stream);
int wx_comm_destroy(wx_comm_t *comm);
int wx_group_start(void);
int wx_group_end(void);
int wx_allreduce(wx_comm_t *comm, void *buf, long count, int dtype, int op, void *stream);
int wx_allgather(wx_comm_t *comm, void *recv, const void *send, long count, int dtype, void *stream);   /* recv: world * count */
int wx_broadcast(wx_comm_t *comm, void *buf, long count, int dtype, int root, void *stream);
/* tree_costs(X, ::JBB | ::LSDB) of the GLOBAL batch from the local shard X(sz,K,Nlocal) (bestbasis/bestbasis_tree.jl:104-207):
 * local reduction kernels -> NCCL exchange -> per-node costs in costs_host (every rank gets the same vector).  comm NULL or a
 * 1-rank communicator: single GPU.  JBB exchanges one all-reduce of 2*sz*K+1 doubles; LSDB exchanges double-double sums by
 * all-gather (combined in rank order, so grid, counts and costs do not depend on the sharding), min / max / bin counts by
 * all-reduce.  Arguments as wx_jbb_costs / wx_lsdb_costs. */
int wx_tree_costs_jbb_f64(wx_comm_t *comm, double *costs_host, const double *X, long m, long n, int K, long Nlocal, int redundant, int cost_kind, double p, void *stream);
int wx_tree_costs_jbb_f32(wx_comm_t *comm, double *costs_host, const float *X, long m, long n, int K, long Nlocal, int redundant, int cost_kind, double p, void *stream);
int wx_tree_costs_lsdb_f64(wx_comm_t *comm, double *costs_host, const double *X, long m, long n, int K, long Nlocal, int redundant, void *stream);
int wx_tree_costs_lsdb_f32(wx_comm_t *comm, double *costs_host, const float *X, long m, long n, int K, long Nlocal, int redundant, void *stream);
/* bestbasistree(X, ::JBB | ::LSDB)  BestBasis.jl:185-217 in one call: costs as above, then bestbasis_treeselection (:min).
 * method 0 = JBB (cost_kind, p as wx_jbb_costs), 1 = LSDB.  tree_out: ntree bytes (n-1, or gettreelength(m,n) for images);
 * costs_host may be NULL, else it receives the node costs as tree_costs returns them (before the selection pass). */
int wx_bestbasistree_f64(wx_comm_t *comm, int method, unsigned char *tree_out, long ntree, double *costs_host, const double *X, long m, long n, int K, long Nlocal, int redundant, int cost_kind, double p, void *stream);
int wx_bestbasistree_f32(wx_comm_t *comm, int method, unsigned char *tree_out, long ntree, double *costs_host, const float *X, long m, long n, int K, long Nlocal, int redundant, int cost_kind, double p, void *stream);
/* the same from ONE host thread driving ndev devices: comms from wx_comm_init_all in rank order, X[i] / Nlocal[i] the shard on
 * comms[i]'s device; launches go to each communicator's own stream (wx_comm_info), NCCL calls are grouped. */
int wx_bestbasistree_multi_f64(wx_comm_t *const *comms, int ndev, int method, unsigned char *tree_out, long ntree, double *costs_host, const double *const *X, const long *Nlocal, long m, long n, int K, int redundant, int cost_kind, double p);
int wx_bestbasistree_multi_f32(wx_comm_t *const *comms, int ndev, int method, unsigned char *tree_out, long ntree, double *costs_host, const float *const *X, const long *Nlocal, long m, long n, int K, int redundant, int cost_kind, double p);

/* ---- host-buffer entry points (the call a drop-in user makes with host arrays) ----------------------- */
/* wpdall with x and y in HOST memory: chunks the batch, overlaps H2D / kernel / D2H on internal streams.
 * Synchronous.  chunk = signals per chunk (0 = auto). */
int wx_wpdall_host_f64(double *y_host, const double *x_host, long n, int L, long N, const double *h, const double *g, int F, long chunk);
int wx_wpdall_host_f32(float *y_host, const float *x_host, long n, int L, long N, const double *h, const double *g, int F, long chunk);
/* wpdall -> bestbasistree(JBB: method 0 | LSDB: method 1) -> getbasiscoefall with x and the best-basis coefficients in HOST
 * memory (the pipeline of paper/paper.md:60-118: dwt/dwt_all.jl:260-282, BestBasis.jl:185-217, Utils.jl:169-197).  The packet
 * table stays in HBM; x (n,N) goes up and coef (n,N) comes back in chunks overlapped with the kernels.  comm NULL = single GPU;
 * otherwise x_host is this rank's shard and the tree is that of the global batch.  tree_out: n-1 bytes.  Synchronous. */
int wx_wpd_bestbasis_host_f64(wx_comm_t *comm, double *coef_host, unsigned char *tree_out, long ntree, const double *x_host, long n, int L, long N, const double *h, const double *g, int F, int method, int cost_kind, double p, long chunk);
int wx_wpd_bestbasis_host_f32(wx_comm_t *comm, float *coef_host, unsigned char *tree_out, long ntree, const float *x_host, long n, int L, long N, const double *h, const double *g, int F, int method, int cost_kind, double p, long chunk);

#ifdef __cplusplus
}
#endif
#endif /* WX_B200_H */

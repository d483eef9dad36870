/* nngp_b200.h -- C ABI of the B200-native NNGP engine (libnngp_b200.so).
 *
 * Plain pointers and sizes only: no R types, no torch types.  This is the seam a maintainer of the
 * reference binds to; every entry point names the reference interface it replaces (paths relative
 * to the spNNGP 1.0.1 source root).  INTEGRATION.md shows the reference-side `.Call` glue.
 *
 * Conventions
 *   - All matrices are column-major (R layout).  `coords` is n x 2: x = coords[0..n), y = coords[n..2n).
 *   - Neighbour structure is the reference's ragged CSR (R/nn.R:29-40, src/util.cpp:55-70):
 *       nnIndx[nIndx] int32 neighbour ids (0-based, ascending distance per location),
 *       nnIndxLU[2n]: [0,n) offsets, [n,2n) counts, nIndx = (1+m)/2*m + (n-m-1)*m (src/rNNGP.cpp:218).
 *   - Covariance model codes follow R/spNNGP.R:172-178: 0 exponential, 1 spherical, 2 matern, 3 gaussian.
 *   - theta = (sigma.sq, tau.sq, phi[, nu]) in the reference's order (src/rNNGP.cpp:182-190).
 *   - Every function returns an int status: 0 ok; >0 = 1-based index of the first location whose
 *     m x m factorisation failed (the reference raises "dpotrf failed", src/rNNGP.cpp:86); <0 = NNGP_E*.
 *     nngp_last_error() gives the text.  The library never falls back to a CPU path: without a CUDA
 *     device every compute call returns NNGP_ENODEV.
 *   - One context drives one GPU.  Multi-GPU = one context per process/GPU, each given a shard
 *     [loc_begin, loc_end) of the locations; partial sums are combined by the caller (NCCL allreduce,
 *     see spnngp_b200/dist.py).  Pointers named *_dev are device pointers, all others host pointers.
 */
#ifndef NNGP_B200_H
#define NNGP_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NNGP_ABI_VERSION 1

enum { NNGP_COV_EXPONENTIAL = 0, NNGP_COV_SPHERICAL = 1, NNGP_COV_MATERN = 2, NNGP_COV_GAUSSIAN = 3 };

enum {
  NNGP_OK = 0,
  NNGP_EINVAL = -1,  /* bad argument (sizes, cov model, m > NNGP_MAX_M, ...) */
  NNGP_ENODEV = -2,  /* no usable CUDA device */
  NNGP_ECUDA = -3,   /* CUDA runtime error, see nngp_last_error */
  NNGP_ENOMEM = -4,
  NNGP_ESTATE = -5,  /* call order (e.g. likelihood before nngp_set_data) */
  NNGP_EINTR = -6    /* the caller's interrupt hook asked to stop */
};

#define NNGP_MAX_M 32     /* neighbours per location handled by the warp-group solver */
#define NNGP_MAX_BATCH 4  /* parameter sets evaluated per launch */

/* search_type, R/spNNGP.R:304-351 */
enum { NNGP_SEARCH_BRUTE = 0, NNGP_SEARCH_CB = 1 };

typedef struct nngp_ctx nngp_ctx;

/* Random-number hooks so the host sampler consumes the caller's stream (R's RNG under .Call:
 * rnorm/runif of Rmath, bracketed by GetRNGstate/PutRNGstate as in src/rNNGP.cpp:270,390). */
typedef double (*nngp_rnorm_fn)(double mu, double sd);
typedef double (*nngp_runif_fn)(double a, double b);
/* progress text hook (Rprintf); may be NULL */
typedef void (*nngp_print_fn)(const char *msg);
/* polled once per MCMC iteration / prediction chunk (R_CheckUserInterrupt); may be NULL */
typedef int (*nngp_interrupt_fn)(void);  /* polled once per iteration; nonzero: stop -- the call unwinds normally and returns NNGP_EINTR */

/* ---- lifecycle ------------------------------------------------------------------------------ */
int nngp_abi_version(void);
int nngp_device_count(void);
/* device < 0: use the current CUDA device.  stream_handle: a cudaStream_t to run on, or NULL for a
 * private non-blocking stream. */
int nngp_create(int device, void *stream_handle, nngp_ctx **out);
void nngp_destroy(nngp_ctx *ctx);
const char *nngp_last_error(const nngp_ctx *ctx); /* ctx may be NULL: last global error */
int nngp_sync(nngp_ctx *ctx);

/* ---- ordered neighbour search: replaces `mkNNIndx` (src/nn.cpp:63-111) and `mkNNIndxCB`
 *      (src/nn.cpp:214-257).  Outputs are written into the caller's arrays exactly like the reference
 *      (R/nn.R:45,71): nnIndx[nIndx], nnDist[nIndx], nnIndxLU[2n].  Neighbour candidates of location i
 *      are locations 0..i-1; ascending distance; distances use the reference's non-fused arithmetic
 *      (sqrt(dx*dx+dy*dy), products rounded separately) so nnDist matches bit for bit.  Every query
 *      runs the reference's own acceptance sequence (strict '<', overwrite-last-then-rsort_with_index)
 *      in one GPU thread, so distance ties resolve exactly as in the reference.  For search_type cb the
 *      u = x + y sort runs on the device; if u has exact ties the reference's serial Shell sort is
 *      replayed on the host instead (n_tie_replayed = number of tied adjacent keys that forced this). */
int nngp_mk_nn_indx(nngp_ctx *ctx, int n, int m, const double *coords, int search_type, int *nnIndx,
                    double *nnDist, int *nnIndxLU, int *n_tie_replayed);

/* `mkUIndx` (src/nn.cpp:25-56): the transposed neighbour graph the latent model walks.  uIndx (nIndx): for each
 * location i, ascending, the locations that have i as a neighbour; uIndxLU (2n): [0,n) offset of i's list, [n,2n) its
 * length; uiIndx (nIndx): the slot i occupies in each of those neighbour lists.  search_type as `u.search.type`
 * (R/spNNGP.R:99): 0 and 2 give ascending lists, 1 the same lists descending. */
int nngp_mk_u_indx(nngp_ctx *ctx, int n, int m, const int *nnIndx, int search_type, int *uIndx, int *uIndxLU,
                   int *uiIndx);

/* unordered query kNN (all n sites are candidates), ascending (d^2, index): the stand-in for
 * RANN::nn2(coords, coords.0, k=m)$nn.idx - 1 (R/generics.R:642, R/spConjNNGP.R:311).
 * nnIndx0 is q x m column-major, 0-based. */
int nngp_nn_query(nngp_ctx *ctx, int n, const double *coords, int q, const double *coords0, int m,
                  int *nnIndx0);

/* ---- model data: uploads what `rNNGP` / `cNNGP` read through REAL()/INTEGER()
 *      (src/rNNGP.cpp:120-128).  X is n x p, y length n.  [loc_begin, loc_end) is this context's
 *      shard of the locations (0, n for a single GPU).  nnIndxLU must be the closed form of
 *      src/util.cpp:55-70 (checked; may be NULL).  nnIndx may be NULL for prediction-only use
 *      (`rNNGPPredict` receives no fit neighbours): then only sites, X and y are uploaded. */
int nngp_set_data(nngp_ctx *ctx, int n, int m, int p, const double *coords, const int *nnIndx,
                  const int *nnIndxLU, const double *X, const double *y, int loc_begin, int loc_end);

/* ---- the hot path: B, F and the fused log-likelihood pieces.
 *      Replaces updateBF (src/rNNGP.cpp:46-101) + Q(B,F,r,r) (src/util.cpp:261-280) with
 *      r = X beta - y (src/rNNGP.cpp:299-303), for `nbatch` parameter sets in ONE launch.
 *      theta: nbatch x 4 row-major (sigma.sq, tau.sq, phi, nu); beta: nbatch x p row-major or NULL
 *      (then only logDet is formed).  B and F of set b are kept on the device in slot slots[b]
 *      (0 or 1) for later nngp_q_forms / nngp_get_bf calls.
 *      out: nbatch x 2 = (sum_i log F_i, Q) over this context's shard.
 *      The conjugate form updateConjBF (src/cNNGP.cpp:24-69) is theta = (1, alpha, phi, nu). */
int nngp_bf_loglik(nngp_ctx *ctx, int cov_model, int nbatch, const double *theta, const double *beta,
                   const int *slots, double *out);

/* The sampler's step in ONE launch: solve theta into slot_new as nngp_bf_loglik does, and with the same residual
 * r = X beta - y also return Q of the factors already resident in slot_cur (the current theta of the chain, whose
 * B and F do not change while only beta moves -- the work src/rNNGP.cpp:297-306 repeats every iteration).
 *      theta: 4 doubles; beta: p doubles; out: (sum_i log F_i(theta), Q(theta), Q(slot_cur)).
 * Kernels without the fused path (no resident distance triangle) run nngp_q_resid behind the launch. */
int nngp_bf_loglik_cur(nngp_ctx *ctx, int cov_model, const double *theta, const double *beta, int slot_new,
                       int slot_cur, double *out);

/* The same launch with the vector to whiten given directly (n doubles) instead of r = X beta - y: Q = u' C~^-1 u.
 * With tau.sq = 0 and u = w this is the latent model's likelihood piece -- its tau.sq-free updateBF plus the
 * OpenMP reductions over b_i^2 / F_i and log F_i (src/sNNGP.cpp:24-65, 323-400). */
int nngp_bf_loglik_vec(nngp_ctx *ctx, int cov_model, int nbatch, const double *theta, const double *u,
                       const int *slots, double *out);

/* Q(B,F,r,r) only, from the B,F already stored in `slot` (the double-buffered sampler uses this for
 * the unchanged current theta instead of re-solving, SURVEY §3.1 note).  out[0] = Q over the shard. */
int nngp_q_resid(nngp_ctx *ctx, int slot, const double *beta, double *out);

/* All cross quadratic forms of the columns [X, y] in one pass over B,F of `slot`:
 * S = [X y]' C~^-1 [X y], (p+1) x (p+1) column-major, full symmetric.  Replaces the p + p(p+1)/2 (+1)
 * separate Q(...) passes of src/rNNGP.cpp:281-286 and src/cNNGP.cpp:227-232,245. */
int nngp_q_forms(nngp_ctx *ctx, int slot, double *S);

/* copy B (nIndx) and F (n) of `slot` back to the host (parity tests; only this shard's range is
 * meaningful).  Either pointer may be NULL. */
int nngp_get_bf(nngp_ctx *ctx, int slot, double *B, double *F);

/* ---- whole entry points behind the reference's `.Call` symbols ------------------------------- */

/* `rNNGP` (src/rNNGP.cpp:105-420) without replicates: Metropolis-within-Gibbs for the response
 * model.  Draw order per iteration is the reference's: p rnorm (beta), rnorm for phi, sigma.sq,
 * tau.sq[, nu], one runif -- so a seeded chain retraces the reference's chain.
 * mode 0 = reference-faithful (two B/F solves per iteration), 1 = double-buffered (one solve plus a
 * Q-only pass; identical results).  Samples: betaSamples p x nSamples, thetaSamples nTheta x nSamples.
 * Uses the data given to nngp_set_data.  accept_out: number of accepted proposals. */
typedef struct {
  int cov_model;
  double sigmaSqIG[2], tauSqIG[2], phiUnif[2], nuUnif[2];
  const double *betaStarting;
  double sigmaSqStarting, tauSqStarting, phiStarting, nuStarting;
  double sigmaSqTuning, tauSqTuning, phiTuning, nuTuning;
  int nSamples;
  int mode;
  int verbose, nReport;
  nngp_rnorm_fn rnorm;
  nngp_runif_fn runif;
  nngp_print_fn print;
  nngp_interrupt_fn interrupt;
  /* optional hook that sums `len` doubles across the contexts/ranks of a sharded fit in place
   * (NCCL allreduce); NULL for a single GPU. */
  void (*allreduce)(double *buf, int len, void *user);
  void *allreduce_user;
  /* fit.rep (src/rNNGP.cpp:203-208,300-313): repIndx[s] != 0 marks the iterations that also draw a replicate
   * y.rep = X beta + v (n extra rnorm calls each, between the beta draw and the proposals); column k of
   * repSamples (n x nRep) receives the k-th marked iteration.  Both NULL: no replicates.  Needs the whole
   * location range on this context (the recurrence is serial in the ordering). */
  const int *repIndx;
  double *repSamples;
} nngp_rnngp_args;

int nngp_rnngp(nngp_ctx *ctx, const nngp_rnngp_args *args, double *betaSamples, double *thetaSamples,
               int *accept_out);

/* One replicated-data draw (updateRep, src/rNNGP.cpp:24-42): yrep = X beta + v with
 * v_i = B_i . v_N(i) + sqrt(F_i) z_i, B and F taken from slot `slot` (as left by nngp_bf_loglik), z = n standard
 * normals in location order.  yrep: n doubles. */
int nngp_rep_draw(nngp_ctx *ctx, int slot, const double *beta, const double *z, double *yrep);

/* `rNNGPReplicated` (src/rNNGPReplicated.cpp:76-222): one replicate per posterior sample.  betaSamples is
 * nSamples x p column-major (R/generics.R:252-259), thetaSamples nTheta x nSamples, repSamples n x nSamples.
 * rnorm is called n times per sample, in the reference's order. */
int nngp_rnngp_replicated(nngp_ctx *ctx, int cov_model, const double *betaSamples, const double *thetaSamples,
                          int nSamples, double (*rnorm)(double mu, double sd), nngp_interrupt_fn interrupt,
                          double *repSamples);

/* `cNNGP` (src/cNNGP.cpp:73-357): conjugate model for g parameter sets.  thetaAlpha is nTheta x g
 * (phi, alpha[, nu]).  Outputs as the reference: betaHat p x g, ab (written with stride 2, reference
 * quirk src/cNNGP.cpp:242-245; allocate nTheta x g), bBInv p*p x g, bb p x g, y0Hat / y0HatVar n0 x g.
 * X0 n0 x p, coords0 n0 x 2, nnIndx0 n0 x m column-major (may be NULL when n0 == 0). */
int nngp_cnngp(nngp_ctx *ctx, int cov_model, int g, const double *thetaAlpha, const double sigmaSqIG[2],
               int n0, const double *X0, const double *coords0, const int *nnIndx0, double *betaHat,
               double *ab, double *bBInv, double *bb, double *y0Hat, double *y0HatVar);

/* `cNNGP` on held-out sites, scored on the device (the k-fold body of R/spConjNNGP.R:322-334 with crps / rmspe of
 * R/util.R:247-255): the n0 x g predictions never leave the GPU.  y0: the n0 held-out responses.
 * crps[k] = mean(-sd (1/sqrt(pi) - 2 dnorm(z) - z (2 pnorm(z) - 1))), rmspe[k] = sqrt(mean((y0 - y0.hat)^2)), k < g. */
int nngp_cnngp_score(nngp_ctx *ctx, int cov_model, int g, const double *thetaAlpha, const double sigmaSqIG[2],
                     int n0, const double *X0, const double *coords0, const int *nnIndx0, const double *y0,
                     double *crps, double *rmspe);

/* The whole k-fold loop of spConjNNGP (R/spConjNNGP.R:286-349) in one call, for the parameter sets
 * [set_begin, set_end) of thetaAlpha (nTheta x g): per fold, drop the hold-out rows, rebuild the ordered neighbour
 * search on the rest (mkNNIndx / mkNNIndxCB, search_type), find the m nearest fitted sites of every held-out site
 * (RANN::nn2 -> nngp_nn_query), fit and score.  y, X (n x p), coords (n x 2) are in NNGP order (R/spConjNNGP.R:116-118).
 * hoIdx: k_fold x n_ho hold-out row ids, 0-based, ascending within a fold (R: sort(sample(1:n, floor(n/k.fold))) - 1,
 * drawn by the caller so R's RNG stream is the reference's).  scores: g x 2 column-major like `k.fold.scores`
 * (column 0 rmspe, column 1 crps), each the mean over the folds; rows outside [set_begin, set_end) are zero, so the
 * shards of a multi-GPU run (one context per GPU, disjoint set ranges) simply add.  Replaces the data of nngp_set_data. */
int nngp_conj_kfold(nngp_ctx *ctx, int n, int p, int m, const double *y, const double *X, const double *coords,
                    int cov_model, int g, const double *thetaAlpha, const double sigmaSqIG[2], int search_type,
                    int k_fold, int n_ho, const int *hoIdx, int set_begin, int set_end, double *scores);

/* `rNNGPPredict` (src/rNNGPPredict.cpp:25-222): posterior predictive draws at q sites.
 * betaSamples p x nSamples, thetaSamples nTheta x nSamples, nnIndx0 q x m column-major,
 * z: q*nSamples standard normals in the reference's single-thread order z[i*nSamples+s]
 * (src/rNNGPPredict.cpp:131-137,176-181); y0: q x nSamples.  Uses X, y, coords of nngp_set_data.
 * [site_begin, site_end) is this context's shard of the sites; rows outside it are left untouched. */
int nngp_rnngp_predict(nngp_ctx *ctx, int cov_model, int q, const double *X0, const double *coords0,
                       const int *nnIndx0, int nSamples, const double *betaSamples,
                       const double *thetaSamples, const double *z, double *y0, int site_begin,
                       int site_end);

/* ---- multi-GPU: one process (context) per GPU, the locations sharded with nngp_set_data's range.
 *      After nngp_comm_init the results of nngp_bf_loglik, nngp_q_resid and nngp_q_forms are summed
 *      over the ranks ON THE DEVICE (NCCL allreduce of <= 12 / (p+1)(p+2)/2 doubles over NVLink, issued
 *      on the context stream right behind the producing kernel), so every rank receives the full
 *      log-likelihood pieces and nngp_rnngp / nngp_cnngp run unchanged and in lock-step on every rank
 *      (same RNG seed on every rank).  The id comes from rank 0 (nngp_comm_unique_id) and is
 *      distributed by the launcher (torch.distributed broadcast in spnngp_b200/dist.py, MPI, files...).
 *      libnccl is loaded lazily; single-GPU use needs no NCCL. */
int nngp_comm_unique_id(char id[128]);
int nngp_comm_init(nngp_ctx *ctx, const char id[128], int rank, int nranks);
int nngp_comm_allreduce_sum(nngp_ctx *ctx, double *buf, int len); /* host buffer, len <= 12; for tests/tools */
/* The same sums for ranks that share one node WITHOUT NCCL: every rank's kernels store their partials, then a sequence
 * stamp, into that rank's slot of a shared-memory mailbox; every host polls all slots and adds them in rank order
 * (deterministic; no collective call, no device-to-host copy, no stream synchronisation on the sampler's critical path).
 * nngp_comm_init sets the same mailbox up next to its NCCL communicator and uses it unless
 * nngp_set_option("reduce_nccl", 1) / NNGP_REDUCE=nccl asks for ncclAllReduce.  key: 128 bytes, identical on all ranks
 * of the job and unique to it (nngp_comm_unique_id provides one).  Ranks may share a GPU. */
int nngp_comm_init_local(nngp_ctx *ctx, const char key[128], int rank, int nranks);
void nngp_comm_destroy(nngp_ctx *ctx);

/* ---- HBM-resident measurement helpers (bench.py `value`): time the likelihood launch with every
 *      input already on the device.  Runs `iters` launches of nngp_bf_loglik's kernel back to back
 *      on the context stream between two CUDA events and returns the mean milliseconds per launch. */
int nngp_time_bf_loglik(nngp_ctx *ctx, int cov_model, int nbatch, const double *theta, const double *beta,
                        int iters, int flush_l2, float *ms_per_launch);
/* FP64 FMA throughput of this device measured with a DFMA micro-benchmark (TFLOP/s, FMA = 2 flops):
 * the roofline denominator for the B/F kernels (MEASURED_PEAKS.json has no FP64 figure). */
int nngp_fp64_peak(nngp_ctx *ctx, double *tflops);
/* Page-locked host memory (cudaHostAlloc).  Optional: every entry point takes ordinary pageable arrays (R vectors) and
 * stages them through an internal pinned ring at the host's memcpy rate; arrays that ARE page-locked -- from this call,
 * or cudaHostRegister'ed by the caller -- are recognised (cudaPointerGetAttributes) and moved by direct DMA at the
 * PCIe rate.  Matters for the q x nSamples streams of nngp_rnngp_predict (src/rNNGPPredict.cpp:131-137,176-181). */
int nngp_host_alloc(nngp_ctx *ctx, size_t bytes, void **out);
int nngp_host_free(nngp_ctx *ctx, void *p);
/* Tuning switches (tests and A/B measurements; defaults are the fast settings).
 *   "matern_short" (default 2): let Matern launches use a truncated series when the data set's neighbourhoods are
 *   dense enough for it (dropped terms < 1e-16, see csrc/matern_series.cuh): 2 = the 5-term series where it holds
 *   (distance-triangle kernel), else the 9-term one; 1 = at most the 9-term series; 0 forces the 14-term series. */
int nngp_set_option(nngp_ctx *ctx, const char *name, int value);
/* reads an option back; "matern_short_used": how many parameter sets of the last likelihood launch ran a truncated
 * series, "matern_tiny_used": how many of them the 5-term one */
int nngp_get_option(nngp_ctx *ctx, const char *name, int *value);
/* number of kernel launches issued by this context so far (bench.py `gpu_launches`) */
long long nngp_launch_count(const nngp_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* NNGP_B200_H */

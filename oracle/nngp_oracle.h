/* oracle/nngp_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the spNNGP 1.0.1 hot path (response / conjugate NNGP): neighbour
 * index layout, the four correlation functions, updateBF / updateConjBF (explicit-inverse route,
 * like the reference), the quadratic form Q, the ordered kNN searches (brute and code-book), the
 * response-model Metropolis sampler, the conjugate per-set fit and the kNN+solve prediction.
 * Each function cites the reference file:line it follows (paths relative to the reference root).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
 * load this library; the product path (spnngp_b200/) never does.
 *
 * Parity pinning: the reference ships no tests or golden vectors, so this restatement
 * is pinned against outputs of the reference's own C++ compiled here (oracle/_ref, see Makefile)
 * -- tests/test_oracle_vs_ref.py -- and against the committed fixtures in tests/golden/ that were
 * generated from oracle/_ref by tests/golden/make_golden.py.  Third-party pieces that are not in
 * the reference tree (R's rsort_with_index, Rmath bessel_k/gammafn, LAPACK dpotrf/dpotri, BLAS)
 * are our restatements of their published algorithms on BOTH sides; parity at those boundaries
 * is unpinned (no R on this box).
 */
#ifndef NNGP_ORACLE_H
#define NNGP_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

/* covariance model codes, R/spNNGP.R:172-178 */
enum { ORC_EXPONENTIAL = 0, ORC_SPHERICAL = 1, ORC_MATERN = 2, ORC_GAUSSIAN = 3 };

typedef double (*orc_rnorm_fn)(double mu, double sd);
typedef double (*orc_runif_fn)(double a, double b);

void orc_set_threads(int n);
int orc_max_threads(void);

/* src/util.cpp:55-70 */
void orc_get_nn_indx(int i, int m, int *offset, int *count);
/* nIndx = (1+m)/2*m + (n-m-1)*m, src/rNNGP.cpp:218 */
long orc_n_indx(int n, int m);
/* src/util.cpp:51-53 */
double orc_dist2(double a1, double a2, double b1, double b2);
/* src/util.cpp:208-245 */
double orc_sp_cor(double D, double phi, double nu, int cov_model);
/* K_nu(x), x>0: trapezoid quadrature of int_0^inf exp(-x cosh t) cosh(nu t) dt (independent of
 * the device algorithm and of libstdc++'s cyl_bessel_k). */
double orc_bessel_k(double x, double nu);
/* R src/main/sort.c rsort_with_index (third-party; restated) */
void orc_rsort_with_index(double *x, int *indx, int n);

/* src/rNNGP.cpp:46-101.  theta = (sigmaSq, tauSq, phi[, nu]).  Returns sum_i log F_i.
 * info_out (may be NULL): 0, or 1-based index of the first location whose factorisation failed. */
double orc_update_bf(double *B, double *F, const double *coords, const int *nnIndx, const int *nnIndxLU,
                     int n, int m, const double *theta, int cov_model, int *info_out);
/* src/cNNGP.cpp:24-69 */
void orc_update_conj_bf(double *B, double *F, const double *coords, const int *nnIndx, const int *nnIndxLU,
                        int n, int m, double phi, double alpha, double nu, int cov_model, int *info_out);
/* src/util.cpp:261-280 */
double orc_q(const double *B, const double *F, const double *u, const double *v, int n, const int *nnIndx,
             const int *nnIndxLU);

/* src/nn.cpp:63-111 (brute) and :136-257 (code-book).  Outputs are filled in place like the reference. */
void orc_mk_nn_indx(int n, int m, const double *coords, int *nnIndx, double *nnDist, int *nnIndxLU);
void orc_mk_nn_indx_cb(int n, int m, const double *coords, int *nnIndx, double *nnDist, int *nnIndxLU);
/* exact unordered kNN of q query sites among n sites, ascending (d^2, index): stands in for
 * RANN::nn2 (third party, R/generics.R:642); 0-based, column-major q x m like nn2$nn.idx-1 */
void orc_nn_query(int n, const double *coords, int q, const double *coords0, int m, int *nnIndx0);

/* src/rNNGP.cpp:105-420 without replicates.  Samples are written p x nSamples / nTheta x nSamples.
 * Returns the number of accepted proposals, or -1 on a factorisation failure. */
int orc_rnngp(const double *y, const double *X, int p, int n, int m, const double *coords, int cov_model,
              const int *nnIndx, const int *nnIndxLU, const double *sigmaSqIG, const double *tauSqIG,
              const double *phiUnif, const double *nuUnif, const double *betaStarting, double sigmaSqStarting,
              double tauSqStarting, double phiStarting, double nuStarting, double sigmaSqTuning,
              double tauSqTuning, double phiTuning, double nuTuning, int nSamples, orc_rnorm_fn rnorm,
              orc_runif_fn runif, double *betaSamples, double *thetaSamples);

/* mkUIndx (src/nn.cpp:25-56): transposed neighbour graph. */
void orc_mk_u_indx(int n, int m, const int *nnIndx, const int *nnIndxLU, int search_type, int *uIndx, int *uIndxLU,
                   int *uiIndx);

/* fit.rep / rNNGPReplicated.  orc_update_rep: src/rNNGP.cpp:24-42; orc_rnngp_rep: orc_rnngp plus the
 * replicate draws of src/rNNGP.cpp:300-313 (repIndx per iteration, repSamples n x nRep; NULL, NULL = orc_rnngp);
 * orc_rnngp_replicated: src/rNNGPReplicated.cpp:183-201 (beta nSamples x p column-major, theta nTheta x nSamples). */
void orc_update_rep(const double *B, const double *F, int n, const int *nnIndx, const int *nnIndxLU, orc_rnorm_fn rnorm,
                    double *v);
int orc_rnngp_rep(const double *y, const double *X, int p, int n, int m, const double *coords, int cov_model,
                  const int *nnIndx, const int *nnIndxLU, const double *sigmaSqIG, const double *tauSqIG,
                  const double *phiUnif, const double *nuUnif, const double *betaStarting, double sigmaSqStarting,
                  double tauSqStarting, double phiStarting, double nuStarting, double sigmaSqTuning,
                  double tauSqTuning, double phiTuning, double nuTuning, int nSamples, orc_rnorm_fn rnorm,
                  orc_runif_fn runif, double *betaSamples, double *thetaSamples, const int *repIndx,
                  double *repSamples);
int orc_rnngp_replicated(const double *X, int p, int n, int m, const double *coords, int cov_model, const int *nnIndx,
                         const int *nnIndxLU, const double *beta, const double *theta, int nSamples,
                         orc_rnorm_fn rnorm, double *repSamples);

/* src/cNNGP.cpp:215-316 for g parameter sets; thetaAlpha is nTheta x g (phi, alpha[, nu]).
 * ab is written with stride 2 per set (reference quirk, :242-245).  y0Hat/y0HatVar may be NULL if n0 == 0. */
int orc_cnngp(const double *y, const double *X, int p, int n, const double *coords, const double *thetaAlpha,
              const int *nnIndx, const int *nnIndxLU, const double *X0, const double *coords0, int n0,
              const int *nnIndx0, int g, int m, const double *sigmaSqIG, int cov_model, double *betaHat,
              double *ab, double *bBInv, double *bb, double *y0Hat, double *y0HatVar);

/* src/rNNGPPredict.cpp:139-183 with the single-thread z assignment z[i*nSamples+s].
 * y0 is q x nSamples.  site_begin/site_end restrict the work to a slice of sites (for timing). */
int orc_rnngp_predict(const double *X, const double *y, const double *coords, int n, int p, int m,
                      const double *X0, const double *coords0, int q, const int *nnIndx0,
                      const double *betaSamples, const double *thetaSamples, int nSamples, int cov_model,
                      const double *z, double *y0, int site_begin, int site_end);

#ifdef __cplusplus
}
#endif
#endif

// spnngp_glue.cpp -- the reference's `.Call` entry points, re-implemented as thin SEXP wrappers over
// the C ABI of the B200 engine (include/nngp_b200.h).  Same symbol names, argument order, in-place
// mutation and result lists as spNNGP 1.0.1:
//   rNNGP(28 args)         src/spNNGP.h:9-13,   src/rNNGP.cpp:105-420
//   cNNGP(18 args)         src/spNNGP.h:21-23,  src/cNNGP.cpp:73-357
//   rNNGPPredict(17 args)  src/spNNGP.h:36-37,  src/rNNGPPredict.cpp:25-222
//   mkNNIndx(7 args)       src/nn.h:25,         src/nn.cpp:63-111
//   mkNNIndxCB(7 args)     src/nn.h:38,         src/nn.cpp:214-257
// The other five symbols of the package (latent / logit / SLGP models) are outside the
// hot path this engine accelerates; they are exported as stubs that raise an R error so that a package
// built with this glue still loads.
//
// Builds against real R headers (R CMD SHLIB) or against spnngp_b200/rshim on hosts
// without R.  R's RNG is consumed in the reference's order through Rmath's rnorm/runif, bracketed by
// GetRNGstate/PutRNGstate, so seeded chains retrace the reference's chain.
#ifndef R_NO_REMAP
#define R_NO_REMAP
#endif
#include <R.h>
#include <Rinternals.h>
#include <Rmath.h>
#include <R_ext/Rdynload.h>
#include <R_ext/Utils.h>

#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/nngp_b200.h"

namespace {

nngp_ctx *g_ctx = nullptr;

nngp_ctx *ctx() {
  if (!g_ctx) {
    const char *dev = getenv("NNGP_DEVICE");
    const int rc = nngp_create(dev ? atoi(dev) : -1, nullptr, &g_ctx);
    if (rc != NNGP_OK) Rf_error("spNNGP (B200 engine): %s", nngp_last_error(nullptr));
  }
  return g_ctx;
}

void check(int rc) {
  if (rc == NNGP_OK) return;
  if (rc > 0) Rf_error("c++ Rf_error: dpotrf failed\n");  // same text as src/rNNGP.cpp:86
  Rf_error("spNNGP (B200 engine): %s", nngp_last_error(g_ctx));
}

double rnorm_hook(double mu, double sd) { return rnorm(mu, sd); }
double runif_hook(double a, double b) { return runif(a, b); }
void print_hook(const char *msg) { Rprintf("%s", msg); }
// R_CheckUserInterrupt longjmps: called from inside the engine it would skip the destructors of live C++ frames
// (host vectors, pooled device buffers) and PutRNGstate.  R_ToplevelExec contains the jump; the hook only reports it,
// the engine unwinds normally with NNGP_EINTR, and the glue raises the R error after its own clean-up.
void check_interrupt_fn(void *) { R_CheckUserInterrupt(); }
int interrupt_hook(void) { return R_ToplevelExec(check_interrupt_fn, nullptr) ? 0 : 1; }

int sampler_mode() {
  // 1 (default): double-buffered -- B, F of the current theta stay resident, one launch per iteration solves the
  // candidate and re-forms Q of the current theta; 0: reference-faithful -- both solved anew every iteration
  // (src/rNNGP.cpp:297-339).  Identical chains either way (tests/test_gpu_golden_and_glue.py).
  const char *e = getenv("NNGP_SAMPLER_MODE");
  return (e && e[0] == '0') ? 0 : 1;
}

SEXP named_list(const std::vector<const char *> &names, const std::vector<SEXP> &vals) {
  const int k = (int)names.size();
  SEXP result_r = PROTECT(Rf_allocVector(VECSXP, k));
  SEXP resultName_r = PROTECT(Rf_allocVector(VECSXP, k));  // a VECSXP of CHARSXP, as the reference builds it
  for (int i = 0; i < k; i++) {
    SET_VECTOR_ELT(result_r, i, vals[i]);
    SET_VECTOR_ELT(resultName_r, i, Rf_mkChar(names[i]));
  }
  Rf_namesgets(result_r, resultName_r);
  UNPROTECT(2);
  return result_r;
}

bool is_matern(int covModel) { return covModel == NNGP_COV_MATERN; }

void check_cov(int covModel) {
  if (covModel < 0 || covModel > 3) Rf_error("c++ Rf_error: cov.model is not correctly specified");
}

SEXP not_provided(const char *name) {
  Rf_error("'%s' is outside the hot path accelerated by the B200 engine; use the CPU build of spNNGP for it", name);
  return R_NilValue;
}

}  // namespace

extern "C" {

SEXP rNNGP(SEXP y_r, SEXP X_r, SEXP p_r, SEXP n_r, SEXP m_r, SEXP coords_r, SEXP covModel_r, SEXP nnIndx_r,
           SEXP nnIndxLU_r, SEXP sigmaSqIG_r, SEXP tauSqIG_r, SEXP phiUnif_r, SEXP nuUnif_r, SEXP betaStarting_r,
           SEXP sigmaSqStarting_r, SEXP tauSqStarting_r, SEXP phiStarting_r, SEXP nuStarting_r, SEXP sigmaSqTuning_r,
           SEXP tauSqTuning_r, SEXP phiTuning_r, SEXP nuTuning_r, SEXP nSamples_r, SEXP nThreads_r, SEXP verbose_r,
           SEXP nReport_r, SEXP nRep_r, SEXP repIndx_r) {
  (void)nThreads_r;
  const int p = INTEGER(p_r)[0], n = INTEGER(n_r)[0], m = INTEGER(m_r)[0];
  const int covModel = INTEGER(covModel_r)[0];
  check_cov(covModel);
  const int nRep = INTEGER(nRep_r)[0];  // fit.rep: number of marked iterations (src/rNNGP.cpp:139-140)
  const int nSamples = INTEGER(nSamples_r)[0];
  const int nTheta = is_matern(covModel) ? 4 : 3;

  nngp_ctx *c = ctx();
  check(nngp_set_data(c, n, m, p, REAL(coords_r), INTEGER(nnIndx_r), INTEGER(nnIndxLU_r), REAL(X_r), REAL(y_r), 0, n));

  nngp_rnngp_args a;
  memset(&a, 0, sizeof a);
  a.cov_model = covModel;
  a.sigmaSqIG[0] = REAL(sigmaSqIG_r)[0]; a.sigmaSqIG[1] = REAL(sigmaSqIG_r)[1];
  a.tauSqIG[0] = REAL(tauSqIG_r)[0]; a.tauSqIG[1] = REAL(tauSqIG_r)[1];
  a.phiUnif[0] = REAL(phiUnif_r)[0]; a.phiUnif[1] = REAL(phiUnif_r)[1];
  if (is_matern(covModel)) {
    a.nuUnif[0] = REAL(nuUnif_r)[0]; a.nuUnif[1] = REAL(nuUnif_r)[1];
    a.nuStarting = REAL(nuStarting_r)[0];
    a.nuTuning = REAL(nuTuning_r)[0];
  }
  a.betaStarting = REAL(betaStarting_r);
  a.sigmaSqStarting = REAL(sigmaSqStarting_r)[0];
  a.tauSqStarting = REAL(tauSqStarting_r)[0];
  a.phiStarting = REAL(phiStarting_r)[0];
  a.sigmaSqTuning = REAL(sigmaSqTuning_r)[0];
  a.tauSqTuning = REAL(tauSqTuning_r)[0];
  a.phiTuning = REAL(phiTuning_r)[0];
  a.nSamples = nSamples;
  a.mode = sampler_mode();
  a.verbose = INTEGER(verbose_r)[0];
  a.nReport = INTEGER(nReport_r)[0];
  a.rnorm = rnorm_hook;
  a.runif = runif_hook;
  a.print = print_hook;
  a.interrupt = interrupt_hook;

  SEXP betaSamples_r = PROTECT(Rf_allocMatrix(REALSXP, p, nSamples));
  SEXP thetaSamples_r = PROTECT(Rf_allocMatrix(REALSXP, nTheta, nSamples));
  SEXP repSamples_r = R_NilValue;
  if (nRep > 0) {
    repSamples_r = PROTECT(Rf_allocMatrix(REALSXP, n, nRep));  // src/rNNGP.cpp:203-208
    a.repIndx = INTEGER(repIndx_r);
    a.repSamples = REAL(repSamples_r);
  }
  GetRNGstate();
  int accept = 0;
  const int rc = nngp_rnngp(c, &a, REAL(betaSamples_r), REAL(thetaSamples_r), &accept);
  PutRNGstate();
  check(rc);
  SEXP out = nRep > 0 ? named_list({"p.beta.samples", "p.theta.samples", "y.rep.samples"},
                                   {betaSamples_r, thetaSamples_r, repSamples_r})
                      : named_list({"p.beta.samples", "p.theta.samples"}, {betaSamples_r, thetaSamples_r});
  UNPROTECT(nRep > 0 ? 3 : 2);
  return out;
}

SEXP cNNGP(SEXP y_r, SEXP X_r, SEXP p_r, SEXP n_r, SEXP coords_r, SEXP thetaAlpha_r, SEXP nnIndx_r, SEXP nnIndxLU_r,
           SEXP X0_r, SEXP coords0_r, SEXP n0_r, SEXP nnIndx0_r, SEXP g_r, SEXP m_r, SEXP sigmaSqIG_r, SEXP covModel_r,
           SEXP nThreads_r, SEXP verbose_r) {
  (void)nThreads_r;
  const int p = INTEGER(p_r)[0], n = INTEGER(n_r)[0], m = INTEGER(m_r)[0], n0 = INTEGER(n0_r)[0], g = INTEGER(g_r)[0];
  const int covModel = INTEGER(covModel_r)[0];
  check_cov(covModel);
  const int nTheta = is_matern(covModel) ? 3 : 2;
  if (INTEGER(verbose_r)[0]) {
    Rprintf("----------------------------------------\n\tModel description\n----------------------------------------\n");
    Rprintf("NNGP Conjugate model fit with %i observations.\n\n", n);
    Rprintf("Number of covariates %i (including intercept if specified).\n\n", p);
    Rprintf("Using %i nearest neighbors.\n\n", m);
    if (n0 > 0) Rprintf("------------\nPredicting at %i locations.\n", n0);
  }
  nngp_ctx *c = ctx();
  check(nngp_set_data(c, n, m, p, REAL(coords_r), INTEGER(nnIndx_r), INTEGER(nnIndxLU_r), REAL(X_r), REAL(y_r), 0, n));
  SEXP beta_r = PROTECT(Rf_allocMatrix(REALSXP, p, g));
  SEXP ab_r = PROTECT(Rf_allocMatrix(REALSXP, nTheta, g));
  SEXP bBInv_r = PROTECT(Rf_allocMatrix(REALSXP, p * p, g));
  SEXP bb_r = PROTECT(Rf_allocMatrix(REALSXP, p, g));
  SEXP y0Hat_r = PROTECT(Rf_allocMatrix(REALSXP, n0, g));
  SEXP y0HatVar_r = PROTECT(Rf_allocMatrix(REALSXP, n0, g));
  check(nngp_cnngp(c, covModel, g, REAL(thetaAlpha_r), REAL(sigmaSqIG_r), n0, n0 > 0 ? REAL(X0_r) : nullptr,
                   n0 > 0 ? REAL(coords0_r) : nullptr, n0 > 0 ? INTEGER(nnIndx0_r) : nullptr, REAL(beta_r), REAL(ab_r),
                   REAL(bBInv_r), REAL(bb_r), REAL(y0Hat_r), REAL(y0HatVar_r)));
  SEXP out;
  if (n0 > 0)
    out = named_list({"beta.hat", "ab", "bB.inv", "bb", "y.0.hat", "y.0.hat.var"},
                     {beta_r, ab_r, bBInv_r, bb_r, y0Hat_r, y0HatVar_r});
  else
    out = named_list({"beta.hat", "ab", "bB.inv", "bb"}, {beta_r, ab_r, bBInv_r, bb_r});
  UNPROTECT(6);
  return out;
}

// NEW symbol (not in the reference): the k-fold loop of R/spConjNNGP.R:286-349 in one call, scores formed on the GPU.
// INTEGRATION.md shows the three-line change in spConjNNGP() that uses it; without that change the reference's R loop
// keeps calling cNNGP per fold and everything still works.  hoIdx_r: n.ho x k.fold integer matrix, 1-based, each
// column sort(sample(1:n, floor(n/k.fold))) drawn in R; searchType_r: 0 brute, 1 cb.
// Returns the g x 2 matrix k.fold.scores/k.fold (columns rmspe, crps).
SEXP cNNGPKFold(SEXP y_r, SEXP X_r, SEXP p_r, SEXP n_r, SEXP coords_r, SEXP thetaAlpha_r, SEXP g_r, SEXP m_r,
                SEXP sigmaSqIG_r, SEXP covModel_r, SEXP searchType_r, SEXP hoIdx_r, SEXP nHo_r, SEXP kFold_r) {
  const int p = INTEGER(p_r)[0], n = INTEGER(n_r)[0], m = INTEGER(m_r)[0], g = INTEGER(g_r)[0];
  const int covModel = INTEGER(covModel_r)[0], nHo = INTEGER(nHo_r)[0], kFold = INTEGER(kFold_r)[0];
  check_cov(covModel);
  std::vector<int> ho((size_t)nHo * kFold);
  for (size_t i = 0; i < ho.size(); i++) ho[i] = INTEGER(hoIdx_r)[i] - 1;
  SEXP scores_r = PROTECT(Rf_allocMatrix(REALSXP, g, 2));
  check(nngp_conj_kfold(ctx(), n, p, m, REAL(y_r), REAL(X_r), REAL(coords_r), covModel, g, REAL(thetaAlpha_r),
                        REAL(sigmaSqIG_r), INTEGER(searchType_r)[0], kFold, nHo, ho.data(), 0, g, REAL(scores_r)));
  UNPROTECT(1);
  return scores_r;
}

SEXP rNNGPPredict(SEXP X_r, SEXP y_r, SEXP coords_r, SEXP n_r, SEXP p_r, SEXP m_r, SEXP X0_r, SEXP coords0_r, SEXP q_r,
                  SEXP nnIndx0_r, SEXP betaSamples_r, SEXP thetaSamples_r, SEXP nSamples_r, SEXP covModel_r,
                  SEXP nThreads_r, SEXP verbose_r, SEXP nReport_r) {
  (void)nThreads_r;
  (void)nReport_r;
  const int n = INTEGER(n_r)[0], p = INTEGER(p_r)[0], m = INTEGER(m_r)[0], q = INTEGER(q_r)[0];
  const int nSamples = INTEGER(nSamples_r)[0], covModel = INTEGER(covModel_r)[0];
  check_cov(covModel);
  if (INTEGER(verbose_r)[0]) {
    Rprintf("----------------------------------------\n\tPrediction description\n----------------------------------------\n");
    Rprintf("NNGP Response model fit with %i observations.\n\n", n);
    Rprintf("Number of MCMC samples %i.\n\nPredicting at %i locations.\n\n", nSamples, q);
  }
  nngp_ctx *c = ctx();
  // prediction needs the fitted sites, X and y, not the fit's neighbour structure
  check(nngp_set_data(c, n, m, p, REAL(coords_r), nullptr, nullptr, REAL(X_r), REAL(y_r), 0, n));
  // all q*nSamples normals up front, in the reference's order (src/rNNGPPredict.cpp:131-137)
  double *z = (double *)R_alloc((size_t)q * nSamples, sizeof(double));
  GetRNGstate();
  for (size_t i = 0; i < (size_t)q * nSamples; i++) z[i] = rnorm(0.0, 1.0);
  PutRNGstate();
  SEXP y0_r = PROTECT(Rf_allocMatrix(REALSXP, q, nSamples));
  check(nngp_rnngp_predict(c, covModel, q, REAL(X0_r), REAL(coords0_r), INTEGER(nnIndx0_r), nSamples,
                           REAL(betaSamples_r), REAL(thetaSamples_r), z, REAL(y0_r), 0, q));
  SEXP out = named_list({"p.y.0"}, {y0_r});
  UNPROTECT(1);
  return out;
}

static SEXP mk_nn(SEXP n_r, SEXP m_r, SEXP coords_r, SEXP nnIndx_r, SEXP nnDist_r, SEXP nnIndxLU_r, int search) {
  int ties = 0;
  check(nngp_mk_nn_indx(ctx(), INTEGER(n_r)[0], INTEGER(m_r)[0], REAL(coords_r), search, INTEGER(nnIndx_r),
                        REAL(nnDist_r), INTEGER(nnIndxLU_r), &ties));
  return R_NilValue;  // outputs are written into the arguments, like the reference (R/nn.R:45,71)
}

SEXP mkNNIndx(SEXP n_r, SEXP m_r, SEXP coords_r, SEXP nnIndx_r, SEXP nnDist_r, SEXP nnIndxLU_r, SEXP nThreads_r) {
  (void)nThreads_r;
  return mk_nn(n_r, m_r, coords_r, nnIndx_r, nnDist_r, nnIndxLU_r, NNGP_SEARCH_BRUTE);
}

SEXP mkNNIndxCB(SEXP n_r, SEXP m_r, SEXP coords_r, SEXP nnIndx_r, SEXP nnDist_r, SEXP nnIndxLU_r, SEXP nThreads_r) {
  (void)nThreads_r;
  return mk_nn(n_r, m_r, coords_r, nnIndx_r, nnDist_r, nnIndxLU_r, NNGP_SEARCH_CB);
}

// ---- symbols outside the accelerated hot path: present so the package loads, error when called ----------
SEXP sNNGP(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP,
           SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP) { return not_provided("sNNGP"); }
SEXP sNNGPLogit(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP,
                SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP) { return not_provided("sNNGPLogit"); }
SEXP cSLGP(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP,
           SEXP, SEXP, SEXP) { return not_provided("cSLGP"); }
SEXP sNNGPPredict(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP,
                  SEXP, SEXP) { return not_provided("sNNGPPredict"); }
SEXP PGLogit(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP) { return not_provided("PGLogit"); }
// src/rNNGPReplicated.cpp:76-222: one replicate of the data per posterior sample
SEXP rNNGPReplicated(SEXP X_r, SEXP p_r, SEXP n_r, SEXP m_r, SEXP coords_r, SEXP covModel_r, SEXP nnIndx_r,
                     SEXP nnIndxLU_r, SEXP beta_r, SEXP theta_r, SEXP nSamples_r, SEXP nThreads_r, SEXP verbose_r,
                     SEXP nReport_r) {
  (void)nThreads_r;
  (void)nReport_r;
  const int p = INTEGER(p_r)[0], n = INTEGER(n_r)[0], m = INTEGER(m_r)[0], nSamples = INTEGER(nSamples_r)[0];
  const int covModel = INTEGER(covModel_r)[0];
  check_cov(covModel);
  if (INTEGER(verbose_r)[0]) {
    Rprintf("----------------------------------------\n\tComputing replicates\n----------------------------------------\n");
    Rprintf("NNGP Response model fit with %i observations.\n\n", n);
    Rprintf("Number of covariates %i (including intercept if specified).\n\n", p);
    Rprintf("Using %i nearest neighbors.\n\n", m);
    Rprintf("Number of MCMC samples %i.\n\n", nSamples);
  }
  nngp_ctx *c = ctx();
  double *y0 = (double *)R_alloc(n, sizeof(double));  // the response does not enter the replicates
  for (int i = 0; i < n; i++) y0[i] = 0.0;
  check(nngp_set_data(c, n, m, p, REAL(coords_r), INTEGER(nnIndx_r), INTEGER(nnIndxLU_r), REAL(X_r), y0, 0, n));
  SEXP repSamples_r = PROTECT(Rf_allocMatrix(REALSXP, n, nSamples));
  GetRNGstate();
  const int rc = nngp_rnngp_replicated(c, covModel, REAL(beta_r), REAL(theta_r), nSamples, rnorm_hook, interrupt_hook,
                                       REAL(repSamples_r));
  PutRNGstate();
  check(rc);
  SEXP out = named_list({"y.rep.samples"}, {repSamples_r});
  UNPROTECT(1);
  return out;
}
// src/nn.cpp:25-56: writes uIndx, uIndxLU, uiIndx in place (R/nn.R:19-23)
SEXP mkUIndx(SEXP n_r, SEXP m_r, SEXP nnIndx_r, SEXP uIndx_r, SEXP uIndxLU_r, SEXP uiIndx_r, SEXP nnIndxLU_r,
             SEXP searchType_r) {
  (void)nnIndxLU_r;
  check(nngp_mk_u_indx(ctx(), INTEGER(n_r)[0], INTEGER(m_r)[0], INTEGER(nnIndx_r), INTEGER(searchType_r)[0],
                       INTEGER(uIndx_r), INTEGER(uIndxLU_r), INTEGER(uiIndx_r)));
  return R_NilValue;
}

// registration of all 12 reference symbols (+ cNNGPKFold) (the reference's init.cpp registers 9 and is never run)
static const R_CallMethodDef CallEntries[] = {
    {"rNNGP", (DL_FUNC)&rNNGP, 28},          {"sNNGP", (DL_FUNC)&sNNGP, 29},
    {"sNNGPLogit", (DL_FUNC)&sNNGPLogit, 27}, {"cNNGP", (DL_FUNC)&cNNGP, 18},
    {"cSLGP", (DL_FUNC)&cSLGP, 21},          {"rNNGPPredict", (DL_FUNC)&rNNGPPredict, 17},
    {"sNNGPPredict", (DL_FUNC)&sNNGPPredict, 19}, {"PGLogit", (DL_FUNC)&PGLogit, 9},
    {"rNNGPReplicated", (DL_FUNC)&rNNGPReplicated, 14}, {"mkNNIndx", (DL_FUNC)&mkNNIndx, 7},
    {"mkNNIndxCB", (DL_FUNC)&mkNNIndxCB, 7}, {"mkUIndx", (DL_FUNC)&mkUIndx, 8},
    {"cNNGPKFold", (DL_FUNC)&cNNGPKFold, 14},  // addition, see above
    {NULL, NULL, 0}};

void R_init_spNNGP(DllInfo *dll) {
  R_registerRoutines(dll, NULL, CallEntries, NULL, NULL);
  R_useDynamicSymbols(dll, (Rboolean)1);  // the R layer calls by string, keep dynamic lookup working
}

void R_unload_spNNGP(DllInfo *) {
  if (g_ctx) nngp_destroy(g_ctx);
  g_ctx = nullptr;
}

}  // extern "C"

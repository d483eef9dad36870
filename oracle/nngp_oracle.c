/* oracle/nngp_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See nngp_oracle.h.
 *
 * CPU restatement (plain C + OpenMP over locations, like the reference) of the spNNGP 1.0.1
 * response/conjugate NNGP hot path.  Written from the behaviour of the reference; citations are
 * reference file:line.  The dense m x m algebra follows the reference's route (Cholesky, explicit
 * inverse, symmetric mat-vec: src/rNNGP.cpp:86-89) with unblocked netlib-style loops.
 */
#include "nngp_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

void orc_set_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

static int thread_id(void) {
#ifdef _OPENMP
  return omp_get_thread_num();
#else
  return 0;
#endif
}

/* ---------------------------------------------------------------- index layout */

/* src/util.cpp:55-70: ragged CSR offsets; the first m locations have i neighbours each */
void orc_get_nn_indx(int i, int m, int *offset, int *count) {
  if (i == 0) {
    *offset = 0;
    *count = 0;
  } else if (i < m) {
    *offset = (int)((double)i / 2 * (i - 1));
    *count = i;
  } else {
    *offset = (int)((double)m / 2 * (m - 1) + (double)(i - m) * m);
    *count = m;
  }
}

long orc_n_indx(int n, int m) { return (long)((double)(1 + m) / 2 * m + (double)(n - m - 1) * m); }

/* ---------------------------------------------------------------- correlation */

/* src/util.cpp:51-53; pow(x,2) is x*x, products rounded separately (no FMA) */
double orc_dist2(double a1, double a2, double b1, double b2) {
  double dx = a1 - b1, dy = a2 - b2;
  double sx = dx * dx, sy = dy * dy;
  return sqrt(sx + sy);
}

double orc_bessel_k(double x, double nu) {
  /* K_nu(x) = int_0^inf exp(-x cosh t) cosh(nu t) dt.  The integrand is entire and decays double-
   * exponentially, so the trapezoid rule converges geometrically: h = 1/16 leaves the discretisation
   * error far below 1e-16 for x >= 1e-8.  All terms are positive (no cancellation). */
  const double h = 0.0625;
  if (!(x > 0.0)) return INFINITY;
  nu = fabs(nu);
  double sum = 0.5 * exp(-x);
  for (int k = 1; k < 100000; k++) {
    double t = k * h;
    double e = -x * cosh(t) + nu * t; /* cosh(nu t) = (e^{nu t} + e^{-nu t})/2 */
    double term = 0.5 * (exp(e) + exp(e - 2.0 * nu * t));
    sum += term;
    if (term < 1e-19 * sum && x * sinh(t) > nu) break;
  }
  return sum * h;
}

/* src/util.cpp:208-245 */
double orc_sp_cor(double D, double phi, double nu, int cov_model) {
  if (cov_model == ORC_EXPONENTIAL) {
    return exp(-phi * D);
  } else if (cov_model == ORC_SPHERICAL) {
    if (D > 0 && D <= 1.0 / phi) {
      return 1.0 - 1.5 * phi * D + 0.5 * pow(phi * D, 3);
    } else if (D >= 1.0 / phi) {
      return 0.0;
    } else {
      return 1.0;
    }
  } else if (cov_model == ORC_MATERN) {
    if (D * phi > 0.0) {
      return pow(D * phi, nu) / (pow(2, nu - 1) * tgamma(nu)) * orc_bessel_k(D * phi, nu);
    } else {
      return 1.0;
    }
  } else { /* gaussian */
    return exp(-1.0 * (pow(phi * D, 2)));
  }
}

/* ---------------------------------------------------------------- sorting (R's Shell sort) */

static int rcmp_nalast(double x, double y) {
  int nax = isnan(x), nay = isnan(y);
  if (nax && nay) return 0;
  if (nax) return 1;
  if (nay) return -1;
  if (x < y) return -1;
  if (x > y) return 1;
  return 0;
}

void orc_rsort_with_index(double *x, int *indx, int n) {
  int h;
  for (h = 1; h <= n / 9; h = 3 * h + 1) {
  }
  for (; h > 0; h /= 3) {
    for (int i = h; i < n; i++) {
      double v = x[i];
      int iv = indx[i];
      int j = i;
      while (j >= h && rcmp_nalast(x[j - h], v) > 0) {
        x[j] = x[j - h];
        indx[j] = indx[j - h];
        j -= h;
      }
      x[j] = v;
      indx[j] = iv;
    }
  }
}

/* ---------------------------------------------------------------- small dense algebra (lower, column-major) */

#define A_(i, j) a[(size_t)(j)*lda + (i)]

static int chol_lower(double *a, int n, int lda) { /* dpotf2 */
  for (int j = 0; j < n; j++) {
    double ajj = A_(j, j);
    for (int k = 0; k < j; k++) ajj -= A_(j, k) * A_(j, k);
    if (!(ajj > 0.0)) return j + 1;
    ajj = sqrt(ajj);
    A_(j, j) = ajj;
    for (int k = 0; k < j; k++) {
      double t = A_(j, k);
      for (int i = j + 1; i < n; i++) A_(i, j) -= t * A_(i, k);
    }
    double r = 1.0 / ajj;
    for (int i = j + 1; i < n; i++) A_(i, j) *= r;
  }
  return 0;
}

static int chol_inverse_lower(double *a, int n, int lda) { /* dtrti2 + dlauu2: lower triangle := inverse */
  for (int j = 0; j < n; j++)
    if (A_(j, j) == 0.0) return j + 1;
  for (int j = n - 1; j >= 0; j--) {
    A_(j, j) = 1.0 / A_(j, j);
    double ajj = -A_(j, j);
    for (int c = n - 1; c > j; c--) {
      double t = A_(c, j);
      if (t != 0.0) {
        for (int i = n - 1; i > c; i--) A_(i, j) += t * A_(i, c);
        A_(c, j) *= A_(c, c);
      }
    }
    for (int i = j + 1; i < n; i++) A_(i, j) *= ajj;
  }
  for (int i = 0; i < n; i++) {
    double aii = A_(i, i);
    if (i < n - 1) {
      double d = 0.0;
      for (int k = i; k < n; k++) d += A_(k, i) * A_(k, i);
      A_(i, i) = d;
      for (int j = 0; j < i; j++) {
        double t = 0.0;
        for (int k = i + 1; k < n; k++) t += A_(k, j) * A_(k, i);
        A_(i, j) = t + aii * A_(i, j);
      }
    } else {
      for (int j = 0; j <= i; j++) A_(i, j) *= aii;
    }
  }
  return 0;
}

static void symv_lower(const double *a, int n, int lda, const double *x, double *y) { /* y = A x */
  for (int i = 0; i < n; i++) y[i] = 0.0;
  for (int j = 0; j < n; j++) {
    double t1 = x[j], t2 = 0.0;
    y[j] += t1 * A_(j, j);
    for (int i = j + 1; i < n; i++) {
      y[i] += t1 * A_(i, j);
      t2 += A_(i, j) * x[i];
    }
    y[j] += t2;
  }
}

static double dot(const double *x, const double *y, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}

#undef A_

/* ---------------------------------------------------------------- B and F */

/* one location of src/rNNGP.cpp:74-93: C = sigmaSq*rho + tauSq*I (lower), c = sigmaSq*rho, B = C^-1 c,
 * F = sigmaSq - B.c + tauSq.  The conjugate variant (src/cNNGP.cpp:46-66) is sigmaSq=1, tauSq=alpha
 * with F = 1 + alpha - B.c (same arithmetic, different association order handled by the caller). */
static int bf_one(int i, double *B, double *F, double *c, double *C, const double *coords, const int *nnIndx,
                  const int *nnIndxLU, int n, double sigmaSq, double tauSq, double phi, double nu, int cov_model,
                  int conj) {
  int off = nnIndxLU[i], cnt = nnIndxLU[n + i];
  for (int k = 0; k < cnt; k++) {
    int nk = nnIndx[off + k];
    double e = orc_dist2(coords[i], coords[n + i], coords[nk], coords[n + nk]);
    double r = orc_sp_cor(e, phi, nu, cov_model);
    c[k] = conj ? r : sigmaSq * r;
    for (int l = 0; l <= k; l++) {
      int nl = nnIndx[off + l];
      e = orc_dist2(coords[nk], coords[n + nk], coords[nl], coords[n + nl]);
      r = orc_sp_cor(e, phi, nu, cov_model);
      C[l * cnt + k] = conj ? r : sigmaSq * r;
      if (l == k) C[l * cnt + k] += tauSq;
    }
  }
  int info = chol_lower(C, cnt, cnt);
  if (info) return info;
  info = chol_inverse_lower(C, cnt, cnt);
  if (info) return info;
  symv_lower(C, cnt, cnt, c, &B[off]);
  if (conj)
    F[i] = 1.0 + tauSq - dot(&B[off], c, cnt);
  else
    F[i] = sigmaSq - dot(&B[off], c, cnt) + tauSq;
  return 0;
}

static void bf_all(double *B, double *F, const double *coords, const int *nnIndx, const int *nnIndxLU, int n, int m,
                   double sigmaSq, double tauSq, double phi, double nu, int cov_model, int conj, int *info_out) {
  int nt = orc_max_threads();
  double *c = (double *)malloc(sizeof(double) * (size_t)m * nt);
  double *C = (double *)malloc(sizeof(double) * (size_t)m * m * nt);
  int fail = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (int i = 0; i < n; i++) {
    int t = thread_id();
    if (i > 0) {
      int info = bf_one(i, B, F, &c[(size_t)m * t], &C[(size_t)m * m * t], coords, nnIndx, nnIndxLU, n, sigmaSq,
                        tauSq, phi, nu, cov_model, conj);
      if (info) {
#ifdef _OPENMP
#pragma omp critical
#endif
        if (fail == 0 || i + 1 < fail) fail = i + 1;
      }
    } else {
      /* the reference also stores B[0] = 0 here, although B[0] belongs to location 1 (benign race,
       * SURVEY §5); location 1 overwrites it, so the restatement leaves B[0] to location 1. */
      F[0] = conj ? 1.0 + tauSq : sigmaSq + tauSq;
    }
  }
  free(c);
  free(C);
  if (info_out) *info_out = fail;
}

double orc_update_bf(double *B, double *F, const double *coords, const int *nnIndx, const int *nnIndxLU, int n, int m,
                     const double *theta, int cov_model, int *info_out) {
  double nu = (cov_model == ORC_MATERN) ? theta[3] : 0.0;
  bf_all(B, F, coords, nnIndx, nnIndxLU, n, m, theta[0], theta[1], theta[2], nu, cov_model, 0, info_out);
  double logDet = 0.0;
  for (int i = 0; i < n; i++) logDet += log(F[i]); /* serial, in index order: src/rNNGP.cpp:96-98 */
  return logDet;
}

void orc_update_conj_bf(double *B, double *F, const double *coords, const int *nnIndx, const int *nnIndxLU, int n,
                        int m, double phi, double alpha, double nu, int cov_model, int *info_out) {
  bf_all(B, F, coords, nnIndx, nnIndxLU, n, m, 1.0, alpha, phi, nu, cov_model, 1, info_out);
}

/* src/util.cpp:261-280 */
double orc_q(const double *B, const double *F, const double *u, const double *v, int n, const int *nnIndx,
             const int *nnIndxLU) {
  double q = 0.0;
#ifdef _OPENMP
#pragma omp parallel for reduction(+ : q) schedule(static)
#endif
  for (int i = 0; i < n; i++) {
    double a = 0.0, b = 0.0;
    int off = nnIndxLU[i], cnt = nnIndxLU[n + i];
    for (int j = 0; j < cnt; j++) {
      a += B[off + j] * u[nnIndx[off + j]];
      b += B[off + j] * v[nnIndx[off + j]];
    }
    q += (u[i] - a) * (v[i] - b) / F[i];
  }
  return q;
}

/* ---------------------------------------------------------------- neighbour search */

/* src/nn.cpp:63-111: scan predecessors in index order, strict '<' against the current worst,
 * overwrite the last slot and re-sort with R's Shell sort (this fixes the tie order). */
void orc_mk_nn_indx(int n, int m, const double *coords, int *nnIndx, double *nnDist, int *nnIndxLU) {
  long nIndx = orc_n_indx(n, m);
  for (long i = 0; i < nIndx; i++) nnDist[i] = INFINITY;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64)
#endif
  for (int i = 0; i < n; i++) {
    int off, cnt;
    orc_get_nn_indx(i, m, &off, &cnt);
    nnIndxLU[i] = off;
    nnIndxLU[n + i] = cnt;
    if (i != 0) {
      for (int j = 0; j < i; j++) {
        double d = orc_dist2(coords[i], coords[n + i], coords[j], coords[n + j]);
        if (d < nnDist[off + cnt - 1]) {
          nnDist[off + cnt - 1] = d;
          nnIndx[off + cnt - 1] = j;
          orc_rsort_with_index(&nnDist[off], &nnIndx[off], cnt);
        }
      }
    }
  }
}

/* src/nn.cpp:128-134 */
static double dmi(const double *x, const double *c, int inc) {
  double t = x[0] + x[inc] - c[0] - c[inc];
  return t * t;
}
static double dei(const double *x, const double *c, int inc) {
  double dx = x[0] - c[0], dy = x[inc] - c[inc];
  double sx = dx * dx, sy = dy * dy;
  return sx + sy;
}

/* src/nn.cpp:136-211: Ra & Kim mean-distance-ordered search along the (x+y)-sorted order */
static void fast_nn(int m, int n, const double *coords, int ui, const int *sIndx, int *rSIndx, double *rSNNDist) {
  if (m == 0) return; /* the reference reads rSNNDist[-1] here (UB); nothing to find */
  for (int i = 0; i < m; i++) rSNNDist[i] = INFINITY;
  int i = ui, j = ui;
  int up = 1, down = 1;
  while (up || down) {
    if (i == 0) down = 0;
    if (j == n - 1) up = 0;
    if (down) {
      i--;
      double dm = dmi(&coords[sIndx[ui]], &coords[sIndx[i]], n);
      if (dm > 2 * rSNNDist[m - 1]) {
        down = 0;
      } else {
        double de = dei(&coords[sIndx[ui]], &coords[sIndx[i]], n);
        if (de < rSNNDist[m - 1] && sIndx[i] < sIndx[ui]) {
          rSNNDist[m - 1] = de;
          rSIndx[m - 1] = sIndx[i];
          orc_rsort_with_index(rSNNDist, rSIndx, m);
        }
      }
    }
    if (up) {
      j++;
      double dm = dmi(&coords[sIndx[ui]], &coords[sIndx[j]], n);
      if (dm > 2 * rSNNDist[m - 1]) {
        up = 0;
      } else {
        double de = dei(&coords[sIndx[ui]], &coords[sIndx[j]], n);
        if (de < rSNNDist[m - 1] && sIndx[j] < sIndx[ui]) {
          rSNNDist[m - 1] = de;
          rSIndx[m - 1] = sIndx[j];
          orc_rsort_with_index(rSNNDist, rSIndx, m);
        }
      }
    }
  }
  for (int k = 0; k < m; k++) rSNNDist[k] = sqrt(rSNNDist[k]);
}

/* src/nn.cpp:214-257 */
void orc_mk_nn_indx_cb(int n, int m, const double *coords, int *nnIndx, double *nnDist, int *nnIndxLU) {
  int *sIndx = (int *)malloc(sizeof(int) * (size_t)n);
  double *u = (double *)malloc(sizeof(double) * (size_t)n);
  for (int i = 0; i < n; i++) {
    sIndx[i] = i;
    u[i] = coords[i] + coords[n + i];
  }
  orc_rsort_with_index(u, sIndx, n);
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 64)
#endif
  for (int i = 0; i < n; i++) {
    int off, cnt;
    orc_get_nn_indx(sIndx[i], m, &off, &cnt);
    nnIndxLU[sIndx[i]] = off;
    nnIndxLU[n + sIndx[i]] = cnt;
    fast_nn(cnt, n, coords, i, sIndx, &nnIndx[off], &nnDist[off]);
  }
  free(sIndx);
  free(u);
}

/* exact kNN of query sites among all n sites, ascending (d^2, index).  Stands in for RANN::nn2. */
void orc_nn_query(int n, const double *coords, int q, const double *coords0, int m, int *nnIndx0) {
#ifdef _OPENMP
#pragma omp parallel
#endif
  {
    double *bd = (double *)malloc(sizeof(double) * (size_t)m);
    int *bi = (int *)malloc(sizeof(int) * (size_t)m);
#ifdef _OPENMP
#pragma omp for schedule(static)
#endif
    for (int i = 0; i < q; i++) {
      int cnt = 0;
      for (int j = 0; j < n; j++) {
        double dx = coords0[i] - coords[j], dy = coords0[q + i] - coords[n + j];
        double sx = dx * dx, sy = dy * dy;
        double d = sx + sy;
        if (cnt < m || d < bd[cnt - 1]) { /* strict: earlier index wins ties */
          int k = (cnt < m) ? cnt++ : m - 1;
          while (k > 0 && bd[k - 1] > d) {
            bd[k] = bd[k - 1];
            bi[k] = bi[k - 1];
            k--;
          }
          bd[k] = d;
          bi[k] = j;
        }
      }
      for (int k = 0; k < m; k++) nnIndx0[(size_t)k * q + i] = (k < cnt) ? bi[k] : -1;
    }
    free(bd);
    free(bi);
  }
}

/* ---------------------------------------------------------------- response sampler */

static double logit(double theta, double a, double b) { return log((theta - a) / (b - theta)); } /* util.cpp:43 */
static double logit_inv(double z, double a, double b) { return b - (b - a) / (1 + exp(z)); }     /* util.cpp:47 */

/* src/rNNGP.cpp:317-324 */
static double log_post(double logDet, double Q, const double *theta, int cov_model, const double *sigmaSqIG,
                       const double *tauSqIG, const double *phiUnif, const double *nuUnif) {
  double lp = -0.5 * logDet - 0.5 * Q;
  lp += log(theta[2] - phiUnif[0]) + log(phiUnif[1] - theta[2]);
  lp += -1.0 * (1.0 + sigmaSqIG[0]) * log(theta[0]) - sigmaSqIG[1] / theta[0] + log(theta[0]);
  lp += -1.0 * (1.0 + tauSqIG[0]) * log(theta[1]) - tauSqIG[1] / theta[1] + log(theta[1]);
  if (cov_model == ORC_MATERN) lp += log(theta[3] - nuUnif[0]) + log(nuUnif[1] - theta[3]);
  return lp;
}

/* mkUIndx, src/nn.cpp:25-56 with mkUIndx2 / crs_csc (src/util.cpp:115-185): transpose of the neighbour graph.
 * uIndx: per location i the locations having i as a neighbour, ascending; uIndxLU: offsets then counts;
 * uiIndx: slot of i in each of those lists (which(), src/util.cpp:248-258).  search_type 0, 2: ascending; 1: descending. */
void orc_mk_u_indx(int n, int m, const int *nnIndx, const int *nnIndxLU, int search_type, int *uIndx, int *uIndxLU,
                   int *uiIndx) {
  long nIndx = orc_n_indx(n, m);
  int *start = (int *)calloc((size_t)n + 1, sizeof(int));
  for (long e = 0; e < nIndx; e++) start[nnIndx[e] + 1]++;          /* count per column */
  for (int i = 0; i < n; i++) start[i + 1] += start[i];             /* cumulative sum */
  for (int i = 0; i < n; i++) {
    uIndxLU[i] = start[i];
    uIndxLU[n + i] = start[i + 1] - start[i];
  }
  int *fill = (int *)malloc(sizeof(int) * (size_t)n);
  for (int i = 0; i < n; i++) fill[i] = start[i];
  for (int row = 0; row < n; row++) {                               /* rows ascending: each list ends up ascending */
    for (int k = 0; k < nnIndxLU[n + row]; k++) {
      int col = nnIndx[nnIndxLU[row] + k];
      uIndx[fill[col]] = row;
      uiIndx[fill[col]] = k;                                        /* == which(col, nn list of row) */
      fill[col]++;
    }
  }
  if (search_type == 1) { /* mkUIndx1 (src/util.cpp:93-111) scans j downwards: descending lists */
    for (int i = 0; i < n; i++)
      for (int a = start[i], b = start[i + 1] - 1; a < b; a++, b--) {
        int t = uIndx[a]; uIndx[a] = uIndx[b]; uIndx[b] = t;
        t = uiIndx[a]; uiIndx[a] = uiIndx[b]; uiIndx[b] = t;
      }
  }
  free(start); free(fill);
}

/* updateRep, src/rNNGP.cpp:24-42 (same loop: src/rNNGPReplicated.cpp:187-197): v_i = B_i . v_N(i) + sqrt(F_i) z_i,
 * one rnorm per location in index order. */
void orc_update_rep(const double *B, const double *F, int n, const int *nnIndx, const int *nnIndxLU, orc_rnorm_fn rnorm,
                    double *v) {
  for (int i = 0; i < n; i++) {
    double z = rnorm(0.0, 1.0);
    if (i == 0) {
      v[i] = sqrt(F[i]) * z;
    } else {
      double d = 0.0; /* ddot over the gathered neighbours, :35-38 */
      for (int j = 0; j < nnIndxLU[n + i]; j++) d += B[nnIndxLU[i] + j] * v[nnIndx[nnIndxLU[i] + j]];
      v[i] = d + sqrt(F[i]) * z;
    }
  }
}

/* rNNGPReplicated sample loop, src/rNNGPReplicated.cpp:183-201.  beta: nSamples x p column-major; theta: nTheta x
 * nSamples; repSamples: n x nSamples. */
int orc_rnngp_replicated(const double *X, int p, int n, int m, const double *coords, int cov_model, const int *nnIndx,
                         const int *nnIndxLU, const double *beta, const double *theta, int nSamples,
                         orc_rnorm_fn rnorm, double *repSamples) {
  int nTheta = (cov_model == ORC_MATERN) ? 4 : 3, info = 0;
  long nIndx = orc_n_indx(n, m);
  double *B = (double *)malloc(sizeof(double) * (size_t)(nIndx > 0 ? nIndx : 1));
  double *F = (double *)malloc(sizeof(double) * (size_t)n);
  for (int s = 0; s < nSamples && !info; s++) {
    double th[4] = {0, 0, 0, 0};
    memcpy(th, &theta[(size_t)s * nTheta], sizeof(double) * nTheta);
    orc_update_bf(B, F, coords, nnIndx, nnIndxLU, n, m, th, cov_model, &info);
    if (info) break;
    double *out = &repSamples[(size_t)s * n];
    orc_update_rep(B, F, n, nnIndx, nnIndxLU, rnorm, out);
    for (int j = 0; j < p; j++) { /* dgemv N, beta with stride nSamples, accumulating, :199-200 */
      double t = beta[(size_t)j * nSamples + s];
      const double *col = &X[(size_t)n * j];
      for (int i = 0; i < n; i++) out[i] += t * col[i];
    }
  }
  free(B); free(F);
  return info ? -1 : 0;
}

int orc_rnngp(const double *y, const double *X, int p, int n, int m, const double *coords, int cov_model,
              const int *nnIndx, const int *nnIndxLU, const double *sigmaSqIG, const double *tauSqIG,
              const double *phiUnif, const double *nuUnif, const double *betaStarting, double sigmaSqStarting,
              double tauSqStarting, double phiStarting, double nuStarting, double sigmaSqTuning, double tauSqTuning,
              double phiTuning, double nuTuning, int nSamples, orc_rnorm_fn rnorm, orc_runif_fn runif,
              double *betaSamples, double *thetaSamples) {
  return orc_rnngp_rep(y, X, p, n, m, coords, cov_model, nnIndx, nnIndxLU, sigmaSqIG, tauSqIG, phiUnif, nuUnif,
                       betaStarting, sigmaSqStarting, tauSqStarting, phiStarting, nuStarting, sigmaSqTuning, tauSqTuning,
                       phiTuning, nuTuning, nSamples, rnorm, runif, betaSamples, thetaSamples, NULL, NULL);
}

/* rNNGP with fit.rep: repIndx[s] != 0 marks iterations that draw a replicate (src/rNNGP.cpp:203-208,300-313);
 * repSamples n x nRep.  Both NULL: plain sampler. */
int orc_rnngp_rep(const double *y, const double *X, int p, int n, int m, const double *coords, int cov_model,
                  const int *nnIndx, const int *nnIndxLU, const double *sigmaSqIG, const double *tauSqIG,
                  const double *phiUnif, const double *nuUnif, const double *betaStarting, double sigmaSqStarting,
                  double tauSqStarting, double phiStarting, double nuStarting, double sigmaSqTuning, double tauSqTuning,
                  double phiTuning, double nuTuning, int nSamples, orc_rnorm_fn rnorm, orc_runif_fn runif,
                  double *betaSamples, double *thetaSamples, const int *repIndx, double *repSamples) {
  int nTheta = (cov_model == ORC_MATERN) ? 4 : 3;
  int repCnt = 0;
  double *tmp_n2 = (double *)malloc(sizeof(double) * (size_t)n);
  long nIndx = orc_n_indx(n, m);
  double theta[4] = {sigmaSqStarting, tauSqStarting, phiStarting, nuStarting};
  double tuning[4] = {sigmaSqTuning, tauSqTuning, phiTuning, nuTuning};
  double thetaCand[4] = {0, 0, 0, 0};
  double *beta = (double *)malloc(sizeof(double) * p);
  double *B = (double *)malloc(sizeof(double) * (size_t)(nIndx > 0 ? nIndx : 1));
  double *F = (double *)malloc(sizeof(double) * (size_t)n);
  double *tmp_pp = (double *)malloc(sizeof(double) * p * p);
  double *tmp_p = (double *)malloc(sizeof(double) * p);
  double *tmp_p2 = (double *)malloc(sizeof(double) * p);
  double *tmp_n = (double *)malloc(sizeof(double) * (size_t)n);
  memcpy(beta, betaStarting, sizeof(double) * p);
  int accept = 0, info = 0, status = 0;
  int thetaUpdate = 1;

  orc_update_bf(B, F, coords, nnIndx, nnIndxLU, n, m, theta, cov_model, &info); /* :255 */
  if (info) status = -1;

  for (int s = 0; s < nSamples && status == 0; s++) {
    if (thetaUpdate) { /* :274-292: beta | theta ~ N((X'C^-1X)^-1 X'C^-1 y, (X'C^-1X)^-1) */
      thetaUpdate = 0;
      for (int i = 0; i < p; i++) {
        tmp_p[i] = orc_q(B, F, &X[(size_t)n * i], y, n, nnIndx, nnIndxLU);
        for (int j = 0; j <= i; j++)
          tmp_pp[j * p + i] = orc_q(B, F, &X[(size_t)n * j], &X[(size_t)n * i], n, nnIndx, nnIndxLU);
      }
      if (chol_lower(tmp_pp, p, p) || chol_inverse_lower(tmp_pp, p, p)) { status = -1; break; }
      symv_lower(tmp_pp, p, p, tmp_p, tmp_p2);
      if (chol_lower(tmp_pp, p, p)) { status = -1; break; }
    }
    /* mvrnorm, src/util.cpp:28-41: des = mu + L z */
    for (int i = 0; i < p; i++) beta[i] = rnorm(0.0, 1.0);
    for (int j = p - 1; j >= 0; j--) { /* dtrmv L,N,N */
      double t = beta[j];
      if (t != 0.0) {
        for (int i = p - 1; i > j; i--) beta[i] += t * tmp_pp[j * p + i];
        beta[j] *= tmp_pp[j * p + j];
      }
    }
    for (int i = 0; i < p; i++) beta[i] += tmp_p2[i];

    /* residual r = X beta - y, :299-303 */
    for (int i = 0; i < n; i++) tmp_n[i] = 0.0;
    for (int j = 0; j < p; j++) {
      double t = beta[j];
      const double *col = &X[(size_t)n * j];
      for (int i = 0; i < n; i++) tmp_n[i] += t * col[i];
    }
    int rep = repIndx && repSamples && repIndx[s];
    if (rep) memcpy(&repSamples[(size_t)repCnt * n], tmp_n, sizeof(double) * (size_t)n); /* X beta, :300-302 */
    for (int i = 0; i < n; i++) tmp_n[i] -= y[i];

    double logDetCur = orc_update_bf(B, F, coords, nnIndx, nnIndxLU, n, m, theta, cov_model, &info); /* :306 */
    if (info) { status = -1; break; }
    if (rep) { /* :309-313 */
      orc_update_rep(B, F, n, nnIndx, nnIndxLU, rnorm, tmp_n2);
      for (int i = 0; i < n; i++) repSamples[(size_t)repCnt * n + i] += tmp_n2[i];
      repCnt++;
    }
    double QCur = orc_q(B, F, tmp_n, tmp_n, n, nnIndx, nnIndxLU);
    double lpCur = log_post(logDetCur, QCur, theta, cov_model, sigmaSqIG, tauSqIG, phiUnif, nuUnif);

    /* proposals in the reference's draw order phi, sigmaSq, tauSq[, nu], :327-333 */
    thetaCand[2] = logit_inv(rnorm(logit(theta[2], phiUnif[0], phiUnif[1]), tuning[2]), phiUnif[0], phiUnif[1]);
    thetaCand[0] = exp(rnorm(log(theta[0]), tuning[0]));
    thetaCand[1] = exp(rnorm(log(theta[1]), tuning[1]));
    if (cov_model == ORC_MATERN)
      thetaCand[3] = logit_inv(rnorm(logit(theta[3], nuUnif[0], nuUnif[1]), tuning[3]), nuUnif[0], nuUnif[1]);

    double logDetCand = orc_update_bf(B, F, coords, nnIndx, nnIndxLU, n, m, thetaCand, cov_model, &info); /* :336 */
    if (info) { status = -1; break; }
    double QCand = orc_q(B, F, tmp_n, tmp_n, n, nnIndx, nnIndxLU);
    double lpCand = log_post(logDetCand, QCand, thetaCand, cov_model, sigmaSqIG, tauSqIG, phiUnif, nuUnif);

    if (runif(0.0, 1.0) <= exp(lpCand - lpCur)) { /* :349 */
      thetaUpdate = 1;
      memcpy(theta, thetaCand, sizeof(double) * nTheta);
      accept++;
    }
    memcpy(&betaSamples[(size_t)s * p], beta, sizeof(double) * p);
    memcpy(&thetaSamples[(size_t)s * nTheta], theta, sizeof(double) * nTheta);
  }
  free(beta); free(B); free(F); free(tmp_pp); free(tmp_p); free(tmp_p2); free(tmp_n); free(tmp_n2);
  return status ? status : accept;
}

/* ---------------------------------------------------------------- conjugate model */

int orc_cnngp(const double *y, const double *X, int p, int n, const double *coords, const double *thetaAlpha,
              const int *nnIndx, const int *nnIndxLU, const double *X0, const double *coords0, int n0,
              const int *nnIndx0, int g, int m, const double *sigmaSqIG, int cov_model, double *betaHat, double *ab,
              double *bBInv, double *bb, double *y0Hat, double *y0HatVar) {
  int nTheta = (cov_model == ORC_MATERN) ? 3 : 2;
  long nIndx = orc_n_indx(n, m);
  int pp = p * p, status = 0;
  double *B = (double *)malloc(sizeof(double) * (size_t)(nIndx > 0 ? nIndx : 1));
  double *F = (double *)malloc(sizeof(double) * (size_t)n);
  double *tmp_pp = (double *)malloc(sizeof(double) * pp);
  double *tmp_p = (double *)malloc(sizeof(double) * p);
  int nt = orc_max_threads();
  double *work = (double *)malloc(sizeof(double) * (size_t)nt * ((size_t)m * m + 3 * m + 2 * p));

  for (int k = 0; k < g && status == 0; k++) {
    double phi = thetaAlpha[k * nTheta], alpha = thetaAlpha[k * nTheta + 1];
    double nu = (cov_model == ORC_MATERN) ? thetaAlpha[k * nTheta + 2] : 0.0;
    int info = 0;
    orc_update_conj_bf(B, F, coords, nnIndx, nnIndxLU, n, m, phi, alpha, nu, cov_model, &info); /* :224 */
    if (info) { status = -1; break; }
    for (int i = 0; i < p; i++) { /* :227-232 */
      tmp_p[i] = orc_q(B, F, &X[(size_t)n * i], y, n, nnIndx, nnIndxLU);
      for (int j = 0; j <= i; j++)
        tmp_pp[j * p + i] = orc_q(B, F, &X[(size_t)n * j], &X[(size_t)n * i], n, nnIndx, nnIndxLU);
    }
    if (chol_lower(tmp_pp, p, p) || chol_inverse_lower(tmp_pp, p, p)) { status = -1; break; } /* :235-236 */
    double *bk_ = &betaHat[(size_t)k * p];
    symv_lower(tmp_pp, p, p, tmp_p, bk_); /* :239 */
    ab[k * 2] = sigmaSqIG[0] + 1.0 * n / 2.0; /* :242 */
    ab[k * 2 + 1] = sigmaSqIG[1] + 0.5 * (orc_q(B, F, y, y, n, nnIndx, nnIndxLU) - dot(bk_, tmp_p, p)); /* :245 */
    memcpy(&bBInv[(size_t)k * pp], tmp_pp, sizeof(double) * pp);
    memcpy(&bb[(size_t)k * p], tmp_p, sizeof(double) * p);

    int fail = 0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int i = 0; i < n0; i++) { /* :255-300 */
      double *C0 = &work[(size_t)thread_id() * ((size_t)m * m + 3 * m + 2 * p)];
      double *c0 = C0 + (size_t)m * m, *w = c0 + m, *tm = w + m, *tp = tm + m, *tp2 = tp + p;
      for (int j = 0; j < m; j++) {
        int nj = nnIndx0[(size_t)j * n0 + i];
        double e = orc_dist2(coords0[i], coords0[n0 + i], coords[nj], coords[n + nj]);
        c0[j] = orc_sp_cor(e, phi, nu, cov_model);
        for (int l = 0; l <= j; l++) {
          int nl = nnIndx0[(size_t)l * n0 + i];
          e = orc_dist2(coords[nj], coords[n + nj], coords[nl], coords[n + nl]);
          C0[l * m + j] = orc_sp_cor(e, phi, nu, cov_model);
          if (l == j) C0[l * m + j] += alpha;
        }
      }
      if (chol_lower(C0, m, m) || chol_inverse_lower(C0, m, m)) { fail = 1; continue; }
      symv_lower(C0, m, m, c0, w);
      for (int j = 0; j < m; j++) { /* residual at the neighbours, :278-280 */
        int nj = nnIndx0[(size_t)j * n0 + i];
        double xb = 0.0;
        for (int c = 0; c < p; c++) xb += X[(size_t)c * n + nj] * bk_[c];
        tm[j] = y[nj] - xb;
      }
      double x0b = 0.0;
      for (int c = 0; c < p; c++) x0b += X0[(size_t)c * n0 + i] * bk_[c];
      y0Hat[(size_t)k * n0 + i] = x0b + dot(w, tm, m); /* :282 */
      for (int c = 0; c < p; c++) { /* u = x0 - X_N' w, :285-293 */
        double t = 0.0;
        for (int j = 0; j < m; j++) t += X[(size_t)c * n + nnIndx0[(size_t)j * n0 + i]] * w[j];
        tp[c] = X0[(size_t)c * n0 + i] - t;
      }
      symv_lower(tmp_pp, p, p, tp, tp2); /* :296 */
      y0HatVar[(size_t)k * n0 + i] =
          ab[k * 2 + 1] * (dot(tp, tp2, p) + 1.0 + alpha - dot(w, c0, m)) / (ab[k * 2] - 1.0); /* :298 */
    }
    if (fail) status = -1;
  }
  free(B); free(F); free(tmp_pp); free(tmp_p); free(work);
  return status;
}

/* ---------------------------------------------------------------- prediction */

int orc_rnngp_predict(const double *X, const double *y, const double *coords, int n, int p, int m, const double *X0,
                      const double *coords0, int q, const int *nnIndx0, const double *betaSamples,
                      const double *thetaSamples, int nSamples, int cov_model, const double *z, double *y0,
                      int site_begin, int site_end) {
  int nTheta = (cov_model == ORC_MATERN) ? 4 : 3;
  int nt = orc_max_threads();
  double *work = (double *)malloc(sizeof(double) * (size_t)nt * ((size_t)m * m + 2 * m));
  int fail = 0;
  for (int i = site_begin; i < site_end; i++) { /* serial over sites, parallel over samples: :139-143 */
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (int s = 0; s < nSamples; s++) {
      double *C = &work[(size_t)thread_id() * ((size_t)m * m + 2 * m)];
      double *c = C + (size_t)m * m, *w = c + m;
      const double *th = &thetaSamples[(size_t)s * nTheta];
      double sigmaSq = th[0], tauSq = th[1], phi = th[2], nu = (cov_model == ORC_MATERN) ? th[3] : 0.0;
      for (int k = 0; k < m; k++) {
        int nk = nnIndx0[i + (size_t)q * k];
        double d = orc_dist2(coords[nk], coords[n + nk], coords0[i], coords0[q + i]);
        c[k] = sigmaSq * orc_sp_cor(d, phi, nu, cov_model);
        for (int l = 0; l < m; l++) {
          int nl = nnIndx0[i + (size_t)q * l];
          d = orc_dist2(coords[nk], coords[n + nk], coords[nl], coords[n + nl]);
          C[l * m + k] = sigmaSq * orc_sp_cor(d, phi, nu, cov_model);
          if (k == l) C[l * m + k] += tauSq;
        }
      }
      if (chol_lower(C, m, m) || chol_inverse_lower(C, m, m)) { fail = 1; continue; }
      symv_lower(C, m, m, c, w);
      double d = 0.0;
      for (int k = 0; k < m; k++) {
        int nk = nnIndx0[i + (size_t)q * k];
        double xb = 0.0;
        for (int j = 0; j < p; j++) xb += X[nk + (size_t)n * j] * betaSamples[(size_t)s * p + j];
        d += w[k] * (y[nk] - xb);
      }
      double x0b = 0.0;
      for (int j = 0; j < p; j++) x0b += X0[i + (size_t)q * j] * betaSamples[(size_t)s * p + j];
      y0[(size_t)s * q + i] = x0b + d + sqrt(sigmaSq + tauSq - dot(w, c, m)) * z[(size_t)i * nSamples + s];
    }
  }
  free(work);
  return fail ? -1 : 0;
}

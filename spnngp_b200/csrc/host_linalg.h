// host_linalg.h -- the p x p (p = number of covariates) dense algebra the host sampler needs.
// Lower triangle, column-major.  Same route as the reference's beta update (Cholesky, explicit
// inverse, symmetric mat-vec, Cholesky of the inverse; src/rNNGP.cpp:288-291, src/util.cpp:28-41)
// so that beta draws agree with the reference to rounding.  Unblocked textbook loops.
#pragma once
#include <cmath>
#include <vector>

namespace nngp {
namespace hl {

// A = L L', lower in place.  Returns 0 or the 1-based index of the failing pivot.
inline int chol(double *a, int n) {
  for (int j = 0; j < n; j++) {
    double d = a[j * n + j];
    for (int k = 0; k < j; k++) d -= a[k * n + j] * a[k * n + j];
    if (!(d > 0.0)) return j + 1;
    d = std::sqrt(d);
    a[j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double s = a[j * n + i];
      for (int k = 0; k < j; k++) s -= a[k * n + i] * a[k * n + j];
      a[j * n + i] = s / d;
    }
  }
  return 0;
}

// lower triangle of A^-1 from its Cholesky factor (in place): inv(L), then inv(L)' inv(L).
inline void chol_inverse(double *a, int n) {
  std::vector<double> li((size_t)n * n, 0.0);  // inv(L), column-major lower
  for (int j = 0; j < n; j++) {
    li[j * n + j] = 1.0 / a[j * n + j];
    for (int i = j + 1; i < n; i++) {
      double s = 0.0;
      for (int k = j; k < i; k++) s += a[k * n + i] * li[j * n + k];
      li[j * n + i] = -s / a[i * n + i];
    }
  }
  for (int j = 0; j < n; j++)
    for (int i = j; i < n; i++) {
      double s = 0.0;
      for (int k = i; k < n; k++) s += li[i * n + k] * li[j * n + k];
      a[j * n + i] = s;
    }
}

// y = A x, A symmetric with the lower triangle stored
inline void symv(const double *a, int n, const double *x, double *y) {
  for (int i = 0; i < n; i++) {
    double s = 0.0;
    for (int j = 0; j < n; j++) s += ((i >= j) ? a[j * n + i] : a[i * n + j]) * x[j];
    y[i] = s;
  }
}

// x := L x (lower, non-unit)
inline void trmv_lower(const double *a, int n, double *x) {
  for (int i = n - 1; i >= 0; i--) {
    double s = 0.0;
    for (int j = 0; j <= i; j++) s += a[j * n + i] * x[j];
    x[i] = s;
  }
}

inline double dot(const double *x, const double *y, int n) {
  double s = 0.0;
  for (int i = 0; i < n; i++) s += x[i] * y[i];
  return s;
}

}  // namespace hl
}  // namespace nngp

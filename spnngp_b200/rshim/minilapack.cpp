// rshim: lock-free, unblocked BLAS/LAPACK subset in the style of the netlib reference
// algorithms (dpotf2; dtrti2 + dlauu2; reference level-1/2 BLAS).  OUR restatement of those
// published algorithms -- not any vendor BLAS -- so that code written against R's BLAS/LAPACK
// interface links without R, and scales under OpenMP (a pthread OpenBLAS serialises).
// Column-major, Fortran calling convention (all arguments by pointer, trailing underscore).
#include <cmath>
#include "include/R_ext/Lapack.h"

static inline bool is_lower(const char *uplo) { return *uplo == 'L' || *uplo == 'l'; }

extern "C" {

double ddot_(const int *n_, const double *x, const int *incx_, const double *y, const int *incy_) {
  int n = *n_, incx = *incx_, incy = *incy_;
  double s = 0.0;
  if (n <= 0) return 0.0;
  int ix = incx < 0 ? (1 - n) * incx : 0, iy = incy < 0 ? (1 - n) * incy : 0;
  for (int i = 0; i < n; i++, ix += incx, iy += incy) s += x[ix] * y[iy];
  return s;
}

void dcopy_(const int *n_, const double *x, const int *incx_, double *y, const int *incy_) {
  int n = *n_, incx = *incx_, incy = *incy_;
  int ix = incx < 0 ? (1 - n) * incx : 0, iy = incy < 0 ? (1 - n) * incy : 0;
  for (int i = 0; i < n; i++, ix += incx, iy += incy) y[iy] = x[ix];
}

void daxpy_(const int *n_, const double *alpha, const double *x, const int *incx_, double *y, const int *incy_) {
  int n = *n_, incx = *incx_, incy = *incy_;
  double a = *alpha;
  if (n <= 0 || a == 0.0) return;
  int ix = incx < 0 ? (1 - n) * incx : 0, iy = incy < 0 ? (1 - n) * incy : 0;
  for (int i = 0; i < n; i++, ix += incx, iy += incy) y[iy] += a * x[ix];
}

void dscal_(const int *n_, const double *alpha, double *x, const int *incx_) {
  int n = *n_, incx = *incx_;
  for (int i = 0, ix = 0; i < n; i++, ix += incx) x[ix] *= *alpha;
}

void dgemv_(const char *trans, const int *m_, const int *n_, const double *alpha_, const double *a, const int *lda_,
            const double *x, const int *incx_, const double *beta_, double *y, const int *incy_) {
  int m = *m_, n = *n_, lda = *lda_, incx = *incx_, incy = *incy_;
  double alpha = *alpha_, beta = *beta_;
  bool notrans = (*trans == 'N' || *trans == 'n');
  int leny = notrans ? m : n, lenx = notrans ? n : m;
  if (m <= 0 || n <= 0) return;
  int kx = incx < 0 ? (1 - lenx) * incx : 0, ky = incy < 0 ? (1 - leny) * incy : 0;
  if (beta != 1.0) {
    for (int i = 0, iy = ky; i < leny; i++, iy += incy) y[iy] = (beta == 0.0) ? 0.0 : beta * y[iy];
  }
  if (alpha == 0.0) return;
  if (notrans) {  // y += alpha*A*x, column sweep
    for (int j = 0, jx = kx; j < n; j++, jx += incx) {
      double t = alpha * x[jx];
      const double *col = a + (size_t)j * lda;
      for (int i = 0, iy = ky; i < m; i++, iy += incy) y[iy] += t * col[i];
    }
  } else {  // y += alpha*A'*x
    for (int j = 0, jy = ky; j < n; j++, jy += incy) {
      const double *col = a + (size_t)j * lda;
      double t = 0.0;
      for (int i = 0, ix = kx; i < m; i++, ix += incx) t += col[i] * x[ix];
      y[jy] += alpha * t;
    }
  }
}

void dsymv_(const char *uplo, const int *n_, const double *alpha_, const double *a, const int *lda_,
            const double *x, const int *incx_, const double *beta_, double *y, const int *incy_) {
  int n = *n_, lda = *lda_, incx = *incx_, incy = *incy_;
  double alpha = *alpha_, beta = *beta_;
  if (n <= 0) return;
  int kx = incx < 0 ? (1 - n) * incx : 0, ky = incy < 0 ? (1 - n) * incy : 0;
  if (beta != 1.0) {
    for (int i = 0, iy = ky; i < n; i++, iy += incy) y[iy] = (beta == 0.0) ? 0.0 : beta * y[iy];
  }
  if (alpha == 0.0) return;
  if (is_lower(uplo)) {
    for (int j = 0, jx = kx, jy = ky; j < n; j++, jx += incx, jy += incy) {
      const double *col = a + (size_t)j * lda;
      double t1 = alpha * x[jx], t2 = 0.0;
      y[jy] += t1 * col[j];
      for (int i = j + 1, ix = jx + incx, iy = jy + incy; i < n; i++, ix += incx, iy += incy) {
        y[iy] += t1 * col[i];
        t2 += col[i] * x[ix];
      }
      y[jy] += alpha * t2;
    }
  } else {
    for (int j = 0, jx = kx, jy = ky; j < n; j++, jx += incx, jy += incy) {
      const double *col = a + (size_t)j * lda;
      double t1 = alpha * x[jx], t2 = 0.0;
      for (int i = 0, ix = kx, iy = ky; i < j; i++, ix += incx, iy += incy) {
        y[iy] += t1 * col[i];
        t2 += col[i] * x[ix];
      }
      y[jy] += t1 * col[j] + alpha * t2;
    }
  }
}

// x := op(A) x, A triangular.  Supports uplo in {L,U}, trans in {N,T}, diag in {N,U}; incx = any.
void dtrmv_(const char *uplo, const char *trans, const char *diag, const int *n_, const double *a, const int *lda_,
            double *x, const int *incx_) {
  int n = *n_, lda = *lda_, incx = *incx_;
  bool lower = is_lower(uplo), notrans = (*trans == 'N' || *trans == 'n'), nounit = (*diag == 'N' || *diag == 'n');
  if (n <= 0) return;
  int kx = incx < 0 ? (1 - n) * incx : 0;
#define X(i) x[kx + (i) * incx]
#define Aij(i, j) a[(size_t)(j) * lda + (i)]
  if (notrans) {
    if (lower) {
      for (int j = n - 1; j >= 0; j--) {
        double t = X(j);
        if (t != 0.0) {
          for (int i = n - 1; i > j; i--) X(i) += t * Aij(i, j);
          if (nounit) X(j) *= Aij(j, j);
        }
      }
    } else {
      for (int j = 0; j < n; j++) {
        double t = X(j);
        if (t != 0.0) {
          for (int i = 0; i < j; i++) X(i) += t * Aij(i, j);
          if (nounit) X(j) *= Aij(j, j);
        }
      }
    }
  } else {
    if (lower) {
      for (int j = 0; j < n; j++) {
        double t = X(j);
        if (nounit) t *= Aij(j, j);
        for (int i = j + 1; i < n; i++) t += Aij(i, j) * X(i);
        X(j) = t;
      }
    } else {
      for (int j = n - 1; j >= 0; j--) {
        double t = X(j);
        if (nounit) t *= Aij(j, j);
        for (int i = j - 1; i >= 0; i--) t += Aij(i, j) * X(i);
        X(j) = t;
      }
    }
  }
#undef X
#undef Aij
}

// Cholesky A = L L' (lower) or U'U (upper), unblocked (dpotf2).
void dpotrf_(const char *uplo, const int *n_, double *a, const int *lda_, int *info) {
  int n = *n_, lda = *lda_;
  *info = 0;
#define Aij(i, j) a[(size_t)(j) * lda + (i)]
  if (is_lower(uplo)) {
    for (int j = 0; j < n; j++) {
      double ajj = Aij(j, j);
      for (int k = 0; k < j; k++) ajj -= Aij(j, k) * Aij(j, k);
      if (!(ajj > 0.0)) { Aij(j, j) = ajj; *info = j + 1; return; }
      ajj = sqrt(ajj);
      Aij(j, j) = ajj;
      if (j < n - 1) {
        for (int k = 0; k < j; k++) {
          double t = Aij(j, k);
          for (int i = j + 1; i < n; i++) Aij(i, j) -= t * Aij(i, k);
        }
        double r = 1.0 / ajj;
        for (int i = j + 1; i < n; i++) Aij(i, j) *= r;
      }
    }
  } else {
    for (int j = 0; j < n; j++) {
      double ajj = Aij(j, j);
      for (int k = 0; k < j; k++) ajj -= Aij(k, j) * Aij(k, j);
      if (!(ajj > 0.0)) { Aij(j, j) = ajj; *info = j + 1; return; }
      ajj = sqrt(ajj);
      Aij(j, j) = ajj;
      for (int i = j + 1; i < n; i++) {
        double t = Aij(j, i);
        for (int k = 0; k < j; k++) t -= Aij(k, j) * Aij(k, i);
        Aij(j, i) = t / ajj;
      }
    }
  }
#undef Aij
}

// inverse from the Cholesky factor: dtrti2 (triangular inverse) then dlauu2 (L'L or UU').
void dpotri_(const char *uplo, const int *n_, double *a, const int *lda_, int *info) {
  int n = *n_, lda = *lda_;
  *info = 0;
#define Aij(i, j) a[(size_t)(j) * lda + (i)]
  for (int j = 0; j < n; j++)
    if (Aij(j, j) == 0.0) { *info = j + 1; return; }
  if (is_lower(uplo)) {
    // dtrti2, lower, non-unit: columns from the last to the first
    for (int j = n - 1; j >= 0; j--) {
      Aij(j, j) = 1.0 / Aij(j, j);
      double ajj = -Aij(j, j);
      if (j < n - 1) {
        // x := T x with T = inverted trailing block (lower, no-trans), x = A(j+1:n, j)
        for (int c = n - 1; c > j; c--) {
          double t = Aij(c, j);
          if (t != 0.0) {
            for (int i = n - 1; i > c; i--) Aij(i, j) += t * Aij(i, c);
            Aij(c, j) *= Aij(c, c);
          }
        }
        for (int i = j + 1; i < n; i++) Aij(i, j) *= ajj;
      }
    }
    // dlauu2, lower: A := L' L
    for (int i = 0; i < n; i++) {
      double aii = Aij(i, i);
      if (i < n - 1) {
        double d = 0.0;
        for (int k = i; k < n; k++) d += Aij(k, i) * Aij(k, i);
        Aij(i, i) = d;
        for (int j = 0; j < i; j++) {
          double t = 0.0;
          for (int k = i + 1; k < n; k++) t += Aij(k, j) * Aij(k, i);
          Aij(i, j) = t + aii * Aij(i, j);
        }
      } else {
        for (int j = 0; j <= i; j++) Aij(i, j) *= aii;
      }
    }
  } else {
    // upper: invert U column by column (dtrti2, upper), then A := U U'
    for (int j = 0; j < n; j++) {
      Aij(j, j) = 1.0 / Aij(j, j);
      double ajj = -Aij(j, j);
      for (int c = 0; c < j; c++) {
        double t = Aij(c, j);
        if (t != 0.0) {
          for (int i = 0; i < c; i++) Aij(i, j) += t * Aij(i, c);
          Aij(c, j) *= Aij(c, c);
        }
      }
      for (int i = 0; i < j; i++) Aij(i, j) *= ajj;
    }
    for (int i = 0; i < n; i++) {
      double aii = Aij(i, i);
      if (i < n - 1) {
        double d = 0.0;
        for (int k = i; k < n; k++) d += Aij(i, k) * Aij(i, k);
        Aij(i, i) = d;
        for (int j = 0; j < i; j++) {
          double t = 0.0;
          for (int k = i + 1; k < n; k++) t += Aij(j, k) * Aij(i, k);
          Aij(j, i) = t + aii * Aij(j, i);
        }
      } else {
        for (int j = 0; j <= i; j++) Aij(j, i) *= aii;
      }
    }
  }
#undef Aij
}

}  // extern "C"

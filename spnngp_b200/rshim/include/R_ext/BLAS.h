/* rshim stand-in for <R_ext/BLAS.h>: the level-1/2 subset used on the NNGP path.
 * Implemented in minilapack.cpp as unblocked netlib-style reference loops (lock-free under OpenMP). */
#ifndef RSHIM_BLAS_H
#define RSHIM_BLAS_H
#include "RS.h"
#ifdef __cplusplus
extern "C" {
#endif
double F77_NAME(ddot)(const int *n, const double *x, const int *incx, const double *y, const int *incy);
void F77_NAME(dcopy)(const int *n, const double *x, const int *incx, double *y, const int *incy);
void F77_NAME(daxpy)(const int *n, const double *alpha, const double *x, const int *incx, double *y, const int *incy);
void F77_NAME(dscal)(const int *n, const double *alpha, double *x, const int *incx);
void F77_NAME(dgemv)(const char *trans, const int *m, const int *n, const double *alpha, const double *a, const int *lda,
                     const double *x, const int *incx, const double *beta, double *y, const int *incy);
void F77_NAME(dsymv)(const char *uplo, const int *n, const double *alpha, const double *a, const int *lda,
                     const double *x, const int *incx, const double *beta, double *y, const int *incy);
void F77_NAME(dtrmv)(const char *uplo, const char *trans, const char *diag, const int *n, const double *a, const int *lda,
                     double *x, const int *incx);
void F77_NAME(dgemm)(const char *transa, const char *transb, const int *m, const int *n, const int *k, const double *alpha,
                     const double *a, const int *lda, const double *b, const int *ldb, const double *beta, double *c, const int *ldc);
void F77_NAME(dsymm)(const char *side, const char *uplo, const int *m, const int *n, const double *alpha,
                     const double *a, const int *lda, const double *b, const int *ldb, const double *beta, double *c, const int *ldc);
#ifdef __cplusplus
}
#endif
#endif

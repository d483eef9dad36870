#ifndef RSHIM_LAPACK_H
#define RSHIM_LAPACK_H
#include "RS.h"
#include "BLAS.h"
#ifdef __cplusplus
extern "C" {
#endif
/* unblocked dpotf2 / dtrti2+dlauum-style reference algorithms, see minilapack.cpp */
void F77_NAME(dpotrf)(const char *uplo, const int *n, double *a, const int *lda, int *info);
void F77_NAME(dpotri)(const char *uplo, const int *n, double *a, const int *lda, int *info);
#ifdef __cplusplus
}
#endif
#endif

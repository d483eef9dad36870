/* rshim stand-in for <Rinternals.h>: a tiny SEXP model (vectors, matrices, named lists). */
#ifndef RSHIM_RINTERNALS_H
#define RSHIM_RINTERNALS_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned int SEXPTYPE;
#define NILSXP   0
#define CHARSXP  9
#define LGLSXP  10
#define INTSXP  13
#define REALSXP 14
#define STRSXP  16
#define VECSXP  19

typedef int R_len_t;
typedef ptrdiff_t R_xlen_t;

typedef struct SEXPREC {
  SEXPTYPE type;
  R_xlen_t len;
  int nrow, ncol;         /* 0,0 when no dim attribute */
  void *data;             /* double* / int* / char* / SEXP* */
  struct SEXPREC *names;  /* names attribute or NULL */
  int owns;               /* data allocated by the shim (vs. borrowed from the caller) */
} SEXPREC, *SEXP;

extern SEXP R_NilValue;

double *REAL(SEXP x);
int *INTEGER(SEXP x);
int *LOGICAL(SEXP x);
const char *R_CHAR(SEXP x);
#define CHAR(x) R_CHAR(x)
SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n);
SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol);
SEXP Rf_mkChar(const char *s);
SEXP Rf_mkString(const char *s);
SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP VECTOR_ELT(SEXP x, R_xlen_t i);
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v);
SEXP STRING_ELT(SEXP x, R_xlen_t i);
SEXP Rf_namesgets(SEXP x, SEXP names);
SEXP Rf_getAttrib_names(SEXP x);
R_len_t Rf_length(SEXP x);
R_xlen_t Rf_xlength(SEXP x);
int Rf_nrows(SEXP x);
int Rf_ncols(SEXP x);
SEXP Rf_protect(SEXP x);
void Rf_unprotect(int n);
SEXP Rf_ScalarInteger(int v);
SEXP Rf_ScalarReal(double v);
#define PROTECT(x) Rf_protect(x)
#define UNPROTECT(n) Rf_unprotect(n)

/* ---- shim-only helpers (not part of R): drive .Call entry points from C/ctypes ---- */
SEXP rshim_wrap_real(double *p, R_xlen_t n);          /* borrow caller memory */
SEXP rshim_wrap_int(int *p, R_xlen_t n);
SEXP rshim_wrap_real_matrix(double *p, int nrow, int ncol);
SEXP rshim_wrap_int_matrix(int *p, int nrow, int ncol);
SEXP rshim_list_get(SEXP list, const char *name);     /* NULL when absent */
void rshim_reset(void);                               /* free every SEXP / R_alloc block */
void rshim_set_seed(unsigned int seed);               /* R-style set.seed (MT19937 + inversion) */
void rshim_set_quiet(int quiet);                      /* drop Rprintf output */
const char *rshim_last_error(void);
/* call fn(args[0..nargs)) under a setjmp guard; returns 0 and *out on success, 1 on Rf_error */
int rshim_call(void *fn, int nargs, SEXP *args, SEXP *out);
#ifdef __cplusplus
}
#endif
#endif

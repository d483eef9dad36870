// rshim runtime: a miniature stand-in for libR so that spNNGP-style `.Call` entry points (the
// reference's, for the oracle build, and this repo's R glue) can be built and driven on hosts
// without R.  Everything here is OUR restatement of R behaviour, written from R's documented
// semantics; nothing is copied from the spNNGP reference.
//
//  * SEXP model: vectors / matrices / named lists with borrowed or owned storage.
//  * RNG: Mersenne-Twister (MT19937) seeded and consumed the way R's RNG.c does it, normal
//    deviates by inversion with Wichura's AS241 (R's defaults "Mersenne-Twister"/"Inversion").
//    Checked in tests against the well-known R values set.seed(1): runif -> 0.2655087,
//    rnorm -> -0.6264538 (6-7 digits is all that is public knowledge here; bit-equality with a
//    real R stream is unverified on this box).
//  * rsort_with_index: R's Shell sort (src/main/sort.c): gaps 1,4,13,... started at h<=n/9,
//    strict '>' comparison with NaN last.  This fixes the neighbour tie order.
#include <cstdarg>
#include <cstdint>
#include <cstring>
#include <csetjmp>
#include <cmath>
#include <vector>
#include <string>
#include <mutex>

#include "include/R.h"
#include "include/Rinternals.h"
#include "include/Rmath.h"

namespace {

std::vector<SEXP> g_nodes;       // every SEXP handed out
std::vector<void *> g_blocks;    // R_alloc blocks
std::mutex g_mu;
int g_quiet = 1;
std::string g_last_error;
thread_local jmp_buf *g_jmp = nullptr;

SEXPREC g_nil = {NILSXP, 0, 0, 0, nullptr, nullptr, 0};

SEXP new_node(SEXPTYPE type, R_xlen_t n, void *data, int owns) {
  SEXP s = (SEXP)calloc(1, sizeof(SEXPREC));
  s->type = type; s->len = n; s->data = data; s->owns = owns; s->names = nullptr;
  std::lock_guard<std::mutex> lk(g_mu);
  g_nodes.push_back(s);
  return s;
}

size_t elt_size(SEXPTYPE t) {
  switch (t) {
    case REALSXP: return sizeof(double);
    case INTSXP: case LGLSXP: return sizeof(int);
    case VECSXP: case STRSXP: return sizeof(SEXP);
    default: return 1;
  }
}

// ---------------- MT19937, R flavour ----------------
const int MT_N = 624, MT_M = 397;
uint32_t g_dummy[625];
bool g_seeded = false;

void mt_sgenrand(uint32_t seed) {
  uint32_t *mt = g_dummy + 1;
  for (int i = 0; i < MT_N; i++) {
    mt[i] = seed & 0xffff0000u; seed = 69069u * seed + 1u;
    mt[i] |= (seed & 0xffff0000u) >> 16; seed = 69069u * seed + 1u;
  }
  g_dummy[0] = MT_N;
}

double mt_genrand() {
  static const uint32_t mag01[2] = {0x0u, 0x9908b0dfu};
  uint32_t *mt = g_dummy + 1;
  uint32_t y;
  int mti = (int)g_dummy[0];
  if (mti >= MT_N) {
    int kk;
    if (mti == MT_N + 1) mt_sgenrand(4357u);
    for (kk = 0; kk < MT_N - MT_M; kk++) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + MT_M] ^ (y >> 1) ^ mag01[y & 1u];
    }
    for (; kk < MT_N - 1; kk++) {
      y = (mt[kk] & 0x80000000u) | (mt[kk + 1] & 0x7fffffffu);
      mt[kk] = mt[kk + (MT_M - MT_N)] ^ (y >> 1) ^ mag01[y & 1u];
    }
    y = (mt[MT_N - 1] & 0x80000000u) | (mt[0] & 0x7fffffffu);
    mt[MT_N - 1] = mt[MT_M - 1] ^ (y >> 1) ^ mag01[y & 1u];
    mti = 0;
  }
  y = mt[mti++];
  y ^= (y >> 11);
  y ^= (y << 7) & 0x9d2c5680u;
  y ^= (y << 15) & 0xefc60000u;
  y ^= (y >> 18);
  g_dummy[0] = (uint32_t)mti;
  return (double)y * 2.3283064365386963e-10;
}

double fixup(double x) {
  const double i2_32m1 = 2.328306437080797e-10;
  if (x <= 0.0) return 0.5 * i2_32m1;
  if ((1.0 - x) <= 0.0) return 1.0 - 0.5 * i2_32m1;
  return x;
}

}  // namespace

extern "C" {

SEXP R_NilValue = &g_nil;

// ---------------- SEXP accessors ----------------
double *REAL(SEXP x) { return (double *)x->data; }
int *INTEGER(SEXP x) { return (int *)x->data; }
int *LOGICAL(SEXP x) { return (int *)x->data; }
const char *R_CHAR(SEXP x) { return (const char *)x->data; }

SEXP Rf_allocVector(SEXPTYPE type, R_xlen_t n) {
  size_t bytes = (size_t)(n > 0 ? n : 1) * elt_size(type);
  void *d = calloc(1, bytes);
  if (!d) Rf_error("rshim: cannot allocate vector of %ld bytes", (long)bytes);
  return new_node(type, n, d, 1);
}

SEXP Rf_allocMatrix(SEXPTYPE type, int nrow, int ncol) {
  SEXP s = Rf_allocVector(type, (R_xlen_t)nrow * ncol);
  s->nrow = nrow; s->ncol = ncol;
  return s;
}

SEXP Rf_mkChar(const char *str) {
  char *c = strdup(str);
  return new_node(CHARSXP, (R_xlen_t)strlen(str), c, 1);
}

SEXP Rf_mkString(const char *str) {
  SEXP s = Rf_allocVector(STRSXP, 1);
  ((SEXP *)s->data)[0] = Rf_mkChar(str);
  return s;
}

SEXP SET_VECTOR_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP *)x->data)[i] = v; return v; }
SEXP VECTOR_ELT(SEXP x, R_xlen_t i) { return ((SEXP *)x->data)[i]; }
void SET_STRING_ELT(SEXP x, R_xlen_t i, SEXP v) { ((SEXP *)x->data)[i] = v; }
SEXP STRING_ELT(SEXP x, R_xlen_t i) { return ((SEXP *)x->data)[i]; }
SEXP Rf_namesgets(SEXP x, SEXP names) { x->names = names; return x; }
SEXP Rf_getAttrib_names(SEXP x) { return x->names ? x->names : R_NilValue; }
R_len_t Rf_length(SEXP x) { return (R_len_t)x->len; }
R_xlen_t Rf_xlength(SEXP x) { return x->len; }
int Rf_nrows(SEXP x) { return x->nrow ? x->nrow : (int)x->len; }
int Rf_ncols(SEXP x) { return x->ncol ? x->ncol : 1; }
SEXP Rf_protect(SEXP x) { return x; }
void Rf_unprotect(int) {}
SEXP Rf_ScalarInteger(int v) { SEXP s = Rf_allocVector(INTSXP, 1); INTEGER(s)[0] = v; return s; }
SEXP Rf_ScalarReal(double v) { SEXP s = Rf_allocVector(REALSXP, 1); REAL(s)[0] = v; return s; }

SEXP rshim_wrap_real(double *p, R_xlen_t n) { return new_node(REALSXP, n, p, 0); }
SEXP rshim_wrap_int(int *p, R_xlen_t n) { return new_node(INTSXP, n, p, 0); }
SEXP rshim_wrap_real_matrix(double *p, int nrow, int ncol) {
  SEXP s = new_node(REALSXP, (R_xlen_t)nrow * ncol, p, 0); s->nrow = nrow; s->ncol = ncol; return s;
}
SEXP rshim_wrap_int_matrix(int *p, int nrow, int ncol) {
  SEXP s = new_node(INTSXP, (R_xlen_t)nrow * ncol, p, 0); s->nrow = nrow; s->ncol = ncol; return s;
}

SEXP rshim_list_get(SEXP list, const char *name) {
  if (!list || list->type != VECSXP || !list->names) return nullptr;
  SEXP nm = list->names;
  for (R_xlen_t i = 0; i < list->len && i < nm->len; i++) {
    SEXP c = ((SEXP *)nm->data)[i];  // names may be a VECSXP or STRSXP of CHARSXP
    if (c && c->type == CHARSXP && strcmp((const char *)c->data, name) == 0) return ((SEXP *)list->data)[i];
  }
  return nullptr;
}

void rshim_reset(void) {
  std::lock_guard<std::mutex> lk(g_mu);
  for (SEXP s : g_nodes) { if (s->owns && s->data) free(s->data); free(s); }
  g_nodes.clear();
  for (void *b : g_blocks) free(b);
  g_blocks.clear();
}

void rshim_set_quiet(int q) { g_quiet = q; }
const char *rshim_last_error(void) { return g_last_error.c_str(); }

// ---------------- memory / console / errors ----------------
char *R_alloc(size_t n, int size) {
  size_t bytes = n * (size_t)size;
  void *b = malloc(bytes ? bytes : 1);
  if (!b) Rf_error("rshim: R_alloc of %ld bytes failed", (long)bytes);
  std::lock_guard<std::mutex> lk(g_mu);
  g_blocks.push_back(b);
  return (char *)b;
}

void Rprintf(const char *fmt, ...) {
  if (g_quiet) return;
  va_list ap; va_start(ap, fmt); vfprintf(stdout, fmt, ap); va_end(ap);
}
void REprintf(const char *fmt, ...) {
  va_list ap; va_start(ap, fmt); vfprintf(stderr, fmt, ap); va_end(ap);
}
void R_FlushConsole(void) { fflush(stdout); }
void R_CheckUserInterrupt(void) {}
// Runs fun(data) in a context that a jump to top level cannot leave; TRUE when fun returned normally.  The stand-in has
// no interrupts to deliver; rshim_raise_interrupt() makes the next call report one (tests).
static int g_pending_interrupt = 0;
void rshim_raise_interrupt(void) { g_pending_interrupt = 1; }
Rboolean R_ToplevelExec(void (*fun)(void *), void *data) {
  if (g_pending_interrupt) {
    g_pending_interrupt = 0;
    return FALSE;
  }
  fun(data);
  return TRUE;
}

void Rf_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  g_last_error = buf;
  if (g_jmp) longjmp(*g_jmp, 1);
  fprintf(stderr, "rshim Rf_error (no handler): %s\n", buf);
  abort();
}

void Rf_warning(const char *fmt, ...) {
  char buf[1024];
  va_list ap; va_start(ap, fmt); vsnprintf(buf, sizeof buf, fmt, ap); va_end(ap);
  if (!g_quiet) fprintf(stderr, "Warning: %s\n", buf);
}

// ---------------- .Call trampoline ----------------
typedef SEXP (*F7)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP);
#define A(i) a[i]
int rshim_call(void *fn, int nargs, SEXP *a, SEXP *out) {
  jmp_buf jb;
  jmp_buf *prev = g_jmp;
  g_jmp = &jb;
  g_last_error.clear();
  if (setjmp(jb)) { g_jmp = prev; return 1; }
  SEXP r = nullptr;
  switch (nargs) {
    case 7: r = ((F7)fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6)); break;
    case 8: r = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7)); break;
    case 9: r = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8)); break;
    case 14: r = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(
                 A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13)); break;
    case 17: r = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(
                 A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16)); break;
    case 18: r = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(
                 A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17)); break;
    case 28: r = ((SEXP(*)(SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP,
                           SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP, SEXP))fn)(
                 A(0), A(1), A(2), A(3), A(4), A(5), A(6), A(7), A(8), A(9), A(10), A(11), A(12), A(13), A(14), A(15), A(16), A(17), A(18), A(19),
                 A(20), A(21), A(22), A(23), A(24), A(25), A(26), A(27)); break;
    default:
      g_jmp = prev;
      g_last_error = "rshim_call: unsupported argument count";
      return 2;
  }
  g_jmp = prev;
  if (out) *out = r;
  return 0;
}
#undef A

// ---------------- sorting ----------------
static inline int rcmp_nalast(double x, double y) {
  int nax = std::isnan(x), nay = std::isnan(y);
  if (nax && nay) return 0;
  if (nax) return 1;
  if (nay) return -1;
  if (x < y) return -1;
  if (x > y) return 1;
  return 0;
}

void rsort_with_index(double *x, int *indx, int n) {
  double v;
  int i, j, h, iv;
  for (h = 1; h <= n / 9; h = 3 * h + 1) {}
  for (; h > 0; h /= 3)
    for (i = h; i < n; i++) {
      v = x[i]; iv = indx[i];
      j = i;
      while (j >= h && rcmp_nalast(x[j - h], v) > 0) {
        x[j] = x[j - h]; indx[j] = indx[j - h]; j -= h;
      }
      x[j] = v; indx[j] = iv;
    }
}

// ---------------- RNG ----------------
void rshim_set_seed(unsigned int seed) {
  uint32_t s = seed;
  for (int j = 0; j < 50; j++) s = 69069u * s + 1u;
  for (int j = 0; j < 625; j++) { s = 69069u * s + 1u; g_dummy[j] = s; }
  g_dummy[0] = MT_N;  // FixupSeeds: force regeneration on first draw
  g_seeded = true;
}

void GetRNGstate(void) { if (!g_seeded) rshim_set_seed(4357u); }
void PutRNGstate(void) {}

double unif_rand(void) {
  if (!g_seeded) rshim_set_seed(4357u);
  return fixup(mt_genrand());
}

double norm_rand(void) {
  const double BIG = 134217728.0;  // 2^27: two uniforms give 53 bits
  double u = unif_rand();
  u = (int)(BIG * u) + unif_rand();
  return qnorm(u / BIG, 0.0, 1.0, 1, 0);
}

double exp_rand(void) { return -log(unif_rand()); }

double rnorm(double mu, double sigma) {
  if (sigma == 0.0) return mu;
  return mu + sigma * norm_rand();
}

double runif(double a, double b) {
  if (a == b) return a;
  double u;
  do { u = unif_rand(); } while (u <= 0 || u >= 1);
  return a + (b - a) * u;
}

// dynload stubs (so an init.cpp can be compiled)
typedef struct _DllInfo DllInfo;
int R_registerRoutines(DllInfo *, const void *, const void *, const void *, const void *) { return 1; }
int R_useDynamicSymbols(DllInfo *, int) { return 1; }
int R_forceSymbols(DllInfo *, int) { return 1; }

}  // extern "C"

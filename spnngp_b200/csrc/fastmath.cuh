// fastmath.cuh -- branch-free FP64 sqrt and exp for the covariance build.
//
// The library sqrt()/exp() carry slow-path branches (BSSY/BSYNC regions in SASS) that stop ptxas from
// interleaving the independent correlation evaluations of one lane; with ~1 warp per scheduler the
// build phase then runs latency-bound (ncu, profiles/).  These versions are straight-line code:
//   nngp_sqrt : MUFU.RSQ64H seed (rsqrt.approx.ftz.f64, ~2^-22) + two coupled Goldschmidt steps +
//               one residual correction -> within 1 ulp for normal positive inputs, exact 0 at 0.
//   nngp_exp_nonpos : x <= 0 only; Cody-Waite reduction with a two-part ln2, degree-13 Taylor
//               polynomial on |r| <= ln2/2, exponent insertion; flushes to 0 below -708 (results
//               < 3e-308 never matter for a correlation).  <= 1 ulp-ish (validated on the host build
//               against libm in tests/test_fastmath_host.py).
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#include "fasttab.cuh"

#if defined(__CUDACC__)
#define NNGP_FM __device__ __forceinline__
#else
#define NNGP_FM inline
#endif

namespace nngp {

// 1/13!, 1/12!, ..., 1/2!, 1, 1: Taylor coefficients of exp, highest degree first
#if defined(__CUDACC__)
static __constant__ double kExpPoly[14] = {
#else
static const double kExpPoly[14] = {
#endif
    1.6059043836821613e-10, 2.08767569878681e-09,  2.505210838544172e-08, 2.755731922398589e-07, 2.7557319223985893e-06,
    2.48015873015873e-05,   1.984126984126984e-04, 1.3888888888888889e-03, 8.333333333333333e-03, 4.1666666666666664e-02,
    1.6666666666666666e-01, 0.5, 1.0, 1.0};

NNGP_FM double nngp_sqrt(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
  double y = (double)(1.0f / sqrtf((float)x));  // host stand-in for the ~22-bit hardware seed
#endif
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double d = fma(-g, g, x);
  g = fma(d, h, g);
  return (x > 0.0) ? g : 0.0;
}

// sqrt for d2 >= 0 that needs no select: adding 1e-300 is exact for every normal d2 and turns an
// exact zero into 1e-150, for which exp(-phi D) and the spherical polynomial still round to exactly 1.
NNGP_FM double nngp_sqrt_eps(double x) {
  x += 1e-300;
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
  double y = (double)(1.0f / sqrtf((float)x));
  if (x < 1e-30) y = 1.0 / sqrt(x);  // host stand-in only: float underflows where the hardware seed does not
#endif
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  const double d = fma(-g, g, x);
  return fma(d, h, g);
}

NNGP_FM double nngp_exp_nonpos(double x) {
  const double L2E = 1.4426950408889634074;
  const double LN2_HI = 6.93147180369123816490e-01, LN2_LO = 1.90821492927058770002e-10;
  const double SHIFTER = 6755399441055744.0;  // 1.5 * 2^52: rounds to nearest integer in the low bits
  const double t = fma(x, L2E, SHIFTER);
  const double n = t - SHIFTER;
  double r = fma(n, -LN2_HI, x);
  r = fma(n, -LN2_LO, r);
  // Horner with the coefficients as constant-bank operands (no per-use immediate materialisation)
  double p = kExpPoly[0];
#pragma unroll
  for (int c = 1; c < 14; ++c) p = fma(p, r, kExpPoly[c]);
#if defined(__CUDA_ARCH__)
  const int ni = __double2loint(t);
  const double s = __hiloint2double((ni + 1023) << 20, 0);
#else
  const int ni = (int)n;
  const uint64_t bits = (uint64_t)(ni + 1023) << 52;
  double s;
  memcpy(&s, &bits, sizeof s);
#endif
  return (x < -708.0) ? 0.0 : p * s;
}


// 1/sqrt(x) for positive normal x, straight-line (hardware seed + two coupled Goldschmidt steps), ~1 ulp.
// Used for the Cholesky pivots: the library rsqrt() carries special-case branches that lengthen the
// dependent chain of the latency-bound solve.
NNGP_FM double nngp_rsqrt_pos(double x) {
#if defined(__CUDA_ARCH__)
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
#else
  double y = (double)(1.0f / sqrtf((float)x));
#endif
  double g = x * y, h = 0.5 * y;
  double r = fma(-g, h, 0.5);
  g = fma(g, r, g);
  h = fma(h, r, h);
  r = fma(-g, h, 0.5);
  h = fma(h, r, h);
  return h + h;
}

// 1/x for positive normal x, straight-line (hardware seed + two Newton steps + residual), < 1 ulp
NNGP_FM double nngp_rcp_pos(double x) {
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
#else
  double r = (double)(1.0f / (float)x);
#endif
  double e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  r = fma(r, e, r);
  e = fma(-x, r, 1.0);
  return fma(r, e, r);
}

// natural log for positive normal x, branch-free (fdlibm's e_log.c scheme: x = 2^k (1+f), s = f/(2+f),
// log(1+f) = f - hfsq + s (hfsq + R(s^2)); the division is a MUFU.RCP64H seed + two Newton steps).
NNGP_FM double nngp_log_pos(double x) {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10;
  const double Lg1 = 6.666666666666735130e-01, Lg2 = 3.999999999940941908e-01, Lg3 = 2.857142874366239149e-01,
               Lg4 = 2.222219843214978396e-01, Lg5 = 1.818357216161805012e-01, Lg6 = 1.531383769920937332e-01,
               Lg7 = 1.479819860511658591e-01;
#if defined(__CUDA_ARCH__)
  int hx = __double2hiint(x);
  const int lx = __double2loint(x);
#else
  uint64_t bits;
  memcpy(&bits, &x, sizeof bits);
  int hx = (int)(bits >> 32);
  const int lx = (int)(bits & 0xffffffffu);
#endif
  int k = (hx >> 20) - 1023;
  hx &= 0x000fffff;
  const int i = (hx + 0x95f64) & 0x100000;  // 1 when the mantissa exceeds sqrt(2): use m/2, k+1
  hx |= (i ^ 0x3ff00000);
  k += (i >> 20);
#if defined(__CUDA_ARCH__)
  const double m = __hiloint2double(hx, lx);
#else
  const uint64_t mb = ((uint64_t)(uint32_t)hx << 32) | (uint32_t)lx;
  double m;
  memcpy(&m, &mb, sizeof m);
#endif
  const double f = m - 1.0;
  const double den = 2.0 + f;
#if defined(__CUDA_ARCH__)
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(den));
#else
  double r = (double)(1.0f / (float)den);
#endif
  double e = fma(-den, r, 1.0);
  r = fma(r, e, r);
  e = fma(-den, r, 1.0);
  r = fma(r, e, r);
  double sv = f * r;
  sv = fma(fma(-den, sv, f), r, sv);  // one residual correction: s = f/den to < 1 ulp
  const double dk = (double)k;
  const double z = sv * sv, w = z * z;
  const double t1 = w * fma(w, fma(w, Lg6, Lg4), Lg2);
  const double t2 = z * fma(w, fma(w, fma(w, Lg7, Lg5), Lg3), Lg1);
  const double R = t2 + t1;
  const double hfsq = 0.5 * f * f;
  return dk * ln2_hi - ((hfsq - (sv * (hfsq + R) + dk * ln2_lo)) - f);
}

// ---- table-driven variants (tables of fasttab.cuh, staged in shared memory by the caller) --------------------
// Same contracts as nngp_exp / nngp_log_pos with ~40% fewer FP64 instructions (12 vs 19, 11 vs 28): the Matern
// producers are bound by the FP64 issue rate, each correlation needs one log and one exp (t^nu).
//   exp: x = (32 k + j) ln2/32 + r, |r| <= ln2/64:  2^k * T[j] * (1 + r + ... + r^6/720)       (truncation 5e-20)
//   log: x = 2^k m, m in [1,2), i = top 6 mantissa bits: r = m * rc_i - 1, |r| <= 2^-7,
//        log x = k ln2 + (-log rc_i) + log1p(r), log1p by its degree-8 Taylor polynomial        (truncation 2e-18)
// Absolute error ~1e-16 (max of the result's ulp and 1e-16); the log is NOT relatively accurate next to x = 1,
// which the callers (t^nu = exp(nu log t), log F summed over locations) do not need.
// Constants of the two table-driven kernels whose low 32 bits are not zero: as literals each costs two UMOV / IMAD.MOV
// per use (seen in SASS: 42 UMOV per batch of four Matern correlations); from the constant bank they are plain
// operands of the DFMA.
#if defined(__CUDACC__)
static __constant__ double kTabC[20] = {
#else
static const double kTabC[20] = {
#endif
    46.166241308446828384,                  // 0: 32 / ln 2
    -6.93147180369123816490e-01 / 32.0,     // 1: -ln2_hi / 32
    -1.90821492927058770002e-10 / 32.0,     // 2: -ln2_lo / 32
    1.0 / 720.0, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0,  // 3..6: exp polynomial
    1.0 / 7.0, -1.0 / 6.0, 1.0 / 5.0, 1.0 / 3.0,      // 7..10: log1p polynomial (-1/8, -1/4, -1/2 are exact immediates)
    6.93147180369123816490e-01, 1.90821492927058770002e-10,  // 11, 12: ln2_hi, ln2_lo
    0.0,
    // 14..19: g^k / k!, k = 6..1, g = ln2 / 32: 2^(f/32) - 1 as a polynomial in f itself (nngp_pow_tab)
    1.434565558413193556893347e-13, 3.973709984549416131798095e-11, 9.172562701824643289546081e-9,
    1.693850972437182005405953e-6, 2.345961982022467893882337e-4, 2.16608493924982909192885e-2};

NNGP_FM double nngp_exp_tab(double x, const double *etab) {
  const double SHIFTER = 6755399441055744.0;
  const double t = fma(x, kTabC[0], SHIFTER);
  const double n = t - SHIFTER;
  double r = fma(n, kTabC[1], x);
  r = fma(n, kTabC[2], r);
  double q = fma(r, kTabC[3], kTabC[4]);
  q = fma(q, r, kTabC[5]);
  q = fma(q, r, kTabC[6]);
  q = fma(q, r, 0.5);
  q = fma(q * r, r, r);  // r + r^2 (1/2 + ...)
#if defined(__CUDA_ARCH__)
  const int ni = __double2loint(t);
#else
  const int ni = (int)n;
#endif
  const double T = etab[ni & 31];
  const double v = fma(T, q, T);
  const int k = ni >> 5;
  // results below 2^-1000 are flushed to zero; decided on the integer side (k), not with an FP64 compare of x
#if defined(__CUDA_ARCH__)
  const bool tiny = k < -1000;
  return __hiloint2double(tiny ? 0 : __double2hiint(v) + (k << 20), tiny ? 0 : __double2loint(v));
#else
  if (k < -1000) return 0.0;
  uint64_t bits;
  memcpy(&bits, &v, sizeof bits);
  bits += (uint64_t)((int64_t)k << 52);
  double res;
  memcpy(&res, &bits, sizeof res);
  return res;
#endif
}

// LSTRIDE: distance, in pairs, between consecutive table entries (8 when the table is replicated per quarter-warp lane
// so that the 8 lanes of one 128-bit shared-memory wavefront never share a bank group; ltab then points at the lane's copy)
template <int LSTRIDE = 1>
NNGP_FM double nngp_log_tab(double x, const double *ltab /* 64 pairs {rc, -log rc} */) {
  const double ln2_hi = kTabC[11], ln2_lo = kTabC[12];
#if defined(__CUDA_ARCH__)
  const int hx = __double2hiint(x);
  const int lx = __double2loint(x);
#else
  uint64_t bits;
  memcpy(&bits, &x, sizeof bits);
  const int hx = (int)(bits >> 32);
  const int lx = (int)(bits & 0xffffffffu);
#endif
  const int k = (hx >> 20) - 1023;
  const int idx = (hx >> 14) & 63;
  const int hm = (hx & 0x000fffff) | 0x3ff00000;
#if defined(__CUDA_ARCH__)
  const double m = __hiloint2double(hm, lx);
  const double2 tb = reinterpret_cast<const double2 *>(ltab)[idx * LSTRIDE];
  const double rc = tb.x, lc = tb.y;
#else
  const uint64_t mb = ((uint64_t)(uint32_t)hm << 32) | (uint32_t)lx;
  double m;
  memcpy(&m, &mb, sizeof m);
  const double rc = ltab[2 * idx * LSTRIDE], lc = ltab[2 * idx * LSTRIDE + 1];
#endif
  const double r = fma(m, rc, -1.0);
  double p = fma(r, -1.0 / 8.0, kTabC[7]);
  p = fma(p, r, kTabC[8]);
  p = fma(p, r, kTabC[9]);
  p = fma(p, r, -1.0 / 4.0);
  p = fma(p, r, kTabC[10]);
  p = fma(p, r, -0.5);
  const double dk = (double)k;
  p = fma(p * r, r, r);  // log1p(r)
  return fma(dk, ln2_hi, lc) + fma(dk, ln2_lo, p);
}

// x^nu for x > 0 with the two table-driven kernels above fused in base 2: x^nu = 2^(y/32), y = 32 nu log2 x =
// A k + B (-log rc_i + log1p(r)) with A = 32 nu, B = 32 nu log2(e) prepared on the host.  The log's k ln2 recombination
// (two-part ln2) and the exp's Cody-Waite reduction by ln2/32 collapse into y = n + f with n = rint(y) and f = y - n
// EXACT, and 2^(f/32) - 1 is a polynomial in f: 13 FP64 instructions from the log1p polynomial on instead of 15.
// Same error as exp(nu * log x): ~1e-16 (|y| ulp) relative.  x = 0 reads as 2^-1023 (x^nu ~ 0 for the callers' nu).
template <int LSTRIDE = 1>
NNGP_FM double nngp_pow_tab(double x, double A, double B, const double *etab, const double *ltab) {
#if defined(__CUDA_ARCH__)
  const int hx = __double2hiint(x);
  const int lx = __double2loint(x);
#else
  uint64_t bits;
  memcpy(&bits, &x, sizeof bits);
  const int hx = (int)(bits >> 32);
  const int lx = (int)(bits & 0xffffffffu);
#endif
  const int k = (hx >> 20) - 1023;
  const int idx = (hx >> 14) & 63;
  const int hm = (hx & 0x000fffff) | 0x3ff00000;
#if defined(__CUDA_ARCH__)
  const double m = __hiloint2double(hm, lx);
  const double2 tb = reinterpret_cast<const double2 *>(ltab)[idx * LSTRIDE];
  const double rc = tb.x, lc = tb.y;
#else
  const uint64_t mb = ((uint64_t)(uint32_t)hm << 32) | (uint32_t)lx;
  double m;
  memcpy(&m, &mb, sizeof m);
  const double rc = ltab[2 * idx * LSTRIDE], lc = ltab[2 * idx * LSTRIDE + 1];
#endif
  const double r = fma(m, rc, -1.0);
  double p = fma(r, -1.0 / 8.0, kTabC[7]);
  p = fma(p, r, kTabC[8]);
  p = fma(p, r, kTabC[9]);
  p = fma(p, r, -1.0 / 4.0);
  p = fma(p, r, kTabC[10]);
  p = fma(p, r, -0.5);
  const double dk = (double)k;
  p = fma(p * r, r, r);  // log1p(r)
  const double y = fma(A, dk, B * (lc + p));
  const double SHIFTER = 6755399441055744.0;
  const double t = y + SHIFTER;
  const double n = t - SHIFTER;
  const double f = y - n;  // exact, |f| <= 1/2
  double q = fma(f, kTabC[14], kTabC[15]);
  q = fma(q, f, kTabC[16]);
  q = fma(q, f, kTabC[17]);
  q = fma(q, f, kTabC[18]);
  q = fma(q, f, kTabC[19]);  // (2^(f/32) - 1) / f
#if defined(__CUDA_ARCH__)
  const int ni = __double2loint(t);
#else
  const int ni = (int)n;
#endif
  const double T = etab[ni & 31];
  const double v = fma(T * f, q, T);
  const int e = ni >> 5;
#if defined(__CUDA_ARCH__)
  const bool tiny = e < -1000;
  return __hiloint2double(tiny ? 0 : __double2hiint(v) + (e << 20), tiny ? 0 : __double2loint(v));
#else
  if (e < -1000) return 0.0;
  uint64_t vb;
  memcpy(&vb, &v, sizeof vb);
  vb += (uint64_t)((int64_t)e << 52);
  double res;
  memcpy(&res, &vb, sizeof res);
  return res;
#endif
}

// exp for any x <= 700 (no overflow handling), flushing to 0 below -708: same kernel as nngp_exp_nonpos
NNGP_FM double nngp_exp(double x) { return nngp_exp_nonpos(x); }

}  // namespace nngp

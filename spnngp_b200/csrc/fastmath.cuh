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

}  // namespace nngp

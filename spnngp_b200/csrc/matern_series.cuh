// matern_series.cuh -- branch-free Matern correlation for x = phi D <= 2, as two polynomials in
// t = x^2/4 = phi^2 d2 / 4 and ONE power t^nu:
//
//   K_nu(x) = pi/2 (I_-nu(x) - I_nu(x)) / sin(nu pi),  I_+-nu(x) = (x/2)^(+-nu) sum_k t^k / (k! Gamma(k+1+-nu))
//   rho(x)  = x^nu K_nu(x) / (2^(nu-1) Gamma(nu)) = sum_k sa[k] t^k  -  t^nu sum_k sb[k] t^k
// (the kernel-parameter copy MaternSeriesParams stores -sb[k], so the last step is one fma without a negation)
//
// with the nu-only coefficients sa, sb prepared on the host (besselk.cuh, matern_prepare).  No square
// root (t comes straight from the squared distance), no data-dependent loop, no Gamma functions on the
// device: 2 Horner chains + one log + one exp, all straight-line FMA code that interleaves across
// entries.  Replaces the Temme iteration for the common case; x > 2 and nu within 1e-3 of an integer
// keep using besselk.cuh (Steed CF2 / Temme).  Error: the two sums cancel like 1/|sin(nu pi)| and the
// series loses ~e^(2x) against K's decay; for x <= 2 that stays below ~1e-13 relative (host test vs mpmath).
#pragma once
#include "besselk.cuh"
#include "fastmath.cuh"

namespace nngp {

// t = (phi^2/4) d2 in (0, 1]; returns rho.  t == 0 gives exactly sa[0] = 1.
template <int NT = kMatSeriesTerms>
NNGP_FM double matern_series(double t, const MaternSeriesParams &mc) {
  const double L = nngp_log_pos(t + 1e-300);
  const double pw = nngp_exp(mc.nu * L);  // t^nu
  double pa = mc.sa[NT - 1], pb = mc.sb[NT - 1];
#pragma unroll
  for (int k = NT - 2; k >= 0; --k) {
    pa = fma(pa, t, mc.sa[k]);
    pb = fma(pb, t, mc.sb[k]);
  }
  return fma(pw, pb, pa);  // MaternSeriesParams::sb holds the NEGATED coefficients
}

// matern_series with the table-driven log / exp (tab: 32 doubles of kExpTab32, then the 64 pairs of kLogTab64, in
// shared memory): 26 fewer FP64 instructions of ~95 per correlation; log and exp fused in base 2 (nngp_pow_tab): 2 more.
#ifndef NNGP_LOGTAB_REP
#define NNGP_LOGTAB_REP 0  // 1: the kernel staged 8 interleaved copies of the log table (one per quarter-warp lane)
#endif
constexpr int kLogTabStride = NNGP_LOGTAB_REP ? 8 : 1;
#ifndef NNGP_POW_FUSED
#define NNGP_POW_FUSED 1  // t^nu by nngp_pow_tab (log and exp fused in base 2) instead of exp_tab(nu * log_tab(t))
#endif

// tab: kExpTab32; ltab: the log table (this lane's copy when replicated)
template <int NT = kMatSeriesTerms>
NNGP_FM double matern_series_tab(double t, const MaternSeriesParams &mc, const double *tab, const double *ltab) {
  // t = 0 (coincident sites) needs no guard here: the table log reads it as 2^-1023, so t^nu ~ 0 and rho = sa[0] = 1
#if NNGP_POW_FUSED
  const double pw = nngp_pow_tab<kLogTabStride>(t, mc.nu32, mc.nu32_log2e, tab, ltab);  // t^nu
#else
  const double pw = nngp_exp_tab(mc.nu * nngp_log_tab<kLogTabStride>(t, ltab), tab);  // t^nu
#endif
  double pa = mc.sa[NT - 1], pb = mc.sb[NT - 1];
#pragma unroll
  for (int k = NT - 2; k >= 0; --k) {
    pa = fma(pa, t, mc.sa[k]);
    pb = fma(pb, t, mc.sb[k]);
  }
  return fma(pw, pb, pa);  // MaternSeriesParams::sb holds the NEGATED coefficients
}

// (A truncated 9-term variant for t <= 1/16 behind a warp-uniform branch was measured and dropped: 0.99 -> 1.33 ms,
// the branch breaks the batch's load / chain / store interleaving.  The short series is therefore a LAUNCH-level
// choice: Cor<NNGP_COV_MATERN_SHORT> (cov.cuh) evaluates kMatShortTerms terms and sends t > t_fast_max to slow(); the
// host picks it when 99 % of the locations keep every pair of their neighbourhood under that bound.)

// The same series with kMatExtTerms terms from the global-memory tables, for 1 < t <= 4 (2 < phi D <= 4; 3 when nu is within ~0.03 of an integer):
// a 32-trip loop instead of the 40-90 divisions-in-the-loop iterations Steed's CF2 needs in that range.
NNGP_HD double matern_series_ext(double t, const MaternConsts &mc) {
  const double pw = exp(mc.nu * log(t));
  double pa = mc.ea[kMatExtTerms - 1], pb = mc.eb[kMatExtTerms - 1];
#pragma unroll 4
  for (int k = kMatExtTerms - 2; k >= 0; --k) {
    pa = fma(pa, t, mc.ea[k]);
    pb = fma(pb, t, mc.eb[k]);
  }
  return fma(-pw, pb, pa);
}

// phi D = x > 2, any nu > 0 (integers included): the integral representation
//   K_nu(x) = sqrt(pi/(2x)) e^-x / Gamma(nu+1/2)  int_0^inf e^-t t^(nu-1/2) (1 + t/(2x))^(nu-1/2) dt
// by the kGlNodes-point Gauss-Laguerre rule for the weight t^(nu-1/2) e^-t (nodes prepared per nu on the host).
// The integrand's only singularity sits at t = -2x, so the rule converges geometrically: 16 nodes give
// 4e-13 relative at x = 2, 2e-15 at x = 3 and ~3e-16 beyond x = 4, for every nu (host test vs mpmath) -- with
// no data-dependent loop, where Steed's continued fraction needs 10-90 iterations with a division each.
//   rho(x) = gl_c x^(nu-1/2) e^-x  sum_i w_i (1 + t_i/(2x))^(nu-1/2)
NNGP_FM double matern_gl(double x, const MaternConsts &mc) {
  const double a = mc.nu - 0.5;
  const double h = 0.5 * nngp_rcp_pos(x);
  double S = 0.0;
#pragma unroll 4
  for (int i = 0; i < kGlNodes; ++i) S = fma(mc.gl_w[i], nngp_exp(a * nngp_log_pos(fma(mc.gl_t[i], h, 1.0))), S);
  return mc.gl_c * nngp_exp(fma(a, nngp_log_pos(x), -x)) * S;
}

}  // namespace nngp

// besselk.cuh -- FP64 modified Bessel function of the second kind K_nu(x) for the Matern
// correlation, written for the device (also compiles as host code for CPU-side validation).
//
// Replaces Rmath's bessel_k_ex as used by spCor (reference src/util.cpp:233-237).  Algorithm:
// Temme's series for x <= 2 and the Steed/Temme continued fraction CF2 for x > 2 give K_mu and
// K_{mu+1} with |mu| <= 1/2, followed by the stable upward recurrence to nu = mu + nl.  Everything
// that depends only on nu (Gamma-function pieces, reciprocals of loop constants) is computed ONCE on
// the host per parameter set (matern_prepare) and handed to the kernel by value, so the device loops
// contain no divisions except the one CF2 needs.
#pragma once
#include <cmath>
#include <math.h>
#include <string.h>

#if defined(__CUDACC__)
#define NNGP_HD __host__ __device__ __forceinline__
#else
#define NNGP_HD inline
#endif

namespace nngp {

constexpr int kMatSeriesTerms = 14;  // t^k/(k!)^2-like decay: 1e-17 relative at t = 1 by k = 13
#ifndef NNGP_MAT_SHORT_TERMS
#define NNGP_MAT_SHORT_TERMS 9
#endif
constexpr int kMatShortTerms = NNGP_MAT_SHORT_TERMS;    // truncated series for small arguments (dense neighbourhoods), see t_short_max
#ifndef NNGP_MAT_TINY_TERMS
#define NNGP_MAT_TINY_TERMS 5
#endif
constexpr int kMatTinyTerms = NNGP_MAT_TINY_TERMS;      // a still shorter one for very dense neighbourhoods (bf_tri_kernel only), see t_tiny_max
constexpr int kMatExtTerms = 32;     // 6.25^k / (k!)^2 < 1e-19 by k = 31
constexpr int kTemmeTerms = 24;  // (x^2/4)^i / i! < 1e-18 well before i = 24 for x <= 2
constexpr int kCf2Tab = 64;      // tabulated reciprocals for the first CF2 iterations
constexpr int kGlNodes = 16;     // generalised Gauss-Laguerre rule for phi D > 2 (see matern_gl)

// The part of the constants the straight-line series path needs; small enough to travel as kernel
// parameters (constant bank, no upload, static operand addresses).
struct MaternSeriesParams {
  double nu;
  double nu32, nu32_log2e;  // 32 nu and 32 nu log2(e): t^nu = 2^(y/32), y = 32 nu log2 t (nngp_pow_tab)
  double t_ext_max;  // largest t = (phi D / 2)^2 served by the extended series (beyond: Gauss-Laguerre, matern_gl)
  double t_fast_max; // largest t the straight-line path serves: 1 for the full series, t_short_max for the short one
  int use_series;
  // high word of the squared distance d2 = t_fast_max / (phi^2/4) from which on an entry MAY need slow(): positive
  // doubles order like their bit patterns, so the producers' per-batch test is one integer compare (matern_set_phi)
  int d2_slow_hi;
  double sa[kMatSeriesTerms];
  double sb[kMatSeriesTerms];  // NEGATED coefficients of the second sum
};

struct MaternConsts {
  double nu;      // smoothness
  double mu;      // nu - nl, in [-1/2, 1/2]
  int nl;         // number of upward recurrence steps
  double gam1;    // (1/Gamma(1-mu) - 1/Gamma(1+mu)) / (2 mu)
  double gam2;    // (1/Gamma(1-mu) + 1/Gamma(1+mu)) / 2
  double gampl;   // 1/Gamma(1+mu)
  double gammi;   // 1/Gamma(1-mu)
  double fact;    // pi mu / sin(pi mu)
  double pref;    // 1 / (2^(nu-1) Gamma(nu)): Matern normalisation
  double a1;      // 1/4 - mu^2
  double r_i2[kTemmeTerms];  // 1/(i^2 - mu^2), i = 1..
  double r_im[kTemmeTerms];  // 1/(i - mu)
  double r_ip[kTemmeTerms];  // 1/(i + mu)
  double r_i[kCf2Tab];       // 1/i (shared by both loops)
  double r_a[kCf2Tab];       // 1/a_i of CF2, a_i = mu^2 - 1/4 - i(i-1)
  // Power-series form of the Matern correlation for x <= 2 (t = x^2/4 <= 1), see matern_series():
  //   rho = sum_k sa[k] t^k - t^nu sum_k sb[k] t^k,
  //   sa[k] = Gamma(1-nu) / (k! Gamma(k+1-nu)),  sb[k] = Gamma(1-nu) / (k! Gamma(k+1+nu)).
  // Only used when nu is not within 1e-3 of an integer (the two sums cancel there): use_series.
  int use_series;
  double sa[kMatSeriesTerms];
  double sb[kMatSeriesTerms];
  // the same series carried to kMatExtTerms terms: valid for 2 < phi D <= 5 (1 < t <= 6.25), where the
  // cancellation against K's decay still leaves ~1e-12; evaluated by a short loop (matern_series_ext)
  double ea[kMatExtTerms];
  double eb[kMatExtTerms];
  // largest t for which the first kMatShortTerms terms of both sums are enough: the dropped terms
  // sum_{k >= kMatShortTerms} (|ea_k| + t^nu |eb_k|) t^k stay below 1e-16 (absolute; rho is O(1) there)
  double t_short_max;
  double t_tiny_max;  // the same bound for the first kMatTinyTerms terms
  // Gauss-Laguerre rule for the weight t^(nu-1/2) e^-t (nodes, weights normalised to sum 1) and
  // gl_c = sqrt(pi/2) / (2^(nu-1) Gamma(nu)): the large-argument path, see matern_gl()
  double gl_t[kGlNodes];
  double gl_w[kGlNodes];
  double gl_c;
};

// Nodes and normalised weights of the n-point Gauss-Laguerre rule for t^a e^-t, a > -1 (Golub-Welsch:
// eigenvalues of the Jacobi matrix by implicit QL; weight = squared first component of the eigenvector).
// Real: the iteration runs in double (nodes to ~1e-15 of the largest, far below the 4e-13 truncation error of the
// 16-node rule it feeds, matern_gl) -- the x87 long double version cost 15 us per new smoothness, which the sampler
// pays on every iteration's critical path; long double remains available for the host tests.
template <typename Real>
inline void gauss_laguerre_rule_t(int n, Real a, double *t, double *w) {
  Real d[64], e[64], z[64];
  const Real eps = (sizeof(Real) > 8) ? (Real)1e-22L : (Real)1.2e-16;
  for (int i = 0; i < n; i++) {
    d[i] = (Real)2 * i + a + (Real)1;
    e[i] = (i + 1 < n) ? std::sqrt((Real)(i + 1) * ((Real)(i + 1) + a)) : (Real)0;  // couples i, i+1
    z[i] = (i == 0) ? (Real)1 : (Real)0;
  }
  for (int l = 0; l < n; l++) {
    for (int iter = 0; iter < 300; iter++) {
      int mm = l;
      for (; mm < n - 1; mm++) {
        const Real dd = std::fabs(d[mm]) + std::fabs(d[mm + 1]);
        if (std::fabs(e[mm]) <= eps * dd) break;
      }
      if (mm == l) break;
      Real g = (d[l + 1] - d[l]) / ((Real)2 * e[l]);
      Real r = std::sqrt(g * g + (Real)1);
      g = d[mm] - d[l] + e[l] / (g + std::copysign(r, g));
      Real sn = 1, cs = 1, pp = 0;
      int i = mm - 1;
      for (; i >= l; i--) {
        Real f = sn * e[i];
        const Real b = cs * e[i];
        r = std::sqrt(f * f + g * g);  // entries are O(n): no overflow guard needed (hypot is 5x slower)
        e[i + 1] = r;
        if (r == (Real)0) {
          d[i + 1] -= pp;
          e[mm] = 0;
          break;
        }
        sn = f / r;
        cs = g / r;
        g = d[i + 1] - pp;
        r = (d[i] - g) * sn + (Real)2 * cs * b;
        pp = sn * r;
        d[i + 1] = g + pp;
        g = cs * r - b;
        f = z[i + 1];
        z[i + 1] = sn * z[i] + cs * f;
        z[i] = cs * z[i] - sn * f;
      }
      if (r == (Real)0 && i >= l) continue;
      d[l] -= pp;
      e[l] = g;
      e[mm] = 0;
    }
  }
  for (int i = 0; i < n; i++) {
    t[i] = (double)d[i];
    w[i] = (double)(z[i] * z[i]);
  }
}
inline void gauss_laguerre_rule(int n, long double a, double *t, double *w) {
#ifdef NNGP_GL_LONG_DOUBLE
  gauss_laguerre_rule_t<long double>(n, a, t, w);
#else
  gauss_laguerre_rule_t<double>(n, (double)a, t, w);
#endif
}

// ---- host side: constants per nu -------------------------------------------------------------
// Coefficients c_k of 1/Gamma(z) = sum_{k>=1} c_k z^k (Abramowitz & Stegun 6.1.34); checked
// against mpmath in tests/test_besselk_host.py.
inline void matern_prepare(double nu, MaternConsts *mc) {
  static const double c[26] = {0.0,
                               1.0000000000000000,  0.5772156649015329,  -0.6558780715202538, -0.0420026350340952,
                               0.1665386113822915,  -0.0421977345555443, -0.0096219715278770, 0.0072189432466630,
                               -0.0011651675918591, -0.0002152416741149, 0.0001280502823882,  -0.0000201348547807,
                               -0.0000012504934821, 0.0000011330272320,  -0.0000002056338417, 0.0000000061160950,
                               0.0000000050020075,  -0.0000000011812746, 0.0000000001043427,  0.0000000000077823,
                               -0.0000000000036968, 0.0000000000005100,  -0.0000000000000206, -0.0000000000000054,
                               0.0000000000000014};
  const double pi = 3.14159265358979323846264338327950288;
  mc->nu = nu;
  int nl = (int)(nu + 0.5);
  double mu = nu - nl;
  mc->nl = nl;
  mc->mu = mu;
  double mu2 = mu * mu;
  // Gamma1, Gamma2 from the even / odd part of the series for 1/Gamma(1+-mu) = sum c_k (+-mu)^(k-1)
  double g1 = 0.0, g2 = 0.0, pw = 1.0;  // pw = mu^(2j)
  for (int j = 0; j < 12; j++) {
    g2 += c[2 * j + 1] * pw;  // c1 + c3 mu^2 + ...
    g1 += c[2 * j + 2] * pw;  // c2 + c4 mu^2 + ...
    pw *= mu2;
  }
  mc->gam1 = -g1;
  mc->gam2 = g2;
  // 1/Gamma(1+mu) = g2 - mu*gam1 ... : sum_k c_k mu^(k-1) = (odd k part) + mu*(even k part)
  mc->gampl = g2 + mu * g1;
  mc->gammi = g2 - mu * g1;
  double pimu = pi * mu;
  mc->fact = (fabs(pimu) < 1e-8) ? 1.0 : pimu / sin(pimu);
  mc->pref = 1.0 / (pow(2.0, nu - 1.0) * tgamma(nu));
  mc->a1 = 0.25 - mu2;
  for (int i = 1; i <= kTemmeTerms; i++) {
    mc->r_i2[i - 1] = 1.0 / ((double)i * i - mu2);
    mc->r_im[i - 1] = 1.0 / ((double)i - mu);
    mc->r_ip[i - 1] = 1.0 / ((double)i + mu);
  }
  {
    // series coefficients in extended precision; Gamma(1-nu)/Gamma(k+1-/+nu) built by recurrence from k = 0
    const long double lnu = (long double)nu;
    const long double frac = lnu - roundl(lnu);
    mc->use_series = (fabsl(frac) > 1e-3L && nu < 20.0) ? 1 : 0;
    if (mc->use_series) {
      const long double G = tgammal(1.0L - lnu);
      long double ca = G / tgammal(1.0L - lnu);  // = 1
      long double cb = G / tgammal(1.0L + lnu);
      for (int k = 0; k < kMatExtTerms; k++) {
        if (k < kMatSeriesTerms) {
          mc->sa[k] = (double)ca;
          mc->sb[k] = (double)cb;
        }
        mc->ea[k] = (double)ca;
        mc->eb[k] = (double)cb;
        ca /= (long double)(k + 1) * ((long double)(k + 1) - lnu);  // Gamma(k+2-nu) = (k+1-nu) Gamma(k+1-nu)
        cb /= (long double)(k + 1) * ((long double)(k + 1) + lnu);
      }
    } else {
      for (int k = 0; k < kMatSeriesTerms; k++) mc->sa[k] = mc->sb[k] = 0.0;
      for (int k = 0; k < kMatExtTerms; k++) mc->ea[k] = mc->eb[k] = 0.0;
    }
    mc->t_short_max = mc->t_tiny_max = 0.0;
    if (mc->use_series) {
      auto tail = [&](double t, int first) {
        double sa_t = 0.0, sb_t = 0.0, pw = 1.0;
        for (int k = 0; k < first; k++) pw *= t;
        for (int k = first; k < kMatExtTerms; k++) {
          sa_t += fabs(mc->ea[k]) * pw;
          sb_t += fabs(mc->eb[k]) * pw;
          pw *= t;
        }
        return sa_t + pow(t, nu) * sb_t;
      };
      auto largest_t = [&](int first) {
        double lo = 0.0, hi = 1.0;
        if (tail(hi, first) <= 1e-16) lo = hi;
        for (int it = 0; it < 16 && lo < hi; it++) {  // `lo` always satisfies the bound: 16 halvings resolve it to 1.5e-5
          const double mid = 0.5 * (lo + hi);
          if (tail(mid, first) <= 1e-16) lo = mid;
          else hi = mid;
        }
        return lo;
      };
      mc->t_short_max = largest_t(kMatShortTerms);
      mc->t_tiny_max = largest_t(kMatTinyTerms);
    }
  }
  gauss_laguerre_rule(kGlNodes, (long double)nu - 0.5L, mc->gl_t, mc->gl_w);
  mc->gl_c = 1.2533141373155002512 * mc->pref;  // sqrt(pi/2)
  double a = -mc->a1;
  for (int i = 1; i <= kCf2Tab; i++) {
    mc->r_i[i - 1] = 1.0 / (double)i;
    if (i >= 2) a -= 2.0 * (i - 1);
    mc->r_a[i - 1] = (a != 0.0) ? 1.0 / a : 0.0;
  }
}

// sinh(e)/e without cancellation for small |e|
NNGP_HD double sinhc(double e, double E /* = exp(e) */) {
  if (fabs(e) > 0.35) return 0.5 * (E - 1.0 / E) / e;
  double e2 = e * e;
  // 1 + e2/6 + e2^2/120 + ... up to e^14/15!
  double s = 1.0 / 1307674368000.0;
  s = s * e2 + 1.0 / 6227020800.0;
  s = s * e2 + 1.0 / 39916800.0;
  s = s * e2 + 1.0 / 362880.0;
  s = s * e2 + 1.0 / 5040.0;
  s = s * e2 + 1.0 / 120.0;
  s = s * e2 + 1.0 / 6.0;
  s = s * e2 + 1.0;
  return s;
}

// K_mu(x), K_{mu+1}(x) for x <= 2 (Temme).  logb = log(x/2) is also returned for the caller's x^nu.
NNGP_HD void besselk_temme(double x, const MaternConsts &mc, double *kmu, double *kmu1, double *logb_out) {
  const double b = 0.5 * x;
  const double logb = log(b);
  const double d = -logb;
  const double e = mc.mu * d;
  const double E = exp(e);
  const double fact2 = sinhc(e, E);
  const double ch = 0.5 * (E + 1.0 / E);
  double ff = mc.fact * (mc.gam1 * ch + mc.gam2 * fact2 * d);
  double sum = ff;
  double p = 0.5 * E / mc.gampl;
  double q = 0.5 / (E * mc.gammi);
  double c = 1.0;
  const double dd = b * b;
  double sum1 = p;
#pragma unroll 1
  for (int i = 1; i <= kTemmeTerms; i++) {
    ff = ((double)i * ff + p + q) * mc.r_i2[i - 1];
    c *= dd * mc.r_i[i - 1];
    p *= mc.r_im[i - 1];
    q *= mc.r_ip[i - 1];
    const double del = c * ff;
    sum += del;
    sum1 += c * (p - (double)i * ff);
    if (fabs(del) < fabs(sum) * 1e-17) break;
  }
  *kmu = sum;
  *kmu1 = sum1 * (2.0 / x);
  *logb_out = logb;
}

// K_mu(x), K_{mu+1}(x) for x > 2 (Steed's algorithm for Temme's CF2), including the exp(-x) factor.
NNGP_HD void besselk_cf2(double x, const MaternConsts &mc, double *kmu, double *kmu1) {
  double b = 2.0 * (1.0 + x);
  double d = 1.0 / b;
  double h = d, delh = d;
  double q1 = 0.0, q2 = 1.0;
  const double a1 = mc.a1;
  double q = a1, c = a1;
  double a = -a1;
  double s = 1.0 + q * delh;
#pragma unroll 1
  for (int i = 2; i < 10000; i++) {
    a -= 2.0 * (i - 1);
    const double ri = (i <= kCf2Tab) ? mc.r_i[i - 1] : 1.0 / (double)i;
    const double ra = (i <= kCf2Tab) ? mc.r_a[i - 1] : 1.0 / a;
    c = -a * c * ri;
    const double qnew = (q1 - b * q2) * ra;
    q1 = q2;
    q2 = qnew;
    q += c * qnew;
    b += 2.0;
    d = 1.0 / (b + a * d);
    delh = (b * d - 1.0) * delh;
    h += delh;
    const double dels = q * delh;
    s += dels;
    if (fabs(dels) < fabs(s) * 1e-17) break;
  }
  h = a1 * h;
  const double rk = sqrt(3.14159265358979323846 / (2.0 * x)) * exp(-x) / s;
  *kmu = rk;
  *kmu1 = rk * (mc.mu + x + 0.5 - h) / x;
}

// K_nu(x), x > 0.
NNGP_HD double besselk(double x, const MaternConsts &mc) {
  double k0, k1, logb;
  if (x <= 2.0) besselk_temme(x, mc, &k0, &k1, &logb);
  else besselk_cf2(x, mc, &k0, &k1);
  const double xi2 = 2.0 / x;
  double mu_j = mc.mu;
  for (int j = 1; j <= mc.nl; j++) {  // K_{mu+j+1} = K_{mu+j-1} + 2(mu+j)/x K_{mu+j}
    const double kn = (mu_j + 1.0) * xi2 * k1 + k0;  // at the top of step j: k0=K_{mu+j-1}, k1=K_{mu+j}
    k0 = k1;
    k1 = kn;
    mu_j += 1.0;
  }
  return k0;
}

// Matern correlation rho(x) = x^nu K_nu(x) / (2^(nu-1) Gamma(nu)), x = D*phi; 1 at x <= 0
// (reference src/util.cpp:233-237).
NNGP_HD double matern_cor(double x, const MaternConsts &mc) {
  if (!(x > 0.0)) return 1.0;
  double k0, k1, logx;
  if (x <= 2.0) {
    double logb;
    besselk_temme(x, mc, &k0, &k1, &logb);
    logx = logb + 0.693147180559945309417232121458;
  } else {
    besselk_cf2(x, mc, &k0, &k1);
    logx = log(x);
  }
  const double xi2 = 2.0 / x;
  double mu_j = mc.mu;
  for (int j = 1; j <= mc.nl; j++) {
    const double kn = (mu_j + 1.0) * xi2 * k1 + k0;
    k0 = k1;
    k1 = kn;
    mu_j += 1.0;
  }
  return mc.pref * exp(mc.nu * logx) * k0;
}


inline void matern_series_params(const MaternConsts &mc, MaternSeriesParams *ms) {
  ms->nu = mc.nu;
  ms->nu32 = 32.0 * mc.nu;
  ms->nu32_log2e = (double)(32.0L * (long double)mc.nu * 1.442695040888963407359925L);
  ms->use_series = mc.use_series;
  // cancellation of the two sums grows like e^(2 phi D) / |sin(nu pi)|: keep ~1e-12 absolute
  const double snp = fabs(sin(3.14159265358979323846 * mc.nu));
  ms->t_ext_max = !mc.use_series ? 1.0 : (snp > 0.1 ? 4.0 : 2.25);
  ms->t_fast_max = 1.0;
  ms->d2_slow_hi = 0;  // every entry is checked until matern_set_phi is called
  for (int k = 0; k < kMatSeriesTerms; k++) {
    ms->sa[k] = mc.sa[k];
    ms->sb[k] = -mc.sb[k];  // negated: rho = pa + t^nu * pb
  }
}

// after t_fast_max is final: the integer threshold of the producers' "may need slow()" test for this phi
inline void matern_set_phi(MaternSeriesParams *ms, double phi) {
  ms->d2_slow_hi = 0;
  if (!ms->use_series) return;  // integer nu: every entry takes slow()
  const double d2 = ms->t_fast_max / (0.25 * phi * phi);
  if (!(d2 > 0.0) || d2 > 1e300) return;
  long long bits;
  memcpy(&bits, &d2, sizeof bits);
  ms->d2_slow_hi = (int)(bits >> 32);
}

}  // namespace nngp

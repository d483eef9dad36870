// Host build of the device Bessel-K / Matern routines (spnngp_b200/csrc/besselk.cuh) so their
// numerics can be validated against mpmath on a machine without a GPU.
#include "../../spnngp_b200/csrc/besselk.cuh"
extern "C" {
double host_besselk(double x, double nu) { nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc); return nngp::besselk(x, mc); }
double host_matern(double x, double nu) { nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc); return nngp::matern_cor(x, mc); }
void host_gammas(double nu, double *out) { nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc);
  out[0]=mc.gam1; out[1]=mc.gam2; out[2]=mc.gampl; out[3]=mc.gammi; out[4]=mc.mu; out[5]=mc.pref; }
}
#include "../../spnngp_b200/csrc/matern_series.cuh"
extern "C" {
// Matern correlation through the series path (x <= 2) exactly as the device evaluates it
double host_matern_series(double x, double nu) {
  nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc);
  if (!mc.use_series) return -1.0;
  nngp::MaternSeriesParams ms; nngp::matern_series_params(mc, &ms);
  return nngp::matern_series(0.25 * x * x, ms);
}
double host_nngp_log_pos(double x) { return nngp::nngp_log_pos(x); }
}
extern "C" double host_matern_series_ext(double x, double nu) {
  nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc);
  if (!mc.use_series) return -1.0;
  return nngp::matern_series_ext(0.25 * x * x, mc);
}
// large-argument path (phi D > 2): Gauss-Laguerre quadrature, any nu
extern "C" double host_matern_gl(double x, double nu) {
  nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc);
  return nngp::matern_gl(x, mc);
}
extern "C" void host_gl_rule(double nu, double *t, double *w) {
  nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc);
  for (int i = 0; i < nngp::kGlNodes; i++) { t[i] = mc.gl_t[i]; w[i] = mc.gl_w[i]; }
}
// short series (kMatShortTerms terms) and the largest t it is used for
extern "C" double host_matern_series_short(double t, double nu) {
  nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc);
  if (!mc.use_series) return -1.0;
  nngp::MaternSeriesParams ms; nngp::matern_series_params(mc, &ms);
  return nngp::matern_series<nngp::kMatShortTerms>(t, ms);
}
extern "C" double host_t_short_max(double nu) {
  nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc);
  return mc.t_short_max;
}
// tiny series (kMatTinyTerms terms, bf_tri_kernel) and the largest t it is used for
extern "C" double host_matern_series_tiny(double t, double nu) {
  nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc);
  if (!mc.use_series) return -1.0;
  nngp::MaternSeriesParams ms; nngp::matern_series_params(mc, &ms);
  return nngp::matern_series<nngp::kMatTinyTerms>(t, ms);
}
extern "C" double host_t_tiny_max(double nu) {
  nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc);
  return mc.t_tiny_max;
}
// integer threshold of the producers' "may need slow()" test (matern_set_phi) and the bound it stands for
extern "C" int host_d2_slow_hi(double nu, double phi, int use_short, double *t_fast_max) {
  nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc);
  nngp::MaternSeriesParams ms; nngp::matern_series_params(mc, &ms);
  if (use_short && mc.use_series) ms.t_fast_max = mc.t_short_max;
  nngp::matern_set_phi(&ms, phi);
  *t_fast_max = ms.t_fast_max;
  return ms.d2_slow_hi;
}
// the series exactly as the producers evaluate it: table-driven t^nu (nngp_pow_tab), NT terms
extern "C" double host_matern_series_tab(double t, double nu, int terms) {
  nngp::MaternConsts mc; nngp::matern_prepare(nu, &mc);
  if (!mc.use_series) return -1.0;
  nngp::MaternSeriesParams ms; nngp::matern_series_params(mc, &ms);
  if (terms == nngp::kMatTinyTerms) return nngp::matern_series_tab<nngp::kMatTinyTerms>(t, ms, nngp::kExpTab32, nngp::kLogTab64);
  if (terms == nngp::kMatShortTerms) return nngp::matern_series_tab<nngp::kMatShortTerms>(t, ms, nngp::kExpTab32, nngp::kLogTab64);
  return nngp::matern_series_tab<nngp::kMatSeriesTerms>(t, ms, nngp::kExpTab32, nngp::kLogTab64);
}

// Host build of spnngp_b200/csrc/fastmath.cuh for CPU-side validation against libm.
#include "../../spnngp_b200/csrc/fastmath.cuh"
extern "C" {
double host_nngp_sqrt(double x) { return nngp::nngp_sqrt(x); }
double host_nngp_exp_nonpos(double x) { return nngp::nngp_exp_nonpos(x); }
}
extern "C" double host_nngp_sqrt_eps(double x) { return nngp::nngp_sqrt_eps(x); }
extern "C" double host_nngp_rsqrt_pos(double x) { return nngp::nngp_rsqrt_pos(x); }
extern "C" double host_nngp_rcp_pos(double x) { return nngp::nngp_rcp_pos(x); }

extern "C" double host_nngp_exp_tab(double x) { return nngp::nngp_exp_tab(x, nngp::kExpTab32); }
extern "C" double host_nngp_log_tab(double x) { return nngp::nngp_log_tab(x, nngp::kLogTab64); }
// x^nu through the fused base-2 kernel (A = 32 nu, B = 32 nu log2 e as matern_series_params prepares them)
extern "C" double host_nngp_pow_tab(double x, double nu) {
  const double A = 32.0 * nu, B = (double)(32.0L * (long double)nu * 1.442695040888963407359925L);
  return nngp::nngp_pow_tab(x, A, B, nngp::kExpTab32, nngp::kLogTab64);
}

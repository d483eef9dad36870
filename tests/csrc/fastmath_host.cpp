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

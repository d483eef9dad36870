// cov.cuh -- the four spNNGP correlation functions (reference src/util.cpp:208-245, spCor) as
// compile-time-selected device functors.  Each maps the SQUARED distance between two sites to the
// correlation, so models that do not need the distance itself can skip the square root.
#pragma once
#include "fastmath.cuh"
#include "nngp_common.cuh"

namespace nngp {

template <int COV>
struct Cor;

template <>
struct Cor<NNGP_COV_EXPONENTIAL> {  // exp(-phi D), src/util.cpp:217
  double phi;
  __device__ __forceinline__ Cor(const BfTheta &t) : phi(t.phi) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *) : phi(phi_) {}
  __device__ __forceinline__ double operator()(double d2) const { return nngp_exp_nonpos(-phi * nngp_sqrt_eps(d2)); }
};

template <>
struct Cor<NNGP_COV_SPHERICAL> {  // src/util.cpp:221-227
  double phi, rng;
  __device__ __forceinline__ Cor(const BfTheta &t) : phi(t.phi), rng(1.0 / t.phi) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *) : phi(phi_), rng(1.0 / phi_) {}
  __device__ __forceinline__ double operator()(double d2) const {
    const double D = nngp_sqrt_eps(d2);  // D = 1e-150 at coincident sites: v rounds to exactly 1
    const double pd = phi * D;
    const double v = 1.0 - 1.5 * pd + 0.5 * (pd * pd * pd);
    return (D <= rng) ? v : 0.0;  // D > 0 always; D == rng gives v = 0 as well (reference: either branch)
  }
};

template <>
struct Cor<NNGP_COV_MATERN> {  // src/util.cpp:233-237 with a device Bessel-K (besselk.cuh)
  double phi;
  const MaternConsts *mc;
  __device__ __forceinline__ Cor(const BfTheta &t) : phi(t.phi), mc(t.mat) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *m) : phi(phi_), mc(m) {}
  __device__ __forceinline__ double operator()(double d2) const { return matern_cor(nngp_sqrt(d2) * phi, *mc); }
};

template <>
struct Cor<NNGP_COV_GAUSSIAN> {  // exp(-(phi D)^2), src/util.cpp:240
  double phi2;  // (phi D)^2 = phi^2 d2: no square root needed
  __device__ __forceinline__ Cor(const BfTheta &t) : phi2(t.phi * t.phi) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *) : phi2(phi_ * phi_) {}
  __device__ __forceinline__ double operator()(double d2) const { return nngp_exp_nonpos(-(phi2 * d2)); }
};

}  // namespace nngp

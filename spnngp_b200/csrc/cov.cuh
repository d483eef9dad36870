// cov.cuh -- the four spNNGP correlation functions (reference src/util.cpp:208-245, spCor) as
// compile-time-selected device functors.  Each maps the SQUARED distance between two sites to the
// correlation, so models that do not need the distance itself can skip the square root.
#pragma once
#include "nngp_common.cuh"

namespace nngp {

template <int COV>
struct Cor;

template <>
struct Cor<NNGP_COV_EXPONENTIAL> {  // exp(-phi D), src/util.cpp:217
  double phi;
  __device__ __forceinline__ Cor(const BfTheta &t) : phi(t.phi) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *) : phi(phi_) {}
  __device__ __forceinline__ double operator()(double d2) const { return exp(-phi * sqrt(d2)); }
};

template <>
struct Cor<NNGP_COV_SPHERICAL> {  // src/util.cpp:221-227
  double phi, rng;
  __device__ __forceinline__ Cor(const BfTheta &t) : phi(t.phi), rng(1.0 / t.phi) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *) : phi(phi_), rng(1.0 / phi_) {}
  __device__ __forceinline__ double operator()(double d2) const {
    const double D = sqrt(d2);
    if (D > 0.0 && D <= rng) {
      const double pd = phi * D;
      return 1.0 - 1.5 * pd + 0.5 * (pd * pd * pd);
    }
    return (D >= rng) ? 0.0 : 1.0;
  }
};

template <>
struct Cor<NNGP_COV_MATERN> {  // src/util.cpp:233-237 with a device Bessel-K (besselk.cuh)
  double phi;
  const MaternConsts *mc;
  __device__ __forceinline__ Cor(const BfTheta &t) : phi(t.phi), mc(t.mat) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *m) : phi(phi_), mc(m) {}
  __device__ __forceinline__ double operator()(double d2) const { return matern_cor(sqrt(d2) * phi, *mc); }
};

template <>
struct Cor<NNGP_COV_GAUSSIAN> {  // exp(-(phi D)^2), src/util.cpp:240
  double phi;
  __device__ __forceinline__ Cor(const BfTheta &t) : phi(t.phi) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *) : phi(phi_) {}
  __device__ __forceinline__ double operator()(double d2) const {
    const double pd = phi * sqrt(d2);
    return exp(-(pd * pd));
  }
};

}  // namespace nngp

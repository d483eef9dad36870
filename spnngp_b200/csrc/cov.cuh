// cov.cuh -- the four spNNGP correlation functions (reference src/util.cpp:208-245, spCor) as
// compile-time-selected device functors.  Each maps the SQUARED distance between two sites to the
// correlation, so models that do not need the distance itself can skip the square root.
#pragma once
#include "fastmath.cuh"
#include "matern_series.cuh"
#include "nngp_common.cuh"

namespace nngp {

// Matern constants of the parameter sets of the current launch (one copy per translation unit;
// uploaded stream-ordered by the launchers: FMA operands then come straight from the constant bank).
static __constant__ MaternConsts c_mat[NNGP_MAX_BATCH];

static inline cudaError_t upload_matern_consts(const MaternConsts *h_mat, int nbatch, cudaStream_t stream) {
  if (h_mat == nullptr) return cudaSuccess;
  return cudaMemcpyToSymbolAsync(c_mat, h_mat, (size_t)nbatch * sizeof(MaternConsts), 0, cudaMemcpyHostToDevice, stream);
}

// Every functor maps the SQUARED distance to the correlation.  fast() is straight-line code valid for
// most arguments; needs_slow() flags the rare arguments that must be re-evaluated with slow().
template <int COV>
struct Cor;

template <>
struct Cor<NNGP_COV_EXPONENTIAL> {  // exp(-phi D), src/util.cpp:217
  double phi;
  __device__ __forceinline__ Cor(const BfTheta &t) : phi(t.phi) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *) : phi(phi_) {}
  __device__ __forceinline__ double operator()(double d2) const { return nngp_exp_nonpos(-phi * nngp_sqrt_eps(d2)); }
  __device__ __forceinline__ double fast(double d2) const { return (*this)(d2); }
  __device__ __forceinline__ bool needs_slow(double) const { return false; }
  __device__ __forceinline__ double slow(double d2) const { return (*this)(d2); }
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
  __device__ __forceinline__ double fast(double d2) const { return (*this)(d2); }
  __device__ __forceinline__ bool needs_slow(double) const { return false; }
  __device__ __forceinline__ double slow(double d2) const { return (*this)(d2); }
};

template <>
struct Cor<NNGP_COV_MATERN> {  // src/util.cpp:233-237 with device Bessel-K code (matern_series.cuh, besselk.cuh)
  double phi, q4;  // q4 = phi^2 / 4: t = q4 d2 = (phi D / 2)^2
  int b;           // parameter set of this CTA (index into c_mat)
  __device__ __forceinline__ Cor(const BfTheta &t) : phi(t.phi), q4(0.25 * t.phi * t.phi), b(blockIdx.y) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *) : phi(phi_), q4(0.25 * phi_ * phi_), b(blockIdx.y) {}
  // power series in t for phi D <= 2 (clamped beyond: the caller re-evaluates those with slow())
  __device__ __forceinline__ double fast(double d2) const { return matern_series(fmin(q4 * d2, 1.0), c_mat[b]); }
  __device__ __forceinline__ bool needs_slow(double d2) const { return !c_mat[b].use_series || q4 * d2 > 1.0; }
  // Temme / Steed CF2 (data-dependent loops): phi D > 2, or nu within 1e-3 of an integer
  __device__ __noinline__ double slow(double d2) const { return matern_cor(nngp_sqrt(d2) * phi, c_mat[b]); }
  __device__ __forceinline__ double operator()(double d2) const { return needs_slow(d2) ? slow(d2) : fast(d2); }
};

template <>
struct Cor<NNGP_COV_GAUSSIAN> {  // exp(-(phi D)^2), src/util.cpp:240
  double phi2;  // (phi D)^2 = phi^2 d2: no square root needed
  __device__ __forceinline__ Cor(const BfTheta &t) : phi2(t.phi * t.phi) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *) : phi2(phi_ * phi_) {}
  __device__ __forceinline__ double operator()(double d2) const { return nngp_exp_nonpos(-(phi2 * d2)); }
  __device__ __forceinline__ double fast(double d2) const { return (*this)(d2); }
  __device__ __forceinline__ bool needs_slow(double) const { return false; }
  __device__ __forceinline__ double slow(double d2) const { return (*this)(d2); }
};

}  // namespace nngp

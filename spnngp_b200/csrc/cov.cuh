// cov.cuh -- the four spNNGP correlation functions (reference src/util.cpp:208-245, spCor) as
// compile-time-selected device functors.  Each maps the SQUARED distance between two sites to the
// correlation, so models that do not need the distance itself can skip the square root.
#pragma once
#include "fastmath.cuh"
#include "matern_series.cuh"
#include "nngp_common.cuh"

namespace nngp {

// Kernel-template code of the short-series Matern variant (not an ABI value: the C ABI's cov_model stays 0..3).
constexpr int NNGP_COV_MATERN_SHORT = 4;
constexpr int NNGP_COV_MATERN_TINY = 5;  // kMatTinyTerms terms (bf_tri_kernel only; the other kernels run SHORT instead)
template <int COV>
constexpr bool kIsMatern = (COV == NNGP_COV_MATERN || COV == NNGP_COV_MATERN_SHORT || COV == NNGP_COV_MATERN_TINY);

// Every functor maps the SQUARED distance to the correlation.  fast() is straight-line code valid for
// most arguments; needs_slow() flags the rare arguments that must be re-evaluated with slow().
template <int COV>
struct Cor;

template <>
struct Cor<NNGP_COV_EXPONENTIAL> {  // exp(-phi D), src/util.cpp:217
  double phi;
  const double *tab;  // shared memory: kExpTab32 when the kernel staged it (table-driven exp: ~8 fewer FP64 instructions)
  __device__ __forceinline__ Cor(const BfTheta &t, const double *tab_ = nullptr) : phi(t.phi), tab(tab_) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *) : phi(phi_), tab(nullptr) {}
  __device__ __forceinline__ double operator()(double d2) const { return nngp_exp_nonpos(-phi * nngp_sqrt_eps(d2)); }
  __device__ __forceinline__ double fast(double d2) const { return (*this)(d2); }
  __device__ __forceinline__ double fast_tab(double d2) const { return nngp_exp_tab(-phi * nngp_sqrt_eps(d2), tab); }
  __device__ __forceinline__ bool needs_slow(double) const { return false; }
  __device__ __forceinline__ double slow(double d2) const { return (*this)(d2); }
};

template <>
struct Cor<NNGP_COV_SPHERICAL> {  // src/util.cpp:221-227
  double phi, rng;
  __device__ __forceinline__ Cor(const BfTheta &t, const double * = nullptr) : phi(t.phi), rng(1.0 / t.phi) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *) : phi(phi_), rng(1.0 / phi_) {}
  __device__ __forceinline__ double operator()(double d2) const {
    const double D = nngp_sqrt_eps(d2);  // D = 1e-150 at coincident sites: v rounds to exactly 1
    const double pd = phi * D;
    const double v = 1.0 - 1.5 * pd + 0.5 * (pd * pd * pd);
    return (D <= rng) ? v : 0.0;  // D > 0 always; D == rng gives v = 0 as well (reference: either branch)
  }
  __device__ __forceinline__ double fast(double d2) const { return (*this)(d2); }
  __device__ __forceinline__ double fast_tab(double d2) const { return (*this)(d2); }
  __device__ __forceinline__ bool needs_slow(double) const { return false; }
  __device__ __forceinline__ double slow(double d2) const { return (*this)(d2); }
};

// Matern (src/util.cpp:233-237) with device Bessel-K code (matern_series.cuh, besselk.cuh).  `ms` refers to
// the series coefficients inside the kernel's __grid_constant__ parameter block: when the functor is built
// from P.th[B] with a compile-time B (one-lane kernels) the 28 coefficients are constant-bank operands of the
// FMAs with static addresses; with a run-time index they cost an LDC plus register moves each (~35% of the
// build's instructions in the ncu source view).
template <>
struct Cor<NNGP_COV_MATERN> {
  double phi, q4;  // q4 = phi^2 / 4: t = q4 d2 = (phi D / 2)^2
  const MaternSeriesParams &ms;
  const MaternConsts *mc;  // global memory: Temme / CF2 tables for the slow path
  const double *tab;       // shared memory: exp table (fasttab.cuh) when the kernel staged the tables, else NULL
  const double *ltab;      // shared memory: log table, this lane's copy when replicated (kLogTabStride)
  __device__ __forceinline__ Cor(const BfTheta &t, const double *tab_ = nullptr)
      : phi(t.phi), q4(0.25 * t.phi * t.phi), ms(t.ms), mc(t.mat), tab(tab_),
        ltab(tab_ ? tab_ + 32 + (kLogTabStride > 1 ? 2 * (threadIdx.x & 7) : 0) : nullptr) {}
  // power series in t for phi D <= 2.  Beyond, the value is garbage (possibly inf / NaN, never a trap: the table
  // indices are masked) and the caller re-evaluates the entry with slow(); no clamp, it cost 7 instructions per entry.
  __device__ __forceinline__ double fast_tab(double d2) const { return matern_series_tab(q4 * d2, ms, tab, ltab); }
  __device__ __forceinline__ double fast(double d2) const { return matern_series(q4 * d2, ms); }
  __device__ __forceinline__ bool needs_slow(double d2) const { return !ms.use_series || q4 * d2 > 1.0; }
  // conservative integer form of needs_slow for the per-batch test (true for every entry that needs slow(), and for
  // the few whose d2 shares the threshold's high word): keeps the test off the FP64 pipe
  __device__ __forceinline__ bool maybe_slow(double d2) const { return __double2hiint(d2) >= ms.d2_slow_hi; }
  // phi D in (2, 4]: the extended series; beyond: Gauss-Laguerre quadrature.  nu within 1e-3 of an integer
  // (the series cancel): Temme's iteration for phi D <= 2, the quadrature beyond.
  __device__ __noinline__ double slow(double d2) const {
    const double t = q4 * d2;
    if (t > ms.t_ext_max) return matern_gl(nngp_sqrt(d2) * phi, *mc);
    if (ms.use_series) return matern_series_ext(t, *mc);
    return matern_cor(nngp_sqrt(d2) * phi, *mc);
  }
  __device__ __forceinline__ double operator()(double d2) const { return needs_slow(d2) ? slow(d2) : fast(d2); }
};

// Matern with the series cut to kMatShortTerms terms: exact (dropped terms < 1e-16) for t <= ms.t_fast_max, which the
// host sets to MaternConsts::t_short_max; larger arguments take slow() (extended series / quadrature) as before.
template <>
struct Cor<NNGP_COV_MATERN_SHORT> : Cor<NNGP_COV_MATERN> {
  __device__ __forceinline__ Cor(const BfTheta &t, const double *tab_ = nullptr) : Cor<NNGP_COV_MATERN>(t, tab_) {}
  __device__ __forceinline__ double fast_tab(double d2) const {
    return matern_series_tab<kMatShortTerms>(q4 * d2, ms, tab, ltab);
  }
  __device__ __forceinline__ double fast(double d2) const { return fast_tab(d2); }
  __device__ __forceinline__ bool needs_slow(double d2) const { return q4 * d2 > ms.t_fast_max; }
  __device__ __forceinline__ double operator()(double d2) const { return needs_slow(d2) ? slow(d2) : fast(d2); }
};

// ... and to kMatTinyTerms terms for t <= MaternConsts::t_tiny_max (very dense neighbourhoods: at n = 10^6, m = 15, phi = 6 every
// pair of a neighbourhood has t ~ 3e-4, where the sixth term is already below 1e-19)
template <>
struct Cor<NNGP_COV_MATERN_TINY> : Cor<NNGP_COV_MATERN> {
  __device__ __forceinline__ Cor(const BfTheta &t, const double *tab_ = nullptr) : Cor<NNGP_COV_MATERN>(t, tab_) {}
  __device__ __forceinline__ double fast_tab(double d2) const {
    return matern_series_tab<kMatTinyTerms>(q4 * d2, ms, tab, ltab);
  }
  __device__ __forceinline__ double fast(double d2) const { return fast_tab(d2); }
  __device__ __forceinline__ bool needs_slow(double d2) const { return q4 * d2 > ms.t_fast_max; }
  __device__ __forceinline__ double operator()(double d2) const { return needs_slow(d2) ? slow(d2) : fast(d2); }
};

template <>
struct Cor<NNGP_COV_GAUSSIAN> {  // exp(-(phi D)^2), src/util.cpp:240
  double phi2;  // (phi D)^2 = phi^2 d2: no square root needed
  const double *tab;  // shared memory: kExpTab32 when the kernel staged it
  __device__ __forceinline__ Cor(const BfTheta &t, const double *tab_ = nullptr) : phi2(t.phi * t.phi), tab(tab_) {}
  __device__ __forceinline__ Cor(double phi_, const MaternConsts *) : phi2(phi_ * phi_), tab(nullptr) {}
  __device__ __forceinline__ double operator()(double d2) const { return nngp_exp_nonpos(-(phi2 * d2)); }
  __device__ __forceinline__ double fast(double d2) const { return (*this)(d2); }
  __device__ __forceinline__ double fast_tab(double d2) const { return nngp_exp_tab(-(phi2 * d2), tab); }
  __device__ __forceinline__ bool needs_slow(double) const { return false; }
  __device__ __forceinline__ double slow(double d2) const { return (*this)(d2); }
};

}  // namespace nngp

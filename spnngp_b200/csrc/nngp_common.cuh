// nngp_common.cuh -- shared declarations for the sm_100a NNGP kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/nngp_b200.h"
#include "besselk.cuh"

namespace nngp {

// Closed form of the reference's ragged neighbour layout (src/util.cpp:55-70): location i has
// min(i, m) neighbours stored at this offset.
__host__ __device__ __forceinline__ long long nn_offset(long long i, int m) {
  return (i < m) ? (i * (i - 1)) / 2 : ((long long)m * (m - 1)) / 2 + (i - m) * (long long)m;
}
__host__ __device__ __forceinline__ int nn_count(long long i, int m) { return i < m ? (int)i : m; }

// Result block of one parameter set of a B/F launch: 8 doubles (one 64-byte line of the rank's mailbox slot)
//   [0] sum log F   [1] Q   [2] 1.0 if a factorisation failed   [3] Q of (Bcur, Fcur)   [4] the launch's sequence number,
// written LAST behind a system-scope fence: the host polls it instead of synchronising the stream (nngp_ctx.h: mailbox)
constexpr int kBfOutStride = 8;
constexpr int kBfOutSeq = 4;
constexpr int kResidBetaMax = 32;  // p up to which beta rides in the residual kernel's parameters

// One parameter set of a batched B/F launch.
struct BfTheta {
  double sigmaSq, tauSq, phi;
  const MaternConsts *mat;  // device pointer (Matern only): full constants for the Temme / CF2 slow path
  MaternSeriesParams ms;    // Matern only: series coefficients, read as constant-bank kernel parameters
  double *B;            // nIndx weights (or nullptr: do not store)
  double *F;            // n conditional variances (or nullptr)
  const double *resid;  // n residuals r = X beta - y (or nullptr: no quadratic form)
  // bf_tri_kernel only: resident factors of ANOTHER parameter set (the sampler's current theta); when given, the launch
  // also returns Q(Bcur, Fcur, resid, resid) -- the residual gathers are shared (double-buffered sampler, one launch)
  const double *Bcur;
  const double *Fcur;
  double seq;           // completion stamp of this launch (an integer value), stored to out[kBfOutSeq] after the results
  double *out;          // kBfOutStride doubles: sum log F, Q over [loc_begin, loc_end), 1.0 if a factorisation failed, Q of (Bcur, Fcur)
  double *partials;     // scratch: 2 doubles per CTA (static kernels) or 4 per 32-location tile (dynamic tiles)
  unsigned int *ticket; // zero-initialised, self-resetting
  unsigned int *tile_counter;  // zero-initialised, self-resetting: next tile to hand out (bf_ws_kernel)
  int *fail;            // device: min over failing locations of (i + 1); INT_MAX when none; self-resetting
  int *fail_out;        // where the last CTA publishes *fail (may be mapped host memory)
};

struct BfLaunch {
  const double2 *coords;  // n interleaved (x, y)
  const int *nnIndx;      // reference CSR
  int n, m;
  int loc_begin, loc_end;
  int nbatch;
  // site mode (prediction paths): "locations" are q query sites with unordered m-neighbour lists;
  // th.B receives the kriging weights as a q x m column-major matrix, th.F the q conditional variances
  int site_mode;
  const double2 *sites;  // q interleaved query sites
  const int *nn0;        // q x m column-major neighbour ids (R/generics.R:642 layout)
  int q;
  // Matern: every parameter set of this launch may use a truncated series (th.ms.t_fast_max set accordingly): 1 = the
  // short one (bf_ws_kernel, bf_tri_kernel), 2 = the tiny one (bf_tri_kernel; bf_ws_kernel runs the short one instead);
  // the other kernels ignore the flag (their t_fast_max is 1)
  int matern_short;
  // 1: the launch directly follows the residual kernel of the same evaluation in the stream and may start beside it
  // (programmatic dependent launch; bf_tri_kernel waits with griddepcontrol.wait before its first residual gather)
  int pdl;
  // resident distance triangle (tri_cache.cu) of the locations [loc_begin, loc_end): tile t (32 locations) holds
  // m(m+1)/2 rows of 32 doubles -- m rows location <-> neighbour k, then the packed pairs (k, l), l < k.  NULL: gather kernels.
  const double *tri;
  int tri_cw;      // bf_tri_kernel: weight of a consumer in the producers' shares, in tenths (set by its launcher)
  int tri_want4;   // bf_tri_kernel: four consumers / six stages when they fit (set by its launcher)
  const int *nnT;  // tile-major neighbour ids behind the triangle: tile t, row k, lane -> N(i, k) or -1
  BfTheta th[NNGP_MAX_BATCH];
};

}  // namespace nngp

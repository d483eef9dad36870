// bf_tpl_kernel.cu -- thread-per-location variant of the NNGP hot path for m <= 16 (the common
// n.neighbors = 5..15): every lane of a warp owns one location end to end, so there is no cross-lane
// communication, no predication and no redundant work -- the instruction stream is almost pure
// FP64 math (ncu on the cooperative kernel of bf_kernel.cu showed it issue-bound at 854 warp
// instructions per location with the FP64 pipe only 34% busy; profiles/).
//
// Same mathematics and the same reference sites as bf_kernel.cu:
//   updateBF src/rNNGP.cpp:46-101, updateConjBF src/cNNGP.cpp:24-69, Q src/util.cpp:261-280,
//   and the kriging solve of the prediction paths (src/rNNGPPredict.cpp:154-169, src/cNNGP.cpp:260-275)
//   when launched in site mode.
//
// Per warp (32 locations), all state in shared memory laid out [entry][lane] so that a warp-wide
// access is one conflict-free row:
//   co [MC][32] double2   neighbour sites
//   A  [MC(MC-1)/2][32]   strict lower triangle of C, later its Cholesky factor V
//   cv [MC][32]           c, later y = V^-1 c, later B for the coalesced store
//   dv [MC][32]           1 / V(j,j)
// Phase A: the m(m-1)/2 + m correlations of the lane's location (loop unrolled x4: four independent
//          exp / Bessel chains in flight per lane keep the FP64 pipe busy at ~1 warp per scheduler).
// Phase B: left-looking Cholesky two columns at a time with the accumulators in registers (one shared
//          load feeds two FMAs), c riding along as row MC, then back substitution; fully unrolled for
//          the compile-time capacity MC in {5,8,10,12,15,16} (m < MC pads with identity rows).
// Phase C: B staged through shared memory for coalesced stores, F, and the fused sum log F / Q terms.
#include <limits.h>
#include <stdlib.h>

#include "tpl_device.cuh"

namespace nngp {

template <int MC, int COV>
__global__ void __launch_bounds__(32) bf_tpl_kernel(const __grid_constant__ BfLaunch P) {
  using Geo = TplGeo<MC>;
  __shared__ int last_flag;
  const int lane = threadIdx.x;
  double2 *co = reinterpret_cast<double2 *>(tpl_smem + Geo::OFF_CO) + lane;
  double *A = reinterpret_cast<double *>(tpl_smem + Geo::OFF_A) + lane;
  double *cv = reinterpret_cast<double *>(tpl_smem + Geo::OFF_CV);
  double *dv = reinterpret_cast<double *>(tpl_smem + Geo::OFF_DV) + lane;

  const BfTheta &th = P.th[blockIdx.y];
  const int m = P.m;
  const bool site = P.site_mode != 0;
  const double sigmaSq = th.sigmaSq, tauSq = th.tauSq;
  const double diag = sigmaSq + tauSq;
  const double *__restrict__ resid = th.resid;

  if (m < MC) {  // identity padding rows: their off-diagonals stay zero for the whole launch
    for (int e = m * (m - 1) / 2; e < Geo::NE; ++e) A[e * 32] = 0.0;
    for (int k = m; k < MC; ++k) cv[k * 32 + lane] = 0.0;
  }
  // work list of phase A (see tpl_build)
  __shared__ TplEnt<COV> ent[kTplMaxEnt];
  const int ne = m * (m - 1) / 2;
  constexpr int kBatch = kTplBatch<COV>;
  const int nent_pad = ((ne + m + kBatch - 1) / kBatch) * kBatch;
  for (int e = lane; e < nent_pad; e += 32) {
    unsigned w;
    if (e < ne) {
      int k = 1;
      while ((k + 1) * k / 2 <= e) ++k;  // row of packed entry e = k(k-1)/2 + l
      const int l = e - k * (k - 1) / 2;
      w = (unsigned)k | ((unsigned)l << 8) | ((unsigned)e << 16);
    } else if (e < ne + m) {
      const int k = e - ne;
      w = (unsigned)k | ((unsigned)MC << 8) | ((unsigned)((Geo::OFF_CV - Geo::OFF_A) / 256 + k) << 16);
    } else {
      w = 0u | ((unsigned)MC << 8) | ((unsigned)((Geo::OFF_DV - Geo::OFF_A) / 256) << 16);  // padding
    }
    tpl_pack_ent(w, &ent[e]);
  }
  __syncwarp();

  // location handled by this lane in the pass starting at `b` (tail lanes redo the last location)
  auto loc_of = [&](long long b) -> long long {
    const long long i = b + lane;
    return (i < P.loc_end) ? i : (long long)P.loc_end - 1;
  };
  // neighbour ids of location i into registers (-1 beyond its m_i neighbours)
  auto load_ids = [&](long long i, int *ids) {
    const int mi = site ? m : nn_count(i, m);
    const long long off = site ? 0 : nn_offset(i, m);
#pragma unroll
    for (int k = 0; k < MC; ++k) {
      int id = -1;
      if (k < mi) id = site ? __ldg(&P.nn0[(size_t)k * P.q + i]) : __ldg(&P.nnIndx[off + k]);
      ids[k] = id;
    }
  };
  // asynchronous gather of the neighbour sites straight into shared memory
  // (a missing neighbour -- ragged start of the ordering -- is stood in by the location's own site so that the
  // masked entries stay on the straight-line correlation path)
  auto fetch_sites = [&](const int *ids, long long i) {
#pragma unroll
    for (int k = 0; k < MC; ++k) {
      if (ids[k] >= 0) cp_async16(&co[k * 32], &P.coords[ids[k]]);
      else if (k < m) cp_async16(&co[k * 32], site ? &P.sites[i] : &P.coords[i]);
    }
  };

  const long long stride = (long long)gridDim.x * 32;
  long long base = P.loc_begin + (long long)blockIdx.x * 32;
  int nid[MC];
  double2 nme = make_double2(0.0, 0.0);
  if (base < P.loc_end) {  // pipeline prologue
    const long long i0 = loc_of(base);
    load_ids(i0, nid);
    nme = site ? __ldg(&P.sites[i0]) : __ldg(&P.coords[i0]);
    fetch_sites(nid, i0);
  }

  double qsum = 0.0, lsum = 0.0;
  int failed = INT_MAX;
  for (; base < P.loc_end; base += stride) {
    const bool active = base + lane < P.loc_end;
    const long long i = loc_of(base);
    const int mi = site ? m : nn_count(i, m);
    const long long off = site ? 0 : nn_offset(i, m);
    const double2 me = nme;
    int ids[MC];
#pragma unroll
    for (int k = 0; k < MC; ++k) ids[k] = nid[k];
    // ---- prefetch stage 1: ids and own site of the NEXT pass (registers), hidden behind phase A ----
    const bool has_next = base + stride < P.loc_end;
    if (has_next) {
      const long long in = loc_of(base + stride);
      load_ids(in, nid);
      nme = site ? __ldg(&P.sites[in]) : __ldg(&P.coords[in]);
    }
    cp_async_wait_all();  // this pass's neighbour sites have landed in co
    co[MC * 32] = me;
    // ---- phase A: correlations ------------------------------------------------------------------------
    tpl_build<COV, kTplBatch<COV>>(P, lane, Geo::OFF_CO, Geo::OFF_A, ent, 0, kTplBatch<COV>, nent_pad);
    if (mi < m) {  // ragged start of the ordering (only the first m locations): identity-pad rows >= m_i
      for (int k = mi; k < m; ++k) {
        cv[k * 32 + lane] = 0.0;
        for (int l = 0; l < k; ++l) A[(k * (k - 1) / 2 + l) * 32] = 0.0;
      }
    }
    // ---- prefetch stage 2: next pass's neighbour sites (co is free now), this pass's residuals ----------
    if (has_next) fetch_sites(nid, loc_of(base + stride));
    double rr[MC];
    if (resid != nullptr) {
#pragma unroll
      for (int k = 0; k < MC; ++k) rr[k] = (ids[k] >= 0) ? __ldg(&resid[ids[k]]) : 0.0;
    }
    // ---- phase B: factorise and solve -------------------------------------------------------------------
    double Fsum;
    const bool ok = tpl_solve<MC>(A, cv + lane, dv, mi, diag, &Fsum);
    if (!ok && active) failed = min(failed, (int)min(i + 1, (long long)INT_MAX - 1));
    const double Fv = diag - Fsum;
    // ---- phase C: outputs -----------------------------------------------------------------------------
    double part = 0.0;
    if (resid != nullptr) {
#pragma unroll
      for (int k = 0; k < MC; ++k) part = fma(cv[k * 32 + lane], rr[k], part);
    }
    if (th.B != nullptr) {
      if (site) {
        if (active) {
#pragma unroll
          for (int k = 0; k < MC; ++k)
            if (k < m) th.B[(size_t)k * P.q + i] = cv[k * 32 + lane];
        }
      } else {
        const bool full = (base >= m) && (base + 32 <= P.loc_end);  // 32 complete rows, contiguous in B
        if (full) {
          __syncwarp();
          double *dst = th.B + nn_offset(base, m);
          const float rm = 1.0f / (float)m;
          for (int t = lane; t < 32 * m; t += 32) {
            const int loc = (int)(((float)t + 0.5f) * rm);
            const int k = t - loc * m;
            dst[t] = cv[k * 32 + loc];
          }
          __syncwarp();
        } else if (active) {
#pragma unroll
          for (int k = 0; k < MC; ++k)
            if (k < mi) th.B[off + k] = cv[k * 32 + lane];
        }
      }
    }
    if (th.F != nullptr && active) th.F[i] = Fv;
    if (!site && active) {
      lsum += (Fv > 0.0) ? nngp_log_pos(Fv) : log(Fv);  // straight-line log on the normal path
      if (resid != nullptr) {
        const double e = __ldg(&resid[i]) - part;
        qsum += e * e * nngp_rcp_pos(Fv);
      }
    }
  }

  // ---- grid reduction (one warp per CTA): per-CTA partials, last CTA adds them in fixed order ----------
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    lsum += __shfl_down_sync(0xffffffffu, lsum, s);
    qsum += __shfl_down_sync(0xffffffffu, qsum, s);
  }
  if (failed != INT_MAX) atomicMin(th.fail, failed);
  __syncwarp();
  if (lane == 0) {
    th.partials[2 * blockIdx.x] = lsum;
    th.partials[2 * blockIdx.x + 1] = qsum;
    __threadfence();
    const unsigned int tk = atomicAdd(th.ticket, 1u);
    last_flag = (tk == gridDim.x - 1);
  }
  __syncwarp();
  if (last_flag) {
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned int j = lane; j < gridDim.x; j += 32) {
      a += __ldcg(&th.partials[2 * j]);
      b += __ldcg(&th.partials[2 * j + 1]);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, s);
      b += __shfl_down_sync(0xffffffffu, b, s);
    }
    if (lane == 0) {
      th.out[0] = a;
      th.out[1] = b;
      const int fv = atomicExch(th.fail, INT_MAX);
      th.out[2] = (fv != INT_MAX) ? 1.0 : 0.0;  // summed across ranks so that every rank sees a failure
      th.out[3] = 0.0;
      *th.fail_out = fv;
      *th.ticket = 0u;
      __threadfence_system();  // results before the stamp, as seen by the polling host (and by peers' hosts)
      *reinterpret_cast<volatile double *>(&th.out[kBfOutSeq]) = th.seq;
    }
  }
}

// ---- launcher ---------------------------------------------------------------------------------------
template <int MC, int COV>
static cudaError_t tpl_launch_one(const BfLaunch &P, int grid_x, cudaStream_t stream) {
  using Geo = TplGeo<MC>;
  static bool configured[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(bf_tpl_kernel<MC, COV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Geo::kWarpBytes);
    if (e != cudaSuccess) return e;
    if (dev < 16) configured[dev] = true;
  }
  dim3 grid(grid_x, P.nbatch, 1);
  bf_tpl_kernel<MC, COV><<<grid, 32, Geo::kWarpBytes, stream>>>(P);
  return cudaGetLastError();
}

template <int MC>
static cudaError_t tpl_launch_cov(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream) {
  switch (cov) {
    case NNGP_COV_EXPONENTIAL: return tpl_launch_one<MC, NNGP_COV_EXPONENTIAL>(P, grid_x, stream);
    case NNGP_COV_SPHERICAL: return tpl_launch_one<MC, NNGP_COV_SPHERICAL>(P, grid_x, stream);
    case NNGP_COV_MATERN: return tpl_launch_one<MC, NNGP_COV_MATERN>(P, grid_x, stream);
    case NNGP_COV_GAUSSIAN: return tpl_launch_one<MC, NNGP_COV_GAUSSIAN>(P, grid_x, stream);
    default: return cudaErrorInvalidValue;
  }
}

int tpl_capacity_for_m(int m) {
  static const int caps[6] = {5, 8, 10, 12, 15, 16};
  for (int c : caps)
    if (m <= c) return c;
  return 0;  // not handled by this kernel
}

static int tpl_warp_bytes(int MC) {
  return (MC + 1) * 32 * 16 + (MC * (MC - 1) / 2) * 32 * 8 + 2 * MC * 32 * 8;
}

// one-warp CTAs resident per SM, limited by shared memory (227 KB usable, 1 KB reserved per CTA)
// An SM has 4 warp schedulers and every CTA here is one warp with a fixed share of the locations: a
// count that is not a multiple of 4 leaves one scheduler with an extra warp, and with FP64-saturating
// warps the whole launch then waits for that scheduler (ncu: 5 warps/SM ran at 62% of 4 warps/SM's
// per-warp speed).  So round down to a multiple of 4.  NNGP_TPL_CTAS overrides (profiling).
int tpl_ctas_per_sm(int m) {
  const int MC = tpl_capacity_for_m(m);
  int c = (227 * 1024) / (tpl_warp_bytes(MC) + 1024);
  if (c > 16) c = 16;
  if (c >= 4) c -= c % 4;
  static int override_ctas = -1;
  if (override_ctas < 0) {
    const char *e = getenv("NNGP_TPL_CTAS");
    override_ctas = e ? atoi(e) : 0;
  }
  if (override_ctas > 0) c = override_ctas;
  return c < 1 ? 1 : c;
}

int tpl_grid_for(int m, long long nloc, int sms) {
  long long need = (nloc + 31) / 32;
  long long g = (long long)sms * tpl_ctas_per_sm(m);
  if (need < g) g = need;
  return (int)(g < 1 ? 1 : g);
}

cudaError_t tpl_launch(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream) {
  switch (tpl_capacity_for_m(P.m)) {
    case 5: return tpl_launch_cov<5>(P, cov, grid_x, stream);
    case 8: return tpl_launch_cov<8>(P, cov, grid_x, stream);
    case 10: return tpl_launch_cov<10>(P, cov, grid_x, stream);
    case 12: return tpl_launch_cov<12>(P, cov, grid_x, stream);
    case 15: return tpl_launch_cov<15>(P, cov, grid_x, stream);
    case 16: return tpl_launch_cov<16>(P, cov, grid_x, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace nngp

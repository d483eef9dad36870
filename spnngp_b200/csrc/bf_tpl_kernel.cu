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

#include "cov.cuh"

namespace nngp {

template <int MC>
struct TplGeo {
  static constexpr int NE = MC * (MC - 1) / 2;
  static constexpr int OFF_CO = 0;                          // bytes; row MC holds the location's own site
  static constexpr int OFF_A = OFF_CO + (MC + 1) * 32 * 16;
  static constexpr int OFF_CV = OFF_A + NE * 32 * 8;
  static constexpr int OFF_DV = OFF_CV + MC * 32 * 8;
  static constexpr int kWarpBytes = OFF_DV + MC * 32 * 8;
};

extern __shared__ __align__(16) unsigned char tpl_smem[];

// ---- phase A: correlations ----------------------------------------------------------------------
// Work list shared by all lanes: entry w = k | l << 8 | dst << 16 means "rho between co[k] and co[l]
// goes to row dst (256-byte rows counted from the start of A)".  It holds the m(m-1)/2 pairs of the
// strict lower triangle (dst = packed index), the m entries of c (l = MC: the location's own site,
// dst = a row of cv) and is padded to a multiple of the batch width with harmless duplicates that
// write into dv (overwritten by the factorisation afterwards).
constexpr int kTplMaxEnt = 144;

// Batches of UNR entries: all loads, then UNR independent evaluation chains, then all stores.
// (A plain unrolled loop is serialised by ptxas: it cannot prove that the store of entry e does not
// alias the site loads of entry e+1, so the chains never overlap -- seen in SASS.)
template <int COV>
__device__ __noinline__ void tpl_build(int off_co, int off_a, const unsigned *ent, int nent_pad, double sigmaSq,
                                       double phi, const MaternConsts *mat) {
  constexpr int UNR = (COV == NNGP_COV_MATERN) ? 4 : 8;
  const int lane = threadIdx.x;
  const unsigned char *cob = tpl_smem + off_co + lane * 16;
  unsigned char *ab = tpl_smem + off_a + lane * 8;
  const Cor<COV> cor(phi, mat);
  for (int e0 = 0; e0 < nent_pad; e0 += UNR) {
    double d2[UNR], v[UNR];
    unsigned dst[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      const unsigned w = ent[e0 + u];
      dst[u] = w >> 16;
      const double2 pk = *reinterpret_cast<const double2 *>(cob + (w & 0xffu) * 512u);
      const double2 pl = *reinterpret_cast<const double2 *>(cob + ((w >> 8) & 0xffu) * 512u);
      const double dx = pk.x - pl.x, dy = pk.y - pl.y;
      d2[u] = dx * dx + dy * dy;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) v[u] = sigmaSq * cor.fast(d2[u]);
    if (COV == NNGP_COV_MATERN) {  // rare arguments outside the straight-line path (phi D > 2, integer nu)
      bool any = false;
#pragma unroll
      for (int u = 0; u < UNR; ++u) any |= cor.needs_slow(d2[u]);
      if (any) {
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          if (cor.needs_slow(d2[u])) v[u] = sigmaSq * cor.slow(d2[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) *reinterpret_cast<double *>(ab + dst[u] * 256u) = v[u];
  }
}

// ---- phase B: Cholesky + solves, fully unrolled -----------------------------------------------------
// Left-looking C = V V' two columns at a time; V(i,j) overwrites A, the right-hand side rides along
// as row MC (cv[j] becomes y_j = (V^-1 c)_j), 1/V(j,j) goes to dv[j].  Then V' B = y.
// On return: Bout[0..MC) kriging weights, *Fsum = c' C^-1 c = |y|^2; false if a pivot is not positive.
#define V_(i, j) A[((i) * ((i)-1) / 2 + (j)) * 32]

// columns J0 and J0+1 (compile-time), then recurse; the memory clobber between pairs keeps ptxas from
// hoisting the next pair's loads across this pair's stores (which would blow up register pressure).
template <int MC, int J0>
__device__ __forceinline__ void tpl_chol_pair(double *A, double *cv, double *dv, int mi, double diag, double &Fsum,
                                              bool &ok) {
  constexpr bool two = (J0 + 1 < MC);
  constexpr int J1 = J0 + 1;
  double a0[MC + 1], a1[MC + 1];  // rows J0..MC of columns J0 and J1 (row MC: right-hand side)
#pragma unroll
  for (int i = J0; i <= MC; ++i) {
    if (i == J0) a0[i] = (J0 < mi) ? diag : 1.0;
    else if (i < MC) a0[i] = V_(i, J0);
    else a0[i] = cv[J0 * 32];
  }
  if (two) {
#pragma unroll
    for (int i = J1; i <= MC; ++i) {
      if (i == J1) a1[i] = (J1 < mi) ? diag : 1.0;
      else if (i < MC) a1[i] = V_(i, J1);
      else a1[i] = cv[J1 * 32];
    }
  }
#pragma unroll
  for (int k = 0; k < J0; ++k) {
    const double v0 = V_(J0, k);
    a0[J0] = fma(-v0, v0, a0[J0]);
    double v1 = 0.0;
    if (two) {
      v1 = V_(J1, k);
      a0[J1] = fma(-v1, v0, a0[J1]);
      a1[J1] = fma(-v1, v1, a1[J1]);
    }
#pragma unroll
    for (int i = J0 + (two ? 2 : 1); i <= MC; ++i) {  // remaining rows; row MC = right-hand side
      const double v = (i < MC) ? V_(i, k) : cv[k * 32];
      a0[i] = fma(-v, v0, a0[i]);
      if (two) a1[i] = fma(-v, v1, a1[i]);
    }
  }
  // finalise column J0
  const double d0 = a0[J0];
  if (J0 < mi && !(d0 > 0.0)) ok = false;
  const double r0 = rsqrt(d0);
  dv[J0 * 32] = r0;
#pragma unroll
  for (int i = J0 + 1; i <= MC; ++i) a0[i] *= r0;
  Fsum = fma(a0[MC], a0[MC], Fsum);
  cv[J0 * 32] = a0[MC];
#pragma unroll
  for (int i = J0 + 1; i < MC; ++i) V_(i, J0) = a0[i];
  if (two) {
    const double v = a0[J1];  // V(J1, J0), still in registers
#pragma unroll
    for (int i = J1; i <= MC; ++i) a1[i] = fma(-a0[i], v, a1[i]);
    const double d1 = a1[J1];
    if (J1 < mi && !(d1 > 0.0)) ok = false;
    const double r1 = rsqrt(d1);
    dv[J1 * 32] = r1;
#pragma unroll
    for (int i = J1 + 1; i <= MC; ++i) a1[i] *= r1;
    Fsum = fma(a1[MC], a1[MC], Fsum);
    cv[J1 * 32] = a1[MC];
#pragma unroll
    for (int i = J1 + 1; i < MC; ++i) V_(i, J1) = a1[i];
  }
  asm volatile("" ::: "memory");
  if constexpr (J0 + 2 < MC) tpl_chol_pair<MC, J0 + 2>(A, cv, dv, mi, diag, Fsum, ok);
}

// On return cv[k] holds the kriging weight B_k (k < MC), *Fsum_out = c' C^-1 c; false if a pivot <= 0.
template <int MC>
__device__ __forceinline__ bool tpl_solve(double *A, double *cv, double *dv, int mi, double diag, double *Fsum_out) {
  double Fsum = 0.0;
  bool ok = true;
  tpl_chol_pair<MC, 0>(A, cv, dv, mi, diag, Fsum, ok);
  // back substitution V' B = y: B_k = (y_k - sum_{i>k} V(i,k) B_i) / V(k,k)
  double w[MC];
#pragma unroll
  for (int k = MC - 1; k >= 0; --k) {
    double acc = cv[k * 32];
#pragma unroll
    for (int i = MC - 1; i > k; --i) acc = fma(-V_(i, k), w[i], acc);
    w[k] = acc * dv[k * 32];
  }
#pragma unroll
  for (int k = 0; k < MC; ++k) cv[k * 32] = w[k];
  *Fsum_out = Fsum;
  return ok;
}
#undef V_

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

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
  __shared__ unsigned ent[kTplMaxEnt];
  const int ne = m * (m - 1) / 2;
  constexpr int kBatch = (COV == NNGP_COV_MATERN) ? 4 : 8;
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
    ent[e] = w;
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
  auto fetch_sites = [&](const int *ids) {
#pragma unroll
    for (int k = 0; k < MC; ++k)
      if (ids[k] >= 0) cp_async16(&co[k * 32], &P.coords[ids[k]]);
  };

  const long long stride = (long long)gridDim.x * 32;
  long long base = P.loc_begin + (long long)blockIdx.x * 32;
  int nid[MC];
  double2 nme = make_double2(0.0, 0.0);
  if (base < P.loc_end) {  // pipeline prologue
    const long long i0 = loc_of(base);
    load_ids(i0, nid);
    nme = site ? __ldg(&P.sites[i0]) : __ldg(&P.coords[i0]);
    fetch_sites(nid);
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
    tpl_build<COV>(Geo::OFF_CO, Geo::OFF_A, ent, nent_pad, sigmaSq, th.phi, th.mat);
    if (mi < m) {  // ragged start of the ordering (only the first m locations): identity-pad rows >= m_i
      for (int k = mi; k < m; ++k) {
        cv[k * 32 + lane] = 0.0;
        for (int l = 0; l < k; ++l) A[(k * (k - 1) / 2 + l) * 32] = 0.0;
      }
    }
    // ---- prefetch stage 2: next pass's neighbour sites (co is free now), this pass's residuals ----------
    if (has_next) fetch_sites(nid);
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
      lsum += log(Fv);
      if (resid != nullptr) {
        const double e = __ldg(&resid[i]) - part;
        qsum += e * e / Fv;
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
      *th.fail_out = fv;
      *th.ticket = 0u;
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
  if (COV == NNGP_COV_MATERN) {
    cudaError_t e = upload_matern_consts(P.h_mat, P.nbatch, stream);
    if (e != cudaSuccess) return e;
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
int tpl_ctas_per_sm(int m) {
  const int MC = tpl_capacity_for_m(m);
  int c = (227 * 1024) / (tpl_warp_bytes(MC) + 1024);
  if (c > 16) c = 16;
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

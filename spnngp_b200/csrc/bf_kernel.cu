// bf_kernel.cu -- the NNGP hot path on sm_100a: per-location neighbour covariance build, FP64
// LDL' factorisation with the forward solve fused in, back substitution, and the fused
// log-likelihood reduction (sum log F_i and the quadratic form Q).
//
// Replaces, for every location i of a shard (reference file:line):
//   updateBF      src/rNNGP.cpp:46-101   c_k = s2 rho(|s_i - s_N(i,k)|), C = s2 rho(.) + t2 I,
//                                        B_i = C^-1 c, F_i = s2 + t2 - B_i.c, logDet = sum log F_i
//   updateConjBF  src/cNNGP.cpp:24-69    the same with (s2, t2) = (1, alpha)
//   Q(B,F,r,r)    src/util.cpp:261-280   sum_i (r_i - B_i.r_N(i))^2 / F_i
//
// Design (B200-first, not the reference's LAPACK route):
//   * A group of T = M/2 threads owns one location; thread r holds rows r and M-1-r of the
//     (padded) M x M lower triangle in registers, which balances the triangle (M+1 entries each).
//     M in {8,16,32}; a 128-thread CTA works on 128/T locations at once, grid-stride persistent.
//   * Covariance build: thread r evaluates columns r and M-1-r (again M+1 entries each, so the
//     expensive exp / Bessel-K evaluations are perfectly balanced) plus its two entries of c,
//     with the column site kept in registers and the row site read from shared memory.
//   * Factorisation: right-looking LDL' (no square roots).  Column k travels through a double-
//     buffered shared-memory vector (one __syncwarp per step); every row update is an FMA on
//     registers with compile-time indices.  The right-hand side c rides along as an extra column,
//     so the forward solve and F_i = s2 + t2 - sum_k y_k^2/d_k come out of the same sweep.
//   * Back substitution L' B = D^-1 y with the L columns re-read from the tile and the solved
//     component broadcast by one shuffle per step.
//   * Epilogue: B, F stored coalesced; the quadratic-form term is reduced inside the group by
//     shuffles; per-thread running sums -> CTA -> per-CTA partials -> the last CTA (ticket) adds
//     the partials in fixed order, so results are run-to-run deterministic.
//   No tensor cores: each solve is a tiny FP64 problem (BASELINE.json north_star).
#include <limits.h>
#include <stdlib.h>
#include <string.h>

#include "cov.cuh"

namespace nngp {

template <int M>
struct BfGeo {
  static constexpr int T = M / 2;         // threads per location
  static constexpr int LD = M + 1;        // leading dimension of the tile; column M holds c
  static constexpr int kThreads = 128;
  static constexpr int G = kThreads / T;  // locations per CTA pass
  static constexpr int OFF_NBR = 0;                // double2[M] neighbour sites
  static constexpr int OFF_A = 2 * M;              // double[M * LD] covariance tile, later L
  static constexpr int OFF_COL = OFF_A + M * LD;   // double[2][M + 2] broadcast column (double-buffered)
  static constexpr int RAW = OFF_COL + 2 * (M + 2);
  static constexpr int SLAB = RAW + (RAW & 1) + 2;  // even (16-byte aligned slabs), skewed by one bank pair
  static constexpr size_t kSmemBytes = (size_t)G * SLAB * sizeof(double);
};

template <int M, int COV>
__global__ void __launch_bounds__(128, (M <= 16 ? 4 : 2)) bf_kernel(const __grid_constant__ BfLaunch P) {
  using Geo = BfGeo<M>;
  constexpr int T = Geo::T, LD = Geo::LD, G = Geo::G;
  extern __shared__ __align__(16) double smem[];
  __shared__ double red[2][4];
  __shared__ int last_flag;

  const int tid = threadIdx.x;
  const int r = tid % T;  // row/column pair index inside the location group
  const int g = tid / T;  // location slot inside the CTA
  const int h = M - 1 - r;
  double *slab = smem + (size_t)g * Geo::SLAB;
  double2 *nbr = reinterpret_cast<double2 *>(slab + Geo::OFF_NBR);
  double *A = slab + Geo::OFF_A;
  double *colb = slab + Geo::OFF_COL;

  const BfTheta &th = P.th[blockIdx.y];
  const Cor<COV> cor(th);
  const int m = P.m;
  const bool site = P.site_mode != 0;  // prediction: query sites with unordered neighbour lists
  const double sigmaSq = th.sigmaSq, tauSq = th.tauSq;
  const double *__restrict__ resid = th.resid;

  double qsum = 0.0, lsum = 0.0;
  int failed = INT_MAX;

  for (long long base = P.loc_begin + (long long)blockIdx.x * G; base < P.loc_end;
       base += (long long)gridDim.x * G) {
    const long long i = base + g;
    const bool active = i < P.loc_end;
    const int mi = active ? (site ? m : nn_count(i, m)) : 0;
    const long long off = (active && !site) ? nn_offset(i, m) : 0;

    // ---- 1. gather this thread's two neighbour sites -------------------------------------------
    int idx_lo = -1, idx_hi = -1;
    if (r < mi) idx_lo = site ? __ldg(&P.nn0[(size_t)r * P.q + i]) : __ldg(&P.nnIndx[off + r]);
    if (h < mi) idx_hi = site ? __ldg(&P.nn0[(size_t)h * P.q + i]) : __ldg(&P.nnIndx[off + h]);
    const double2 zero2 = make_double2(0.0, 0.0);
    const double2 me = active ? (site ? __ldg(&P.sites[i]) : __ldg(&P.coords[i])) : zero2;
    // a missing neighbour is stood in by the location's own site: masked entries stay on the cheap path
    const double2 plo = (idx_lo >= 0) ? __ldg(&P.coords[idx_lo]) : me;
    const double2 phi_ = (idx_hi >= 0) ? __ldg(&P.coords[idx_hi]) : me;
    nbr[r] = plo;
    nbr[h] = phi_;
    __syncwarp();

    // ---- 2. covariance entries: columns r and h of C (rows >= column), then c_r and c_h ----------
    {
      const int nlo = max(mi - r, 0), nhi = max(mi - h, 0);
      const int e1 = nlo + nhi;               // matrix entries
      const int e2 = e1 + (r < mi ? 1 : 0);   // + c_r
      const int e3 = e2 + (h < mi ? 1 : 0);   // + c_h
      const int nloop = __reduce_max_sync(0xffffffffu, e3);
#pragma unroll 1
      for (int t = 0; t < nloop; ++t) {
        if (t < e3) {
          const bool is_lo = t < nlo;
          const bool is_mat = t < e1;
          const bool c_is_lo = (t == e1) && (r < mi);  // first c entry is c_r when row r exists
          // site a: the column site (matrix) or the neighbour whose c entry is formed
          const bool a_lo = is_mat ? is_lo : c_is_lo;
          const double2 pa = a_lo ? plo : phi_;
          const int ra = a_lo ? r : h;  // row index of site a
          int row, col;
          double2 pb;
          if (is_mat) {
            row = is_lo ? r + t : h + (t - nlo);
            col = ra;
            pb = nbr[row];
          } else {
            row = ra;
            col = M;  // the c column of the tile
            pb = me;
          }
          const double dx = pb.x - pa.x, dy = pb.y - pa.y;
          const double d2 = dx * dx + dy * dy;
          double v = sigmaSq * cor(d2);
          if (row == col) v += tauSq;
          A[row * LD + col] = v;
        }
      }
    }
    __syncwarp();

    // ---- 3. load the two rows into registers (identity rows beyond m_i pad the system) -----------
    double lo[T], hi[M];
    const bool lo_real = r < mi, hi_real = h < mi;
#pragma unroll
    for (int j = 0; j < T; ++j) lo[j] = (lo_real && j <= r) ? A[r * LD + j] : ((j == r) ? 1.0 : 0.0);
#pragma unroll
    for (int j = 0; j < M; ++j) hi[j] = (hi_real && j <= h) ? A[h * LD + j] : ((j == h) ? 1.0 : 0.0);
    double c_lo = lo_real ? A[r * LD + M] : 0.0;
    double c_hi = hi_real ? A[h * LD + M] : 0.0;
    __syncwarp();  // the tile is reused for L from here on

    // ---- 4. LDL' elimination with the right-hand side riding along ------------------------------
    double Fsum = 0.0, w_lo = 0.0, w_hi = 0.0;
#pragma unroll
    for (int k = 0; k < M; ++k) {
      if (k < m) {  // uniform: steps beyond m only see identity rows
        double *cb = colb + (k & 1) * (M + 2);
        if (k < T) cb[r] = lo[k];
        cb[h] = hi[k];
        if (k < T) {
          if (r == k) cb[M] = c_lo;
        } else {
          if (h == k) cb[M] = c_hi;
        }
        __syncwarp();
        double colv[M];
        const double2 *cb2 = reinterpret_cast<const double2 *>(cb);
#pragma unroll
        for (int pp = k / 2; pp < M / 2; ++pp) {
          const double2 v = cb2[pp];
          colv[2 * pp] = v.x;
          colv[2 * pp + 1] = v.y;
        }
        const double piv = colv[k];
        const double yk = cb[M];
        if (k < mi && !(piv > 0.0)) failed = min(failed, (int)min(i + 1, (long long)INT_MAX - 1));
        const double inv = 1.0 / piv;
        const double wk = yk * inv;
        Fsum = fma(wk, yk, Fsum);
        if (k < T) {
          if (r == k) w_lo = wk;
        } else {
          if (h == k) w_hi = wk;
        }
        if (k < T - 1) {  // rows r > k of the low half
          const double t = lo[k] * inv;
          c_lo = fma(-t, yk, c_lo);
#pragma unroll
          for (int j = k + 1; j < T; ++j) lo[j] = fma(-t, colv[j], lo[j]);
          if (r > k) A[r * LD + k] = t;
        }
        if (k < M - 1) {  // rows h > k of the high half
          const double t = hi[k] * inv;
          c_hi = fma(-t, yk, c_hi);
#pragma unroll
          for (int j = k + 1; j < M; ++j) hi[j] = fma(-t, colv[j], hi[j]);
          if (h > k) A[h * LD + k] = t;
        }
      }
    }
    __syncwarp();

    // ---- 5. back substitution L' B = D^-1 y: thread r owns columns r and h of L -----------------
    {
      double cl_lo[M], cl_hi[T];
      const bool lo_col = r < m, hi_col = h < m;
#pragma unroll
      for (int ii = 1; ii < M; ++ii) cl_lo[ii] = (lo_col && ii > r && ii < m) ? A[ii * LD + r] : 0.0;
#pragma unroll
      for (int ii = T; ii < M; ++ii) cl_hi[ii - T] = (hi_col && ii > h && ii < m) ? A[ii * LD + h] : 0.0;
#pragma unroll
      for (int ii = M - 1; ii >= 1; --ii) {
        if (ii < m) {
          const double src = (ii < T) ? w_lo : w_hi;
          const int srcl = (ii < T) ? ii : (M - 1 - ii);
          const double Bi = __shfl_sync(0xffffffffu, src, srcl, T);
          w_lo = fma(-cl_lo[ii], Bi, w_lo);
          if (ii >= T) w_hi = fma(-cl_hi[ii - T], Bi, w_hi);
        }
      }
    }

    // ---- 6. outputs and the fused likelihood terms -----------------------------------------------
    const double Fv = sigmaSq + tauSq - Fsum;
    if (th.B != nullptr) {
      if (site) {  // q x m column-major weights
        if (lo_real) th.B[(size_t)r * P.q + i] = w_lo;
        if (hi_real) th.B[(size_t)h * P.q + i] = w_hi;
      } else {
        if (lo_real) th.B[off + r] = w_lo;
        if (hi_real) th.B[off + h] = w_hi;
      }
    }
    if (th.F != nullptr && r == 0 && active) th.F[i] = Fv;
    double part = 0.0;
    if (resid != nullptr) {
      if (lo_real) part = w_lo * __ldg(&resid[idx_lo]);
      if (hi_real) part = fma(w_hi, __ldg(&resid[idx_hi]), part);
    }
#pragma unroll
    for (int s = T / 2; s >= 1; s >>= 1) part += __shfl_xor_sync(0xffffffffu, part, s, T);
    if (r == 0 && active && !site) {
      lsum += log(Fv);
      if (resid != nullptr) {
        const double e = __ldg(&resid[i]) - part;
        qsum += e * e / Fv;
      }
    }
  }

  // ---- grid reduction: warp -> CTA -> per-CTA partials -> last CTA sums in fixed order -----------
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    lsum += __shfl_down_sync(0xffffffffu, lsum, s);
    qsum += __shfl_down_sync(0xffffffffu, qsum, s);
  }
  if ((tid & 31) == 0) {
    red[0][tid >> 5] = lsum;
    red[1][tid >> 5] = qsum;
  }
  if (failed != INT_MAX) atomicMin(th.fail, failed);
  __syncthreads();
  if (tid == 0) {
    th.partials[2 * blockIdx.x] = (red[0][0] + red[0][1]) + (red[0][2] + red[0][3]);
    th.partials[2 * blockIdx.x + 1] = (red[1][0] + red[1][1]) + (red[1][2] + red[1][3]);
    __threadfence();
    const unsigned int tk = atomicAdd(th.ticket, 1u);
    last_flag = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (last_flag && tid < 32) {
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned int j = tid; j < gridDim.x; j += 32) {
      a += __ldcg(&th.partials[2 * j]);
      b += __ldcg(&th.partials[2 * j + 1]);
    }
#pragma unroll
    for (int s = 16; s >= 1; s >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, s);
      b += __shfl_down_sync(0xffffffffu, b, s);
    }
    if (tid == 0) {
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

// ---- launcher -------------------------------------------------------------------------------------

template <int M, int COV>
static cudaError_t launch_one(const BfLaunch &P, int grid_x, cudaStream_t stream) {
  using Geo = BfGeo<M>;
  // the attribute is per device and cheap to set: set it on every launch (a process may drive several GPUs)
  cudaError_t e = cudaFuncSetAttribute(bf_kernel<M, COV>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)Geo::kSmemBytes);
  if (e != cudaSuccess) return e;
  dim3 grid(grid_x, P.nbatch, 1);
  bf_kernel<M, COV><<<grid, Geo::kThreads, Geo::kSmemBytes, stream>>>(P);
  return cudaGetLastError();
}

template <int M>
static cudaError_t launch_cov(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream) {
  switch (cov) {
    case NNGP_COV_EXPONENTIAL: return launch_one<M, NNGP_COV_EXPONENTIAL>(P, grid_x, stream);
    case NNGP_COV_SPHERICAL: return launch_one<M, NNGP_COV_SPHERICAL>(P, grid_x, stream);
    case NNGP_COV_MATERN: return launch_one<M, NNGP_COV_MATERN>(P, grid_x, stream);
    case NNGP_COV_GAUSSIAN: return launch_one<M, NNGP_COV_GAUSSIAN>(P, grid_x, stream);
    default: return cudaErrorInvalidValue;
  }
}

// Tile capacity used for m neighbours.
int bf_tile_for_m(int m) { return m <= 8 ? 8 : (m <= 16 ? 16 : 32); }

int bf_locations_per_cta(int m) {
  const int M = bf_tile_for_m(m);
  return 128 / (M / 2);
}

// Maximum useful grid.x (CTAs per parameter set) on a device with `sms` SMs.
int bf_grid_for(int m, long long nloc, int sms) {
  const int M = bf_tile_for_m(m);
  const int per_sm = (M <= 16) ? 4 : 2;
  const long long need = (nloc + bf_locations_per_cta(m) - 1) / bf_locations_per_cta(m);
  long long g = (long long)sms * per_sm;
  if (need < g) g = need;
  if (g < 1) g = 1;
  return (int)g;
}

cudaError_t coop_launch(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream) {
  const int M = bf_tile_for_m(P.m);
  if (M == 8) return launch_cov<8>(P, cov, grid_x, stream);
  if (M == 16) return launch_cov<16>(P, cov, grid_x, stream);
  return launch_cov<32>(P, cov, grid_x, stream);
}

// bf_tpl_kernel.cu / bf_ws_kernel.cu
int tpl_capacity_for_m(int m);
int tpl_grid_for(int m, long long nloc, int sms);
cudaError_t tpl_launch(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream);
int ws_capacity_for_m(int m);
int ws_grid_for(int m, long long nloc, int sms);
cudaError_t ws_launch(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream);

// Kernel choice.  m <= 16: the warp-specialised one-lane-per-location kernel (bf_ws_kernel.cu);
// larger m: the cooperative warp-group kernel above.  NNGP_KERNEL=coop|tpl|ws forces one (A/B profiling;
// tpl = the single-role one-lane-per-location kernel of bf_tpl_kernel.cu).
enum { KERNEL_COOP = 0, KERNEL_TPL = 1, KERNEL_WS = 2 };

static int kernel_for(int m, int cov) {
  static int forced = -2;
  if (forced == -2) {
    const char *e = getenv("NNGP_KERNEL");
    forced = -1;
    if (e) {
      if (!strcmp(e, "coop")) forced = KERNEL_COOP;
      else if (!strcmp(e, "tpl")) forced = KERNEL_TPL;
      else if (!strcmp(e, "ws")) forced = KERNEL_WS;
    }
  }
  (void)cov;
  const bool tpl_ok = tpl_capacity_for_m(m) > 0, ws_ok = ws_capacity_for_m(m) > 0;
  if (forced == KERNEL_COOP) return forced;
  if (forced == KERNEL_TPL && tpl_ok) return forced;
  if (forced == KERNEL_WS && ws_ok) return forced;
  // small tiles do not amortise the producer/consumer hand-off (measured at n = 10^6, exponential: m = 5 0.112 ms
  // single-role vs 0.133 ms specialised; m = 8 0.243 vs 0.223; m = 10 0.377 vs 0.313 ms); beyond m = 24 two stages of
  // the specialised kernel no longer fit an SM's shared memory
  if (m <= 5) return KERNEL_TPL;
  return ws_ok ? KERNEL_WS : KERNEL_COOP;
}

int bf_grid_any(int m, int cov, long long nloc, int sms) {
  switch (kernel_for(m, cov)) {
    case KERNEL_WS: return ws_grid_for(m, nloc, sms);
    case KERNEL_TPL: return tpl_grid_for(m, nloc, sms);
    default: return bf_grid_for(m, nloc, sms);
  }
}

bool bf_has_matern_short(int m) { return kernel_for(m, NNGP_COV_MATERN) == KERNEL_WS; }

// upper bound of bf_grid_any over all kernels (for sizing the partial-sum scratch)
int bf_grid_max(int m, long long nloc, int sms) {
  int a = bf_grid_for(m, nloc, sms);
  if (tpl_capacity_for_m(m) > 0) {
    const int b = tpl_grid_for(m, nloc, sms), c = ws_grid_for(m, nloc, sms);
    if (b > a) a = b;
    if (c > a) a = c;
  }
  return a;
}

int tri_capacity_for_m(int m);
cudaError_t tri_launch(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream);

cudaError_t bf_launch(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream) {
  // resident distance triangle given (nngp_set_data built it): stream it instead of gathering sites (bf_tri_kernel.cu)
  if (P.tri != nullptr && !P.site_mode && tri_capacity_for_m(P.m) > 0) return tri_launch(P, cov, grid_x, stream);
  switch (kernel_for(P.m, cov)) {
    case KERNEL_WS: return ws_launch(P, cov, grid_x, stream);
    case KERNEL_TPL: return tpl_launch(P, cov, grid_x, stream);
    default: return coop_launch(P, cov, grid_x, stream);
  }
}

}  // namespace nngp

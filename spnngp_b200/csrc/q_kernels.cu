// q_kernels.cu -- HBM-bound passes over stored B, F: residual formation, the quadratic form
// Q(B,F,r,r) alone, and all cross quadratic forms of the columns [X y] in one pass.
//
// Replaces (reference file:line):
//   r = X beta - y                      src/rNNGP.cpp:299-303 (dgemv + daxpy)
//   Q(B,F,u,v)                          src/util.cpp:261-280
//   the p + p(p+1)/2 (+1) calls of Q    src/rNNGP.cpp:281-286, src/cNNGP.cpp:227-232,245
//
// Layout: LP = 8/16/32 lanes cooperate on one location so that B and nnIndx (contiguous per
// location, consecutive locations adjacent) are read fully coalesced; the u[N(i,j)] gathers hit L2.
// Reductions are fixed-order trees (shuffle -> CTA -> per-CTA partials -> second tiny kernel), so
// results are run-to-run deterministic.
#include "nngp_common.cuh"

namespace nngp {

// beta travels as a kernel parameter when p <= kResidBetaMax: no host-to-device copy in front of every evaluation
struct ResidBeta {
  double b[kResidBetaMax];
};
__global__ void __launch_bounds__(256) resid_val_kernel(const double *__restrict__ X, const double *__restrict__ y,
                                                        const __grid_constant__ ResidBeta beta, int n, int p,
                                                        double *__restrict__ r, int i0, int i1) {
  // the B/F launch behind this kernel may set itself up meanwhile (it waits for this grid before it reads r)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += stride) {
    double acc = 0.0;
    for (int c = 0; c < p; ++c) acc = fma(__ldg(&X[(size_t)c * n + i]), beta.b[c], acc);
    r[i] = acc - __ldg(&y[i]);
  }
}

__global__ void __launch_bounds__(256) resid_kernel(const double *__restrict__ X, const double *__restrict__ y,
                                                    const double *__restrict__ beta, int n, int p,
                                                    double *__restrict__ r, int i0, int i1) {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = i0 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < i1; i += stride) {
    double acc = 0.0;
    for (int c = 0; c < p; ++c) acc = fma(__ldg(&X[(size_t)c * n + i]), __ldg(&beta[c]), acc);
    r[i] = acc - __ldg(&y[i]);
  }
}

// ---- sums of V values per lane over groups of LP lanes, by a butterfly that HALVES the value count at each step ----
// Step with mask M: a lane keeps the upper half of its values if (lane & M), the lower half otherwise, and adds the
// partner's copy of the same half: V/2 shuffles instead of V.  After log2(LP) steps the V sums are spread over the
// lanes: V - 1 (+ a few) 64-bit shuffles in place of V log2(LP) -- shuffles run through the LSU data pipe, which is
// what bounded these passes (ncu: l1tex data-pipe wavefronts 81%, SHFL on short scoreboard / mio).
// On return lane l holds, in v[0 .. max(1, V/LP)), the sums with original indices
//   reduce_index<V, LP>(l) + j,  j < max(1, V/LP)       (replicated over lanes that agree in their low log2(V) bits).
template <int V, int LP>
__device__ __forceinline__ int reduce_index(int lane) {
  int idx = 0, c = V, s = 0;
#pragma unroll
  for (int M = 1; M < LP; M <<= 1, ++s) {
    if (c > 1) {
      c >>= 1;
      idx |= ((lane >> s) & 1) * c;
    }
  }
  return idx;
}

template <int V, int LP>
__device__ __forceinline__ void lanes_reduce(double (&v)[V], int lane) {
  int c = V;  // compile-time after unrolling
#pragma unroll
  for (int M = 1; M < LP; M <<= 1) {
    if (c > 1) {
      c >>= 1;
      const bool upper = (lane & M) != 0;
#pragma unroll
      for (int j = 0; j < V / 2; ++j) {
        if (j < c) {
          const double send = upper ? v[j] : v[j + c];
          const double keep = upper ? v[j + c] : v[j];
          v[j] = keep + __shfl_xor_sync(0xffffffffu, send, M, LP);
        }
      }
    } else {
      v[0] += __shfl_xor_sync(0xffffffffu, v[0], M, LP);
    }
  }
}

// ---- Q(B,F,r,r) over a shard ----------------------------------------------------------------------
// UN locations per lane group and trip: their id/weight loads, then their gathers, go out together (with one
// location per trip the id -> gather chain leaves too few bytes in flight to fill HBM: 156 us vs 55 us at n = 10^6).
template <int LP>
__global__ void __launch_bounds__(256) q_resid_kernel(const double *__restrict__ B, const double *__restrict__ F,
                                                      const int *__restrict__ nnIndx,
                                                      const double *__restrict__ resid, int m, int loc_begin,
                                                      int loc_end, double *__restrict__ partials) {
  constexpr int G = 256 / LP;
  constexpr int UN = 4;
  __shared__ double red[8];
  const int tid = threadIdx.x, lane = tid % LP, g = tid / LP;
  double qsum = 0.0;
  for (long long base = loc_begin + (long long)blockIdx.x * (G * UN); base < loc_end; base += (long long)gridDim.x * (G * UN)) {
    double b[UN];
    int id[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long i = base + u * G + g;  // consecutive groups take consecutive locations: coalesced B, nnIndx
      const bool act = i < loc_end;
      const int mi = act ? nn_count(i, m) : 0;
      const long long off = act ? nn_offset(i, m) : 0;
      const bool ok = lane < mi;
      b[u] = ok ? __ldg(&B[off + lane]) : 0.0;
      id[u] = ok ? __ldg(&nnIndx[off + lane]) : 0;
    }
    // lanes 0..UN-1 of the group each finish one row: the one whose sum the butterfly leaves with them
    const int umine = reduce_index<UN, LP>(lane);
    const long long imine = base + (long long)umine * G + g;
    const bool mine = lane < UN && imine < loc_end;
    const double ri = mine ? __ldg(&resid[imine]) : 0.0;
    const double fi = mine ? __ldg(&F[imine]) : 1.0;
    double a[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) a[u] = b[u] * __ldg(&resid[id[u]]);  // weight 0 for missing neighbours
    lanes_reduce<UN, LP>(a, lane);
    if (mine) {
      const double e = ri - a[0];
      qsum += e * e / fi;
    }
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) qsum += __shfl_down_sync(0xffffffffu, qsum, s);
  if ((tid & 31) == 0) red[tid >> 5] = qsum;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    partials[blockIdx.x] = t;
  }
}

// sums `count` partial vectors of length `len` (partials[j*len + e]) in fixed order: one CTA per entry
__global__ void __launch_bounds__(256) sum_partials_kernel(const double *__restrict__ partials, int count, int len,
                                                           double *__restrict__ out) {
  __shared__ double red[8];
  const int e = blockIdx.x, tid = threadIdx.x;
  double a = 0.0;
  for (int j = tid; j < count; j += 256) a += partials[(size_t)j * len + e];
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) a += __shfl_down_sync(0xffffffffu, a, s);
  if ((tid & 31) == 0) red[tid >> 5] = a;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[e] = t;
  }
}

// ---- all cross forms S = [X y]' C~^-1 [X y] ---------------------------------------------------------
// CTA = 256 threads handles CH = 128 consecutive locations: phase 1 writes the whitened columns
// e_c(i) = (u_c[i] - B_i . u_c[N(i)]) / sqrt(F_i) of its chunk to shared memory; phase 2 forms the
// K(K+1)/2 cross products of the chunk; per-CTA partials are summed by sum_partials_kernel.
// Xy: the columns [X y] as ROWS (n x KS row-major, KS = K rounded up to 4, zero padded; xy_rows_kernel), so that the
// cross-form pass fetches four columns of a neighbour with ONE 32-byte load.  ncu on the column-wise version: the
// pass is bound by L1 wavefronts (l1tex data pipe 85%: every 8-byte gather of a warp is its own wavefront), not by
// HBM or L2 -- a quarter of the gather instructions is a quarter of the wavefronts.
__host__ __device__ inline int xy_row_stride(int p) { return ((p + 1) + 3) & ~3; }

__global__ void __launch_bounds__(256) xy_rows_kernel(const double *__restrict__ X, const double *__restrict__ y, int n,
                                                      int p, double *__restrict__ Xy) {
  const int KS = xy_row_stride(p);
  const long long total = (long long)n * KS, stride = (long long)gridDim.x * blockDim.x;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const long long i = t / KS;
    const int c = (int)(t - i * KS);
    Xy[t] = (c < p) ? __ldg(&X[(size_t)c * n + i]) : (c == p ? __ldg(&y[i]) : 0.0);
  }
}

// 32-byte read-only load (sm_100: 256-bit vector loads)
__device__ __forceinline__ void ldg256(const double *p, double (&v)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

size_t xy_rows_count(int n, int p) { return (size_t)n * xy_row_stride(p); }

cudaError_t xy_rows_launch(const double *X, const double *y, int n, int p, double *Xy, int sms, cudaStream_t stream) {
  long long g = ((long long)n * xy_row_stride(p) + 255) / 256;
  if (g > (long long)sms * 16) g = (long long)sms * 16;
  xy_rows_kernel<<<(int)(g < 1 ? 1 : g), 256, 0, stream>>>(X, y, n, p, Xy);
  return cudaGetLastError();
}

template <int LP>
__global__ void __launch_bounds__(256) q_forms_kernel(const double *__restrict__ B, const double *__restrict__ F,
                                                      const int *__restrict__ nnIndx, const double *__restrict__ Xy,
                                                      int n, int p, int m, int loc_begin, int loc_end,
                                                      double *__restrict__ partials) {
  constexpr int CH = 128;
  constexpr int G = 256 / LP;
  extern __shared__ double W[];  // CH x K, column c at W[c*CH ...]
  const int K = p + 1;
  const int npairs = K * (K + 1) / 2;
  const int tid = threadIdx.x, lane = tid % LP, g = tid / LP;
  const long long chunk0 = loc_begin + (long long)blockIdx.x * CH;

  constexpr int UN = 4;  // locations per lane group and trip: ids/weights first, then all gathers (bytes in flight)
  static_assert(CH % (G * UN) == 0, "chunk must be a whole number of trips");
  for (int s0 = 0; s0 < CH; s0 += G * UN) {
    double b[UN];
    int idx[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const long long i = chunk0 + s0 + u * G + g;
      const bool act = i < loc_end;
      const int mi = act ? nn_count(i, m) : 0;
      const long long off = act ? nn_offset(i, m) : 0;
      const bool ok = lane < mi;
      b[u] = ok ? __ldg(&B[off + lane]) : 0.0;
      idx[u] = ok ? __ldg(&nnIndx[off + lane]) : 0;
    }
    // KC columns at a time: their UN x KC products per lane are summed over the group by one halving butterfly,
    // which leaves sum (row u, column cc) with lane reduce_index(lane) = u KC + cc; that lane finishes the entry.
    constexpr int KC = 4, V = UN * KC, HELD = (V / LP > 1) ? V / LP : 1;
    const int KS = xy_row_stride(p);
    const int vidx = reduce_index<V, LP>(lane);
    const int umine = vidx / KC, ccmine = vidx % KC;  // HELD consecutive columns of row umine, starting at ccmine
    const int smine = s0 + umine * G + g;
    const long long imine = chunk0 + smine;
    const bool fin = lane < V;  // lanes V.. of a 32-lane group hold replicas
    const bool mine = fin && imine < loc_end;
    const double rsq = mine ? rsqrt(__ldg(&F[imine])) : 0.0;
    for (int c0 = 0; c0 < K; c0 += KC) {
      double prod[V], row[KC], own[HELD];
#pragma unroll
      for (int h = 0; h < HELD; ++h) own[h] = (mine && c0 + ccmine + h < K) ? __ldg(&Xy[(size_t)imine * KS + c0 + ccmine + h]) : 0.0;
#pragma unroll
      for (int u = 0; u < UN; ++u) {
        ldg256(&Xy[(size_t)idx[u] * KS + c0], row);  // row 0 with weight 0: missing neighbour
#pragma unroll
        for (int cc = 0; cc < KC; ++cc) prod[u * KC + cc] = b[u] * row[cc];
      }
      lanes_reduce<V, LP>(prod, lane);
#pragma unroll
      for (int h = 0; h < HELD; ++h)
        if (fin && c0 + ccmine + h < K) W[(c0 + ccmine + h) * CH + smine] = (own[h] - prod[h]) * rsq;  // 0 beyond the shard
    }
  }
  __syncthreads();
  // phase 2: pair e -> (c, d), c >= d, column-major lower packing; 8 lanes per pair
  const int sub = tid & 7;
  for (int e0 = 0; e0 < npairs; e0 += 32) {  // uniform trip count: the shuffles below need full warps
    const int e = e0 + (tid >> 3);
    const bool ok = e < npairs;
    int d = 0, rem = ok ? e : 0;
    while (rem >= K - d) {
      rem -= K - d;
      ++d;
    }
    const int c = d + rem;
    double a = 0.0;
    if (ok)
      for (int s = sub; s < CH; s += 8) a = fma(W[c * CH + s], W[d * CH + s], a);
    a += __shfl_xor_sync(0xffffffffu, a, 4, 8);
    a += __shfl_xor_sync(0xffffffffu, a, 2, 8);
    a += __shfl_xor_sync(0xffffffffu, a, 1, 8);
    if (ok && sub == 0) partials[(size_t)blockIdx.x * npairs + e] = a;
  }
}

// ---- launchers --------------------------------------------------------------------------------------
static int lanes_for_m(int m) { return m <= 8 ? 8 : (m <= 16 ? 16 : 32); }

// r[i] = x_i' beta - y_i for i in [i0, i1): a shard only needs its own locations and their neighbours (all earlier in the
// ordering), so a sharded fit does not recompute the whole vector on every rank and every iteration
cudaError_t resid_launch(const double *X, const double *y, const double *beta, int n, int p, double *r, int i0, int i1, int sms,
                         cudaStream_t stream) {
  int grid = (i1 - i0 + 255) / 256;
  if (grid > sms * 8) grid = sms * 8;
  if (grid < 1) grid = 1;
  resid_kernel<<<grid, 256, 0, stream>>>(X, y, beta, n, p, r, i0, i1);
  return cudaGetLastError();
}

// beta_host: p <= kResidBetaMax doubles on the host
cudaError_t resid_launch_val(const double *X, const double *y, const double *beta_host, int n, int p, double *r, int i0, int i1,
                             int sms, cudaStream_t stream) {
  int grid = (i1 - i0 + 255) / 256;
  if (grid > sms * 8) grid = sms * 8;
  if (grid < 1) grid = 1;
  ResidBeta b;
  for (int c = 0; c < kResidBetaMax; ++c) b.b[c] = c < p ? beta_host[c] : 0.0;
  resid_val_kernel<<<grid, 256, 0, stream>>>(X, y, b, n, p, r, i0, i1);
  return cudaGetLastError();
}

int q_resid_grid(int m, long long nloc, int sms) {
  const int G = 4 * (256 / lanes_for_m(m));  // locations per CTA and trip (UN = 4)
  long long need = (nloc + G - 1) / G;
  long long g = (long long)sms * 8;
  if (need < g) g = need;
  return (int)(g < 1 ? 1 : g);
}

// partials must hold q_resid_grid() doubles; out[0] receives Q.
cudaError_t q_resid_launch(const double *B, const double *F, const int *nnIndx, const double *resid, int m,
                           int loc_begin, int loc_end, double *partials, double *out, int sms,
                           cudaStream_t stream) {
  const int grid = q_resid_grid(m, (long long)loc_end - loc_begin, sms);
  const int lp = lanes_for_m(m);
  if (lp == 8) q_resid_kernel<8><<<grid, 256, 0, stream>>>(B, F, nnIndx, resid, m, loc_begin, loc_end, partials);
  else if (lp == 16) q_resid_kernel<16><<<grid, 256, 0, stream>>>(B, F, nnIndx, resid, m, loc_begin, loc_end, partials);
  else q_resid_kernel<32><<<grid, 256, 0, stream>>>(B, F, nnIndx, resid, m, loc_begin, loc_end, partials);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sum_partials_kernel<<<1, 256, 0, stream>>>(partials, grid, 1, out);
  return cudaGetLastError();
}

int q_forms_grid(long long nloc) { return (int)((nloc + 127) / 128 < 1 ? 1 : (nloc + 127) / 128); }

// partials must hold q_forms_grid() * (p+1)(p+2)/2 doubles; out receives the packed lower triangle
// (column-major: (0,0),(1,0),..,(K-1,0),(1,1),...).
cudaError_t q_forms_launch(const double *B, const double *F, const int *nnIndx, const double *Xy, int n, int p, int m,
                           int loc_begin, int loc_end, double *partials, double *out, cudaStream_t stream) {
  const int K = p + 1, npairs = K * (K + 1) / 2;
  const int grid = q_forms_grid((long long)loc_end - loc_begin);
  const size_t smem = (size_t)128 * K * sizeof(double);
  const int lp = lanes_for_m(m);
  cudaError_t e;
#define NNGP_QF(LPV)                                                                                         \
  do {                                                                                                       \
    if (smem > 48 * 1024) {                                                                                  \
      e = cudaFuncSetAttribute(q_forms_kernel<LPV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
      if (e != cudaSuccess) return e;                                                                        \
    }                                                                                                        \
    q_forms_kernel<LPV><<<grid, 256, smem, stream>>>(B, F, nnIndx, Xy, n, p, m, loc_begin, loc_end, partials); \
  } while (0)
  if (lp == 8) NNGP_QF(8);
  else if (lp == 16) NNGP_QF(16);
  else NNGP_QF(32);
#undef NNGP_QF
  e = cudaGetLastError();
  if (e != cudaSuccess) return e;
  sum_partials_kernel<<<npairs, 256, 0, stream>>>(partials, grid, npairs, out);
  return cudaGetLastError();
}

}  // namespace nngp

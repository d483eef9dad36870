// rep_kernels.cu -- replicated-data draws: v_i = B_i . v_N(i) + sqrt(F_i) z_i, y.rep = X beta + v.
//
// Reference: updateRep src/rNNGP.cpp:24-42 (called at :309-313 when fit.rep is on) and the sample loop of
// rNNGPReplicated src/rNNGPReplicated.cpp:183-201.  The recurrence is a sparse unit-lower-triangular solve
// (I - B) v = sqrt(F) z: location i needs v at its neighbours, all earlier in the ordering.  The reference
// walks i = 0..n-1 serially.  Here the neighbour graph is levelled once per data set (level(i) = 1 + max
// level over N(i); host pass over the ids, 50 ms at n = 10^6): locations of one level are independent.
// An x-sorted NNGP ordering gives ~8.5 sqrt(n) levels of ~sqrt(n)/8 locations each (n = 10^6: 8,454 levels,
// 120 wide) -- far too narrow to feed 148 SMs, so ONE CTA walks the levels with a block barrier between them,
// 4 lanes per location, everything that does not depend on v (ids, weights, F, z of the NEXT level)
// prefetched into registers before the barrier, so a level costs one L2 round trip for the v gathers.
#include <vector>

#include "nngp_ctx.h"

namespace nngp {

constexpr int kRepThreads = 1024;
constexpr int kRepLanes = 4;  // lanes per location
constexpr int kRepGroups = kRepThreads / kRepLanes;

template <int KMAX>  // neighbour slots per lane: ceil(m / 4)
__global__ void __launch_bounds__(kRepThreads, 1)
rep_solve_kernel(const int *__restrict__ perm, const int *__restrict__ levoff, int nlev, const int *__restrict__ nnIndx,
                 const double *__restrict__ B, const double *__restrict__ F, const double *__restrict__ z, double *v, int m) {
  const int tid = threadIdx.x, sub = tid & (kRepLanes - 1), g = tid / kRepLanes;
  const unsigned gmask = 0xFu << ((tid & 31) & ~3);
  // prefetched state of this group's first location in the level about to be processed
  int pi = -1, pid[KMAX];
  double pb[KMAX], pfz = 0.0;
  auto prefetch = [&](int L) {
    pi = -1;
    if (L >= nlev) return;
    const int q = __ldg(&levoff[L]) + g;
    if (q >= __ldg(&levoff[L + 1])) return;
    pi = __ldg(&perm[q]);
    const int mi = nn_count(pi, m);
    const long long off = nn_offset(pi, m);
#pragma unroll
    for (int t = 0; t < KMAX; ++t) {
      const int k = sub + t * kRepLanes;
      pid[t] = -1;
      pb[t] = 0.0;
      if (k < mi) {
        pid[t] = __ldg(&nnIndx[off + k]);
        pb[t] = __ldg(&B[off + k]);
      }
    }
    if (sub == 0) pfz = sqrt(__ldg(&F[pi])) * __ldg(&z[pi]);
  };
  prefetch(0);
  for (int L = 0; L < nlev; ++L) {
    const int beg = __ldg(&levoff[L]), end = __ldg(&levoff[L + 1]);
    // first pass: the prefetched location
    const int i0 = pi;
    double acc = 0.0;
    if (i0 >= 0) {
#pragma unroll
      for (int t = 0; t < KMAX; ++t)
        if (pid[t] >= 0) acc = fma(pb[t], __ldcg(&v[pid[t]]), acc);
    }
    const double fz0 = pfz;
    prefetch(L + 1);  // static data of the next level: in flight across the reduction and the barrier
    if (i0 >= 0) {
      acc += __shfl_xor_sync(gmask, acc, 1);
      acc += __shfl_xor_sync(gmask, acc, 2);
      if (sub == 0) v[i0] = acc + fz0;
    }
    // levels wider than the CTA's 256 groups (not seen on x-sorted orderings): plain passes
    for (int q = beg + g + kRepGroups; q < end; q += kRepGroups) {
      const int i = __ldg(&perm[q]);
      const int mi = nn_count(i, m);
      const long long off = nn_offset(i, m);
      double a = 0.0;
      for (int k = sub; k < mi; k += kRepLanes) a = fma(__ldg(&B[off + k]), __ldcg(&v[__ldg(&nnIndx[off + k])]), a);
      a += __shfl_xor_sync(gmask, a, 1);
      a += __shfl_xor_sync(gmask, a, 2);
      if (sub == 0) v[i] = a + sqrt(__ldg(&F[i])) * __ldg(&z[i]);
    }
    __syncthreads();
  }
}

// yrep_i = v_i + x_i' beta (src/rNNGP.cpp:300-302,312; src/rNNGPReplicated.cpp:199-200); in place over v
struct RepBeta {
  double b[64];
};
__global__ void __launch_bounds__(256) rep_add_mean_kernel(const double *__restrict__ X, RepBeta beta, int n, int p, double *v) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    double s = 0.0;
    for (int j = 0; j < p; ++j) s = fma(X[(size_t)j * n + i], beta.b[j], s);
    v[i] += s;
  }
}

// Level structure of the neighbour graph (host pass over the ids taken back from the device).
int rep_prepare(nngp_ctx *c) {
  if (c->rep_ready) return NNGP_OK;
  const int n = c->n, m = c->m;
  std::vector<int> idx((size_t)(c->nIndx > 0 ? c->nIndx : 1)), lev((size_t)n), perm((size_t)n);
  if (c->nIndx > 0)
    NNGP_CUDA(c, cudaMemcpy(idx.data(), c->d_nnIndx, (size_t)c->nIndx * sizeof(int), cudaMemcpyDeviceToHost));
  int depth = 0;
  for (int i = 0; i < n; i++) {
    const int mi = nn_count(i, m);
    const long long off = nn_offset(i, m);
    int l = 0;
    for (int k = 0; k < mi; k++) {
      const int j = idx[off + k];
      if (j >= i) return ctx_fail(c, NNGP_EINVAL, "replicates: neighbour ids must precede their location in the ordering");
      l = lev[j] + 1 > l ? lev[j] + 1 : l;
    }
    lev[i] = l;
    depth = l + 1 > depth ? l + 1 : depth;
  }
  std::vector<int> off((size_t)depth + 1, 0);
  for (int i = 0; i < n; i++) off[lev[i] + 1]++;
  for (int l = 0; l < depth; l++) off[l + 1] += off[l];
  {
    std::vector<int> cur(off.begin(), off.end() - 1);
    for (int i = 0; i < n; i++) perm[cur[lev[i]]++] = i;  // stable: ascending index inside a level
  }
  if (c->d_rep_perm) cudaFree(c->d_rep_perm);
  if (c->d_rep_levoff) cudaFree(c->d_rep_levoff);
  c->d_rep_perm = c->d_rep_levoff = nullptr;
  NNGP_CUDA(c, cudaMalloc(&c->d_rep_perm, (size_t)n * sizeof(int)));
  NNGP_CUDA(c, cudaMalloc(&c->d_rep_levoff, ((size_t)depth + 1) * sizeof(int)));
  NNGP_CUDA(c, cudaMemcpy(c->d_rep_perm, perm.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
  NNGP_CUDA(c, cudaMemcpy(c->d_rep_levoff, off.data(), ((size_t)depth + 1) * sizeof(int), cudaMemcpyHostToDevice));
  c->rep_levels = depth;
  if (c->h_z_cap < (size_t)n) {
    if (c->h_z) cudaFreeHost(c->h_z);
    c->h_z = nullptr;
    NNGP_CUDA(c, cudaHostAlloc(&c->h_z, (size_t)n * sizeof(double), cudaHostAllocDefault));
    c->h_z_cap = (size_t)n;
  }
  c->rep_ready = true;
  return NNGP_OK;
}

// z (host, n normals in location order; may be c->h_z itself) -> yrep (host, n).  B, F of `slot` must be current.
int rep_draw(nngp_ctx *c, int slot, const double *beta, const double *z, double *yrep) {
  if (c->n == 0 || !c->has_nn) return ctx_fail(c, NNGP_ESTATE, "replicates requested before nngp_set_data (with neighbours)");
  if (c->loc_begin != 0 || c->loc_end != c->n)
    return ctx_fail(c, NNGP_ESTATE, "replicates need the whole location range on one GPU (the recurrence is serial in the ordering)");
  if (c->p > 64) return ctx_fail(c, NNGP_EINVAL, "replicates: p > 64 not supported");
  int rc = rep_prepare(c);
  if (rc) return rc;
  const int n = c->n;
  double *d_z = c->d_resid[2], *d_v = c->d_resid[3];  // batch entries 2, 3 are idle outside 4-set launches
  if (c->p == 0 || !d_z || !d_v) return ctx_fail(c, NNGP_ESTATE, "replicates need a model with covariates");
  if (z != c->h_z) memcpy(c->h_z, z, (size_t)n * sizeof(double));
  NNGP_CUDA(c, cudaMemcpyAsync(d_z, c->h_z, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  if (c->m <= 16)
    rep_solve_kernel<4><<<1, kRepThreads, 0, c->stream>>>(c->d_rep_perm, c->d_rep_levoff, c->rep_levels, c->d_nnIndx, c->d_B[slot],
                                                          c->d_F[slot], d_z, d_v, c->m);
  else
    rep_solve_kernel<8><<<1, kRepThreads, 0, c->stream>>>(c->d_rep_perm, c->d_rep_levoff, c->rep_levels, c->d_nnIndx, c->d_B[slot],
                                                          c->d_F[slot], d_z, d_v, c->m);
  NNGP_CUDA(c, cudaGetLastError());
  RepBeta rb;
  for (int j = 0; j < 64; j++) rb.b[j] = (beta && j < c->p) ? beta[j] : 0.0;
  int grid = (n + 255) / 256;
  if (grid > c->sms * 8) grid = c->sms * 8;
  rep_add_mean_kernel<<<grid, 256, 0, c->stream>>>(c->d_X, rb, n, c->p, d_v);
  NNGP_CUDA(c, cudaGetLastError());
  c->launches += 2;
  NNGP_CUDA(c, cudaMemcpyAsync(c->h_z, d_v, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(yrep, c->h_z, (size_t)n * sizeof(double));
  return NNGP_OK;
}

}  // namespace nngp

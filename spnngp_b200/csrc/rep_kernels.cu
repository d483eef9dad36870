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

template <int KMAX>
struct RepSlot {  // static data of one location's share held by one lane
  int id[KMAX];
  double w[KMAX];
  double f, z;
};

template <int KMAX>  // neighbour slots per lane: ceil(m / 4)
__global__ void __launch_bounds__(kRepThreads, 1)
rep_solve_kernel(const int *__restrict__ perm, const int *__restrict__ levoff, int nlev, const int *__restrict__ nnIndx,
                 const double *__restrict__ B, const double *__restrict__ F, const double *__restrict__ z, double *v, int m) {
  const int tid = threadIdx.x, sub = tid & (kRepLanes - 1), g = tid / kRepLanes;
  const unsigned gmask = 0xFu << ((tid & 31) & ~3);
  // Software pipeline over levels (everything here is independent of v):
  //   location index of this group's slot in level L+2  <- perm        (one round trip ahead of ...)
  //   ids, weights, F, z of its slot in level L+1       <- nnIndx, B, F, z   (loaded, not used, during level L)
  // so that in level L only the v gathers (and the barrier) are exposed.  Missing neighbours gather v[0] with
  // weight 0, so the gathers are unconditional and go out together.
  auto slot_index = [&](int L) -> int {
    if (L >= nlev) return -1;
    const int q = __ldg(&levoff[L]) + g;
    return (q < __ldg(&levoff[L + 1])) ? __ldg(&perm[q]) : -1;
  };
  auto load_static = [&](int i, RepSlot<KMAX> &S) {
    const int mi = (i >= 0) ? nn_count(i, m) : 0;
    const long long off = (i >= 0) ? nn_offset(i, m) : 0;
#pragma unroll
    for (int t = 0; t < KMAX; ++t) {
      const int k = sub + t * kRepLanes;
      const bool ok = k < mi;
      S.id[t] = ok ? __ldg(&nnIndx[off + k]) : 0;
      S.w[t] = ok ? __ldg(&B[off + k]) : 0.0;
    }
    const int ii = i >= 0 ? i : 0;
    S.f = __ldg(&F[ii]);
    S.z = __ldg(&z[ii]);
  };
  // one level: `cur` holds its static data, `nxt` receives the next level's
  auto level = [&](int L, int i_c, int i_n, int &i_nn, const RepSlot<KMAX> &cur, RepSlot<KMAX> &nxt) {
    double vv[KMAX];
#pragma unroll
    for (int t = 0; t < KMAX; ++t) vv[t] = __ldcg(&v[cur.id[t]]);  // the only loads that depend on earlier levels
    load_static(i_n, nxt);
    i_nn = slot_index(L + 2);
    double acc = 0.0;
#pragma unroll
    for (int t = 0; t < KMAX; ++t) acc = fma(cur.w[t], vv[t], acc);
    acc += __shfl_xor_sync(gmask, acc, 1);
    acc += __shfl_xor_sync(gmask, acc, 2);
    if (i_c >= 0 && sub == 0) v[i_c] = acc + sqrt(cur.f) * cur.z;
    // levels wider than the CTA's 256 groups (not seen on x-sorted orderings): plain passes
    const int beg = __ldg(&levoff[L]), end = __ldg(&levoff[L + 1]);
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
  };
  RepSlot<KMAX> SA, SB;
  int i0 = slot_index(0), i1 = slot_index(1), i2 = -1;
  load_static(i0, SA);
  for (int L = 0; L < nlev; L += 2) {  // two levels per trip: the slot sets swap roles without register moves
    level(L, i0, i1, i2, SA, SB);
    i0 = i1; i1 = i2;
    if (L + 1 >= nlev) break;
    level(L + 1, i0, i1, i2, SB, SA);
    i0 = i1; i1 = i2;
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
  if (c->loc_begin != 0 || c->loc_end != c->n)
    return ctx_fail(c, NNGP_ESTATE, "replicates need the whole location range on one GPU (the recurrence is serial in the ordering)");
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
  if (c->d_rep_z) cudaFree(c->d_rep_z);
  if (c->d_rep_v) cudaFree(c->d_rep_v);
  c->d_rep_z = c->d_rep_v = nullptr;
  NNGP_CUDA(c, cudaMalloc(&c->d_rep_z, (size_t)n * sizeof(double)));
  NNGP_CUDA(c, cudaMalloc(&c->d_rep_v, (size_t)n * sizeof(double)));
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
  double *d_z = c->d_rep_z, *d_v = c->d_rep_v;
  if (c->p == 0) return ctx_fail(c, NNGP_ESTATE, "replicates need a model with covariates");
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

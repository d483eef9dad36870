// knn_kernels.cu -- ordered nearest-neighbour search on sm_100a.
//
// Replaces (reference file:line):
//   mkNNIndx    src/nn.cpp:63-111    brute force over predecessors, strict '<' + rsort_with_index
//   fastNN /    src/nn.cpp:136-257   Ra & Kim mean-distance-ordered walk along the (x+y)-sorted order
//   mkNNIndxCB
//   RANN::nn2   R/generics.R:642     unordered query kNN (third party; our own exact search)
//
// Bit-exactness strategy: every query runs the reference's OWN acceptance sequence (scan order,
// strict comparisons, "overwrite the last slot then Shell-sort" update with R's rsort_with_index
// gap rule) inside one thread, with the reference's non-fused arithmetic (__dmul_rn/__dadd_rn/
// __dsqrt_rn: products rounded separately, IEEE sqrt).  Neighbour ids, their order under ties and
// the nnDist bits therefore match by construction; the GPU only changes WHO runs each query.
// The hot loops keep just the current worst distance in a register; the m-entry list lives in
// local memory and is touched only on the (rare, ~m ln(i/m)) accepted candidates.
#include <math.h>

#include "nngp_common.cuh"

namespace nngp {

// R's rsort_with_index (src/main/sort.c): Shell sort, gaps 1,4,13,.. started at h <= n/9,
// ascending, strict '>' with NaN last.  Restated for the device (third-party algorithm).
__device__ __forceinline__ bool r_greater(double x, double y) {
  const bool nax = isnan(x), nay = isnan(y);
  if (nax && nay) return false;
  if (nax) return true;
  if (nay) return false;
  return x > y;
}

__device__ void rsort_with_index_dev(double *x, int *indx, int n) {
  int h;
  for (h = 1; h <= n / 9; h = 3 * h + 1) {
  }
  for (; h > 0; h /= 3)
    for (int i = h; i < n; i++) {
      const double v = x[i];
      const int iv = indx[i];
      int j = i;
      while (j >= h && r_greater(x[j - h], v)) {
        x[j] = x[j - h];
        indx[j] = indx[j - h];
        j -= h;
      }
      x[j] = v;
      indx[j] = iv;
    }
}

// ---- brute force: one thread per query location, candidate tiles staged in shared memory -----------
constexpr int kBruteThreads = 256;

__global__ void __launch_bounds__(kBruteThreads) knn_brute_kernel(const double2 *__restrict__ coords, int n, int m,
                                                                  int *__restrict__ nnIndx,
                                                                  double *__restrict__ nnDist) {
  __shared__ double2 tile[kBruteThreads];
  const int nchunks = (n + kBruteThreads - 1) / kBruteThreads;
  const int chunk = nchunks - 1 - blockIdx.x;  // longest scans first
  const int i0 = chunk * kBruteThreads;
  const int i = i0 + threadIdx.x;
  const bool valid = i < n;
  const int mi = valid ? nn_count(i, m) : 0;
  const double2 me = valid ? coords[i] : make_double2(0.0, 0.0);
  double ld[NNGP_MAX_M];
  int li[NNGP_MAX_M];
  for (int k = 0; k < NNGP_MAX_M; ++k) {
    ld[k] = INFINITY;
    li[k] = 0;
  }
  double worst = INFINITY;  // ld[mi-1]
  double thr2 = INFINITY;   // upper bound of worst^2: d2 > thr2 implies sqrt(d2) >= worst
  const int jend = min(i0 + kBruteThreads - 1, n - 1);  // candidates needed by this CTA: j < i <= jend
  for (int j0 = 0; j0 < jend; j0 += kBruteThreads) {
    __syncthreads();
    if (j0 + threadIdx.x < n) tile[threadIdx.x] = coords[j0 + threadIdx.x];
    __syncthreads();
    const int lim = min(kBruteThreads, i - j0);  // this thread scans j0 .. j0+lim-1 (j < i)
    if (mi > 0) {
#pragma unroll 4
      for (int jj = 0; jj < lim; ++jj) {
        const double2 pj = tile[jj];
        const double dx = me.x - pj.x, dy = me.y - pj.y;
        const double d2 = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (!(d2 > thr2)) {
          const double d = __dsqrt_rn(d2);
          if (d < worst) {  // src/nn.cpp:101-105
            ld[mi - 1] = d;
            li[mi - 1] = j0 + jj;
            rsort_with_index_dev(ld, li, mi);
            worst = ld[mi - 1];
            thr2 = __dmul_ru(worst, worst);
          }
        }
      }
    }
  }
  if (mi > 0) {
    const long long off = nn_offset(i, m);
    for (int k = 0; k < mi; ++k) {
      nnIndx[off + k] = li[k];
      nnDist[off + k] = ld[k];
    }
  }
}

// ---- code-book search: one thread per position of the (x+y)-sorted order ---------------------------
// sxy / sidx: sites and their NNGP index in ascending u = x + y order.
__global__ void __launch_bounds__(128) knn_cb_kernel(const double2 *__restrict__ sxy, const int *__restrict__ sidx,
                                                     int n, int m, int *__restrict__ nnIndx,
                                                     double *__restrict__ nnDist) {
  const int ui = blockIdx.x * blockDim.x + threadIdx.x;
  if (ui >= n) return;
  const int myidx = sidx[ui];
  const int mi = nn_count(myidx, m);
  if (mi == 0) return;  // the reference calls fastNN(m=0) here (reads rSNNDist[-1]; writes nothing)
  const double2 me = sxy[ui];
  double ld[NNGP_MAX_M];
  int li[NNGP_MAX_M];
  for (int k = 0; k < NNGP_MAX_M; ++k) {
    ld[k] = INFINITY;
    li[k] = 0;
  }
  double worst = INFINITY;  // squared Euclidean distance of the current m-th neighbour
  int i = ui, j = ui;
  bool up = true, down = true;
  while (up || down) {  // src/nn.cpp:151-204
    if (i == 0) down = false;
    if (j == n - 1) up = false;
    if (down) {
      i--;
      const double2 pc = sxy[i];
      // dmi: pow(x0 + x1 - c0 - c1, 2), evaluated left to right (src/nn.cpp:128-130)
      const double t = __dsub_rn(__dsub_rn(__dadd_rn(me.x, me.y), pc.x), pc.y);
      const double dm = __dmul_rn(t, t);
      if (dm > 2.0 * worst) {
        down = false;
      } else {
        const double dx = me.x - pc.x, dy = me.y - pc.y;
        const double de = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (de < worst) {
          const int cid = sidx[i];
          if (cid < myidx) {
            ld[mi - 1] = de;
            li[mi - 1] = cid;
            rsort_with_index_dev(ld, li, mi);
            worst = ld[mi - 1];
          }
        }
      }
    }
    if (up) {
      j++;
      const double2 pc = sxy[j];
      const double t = __dsub_rn(__dsub_rn(__dadd_rn(me.x, me.y), pc.x), pc.y);
      const double dm = __dmul_rn(t, t);
      if (dm > 2.0 * worst) {
        up = false;
      } else {
        const double dx = me.x - pc.x, dy = me.y - pc.y;
        const double de = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
        if (de < worst) {
          const int cid = sidx[j];
          if (cid < myidx) {
            ld[mi - 1] = de;
            li[mi - 1] = cid;
            rsort_with_index_dev(ld, li, mi);
            worst = ld[mi - 1];
          }
        }
      }
    }
  }
  const long long off = nn_offset(myidx, m);
  for (int k = 0; k < mi; ++k) {
    nnIndx[off + k] = li[k];
    nnDist[off + k] = __dsqrt_rn(ld[k]);  // src/nn.cpp:206-208
  }
}

// u = x + y keys for the device sort, and the gather of sites into sorted order
__global__ void knn_keys_kernel(const double2 *__restrict__ coords, int n, double *__restrict__ u, int *__restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const double2 p = coords[i];
    u[i] = __dadd_rn(p.x, p.y);
    idx[i] = i;
  }
}

__global__ void knn_gather_kernel(const double2 *__restrict__ coords, const int *__restrict__ sidx, int n,
                                  double2 *__restrict__ sxy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) sxy[i] = coords[sidx[i]];
}

// counts adjacent equal keys (ties in u make the reference's Shell-sort order matter)
__global__ void knn_count_ties_kernel(const double *__restrict__ u_sorted, int n, int *__restrict__ count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i + 1 < n && u_sorted[i] == u_sorted[i + 1]) atomicAdd(count, 1);
}

// ---- unordered query kNN (nn2 stand-in): strip walk over the (x+y)-sorted sites ---------------------
// ascending (d^2, site index); exact (the same Cauchy-Schwarz pruning as the code-book search).
__global__ void __launch_bounds__(128) knn_query_kernel(const double2 *__restrict__ sxy, const int *__restrict__ sidx,
                                                        const double *__restrict__ u_sorted, int n,
                                                        const double2 *__restrict__ q, int nq, int m,
                                                        int *__restrict__ out /* nq x m column-major */) {
  const int qi = blockIdx.x * blockDim.x + threadIdx.x;
  if (qi >= nq) return;
  const double2 me = q[qi];
  const double uq = me.x + me.y;
  int lo = 0, hi = n;  // first sorted position with u >= uq
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (u_sorted[mid] < uq) lo = mid + 1;
    else hi = mid;
  }
  double ld[NNGP_MAX_M];
  int li[NNGP_MAX_M];
  const int mm = min(m, n);
  for (int k = 0; k < NNGP_MAX_M; ++k) {
    ld[k] = INFINITY;
    li[k] = 0x7fffffff;
  }
  double worst = INFINITY;
  int worst_id = 0x7fffffff;
  int i = lo, j = lo - 1;  // walk down from lo-1 and up from lo
  bool up = true, down = true;
  while (up || down) {
    if (i == 0) down = false;
    if (j == n - 1) up = false;
#pragma unroll
    for (int dir = 0; dir < 2; ++dir) {
      if (dir == 0 ? !down : !up) continue;
      const int pos = (dir == 0) ? --i : ++j;
      const double2 pc = sxy[pos];
      const double t = uq - (pc.x + pc.y);
      if (t * t > 2.0 * worst * (1.0 + 1e-12)) {  // conservative: keeps exactness of the pruning under rounding
        if (dir == 0) down = false;
        else up = false;
        continue;
      }
      const double dx = me.x - pc.x, dy = me.y - pc.y;
      const double de = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
      const int cid = sidx[pos];
      if (de < worst || (de == worst && cid < worst_id)) {
        int k = mm - 1;  // insertion keeping (d2, id) ascending
        while (k > 0 && (ld[k - 1] > de || (ld[k - 1] == de && li[k - 1] > cid))) {
          ld[k] = ld[k - 1];
          li[k] = li[k - 1];
          --k;
        }
        ld[k] = de;
        li[k] = cid;
        worst = ld[mm - 1];
        worst_id = li[mm - 1];
      }
    }
  }
  for (int k = 0; k < m; ++k) out[(size_t)k * nq + qi] = (k < mm) ? li[k] : -1;
}

// ---- launchers ------------------------------------------------------------------------------------
cudaError_t knn_brute_launch(const double2 *coords, int n, int m, int *nnIndx, double *nnDist, cudaStream_t stream) {
  const int nchunks = (n + kBruteThreads - 1) / kBruteThreads;
  knn_brute_kernel<<<nchunks, kBruteThreads, 0, stream>>>(coords, n, m, nnIndx, nnDist);
  return cudaGetLastError();
}

cudaError_t knn_cb_launch(const double2 *sxy, const int *sidx, int n, int m, int *nnIndx, double *nnDist,
                          cudaStream_t stream) {
  knn_cb_kernel<<<(n + 127) / 128, 128, 0, stream>>>(sxy, sidx, n, m, nnIndx, nnDist);
  return cudaGetLastError();
}

cudaError_t knn_keys_launch(const double2 *coords, int n, double *u, int *idx, cudaStream_t stream) {
  knn_keys_kernel<<<(n + 255) / 256, 256, 0, stream>>>(coords, n, u, idx);
  return cudaGetLastError();
}

cudaError_t knn_gather_launch(const double2 *coords, const int *sidx, int n, double2 *sxy, cudaStream_t stream) {
  knn_gather_kernel<<<(n + 255) / 256, 256, 0, stream>>>(coords, sidx, n, sxy);
  return cudaGetLastError();
}

cudaError_t knn_count_ties_launch(const double *u_sorted, int n, int *count, cudaStream_t stream) {
  knn_count_ties_kernel<<<(n + 255) / 256, 256, 0, stream>>>(u_sorted, n, count);
  return cudaGetLastError();
}

cudaError_t knn_query_launch(const double2 *sxy, const int *sidx, const double *u_sorted, int n, const double2 *q,
                             int nq, int m, int *out, cudaStream_t stream) {
  knn_query_kernel<<<(nq + 127) / 128, 128, 0, stream>>>(sxy, sidx, u_sorted, n, q, nq, m, out);
  return cudaGetLastError();
}

}  // namespace nngp

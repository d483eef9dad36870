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

// ---- the running neighbour list, in registers -------------------------------------------------------
// The reference keeps dist[0..n) ascending and, on acceptance, overwrites the LAST slot and calls
// rsort_with_index(dist, idx, n) (src/nn.cpp:102-104,174-176,195-197).  With only the last element out of
// place R's Shell sort (gaps 4, 1 for 9 <= n <= 35; plain insertion for n <= 8) does exactly this:
//   gap 4: the new value moves down its own chain n-1, n-5, ... while the element 4 below is greater;
//          every other chain is already ordered and does not move;
//   gap 1: only the chain positions can be out of place, each by at most 3 (the three elements between
//          two chain positions), and they are visited in ascending order.
// So the call is 8 + 24 compare-exchanges at positions that are STATIC when the list is stored right-aligned
// (element k at slot k + MC - n, -inf below the list start so nothing moves past it).  That keeps the list in
// registers -- the generic sort indexes its arrays dynamically, which puts them in local memory: 0.53 s for the
// brute search at n = 10^5, m = 15 -- and reproduces the reference's order under ties (tests: tied grids,
// bit-exact).  Checked against the generic sort on 2*10^5 random lists with ties (tools/shell_check.py).
template <int MC>
struct NeighbourList {
  double d[MC];
  int id[MC];
  __device__ __forceinline__ void init(int n) {
#pragma unroll
    for (int s = 0; s < MC; ++s) {
      d[s] = (s < MC - n) ? -INFINITY : INFINITY;
      id[s] = 0;
    }
  }
  __device__ __forceinline__ double worst() const { return d[MC - 1]; }
  __device__ __forceinline__ bool cswap(int a, int b) {  // a < b static; strict >, like rsort_with_index
    const bool sw = d[a] > d[b];
    const double da = d[a], db = d[b];
    const int ia = id[a], ib = id[b];
    d[a] = sw ? db : da;
    d[b] = sw ? da : db;
    id[a] = sw ? ib : ia;
    id[b] = sw ? ia : ib;
    return sw;
  }
  // dist[n-1] = v; idx[n-1] = iv; rsort_with_index(dist, idx, n)
  __device__ __forceinline__ void replace_worst(double v, int iv, int n) {
    d[MC - 1] = v;
    id[MC - 1] = iv;
    if (n > 8) {
      bool moving = true;
#pragma unroll
      for (int c = MC - 1; c >= 4; c -= 4) {
        if (moving) moving = cswap(c - 4, c);
      }
#pragma unroll
      for (int c = (MC - 1) % 4; c < MC; c += 4) {
        if (c >= 1) cswap(c - 1, c);
        if (c >= 2) cswap(c - 2, c - 1);
        if (c >= 3) cswap(c - 3, c - 2);
      }
    } else {
#pragma unroll
      for (int c = MC - 1; c >= 1 && c >= MC - 8; --c) cswap(c - 1, c);
    }
  }
  // list element k (0 = nearest), k < n
  __device__ __forceinline__ void store(int n, int *__restrict__ oi, double *__restrict__ od, bool take_sqrt) const {
#pragma unroll
    for (int s = 0; s < MC; ++s) {
      const int k = s - (MC - n);
      if (k >= 0) {
        oi[k] = id[s];
        od[k] = take_sqrt ? __dsqrt_rn(d[s]) : d[s];
      }
    }
  }
};

// ---- brute force: one thread per query location, candidate tiles staged in shared memory -----------
constexpr int kBruteThreads = 256;

template <int MC>
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
  NeighbourList<MC> L;
  L.init(mi);
  double worst = INFINITY;  // dist[mi-1]
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
            L.replace_worst(d, j0 + jj, mi);
            worst = L.worst();
            thr2 = __dmul_ru(worst, worst);
          }
        }
      }
    }
  }
  if (mi > 0) {
    const long long off = nn_offset(i, m);
    L.store(mi, nnIndx + off, nnDist + off, false);
  }
}

// ---- code-book search: one thread per position of the (x+y)-sorted order ---------------------------
// sxy / sidx: sites and their NNGP index in ascending u = x + y order.
// Locations with NNGP index < head are left to knn_cb_head_kernel (below).
template <int MC>
__global__ void __launch_bounds__(128) knn_cb_kernel(const double2 *__restrict__ sxy, const int *__restrict__ sidx,
                                                     int n, int m, int head, int *__restrict__ nnIndx,
                                                     double *__restrict__ nnDist) {
  const int ui = blockIdx.x * blockDim.x + threadIdx.x;
  if (ui >= n) return;
  const int myidx = sidx[ui];
  if (myidx < head) return;
  const int mi = nn_count(myidx, m);
  if (mi == 0) return;  // the reference calls fastNN(m=0) here (reads rSNNDist[-1]; writes nothing)
  const double2 me = sxy[ui];
  NeighbourList<MC> L;
  L.init(mi);
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
            L.replace_worst(de, cid, mi);
            worst = L.worst();
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
            L.replace_worst(de, cid, mi);
            worst = L.worst();
          }
        }
      }
    }
  }
  const long long off = nn_offset(myidx, m);
  L.store(mi, nnIndx + off, nnDist + off, true);  // sqrt at the end, src/nn.cpp:206-208
}

// ---- code-book search, head of the ordering: one WARP per location -----------------------------------
// Location i may only take neighbours among its i predecessors.  For small i those are a tiny fraction of the
// sites, so the walk along the u-sorted order runs far before m of them are met (i < m: the whole array, both
// ways) -- as single threads these few walks were a 0.2 s tail at n = 10^6 (of 0.23 s) and 3 s at n = 10^7.
// Here a warp evaluates 32 rounds of the reference's loop (src/nn.cpp:151-204: one candidate down, one up) at
// once and ballots for the first EVENT in the reference's visiting order -- a direction stopping (dmi > 2
// worst, or the array end) or a candidate accepted.  Rounds before it cannot have changed any state, so they
// are consumed wholesale; the event is applied (every lane keeps an identical copy of the list) and the block
// is re-evaluated from there with the new bound.  Acceptance sequence, hence the result under ties, is the
// reference's.
__global__ void knn_head_positions_kernel(const int *__restrict__ sidx, int n, int head, int *__restrict__ headpos) {
  const int ui = blockIdx.x * blockDim.x + threadIdx.x;
  if (ui < n && sidx[ui] < head) headpos[sidx[ui]] = ui;
}

template <int MC>
__global__ void __launch_bounds__(128) knn_cb_head_kernel(const double2 *__restrict__ sxy, const int *__restrict__ sidx,
                                                          const int *__restrict__ headpos, int head, int n, int m,
                                                          int *__restrict__ nnIndx, double *__restrict__ nnDist) {
  const int myidx = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (myidx >= head) return;
  const int mi = nn_count(myidx, m);
  if (mi == 0) return;
  const int ui = headpos[myidx];
  const double2 me = sxy[ui];
  const double su = __dadd_rn(me.x, me.y);
  NeighbourList<MC> L;
  L.init(mi);
  double worst = INFINITY;
  int i = ui, j = ui;  // last positions consumed downwards / upwards
  bool up = true, down = true;
  // one candidate under the bound w: stop (direction ends, incl. running off the array), accept, or neither
  auto eval = [&](int pos, double w, bool &stop, bool &acc, double &de, int &cid) {
    stop = (pos < 0) || (pos > n - 1);
    acc = false;
    de = 0.0;
    cid = 0;
    if (!stop) {
      const double2 pc = sxy[pos];
      const double t = __dsub_rn(__dsub_rn(su, pc.x), pc.y);  // dmi, src/nn.cpp:128-130
      stop = __dmul_rn(t, t) > 2.0 * w;
      const double dx = me.x - pc.x, dy = me.y - pc.y;
      de = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
      cid = sidx[pos];
      acc = !stop && de < w && cid < myidx;
    }
  };
  auto apply = [&](bool stop, double de, int cid, bool &dir_flag) {  // uniform over the warp
    if (stop) {
      dir_flag = false;
    } else {
      L.replace_worst(de, cid, mi);
      worst = L.worst();
    }
  };
  const unsigned full = 0xffffffffu;
  while (up || down) {
    if (up && down) {
      bool sD, aD, sU, aU;
      double deD, deU;
      int cD, cU;
      eval(i - 1 - lane, worst, sD, aD, deD, cD);  // round `lane`: down candidate, then up candidate
      eval(j + 1 + lane, worst, sU, aU, deU, cU);
      const unsigned eD = __ballot_sync(full, sD || aD), eU = __ballot_sync(full, sU || aU);
      const int fD = eD ? __ffs(eD) - 1 : 32, fU = eU ? __ffs(eU) - 1 : 32;
      if (fD == 32 && fU == 32) {
        i -= 32;
        j += 32;
      } else if (fD <= fU) {  // the down candidate of round fD is the first event
        apply(__shfl_sync(full, (int)sD, fD) != 0, __shfl_sync(full, deD, fD), __shfl_sync(full, cD, fD), down);
        i -= fD + 1;
        j += fD;
        // the same round's up candidate, now under the new bound, before rounds realign
        bool s1, a1;
        double d1;
        int c1;
        eval(j + 1, worst, s1, a1, d1, c1);
        if (s1) {
          up = false;
        } else {
          j += 1;
          if (a1) apply(false, d1, c1, up);
        }
      } else {  // the up candidate of round fU is the first event (its round's down candidate was no event)
        apply(__shfl_sync(full, (int)sU, fU) != 0, __shfl_sync(full, deU, fU), __shfl_sync(full, cU, fU), up);
        i -= fU + 1;
        j += fU + 1;
      }
    } else if (down) {
      bool sD, aD;
      double deD;
      int cD;
      eval(i - 1 - lane, worst, sD, aD, deD, cD);
      const unsigned eD = __ballot_sync(full, sD || aD);
      if (!eD) {
        i -= 32;
      } else {
        const int f = __ffs(eD) - 1;
        apply(__shfl_sync(full, (int)sD, f) != 0, __shfl_sync(full, deD, f), __shfl_sync(full, cD, f), down);
        i -= f + 1;
      }
    } else {
      bool sU, aU;
      double deU;
      int cU;
      eval(j + 1 + lane, worst, sU, aU, deU, cU);
      const unsigned eU = __ballot_sync(full, sU || aU);
      if (!eU) {
        j += 32;
      } else {
        const int f = __ffs(eU) - 1;
        apply(__shfl_sync(full, (int)sU, f) != 0, __shfl_sync(full, deU, f), __shfl_sync(full, cU, f), up);
        j += f + 1;
      }
    }
  }
  if (lane == 0) {
    const long long off = nn_offset(myidx, m);
    L.store(mi, nnIndx + off, nnDist + off, true);
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

// ---- unordered query kNN on a uniform grid (the default nn2 stand-in) ---------------------------------------------
// The strip walk above inspects every site whose x + y lies within sqrt(2) of the current radius -- a band across the
// whole domain, ~6,000 candidates per query at n = 10^6 -- and its threads walk unrelated strips (queries arrive in user
// order).  Here the sites are binned into square cells of ~4 sites, the queries are processed in cell order (one warp =
// neighbouring queries = the same cells: broadcast loads), and a query visits its own cell, then ring after ring of
// cells until the m-th best distance is below the distance to everything not yet visited.  Exact; ties resolve by
// ascending (d^2, site index) exactly as in knn_query_kernel and the oracle's exhaustive search.
struct GridGeo {
  double x0, y0, inv_h, h;
  int gx, gy;
};

__device__ __forceinline__ int grid_cell_1d(double v, double v0, double inv_h, int g) {
  const int c = (int)floor((v - v0) * inv_h);
  return c < 0 ? 0 : (c >= g ? g - 1 : c);
}

__global__ void __launch_bounds__(256) grid_cell_kernel(const double2 *__restrict__ pts, int n, GridGeo G,
                                                        int *__restrict__ cell, int *__restrict__ idx) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double2 p = pts[i];
  cell[i] = grid_cell_1d(p.y, G.y0, G.inv_h, G.gy) * G.gx + grid_cell_1d(p.x, G.x0, G.inv_h, G.gx);
  idx[i] = i;
}

// start[c] = first sorted position whose cell id is >= c, c = 0 .. ncells (binary search per cell)
__global__ void __launch_bounds__(256) grid_start_kernel(const int *__restrict__ sorted_cell, int n, int ncells,
                                                         int *__restrict__ start) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > ncells) return;
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sorted_cell[mid] < c) lo = mid + 1;
    else hi = mid;
  }
  start[c] = lo;
}

__global__ void __launch_bounds__(128) knn_grid_query_kernel(const double2 *__restrict__ sxy /* sites in cell order */,
                                                             const int *__restrict__ sidx /* their site ids */,
                                                             const int *__restrict__ start, GridGeo G, int n,
                                                             const double2 *__restrict__ q,
                                                             const int *__restrict__ qorder /* queries in cell order */,
                                                             int nq, int m, int *__restrict__ out /* nq x m col-major */) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nq) return;
  const int qi = qorder[t];
  const double2 me = q[qi];
  const int cx = grid_cell_1d(me.x, G.x0, G.inv_h, G.gx), cy = grid_cell_1d(me.y, G.y0, G.inv_h, G.gy);
  double ld[NNGP_MAX_M];
  int li[NNGP_MAX_M];
  const int mm = min(m, n);
  for (int k = 0; k < NNGP_MAX_M; ++k) {
    ld[k] = INFINITY;
    li[k] = 0x7fffffff;
  }
  double worst = INFINITY;
  int worst_id = 0x7fffffff;
  // (a query outside the sites' bounding box is clamped into a border cell; every unvisited cell then lies on the far
  // side of the visited block, so the bound ring * h below still holds)
  const int rmax = max(max(cx, G.gx - 1 - cx), max(cy, G.gy - 1 - cy));
  for (int ring = 0; ring <= rmax; ++ring) {
    const int y_lo = cy - ring, y_hi = cy + ring;
    for (int yy = max(y_lo, 0); yy <= min(y_hi, G.gy - 1); ++yy) {
      const bool edge_row = (yy == y_lo || yy == y_hi);
      // edge rows: the whole span of the ring; inner rows: only its two end cells
      const int x_lo = cx - ring, x_hi = cx + ring;
      for (int pass = 0; pass < 2; ++pass) {
        int xa, xb;
        if (edge_row) {
          if (pass == 1) break;
          xa = max(x_lo, 0);
          xb = min(x_hi, G.gx - 1);
        } else {
          const int xc = pass == 0 ? x_lo : x_hi;
          if (xc < 0 || xc >= G.gx || (pass == 1 && ring == 0)) continue;
          xa = xb = xc;
        }
        if (xa > xb) continue;
        const int p0 = start[yy * G.gx + xa], p1 = start[yy * G.gx + xb + 1];  // cells of one row are contiguous
        for (int pos = p0; pos < p1; ++pos) {
          const double2 pc = sxy[pos];
          const double dx = me.x - pc.x, dy = me.y - pc.y;
          const double de = __dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy));
          if (de > worst) continue;
          const int cid = sidx[pos];
          if (de < worst || cid < worst_id) {
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
    }
    // everything not yet visited is farther than reach: stop once the m-th best is strictly inside it
    const double reach = (double)ring * G.h;
    if (worst < reach * reach * (1.0 - 1e-9)) break;
  }
  for (int k = 0; k < m; ++k) out[(size_t)k * nq + qi] = (k < mm) ? li[k] : -1;
}

cudaError_t grid_cell_launch(const double2 *pts, int n, double x0, double y0, double h, int gx, int gy, int *cell, int *idx,
                             cudaStream_t stream) {
  GridGeo G{x0, y0, 1.0 / h, h, gx, gy};
  grid_cell_kernel<<<(n + 255) / 256, 256, 0, stream>>>(pts, n, G, cell, idx);
  return cudaGetLastError();
}

cudaError_t grid_start_launch(const int *sorted_cell, int n, int ncells, int *start, cudaStream_t stream) {
  grid_start_kernel<<<(ncells + 1 + 255) / 256, 256, 0, stream>>>(sorted_cell, n, ncells, start);
  return cudaGetLastError();
}

cudaError_t knn_grid_query_launch(const double2 *sxy, const int *sidx, const int *start, double x0, double y0, double h,
                                  int gx, int gy, int n, const double2 *q, const int *qorder, int nq, int m, int *out,
                                  cudaStream_t stream) {
  GridGeo G{x0, y0, 1.0 / h, h, gx, gy};
  knn_grid_query_kernel<<<(nq + 127) / 128, 128, 0, stream>>>(sxy, sidx, start, G, n, q, qorder, nq, m, out);
  return cudaGetLastError();
}

// ---- mkUIndx: the transposed neighbour graph (src/nn.cpp:25-56, crs_csc src/util.cpp:151-185) ------------
// uIndx lists, for every location i, the locations that have i as a neighbour (ascending), uiIndx the slot i
// occupies in each of those neighbour lists.  The reference builds it serially; here the nIndx (neighbour id,
// entry number) pairs are sorted by id with a STABLE radix sort -- entries are numbered row by row, so equal
// ids stay in ascending row order, which is the reference's order -- and the row / slot of an entry come from
// the closed form of the ragged layout (src/util.cpp:55-70).
__global__ void uindx_keys_kernel(const int *__restrict__ nnIndx, long long nIndx, int *__restrict__ keys,
                                  int *__restrict__ vals, int *__restrict__ counts) {
  const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nIndx) return;
  const int j = nnIndx[e];
  keys[e] = j;
  vals[e] = (int)e;
  atomicAdd(&counts[j], 1);
}

__global__ void uindx_scatter_kernel(const int *__restrict__ sorted_entry, long long nIndx, int m, int *__restrict__ uIndx,
                                     int *__restrict__ uiIndx) {
  const long long pos = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (pos >= nIndx) return;
  const long long e = sorted_entry[pos];
  const long long head = (long long)m * (m - 1) / 2;  // entries of the ragged rows 1 .. m-1
  long long i, k;
  if (e >= head) {
    i = m + (e - head) / m;
    k = (e - head) % m;
  } else {  // row i holds entries i(i-1)/2 .. i(i+1)/2 - 1
    i = (long long)((1.0 + sqrt(1.0 + 8.0 * (double)e)) * 0.5);
    while (i * (i - 1) / 2 > e) --i;
    while (i * (i + 1) / 2 <= e) ++i;
    k = e - i * (i - 1) / 2;
  }
  uIndx[pos] = (int)i;
  uiIndx[pos] = (int)k;
}

cudaError_t uindx_keys_launch(const int *nnIndx, long long nIndx, int *keys, int *vals, int *counts, cudaStream_t stream) {
  uindx_keys_kernel<<<(unsigned)((nIndx + 255) / 256), 256, 0, stream>>>(nnIndx, nIndx, keys, vals, counts);
  return cudaGetLastError();
}

cudaError_t uindx_scatter_launch(const int *sorted_entry, long long nIndx, int m, int *uIndx, int *uiIndx, cudaStream_t stream) {
  uindx_scatter_kernel<<<(unsigned)((nIndx + 255) / 256), 256, 0, stream>>>(sorted_entry, nIndx, m, uIndx, uiIndx);
  return cudaGetLastError();
}

// ---- launchers ------------------------------------------------------------------------------------
cudaError_t knn_brute_launch(const double2 *coords, int n, int m, int *nnIndx, double *nnDist, cudaStream_t stream) {
  const int nchunks = (n + kBruteThreads - 1) / kBruteThreads;
  if (m <= 8) knn_brute_kernel<8><<<nchunks, kBruteThreads, 0, stream>>>(coords, n, m, nnIndx, nnDist);
  else if (m <= 16) knn_brute_kernel<16><<<nchunks, kBruteThreads, 0, stream>>>(coords, n, m, nnIndx, nnDist);
  else if (m <= 24) knn_brute_kernel<24><<<nchunks, kBruteThreads, 0, stream>>>(coords, n, m, nnIndx, nnDist);
  else knn_brute_kernel<32><<<nchunks, kBruteThreads, 0, stream>>>(coords, n, m, nnIndx, nnDist);
  return cudaGetLastError();
}

// headpos: scratch of knn_cb_head_count(n, m) ints
int knn_cb_head_count(int n, int m) {
  long long h = (long long)n * m / 2048;  // a single thread's walk at index `head` is ~2 n m / head ~ 4k candidates
  if (h < 1024) h = 1024;
  if (h > n) h = n;
  return (int)h;
}

cudaError_t knn_cb_launch(const double2 *sxy, const int *sidx, int n, int m, int *headpos, int *nnIndx, double *nnDist,
                          cudaStream_t stream) {
  const int grid = (n + 127) / 128;
  const int head = knn_cb_head_count(n, m);
  const int hgrid = (int)(((long long)head * 32 + 127) / 128);
  knn_head_positions_kernel<<<(n + 255) / 256, 256, 0, stream>>>(sidx, n, head, headpos);
#define NNGP_CB(MCV)                                                                                     \
  do {                                                                                                   \
    knn_cb_head_kernel<MCV><<<hgrid, 128, 0, stream>>>(sxy, sidx, headpos, head, n, m, nnIndx, nnDist);  \
    knn_cb_kernel<MCV><<<grid, 128, 0, stream>>>(sxy, sidx, n, m, head, nnIndx, nnDist);                 \
  } while (0)
  if (m <= 8) NNGP_CB(8);
  else if (m <= 16) NNGP_CB(16);
  else if (m <= 24) NNGP_CB(24);
  else NNGP_CB(32);
#undef NNGP_CB
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

// nngp_abi_knn.cu -- C ABI for the neighbour searches (include/nngp_b200.h): nngp_mk_nn_indx replaces
// the reference's `mkNNIndx` / `mkNNIndxCB` .Call entry points (src/nn.cpp:63-111, 214-257);
// nngp_nn_query stands in for RANN::nn2 (R/generics.R:642).
#include <cub/cub.cuh>

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include <algorithm>
#include <cmath>
#include <vector>

#include "nngp_ctx.h"
#include "host_stage.h"

namespace nngp {
cudaError_t knn_brute_launch(const double2 *coords, int n, int m, int *nnIndx, double *nnDist, cudaStream_t stream);
int knn_cb_head_count(int n, int m);
cudaError_t knn_cb_launch(const double2 *sxy, const int *sidx, int n, int m, int *headpos, int *nnIndx, double *nnDist,
                          cudaStream_t stream);
cudaError_t uindx_keys_launch(const int *nnIndx, long long nIndx, int *keys, int *vals, int *counts, cudaStream_t stream);
cudaError_t uindx_scatter_launch(const int *sorted_entry, long long nIndx, int m, int *uIndx, int *uiIndx, cudaStream_t stream);
cudaError_t knn_keys_launch(const double2 *coords, int n, double *u, int *idx, cudaStream_t stream);
cudaError_t knn_gather_launch(const double2 *coords, const int *sidx, int n, double2 *sxy, cudaStream_t stream);
cudaError_t knn_count_ties_launch(const double *u_sorted, int n, int *count, cudaStream_t stream);
cudaError_t knn_query_launch(const double2 *sxy, const int *sidx, const double *u_sorted, int n, const double2 *q,
                             int nq, int m, int *out, cudaStream_t stream);
cudaError_t grid_cell_launch(const double2 *pts, int n, double x0, double y0, double h, int gx, int gy, int *cell, int *idx,
                             cudaStream_t stream);
cudaError_t grid_start_launch(const int *sorted_cell, int n, int ncells, int *start, cudaStream_t stream);
cudaError_t knn_grid_query_launch(const double2 *sxy, const int *sidx, const int *start, double x0, double y0, double h,
                                  int gx, int gy, int n, const double2 *q, const int *qorder, int nq, int m, int *out,
                                  cudaStream_t stream);

// R's rsort_with_index on the host (src/main/sort.c; third-party algorithm restated): only used
// when u = x + y has exact ties, where the reference's Shell-sort order decides the scan order.
static void host_rsort_with_index(double *x, int *indx, int n) {
  int h;
  for (h = 1; h <= n / 9; h = 3 * h + 1) {
  }
  for (; h > 0; h /= 3)
    for (int i = h; i < n; i++) {
      const double v = x[i];
      const int iv = indx[i];
      int j = i;
      while (j >= h && (x[j - h] > v || (x[j - h] != x[j - h] && v == v))) {
        x[j] = x[j - h];
        indx[j] = indx[j - h];
        j -= h;
      }
      x[j] = v;
      indx[j] = iv;
    }
}

// sort the sites by u = x + y on the device; falls back to the reference's exact Shell sort on the
// host when u has ties.  Outputs: d_sidx (NNGP index per sorted position), d_u (sorted keys).
static int sort_by_u(nngp_ctx *c, const double2 *d_coords, int n, DevBuf &d_sidx, DevBuf &d_u, int *n_ties) {
  DevBuf u_in, idx_in, tmp, ties;
  NNGP_CUDA(c, u_in.alloc((size_t)n * sizeof(double)));
  NNGP_CUDA(c, idx_in.alloc((size_t)n * sizeof(int)));
  NNGP_CUDA(c, d_u.alloc((size_t)n * sizeof(double)));
  NNGP_CUDA(c, d_sidx.alloc((size_t)n * sizeof(int)));
  NNGP_CUDA(c, ties.alloc(sizeof(int)));
  NNGP_CUDA(c, knn_keys_launch(d_coords, n, u_in.as<double>(), idx_in.as<int>(), c->stream));
  size_t tmp_bytes = 0;
  NNGP_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, u_in.as<double>(), d_u.as<double>(),
                                               idx_in.as<int>(), d_sidx.as<int>(), n, 0, 64, c->stream));
  NNGP_CUDA(c, tmp.alloc(tmp_bytes));
  NNGP_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, u_in.as<double>(), d_u.as<double>(),
                                               idx_in.as<int>(), d_sidx.as<int>(), n, 0, 64, c->stream));
  NNGP_CUDA(c, cudaMemsetAsync(ties.p, 0, sizeof(int), c->stream));
  NNGP_CUDA(c, knn_count_ties_launch(d_u.as<double>(), n, ties.as<int>(), c->stream));
  c->launches += 4;
  int nt = 0;
  NNGP_CUDA(c, cudaMemcpyAsync(&nt, ties.p, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  *n_ties = nt;
  if (nt > 0) {  // exact tie order: replay the reference's serial sort (src/nn.cpp:238-243)
    std::vector<double> u((size_t)n);
    std::vector<int> idx((size_t)n);
    NNGP_CUDA(c, cudaMemcpy(u.data(), u_in.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost));
    for (int i = 0; i < n; i++) idx[i] = i;
    host_rsort_with_index(u.data(), idx.data(), n);
    NNGP_CUDA(c, cudaMemcpy(d_u.p, u.data(), (size_t)n * sizeof(double), cudaMemcpyHostToDevice));
    NNGP_CUDA(c, cudaMemcpy(d_sidx.p, idx.data(), (size_t)n * sizeof(int), cudaMemcpyHostToDevice));
  }
  return NNGP_OK;
}

static int ensure_stager(nngp_ctx *c) {
  if (!c->stager) {
    c->stager = new HostStager();
    NNGP_CUDA(c, c->stager->init(c->device));
  }
  return NNGP_OK;
}

// column-major (x[n], y[n]) host sites -> interleaved double2 on the device, through the pinned ring
static int upload_sites(nngp_ctx *c, const double *coords, int n, DevBuf &d) {
  NNGP_CUDA(c, d.alloc((size_t)n * sizeof(double2)));
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));  // the block may be recycled: earlier work on it has to be over before
  int rc = ensure_stager(c);                        // the stager's own streams write into it
  if (rc) return rc;
  std::vector<StageJob> up{{StageJob::XY_INTERLEAVE, d.p, coords, (size_t)n * sizeof(double2), n}};
  NNGP_CUDA(c, c->stager->run(up, nullptr));
  return NNGP_OK;
}

}  // namespace nngp

using namespace nngp;

extern "C" {

int nngp_mk_nn_indx(nngp_ctx *c, int n, int m, const double *coords, int search_type, int *nnIndx, double *nnDist,
                    int *nnIndxLU, int *n_tie_replayed) {
  if (!c) return NNGP_EINVAL;
  if (n < 1 || m < 1 || !coords || !nnIndx || !nnDist || !nnIndxLU) return ctx_fail(c, NNGP_EINVAL, "nngp_mk_nn_indx: bad arguments");
  if (m > NNGP_MAX_M) return ctx_fail(c, NNGP_EINVAL, "nngp_mk_nn_indx: n.neighbors > 32 is not supported");
  if (search_type != NNGP_SEARCH_BRUTE && search_type != NNGP_SEARCH_CB) return ctx_fail(c, NNGP_EINVAL, "search.type must be brute (0) or cb (1)");
  NNGP_CUDA(c, cudaSetDevice(c->device));
  ScratchScope scratch(c->stream);
  const long long nIndx = nn_offset(n, m);
  DevBuf d_coords, d_idx, d_dist;
  int rc = upload_sites(c, coords, n, d_coords);
  if (rc) return rc;
  NNGP_CUDA(c, d_idx.alloc((size_t)(nIndx > 0 ? nIndx : 1) * sizeof(int)));
  NNGP_CUDA(c, d_dist.alloc((size_t)(nIndx > 0 ? nIndx : 1) * sizeof(double)));
  int nties = 0;
  if (search_type == NNGP_SEARCH_BRUTE) {
    NNGP_CUDA(c, knn_brute_launch(d_coords.as<double2>(), n, m, d_idx.as<int>(), d_dist.as<double>(), c->stream));
    c->launches++;
  } else {
    DevBuf d_sidx, d_u, d_sxy, d_head;
    rc = sort_by_u(c, d_coords.as<double2>(), n, d_sidx, d_u, &nties);
    if (rc) return rc;
    NNGP_CUDA(c, d_sxy.alloc((size_t)n * sizeof(double2)));
    NNGP_CUDA(c, d_head.alloc((size_t)knn_cb_head_count(n, m) * sizeof(int)));
    NNGP_CUDA(c, knn_gather_launch(d_coords.as<double2>(), d_sidx.as<int>(), n, d_sxy.as<double2>(), c->stream));
    NNGP_CUDA(c, knn_cb_launch(d_sxy.as<double2>(), d_sidx.as<int>(), n, m, d_head.as<int>(), d_idx.as<int>(),
                               d_dist.as<double>(), c->stream));
    c->launches += 4;
    NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  }
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  if (nIndx > 0) {  // 12 bytes per neighbour slot back to pageable arrays: threaded pinned ring (host_stage.h)
    rc = ensure_stager(c);
    if (rc) return rc;
    std::vector<StageJob> down{{StageJob::RAW, nnIndx, d_idx.p, (size_t)nIndx * sizeof(int), 0},
                               {StageJob::RAW, nnDist, d_dist.p, (size_t)nIndx * sizeof(double), 0}};
    NNGP_CUDA(c, c->stager->run_d2h(down));
  }
  for (int i = 0; i < n; i++) {  // src/util.cpp:55-70 via src/nn.cpp:95-97
    nnIndxLU[i] = (i == 0) ? 0 : (int)nn_offset(i, m);
    nnIndxLU[n + i] = nn_count(i, m);
  }
  if (n_tie_replayed) *n_tie_replayed = nties;
  return NNGP_OK;
}

int nngp_mk_u_indx(nngp_ctx *c, int n, int m, const int *nnIndx, int search_type, int *uIndx, int *uIndxLU, int *uiIndx) {
  if (!c) return NNGP_EINVAL;
  if (n < 1 || m < 1 || !nnIndx || !uIndx || !uIndxLU || !uiIndx) return ctx_fail(c, NNGP_EINVAL, "nngp_mk_u_indx: bad arguments");
  NNGP_CUDA(c, cudaSetDevice(c->device));
  ScratchScope scratch(c->stream);
  const long long nIndx = nn_offset(n, m);
  if (nIndx > 0x7fffffffLL) return ctx_fail(c, NNGP_EINVAL, "nngp_mk_u_indx: more than 2^31 neighbour slots");
  for (long long e = 0; e < nIndx; e++)
    if (nnIndx[e] < 0 || nnIndx[e] >= n) return ctx_fail(c, NNGP_EINVAL, "nngp_mk_u_indx: neighbour index out of range");
  int rc = ensure_stager(c);
  if (rc) return rc;
  std::vector<int> counts((size_t)n);
  if (nIndx > 0) {
    DevBuf d_nn, d_keys, d_vals, d_keys2, d_vals2, d_counts, d_u, d_ui, tmp;
    const size_t ib = (size_t)nIndx * sizeof(int);
    NNGP_CUDA(c, d_nn.alloc(ib)); NNGP_CUDA(c, d_keys.alloc(ib)); NNGP_CUDA(c, d_vals.alloc(ib));
    NNGP_CUDA(c, d_keys2.alloc(ib)); NNGP_CUDA(c, d_vals2.alloc(ib)); NNGP_CUDA(c, d_u.alloc(ib)); NNGP_CUDA(c, d_ui.alloc(ib));
    NNGP_CUDA(c, d_counts.alloc((size_t)n * sizeof(int)));
    std::vector<StageJob> up{{StageJob::RAW, d_nn.p, nnIndx, ib, 0}};
    NNGP_CUDA(c, cudaStreamSynchronize(c->stream));  // recycled pool block: the stager writes from its own streams
    NNGP_CUDA(c, c->stager->run(up, nullptr));
    NNGP_CUDA(c, cudaMemsetAsync(d_counts.p, 0, (size_t)n * sizeof(int), c->stream));
    NNGP_CUDA(c, uindx_keys_launch(d_nn.as<int>(), nIndx, d_keys.as<int>(), d_vals.as<int>(), d_counts.as<int>(), c->stream));
    int bits = 1;
    while ((1LL << bits) < n) ++bits;
    size_t tmp_bytes = 0;
    NNGP_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys.as<int>(), d_keys2.as<int>(), d_vals.as<int>(),
                                                 d_vals2.as<int>(), (int)nIndx, 0, bits, c->stream));
    NNGP_CUDA(c, tmp.alloc(tmp_bytes));
    NNGP_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, d_keys.as<int>(), d_keys2.as<int>(), d_vals.as<int>(),
                                                 d_vals2.as<int>(), (int)nIndx, 0, bits, c->stream));
    NNGP_CUDA(c, uindx_scatter_launch(d_vals2.as<int>(), nIndx, m, d_u.as<int>(), d_ui.as<int>(), c->stream));
    c->launches += 3;
    NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
    std::vector<StageJob> down{{StageJob::RAW, uIndx, d_u.p, ib, 0}, {StageJob::RAW, uiIndx, d_ui.p, ib, 0},
                               {StageJob::RAW, counts.data(), d_counts.p, (size_t)n * sizeof(int), 0}};
    NNGP_CUDA(c, c->stager->run_d2h(down));
  }
  long long run = 0;  // offsets: running sum of the per-location counts (crs_csc, src/util.cpp:163-168)
  for (int i = 0; i < n; i++) {
    uIndxLU[i] = (int)run;
    uIndxLU[n + i] = counts[i];
    run += counts[i];
  }
  if (search_type == 1) {  // mkUIndx1 walks j = n-1 .. i+1: the same lists, descending (src/util.cpp:93-111)
    for (int i = 0; i < n; i++) {
      int a = uIndxLU[i], b = a + uIndxLU[n + i] - 1;
      for (; a < b; a++, b--) {
        std::swap(uIndx[a], uIndx[b]);
        std::swap(uiIndx[a], uiIndx[b]);
      }
    }
  }
  return NNGP_OK;
}

// cell ids of `cnt` points -> (sorted cell ids, point ids in cell order)
static int grid_sort(nngp_ctx *c, const double2 *d_pts, int cnt, double x0, double y0, double h, int gx, int gy,
                     DevBuf &d_cell_sorted, DevBuf &d_order) {
  DevBuf d_cell, d_idx, tmp;
  const size_t ib = (size_t)cnt * sizeof(int);
  NNGP_CUDA(c, d_cell.alloc(ib)); NNGP_CUDA(c, d_idx.alloc(ib));
  NNGP_CUDA(c, d_cell_sorted.alloc(ib)); NNGP_CUDA(c, d_order.alloc(ib));
  NNGP_CUDA(c, grid_cell_launch(d_pts, cnt, x0, y0, h, gx, gy, d_cell.as<int>(), d_idx.as<int>(), c->stream));
  int bits = 1;
  while ((1LL << bits) < (long long)gx * gy) ++bits;
  size_t tmp_bytes = 0;
  NNGP_CUDA(c, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_cell.as<int>(), d_cell_sorted.as<int>(), d_idx.as<int>(),
                                               d_order.as<int>(), cnt, 0, bits, c->stream));
  NNGP_CUDA(c, tmp.alloc(tmp_bytes));
  NNGP_CUDA(c, cub::DeviceRadixSort::SortPairs(tmp.p, tmp_bytes, d_cell.as<int>(), d_cell_sorted.as<int>(), d_idx.as<int>(),
                                               d_order.as<int>(), cnt, 0, bits, c->stream));
  c->launches += 2;
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));  // tmp and the unsorted buffers are freed on return
  return NNGP_OK;
}

int nngp_nn_query(nngp_ctx *c, int n, const double *coords, int q, const double *coords0, int m, int *nnIndx0) {
  if (!c) return NNGP_EINVAL;
  if (n < 1 || q < 0 || m < 1 || !coords || (q > 0 && (!coords0 || !nnIndx0))) return ctx_fail(c, NNGP_EINVAL, "nngp_nn_query: bad arguments");
  if (m > NNGP_MAX_M) return ctx_fail(c, NNGP_EINVAL, "nngp_nn_query: k > 32 is not supported");
  if (q == 0) return NNGP_OK;
  NNGP_CUDA(c, cudaSetDevice(c->device));
  ScratchScope scratch(c->stream);
  PhaseTrace tr("nngp_nn_query");
  DevBuf d_coords, d_q, d_sidx, d_u, d_sxy, d_out;
  int rc = upload_sites(c, coords, n, d_coords);
  if (rc) return rc;
  rc = upload_sites(c, coords0, q, d_q);
  if (rc) return rc;
  tr.mark("upload");
  NNGP_CUDA(c, d_sxy.alloc((size_t)n * sizeof(double2)));
  NNGP_CUDA(c, d_out.alloc((size_t)q * m * sizeof(int)));
  tr.mark("alloc");
  static int use_strip = -1;  // NNGP_NN_QUERY=strip: the (x+y)-strip walk (kept for cross-checks)
  if (use_strip < 0) {
    const char *e = getenv("NNGP_NN_QUERY");
    use_strip = (e && !strcmp(e, "strip")) ? 1 : 0;
  }
  if (use_strip) {
    int nties = 0;
    rc = sort_by_u(c, d_coords.as<double2>(), n, d_sidx, d_u, &nties);
    if (rc) return rc;
    NNGP_CUDA(c, knn_gather_launch(d_coords.as<double2>(), d_sidx.as<int>(), n, d_sxy.as<double2>(), c->stream));
    NNGP_CUDA(c, knn_query_launch(d_sxy.as<double2>(), d_sidx.as<int>(), d_u.as<double>(), n, d_q.as<double2>(), q, m,
                                  d_out.as<int>(), c->stream));
    c->launches += 2;
  } else {
    // uniform grid over the sites' bounding box, ~4 sites per cell (knn_kernels.cu: knn_grid_query_kernel)
    double x0 = coords[0], x1 = coords[0], y0 = coords[n], y1 = coords[n];
    for (int i = 1; i < n; i++) {
      x0 = std::min(x0, coords[i]); x1 = std::max(x1, coords[i]);
      y0 = std::min(y0, coords[(size_t)n + i]); y1 = std::max(y1, coords[(size_t)n + i]);
    }
    if (!(x1 - x0 >= 0.0) || !(y1 - y0 >= 0.0) || !std::isfinite(x1 - x0) || !std::isfinite(y1 - y0))
      return ctx_fail(c, NNGP_EINVAL, "nngp_nn_query: coordinates must be finite");
    const double wx = x1 - x0, wy = y1 - y0;
    double h;
    if (wx > 0.0 && wy > 0.0) h = sqrt(wx * wy * 4.0 / n);
    else h = std::max(wx, wy) * 4.0 / n;  // collinear sites: a 1-D grid
    if (!(h > 0.0)) h = 1.0;              // all sites coincide: one cell
    long long gx, gy;
    for (;;) {
      gx = (long long)floor(wx / h) + 1;
      gy = (long long)floor(wy / h) + 1;
      if (gx * gy <= (1LL << 26) && gx < (1LL << 20) && gy < (1LL << 20)) break;
      h *= 1.5;  // very elongated boxes: coarser cells keep the table small
    }
    tr.mark("bounding box");
    DevBuf d_cell_sorted, d_qcell_sorted, d_qorder, d_start;
    rc = grid_sort(c, d_coords.as<double2>(), n, x0, y0, h, (int)gx, (int)gy, d_cell_sorted, d_sidx);
    if (rc) return rc;
    rc = grid_sort(c, d_q.as<double2>(), q, x0, y0, h, (int)gx, (int)gy, d_qcell_sorted, d_qorder);
    if (rc) return rc;
    tr.mark("cell sorts");
    const int ncells = (int)(gx * gy);
    NNGP_CUDA(c, d_start.alloc(((size_t)ncells + 1) * sizeof(int)));
    NNGP_CUDA(c, grid_start_launch(d_cell_sorted.as<int>(), n, ncells, d_start.as<int>(), c->stream));
    NNGP_CUDA(c, knn_gather_launch(d_coords.as<double2>(), d_sidx.as<int>(), n, d_sxy.as<double2>(), c->stream));
    NNGP_CUDA(c, knn_grid_query_launch(d_sxy.as<double2>(), d_sidx.as<int>(), d_start.as<int>(), x0, y0, h, (int)gx, (int)gy,
                                       n, d_q.as<double2>(), d_qorder.as<int>(), q, m, d_out.as<int>(), c->stream));
    c->launches += 3;
  }
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  tr.mark("search kernels");
  NNGP_CUDA(c, cudaMemcpy(nnIndx0, d_out.p, (size_t)q * m * sizeof(int), cudaMemcpyDeviceToHost));
  tr.mark("result copy");
  return NNGP_OK;
}

}  // extern "C"

// nngp_abi.cu -- C ABI (include/nngp_b200.h): lifecycle, model data, the likelihood hot path and
// the response-model sampler `rNNGP` (reference src/rNNGP.cpp:105-420) driven from the host with
// the caller's RNG.  There is no CPU fallback anywhere in this library.
#include <limits.h>
#include <math.h>
#include <stdio.h>
#include <string.h>

#include <chrono>
#include <functional>
#include <future>

#include <mutex>
#include <string>
#include <vector>

#include "host_linalg.h"
#include "nngp_ctx.h"
#include "host_stage.h"

namespace nngp {

static std::mutex g_err_mu;
static std::string g_err;

void set_global_error(const std::string &msg) {
  std::lock_guard<std::mutex> lk(g_err_mu);
  g_err = msg;
}

int ctx_fail(nngp_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg;
  set_global_error(msg);
  return code;
}

int ctx_cuda(nngp_ctx *ctx, cudaError_t e, const char *what) {
  std::string msg = std::string("CUDA error: ") + cudaGetErrorString(e) + " in " + what;
  cudaGetLastError();  // clear sticky state where possible
  return ctx_fail(ctx, (e == cudaErrorMemoryAllocation) ? NNGP_ENOMEM : NNGP_ECUDA, msg);
}

template <typename Tp>
static void free_dev(Tp *&p) {
  if (p) cudaFree(p);
  p = nullptr;
}

// grow-only device buffer: reallocates when `count` elements no longer fit
template <typename Tp>
static int ensure_dev(nngp_ctx *c, Tp *&p, size_t &cap_bytes, size_t count) {
  const size_t bytes = (count ? count : 1) * sizeof(Tp);
  if (p && bytes <= cap_bytes) return NNGP_OK;
  if (p) cudaFree(p);
  p = nullptr;
  cap_bytes = 0;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e == cudaErrorMemoryAllocation) {
    // the scratch pool (DevBuf, nngp_ctx.h) keeps freed blocks cached: hand them back and try once more
    cudaGetLastError();
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) {
      cudaStreamSynchronize(c->stream);
      cudaMemPoolTrimTo(pool, 0);
    }
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) return ctx_cuda(c, e, "cudaMalloc (model buffer)");
  cap_bytes = bytes;
  return NNGP_OK;
}

static void free_model(nngp_ctx *c) {
  free_dev(c->d_coords);
  free_dev(c->d_nnIndx_store);
  c->d_nnIndx = nullptr;
  free_dev(c->d_X);
  free_dev(c->d_y);
  for (int s = 0; s < 2; s++) {
    free_dev(c->d_B_store[s]);
    free_dev(c->d_F_store[s]);
    c->d_B[s] = c->d_F[s] = nullptr;
  }
  free_dev(c->d_tri);
  c->tri_cap = 0;
  for (int b = 0; b < NNGP_MAX_BATCH; b++) free_dev(c->d_resid[b]);
  free_dev(c->d_beta);
  free_dev(c->d_partials);
  free_dev(c->d_out);
  if (c->h_beta) cudaFreeHost(c->h_beta);
  c->h_beta = nullptr;
  c->h_beta_cap = 0;
  for (size_t &x : c->cap) x = 0;
  c->n = c->m = c->p = 0;
}

}  // namespace nngp

using namespace nngp;

extern "C" {

int nngp_abi_version(void) { return NNGP_ABI_VERSION; }

int nngp_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

const char *nngp_last_error(const nngp_ctx *ctx) {
  if (ctx) return ctx->err.c_str();
  static thread_local std::string copy;
  std::lock_guard<std::mutex> lk(g_err_mu);
  copy = g_err;
  return copy.c_str();
}

static int create_init(nngp_ctx *c, int device, void *stream_handle) {
  NNGP_CUDA(c, cudaSetDevice(device));
  int major = 0, sms = 0;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
  if (major < 10) return ctx_fail(nullptr, NNGP_ENODEV, "nngp_create: device is not sm_100-class (kernels are built for sm_100a only)");
  c->sms = sms > 0 ? sms : 148;
  if (stream_handle) {
    c->stream = (cudaStream_t)stream_handle;
  } else {
    NNGP_CUDA(c, cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    c->own_stream = true;
  }
  {  // keep freed scratch blocks cached in the stream-ordered pool (DevBuf, nngp_ctx.h)
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, c->device) == cudaSuccess) {
      unsigned long long thr = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr);
    }
  }
  NNGP_CUDA(c, cudaMalloc(&c->d_ticket, 2 * NNGP_MAX_BATCH * sizeof(unsigned int)));  // tickets, tile counters
  NNGP_CUDA(c, cudaMemset(c->d_ticket, 0, 2 * NNGP_MAX_BATCH * sizeof(unsigned int)));
  NNGP_CUDA(c, cudaMalloc(&c->d_fail, NNGP_MAX_BATCH * sizeof(int)));
  {
    int init[NNGP_MAX_BATCH];
    for (int b = 0; b < NNGP_MAX_BATCH; b++) init[b] = INT_MAX;
    NNGP_CUDA(c, cudaMemcpy(c->d_fail, init, sizeof init, cudaMemcpyHostToDevice));
  }
  NNGP_CUDA(c, cudaMalloc(&c->d_mat, NNGP_MAX_BATCH * sizeof(MaternConsts)));
  NNGP_CUDA(c, cudaHostAlloc(&c->h_out, kBfOutStride * NNGP_MAX_BATCH * sizeof(double), cudaHostAllocMapped));
  NNGP_CUDA(c, cudaMalloc(&c->d_red, kBfOutStride * NNGP_MAX_BATCH * sizeof(double)));
  NNGP_CUDA(c, cudaHostAlloc(&c->h_red, kBfOutStride * NNGP_MAX_BATCH * sizeof(double), cudaHostAllocDefault));
  NNGP_CUDA(c, cudaHostAlloc(&c->h_fail, NNGP_MAX_BATCH * sizeof(int), cudaHostAllocMapped));
  {
    int rc = mbox_local_init(c);
    if (rc) return rc;
  }
  NNGP_CUDA(c, cudaHostAlloc(&c->h_mat, NNGP_MAX_BATCH * sizeof(MaternConsts), cudaHostAllocDefault));
  NNGP_CUDA(c, cudaEventCreate(&c->ev0));
  NNGP_CUDA(c, cudaEventCreate(&c->ev1));
  return NNGP_OK;
}

int nngp_create(int device, void *stream_handle, nngp_ctx **out) {
  if (!out) return ctx_fail(nullptr, NNGP_EINVAL, "nngp_create: out is NULL");
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    return ctx_fail(nullptr, NNGP_ENODEV,
                    "nngp_create: no CUDA device visible (this library has no CPU fallback)");
  }
  if (device < 0) {
    if (cudaGetDevice(&device) != cudaSuccess) device = 0;
  }
  if (device >= ndev) return ctx_fail(nullptr, NNGP_EINVAL, "nngp_create: device index out of range");
  nngp_ctx *c = new nngp_ctx();
  c->device = device;
  const int rc = create_init(c, device, stream_handle);
  if (rc) {  // nothing half-built survives a failed create
    nngp_destroy(c);
    return rc;
  }
  *out = c;
  return NNGP_OK;
}

void nngp_destroy(nngp_ctx *c) {
  if (!c) return;
  cudaSetDevice(c->device);
  if (c->stream) cudaStreamSynchronize(c->stream);
  free_model(c);
  free_dev(c->d_ticket);
  free_dev(c->d_hist);
  if (c->h_hist) cudaFreeHost(c->h_hist);
  if (c->ev_hist) cudaEventDestroy(c->ev_hist);
  free_dev(c->d_Xy);
  free_dev(c->d_fail);
  free_dev(c->d_mat);
  if (c->d_flush) cudaFree(c->d_flush);
  nngp_comm_destroy(c);
  mbox_release(c);
  if (c->h_out) cudaFreeHost(c->h_out);
  if (c->h_red) cudaFreeHost(c->h_red);
  free_dev(c->d_red);
  if (c->h_fail) cudaFreeHost(c->h_fail);
  if (c->h_mat) cudaFreeHost(c->h_mat);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  delete c->stager;
  delete c->stager2;
  if (c->d_rep_perm) cudaFree(c->d_rep_perm);
  if (c->d_rep_levoff) cudaFree(c->d_rep_levoff);
  if (c->d_rep_z) cudaFree(c->d_rep_z);
  if (c->d_rep_v) cudaFree(c->d_rep_v);
  if (c->h_z) cudaFreeHost(c->h_z);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
}

int nngp_sync(nngp_ctx *c) {
  if (!c) return NNGP_EINVAL;
  NNGP_CUDA(c, cudaSetDevice(c->device));
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  return NNGP_OK;
}

long long nngp_launch_count(const nngp_ctx *c) { return c ? c->launches : 0; }

int nngp_set_option(nngp_ctx *c, const char *name, int value) {
  if (!c || !name) return NNGP_EINVAL;
  if (!strcmp(name, "matern_short")) {
    c->opt_matern_short = value < 0 ? 0 : (value > 2 ? 2 : value);  // 0: full series; 1: at most the short one; 2: tiny allowed
    return NNGP_OK;
  }
  if (!strcmp(name, "tri_cache")) {  // takes effect at the next nngp_set_data
    c->opt_tri = value ? 1 : 0;
    return NNGP_OK;
  }
  if (!strcmp(name, "tri_force")) {  // 1: the distance-triangle kernel for every 2 <= m <= 32, not only where it is fastest
    tri_set_force(value ? 1 : 0);    // (process-wide; takes effect at the next nngp_set_data)
    return NNGP_OK;
  }
  if (!strcmp(name, "reduce_nccl")) {  // 1: sums over ranks go through ncclAllReduce instead of the shared mailbox
    c->opt_reduce_nccl = value ? 1 : 0;
    return NNGP_OK;
  }
  return ctx_fail(c, NNGP_EINVAL, std::string("nngp_set_option: unknown option ") + name);
}

int nngp_get_option(nngp_ctx *c, const char *name, int *value) {
  if (!c || !name || !value) return NNGP_EINVAL;
  if (!strcmp(name, "matern_short")) {
    *value = c->opt_matern_short;
    return NNGP_OK;
  }
  if (!strcmp(name, "tri_cache")) {
    *value = c->opt_tri;
    return NNGP_OK;
  }
  if (!strcmp(name, "reduce_nccl")) {
    *value = c->opt_reduce_nccl;
    return NNGP_OK;
  }
  if (!strcmp(name, "mailbox_ranks")) {  // ranks whose slots this context reads (1: private mailbox)
    *value = c->mbox_ranks;
    return NNGP_OK;
  }
  if (!strcmp(name, "tri_cache_used")) {  // 1 when the B/F launches of this data set stream the resident distance triangle
    *value = c->d_tri != nullptr && c->tri_on;
    return NNGP_OK;
  }
  if (!strcmp(name, "matern_short_used")) {  // parameter sets of the last likelihood launch that ran the short series
    *value = 0;
    for (int b = 0; b < NNGP_MAX_BATCH; b++) *value += c->bf_short[b] != 0;
    return NNGP_OK;
  }
  if (!strcmp(name, "matern_tiny_used")) {  // ... and how many of them the kMatTinyTerms-term one (bf_tri_kernel)
    *value = 0;
    for (int b = 0; b < NNGP_MAX_BATCH; b++) *value += c->bf_short[b] == 2;
    return NNGP_OK;
  }
  return ctx_fail(c, NNGP_EINVAL, std::string("nngp_get_option: unknown option ") + name);
}

// Second half of the pair-distance quantile (see nngp_set_data): histogram -> [sum over ranks] -> 99 % quantile.
static int pair_quantile_finish(nngp_ctx *c) {
  if (!c->hist_pending) return NNGP_OK;
  c->hist_pending = false;
  const int nb = pair_d2_hist_bins();
  NNGP_CUDA(c, cudaEventSynchronize(c->ev_hist));
  const unsigned int *hist = c->h_hist;
  std::vector<double> cnt(hist, hist + nb);
  if (c->mbox_shared) {
    // only the occupied range travels: every rank posts its bounds, all take the union
    int lo = nb, hi = -1;
    for (int b = 0; b < nb; b++)
      if (hist[b]) {
        if (b < lo) lo = b;
        hi = b;
      }
    std::vector<double> bounds((size_t)2 * c->mbox_ranks, 0.0);
    bounds[2 * c->mbox_rank] = (double)lo;
    bounds[2 * c->mbox_rank + 1] = (double)hi;
    int rc = mbox_host_allreduce(c, bounds.data(), 2 * c->mbox_ranks);
    if (rc) return rc;
    for (int r = 0; r < c->mbox_ranks; r++) {
      if ((int)bounds[2 * r] < lo) lo = (int)bounds[2 * r];
      if ((int)bounds[2 * r + 1] > hi) hi = (int)bounds[2 * r + 1];
    }
    const int chunk = kMboxSlotDoubles - kMboxPayload;
    for (int b0 = lo; b0 <= hi; b0 += chunk) {
      const int len = (hi + 1 - b0 < chunk) ? hi + 1 - b0 : chunk;
      rc = mbox_host_allreduce(c, cnt.data() + b0, len);
      if (rc) return rc;
    }
  }
  double total = 0, above = 0;
  for (int b = 0; b < nb; b++) total += cnt[b];
  int b = nb - 1;
  for (; b > 0; b--) {
    if ((above + cnt[b]) * 100.0 > total) break;  // more than 1 % of the locations at or above this bin
    above += cnt[b];
  }
  const unsigned long long edge = (unsigned long long)(b + 1) << 48;  // upper edge of bin b
  double v;
  memcpy(&v, &edge, sizeof v);
  if (total > 0 && b + 1 < 0x7ff0) c->pair_d2_q99 = v;
  return NNGP_OK;
}

int nngp_set_data(nngp_ctx *c, int n, int m, int p, const double *coords, const int *nnIndx, const int *nnIndxLU,
                  const double *X, const double *y, int loc_begin, int loc_end) {
  if (!c) return NNGP_EINVAL;
  if (n < 1 || m < 1 || p < 0 || !coords) return ctx_fail(c, NNGP_EINVAL, "nngp_set_data: bad sizes or NULL input");
  if (m > NNGP_MAX_M) return ctx_fail(c, NNGP_EINVAL, "nngp_set_data: n.neighbors > 32 is not supported by the warp-group solver");
  if (p > 0 && (!X || !y)) return ctx_fail(c, NNGP_EINVAL, "nngp_set_data: X / y missing");
  if (loc_begin < 0 || loc_end > n || loc_begin > loc_end) return ctx_fail(c, NNGP_EINVAL, "nngp_set_data: bad shard range");
  PhaseTrace tr("nngp_set_data");
  // nnIndxLU must be the closed form of src/util.cpp:55-70 (offset = running sum of the counts min(i, m)); the pass over
  // its 8 n bytes runs on its own thread next to the staged upload and is collected behind it
  std::future<int> lu_bad;
  if (nnIndxLU) {
    lu_bad = std::async(std::launch::async, [=] {
      long long run = 0;
      int bad = 0;
      for (int i = 1; i < n; i++) {
        run += (i - 1 < m) ? i - 1 : m;
        bad |= (nnIndxLU[i] != (int)run) | (nnIndxLU[n + i] != ((i < m) ? i : m));
      }
      return bad;
    });
  }
  NNGP_CUDA(c, cudaSetDevice(c->device));
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  tr.mark("validate + sync");
  // device buffers are kept across calls and only grown (cudaMalloc/cudaFree of the ~0.5 GB working set
  // would otherwise dominate a short fit: the whole-call path is measured end to end by bench.py)
  c->n = n; c->m = m; c->p = p; c->loc_begin = loc_begin; c->loc_end = loc_end;
  c->rep_ready = false;
  c->tri_on = false;
  c->has_nn = (nnIndx != nullptr);  // NULL: prediction-only use (sites, X, y), no B/F storage
  c->nIndx = c->has_nn ? nn_offset(n, m) : 0;  // == (1+m)/2*m + (n-m-1)*m for n > m
  // this shard's slice of the CSR arrays; the virtual bases (nngp_ctx.h) make global offsets work on the slice
  c->csr_begin = c->has_nn ? nn_offset(loc_begin, m) : 0;
  c->csr_end = c->has_nn ? nn_offset(loc_end, m) : 0;
  const long long nI = c->csr_end > c->csr_begin ? c->csr_end - c->csr_begin : 1;
  const long long nF = loc_end > loc_begin ? loc_end - loc_begin : 1;
  // B: complete rows of a 32-location tile must start 16-byte aligned (bulk copies): one element of padding when the
  // CSR offset of the first such tile, relative to the shard, is odd (all later tiles follow in steps of 32 m)
  const long long first_full = loc_begin < m ? loc_begin + ((m - loc_begin + 31) / 32) * 32 : loc_begin;
  const long long b_pad = c->has_nn ? ((nn_offset(first_full, m) - c->csr_begin) & 1) : 0;
  int rc;
#define NNGP_ENSURE(ptr, slot, count)                                     \
  do {                                                                    \
    rc = ensure_dev(c, ptr, c->cap[slot], (size_t)(count));               \
    if (rc) return rc;                                                    \
  } while (0)
  NNGP_ENSURE(c->d_coords, 0, n);
  NNGP_ENSURE(c->d_nnIndx_store, 1, nI);
  c->d_nnIndx = c->d_nnIndx_store - c->csr_begin;
  if (p > 0) {
    NNGP_ENSURE(c->d_X, 2, (size_t)n * p);
    NNGP_ENSURE(c->d_y, 3, n);
  }
  {
    // one threaded pass through pinned chunks (host_stage.h): coords interleaved to (x, y) pairs, neighbour ids
    // range-checked on the way
    if (!c->stager) {
      c->stager = new HostStager();
      NNGP_CUDA(c, c->stager->init(c->device));
    }
    // The REPLICATED arrays (sites, X, y) of a sharded fit: every rank uploads 1 / N of them from its host copy and
    // the ranks exchange the pieces over NVLink (grouped ncclBroadcast, one root per piece) -- host-to-device traffic per
    // rank drops N-fold, which is what made nngp_set_data slower with every added GPU (8 ranks x 40 MB through one host).
    const char *gmin = getenv("NNGP_GATHER_MIN");  // sites from which the exchange pays (tests lower it)
    const bool gather = c->comm != nullptr && c->comm_size > 1 && n >= (gmin ? atoi(gmin) : (1 << 16)) && !getenv("NNGP_NO_GATHER");
    const long long s0 = gather ? (long long)n * c->comm_rank / c->comm_size : 0;
    const long long s1 = gather ? (long long)n * (c->comm_rank + 1) / c->comm_size : n;
    std::vector<StageJob> jobs;
    // (XY_INTERLEAVE reads y at src + n doubles: a shifted src addresses the same pair of columns from site s0 on)
    jobs.push_back({StageJob::XY_INTERLEAVE, c->d_coords + s0, coords + s0, (size_t)(s1 - s0) * sizeof(double2), n});
    if (c->csr_end > c->csr_begin)
      jobs.push_back({StageJob::INT_RANGE, c->d_nnIndx_store, nnIndx + c->csr_begin,
                      (size_t)(c->csr_end - c->csr_begin) * sizeof(int), n});
    if (p > 0) {
      for (int col = 0; col < p; col++)
        jobs.push_back({StageJob::RAW, c->d_X + (size_t)col * n + s0, X + (size_t)col * n + s0, (size_t)(s1 - s0) * sizeof(double), 0});
      jobs.push_back({StageJob::RAW, c->d_y + s0, y + s0, (size_t)(s1 - s0) * sizeof(double), 0});
    }
    bool bad_id = false;
    int min_id = INT_MAX, max_id = -1;
    tr.mark("buffers");
    NNGP_CUDA(c, c->stager->run(jobs, &bad_id, &min_id, &max_id));
    // residuals are read at the shard's own locations and at their neighbours (in the reference's ordered lists: earlier
    // locations, so the range ends at loc_end; taken from the ids themselves, not assumed)
    c->resid_begin = (min_id >= 0 && min_id < loc_begin) ? min_id : loc_begin;
    c->resid_end = (max_id >= loc_end && max_id < n) ? max_id + 1 : loc_end;
    if (!c->has_nn || c->csr_end <= c->csr_begin) {
      c->resid_begin = 0;
      c->resid_end = n;
    }
    tr.mark("staged upload");
    if (gather) {
      rc = comm_exchange_replicated(c, c->d_coords, c->d_X, c->d_y, n, p);
      if (rc) return rc;
      tr.mark("exchange of the replicated arrays");
    }
    if (bad_id) {
      c->n = 0;
      return ctx_fail(c, NNGP_EINVAL, "nngp_set_data: neighbour index out of range");
    }
    if (lu_bad.valid() && lu_bad.get()) {
      c->n = 0;
      return ctx_fail(c, NNGP_EINVAL, "nngp_set_data: nnIndxLU is not the reference's closed-form layout");
    }
  }
  if (p > 0) {
    NNGP_ENSURE(c->d_beta, 4, (size_t)NNGP_MAX_BATCH * p);
    if (c->h_beta_cap < (size_t)NNGP_MAX_BATCH * p) {
      if (c->h_beta) cudaFreeHost(c->h_beta);
      c->h_beta = nullptr;
      NNGP_CUDA(c, cudaHostAlloc(&c->h_beta, (size_t)NNGP_MAX_BATCH * p * sizeof(double), cudaHostAllocDefault));
      c->h_beta_cap = (size_t)NNGP_MAX_BATCH * p;
    }
    for (int b = 0; b < NNGP_MAX_BATCH; b++) NNGP_ENSURE(c->d_resid[b], 5 + b, n);
    NNGP_ENSURE(c->d_Xy, 15, xy_rows_count(n, p));
    NNGP_CUDA(c, xy_rows_launch(c->d_X, c->d_y, n, p, c->d_Xy, c->sms, c->stream));
  }
  for (int s = 0; s < 2 && c->has_nn; s++) {
    NNGP_ENSURE(c->d_B_store[s], 9 + s, nI + 2);
    NNGP_ENSURE(c->d_F_store[s], 11 + s, nF);
    c->d_B[s] = c->d_B_store[s] + b_pad - c->csr_begin;
    c->d_F[s] = c->d_F_store[s] - loc_begin;
  }
  if (c->has_nn && c->opt_tri && tri_capacity_for_m(m) > 0 && loc_end > loc_begin) {
    // resident distance triangle (tri_cache.cu): m(m+1)/2 doubles per location, built once per data set
    const size_t need = tri_cache_doubles(m, (long long)loc_end - loc_begin);
    size_t free_b = 0, total_b = 0;
    cudaMemGetInfo(&free_b, &total_b);
    const bool fits = c->tri_cap >= need * sizeof(double) || need * sizeof(double) + ((size_t)2 << 30) < free_b + c->tri_cap;
    if (fits) {
      rc = ensure_dev(c, c->d_tri, c->tri_cap, need);
      if (rc == NNGP_OK) {
        NNGP_CUDA(c, tri_build_launch(c->d_coords, c->d_nnIndx, m, loc_begin, loc_end, c->d_tri, c->stream));
        c->launches++;
        c->tri_on = true;
      } else if (rc != NNGP_ENOMEM) {
        return rc;
      }
    }
    tr.mark("distance triangle");
  }
  c->pair_d2_q99 = 1e300;
  c->hist_pending = false;
  const bool have_pairs = c->has_nn && (long long)loc_end > m && loc_end > loc_begin;
  if (have_pairs || (c->has_nn && c->mbox_shared)) {
    // 99 % quantile of the largest squared pair distance inside a neighbourhood (stat_kernels.cu): only the Matern
    // series choice needs it, so an unsharded context launches the histogram here and reads it when the first Matern
    // likelihood is prepared (pair_quantile_finish).  In a sharded fit the histograms of all ranks are added first
    // (through the mailbox, a few rounds of < 2,300 counts), so that every rank -- and every shard layout -- takes the
    // same decision as the unsharded fit; that exchange is collective and stays here.
    const int nb = pair_d2_hist_bins();
    if (!c->h_hist) NNGP_CUDA(c, cudaHostAlloc(&c->h_hist, nb * sizeof(unsigned int), cudaHostAllocDefault));
    if (!c->ev_hist) NNGP_CUDA(c, cudaEventCreateWithFlags(&c->ev_hist, cudaEventDisableTiming));
    memset(c->h_hist, 0, nb * sizeof(unsigned int));
    if (have_pairs) {
      if (!c->d_hist) NNGP_CUDA(c, cudaMalloc(&c->d_hist, nb * sizeof(unsigned int)));
      NNGP_CUDA(c, pair_d2_hist_launch(c->d_coords, c->d_nnIndx, m, loc_begin, loc_end, c->d_hist, c->stream));
      c->launches++;
      NNGP_CUDA(c, cudaMemcpyAsync(c->h_hist, c->d_hist, nb * sizeof(unsigned int), cudaMemcpyDeviceToHost, c->stream));
      NNGP_CUDA(c, cudaEventRecord(c->ev_hist, c->stream));
    }
    c->hist_pending = true;
    if (c->mbox_shared) {
      rc = pair_quantile_finish(c);
      if (rc) return rc;
    }
    tr.mark("row layout + pair-distance histogram");
  }
  const long long nloc = (long long)loc_end - loc_begin;
  const int K = p + 1, npairs = K * (K + 1) / 2;
  // per parameter set: 2 doubles per CTA or 4 per 32-location tile, whichever is more (site-mode chunks are <= 2^20)
  const long long tiles_max = ((nloc > (1 << 20) ? nloc : (1 << 20)) + 31) / 32;
  const long long per_set = 4 * (tiles_max > bf_grid_max(m, (long long)1 << 40, c->sms) ? tiles_max : bf_grid_max(m, (long long)1 << 40, c->sms));
  c->partials_stride = (size_t)per_set;
  size_t need = (size_t)NNGP_MAX_BATCH * per_set;
  size_t need2 = (size_t)q_forms_grid(nloc) * npairs;
  size_t need3 = (size_t)q_resid_grid(m, nloc, c->sms);
  c->partials_len = need > need2 ? need : need2;
  if (need3 > c->partials_len) c->partials_len = need3;
  NNGP_ENSURE(c->d_partials, 13, c->partials_len);
  c->out_len = (size_t)npairs + 8;
  NNGP_ENSURE(c->d_out, 14, c->out_len);
#undef NNGP_ENSURE
  return NNGP_OK;
}

// Build the launch descriptor for `nbatch` parameter sets.  resid_on: form r = X beta - y first.
static int prepare_bf(nngp_ctx *c, int cov, int nbatch, const double *theta, const double *beta, const int *slots,
                      bool store, BfLaunch *P, int *grid_out, int cur_slot = -1) {
  if (c->n == 0 || !c->has_nn) return ctx_fail(c, NNGP_ESTATE, "likelihood requested before nngp_set_data (with neighbours)");
  if (nbatch < 1 || nbatch > NNGP_MAX_BATCH) return ctx_fail(c, NNGP_EINVAL, "nbatch must be in 1..4");
  if (cov < 0 || cov > 3) return ctx_fail(c, NNGP_EINVAL, "cov.model is not correctly specified");
  if (beta && c->p == 0) return ctx_fail(c, NNGP_EINVAL, "beta given but the model has no covariates");
  const bool tri = c->tri_on && c->d_tri != nullptr;
  const int grid = tri ? tri_grid_for((long long)c->loc_end - c->loc_begin, c->sms)
                       : bf_grid_any(c->m, cov, (long long)c->loc_end - c->loc_begin, c->sms);
  memset(P, 0, sizeof *P);
  P->tri = tri ? c->d_tri : nullptr;
  P->nnT = tri ? reinterpret_cast<const int *>(c->d_tri + tri_ids_offset(c->m, (long long)c->loc_end - c->loc_begin)) : nullptr;
  memset(c->bf_short, 0, sizeof c->bf_short);
  P->coords = c->d_coords;
  P->nnIndx = c->d_nnIndx;
  P->n = c->n;
  P->m = c->m;
  P->loc_begin = c->loc_begin;
  P->loc_end = c->loc_end;
  P->nbatch = nbatch;
  if (!c->h_fail_dev) NNGP_CUDA(c, cudaHostGetDevicePointer((void **)&c->h_fail_dev, c->h_fail, 0));
  int *fail_dev = c->h_fail_dev;
  // results go to this rank's mailbox slot, stamped with a fresh sequence number (NCCL path: to d_red, reduced there)
  const unsigned long long seq = ++c->seq;
  double *out_dev = mbox_slot_dev(c, seq) + kMboxPayload;
  for (int b = 0; b < nbatch; b++) {
    const double *t = theta + 4 * b;
    if (!(t[0] > 0.0) || !(t[1] >= 0.0) || !(t[2] > 0.0)) return ctx_fail(c, NNGP_EINVAL, "theta out of range");
    BfTheta &th = P->th[b];
    th.sigmaSq = t[0];
    th.tauSq = t[1];
    th.phi = t[2];
    th.mat = nullptr;
    if (cov == NNGP_COV_MATERN) {
      if (!(t[3] > 0.0)) return ctx_fail(c, NNGP_EINVAL, "matern nu must be > 0");
      if (c->hist_pending) {
        int rcq = pair_quantile_finish(c);
        if (rcq) return rcq;
      }
      int hit = -1;
      for (int k = 0; k < 4; k++)
        if (c->mat_cache[k].nu == t[3]) hit = k;
      if (hit < 0) {  // 10 us of host time (the Gauss-Laguerre rule): skipped for a smoothness seen recently
        hit = c->mat_cache_next;
        c->mat_cache_next = (c->mat_cache_next + 1) & 3;
        matern_prepare(t[3], &c->mat_cache[hit].mc);
        c->mat_cache[hit].nu = t[3];
      }
      c->h_mat[b] = c->mat_cache[hit].mc;
      matern_series_params(c->h_mat[b], &th.ms);
      th.mat = c->d_mat + b;
      // short series: 99 % of the locations keep t = phi^2 d^2 / 4 of every pair inside its range
      const double t99 = 0.25 * t[2] * t[2] * c->pair_d2_q99;
      c->bf_short[b] = (c->opt_matern_short && th.ms.use_series && (tri || bf_has_matern_short(c->m)) &&
                        t99 <= c->h_mat[b].t_short_max) ? 1 : 0;
      if (c->bf_short[b] && c->opt_matern_short >= 2 && tri && t99 <= c->h_mat[b].t_tiny_max) c->bf_short[b] = 2;
      if (c->bf_short[b]) th.ms.t_fast_max = (c->bf_short[b] == 2) ? c->h_mat[b].t_tiny_max : c->h_mat[b].t_short_max;
      matern_set_phi(&th.ms, t[2]);
    }
    const int slot = slots ? slots[b] : (b & 1);
    if (slot < 0 || slot > 1) return ctx_fail(c, NNGP_EINVAL, "slot must be 0 or 1");
    th.B = store ? c->d_B[slot] : nullptr;
    th.F = store ? c->d_F[slot] : nullptr;
    th.resid = beta ? c->d_resid[b] : nullptr;
    if (cur_slot >= 0 && tri && beta) {  // also return Q of the resident factors of `cur_slot` (bf_tri_kernel only)
      th.Bcur = c->d_B[cur_slot];
      th.Fcur = c->d_F[cur_slot];
    }
    th.out = (mbox_reduces(c) ? out_dev : c->d_red) + kBfOutStride * b;
    th.seq = (double)seq;
    th.partials = c->d_partials + c->partials_stride * b;
    th.ticket = c->d_ticket + b;
    th.tile_counter = c->d_ticket + NNGP_MAX_BATCH + b;
    th.fail = c->d_fail + b;
    th.fail_out = fail_dev + b;
  }
  if (cov == NNGP_COV_MATERN) {  // full tables for the rare Temme / CF2 path (the series coefficients travel as parameters)
    for (int b = 0; b < nbatch; b++) {
      if (c->d_mat_nu[b] == theta[4 * b + 3]) continue;  // this slot already holds the tables of that smoothness
      NNGP_CUDA(c, cudaMemcpyAsync(c->d_mat + b, c->h_mat + b, sizeof(MaternConsts), cudaMemcpyHostToDevice, c->stream));
      c->d_mat_nu[b] = theta[4 * b + 3];
    }
  }
  if (beta) {
    memcpy(c->h_beta, beta, (size_t)nbatch * c->p * sizeof(double));
    if (c->p > kResidBetaMax)  // otherwise beta rides in the residual kernel's parameters
      NNGP_CUDA(c, cudaMemcpyAsync(c->d_beta, c->h_beta, (size_t)nbatch * c->p * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  }
  *grid_out = grid;
  return NNGP_OK;
}

static int enqueue_bf(nngp_ctx *c, int cov, const BfLaunch &P0, int grid, bool with_resid) {
  BfLaunch P = P0;
  bool resid_launched = false;
  if (with_resid) {
    for (int b = 0; b < P.nbatch; b++) {
      // identical betas share one residual vector
      bool dup = false;
      for (int a = 0; a < b && !dup; a++)
        if (memcmp(c->h_beta + (size_t)a * c->p, c->h_beta + (size_t)b * c->p, c->p * sizeof(double)) == 0) dup = true;
      if (dup) continue;
      if (c->p <= kResidBetaMax)
        NNGP_CUDA(c, resid_launch_val(c->d_X, c->d_y, c->h_beta + (size_t)b * c->p, c->n, c->p, c->d_resid[b], c->resid_begin, c->resid_end, c->sms, c->stream));
      else
        NNGP_CUDA(c, resid_launch(c->d_X, c->d_y, c->d_beta + (size_t)b * c->p, c->n, c->p, c->d_resid[b], c->resid_begin, c->resid_end, c->sms, c->stream));
      c->launches++;
      resid_launched = true;
    }
  }
  // NNGP_PDL=1: the first B/F launch behind a residual kernel starts beside it (programmatic dependent launch).  Opt-in:
  // measured per sampler iteration on one GPU (tools/iter_probe.py), it buys 2.5 % on an inner 125k-location shard (Matern
  // 139.8 -> 136.4 us, exponential 106.9 -> 104.2), nothing at n = 10^6, and COSTS 10 % on the Matern head shard
  // [0, 125000) (151.9 -> 168.2 us) -- the rank that paces a sharded fit.
  static const bool use_pdl = getenv("NNGP_PDL") != nullptr;
  P.pdl = (resid_launched && use_pdl) ? 1 : 0;
  // Matern: one launch per series variant present in the batch (0 full, 1 short, 2 tiny), so that the result of a
  // parameter set never depends on what it was batched with (the sampler's current and candidate phi may straddle a
  // variant's range); a homogeneous batch -- the usual case -- is one launch
  int have[3] = {0, 0, 0};
  for (int b = 0; b < P.nbatch; b++) have[cov == NNGP_COV_MATERN ? c->bf_short[b] : 0]++;
  for (int v = 0; v < 3; v++) {
    if (!have[v]) continue;
    if (have[v] == P.nbatch) {
      if (v == 0) {
        NNGP_CUDA(c, bf_launch(P, cov, grid, c->stream));
      } else {
        BfLaunch Ps = P;
        Ps.matern_short = v;
        NNGP_CUDA(c, bf_launch(Ps, cov, grid, c->stream));
      }
      c->launches++;
      return NNGP_OK;
    }
    BfLaunch Pv = P;
    Pv.nbatch = 0;
    for (int b = 0; b < P.nbatch; b++)
      if (c->bf_short[b] == v) Pv.th[Pv.nbatch++] = P.th[b];
    Pv.matern_short = v;
    NNGP_CUDA(c, bf_launch(Pv, cov, grid, c->stream));
    c->launches++;
    P.pdl = 0;  // only the launch right behind the residual kernel
  }
  return NNGP_OK;
}

// after prepare_bf: point duplicate-beta sets at the first set's residual vector
static void share_resid(nngp_ctx *c, BfLaunch *P) {
  for (int b = 1; b < P->nbatch; b++)
    for (int a = 0; a < b; a++)
      if (P->th[b].resid && memcmp(c->h_beta + (size_t)a * c->p, c->h_beta + (size_t)b * c->p, c->p * sizeof(double)) == 0) {
        P->th[b].resid = P->th[a].resid;
        break;
      }
}

// after the launch: sum over ranks if sharded, one stream sync, unpack (logDet, Q) per set, report failures
static double wall_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

static int finish_bf(nngp_ctx *c, int nbatch, bool have_q, double *out) {
  const double t_wait0 = wall_s();
  struct WaitClock {
    nngp_ctx *c;
    double t0;
    ~WaitClock() { c->trace_wait_s += wall_s() - t0; }
  } wait_clock{c, t_wait0};
  double sums[kBfOutStride * NNGP_MAX_BATCH];
  const double *res = sums;
  if (mbox_reduces(c)) {  // poll the stamps of every rank's result blocks, add them in rank order
    int rc = mbox_wait_sets(c, c->seq, nbatch, sums);
    if (rc) return rc;
  } else {  // NCCL: sum (logDet, Q, fail flag) over the ranks on the device, then one small copy back
    int rc = comm_allreduce_dev(c, c->d_red, kBfOutStride * nbatch);
    if (rc) return rc;
    NNGP_CUDA(c, cudaMemcpyAsync(c->h_red, c->d_red, kBfOutStride * nbatch * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
    res = c->h_red;
  }
  int status = 0;
  bool any_rank_failed = false;
  for (int b = 0; b < nbatch; b++) {
    out[2 * b] = res[kBfOutStride * b];
    out[2 * b + 1] = have_q ? res[kBfOutStride * b + 1] : 0.0;
    c->last_qcur[b] = res[kBfOutStride * b + 3];
    if (res[kBfOutStride * b + 2] > 0.0) any_rank_failed = true;
    if (c->h_fail[b] != INT_MAX && (status == 0 || c->h_fail[b] < status)) status = c->h_fail[b];
  }
  if (status == 0 && any_rank_failed) status = INT_MAX;  // the failing location lives on another rank
  if (status > 0) return ctx_fail(c, status, "c++ error: dpotrf failed (neighbour covariance not positive definite)");
  return NNGP_OK;
}

// `len` doubles at d_src (this rank's part) -> their sum over the ranks, on the host
static int reduce_small(nngp_ctx *c, double *d_src, int len, double *host) {
  if (mbox_reduces(c) && len <= kMboxSlotDoubles - kMboxPayload) {
    const unsigned long long seq = ++c->seq;
    int rc = mbox_post_dev(c, d_src, len, seq);
    if (rc) return rc;
    return mbox_wait_payload(c, seq, len, host);
  }
  if (c->comm) {
    int rc = comm_allreduce_dev(c, d_src, len);
    if (rc) return rc;
  }
  NNGP_CUDA(c, cudaMemcpyAsync(host, d_src, len * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  return NNGP_OK;
}

int nngp_bf_loglik(nngp_ctx *c, int cov, int nbatch, const double *theta, const double *beta, const int *slots,
                   double *out) {
  if (!c || !theta || !out) return NNGP_EINVAL;
  const double t_enq0 = wall_s();
  NNGP_CUDA(c, cudaSetDevice(c->device));
  BfLaunch P;
  int grid = 0;
  int rc = prepare_bf(c, cov, nbatch, theta, beta, slots, true, &P, &grid);
  if (rc) return rc;
  share_resid(c, &P);
  rc = enqueue_bf(c, cov, P, grid, beta != nullptr);
  if (rc) return rc;
  c->trace_enqueue_s += wall_s() - t_enq0;
  return finish_bf(c, nbatch, beta != nullptr, out);
}

// make sure the Matern tables of smoothness nu sit in the context's small cache (what prepare_bf looks up)
static void matern_cache_warm(nngp_ctx *c, double nu) {
  if (!(nu > 0.0)) return;
  for (int k = 0; k < 4; k++)
    if (c->mat_cache[k].nu == nu) return;
  const int hit = c->mat_cache_next;
  c->mat_cache_next = (c->mat_cache_next + 1) & 3;
  matern_prepare(nu, &c->mat_cache[hit].mc);
  c->mat_cache[hit].nu = nu;
}

// `between` (may be empty) runs on the host after the launch has been queued and before its result is awaited
static int bf_loglik_cur_impl(nngp_ctx *c, int cov, const double *theta, const double *beta, int slot_new, int slot_cur,
                              double *out, const std::function<void()> &between) {
  if (!c || !theta || !beta || !out || slot_new < 0 || slot_new > 1 || slot_cur < 0 || slot_cur > 1 || slot_new == slot_cur)
    return NNGP_EINVAL;
  const double t_enq0 = wall_s();
  NNGP_CUDA(c, cudaSetDevice(c->device));
  BfLaunch P;
  int grid = 0;
  int rc = prepare_bf(c, cov, 1, theta, beta, &slot_new, true, &P, &grid, slot_cur);
  if (rc) return rc;
  rc = enqueue_bf(c, cov, P, grid, true);
  if (rc) return rc;
  c->trace_enqueue_s += wall_s() - t_enq0;
  if (between) between();
  double o2[2];
  rc = finish_bf(c, 1, true, o2);
  if (rc) return rc;
  out[0] = o2[0];
  out[1] = o2[1];
  if (P.th[0].Bcur != nullptr) {  // the launch carried the resident factors along
    out[2] = c->last_qcur[0];
    return NNGP_OK;
  }
  return nngp_q_resid(c, slot_cur, beta, out + 2);  // kernels without that path: one more pass over B, F
}

int nngp_bf_loglik_cur(nngp_ctx *c, int cov, const double *theta, const double *beta, int slot_new, int slot_cur,
                       double *out) {
  return bf_loglik_cur_impl(c, cov, theta, beta, slot_new, slot_cur, out, nullptr);
}

int nngp_bf_loglik_vec(nngp_ctx *c, int cov, int nbatch, const double *theta, const double *u, const int *slots,
                       double *out) {
  if (!c || !theta || !u || !out) return NNGP_EINVAL;
  NNGP_CUDA(c, cudaSetDevice(c->device));
  if (c->n == 0 || !c->has_nn) return ctx_fail(c, NNGP_ESTATE, "likelihood requested before nngp_set_data (with neighbours)");
  int rc = ensure_dev(c, c->d_resid[0], c->cap[5], (size_t)c->n);
  if (rc) return rc;
  if (!c->stager) {
    c->stager = new HostStager();
    NNGP_CUDA(c, c->stager->init(c->device));
  }
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  std::vector<StageJob> up{{StageJob::RAW, c->d_resid[0], u, (size_t)c->n * sizeof(double), 0}};
  NNGP_CUDA(c, c->stager->run(up, nullptr));
  BfLaunch P;
  int grid = 0;
  rc = prepare_bf(c, cov, nbatch, theta, nullptr, slots, true, &P, &grid);
  if (rc) return rc;
  for (int b = 0; b < nbatch; b++) P.th[b].resid = c->d_resid[0];  // every set whitens the same vector
  rc = enqueue_bf(c, cov, P, grid, false);
  if (rc) return rc;
  return finish_bf(c, nbatch, true, out);
}

int nngp_time_bf_loglik(nngp_ctx *c, int cov, int nbatch, const double *theta, const double *beta, int iters,
                        int flush_l2, float *ms_per_launch) {
  if (!c || !theta || !ms_per_launch || iters < 1) return NNGP_EINVAL;
  NNGP_CUDA(c, cudaSetDevice(c->device));
  BfLaunch P;
  int grid = 0;
  int rc = prepare_bf(c, cov, nbatch, theta, beta, nullptr, true, &P, &grid);
  if (rc) return rc;
  share_resid(c, &P);
  if (flush_l2 && !c->d_flush) {
    c->flush_bytes = (size_t)256 << 20;  // > 126 MB L2
    NNGP_CUDA(c, cudaMalloc(&c->d_flush, c->flush_bytes));
  }
  double total = 0.0;
  if (!flush_l2) {
    NNGP_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    for (int it = 0; it < iters; it++) {
      rc = enqueue_bf(c, cov, P, grid, beta != nullptr);
      if (rc) return rc;
    }
    NNGP_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    NNGP_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    total = ms;
  } else {
    for (int it = 0; it < iters; it++) {
      NNGP_CUDA(c, cudaMemsetAsync(c->d_flush, it & 0xff, c->flush_bytes, c->stream));
      NNGP_CUDA(c, cudaEventRecord(c->ev0, c->stream));
      rc = enqueue_bf(c, cov, P, grid, beta != nullptr);
      if (rc) return rc;
      NNGP_CUDA(c, cudaEventRecord(c->ev1, c->stream));
      NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
      float ms = 0.f;
      NNGP_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
      total += ms;
    }
  }
  *ms_per_launch = (float)(total / iters);
  return NNGP_OK;
}

int nngp_q_resid(nngp_ctx *c, int slot, const double *beta, double *out) {
  if (!c || !beta || !out || slot < 0 || slot > 1) return NNGP_EINVAL;
  if (c->n == 0 || c->p == 0 || !c->has_nn) return ctx_fail(c, NNGP_ESTATE, "nngp_q_resid before nngp_set_data");
  NNGP_CUDA(c, cudaSetDevice(c->device));
  memcpy(c->h_beta, beta, c->p * sizeof(double));
  if (c->p <= kResidBetaMax) {
    NNGP_CUDA(c, resid_launch_val(c->d_X, c->d_y, c->h_beta, c->n, c->p, c->d_resid[0], c->resid_begin, c->resid_end, c->sms, c->stream));
  } else {
    NNGP_CUDA(c, cudaMemcpyAsync(c->d_beta, c->h_beta, c->p * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    NNGP_CUDA(c, resid_launch(c->d_X, c->d_y, c->d_beta, c->n, c->p, c->d_resid[0], c->resid_begin, c->resid_end, c->sms, c->stream));
  }
  NNGP_CUDA(c, q_resid_launch(c->d_B[slot], c->d_F[slot], c->d_nnIndx, c->d_resid[0], c->m, c->loc_begin, c->loc_end,
                              c->d_partials, c->d_out, c->sms, c->stream));
  c->launches += 3;
  return reduce_small(c, c->d_out, 1, out);
}

int nngp_q_forms(nngp_ctx *c, int slot, double *S) {
  if (!c || !S || slot < 0 || slot > 1) return NNGP_EINVAL;
  if (c->n == 0 || c->p == 0 || !c->has_nn) return ctx_fail(c, NNGP_ESTATE, "nngp_q_forms before nngp_set_data");
  NNGP_CUDA(c, cudaSetDevice(c->device));
  const int K = c->p + 1, npairs = K * (K + 1) / 2;
  NNGP_CUDA(c, q_forms_launch(c->d_B[slot], c->d_F[slot], c->d_nnIndx, c->d_Xy, c->n, c->p, c->m, c->loc_begin,
                              c->loc_end, c->d_partials, c->d_out, c->stream));
  c->launches += 2;
  std::vector<double> packed(npairs);
  int rc = reduce_small(c, c->d_out, npairs, packed.data());
  if (rc) return rc;
  int e = 0;
  for (int d = 0; d < K; d++)
    for (int r = d; r < K; r++, e++) {
      S[(size_t)d * K + r] = packed[e];
      S[(size_t)r * K + d] = packed[e];
    }
  return NNGP_OK;
}

int nngp_get_bf(nngp_ctx *c, int slot, double *B, double *F) {
  if (!c || slot < 0 || slot > 1) return NNGP_EINVAL;
  if (c->n == 0 || !c->has_nn) return ctx_fail(c, NNGP_ESTATE, "nngp_get_bf before nngp_set_data");
  NNGP_CUDA(c, cudaSetDevice(c->device));
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  // storage is sharded: this context holds [csr_begin, csr_end) of B and [loc_begin, loc_end) of F; the rest of the
  // caller's arrays is left untouched
  if (B && c->csr_end > c->csr_begin)
    NNGP_CUDA(c, cudaMemcpy(B + c->csr_begin, c->d_B[slot] + c->csr_begin, (size_t)(c->csr_end - c->csr_begin) * sizeof(double),
                            cudaMemcpyDeviceToHost));
  if (F && c->loc_end > c->loc_begin)
    NNGP_CUDA(c, cudaMemcpy(F + c->loc_begin, c->d_F[slot] + c->loc_begin, (size_t)(c->loc_end - c->loc_begin) * sizeof(double),
                            cudaMemcpyDeviceToHost));
  return NNGP_OK;
}

// ---------------------------------------------------------------------------------------------------
// `rNNGP`: the response-model sampler (reference src/rNNGP.cpp:255-378), host loop + device hot path.
// ---------------------------------------------------------------------------------------------------

static double logit_(double theta, double a, double b) { return log((theta - a) / (b - theta)); }  // src/util.cpp:43
static double logit_inv_(double z, double a, double b) { return b - (b - a) / (1 + exp(z)); }      // src/util.cpp:47

// src/rNNGP.cpp:317-324
static double log_post_(double logDet, double Q, const double *th, const nngp_rnngp_args *a) {
  double lp = -0.5 * logDet - 0.5 * Q;
  lp += log(th[2] - a->phiUnif[0]) + log(a->phiUnif[1] - th[2]);
  lp += -1.0 * (1.0 + a->sigmaSqIG[0]) * log(th[0]) - a->sigmaSqIG[1] / th[0] + log(th[0]);
  lp += -1.0 * (1.0 + a->tauSqIG[0]) * log(th[1]) - a->tauSqIG[1] / th[1] + log(th[1]);
  if (a->cov_model == NNGP_COV_MATERN) lp += log(th[3] - a->nuUnif[0]) + log(a->nuUnif[1] - th[3]);
  return lp;
}

int nngp_rnngp(nngp_ctx *c, const nngp_rnngp_args *a, double *betaSamples, double *thetaSamples, int *accept_out) {
  if (!c || !a || !betaSamples || !thetaSamples) return NNGP_EINVAL;
  if (c->n == 0 || c->p == 0) return ctx_fail(c, NNGP_ESTATE, "nngp_rnngp before nngp_set_data");
  if (!a->rnorm || !a->runif || !a->betaStarting) return ctx_fail(c, NNGP_EINVAL, "nngp_rnngp: RNG hooks / starting values missing");
  const int p = c->p, K = p + 1, cov = a->cov_model;
  const int nTheta = (cov == NNGP_COV_MATERN) ? 4 : 3;
  char line[256];
  if (a->verbose && a->print) {  // banner text of src/rNNGP.cpp:157-179
    static const char *names[4] = {"exponential", "spherical", "matern", "gaussian"};
    a->print("----------------------------------------\n\tModel description\n----------------------------------------\n");
    snprintf(line, sizeof line, "NNGP Response model fit with %i observations.\n\n", c->n); a->print(line);
    snprintf(line, sizeof line, "Number of covariates %i (including intercept if specified).\n\n", p); a->print(line);
    snprintf(line, sizeof line, "Using the %s spatial correlation model.\n\n", names[cov]); a->print(line);
    snprintf(line, sizeof line, "Using %i nearest neighbors.\n\n", c->m); a->print(line);
    snprintf(line, sizeof line, "Number of MCMC samples %i.\n\n", a->nSamples); a->print(line);
    a->print("Priors and hyperpriors:\n\tbeta flat.\n");
    snprintf(line, sizeof line, "\tsigma.sq IG hyperpriors shape=%.5f and scale=%.5f\n", a->sigmaSqIG[0], a->sigmaSqIG[1]); a->print(line);
    snprintf(line, sizeof line, "\ttau.sq IG hyperpriors shape=%.5f and scale=%.5f\n", a->tauSqIG[0], a->tauSqIG[1]); a->print(line);
    snprintf(line, sizeof line, "\tphi Unif hyperpriors a=%.5f and b=%.5f\n", a->phiUnif[0], a->phiUnif[1]); a->print(line);
    if (cov == NNGP_COV_MATERN) {
      snprintf(line, sizeof line, "\tnu Unif hyperpriors a=%.5f and b=%.5f\n", a->nuUnif[0], a->nuUnif[1]); a->print(line);
    }
    snprintf(line, sizeof line, "\nSource compiled for sm_100a; model fit on CUDA device %i (%s sampler).\n", c->device,
             a->mode ? "double-buffered" : "reference-faithful");
    a->print(line);
    a->print("----------------------------------------\n\t\tSampling\n----------------------------------------\n");
  }

  double theta[4] = {a->sigmaSqStarting, a->tauSqStarting, a->phiStarting, (cov == NNGP_COV_MATERN) ? a->nuStarting : 0.0};
  double tuning[4] = {a->sigmaSqTuning, a->tauSqTuning, a->phiTuning, a->nuTuning};
  double cand[4] = {0, 0, 0, 0};
  std::vector<double> beta(a->betaStarting, a->betaStarting + p), S((size_t)K * K), tmp_pp((size_t)p * p), tmp_p(p),
      mean(p);
  int cur = 0;  // slot holding B,F of the current theta
  double out[2 * NNGP_MAX_BATCH], red[4];
  int slots[2];
  int rc;

  // initial B,F at the starting values (src/rNNGP.cpp:255)
  slots[0] = cur;
  rc = nngp_bf_loglik(c, cov, 1, theta, nullptr, slots, out);
  if (rc) return rc;
  double logDetCur = out[0];
  if (a->allreduce) a->allreduce(&logDetCur, 1, a->allreduce_user);

  // NNGP_TRACE_ITER=1: where one iteration's wall time goes (host clock; printed once per call on stderr)
  const bool trace_iter = getenv("NNGP_TRACE_ITER") != nullptr;
  double tr_forms = 0.0, tr_host = 0.0, tr_enqueue = 0.0, tr_wait = 0.0, tr_tail = 0.0;
  auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  c->trace_enqueue_s = c->trace_wait_s = 0.0;

  // Draws of the NEXT iteration taken while this iteration's kernel runs (mode 1).  The reference consumes a fixed
  // sequence per iteration -- p normals for beta, [n for a replicate,] 3 or 4 for the proposals, one uniform
  // (src/rNNGP.cpp:295-349) -- none of which depends on the accept decision, so taking the uniform of iteration s and
  // the raw normals of iteration s + 1 early leaves the stream, and the chain, unchanged; rnorm(mu, sd) is
  // mu + sd * norm_rand() (nmath/rnorm.c), formed below without contraction.  With them both possible candidates of
  // iteration s + 1 are known, and the Matern tables of both smoothness values are prepared while the GPU works
  // (8-14 us each, otherwise on the critical path of every iteration).
  struct Predrawn {
    bool have = false;
    double u = 0.0, zp[4] = {0, 0, 0, 0};
    std::vector<double> zb;
  } pre;
  pre.zb.resize(p);
  const bool can_predraw = (a->mode != 0) && getenv("NNGP_NO_PREDRAW") == nullptr;
  auto mul_rn = [](double x, double y) { volatile double r = x * y; return (double)r; };  // product rounded on its own
  auto propose = [&](const double *from, const double *z, double *to) {  // src/rNNGP.cpp:327-333 with the normals given
    to[2] = logit_inv_(logit_(from[2], a->phiUnif[0], a->phiUnif[1]) + mul_rn(tuning[2], z[0]), a->phiUnif[0], a->phiUnif[1]);
    to[0] = exp(log(from[0]) + mul_rn(tuning[0], z[1]));
    to[1] = exp(log(from[1]) + mul_rn(tuning[1], z[2]));
    to[3] = 0.0;
    if (cov == NNGP_COV_MATERN)
      to[3] = logit_inv_(logit_(from[3], a->nuUnif[0], a->nuUnif[1]) + mul_rn(tuning[3], z[3]), a->nuUnif[0], a->nuUnif[1]);
  };

  bool thetaUpdate = true;
  int accept = 0, batchAccept = 0, status = 0, repCnt = 0;
  for (int s = 0; s < a->nSamples; s++) {
    const double t_it0 = trace_iter ? now() : 0.0;
    if (thetaUpdate) {  // src/rNNGP.cpp:274-292
      thetaUpdate = false;
      rc = nngp_q_forms(c, cur, S.data());
      if (rc) return rc;
      if (a->allreduce) a->allreduce(S.data(), K * K, a->allreduce_user);
      for (int i = 0; i < p; i++) {
        tmp_p[i] = S[(size_t)p * K + i];  // X_i' C^-1 y
        for (int j = 0; j <= i; j++) tmp_pp[(size_t)j * p + i] = S[(size_t)j * K + i];
      }
      if (hl::chol(tmp_pp.data(), p)) return ctx_fail(c, NNGP_EINVAL, "c++ error: dpotrf failed (X'C^-1X)");
      hl::chol_inverse(tmp_pp.data(), p);
      hl::symv(tmp_pp.data(), p, tmp_p.data(), mean.data());
      if (hl::chol(tmp_pp.data(), p)) return ctx_fail(c, NNGP_EINVAL, "c++ error: dpotrf failed (inverse of X'C^-1X)");
    }
    const double t_it1 = trace_iter ? now() : 0.0;
    // beta ~ N(mean, (X'C^-1X)^-1): mvrnorm, src/util.cpp:28-41
    const bool have_pre = pre.have;  // this iteration's draws were taken during the previous kernel
    pre.have = false;
    for (int i = 0; i < p; i++) beta[i] = have_pre ? pre.zb[i] : a->rnorm(0.0, 1.0);
    hl::trmv_lower(tmp_pp.data(), p, beta.data());
    for (int i = 0; i < p; i++) beta[i] += mean[i];

    // fit.rep: the n normals of updateRep come between the beta draw and the proposals (src/rNNGP.cpp:309-313)
    const bool rep = a->repIndx && a->repSamples && a->repIndx[s];
    if (rep) {
      rc = rep_prepare(c);
      if (rc) return rc;
      for (int i = 0; i < c->n; i++) c->h_z[i] = a->rnorm(0.0, 1.0);
    }

    // proposals, in the reference's draw order phi, sigma.sq, tau.sq[, nu] (src/rNNGP.cpp:327-333)
    if (have_pre) {
      propose(theta, pre.zp, cand);
    } else {
      cand[2] = logit_inv_(a->rnorm(logit_(theta[2], a->phiUnif[0], a->phiUnif[1]), tuning[2]), a->phiUnif[0], a->phiUnif[1]);
      cand[0] = exp(a->rnorm(log(theta[0]), tuning[0]));
      cand[1] = exp(a->rnorm(log(theta[1]), tuning[1]));
      if (cov == NNGP_COV_MATERN)
        cand[3] = logit_inv_(a->rnorm(logit_(theta[3], a->nuUnif[0], a->nuUnif[1]), tuning[3]), a->nuUnif[0], a->nuUnif[1]);
    }
    // what the host does while the kernel runs: this iteration's uniform, the next iteration's normals, both of its
    // possible Matern preparations
    const bool next_rep = a->repIndx && a->repSamples && s + 1 < a->nSamples && a->repIndx[s + 1];
    const bool predraw = can_predraw && !rep && !next_rep && s + 1 < a->nSamples;
    double u_accept = 0.0;
    bool have_u = false;
    auto between = [&]() {
      if (!predraw) return;
      u_accept = a->runif(0.0, 1.0);
      have_u = true;
      for (int i = 0; i < p; i++) pre.zb[i] = a->rnorm(0.0, 1.0);
      // draw order phi, sigma.sq, tau.sq[, nu]; rnorm(mu, 0) returns mu WITHOUT consuming a normal (nmath/rnorm.c)
      const double sd[4] = {tuning[2], tuning[0], tuning[1], tuning[3]};
      for (int k = 0; k < nTheta; k++) pre.zp[k] = (sd[k] == 0.0) ? 0.0 : a->rnorm(0.0, 1.0);
      pre.have = true;
      if (cov == NNGP_COV_MATERN) {
        double t1[4], t2[4];
        propose(theta, pre.zp, t1);  // if this iteration's candidate is rejected
        propose(cand, pre.zp, t2);   // if it is accepted
        matern_cache_warm(c, t1[3]);
        matern_cache_warm(c, t2[3]);
      }
    };

    double QCur, logDetCand, QCand;
    const double t_it2 = trace_iter ? now() : 0.0;
    const double e0 = c->trace_enqueue_s, w0 = c->trace_wait_s;
    if (a->mode == 0) {
      // reference-faithful: re-solve the current theta and solve the candidate, one launch
      double th2[8], b2[2 * 64];
      if (p > 64) return ctx_fail(c, NNGP_EINVAL, "p > 64 not supported by the batched launch");
      memcpy(th2, theta, sizeof theta);
      memcpy(th2 + 4, cand, sizeof cand);
      memcpy(b2, beta.data(), p * sizeof(double));
      memcpy(b2 + p, beta.data(), p * sizeof(double));
      slots[0] = cur;
      slots[1] = 1 - cur;
      rc = nngp_bf_loglik(c, cov, 2, th2, b2, slots, out);
      if (rc) return rc;
      red[0] = out[0]; red[1] = out[1]; red[2] = out[2]; red[3] = out[3];
      if (a->allreduce) a->allreduce(red, 4, a->allreduce_user);
      logDetCur = red[0]; QCur = red[1]; logDetCand = red[2]; QCand = red[3];
    } else {
      // double-buffered: B,F of the current theta are still resident; only Q changes with beta
      // one launch: solve the candidate and, with the same residual gathers, Q of the resident factors
      double o3[3];
      rc = bf_loglik_cur_impl(c, cov, cand, beta.data(), 1 - cur, cur, o3, between);
      if (rc) return rc;
      red[2] = o3[0]; red[3] = o3[1];
      red[0] = 0.0; red[1] = o3[2];
      if (a->allreduce) a->allreduce(red, 4, a->allreduce_user);
      QCur = red[1]; logDetCand = red[2]; QCand = red[3];
    }
    const double t_it3 = trace_iter ? now() : 0.0;
    if (rep) {  // B, F of the current theta are in slot `cur` in both modes
      rc = rep_draw(c, cur, beta.data(), c->h_z, a->repSamples + (size_t)repCnt * c->n);
      if (rc) return rc;
      repCnt++;
    }
    const double lpCur = log_post_(logDetCur, QCur, theta, a);
    const double lpCand = log_post_(logDetCand, QCand, cand, a);
    if ((have_u ? u_accept : a->runif(0.0, 1.0)) <= exp(lpCand - lpCur)) {  // src/rNNGP.cpp:349
      thetaUpdate = true;
      memcpy(theta, cand, sizeof theta);
      cur = 1 - cur;
      logDetCur = logDetCand;
      accept++;
      batchAccept++;
    }
    memcpy(betaSamples + (size_t)s * p, beta.data(), p * sizeof(double));
    memcpy(thetaSamples + (size_t)s * nTheta, theta, nTheta * sizeof(double));

    if (status == a->nReport) {  // src/rNNGP.cpp:361-373
      if (a->verbose && a->print) {
        snprintf(line, sizeof line, "Sampled: %i of %i, %3.2f%%\n", s, a->nSamples, 100.0 * s / a->nSamples); a->print(line);
        snprintf(line, sizeof line, "Report interval Metrop. Acceptance rate: %3.2f%%\n", 100.0 * batchAccept / a->nReport); a->print(line);
        snprintf(line, sizeof line, "Overall Metrop. Acceptance rate: %3.2f%%\n", 100.0 * accept / s); a->print(line);
        a->print("-------------------------------------------------\n");
      }
      batchAccept = 0;
      status = 0;
    }
    status++;
    if (a->interrupt && a->interrupt()) return ctx_fail(c, NNGP_EINTR, "interrupted by the caller");
    if (trace_iter) {
      tr_forms += t_it1 - t_it0;
      tr_host += t_it2 - t_it1;
      tr_enqueue += c->trace_enqueue_s - e0;
      tr_wait += c->trace_wait_s - w0;
      tr_tail += now() - t_it3;
    }
  }
  if (trace_iter && a->nSamples > 0) {
    const double k = 1e6 / a->nSamples;
    fprintf(stderr,
            "{\"trace\": \"nngp_rnngp\", \"rank\": %d, \"ranks\": %d, \"iterations\": %d, \"accepted\": %d, \"us_per_iteration\": "
            "{\"q_forms_on_accepts\": %.1f, \"host_draws\": %.1f, \"prepare_and_launch\": %.1f, \"wait_for_stamps\": %.1f, "
            "\"accept_and_store\": %.1f}}\n",
            c->comm_rank, c->comm_size, a->nSamples, accept, tr_forms * k, tr_host * k, tr_enqueue * k, tr_wait * k, tr_tail * k);
  }
  if (a->verbose && a->print) {
    snprintf(line, sizeof line, "Sampled: %i of %i, %3.2f%%\n", a->nSamples, a->nSamples, 100.0); a->print(line);
    snprintf(line, sizeof line, "Report interval Metrop. Acceptance rate: %3.2f%%\n", 100.0 * batchAccept / a->nReport); a->print(line);
    snprintf(line, sizeof line, "Overall Metrop. Acceptance rate: %3.2f%%\n", 100.0 * accept / a->nSamples); a->print(line);
    a->print("-------------------------------------------------\n");
  }
  if (accept_out) *accept_out = accept;
  return NNGP_OK;
}

int nngp_rep_draw(nngp_ctx *c, int slot, const double *beta, const double *z, double *yrep) {
  if (!c || !beta || !z || !yrep || slot < 0 || slot > 1) return NNGP_EINVAL;
  NNGP_CUDA(c, cudaSetDevice(c->device));
  return rep_draw(c, slot, beta, z, yrep);
}

int nngp_rnngp_replicated(nngp_ctx *c, int cov, const double *betaSamples, const double *thetaSamples, int nSamples,
                          double (*rnorm)(double, double), nngp_interrupt_fn interrupt, double *repSamples) {
  if (!c || !betaSamples || !thetaSamples || !rnorm || !repSamples || nSamples < 0) return NNGP_EINVAL;
  if (c->n == 0 || c->p == 0 || !c->has_nn) return ctx_fail(c, NNGP_ESTATE, "nngp_rnngp_replicated before nngp_set_data");
  if (cov < 0 || cov > 3) return ctx_fail(c, NNGP_EINVAL, "cov.model is not correctly specified");
  NNGP_CUDA(c, cudaSetDevice(c->device));
  const int p = c->p, nTheta = (cov == NNGP_COV_MATERN) ? 4 : 3;
  int rc = rep_prepare(c);
  if (rc) return rc;
  std::vector<double> beta(p);
  double prev[4] = {-1, -1, -1, -1}, out[2];
  int slot = 0;
  for (int s = 0; s < nSamples; s++) {
    // B, F at theta_s (src/rNNGPReplicated.cpp:185); rejected proposals repeat columns: solve once per run of equal theta
    double th[4] = {0, 0, 0, 0};
    memcpy(th, thetaSamples + (size_t)s * nTheta, nTheta * sizeof(double));
    if (memcmp(th, prev, sizeof th) != 0) {
      rc = nngp_bf_loglik(c, cov, 1, th, nullptr, &slot, out);  // asynchronous to the host draws below until its sync
      if (rc) return rc;
      memcpy(prev, th, sizeof th);
    }
    for (int i = 0; i < c->n; i++) c->h_z[i] = rnorm(0.0, 1.0);
    for (int j = 0; j < p; j++) beta[j] = betaSamples[(size_t)j * nSamples + s];  // nSamples x p (R/generics.R:252)
    rc = rep_draw(c, slot, beta.data(), c->h_z, repSamples + (size_t)s * c->n);
    if (rc) return rc;
    if (interrupt && interrupt()) return ctx_fail(c, NNGP_EINTR, "interrupted by the caller");
  }
  return NNGP_OK;
}

}  // extern "C"

// nngp_ctx.h -- the engine context behind the C ABI (one context drives one GPU).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>

#include <string>
#include <vector>

#include "nngp_common.cuh"

namespace nngp {
class HostStager;
}

struct nngp_ctx {
  int device = 0;
  int sms = 148;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  std::string err;
  long long launches = 0;

  // model data (device)
  int n = 0, m = 0, p = 0;
  int loc_begin = 0, loc_end = 0;
  long long nIndx = 0;
  bool has_nn = false;  // false when only sites/X/y were given (prediction-only use)
  double2 *d_coords = nullptr;  // n interleaved sites
  // Storage is SHARDED: only the slice of nnIndx / B / F that belongs to [loc_begin, loc_end) is allocated and
  // uploaded (n = 10^7, m = 32 is 1.28 + 2 x 2.56 GB when replicated).  The kernels keep indexing with the
  // reference's global CSR offsets (src/util.cpp:55-70) through VIRTUAL base pointers = allocation - first index of
  // the shard, valid for [csr_begin, csr_end) resp. [loc_begin, loc_end) only.
  long long csr_begin = 0, csr_end = 0;
  int *d_nnIndx_store = nullptr;
  double *d_B_store[2] = {nullptr, nullptr};
  double *d_F_store[2] = {nullptr, nullptr};
  int *d_nnIndx = nullptr;      // reference CSR (virtual base)
  double *d_X = nullptr;        // n x p column-major
  double *d_y = nullptr;
  double *d_Xy = nullptr;       // [X y] as rows, n x roundup4(p+1) row-major, zero padded (q_forms gathers)
  double *d_B[2] = {nullptr, nullptr};  // two slots: current / candidate (virtual bases; complete rows of a 32-location
                                        // tile start 16-byte aligned, so bulk copies can move them)
  double *d_F[2] = {nullptr, nullptr};
  // resident distance triangle of the shard (tri_cache.cu), NULL when disabled / m not served / out of memory
  double *d_tri = nullptr;
  size_t tri_cap = 0;
  int opt_tri = 1;  // nngp_set_option("tri_cache", 0/1)
  bool tri_on = false;  // d_tri describes the current data set
  double *d_resid[NNGP_MAX_BATCH] = {};
  double *d_beta = nullptr;             // NNGP_MAX_BATCH x p
  nngp::MaternConsts *d_mat = nullptr;  // NNGP_MAX_BATCH
  double *d_partials = nullptr;         // scratch for the grid reductions
  size_t partials_len = 0;
  size_t partials_stride = 0;  // doubles per parameter set
  unsigned int *d_ticket = nullptr;  // NNGP_MAX_BATCH
  int *d_fail = nullptr;             // NNGP_MAX_BATCH
  double *d_out = nullptr;           // device result slots
  size_t out_len = 0;

  // pinned host staging
  double *h_out = nullptr;
  int *h_fail = nullptr;
  int *h_fail_dev = nullptr;  // device alias of h_fail
  double *h_beta = nullptr;
  size_t h_beta_cap = 0;
  // replicated-data draws (rep_kernels.cu): level structure of the neighbour graph, built on first use
  bool rep_ready = false;
  int rep_levels = 0;
  int *d_rep_perm = nullptr, *d_rep_levoff = nullptr;
  double *d_rep_z = nullptr, *d_rep_v = nullptr;  // n doubles each: the normals, the solved v / y.rep
  double *h_z = nullptr;  // pinned, n doubles
  size_t h_z_cap = 0;
  nngp::HostStager *stager = nullptr;  // pinned upload ring of nngp_set_data (host_stage.h)
  nngp::HostStager *stager2 = nullptr; // second ring: downloads that overlap uploads (nngp_rnngp_predict)
  size_t cap[16] = {};  // capacities (bytes) of the grow-only model buffers, see nngp_set_data
  nngp::MaternConsts *h_mat = nullptr;
  // last few prepared smoothness values (the sampler re-evaluates its current theta every iteration)
  struct MatCache {
    double nu = -1.0;
    nngp::MaternConsts mc;
  } mat_cache[4];
  int mat_cache_next = 0;
  double d_mat_nu[NNGP_MAX_BATCH] = {-1.0, -1.0, -1.0, -1.0};  // smoothness whose tables sit in d_mat[b]

  // Matern short series (matern_series.cuh): chosen per parameter set when phi^2/4 * pair_d2_q99 <= t_short_max(nu)
  int opt_matern_short = 2;      // nngp_set_option("matern_short", 0/1/2): 0 full series, 1 at most the short one, 2 tiny allowed
  int resid_begin = 0, resid_end = 0;  // residuals are formed for [resid_begin, resid_end): the shard's locations and every neighbour of theirs
  double pair_d2_q99 = 1e300;    // 99 % quantile over this shard's locations of the largest squared pair distance
  unsigned int *d_hist = nullptr;
  unsigned int *h_hist = nullptr;   // pinned copy, filled asynchronously behind the histogram kernel
  cudaEvent_t ev_hist = nullptr;
  bool hist_pending = false;        // the quantile has not been formed from h_hist yet (pair_quantile_finish)
  unsigned char bf_short[NNGP_MAX_BATCH] = {};  // decision of the last prepare_bf, per parameter set
  double last_qcur[NNGP_MAX_BATCH] = {};  // Q of the resident factors returned by the last launch (BfTheta::Bcur)

  // multi-GPU: NCCL communicator over the ranks of a sharded fit (nngp_comm_init); when set, the
  // results of nngp_bf_loglik / nngp_q_resid / nngp_q_forms are summed over the ranks on the device
  void *comm = nullptr;
  int comm_rank = 0, comm_size = 1;
  double *d_red = nullptr;  // kBfOutStride * NNGP_MAX_BATCH doubles: kernel results, reduced in place (NCCL path)
  double *h_red = nullptr;  // pinned copy

  // MAILBOX (nngp_abi_misc.cu): how small results leave the device.  A ring of two slots per rank in mapped host memory;
  // the kernel that finishes a reduction stores its numbers and then a sequence stamp straight into this rank's slot,
  // the host polls the stamp -- no cudaMemcpy, no stream synchronisation.  With nngp_comm_init the ring lives in a
  // POSIX shared-memory object registered by every rank of the node: each host reads ALL ranks' slots and adds them in
  // rank order (deterministic, identical on every rank), which replaces ncclAllReduce + copy + sync for the 48..100
  // bytes per iteration of a sharded fit.  nngp_set_option("reduce_nccl", 1) keeps the NCCL path (tested alternative).
  double *mbox_h = nullptr;   // [2][mbox_ranks][kMboxSlotDoubles]
  double *mbox_d = nullptr;   // device alias
  size_t mbox_bytes = 0;
  int mbox_ranks = 1, mbox_rank = 0;
  bool mbox_shared = false;   // true: shm-backed (mbox_h came from mmap), false: cudaHostAlloc
  int opt_reduce_nccl = 0;
  unsigned long long seq = 0; // collective counter: the stamp of the most recent posting
  double trace_enqueue_s = 0.0, trace_wait_s = 0.0;  // host seconds spent preparing + launching / waiting for stamps

  // timing
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  void *d_flush = nullptr;
  size_t flush_bytes = 0;
};

namespace nngp {

// Transient device buffers of the search / prediction entry points.  They come from the device's stream-ordered pool
// (cudaMallocAsync on the context stream; nngp_create raises the pool's release threshold so freed blocks stay cached):
// plain cudaMalloc / cudaFree cost 50-650 ms per call once a few GB had been allocated and released by an earlier
// prediction (traced at n = 10^6: 16 ms of work inside a 340 ms nngp_nn_query).  ScratchScope names the stream for the
// buffers declared below it; every entry point that declares a DevBuf opens one.
inline cudaStream_t &scratch_stream() {
  static thread_local cudaStream_t s = nullptr;
  return s;
}
struct ScratchScope {
  cudaStream_t prev;
  explicit ScratchScope(cudaStream_t s) : prev(scratch_stream()) { scratch_stream() = s; }
  ~ScratchScope() { scratch_stream() = prev; }
};
struct DevBuf {
  void *p = nullptr;
  cudaStream_t s = nullptr;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  DevBuf(DevBuf &&o) noexcept : p(o.p), s(o.s) { o.p = nullptr; }
  ~DevBuf() { release(); }
  void release() {
    if (p) cudaFreeAsync(p, s);
    p = nullptr;
  }
  cudaError_t alloc(size_t bytes) {
    release();
    s = scratch_stream();
    return cudaMallocAsync(&p, bytes ? bytes : 1, s);
  }
  template <typename Tp>
  Tp *as() const {
    return reinterpret_cast<Tp *>(p);
  }
};

// NNGP_TRACE=1: phase timings of the host-side entry points on stderr
struct PhaseTrace {
  bool on;
  double t0;
  const char *what;
  static double now() {
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec + 1e-9 * ts.tv_nsec;
  }
  explicit PhaseTrace(const char *w) : what(w) {
    static int flag = -1;
    if (flag < 0) {
      const char *e = getenv("NNGP_TRACE");
      flag = (e && e[0] == '1') ? 1 : 0;
    }
    on = flag == 1;
    t0 = on ? now() : 0.0;
  }
  void mark(const char *phase) {
    if (!on) return;
    const double t = now();
    fprintf(stderr, "[nngp trace] %s: %s %.3f ms\n", what, phase, (t - t0) * 1e3);
    t0 = t;
  }
};

// bf_kernel.cu
int bf_tile_for_m(int m);
int bf_locations_per_cta(int m);
int bf_grid_for(int m, long long nloc, int sms);
int bf_grid_any(int m, int cov, long long nloc, int sms);
int bf_grid_max(int m, long long nloc, int sms);
cudaError_t bf_launch(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream);

// bf_tri_kernel.cu / tri_cache.cu
int tri_capacity_for_m(int m);
void tri_set_force(int on);
int tri_grid_for(long long nloc, int sms);
cudaError_t tri_launch(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream);
size_t tri_cache_doubles(int m, long long nloc);
size_t tri_ids_offset(int m, long long nloc);
cudaError_t tri_build_launch(const double2 *coords, const int *nnIndx, int m, int loc_begin, int loc_end, double *tri,
                             cudaStream_t stream);

// nngp_abi_misc.cu: every rank holds sites [n r / N, n (r+1) / N) of the replicated arrays -> every rank holds all of them
int comm_exchange_replicated(nngp_ctx *c, double2 *coords, double *X, double *y, int n, int p);

// nngp_abi_misc.cu: mailbox
constexpr int kMboxSlotDoubles = 2304;  // [0] stamp of a generic posting, [8 ..] payload (B/F sets: 8 doubles each, own stamps)
constexpr int kMboxPayload = 8;
int mbox_local_init(nngp_ctx *c);
void mbox_release(nngp_ctx *c);
inline bool mbox_reduces(const nngp_ctx *c) { return c->comm == nullptr || (c->mbox_shared && !c->opt_reduce_nccl); }
inline double *mbox_slot_dev(nngp_ctx *c, unsigned long long seq) {
  return c->mbox_d + ((size_t)(seq & 1) * c->mbox_ranks + c->mbox_rank) * kMboxSlotDoubles;
}
// B/F form: waits for the stamps of `nbatch` result blocks of every rank, res[8 b + j] = sum over ranks (j < 4)
int mbox_wait_sets(nngp_ctx *c, unsigned long long seq, int nbatch, double *res);
// generic form: the device posts len doubles from d_src (a tiny kernel on the context stream); wait adds all ranks' copies
int mbox_post_dev(nngp_ctx *c, const double *d_src, int len, unsigned long long seq);
int mbox_wait_payload(nngp_ctx *c, unsigned long long seq, int len, double *sum);
int mbox_host_allreduce(nngp_ctx *c, double *buf, int len);

// q_kernels.cu
cudaError_t resid_launch(const double *X, const double *y, const double *beta, int n, int p, double *r, int i0, int i1, int sms,
                         cudaStream_t stream);
cudaError_t resid_launch_val(const double *X, const double *y, const double *beta_host, int n, int p, double *r, int i0, int i1, int sms,
                             cudaStream_t stream);
int q_resid_grid(int m, long long nloc, int sms);
cudaError_t q_resid_launch(const double *B, const double *F, const int *nnIndx, const double *resid, int m,
                           int loc_begin, int loc_end, double *partials, double *out, int sms, cudaStream_t stream);
int q_forms_grid(long long nloc);
cudaError_t q_forms_launch(const double *B, const double *F, const int *nnIndx, const double *Xy, int n, int p, int m,
                           int loc_begin, int loc_end, double *partials, double *out, cudaStream_t stream);
size_t xy_rows_count(int n, int p);
cudaError_t xy_rows_launch(const double *X, const double *y, int n, int p, double *Xy, int sms, cudaStream_t stream);

// bf_kernel.cu: true when the kernel that serves m neighbours has the short-series Matern variant
bool bf_has_matern_short(int m);
// stat_kernels.cu
int pair_d2_hist_bins();
cudaError_t pair_d2_hist_launch(const double2 *coords, const int *nnIndx, int m, int loc_begin, int loc_end,
                                unsigned int *hist, cudaStream_t stream);

// rep_kernels.cu
int rep_prepare(nngp_ctx *c);  // levels + pinned z buffer (c->h_z), idempotent
int rep_draw(nngp_ctx *c, int slot, const double *beta, const double *z, double *yrep);

// shared helpers (nngp_abi.cu)
int ctx_fail(nngp_ctx *ctx, int code, const std::string &msg);
int ctx_cuda(nngp_ctx *ctx, cudaError_t e, const char *what);
void set_global_error(const std::string &msg);
// nngp_abi_misc.cu: in-place sum over ranks of `len` doubles in device memory, on the context stream
int comm_allreduce_dev(nngp_ctx *ctx, double *d_buf, int len);

}  // namespace nngp

#define NNGP_CUDA(ctx, call)                                     \
  do {                                                           \
    cudaError_t e__ = (call);                                    \
    if (e__ != cudaSuccess) return nngp::ctx_cuda(ctx, e__, #call); \
  } while (0)

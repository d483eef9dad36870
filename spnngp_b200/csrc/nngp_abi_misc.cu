// nngp_abi_misc.cu -- measurement helpers of the C ABI: the FP64 FMA peak of this device (roofline
// denominator for the B/F kernels; MEASURED_PEAKS.json only has HBM and bf16 numbers).
#include <string.h>

#include "nngp_ctx.h"

namespace nngp {

// 8 independent DFMA chains per thread: enough ILP to saturate the FP64 pipe at any occupancy
__global__ void __launch_bounds__(256) fp64_peak_kernel(double *out, int iters, double a, double b) {
  double x0 = threadIdx.x * 1e-3, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
#pragma unroll 1
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      x0 = fma(x0, a, b); x1 = fma(x1, a, b); x2 = fma(x2, a, b); x3 = fma(x3, a, b);
      x4 = fma(x4, a, b); x5 = fma(x5, a, b); x6 = fma(x6, a, b); x7 = fma(x7, a, b);
    }
  }
  const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
  if (s == 123.456) out[0] = s;  // never true: keeps the chains alive
}

}  // namespace nngp

using namespace nngp;

// Page-locked host arrays for callers that can place their big inputs / outputs there (the q x S normals and draws of
// nngp_rnngp_predict): the transfer paths DMA straight from / to such arrays instead of staging through the pinned ring.
extern "C" int nngp_host_alloc(nngp_ctx *c, size_t bytes, void **out) {
  if (!c || !out) return NNGP_EINVAL;
  NNGP_CUDA(c, cudaSetDevice(c->device));
  NNGP_CUDA(c, cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable));
  return NNGP_OK;
}

extern "C" int nngp_host_free(nngp_ctx *c, void *p) {
  if (!c) return NNGP_EINVAL;
  if (p) NNGP_CUDA(c, cudaFreeHost(p));
  return NNGP_OK;
}

extern "C" int nngp_fp64_peak(nngp_ctx *c, double *tflops) {
  if (!c || !tflops) return NNGP_EINVAL;
  NNGP_CUDA(c, cudaSetDevice(c->device));
  double *d = nullptr;
  NNGP_CUDA(c, cudaMalloc(&d, sizeof(double)));
  const int grid = c->sms * 8, iters = 2048;
  fp64_peak_kernel<<<grid, 256, 0, c->stream>>>(d, 64, 0.999999, 1e-9);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; rep++) {
    NNGP_CUDA(c, cudaEventRecord(c->ev0, c->stream));
    fp64_peak_kernel<<<grid, 256, 0, c->stream>>>(d, iters, 0.999999, 1e-9);
    NNGP_CUDA(c, cudaEventRecord(c->ev1, c->stream));
    NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
    float ms = 0.f;
    NNGP_CUDA(c, cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    const double flops = (double)grid * 256 * (double)iters * 64 * 2.0;
    const double tf = flops / (ms * 1e-3) / 1e12;
    if (tf > best) best = tf;
  }
  c->launches += 6;
  cudaFree(d);
  *tflops = best;
  return NNGP_OK;
}

// ---------------------------------------------------------------------------------------------------
// Multi-GPU: an NCCL communicator over the ranks of a sharded fit.  libnccl is loaded lazily with
// dlopen (single-GPU users need no NCCL; inside a torch process the already-loaded libnccl.so.2 is
// reused).  The only collective of the fit path is an in-place sum of <= 3*4 doubles (log-likelihood
// partials) or (p+1)(p+2)/2 doubles (cross quadratic forms), issued on the context stream right behind
// the kernel that produced them, so the host sees one synchronisation per likelihood evaluation.
// ---------------------------------------------------------------------------------------------------
#include <dlfcn.h>
#include <errno.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <chrono>

#include <mutex>

// The handful of NCCL declarations this file needs (values fixed by NCCL's ABI, nccl.h 2.x), so that the engine builds
// without the NCCL headers: the library itself is only ever dlopen'ed.
extern "C" {
typedef struct ncclComm *ncclComm_t;
#define NCCL_UNIQUE_ID_BYTES 128
typedef struct { char internal[NCCL_UNIQUE_ID_BYTES]; } ncclUniqueId;
typedef enum { ncclSuccess = 0 } ncclResult_t;
typedef enum { ncclChar = 0, ncclDouble = 8 } ncclDataType_t;
typedef enum { ncclSum = 0 } ncclRedOp_t;
}

namespace nngp {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void *, void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi *nccl_api(std::string *why) {
  static NcclApi api;
  static std::once_flag once;
  static std::string err;
  std::call_once(once, [] {
    const char *names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char *nm : names) {
      api.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (api.handle) break;
    }
    if (!api.handle) {
      err = std::string("cannot load libnccl: ") + dlerror();
      return;
    }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.handle, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.handle, "ncclCommInitRank");
    api.AllReduce = (decltype(api.AllReduce))dlsym(api.handle, "ncclAllReduce");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.handle, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.handle, "ncclGetErrorString");
    api.Broadcast = (decltype(api.Broadcast))dlsym(api.handle, "ncclBroadcast");
    api.GroupStart = (decltype(api.GroupStart))dlsym(api.handle, "ncclGroupStart");
    api.GroupEnd = (decltype(api.GroupEnd))dlsym(api.handle, "ncclGroupEnd");
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce) err = "libnccl lacks the expected symbols";
  });
  if (!err.empty()) {
    if (why) *why = err;
    return nullptr;
  }
  return &api;
}

int comm_allreduce_dev(nngp_ctx *c, double *d_buf, int len);

// ---- mailbox ------------------------------------------------------------------------------------------
__global__ void mbox_post_kernel(const double *__restrict__ src, int len, double *slot, double seq) {
  for (int x = threadIdx.x; x < len; x += blockDim.x) slot[kMboxPayload + x] = src[x];
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) *reinterpret_cast<volatile double *>(slot) = seq;
}

int mbox_local_init(nngp_ctx *c) {
  c->mbox_ranks = 1;
  c->mbox_rank = 0;
  c->mbox_shared = false;
  c->mbox_bytes = (size_t)2 * kMboxSlotDoubles * sizeof(double);
  NNGP_CUDA(c, cudaHostAlloc(&c->mbox_h, c->mbox_bytes, cudaHostAllocMapped));
  memset(c->mbox_h, 0, c->mbox_bytes);
  NNGP_CUDA(c, cudaHostGetDevicePointer((void **)&c->mbox_d, c->mbox_h, 0));
  return NNGP_OK;
}

void mbox_release(nngp_ctx *c) {
  if (!c->mbox_h) return;
  if (c->mbox_shared) {
    cudaHostUnregister(c->mbox_h);
    munmap(c->mbox_h, c->mbox_bytes);
  } else {
    cudaFreeHost(c->mbox_h);
  }
  c->mbox_h = c->mbox_d = nullptr;
  c->mbox_shared = false;
  c->mbox_ranks = 1;
  c->mbox_rank = 0;
}

// Every rank of the node maps the same object; the name comes from the (random, per-job) NCCL id.
static int mbox_shared_init(nngp_ctx *c, const char id[128], int rank, int nranks) {
  unsigned long long h = 1469598103934665603ull;
  for (int i = 0; i < 128; i++) h = (h ^ (unsigned char)id[i]) * 1099511628211ull;
  char name[64];
  snprintf(name, sizeof name, "/nngp_mbox_%016llx", h);
  const size_t bytes = (size_t)2 * nranks * kMboxSlotDoubles * sizeof(double);
  int fd = shm_open(name, O_CREAT | O_RDWR, 0600);
  if (fd < 0) return ctx_fail(c, NNGP_ESTATE, std::string("mailbox: shm_open failed: ") + strerror(errno));
  if (ftruncate(fd, (off_t)bytes) != 0) {
    close(fd);
    return ctx_fail(c, NNGP_ESTATE, std::string("mailbox: ftruncate failed: ") + strerror(errno));
  }
  void *p = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) return ctx_fail(c, NNGP_ESTATE, std::string("mailbox: mmap failed: ") + strerror(errno));
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterMapped | cudaHostRegisterPortable);
  if (e != cudaSuccess) {
    munmap(p, bytes);
    return ctx_fail(c, NNGP_ECUDA, std::string("mailbox: cudaHostRegister: ") + cudaGetErrorString(e));
  }
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  mbox_release(c);
  c->mbox_h = (double *)p;
  c->mbox_bytes = bytes;
  NNGP_CUDA(c, cudaHostGetDevicePointer((void **)&c->mbox_d, p, 0));
  c->mbox_ranks = nranks;
  c->mbox_rank = rank;
  c->mbox_shared = true;
  c->seq = 0;  // every rank starts counting at the same point
  // everybody has mapped the object once this sum (posted by the hosts through the mailbox itself) returns: the name
  // can go.  A rank that arrives late finds the object by name until then; ranks never outrun each other by more than
  // one posting, so nobody unlinks before the slowest rank has opened it.
  double one = 1.0;
  int rc = mbox_host_allreduce(c, &one, 1);
  if (rc == NNGP_OK && one != (double)nranks) rc = ctx_fail(c, NNGP_ESTATE, "mailbox: rank count mismatch");
  shm_unlink(name);
  return rc;
}

// spin on a stamp; now and then make sure the stream is still alive, give up on a peer after two minutes
static int mbox_spin(nngp_ctx *c, const volatile double *stamp, double seq, bool own) {
  const auto t0 = std::chrono::steady_clock::now();
  for (unsigned long long it = 1;; ++it) {
    if (*stamp == seq) {
      std::atomic_thread_fence(std::memory_order_acquire);
      return NNGP_OK;
    }
    if ((it & 0x3fff) == 0) {
      const cudaError_t e = cudaStreamQuery(c->stream);
      if (e != cudaSuccess && e != cudaErrorNotReady) return ctx_fail(c, NNGP_ECUDA, std::string("CUDA error: ") + cudaGetErrorString(e) + " while waiting for a kernel result");
      if (e == cudaSuccess && own && *stamp != seq) {
        std::atomic_thread_fence(std::memory_order_seq_cst);
        if (*stamp != seq) return ctx_fail(c, NNGP_ESTATE, "mailbox: the stream drained without a result stamp");
      }
      const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (dt > 120.0) return ctx_fail(c, NNGP_ESTATE, "mailbox: timed out waiting for a peer rank's result");
      if (dt > 0.05) sched_yield();
    }
#if defined(__x86_64__)
    __builtin_ia32_pause();
#endif
  }
}

int mbox_wait_sets(nngp_ctx *c, unsigned long long seq, int nbatch, double *res) {
  for (int x = 0; x < kBfOutStride * nbatch; x++) res[x] = 0.0;
  for (int r = 0; r < c->mbox_ranks; r++) {
    const double *slot = c->mbox_h + ((size_t)(seq & 1) * c->mbox_ranks + r) * kMboxSlotDoubles + kMboxPayload;
    for (int b = 0; b < nbatch; b++) {
      int rc = mbox_spin(c, slot + kBfOutStride * b + kBfOutSeq, (double)seq, r == c->mbox_rank);
      if (rc) return rc;
      for (int j = 0; j < 4; j++) res[kBfOutStride * b + j] += slot[kBfOutStride * b + j];
    }
  }
  return NNGP_OK;
}

int mbox_post_dev(nngp_ctx *c, const double *d_src, int len, unsigned long long seq) {
  if (len > kMboxSlotDoubles - kMboxPayload) return ctx_fail(c, NNGP_EINVAL, "mailbox: payload too long");
  mbox_post_kernel<<<1, 256, 0, c->stream>>>(d_src, len, mbox_slot_dev(c, seq), (double)seq);
  NNGP_CUDA(c, cudaGetLastError());
  c->launches++;
  return NNGP_OK;
}

int mbox_wait_payload(nngp_ctx *c, unsigned long long seq, int len, double *sum) {
  for (int x = 0; x < len; x++) sum[x] = 0.0;
  for (int r = 0; r < c->mbox_ranks; r++) {
    const double *slot = c->mbox_h + ((size_t)(seq & 1) * c->mbox_ranks + r) * kMboxSlotDoubles;
    int rc = mbox_spin(c, slot, (double)seq, r == c->mbox_rank);
    if (rc) return rc;
    for (int x = 0; x < len; x++) sum[x] += slot[kMboxPayload + x];
  }
  return NNGP_OK;
}

int mbox_host_allreduce(nngp_ctx *c, double *buf, int len) {
  if (len > kMboxSlotDoubles - kMboxPayload) return ctx_fail(c, NNGP_EINVAL, "mailbox: payload too long");
  const unsigned long long seq = ++c->seq;
  double *slot = c->mbox_h + ((size_t)(seq & 1) * c->mbox_ranks + c->mbox_rank) * kMboxSlotDoubles;
  memcpy(slot + kMboxPayload, buf, len * sizeof(double));
  std::atomic_thread_fence(std::memory_order_release);
  *reinterpret_cast<volatile double *>(slot) = (double)seq;
  return mbox_wait_payload(c, seq, len, buf);
}

int comm_allreduce_dev(nngp_ctx *c, double *d_buf, int len) {
  if (!c->comm) return NNGP_OK;
  std::string why;
  NcclApi *api = nccl_api(&why);
  if (!api) return ctx_fail(c, NNGP_ESTATE, why);
  ncclResult_t r = api->AllReduce(d_buf, d_buf, (size_t)len, ncclDouble, ncclSum, (ncclComm_t)c->comm, c->stream);
  if (r != ncclSuccess) return ctx_fail(c, NNGP_ECUDA, std::string("ncclAllReduce: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"));
  return NNGP_OK;
}

int comm_exchange_replicated(nngp_ctx *c, double2 *coords, double *X, double *y, int n, int p) {
  std::string why;
  NcclApi *api = nccl_api(&why);
  if (!api || !api->Broadcast || !api->GroupStart || !api->GroupEnd) return ctx_fail(c, NNGP_ESTATE, "libnccl lacks ncclBroadcast / ncclGroupStart");
  ncclComm_t comm = (ncclComm_t)c->comm;
  const int N = c->comm_size;
  bool ok = api->GroupStart() == ncclSuccess;
  for (int r = 0; r < N && ok; r++) {  // piece of root r, in place on every rank
    const long long s0 = (long long)n * r / N, s1 = (long long)n * (r + 1) / N;
    if (s1 <= s0) continue;
    ok = ok && api->Broadcast(coords + s0, coords + s0, (size_t)(s1 - s0) * sizeof(double2), ncclChar, r, comm, c->stream) == ncclSuccess;
    for (int col = 0; col < p && ok; col++)
      ok = api->Broadcast(X + (size_t)col * n + s0, X + (size_t)col * n + s0, (size_t)(s1 - s0) * sizeof(double), ncclChar, r, comm, c->stream) == ncclSuccess;
    if (p > 0) ok = ok && api->Broadcast(y + s0, y + s0, (size_t)(s1 - s0) * sizeof(double), ncclChar, r, comm, c->stream) == ncclSuccess;
  }
  ok = (api->GroupEnd() == ncclSuccess) && ok;
  if (!ok) return ctx_fail(c, NNGP_ECUDA, "ncclBroadcast of the replicated arrays failed");
  c->launches += 1;
  return NNGP_OK;
}

}  // namespace nngp

extern "C" {

int nngp_comm_unique_id(char id[128]) {
  std::string why;
  NcclApi *api = nccl_api(&why);
  if (!api) return ctx_fail(nullptr, NNGP_ESTATE, why);
  ncclUniqueId uid;
  if (api->GetUniqueId(&uid) != ncclSuccess) return ctx_fail(nullptr, NNGP_ECUDA, "ncclGetUniqueId failed");
  memcpy(id, uid.internal, NCCL_UNIQUE_ID_BYTES);
  return NNGP_OK;
}

int nngp_comm_init(nngp_ctx *c, const char id[128], int rank, int nranks) {
  if (!c || !id || nranks < 1 || rank < 0 || rank >= nranks) return NNGP_EINVAL;
  if (nranks == 1) return NNGP_OK;
  std::string why;
  NcclApi *api = nccl_api(&why);
  if (!api) return ctx_fail(c, NNGP_ESTATE, why);
  NNGP_CUDA(c, cudaSetDevice(c->device));
  ncclUniqueId uid;
  memcpy(uid.internal, id, NCCL_UNIQUE_ID_BYTES);
  ncclComm_t comm = nullptr;
  ncclResult_t r = api->CommInitRank(&comm, nranks, uid, rank);
  if (r != ncclSuccess) return ctx_fail(c, NNGP_ECUDA, std::string("ncclCommInitRank: ") + (api->GetErrorString ? api->GetErrorString(r) : "error"));
  c->comm = comm;
  c->comm_rank = rank;
  c->comm_size = nranks;
  const char *env = getenv("NNGP_REDUCE");  // "nccl": keep the collective on NCCL
  if (env && !strcmp(env, "nccl")) c->opt_reduce_nccl = 1;
  return mbox_shared_init(c, id, rank, nranks);
}

// Ranks that share one node need no NCCL at all for these sums: the mailbox alone (several ranks may even share a GPU,
// which NCCL refuses).  key: 128 bytes agreed by all ranks of the job and unique to it (e.g. nngp_comm_unique_id's).
int nngp_comm_init_local(nngp_ctx *c, const char key[128], int rank, int nranks) {
  if (!c || !key || nranks < 1 || rank < 0 || rank >= nranks) return NNGP_EINVAL;
  if (nranks == 1) return NNGP_OK;
  NNGP_CUDA(c, cudaSetDevice(c->device));
  c->comm_rank = rank;
  c->comm_size = nranks;
  return mbox_shared_init(c, key, rank, nranks);
}

int nngp_comm_allreduce_sum(nngp_ctx *c, double *buf, int len) {
  if (!c || !buf || len < 1 || len > nngp::kMboxSlotDoubles - nngp::kMboxPayload) return NNGP_EINVAL;
  if (!c->comm && !c->mbox_shared) return NNGP_OK;
  if (nngp::mbox_reduces(c)) return nngp::mbox_host_allreduce(c, buf, len);
  if (len > nngp::kBfOutStride * NNGP_MAX_BATCH) return NNGP_EINVAL;
  NNGP_CUDA(c, cudaSetDevice(c->device));
  memcpy(c->h_red, buf, len * sizeof(double));
  NNGP_CUDA(c, cudaMemcpyAsync(c->d_red, c->h_red, len * sizeof(double), cudaMemcpyHostToDevice, c->stream));
  int rc = comm_allreduce_dev(c, c->d_red, len);
  if (rc) return rc;
  NNGP_CUDA(c, cudaMemcpyAsync(c->h_red, c->d_red, len * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  memcpy(buf, c->h_red, len * sizeof(double));
  return NNGP_OK;
}

void nngp_comm_destroy(nngp_ctx *c) {
  if (!c || (!c->comm && !c->mbox_shared)) return;
  if (c->comm) {
    NcclApi *api = nccl_api(nullptr);
    if (api && api->CommDestroy) api->CommDestroy((ncclComm_t)c->comm);
  }
  c->comm = nullptr;
  c->comm_size = 1;
  c->comm_rank = 0;
  if (c->mbox_shared) {  // back to a private mailbox
    cudaStreamSynchronize(c->stream);
    nngp::mbox_release(c);
    nngp::mbox_local_init(c);
  }
}

}  // extern "C"

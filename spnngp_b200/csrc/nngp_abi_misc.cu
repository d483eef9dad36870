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
#include <nccl.h>

#include <mutex>

namespace nngp {

struct NcclApi {
  void *handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
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
    if (!api.GetUniqueId || !api.CommInitRank || !api.AllReduce) err = "libnccl lacks the expected symbols";
  });
  if (!err.empty()) {
    if (why) *why = err;
    return nullptr;
  }
  return &api;
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
  return NNGP_OK;
}

int nngp_comm_allreduce_sum(nngp_ctx *c, double *buf, int len) {
  if (!c || !buf || len < 1 || len > nngp::kBfOutStride * NNGP_MAX_BATCH) return NNGP_EINVAL;
  if (!c->comm) return NNGP_OK;
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
  if (!c || !c->comm) return;
  NcclApi *api = nccl_api(nullptr);
  if (api && api->CommDestroy) api->CommDestroy((ncclComm_t)c->comm);
  c->comm = nullptr;
  c->comm_size = 1;
}

}  // extern "C"

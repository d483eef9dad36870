// nngp_abi_misc.cu -- measurement helpers of the C ABI: the FP64 FMA peak of this device (roofline
// denominator for the B/F kernels; MEASURED_PEAKS.json only has HBM and bf16 numbers).
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

// Which warps of a CTA share a scheduler (SM sub-partition)?  Two warps run a saturating DFMA stream; if they sit
// on the same sub-partition they share its FP64 pipe and the pair takes ~2x as long as on different ones.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int wa, int wb, double *out, long long *cyc) {
  const int w = threadIdx.x >> 5;
  if (w != wa && w != wb) return;
  double a[8];
  for (int i = 0; i < 8; i++) a[i] = threadIdx.x + i;
  const double m = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < 20000; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = fma(a[i], m, c);
  }
  long long t1 = clock64();
  double s = 0;
  for (int i = 0; i < 8; i++) s += a[i];
  out[threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[w] = t1 - t0;
}
int main() {
  double *out; long long *cyc;
  cudaMalloc(&out, 1024 * 8); cudaMallocManaged(&cyc, 32 * 8);
  for (int wb = 0; wb < 8; wb++) {
    k<<<1, 256>>>(0, wb, out, cyc);
    cudaDeviceSynchronize();
    printf("warps 0 and %d: %lld cycles\n", wb, cyc[0]);
  }
  // two CTAs of 128 threads on one SM? launch 2 CTAs x 128 with 100KB smem each so both fit one SM is not controllable; skip
  return 0;
}

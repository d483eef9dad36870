// fp64_probe.cu -- micro-benchmarks behind the design of the B/F kernels (B200, sm_100a):
//   1. dependent DFMA latency, 2. DFMA throughput of one SM sub-partition vs. warps x independent chains,
//   3. shared-memory load latency, 4. DFMA + LDS mix, 5. 1-D bulk-copy (TMA) streaming rate into shared memory.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/cu/bin/fp64_probe tools/cu/fp64_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <stdint.h>

template <int CH>
__global__ void dfma_chain(double *out, long long *cyc, int iters, int warps_mask_stride) {
  // warps w with (w % 4 == 0) sit on sub-partition 0; only those run when warps_mask_stride == 4
  const int w = threadIdx.x >> 5;
  if (w % warps_mask_stride != 0) return;
  double a[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) a[i] = threadIdx.x + i;
  const double m = 1.0000001, c = 1e-9;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < CH; i++) a[i] = fma(a[i], m, c);
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * 32 + w] = t1 - t0;
}

__global__ void lds_latency(double *out, long long *cyc, int iters) {
  __shared__ int nxt[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i + 33) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) p = nxt[p];
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

// each of CH chains: one LDS.64 then NF dependent DFMAs on the loaded value (like the correlation build)
template <int CH, int NF>
__global__ void lds_dfma_mix(double *out, long long *cyc, int iters, int stride) {
  extern __shared__ double sm[];
  const int w = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = 1.0 + 1e-9 * i;
  __syncthreads();
  if (w % stride != 0) return;
  double acc[CH];
#pragma unroll
  for (int i = 0; i < CH; i++) acc[i] = 0.0;
  const int lane = threadIdx.x & 31;
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
    double v[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) v[i] = sm[((it * CH + i) * 32 + lane) & 4095];
#pragma unroll
    for (int f = 0; f < NF; f++) {
#pragma unroll
      for (int i = 0; i < CH; i++) v[i] = fma(v[i], 1.0000001, 1e-9);
    }
#pragma unroll
    for (int i = 0; i < CH; i++) acc[i] += v[i];
  }
  long long t1 = clock64();
  double s = 0;
#pragma unroll
  for (int i = 0; i < CH; i++) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (lane == 0) cyc[blockIdx.x * 32 + w] = t1 - t0;
}

// ---- bulk copy streaming: every CTA pulls `chunk` bytes per stage through an mbarrier ring --------------------
__device__ __forceinline__ void mbar_init(uint64_t *b, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *b, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"((unsigned)__cvta_generic_to_shared(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *b, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra DONE_%=;\n bra WAIT_%=;\n DONE_%=:\n}\n" ::"r"(
          (unsigned)__cvta_generic_to_shared(b)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, uint64_t *b) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   (unsigned)__cvta_generic_to_shared(dst)),
               "l"(src), "r"(bytes), "r"((unsigned)__cvta_generic_to_shared(b))
               : "memory");
}

template <int STAGES>
__global__ void bulk_stream(const char *src, size_t total_bytes, unsigned chunk, double *out) {
  extern __shared__ __align__(128) char sbuf[];
  __shared__ uint64_t full[STAGES];
  const size_t nchunks = total_bytes / chunk;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; s++) mbar_init(&full[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  double acc = 0.0;
  // chunk c of this CTA: blockIdx.x + c * gridDim.x
  size_t my = (nchunks > blockIdx.x) ? (nchunks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES && (size_t)s < my; s++) {
      mbar_expect_tx(&full[s], chunk);
      bulk_g2s(sbuf + (size_t)s * chunk, src + (blockIdx.x + (size_t)s * gridDim.x) * chunk, chunk, &full[s]);
    }
  }
  for (size_t c = 0; c < my; c++) {
    const int s = c % STAGES;
    mbar_wait(&full[s], (c / STAGES) & 1);
    const double *d = reinterpret_cast<const double *>(sbuf + (size_t)s * chunk);
    for (unsigned i = threadIdx.x; i < chunk / 8; i += blockDim.x * 8) acc += d[i];  // touch 1/8 of it
    __syncthreads();
    if (threadIdx.x == 0 && c + STAGES < my) {
      mbar_expect_tx(&full[s], chunk);
      bulk_g2s(sbuf + (size_t)s * chunk, src + (blockIdx.x + (c + STAGES) * gridDim.x) * chunk, chunk, &full[s]);
    }
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

__global__ void __launch_bounds__(384, 1) setmaxnreg_probe(int *out) {
  const int wg = threadIdx.x >> 7;
  if (wg == 0) {
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
  } else {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 104;");
  }
  if (threadIdx.x == 0) out[0] = 1;
}

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <int CH>
static void run_chain(int warps_on_smsp, double *out, long long *cyc) {
  const int iters = 20000;
  // block of 4 * warps_on_smsp warps, only warps with w % 4 == 0 active -> all on sub-partition 0
  dfma_chain<CH><<<1, 128 * warps_on_smsp>>>(out, cyc, iters, 4);
  CK(cudaDeviceSynchronize());
  long long mx = 0;
  for (int w = 0; w < 4 * warps_on_smsp; w += 4) mx = cyc[w] > mx ? cyc[w] : mx;
  const double per = (double)mx / iters;  // cycles per trip (CH DFMAs per warp)
  printf("  warps/SMSP %d chains %2d : %7.2f cyc/trip  -> %5.3f DFMA warp-inst/cyc/SMSP\n", warps_on_smsp, CH, per,
         warps_on_smsp * CH / per);
}

template <int CH, int NF>
static void run_mix(int warps_on_smsp, double *out, long long *cyc) {
  const int iters = 5000;
  lds_dfma_mix<CH, NF><<<1, 128 * warps_on_smsp, 4096 * 8>>>(out, cyc, iters, 4);
  CK(cudaDeviceSynchronize());
  long long mx = 0;
  for (int w = 0; w < 4 * warps_on_smsp; w += 4) mx = cyc[w] > mx ? cyc[w] : mx;
  const double per = (double)mx / iters;
  printf("  mix warps/SMSP %d chains %d x (1 LDS + %2d dep DFMA + 1 DADD): %7.2f cyc/trip -> FP64 %5.3f inst/cyc/SMSP\n",
         warps_on_smsp, CH, NF, per, warps_on_smsp * CH * (NF + 1) / per);
}

int main() {
  double *out;
  long long *cyc;
  CK(cudaMalloc(&out, 1 << 22));
  CK(cudaMallocManaged(&cyc, 4096 * 8));
  printf("== dependent DFMA chains on ONE sub-partition ==\n");
  run_chain<1>(1, out, cyc);
  run_chain<2>(1, out, cyc);
  run_chain<4>(1, out, cyc);
  run_chain<8>(1, out, cyc);
  run_chain<1>(2, out, cyc);
  run_chain<2>(2, out, cyc);
  run_chain<4>(2, out, cyc);
  run_chain<1>(3, out, cyc);
  run_chain<2>(3, out, cyc);
  run_chain<4>(3, out, cyc);
  run_chain<1>(4, out, cyc);
  run_chain<2>(4, out, cyc);
  run_chain<1>(8, out, cyc);
  printf("== shared-memory pointer chase ==\n");
  lds_latency<<<1, 32>>>(out, cyc, 20000);
  CK(cudaDeviceSynchronize());
  printf("  LDS dependent latency: %.1f cycles\n", (double)cyc[0] / 20000);
  printf("== LDS + DFMA mix on one sub-partition ==\n");
  run_mix<4, 8>(1, out, cyc);
  run_mix<8, 8>(1, out, cyc);
  run_mix<4, 8>(2, out, cyc);
  run_mix<8, 8>(2, out, cyc);
  run_mix<4, 8>(3, out, cyc);
  run_mix<8, 8>(3, out, cyc);
  run_mix<4, 20>(1, out, cyc);
  run_mix<8, 20>(1, out, cyc);
  run_mix<4, 20>(2, out, cyc);
  run_mix<8, 20>(3, out, cyc);
  printf("== setmaxnreg ==\n");
  int *flag;
  CK(cudaMallocManaged(&flag, 4));
  *flag = 0;
  setmaxnreg_probe<<<1, 384>>>(flag);
  cudaError_t e = cudaDeviceSynchronize();
  printf("  setmaxnreg inc 232 / dec 104 on a 384-thread CTA: %s (flag %d)\n", cudaGetErrorString(e), *flag);
  printf("== 1-D bulk copy (TMA) streaming global -> shared ==\n");
  const size_t total = (size_t)2 << 30;
  char *src;
  CK(cudaMalloc(&src, total));
  CK(cudaMemset(src, 0, total));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (unsigned chunk : {8192u, 16384u, 30720u}) {
    for (int ctas_per_sm : {1, 2}) {
      const int stages = 4;
      const size_t smem = (size_t)stages * chunk;
      CK(cudaFuncSetAttribute(bulk_stream<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      const size_t use = total / chunk * chunk;
      bulk_stream<4><<<148 * ctas_per_sm, 128, smem>>>(src, use, chunk, out);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      bulk_stream<4><<<148 * ctas_per_sm, 128, smem>>>(src, use, chunk, out);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      printf("  chunk %6u B, %d CTA/SM, 4 stages: %.3f ms -> %.0f GB/s\n", chunk, ctas_per_sm, ms, use / ms * 1e-6);
    }
  }
  return 0;
}

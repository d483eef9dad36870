// stat_kernels.cu -- one-time statistics of a data set that steer kernel choices (nngp_set_data).
//
// pair_d2_hist_kernel: for every location with a full neighbour row, the largest SQUARED distance between any two
// of its m + 1 sites (the arguments updateBF will feed the correlation function, src/rNNGP.cpp:76-84), binned by the
// top 16 bits of the double (monotone in the value: 16 bins per octave).  The host reads the 99 % quantile off the
// histogram; the Matern launch uses its short series when phi^2/4 times that bound stays inside the series' range
// (matern_series.cuh).  ~0.1 ms at n = 10^6, once per nngp_set_data.
#include "nngp_common.cuh"

namespace nngp {

constexpr int kPairHistBins = 0x8000;  // positive doubles: sign bit clear

__global__ void __launch_bounds__(128) pair_d2_hist_kernel(const double2 *__restrict__ coords,
                                                           const int *__restrict__ nnIndx, int m, int loc_begin,
                                                           int loc_end, unsigned int *__restrict__ hist) {
  const long long i = (long long)loc_begin + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= loc_end || i < m) return;
  double2 s[NNGP_MAX_M + 1];
  const long long off = nn_offset(i, m);
  for (int k = 0; k < m; ++k) s[k] = __ldg(&coords[nnIndx[off + k]]);
  s[m] = __ldg(&coords[i]);
  double mx = 0.0;
  for (int k = 1; k <= m; ++k)
    for (int l = 0; l < k; ++l) {
      const double dx = s[k].x - s[l].x, dy = s[k].y - s[l].y;
      mx = fmax(mx, dx * dx + dy * dy);
    }
  atomicAdd(&hist[(unsigned)((unsigned long long)__double_as_longlong(mx) >> 48) & (kPairHistBins - 1)], 1u);
}

int pair_d2_hist_bins() { return kPairHistBins; }

cudaError_t pair_d2_hist_launch(const double2 *coords, const int *nnIndx, int m, int loc_begin, int loc_end,
                                unsigned int *hist, cudaStream_t stream) {
  cudaError_t e = cudaMemsetAsync(hist, 0, kPairHistBins * sizeof(unsigned int), stream);
  if (e != cudaSuccess) return e;
  const long long nloc = (long long)loc_end - loc_begin;
  if (nloc <= 0) return cudaSuccess;
  pair_d2_hist_kernel<<<(unsigned)((nloc + 127) / 128), 128, 0, stream>>>(coords, nnIndx, m, loc_begin, loc_end, hist);
  return cudaGetLastError();
}

}  // namespace nngp

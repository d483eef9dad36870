// predict_kernels.cu -- epilogues of the kNN + solve prediction paths.  The m x m kriging solves run in
// the B/F kernels' site mode (bf_tpl_kernel.cu / bf_kernel.cu); these kernels turn the weights w into
// predictions.
//
// Replaces (reference file:line):
//   rNNGPPredict  src/rNNGPPredict.cpp:171-181  y0 = x0'b + sum_k w_k (y_Nk - x_Nk'b) + sqrt(s2+t2-w'c) z
//   cNNGP         src/cNNGP.cpp:278-298         y0.hat = x0'bh + w'(y_N - X_N bh),
//                                               var = b (u'Vu + 1 + alpha - w'c0)/(a - 1), u = x0 - X_N'w
// Both use the same per-site reductions  a_i = w . y_N  and  u_i = x0_i - X_N' w, so that a prediction for
// ANY beta is u_i . beta + a_i: the solve is done once per distinct theta, not once per posterior sample.
#include "nngp_common.cuh"

namespace nngp {

// thread per site; W, nn0, X0, U are column-major over the cs sites of the chunk (coalesced)
__global__ void __launch_bounds__(256) site_reduce_kernel(const double *__restrict__ W, const int *__restrict__ nn0,
                                                          const double *__restrict__ X, const double *__restrict__ y,
                                                          const double *__restrict__ X0, int n, int p, int m, int cs,
                                                          double *__restrict__ a, double *__restrict__ U) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cs) return;
  double acc = 0.0;
  for (int k = 0; k < m; ++k) acc = fma(W[(size_t)k * cs + i], __ldg(&y[nn0[(size_t)k * cs + i]]), acc);
  a[i] = acc;
  for (int c = 0; c < p; ++c) {
    const double *__restrict__ Xc = X + (size_t)c * n;
    double t = 0.0;
    for (int k = 0; k < m; ++k) t = fma(W[(size_t)k * cs + i], __ldg(&Xc[nn0[(size_t)k * cs + i]]), t);
    U[(size_t)c * cs + i] = X0[(size_t)c * cs + i] - t;
  }
}

// y0c[s_local... ] : for the samples listed in `samples` (all sharing the theta of this group)
// y0c is S x cs (sample-major, coalesced over sites); z is cs x S (site-major, reference order)
__global__ void __launch_bounds__(256) predict_emit_kernel(const double *__restrict__ a, const double *__restrict__ U,
                                                           const double *__restrict__ F, const double *__restrict__ beta,
                                                           const double *__restrict__ z, const int *__restrict__ samples,
                                                           int nsamp, int p, int cs, int S, double *__restrict__ y0c) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cs) return;
  const double sd = sqrt(F[i]);
  const double ai = a[i];
  for (int j = blockIdx.y; j < nsamp; j += gridDim.y) {
    const int s = samples[j];
    double v = ai;
    for (int c = 0; c < p; ++c) v = fma(U[(size_t)c * cs + i], __ldg(&beta[(size_t)s * p + c]), v);
    y0c[(size_t)s * cs + i] = fma(sd, __ldg(&z[(size_t)i * S + s]), v);
  }
}

// conjugate model: point prediction and its variance for one parameter set
__global__ void __launch_bounds__(256) conj_emit_kernel(const double *__restrict__ a, const double *__restrict__ U,
                                                        const double *__restrict__ F, const double *__restrict__ betaHat,
                                                        const double *__restrict__ V /* p x p full symmetric */, int p,
                                                        int cs, double bIG, double aIG, double *__restrict__ y0Hat,
                                                        double *__restrict__ y0HatVar) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= cs) return;
  double v = a[i];
  for (int c = 0; c < p; ++c) v = fma(U[(size_t)c * cs + i], __ldg(&betaHat[c]), v);
  y0Hat[i] = v;
  double quad = 0.0;
  for (int c = 0; c < p; ++c) {
    double t = 0.0;
    for (int d = 0; d < p; ++d) t = fma(__ldg(&V[(size_t)d * p + c]), U[(size_t)d * cs + i], t);
    quad = fma(U[(size_t)c * cs + i], t, quad);
  }
  y0HatVar[i] = bIG * (quad + F[i]) / (aIG - 1.0);
}

// Hold-out scores of one parameter set (R/util.R:247-255, summed in R/spConjNNGP.R:330-334):
//   crps  = mean(-sd (1/sqrt(pi) - 2 dnorm(z) - z (2 pnorm(z) - 1))),  sd = sqrt(var), z = (y - yhat)/sd
//   rmspe = sqrt(mean((y - yhat)^2))
// This kernel leaves the two SUMS over a chunk of sites in out[0..1] (added to what is there, so chunks
// accumulate); the caller divides by n0 and takes the root.  Fixed grid, per-block partials added by the last
// block in block order: run-to-run deterministic.
constexpr int kScoreBlocks = 148;
__global__ void __launch_bounds__(256) conj_score_kernel(const double *__restrict__ y0, const double *__restrict__ y0Hat,
                                                         const double *__restrict__ y0HatVar, int cs,
                                                         double *__restrict__ partials, unsigned int *__restrict__ ticket,
                                                         double *__restrict__ out) {
  __shared__ double red[2][8];
  __shared__ int last;
  double sc = 0.0, se = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < cs; i += gridDim.x * blockDim.x) {
    const double e = y0[i] - y0Hat[i];
    const double sd = sqrt(y0HatVar[i]);
    const double z = e / sd;
    const double dn = 0.3989422804014326779 * exp(-0.5 * z * z);
    sc += -sd * (0.5641895835477562869 - 2.0 * dn - z * (2.0 * normcdf(z) - 1.0));
    se = fma(e, e, se);
  }
#pragma unroll
  for (int sft = 16; sft >= 1; sft >>= 1) {
    sc += __shfl_down_sync(0xffffffffu, sc, sft);
    se += __shfl_down_sync(0xffffffffu, se, sft);
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) {
    red[0][warp] = sc;
    red[1][warp] = se;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double a = 0.0, b = 0.0;
    for (int w = 0; w < 8; ++w) {
      a += red[0][w];
      b += red[1][w];
    }
    partials[2 * blockIdx.x] = a;
    partials[2 * blockIdx.x + 1] = b;
    __threadfence();
    last = (atomicAdd(ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (last && threadIdx.x == 0) {
    __threadfence();
    double a = 0.0, b = 0.0;
    for (unsigned k = 0; k < gridDim.x; ++k) {
      a += __ldcg(&partials[2 * k]);
      b += __ldcg(&partials[2 * k + 1]);
    }
    out[0] += a;
    out[1] += b;
    *ticket = 0u;
  }
}

cudaError_t site_reduce_launch(const double *W, const int *nn0, const double *X, const double *y, const double *X0,
                               int n, int p, int m, int cs, double *a, double *U, cudaStream_t stream) {
  site_reduce_kernel<<<(cs + 255) / 256, 256, 0, stream>>>(W, nn0, X, y, X0, n, p, m, cs, a, U);
  return cudaGetLastError();
}

cudaError_t predict_emit_launch(const double *a, const double *U, const double *F, const double *beta, const double *z,
                                const int *samples, int nsamp, int p, int cs, int S, double *y0c, cudaStream_t stream) {
  dim3 grid((cs + 255) / 256, nsamp < 64 ? nsamp : 64, 1);
  predict_emit_kernel<<<grid, 256, 0, stream>>>(a, U, F, beta, z, samples, nsamp, p, cs, S, y0c);
  return cudaGetLastError();
}

cudaError_t conj_emit_launch(const double *a, const double *U, const double *F, const double *betaHat, const double *V,
                             int p, int cs, double bIG, double aIG, double *y0Hat, double *y0HatVar,
                             cudaStream_t stream) {
  conj_emit_kernel<<<(cs + 255) / 256, 256, 0, stream>>>(a, U, F, betaHat, V, p, cs, bIG, aIG, y0Hat, y0HatVar);
  return cudaGetLastError();
}

// scratch: 2 * kScoreBlocks doubles of partials followed by one zeroed unsigned ticket (see conj_score_scratch_bytes)
size_t conj_score_scratch_bytes() { return 2 * kScoreBlocks * sizeof(double) + 16; }

cudaError_t conj_score_launch(const double *y0, const double *y0Hat, const double *y0HatVar, int cs, void *scratch,
                              double *out, cudaStream_t stream) {
  double *partials = reinterpret_cast<double *>(scratch);
  unsigned int *ticket = reinterpret_cast<unsigned int *>(partials + 2 * kScoreBlocks);
  conj_score_kernel<<<kScoreBlocks, 256, 0, stream>>>(y0, y0Hat, y0HatVar, cs, partials, ticket, out);
  return cudaGetLastError();
}

}  // namespace nngp

// tri_cache.cu -- the resident DISTANCE TRIANGLE of a data set (built once by nngp_set_data, streamed by
// bf_tri_kernel.cu on every likelihood evaluation).
//
// For location i with neighbours N(i,0..m_i-1) the reference evaluates, on every call of updateBF, the m distances
// location <-> neighbour (src/rNNGP.cpp:76-77) and the m(m-1)/2 distances between neighbours (:78-84) with
//   dist2(a1,a2,b1,b2) = sqrt(pow(a1-b1,2) + pow(a2-b2,2))            src/util.cpp:51-53
// They depend on the geometry only.  Layout: tiles of 32 consecutive locations of the shard; tile t holds
// R = m(m+1)/2 rows of 32 doubles (lane = location in the tile): rows 0..m-1 = D(i, N(i,k)), rows m + k(k-1)/2 + l =
// D(N(i,k), N(i,l)), l < k.  One tile is one contiguous block of R x 256 bytes -- the unit of the kernel's bulk copies.
// Missing neighbours of the ragged head (k >= m_i) get D = 0 (the solver pads those rows with the identity); the lanes
// beyond loc_end in the last tile replicate the last location.  Arithmetic: __dsub_rn / __dmul_rn / __dadd_rn /
// __dsqrt_rn -- no contraction, IEEE sqrt -- so every stored distance has the bits the reference's dist2 produces.
// Behind the distances the same allocation holds the neighbour ids TILE-MAJOR (tri_ids_offset): tile t has m rows of 32
// ints, row k = N(i, k) of the tile's 32 locations, -1 where a location has fewer than k + 1 neighbours -- what the
// kernel's loader warp reads coalesced (lane = location) to gather the residuals.
// HBM-write bound: 960 + 60 bytes per location at m = 15 (n = 10^6: 1.02 GB).
#include "nngp_common.cuh"

namespace nngp {

size_t tri_ids_offset(int m, long long nloc) {  // doubles in front of the tile-major ids
  const long long tiles = (nloc + 31) / 32;
  return (size_t)tiles * (size_t)(m * (m + 1) / 2) * 32;
}
size_t tri_cache_doubles(int m, long long nloc) {
  const long long tiles = (nloc + 31) / 32;
  return tri_ids_offset(m, nloc) + ((size_t)tiles * m * 32 + 1) / 2;
}

constexpr int kTriWarps = 2;

__global__ void __launch_bounds__(32 * kTriWarps) tri_build_kernel(const double2 *__restrict__ coords,
                                                                  const int *__restrict__ nnIndx, int m, int loc_begin,
                                                                  int loc_end, double *__restrict__ tri, int *__restrict__ nnT,
                                                                  long long ntiles) {
  extern __shared__ double2 tri_co[];  // [warp][k][lane], k <= m (row m: the location itself)
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  double2 *co = tri_co + (size_t)w * (m + 1) * 32;
  const int rows = m * (m + 1) / 2;
  for (long long tile = (long long)blockIdx.x * kTriWarps + w; tile < ntiles; tile += (long long)gridDim.x * kTriWarps) {
    long long i = (long long)loc_begin + tile * 32 + lane;
    if (i >= loc_end) i = (long long)loc_end - 1;
    const int mi = nn_count(i, m);
    const long long off = nn_offset(i, m);
    const double2 own = __ldg(&coords[i]);
    for (int k = 0; k < m; ++k) {
      const int id = (k < mi) ? __ldg(&nnIndx[off + k]) : -1;
      nnT[(size_t)tile * m * 32 + k * 32 + lane] = id;
      co[k * 32 + lane] = (id >= 0) ? __ldg(&coords[id]) : own;
    }
    co[m * 32 + lane] = own;
    __syncwarp();
    double *dst = tri + (size_t)tile * rows * 32 + lane;
    for (int k = 0; k < m; ++k) {
      const double2 a = co[k * 32 + lane];
      const double dx = __dsub_rn(own.x, a.x), dy = __dsub_rn(own.y, a.y);
      const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
      dst[(size_t)k * 32] = (k < mi) ? d : 0.0;
    }
    for (int k = 1; k < m; ++k) {
      const double2 a = co[k * 32 + lane];
      for (int l = 0; l < k; ++l) {
        const double2 b = co[l * 32 + lane];
        const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y);
        const double d = __dsqrt_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)));
        dst[(size_t)(m + k * (k - 1) / 2 + l) * 32] = (k < mi) ? d : 0.0;
      }
    }
    __syncwarp();
  }
}

cudaError_t tri_build_launch(const double2 *coords, const int *nnIndx, int m, int loc_begin, int loc_end, double *tri,
                             cudaStream_t stream) {
  const long long nloc = (long long)loc_end - loc_begin;
  if (nloc <= 0) return cudaSuccess;
  const long long ntiles = (nloc + 31) / 32;
  const size_t smem = (size_t)kTriWarps * (m + 1) * 32 * sizeof(double2);
  long long grid = (ntiles + kTriWarps - 1) / kTriWarps;
  if (grid > 148 * 32) grid = 148 * 32;
  int *nnT = reinterpret_cast<int *>(tri + tri_ids_offset(m, nloc));
  tri_build_kernel<<<(unsigned)grid, 32 * kTriWarps, smem, stream>>>(coords, nnIndx, m, loc_begin, loc_end, tri, nnT, ntiles);
  return cudaGetLastError();
}

}  // namespace nngp

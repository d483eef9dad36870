// bf_tri_kernel.cu -- B/F + log-likelihood kernel over the resident DISTANCE TRIANGLE (2 <= m <= 32).
//
// Same mathematics as the other B/F kernels (reference: updateBF src/rNNGP.cpp:46-101, updateConjBF
// src/cNNGP.cpp:24-69, Q src/util.cpp:261-280), different data path.  The distances between a location's
// m + 1 sites -- the arguments of spCor (src/util.cpp:208-245; dist2 src/util.cpp:51-53) -- do not depend on
// theta, and a sampler evaluates the same geometry thousands of times.  nngp_set_data therefore builds them
// once (tri_cache.cu: m(m+1)/2 doubles per location, tile-major, bit-equal to the reference's dist2) and this
// kernel STREAMS them: on B200 the FP64 pipe (34 TFLOP/s) is the scarce resource and HBM (6.5 TB/s) is not --
// recomputing a distance costs 15 FP64 instructions (gather, d^2, sqrt), reading it costs 8 bytes.
//
//   n = 10^6, m = 15:  0.96 GB of distances + 0.27 GB of ids / weights / variances per launch (HBM floor 0.19 ms),
//                      ~3,700 (exponential) / ~7,700 (Matern) FP64 warp instructions per 32-location tile
//                      (FP64 floor 0.20 / 0.42 ms).
//
// One persistent CTA per SM, 12 warps in three warpgroups with `setmaxnreg`:
//   warpgroup 0 (232 registers): warps 0..NC-1 CONSUMERS -- one tile each: ragged-row fix-up, blocked Cholesky +
//                                solves (tri_solve below), log F, the quadratic-form terms, weights out through a
//                                bulk store;
//                                warp 3 LOADER -- claims tiles (one atomic per tile), issues the 1-D bulk copies
//                                (TMA) of the tile's distances [and of the resident weights of the current theta]
//                                and the cp.async gathers of the neighbours' residuals
//   warpgroups 1, 2 (136 registers; 128 x 232 + 256 x 136 = the 384 x 168 the CTA was launched with):
//                                8 PRODUCER warps -- turn the stage's distances into covariances in place,
//                                sigma^2 rho(D): rows of 32 doubles, NU independent chains per batch
// Stages of one tile each travel  loader -> producers -> consumer -> loader  on mbarriers (full / built / empty).
// The geometry (rows per stage, number of stages S, consumers NC) is a run-time function of m (tri_geometry): one
// instantiation per correlation model serves every m; nothing is padded.
// Per-tile partial sums are written to a tile-indexed array and added in tile order by the last CTA, so the result
// does not depend on which CTA drew which tile.
//
// The solve is a BLOCKED Cholesky with rolled loops (4 x 4 register tiles over the per-lane matrix in shared memory),
// not the fully unrolled column sweep of tpl_device.cuh: ncu showed the unrolled solve (36 KB of straight-line code
// per consumer, executed once per tile) stalled on instruction fetch for 55 % of its samples; the rolled form is
// ~700 instructions.
#include <stdlib.h>

#include "tpl_device.cuh"

namespace nngp {

#ifndef NNGP_TRI_UNR
#define NNGP_TRI_UNR 5  // rows (independent chains) per producer batch
#endif
#ifndef NNGP_TRI_UNR_MATERN
#define NNGP_TRI_UNR_MATERN 4
#endif
#ifndef NNGP_TRI_SIG_FOLD
#define NNGP_TRI_SIG_FOLD 1  // Matern: sigma^2 multiplied into the series coefficients on the host (one FP64 instruction per entry)
#endif
#ifndef NNGP_TRI_MATERN_COST
#define NNGP_TRI_MATERN_COST 52  // FP64 instructions per Matern correlation (9-term series), for the producers' shares
#endif

constexpr int kTriMaxM = 32;       // 27..32: ONE stage fits shared memory (loader, producers and the consumer take turns)
constexpr int kTriNarrowMaxM = 19;  // largest m with two or more consumers (tri_geometry): the narrow instantiation
constexpr int kTriMaxStages = 6;
constexpr int kTriThreads = 384;
constexpr int kTriProducers = 8;  // at most (seven when the loader takes warp 4)
// Register split of the CTA's pool (384 x 168): consumers / producers 232 / 136, or "slim" 200 / 152 for models whose
// producers are the critical path and want one more evaluation chain per batch (NNGP_TRI_SLIM=1 builds: Matern with five
// chains at 152 registers and a 3-column consumer at 200 measured 0.665 vs 0.648 ms -- not taken)
#ifndef NNGP_TRI_SLIM
#define NNGP_TRI_SLIM 0
#endif
template <int COV>
constexpr bool kTriSlim = NNGP_TRI_SLIM && kIsMatern<COV>;
template <int COV>
constexpr int kTriConsumerWeight = 10;  // tenths; see the shares in the kernel's set-up.  Measured (n = 10^6, m = 15, ms,
                                        // exponential / spherical / Matern): 0 (even split) 0.473 / 0.432 / 0.667; 10: 0.467 / 0.414 /
                                        // 0.647; 18: 0.501 / 0.445 / 0.693; 28: 0.597 / 0.447 / 0.854 (scheduler 3's two warps pace the tile)
template <int COV>
constexpr bool kTriWant4 = !kIsMatern<COV>;  // Matern keeps three consumers and eight producers
constexpr int kTriTabBytes = (32 + 128) * 8;
constexpr int kTriBarBytes = 512;  // 4 x 6 mbarriers, then the tile ring (6 long long)

// Shared-memory geometry of the kernel for a given m (bytes; every region is a whole number of 256-byte rows).
struct TriGeom {
  int off_cv, off_a, off_rr, stage_bytes;  // within a stage: c -> y -> weights | packed strict lower triangle |
                                           // residuals of the neighbours, own in row m
  int cons_bytes;                          // per consumer: 1/diag rows during the solve, then the weight staging [lane][k]
  int S, NC;
  int off_cons, off_tab, off_bar, smem_bytes;
};
// want4: four consumers when six stages fit (models whose consumers are the critical path); the loader then moves to
// warp 4 and seven producer warps remain.
__host__ __device__ inline TriGeom tri_geometry(int m, bool want4) {
  TriGeom g;
  const int ne = m * (m - 1) / 2;
  g.off_cv = 0;
  g.off_a = m * 256;
  g.off_rr = g.off_a + ne * 256;
  g.stage_bytes = g.off_rr + (m + 1) * 256;
  g.cons_bytes = m * 256;
  const int avail = 227 * 1024 - kTriTabBytes - kTriBarBytes - 1024;  // 1 KB: the kernel's static shared memory
  g.NC = 4;
  g.S = (avail - 4 * g.cons_bytes) / g.stage_bytes;
  if (!want4 || g.S < 6) {
    g.NC = 3;
    g.S = (avail - 3 * g.cons_bytes) / g.stage_bytes;
    if (g.S < 5) {
      g.NC = 2;
      g.S = (avail - 2 * g.cons_bytes) / g.stage_bytes;
      if (g.S < 4) {
        g.NC = 1;
        g.S = (avail - g.cons_bytes) / g.stage_bytes;
      }
    }
  }
  if (g.S > kTriMaxStages) g.S = kTriMaxStages;
  g.off_cons = g.S * g.stage_bytes;
  g.off_tab = g.off_cons + g.NC * g.cons_bytes;
  g.off_bar = g.off_tab + kTriTabBytes;
  g.smem_bytes = g.off_bar + kTriBarBytes;
  return g;
}

// ---- mbarrier / bulk-copy primitives (PTX) ---------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n .reg .pred p;\n TRI_WAIT_%=:\n mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n @p bra TRI_DONE_%=;\n"
      " bra TRI_WAIT_%=;\n TRI_DONE_%=:\n}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// arrive on `bar` once all cp.async of this thread issued so far have landed (counts as one pending arrival)
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(bar) : "memory");
}
// 8-byte cp.async that writes zeros when `valid` is false (src-size 0: nothing is read)
__device__ __forceinline__ void cp_async8_zfill(void *smem_dst, const void *gsrc, bool valid) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gsrc), "r"(valid ? 8 : 0) : "memory");
}
__device__ __forceinline__ void bulk_g2s(unsigned dst, const void *src, unsigned bytes, unsigned bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst), "l"(src),
               "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void *dst, unsigned src, unsigned bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst), "r"(src), "r"(bytes) : "memory");
}
// pulls a contiguous block into L2 ahead of its first use (no destination, no completion to wait for)
__device__ __forceinline__ void bulk_prefetch_l2(const void *src, unsigned bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// tile of 32 complete (m-neighbour) rows: its weights are one contiguous run of 32 m doubles in B
__device__ __forceinline__ bool tri_tile_is_full(long long base, int m, long long loc_end) {
  return base >= m && base + 32 <= loc_end;
}

// ---- sigma^2 rho(D) from the DISTANCE --------------------------------------------------------------
template <int COV>
struct FromDist;
template <>
struct FromDist<NNGP_COV_EXPONENTIAL> {
  double nphi, sig;
  __device__ __forceinline__ FromDist(const BfTheta &t, const double *) : nphi(-t.phi), sig(t.sigmaSq) {}
  __device__ __forceinline__ double operator()(double D) const { return sig * nngp_exp_nonpos(nphi * D); }
};
template <>
struct FromDist<NNGP_COV_GAUSSIAN> {
  double nphi2, sig;
  __device__ __forceinline__ FromDist(const BfTheta &t, const double *) : nphi2(-(t.phi * t.phi)), sig(t.sigmaSq) {}
  __device__ __forceinline__ double operator()(double D) const { return sig * nngp_exp_nonpos(nphi2 * (D * D)); }
};
// the same with the table-driven exp (kExpTab32 in shared memory, fastmath.cuh: 12 FP64 instructions instead of 19): used
// where the producers' phase is not hidden behind the consumers (team geometries)
template <int COV>
struct FromDistTab;
template <>
struct FromDistTab<NNGP_COV_EXPONENTIAL> {
  double nphi, sig;
  const double *tab;
  __device__ __forceinline__ FromDistTab(const BfTheta &t, const double *tab_) : nphi(-t.phi), sig(t.sigmaSq), tab(tab_) {}
  __device__ __forceinline__ double operator()(double D) const { return sig * nngp_exp_tab(nphi * D, tab); }
};
template <>
struct FromDistTab<NNGP_COV_GAUSSIAN> {
  double nphi2, sig;
  const double *tab;
  __device__ __forceinline__ FromDistTab(const BfTheta &t, const double *tab_) : nphi2(-(t.phi * t.phi)), sig(t.sigmaSq), tab(tab_) {}
  __device__ __forceinline__ double operator()(double D) const { return sig * nngp_exp_tab(nphi2 * (D * D), tab); }
};
template <>
struct FromDist<NNGP_COV_SPHERICAL> {
  double phi, rng, sig;
  __device__ __forceinline__ FromDist(const BfTheta &t, const double *) : phi(t.phi), rng(1.0 / t.phi), sig(t.sigmaSq) {}
  __device__ __forceinline__ double operator()(double D) const {  // src/util.cpp:221-227; D = 0 gives exactly 1
    const double pd = phi * D;
    const double v = 1.0 - 1.5 * pd + 0.5 * (pd * pd * pd);
    return (D <= rng) ? sig * v : 0.0;
  }
};

// One batch of NU rows of 32 doubles (this lane's element of each): all loads, NU independent chains, all stores.
template <int COV, int NU, typename F>
__device__ __forceinline__ void tri_rows(unsigned char *const (&p)[NU], const F &f) {
  double v[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) v[u] = *reinterpret_cast<const double *>(p[u]);
#pragma unroll
  for (int u = 0; u < NU; ++u) v[u] = f(v[u]);
#pragma unroll
  for (int u = 0; u < NU; ++u) *reinterpret_cast<double *>(p[u]) = v[u];
}
// Matern: the series in t = (phi D / 2)^2 (matern_series.cuh) with the table-driven log / exp; the rare arguments
// outside the straight-line path are re-evaluated afterwards, like the gather kernels do (tpl_device.cuh)
template <int COV, int NU>
__device__ __forceinline__ void tri_rows_matern(unsigned char *const (&p)[NU], const Cor<COV> &cor, double sig) {
  double d2[NU], v[NU];
#pragma unroll
  for (int u = 0; u < NU; ++u) {
    const double D = *reinterpret_cast<const double *>(p[u]);
    d2[u] = D * D;
  }
#pragma unroll
  for (int u = 0; u < NU; ++u) v[u] = NNGP_TRI_SIG_FOLD ? cor.fast_tab(d2[u]) : sig * cor.fast_tab(d2[u]);  // folded: sigma^2 rides in the series coefficients (tri_launch_one)
  bool any = false;
#pragma unroll
  for (int u = 0; u < NU; ++u) any |= cor.maybe_slow(d2[u]);
  if (any) {
#pragma unroll
    for (int u = 0; u < NU; ++u)
      if (cor.needs_slow(d2[u])) v[u] = sig * cor.slow(d2[u]);
  }
#pragma unroll
  for (int u = 0; u < NU; ++u) *reinterpret_cast<double *>(p[u]) = v[u];
}

template <int COV>
constexpr int kTriUnr = kIsMatern<COV> ? (NNGP_TRI_SLIM ? NNGP_TRI_UNR_MATERN + 1 : NNGP_TRI_UNR_MATERN) : NNGP_TRI_UNR;
// FP64 instructions per correlation (SASS counts), for the producers' shares
template <int COV>
constexpr int kTriRowCost = COV == NNGP_COV_MATERN_TINY ? NNGP_TRI_MATERN_COST - 8 : (kIsMatern<COV> ? NNGP_TRI_MATERN_COST : (COV == NNGP_COV_SPHERICAL ? 7 : 21));

// Whole batches [b0, b1) of the stage's m(m+1)/2 rows (c rows, then the packed triangle: contiguous);
// `tail`: this warp also takes the rows left over after the last whole batch.
template <int COV, typename Fn1, typename FnN>
__device__ __forceinline__ void tri_transform(unsigned char *stage, int lane, int b0, int b1, bool tail, int m,
                                              const Fn1 &one, const FnN &many) {
  constexpr int NU = kTriUnr<COV>;
  unsigned char *base = stage + lane * 8;
  for (int b = b0; b < b1; ++b) {
    unsigned char *p[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) p[u] = base + (b * NU + u) * 256;
    many(p);
  }
  if (tail) {
    const int nrows = m * (m + 1) / 2;
    for (int r = (nrows / NU) * NU; r < nrows; ++r) {
      unsigned char *p[1] = {base + r * 256};
      one(p);
    }
  }
}

// ---- the consumer's solve ----------------------------------------------------------------------------
// Per lane: the (m+1) x (m+1) matrix  M = [C c; c' diag]  with C = sigma^2 rho(neighbours) + tau^2 I (strict lower
// triangle packed in A, row i at A + i(i-1)/2 rows; unit rows for i >= mi), c in cv.  Left-looking Cholesky in 4 x 4
// register tiles, block column by block column; the right-hand side rides along as row m, so row m ends up as
// y = V^-1 c and its pivot as F = diag - |y|^2.  1 / V(j,j) goes to dv.  Rows past m in the last block are register-only
// dummies (clamped loads, no stores).  Then V' B = y by blocks from the bottom; B overwrites cv.
// Returns false if a pivot of C is not positive.
__device__ __forceinline__ bool tri_solve(double *__restrict__ A, double *__restrict__ cv, double *__restrict__ dv,
                                          int m, int mi, double diag, double *F_out) {
  bool ok = true;
  double F = diag;
  auto rowp = [&](int i) -> double * { return (i < m) ? A + (i * (i - 1) / 2) * 32 : cv; };  // row m (and dummies): cv
  const int nrb = (m + 4) >> 2;  // row blocks over rows 0..m
#pragma unroll 1
  for (int jb = 0; jb < nrb; ++jb) {
    const int j0 = 4 * jb;
    double *pj[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) pj[r] = rowp(min(j0 + r, m));
    // ---- diagonal block ----
    double d[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
#pragma unroll
      for (int c = 0; c < r; ++c) d[r][c] = pj[r][(j0 + c) * 32];
      d[r][r] = (j0 + r < mi || j0 + r == m) ? diag : 1.0;
    }
#pragma unroll 1
    for (int k = 0; k < j0; k += 4) {
      double a[4][4];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int r = 0; r < 4; ++r) a[kk][r] = pj[r][(k + kk) * 32];
#pragma unroll
      for (int kk = 0; kk < 4; ++kk)
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int c = 0; c <= r; ++c) d[r][c] = fma(-a[kk][r], a[kk][c], d[r][c]);
    }
    double ri[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double piv = d[c][c];
      if (j0 + c < mi && !(piv > 0.0)) ok = false;
      if (j0 + c == m) F = piv;
      ri[c] = nngp_rsqrt_pos(piv);
#pragma unroll
      for (int r = c + 1; r < 4; ++r) d[r][c] *= ri[c];
#pragma unroll
      for (int c2 = c + 1; c2 < 4; ++c2)
#pragma unroll
        for (int r = c2; r < 4; ++r) d[r][c2] = fma(-d[r][c], d[c2][c], d[r][c2]);
    }
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      if (j0 + r < m) dv[(j0 + r) * 32] = ri[r];
      if (j0 + r <= m) {
#pragma unroll
        for (int c = 0; c < r; ++c) pj[r][(j0 + c) * 32] = d[r][c];
      }
    }
    // ---- blocks below ----
#pragma unroll 1
    for (int ib = jb + 1; ib < nrb; ++ib) {
      const int i0 = 4 * ib;
      double *pi[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) pi[r] = rowp(min(i0 + r, m));
      double e[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) e[r][c] = pi[r][(j0 + c) * 32];
#pragma unroll 1
      for (int k = 0; k < j0; k += 4) {
        double ai[4][4], aj[4][4];
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            ai[kk][r] = pi[r][(k + kk) * 32];
            aj[kk][r] = pj[r][(k + kk) * 32];
          }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) e[r][c] = fma(-ai[kk][r], aj[kk][c], e[r][c]);
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {  // e <- e L^-T with the diagonal block still in registers
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
          for (int c2 = 0; c2 < c; ++c2) e[r][c] = fma(-e[r][c2], d[c][c2], e[r][c]);
          e[r][c] *= ri[c];
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (i0 + r <= m) {
#pragma unroll
          for (int c = 0; c < 4; ++c) pi[r][(j0 + c) * 32] = e[r][c];
        }
    }
  }
  // ---- V' B = y, blocks of four unknowns from the bottom; B_i = 0 stands in for the dummies i >= m ----
  const int nb = (m + 3) >> 2;
#pragma unroll 1
  for (int ib = nb - 1; ib >= 0; --ib) {
    const int i0 = 4 * ib;
    double w[4];
#pragma unroll
    for (int c = 0; c < 4; ++c) w[c] = (i0 + c < m) ? cv[(i0 + c) * 32] : 0.0;
#pragma unroll 1
    for (int lb = nb - 1; lb > ib; --lb) {
      const int l0 = 4 * lb;
      double wl[4], v[4][4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = min(l0 + r, m - 1);
        wl[r] = (l0 + r < m) ? cv[i * 32] : 0.0;
        const double *pr = A + (i * (i - 1) / 2 + i0) * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) v[r][c] = pr[c * 32];
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) w[c] = fma(-v[r][c], wl[r], w[c]);
    }
    double t[4][4], rv[4];  // the diagonal block: t[r][c] = V(i0 + r, i0 + c), c < r
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = min(i0 + r, m - 1);
      rv[r] = dv[i * 32];
      const double *pr = A + (i * (i - 1) / 2 + i0) * 32;
#pragma unroll
      for (int c = 0; c < r; ++c) t[r][c] = (i0 + r < m) ? pr[c * 32] : 0.0;
    }
#pragma unroll
    for (int c = 3; c >= 0; --c) {
#pragma unroll
      for (int r = c + 1; r < 4; ++r) w[c] = fma(-t[r][c], w[r], w[c]);
      w[c] = (i0 + c < m) ? w[c] * rv[c] : 0.0;
    }
#pragma unroll
    for (int c = 0; c < 4; ++c)
      if (i0 + c < m) cv[(i0 + c) * 32] = w[c];
  }
  *F_out = F;
  return ok;
}

// ---- the same solve by a TEAM of warps on one tile (single-stage geometry, m >= 27) -----------------------------------
// With one stage in shared memory the loader, the producers and the consumer take turns, and a lone consumer warp needs
// ~60k cycles for the 33 x 33 factorisation (loads and dependent FMA chains of one warp).  The blocks of a block column
// are independent given the columns before it, so every warp of the CTA but the loader joins in: team rank tw takes block
// row jb + tw of column jb (rank 0 the diagonal block, which it also factors; (m + 4) / 4 <= 9 block rows, one block per
// warp at most).  Two team barriers per column: diagonal factor published -> the others finish their block (e L^-T) ->
// column published.  V' B = y runs right-looking with one barrier per block (rank 0 solves the 4 x 4 diagonal systems and
// carries the block right above; the others subtract V(ib, jb)' w_ib from the blocks further up).  Every element sees the
// same operations in the same order as in tri_solve: results are bit-identical to the one-warp solve.
__device__ __forceinline__ void team_bar(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

__device__ __forceinline__ bool tri_solve_team(double *__restrict__ A, double *__restrict__ cv, double *__restrict__ dv, int m,
                                               int mi, double diag, double *F_out, int tw, int nthreads) {
  bool ok = true;
  double F = diag;
  auto rowp = [&](int i) -> double * { return (i < m) ? A + (i * (i - 1) / 2) * 32 : cv; };  // row m (and dummies): cv
  const int nrb = (m + 4) >> 2;  // row blocks over rows 0..m
#pragma unroll 1
  for (int jb = 0; jb < nrb; ++jb) {
    const int j0 = 4 * jb;
    const int ib = jb + tw;
    const bool have = ib < nrb;
    double *pj[4], *pi[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      pj[r] = rowp(min(j0 + r, m));
      pi[r] = rowp(min(4 * ib + r, m));
    }
    double e[4][4];
    if (tw == 0) {
      // ---- diagonal block: update (lower part), factor, publish ----
#pragma unroll
      for (int r = 0; r < 4; ++r) {
#pragma unroll
        for (int c = 0; c < 4; ++c) e[r][c] = 0.0;
#pragma unroll
        for (int c = 0; c < r; ++c) e[r][c] = pj[r][(j0 + c) * 32];
        e[r][r] = (j0 + r < mi || j0 + r == m) ? diag : 1.0;
      }
#pragma unroll 1
      for (int k = 0; k < j0; k += 2) {
        double a[2][4];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int r = 0; r < 4; ++r) a[kk][r] = pj[r][(k + kk) * 32];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c <= r; ++c) e[r][c] = fma(-a[kk][r], a[kk][c], e[r][c]);
      }
      double ri[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        const double piv = e[c][c];
        if (j0 + c < mi && !(piv > 0.0)) ok = false;
        if (j0 + c == m) F = piv;
        ri[c] = nngp_rsqrt_pos(piv);
#pragma unroll
        for (int r = c + 1; r < 4; ++r) e[r][c] *= ri[c];
#pragma unroll
        for (int c2 = c + 1; c2 < 4; ++c2)
#pragma unroll
          for (int r = c2; r < 4; ++r) e[r][c2] = fma(-e[r][c], e[c2][c], e[r][c2]);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        if (j0 + r < m) dv[(j0 + r) * 32] = ri[r];
        if (j0 + r <= m) {
#pragma unroll
          for (int c = 0; c < r; ++c) pj[r][(j0 + c) * 32] = e[r][c];
        }
      }
    } else if (have) {
      // ---- a block below: update with the columns before j0 ----
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) e[r][c] = pi[r][(j0 + c) * 32];
#pragma unroll 1
      for (int k = 0; k < j0; k += 2) {
        double ai[2][4], aj[2][4];
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int r = 0; r < 4; ++r) {
            ai[kk][r] = pi[r][(k + kk) * 32];
            aj[kk][r] = pj[r][(k + kk) * 32];
          }
#pragma unroll
        for (int kk = 0; kk < 2; ++kk)
#pragma unroll
          for (int r = 0; r < 4; ++r)
#pragma unroll
            for (int c = 0; c < 4; ++c) e[r][c] = fma(-ai[kk][r], aj[kk][c], e[r][c]);
      }
    }
    team_bar(nthreads);  // the factored diagonal block and its 1 / pivots are in shared memory
    if (tw != 0 && have) {
      double d[4][4], ri[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        ri[c] = dv[(j0 + c) * 32];
#pragma unroll
        for (int c2 = 0; c2 < c; ++c2) d[c][c2] = pj[c][(j0 + c2) * 32];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) {  // e <- e L^-T
#pragma unroll
        for (int r = 0; r < 4; ++r) {
#pragma unroll
          for (int c2 = 0; c2 < c; ++c2) e[r][c] = fma(-e[r][c2], d[c][c2], e[r][c]);
          e[r][c] *= ri[c];
        }
      }
#pragma unroll
      for (int r = 0; r < 4; ++r)
        if (4 * ib + r <= m) {
#pragma unroll
          for (int c = 0; c < 4; ++c) pi[r][(j0 + c) * 32] = e[r][c];
        }
    }
    team_bar(nthreads);  // column jb is final
  }
  // ---- V' B = y, right-looking from the bottom ----
  const int nb = (m + 3) >> 2;
#pragma unroll 1
  for (int ib = nb - 1; ib >= 0; --ib) {
    const int i0 = 4 * ib;
    if (tw == 0) {
      double w[4], t[4][4], rv[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) w[c] = (i0 + c < m) ? cv[(i0 + c) * 32] : 0.0;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = min(i0 + r, m - 1);
        rv[r] = dv[i * 32];
        const double *pr = A + (i * (i - 1) / 2 + i0) * 32;
#pragma unroll
        for (int c = 0; c < r; ++c) t[r][c] = (i0 + r < m) ? pr[c * 32] : 0.0;
      }
#pragma unroll
      for (int c = 3; c >= 0; --c) {
#pragma unroll
        for (int r = c + 1; r < 4; ++r) w[c] = fma(-t[r][c], w[r], w[c]);
        w[c] = (i0 + c < m) ? w[c] * rv[c] : 0.0;
      }
#pragma unroll
      for (int c = 0; c < 4; ++c)
        if (i0 + c < m) cv[(i0 + c) * 32] = w[c];
    }
    if (ib == 0) break;
    team_bar(nthreads);  // w_ib is final
    // blocks above: y_jb -= V(ib, jb)' w_ib.  Rank 0 takes the block right above (its next system), rank tw >= 1 block tw - 1
    const int jb = (tw == 0) ? ib - 1 : tw - 1;
    if (jb < ib - (tw == 0 ? 0 : 1)) {
      const int c0 = 4 * jb;
      double wl[4], v[4][4], y[4];
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        const int i = min(i0 + r, m - 1);
        wl[r] = (i0 + r < m) ? cv[i * 32] : 0.0;
        const double *pr = A + (i * (i - 1) / 2 + c0) * 32;
#pragma unroll
        for (int c = 0; c < 4; ++c) v[r][c] = pr[c * 32];
      }
#pragma unroll
      for (int c = 0; c < 4; ++c) y[c] = cv[(c0 + c) * 32];
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int c = 0; c < 4; ++c) y[c] = fma(-v[r][c], wl[r], y[c]);
#pragma unroll
      for (int c = 0; c < 4; ++c) cv[(c0 + c) * 32] = y[c];
    }
  }
  *F_out = F;
  return ok;
}

template <int MC, bool kSlim>
__device__ __forceinline__ bool tri_solve_unrolled(double *A, double *cv, double *dv, int mi, double diag, double *F_out) {
  double Fsum;
  constexpr int kCholW = (MC >= 12 && MC <= 16) ? (kSlim ? 3 : 4) : 2;  // kSlim: the consumer runs at 200 registers
  const bool ok = tpl_solve<MC, kCholW>(A, cv, dv, mi, diag, &Fsum);
  *F_out = diag - Fsum;
  return ok;
}

// WIDE: the one-consumer geometries (m >= 20) with the team solve -- an instantiation of its own, so that the hot
// configurations (m <= 19: two to four consumers, six stages at m = 15) keep exactly the registers and the code they had
// before the team path existed (sharing one instantiation cost the m = 15 Matern launch 13 %: 0.664 -> 0.753 ms).
// MS: the consumer's solve -- 10 / 15: the fully unrolled sweep for exactly that m, 0: the rolled blocked solve (any m).
// One solve per instantiation: the unused ones were dead weight in the hot kernels (code size, register allocation).
template <int COV, bool WIDE, int MS>
__global__ void __launch_bounds__(kTriThreads, 1) bf_tri_kernel(const __grid_constant__ BfLaunch P) {
  constexpr bool kMat = kIsMatern<COV>;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const BfTheta &th = P.th[blockIdx.y];
  const int m = P.m;
  const int ne = m * (m - 1) / 2;
  const TriGeom G = tri_geometry(m, P.tri_want4 != 0);
  const int S = G.S, NC = G.NC;
  const int loader_warp = (NC == 4) ? 4 : 3, first_producer = loader_warp + 1, nprod = 12 - first_producer;
  // one consumer only (m >= 20; from 27 on a single stage, where the phases of a tile cannot overlap at all): a lone warp
  // needs tens of thousands of cycles per factorisation, so every warp but the loader joins the solve (tri_solve_team)
  constexpr bool wide = WIDE;   // host side: WIDE == (tri_geometry(m).NC == 1)
  constexpr int kMaxM = WIDE ? kTriMaxM : kTriNarrowMaxM;
  constexpr int kTeamThreads = 11 * 32;
  const double *__restrict__ resid = th.resid;
  const bool with_cur = th.Bcur != nullptr && resid != nullptr;
  const long long nloc = (long long)P.loc_end - P.loc_begin;
  const long long ntiles = (nloc + 31) / 32;
  unsigned char *smem = tpl_smem;
  const unsigned bar0 = smem_u32(smem + G.off_bar);
  auto bar_full = [&](int s) { return bar0 + 8u * s; };
  auto bar_built = [&](int s) { return bar0 + 8u * (kTriMaxStages + s); };
  auto bar_empty = [&](int s) { return bar0 + 8u * (2 * kTriMaxStages + s); };
  auto bar_rr = [&](int s) { return bar0 + 8u * (3 * kTriMaxStages + s); };
  volatile long long *ring = reinterpret_cast<volatile long long *>(smem + G.off_bar + 8 * 4 * kTriMaxStages);
  double *s_tab = reinterpret_cast<double *>(smem + G.off_tab);
  __shared__ double red[3][12];
  __shared__ int last_flag;
  __shared__ int bbeg[kTriProducers + 1];  // batches of producer warp pw: [bbeg[pw], bbeg[pw + 1])

  // ---- one-time setup ---------------------------------------------------------------------------------
  if (tid == 0) {
    for (int s = 0; s < S; ++s) {
      mbar_init(bar_full(s), 1);                 // expect_tx arrival of the loader's lane 0: the distances have landed
      mbar_init(bar_rr(s), 32);                  // 32 cp.async arrivals: the residual gathers have landed -- a barrier of
                                                 // their own, only the consumer needs them: behind bar_full their L2
                                                 // latency sat on the producers' critical path (spherical 0.47 -> ms)
      mbar_init(bar_built(s), nprod);            // one arrival per producer warp
      mbar_init(bar_empty(s), 1);                // the consumer
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // Shares of the producer warps, so that the four schedulers carry equal FP64 work: warp w issues on scheduler
    // w & 3; schedulers 0..NC-1 also run a consumer (C instructions per tile, one tile in NC); a scheduler whose only
    // producer would carry a whole quarter alone gets less (that warp would pace every tile).
    const int nrows = m * (m + 1) / 2, nb = nrows / kTriUnr<COV>;
    // tri_cw: the consumer's weight in tenths (its dependent chains suffer from sharing a scheduler more than its
    // instruction count says)
    const float C = wide ? 0.f : ((float)(m * m * m) / 6.f + 2.f * m * m + 25.f * m) * 0.1f * (float)P.tri_cw;
    const float Tp = (float)nrows * kTriRowCost<COV>;
    int np[4] = {0, 0, 0, 0};
    for (int w = first_producer; w < 12; ++w) np[w & 3]++;
    float f[4], fs = 0.f;
    for (int q = 0; q < 4; ++q) {
      f[q] = fmaxf(0.25f * NC * (C + Tp) - (q < NC ? C : 0.f), 0.f);
      if (np[q] == 1) f[q] *= 0.6f;
      if (np[q] == 0) f[q] = 0.f;
      fs += f[q];
    }
    float acc = 0.f;
    int prev = 0, cnt[kTriProducers] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int q = 0; q < 4; ++q) {
      acc += f[q] / fs;
      const int upto = (q == 3) ? nb : min(nb, (int)(acc * nb + 0.5f));
      int nq = upto - prev, left = np[q];
      prev = upto;
      for (int w = first_producer; w < 12; ++w)
        if ((w & 3) == q) {
          const int take = nq / left;
          cnt[w - first_producer] = take;
          nq -= take;
          --left;
        }
    }
    bbeg[0] = 0;
    for (int w = 0; w < kTriProducers; ++w) bbeg[w + 1] = bbeg[w] + cnt[w];
  }
  if (kMat || wide) {
    for (int x = tid; x < 32; x += kTriThreads) s_tab[x] = kExpTab32[x];
    if (kMat)
      for (int x = tid; x < 128; x += kTriThreads) s_tab[32 + x] = kLogTab64[x];
  }
  for (int s = 0; s < S; ++s) {  // residual rows: stay zero when the launch has no quadratic form
    double *rr = reinterpret_cast<double *>(smem + s * G.stage_bytes + G.off_rr);
    for (int x = tid; x < (m + 1) * 32; x += kTriThreads) rr[x] = 0.0;
  }
  __syncthreads();

  auto run_loader = [&]() {
    // ================================ loader =========================================================
    // Tile sequence of this CTA: T(j) = blockIdx.x + j gridDim.x for j < S (static: spreads the slow head of the
    // ordering over the CTAs), then S gridDim.x + (atomic counter).  Software pipeline so that the warp never waits
    // for global memory: at trip j the copies of T(j) are issued with neighbour ids that were loaded during trip
    // j - 1, the ids of T(j+1) are requested, and the atomic that claims T(j+2) is outstanding in lane 0.
    const unsigned tri_bytes = (unsigned)(ne + m) * 256u;
    const size_t tile_doubles = (size_t)(ne + m) * 32;
    int ids_a[kMaxM], ids_b[kMaxM];  // ping-pong: the ids a trip loads are first READ one trip later (a register
                                           // copy at the end of the trip would wait for the loads right away)
    auto load_ids = [&](long long tile, int(&ids)[kMaxM]) {
      const int *src = P.nnT + (size_t)tile * m * 32 + lane;
#pragma unroll
      for (int k = 0; k < kMaxM; ++k) ids[k] = (k < m) ? __ldg(src + k * 32) : -1;
    };
    long long t_cur = blockIdx.x, t_nxt = (long long)blockIdx.x + gridDim.x;
    unsigned claim = 0u;
    bool pending = false;
    if (resid != nullptr && t_cur < ntiles) load_ids(t_cur, ids_a);
    // launched beside the residual kernel (BfLaunch::pdl): everything up to here -- barriers, tables, zeroed residual rows,
    // the first neighbour ids -- needed nothing of it; the gathers below do
    if (P.pdl) asm volatile("griddepcontrol.wait;" ::: "memory");
    int done_markers = 0, s = 0, use = 0, j = 0;  // stage of trip j, and how often it has been used before
    // one trip: issue T(j) with `ids_use`, request the ids of T(j+1) into `ids_load`; false when the work is over
    auto trip = [&](int(&ids_use)[kMaxM], int(&ids_load)[kMaxM]) -> bool {
      if (j > 0) {  // T(j+1): static, or the claim issued one trip ago
        if (pending) t_nxt = (long long)S * gridDim.x + __shfl_sync(0xffffffffu, claim, 0);
        else t_nxt = (long long)blockIdx.x + (long long)(j + 1) * gridDim.x;
      }
      pending = false;
      if (j + 2 >= S) {  // claim T(j+2); past the end: stay there without touching the counter
        claim = 0x7fffffffu;
        if (t_nxt < ntiles && lane == 0) claim = atomicAdd(th.tile_counter, 1u);
        pending = true;
      }
      if (resid != nullptr && t_nxt < ntiles) load_ids(t_nxt, ids_load);
      if (use > 0) mbar_wait(bar_empty(s), (use - 1) & 1);
      if (t_cur >= ntiles) {  // end of work: NC consecutive markers, one per consumer (they take the stages round-robin)
        if (lane == 0) ring[s] = -1;
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full(s));
        if (++done_markers == NC) return false;
      } else {
        unsigned char *st = smem + s * G.stage_bytes;
        const long long base = P.loc_begin + t_cur * 32;
        if (lane == 0) {
          ring[s] = t_cur;
          // a single stage cannot hide the copy behind the solve: have the NEXT tile's block in L2 by then
          if (S == 1 && t_nxt < ntiles) bulk_prefetch_l2(P.tri + (size_t)t_nxt * tile_doubles, tri_bytes);
          mbar_expect_tx(bar_full(s), tri_bytes);
          // c rows and the packed triangle are adjacent in the cache and in the stage: one copy
          bulk_g2s(smem_u32(st + G.off_cv), P.tri + (size_t)t_cur * tile_doubles, tri_bytes, bar_full(s));
          // the resident weights of the current theta are read by the consumer two stages from now: have them in L2
          if (with_cur && tri_tile_is_full(base, m, P.loc_end)) bulk_prefetch_l2(th.Bcur + nn_offset(base, m), (unsigned)m * 256u);
        }
        if (resid != nullptr) {
          double *rr = reinterpret_cast<double *>(st + G.off_rr) + lane;
          long long i = base + lane;
          if (i >= P.loc_end) i = (long long)P.loc_end - 1;
#pragma unroll
          for (int k = 0; k < kMaxM; ++k)
            if (k < m) cp_async8_zfill(&rr[k * 32], &resid[ids_use[k] >= 0 ? ids_use[k] : 0], ids_use[k] >= 0);
          cp_async8(&rr[m * 32], &resid[i]);
        }
        cp_async_mbar_arrive(bar_rr(s));
      }
      t_cur = t_nxt;
      if (++s == S) {
        s = 0;
        ++use;
      }
      ++j;
      return true;
    };
    for (;;) {
      if (!trip(ids_a, ids_b)) break;
      if (!trip(ids_b, ids_a)) break;
    }
  };

  if (warp < 4) {
    if constexpr (kTriSlim<COV>) asm volatile("setmaxnreg.inc.sync.aligned.u32 200;");
    else asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    if (warp == loader_warp) {
      run_loader();
    } else if (warp < NC || wide) {
      // ================================ consumers ======================================================
      // (wide: warp 0 is the tile's consumer, warps 1 and 2 only take part in the team solve)
      unsigned char *cons = smem + G.off_cons + (wide ? 0 : warp) * G.cons_bytes;
      double *dv = reinterpret_cast<double *>(cons) + lane;  // [k][lane]: 1 / V(k,k) during the solve ...
      double *bo = reinterpret_cast<double *>(cons);         // ... then [lane][k], the image of the tile's block of B
      const double diag = th.sigmaSq + th.tauSq;
      int failed = INT_MAX;
      bool store_pending = false;
      int s = wide ? 0 : warp, use = 0;
      while (s >= S) {  // (NC <= S always; kept general)
        s -= S;
        ++use;
      }
      for (;;) {
        unsigned char *st = smem + s * G.stage_bytes;
        double *A = reinterpret_cast<double *>(st + G.off_a) + lane;
        double *cv = reinterpret_cast<double *>(st + G.off_cv) + lane;
        const double *rr = reinterpret_cast<const double *>(st + G.off_rr) + lane;
        mbar_wait(bar_built(s), use & 1);
        const long long tile = ring[s];
        if (tile < 0) break;
        const long long base = P.loc_begin + tile * 32;
        const bool active = base + lane < P.loc_end;
        const long long i = active ? base + lane : (long long)P.loc_end - 1;
        const int mi = nn_count(i, m);
        const long long off = nn_offset(i, m);
        const bool full = tri_tile_is_full(base, m, P.loc_end);
        if (wide && warp != 0) {  // helper of the team solve
          double Fh;
          team_bar(kTeamThreads);
          tri_solve_team(A, cv, dv, m, mi, diag, &Fh, warp, kTeamThreads);
          if (++s >= S) {
            s = 0;
            ++use;
          }
          continue;
        }
        double fcur = 1.0;
        if (with_cur) fcur = __ldg(&th.Fcur[i]);
        const unsigned rr_parity = use & 1;
        if (mi < m) {  // ragged start of the ordering: unit rows k >= m_i
          for (int k = mi; k < m; ++k) {
            cv[k * 32] = 0.0;
            for (int l = 0; l < k; ++l) A[(k * (k - 1) / 2 + l) * 32] = 0.0;
          }
        }
        // Q term of the RESIDENT factors (current theta of the sampler) with the same residuals
        double qcur_part = 0.0;
        if (with_cur) {
          mbar_wait(bar_rr(s), rr_parity);
          // the location's m resident weights are 8 m contiguous bytes (the 32 rows of a tile share ~30 cache lines):
          // all loads first -- one memory round trip instead of one per unrolled group -- and done before the solve
          // claims the registers (fused launch, exponential: 0.550 -> 0.515 ms together with the loader's L2 prefetch).
          // Forming this sum in the loader warp instead, one trip behind its gathers, was measured and dropped: the
          // loader then paces the tiles (0.918 ms).
          double bc[kMaxM];
#pragma unroll
          for (int k = 0; k < kMaxM; ++k) bc[k] = (k < mi) ? __ldg(&th.Bcur[off + k]) : 0.0;
#pragma unroll
          for (int k = 0; k < kMaxM; ++k)
            if (k < m) qcur_part = fma(bc[k], rr[k * 32], qcur_part);
        }
        if (store_pending) {  // the staging area doubles as the solve's 1 / diag rows: the last bulk store must have read it
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
          store_pending = false;
        }
        // The usual neighbour counts take the fully unrolled column sweep (tpl_device.cuh; same layout when m equals its
        // capacity): 12.4k cycles per tile against 16.4k for the rolled blocks at m = 15 (stall_wait on the loop-carried
        // chains), in spite of its instruction-fetch stalls.  Every other m: the rolled blocked solve.
        double Fv;
        bool ok;
        if (wide) {
          team_bar(kTeamThreads);  // the unit rows of a ragged tile are in place
          ok = tri_solve_team(A, cv, dv, m, mi, diag, &Fv, 0, kTeamThreads);
        } else
        {
          if constexpr (MS == 10) ok = tri_solve_unrolled<10, kTriSlim<COV>>(A, cv, dv, mi, diag, &Fv);
          else if constexpr (MS == 15) ok = tri_solve_unrolled<15, kTriSlim<COV>>(A, cv, dv, mi, diag, &Fv);
          else ok = tri_solve(A, cv, dv, m, mi, diag, &Fv);
        }
        if (!ok && active) failed = min(failed, (int)min(i + 1, (long long)INT_MAX - 1));
        if (!with_cur) mbar_wait(bar_rr(s), rr_parity);  // long landed by now: the solve took microseconds
        const bool to_bulk = th.B != nullptr && full;
        __syncwarp();  // every lane is done with the 1 / diag rows
        double part = 0.0;
#pragma unroll 4
        for (int k = 0; k < m; ++k) {  // residual rows are zero when the launch has no quadratic form
          const double wk = cv[k * 32];
          part = fma(wk, rr[k * 32], part);
          if (to_bulk) bo[lane * m + k] = wk;
        }
        const double r_own = rr[m * 32];
        if (to_bulk) {
          fence_proxy_async();
          __syncwarp();
          if (lane == 0) {
            bulk_s2g(th.B + nn_offset(base, m), smem_u32(bo), (unsigned)m * 256u);
            bulk_commit();
          }
          store_pending = true;
        } else if (th.B != nullptr && active) {
          for (int k = 0; k < mi; ++k) th.B[off + k] = cv[k * 32];
        }
        fence_proxy_async();  // this warp's generic-proxy accesses to the stage precede the next bulk copy into it
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_empty(s));
        if (th.F != nullptr && active) th.F[i] = Fv;
        double lt = 0.0, qt = 0.0, qc = 0.0;
        if (active) {
          lt = (Fv > 0.0) ? nngp_log_pos(Fv) : log(Fv);
          if (resid != nullptr) {
            const double e = r_own - part;
            qt = e * e * nngp_rcp_pos(Fv);
            if (with_cur) {
              const double ec = r_own - qcur_part;
              qc = ec * ec / fcur;
            }
          }
        }
#pragma unroll
        for (int sft = 16; sft >= 1; sft >>= 1) {
          lt += __shfl_down_sync(0xffffffffu, lt, sft);
          qt += __shfl_down_sync(0xffffffffu, qt, sft);
          qc += __shfl_down_sync(0xffffffffu, qc, sft);
        }
        if (lane == 0) {
          th.partials[4 * tile] = lt;
          th.partials[4 * tile + 1] = qt;
          th.partials[4 * tile + 2] = qc;
        }
        s += NC;
        if (s >= S) {
          s -= S;
          ++use;
        }
      }
      if (store_pending && lane == 0) bulk_wait0();
      if (failed != INT_MAX) atomicMin(th.fail, failed);
    }
  } else {
    // ================================ producers ========================================================
    if constexpr (kTriSlim<COV>) asm volatile("setmaxnreg.dec.sync.aligned.u32 152;");
    else asm volatile("setmaxnreg.dec.sync.aligned.u32 136;");
    if (warp == loader_warp) {
      run_loader();
    } else {
    const int pw = warp - first_producer;
    const int b0 = bbeg[pw], b1 = bbeg[pw + 1];
    const bool tail = (pw == 0);
    int markers = 0, s = 0, use = 0;
    for (;;) {
      mbar_wait(bar_full(s), use & 1);
      const long long tile = ring[s];
      if (tile >= 0) {
        unsigned char *st = smem + s * G.stage_bytes;
        if constexpr (kMat) {
          auto body = [&](auto idx) {
            constexpr int B = decltype(idx)::value;
            const Cor<COV> cor(P.th[B], s_tab);
            const double sig = P.th[B].sigmaSq;
            tri_transform<COV>(
                st, lane, b0, b1, tail, m, [&](unsigned char *const(&p)[1]) { tri_rows_matern<COV, 1>(p, cor, sig); },
                [&](unsigned char *const(&p)[kTriUnr<COV>]) { tri_rows_matern<COV, kTriUnr<COV>>(p, cor, sig); });
          };
          switch (blockIdx.y) {  // compile-time set index: the series coefficients stay constant-bank operands
            case 0: body(std::integral_constant<int, 0>()); break;
            case 1: body(std::integral_constant<int, 1>()); break;
            case 2: body(std::integral_constant<int, 2>()); break;
            default: body(std::integral_constant<int, 3>()); break;
          }
        } else {
          if constexpr (COV == NNGP_COV_EXPONENTIAL || COV == NNGP_COV_GAUSSIAN) {
            if (wide) {
              const FromDistTab<COV> f(th, s_tab);
              tri_transform<COV>(
                  st, lane, b0, b1, tail, m, [&](unsigned char *const(&p)[1]) { tri_rows<COV, 1>(p, f); },
                  [&](unsigned char *const(&p)[kTriUnr<COV>]) { tri_rows<COV, kTriUnr<COV>>(p, f); });
            } else {
              const FromDist<COV> f(th, nullptr);
              tri_transform<COV>(
                  st, lane, b0, b1, tail, m, [&](unsigned char *const(&p)[1]) { tri_rows<COV, 1>(p, f); },
                  [&](unsigned char *const(&p)[kTriUnr<COV>]) { tri_rows<COV, kTriUnr<COV>>(p, f); });
            }
          } else {
            const FromDist<COV> f(th, nullptr);
            tri_transform<COV>(
                st, lane, b0, b1, tail, m, [&](unsigned char *const(&p)[1]) { tri_rows<COV, 1>(p, f); },
                [&](unsigned char *const(&p)[kTriUnr<COV>]) { tri_rows<COV, kTriUnr<COV>>(p, f); });
          }
        }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(bar_built(s));
      if (wide && tile >= 0) {  // single stage: nothing to build until this tile is solved -- join the solve
        mbar_wait(bar_built(s), use & 1);
        unsigned char *st = smem + s * G.stage_bytes;
        const long long base = P.loc_begin + tile * 32;
        const long long i = (base + lane < P.loc_end) ? base + lane : (long long)P.loc_end - 1;
        double Fh;
        team_bar(kTeamThreads);
        tri_solve_team(reinterpret_cast<double *>(st + G.off_a) + lane, reinterpret_cast<double *>(st + G.off_cv) + lane,
                       reinterpret_cast<double *>(smem + G.off_cons) + lane, m, nn_count(i, m), th.sigmaSq + th.tauSq, &Fh,
                       warp - 1, kTeamThreads);
      }
      if (tile < 0 && ++markers == NC) break;  // the markers pass through, so that every consumer sees one
      if (++s == S) {
        s = 0;
        ++use;
      }
    }
    }
  }

  // ---- grid reduction: the last CTA to finish adds the per-tile partials in tile order --------------------
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned int tk = atomicAdd(th.ticket, 1u);
    last_flag = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (last_flag) {
    __threadfence();
    double a = 0.0, b = 0.0, c = 0.0;
    for (long long t = tid; t < ntiles; t += kTriThreads) {
      a += __ldcg(&th.partials[4 * t]);
      b += __ldcg(&th.partials[4 * t + 1]);
      c += __ldcg(&th.partials[4 * t + 2]);
    }
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, sft);
      b += __shfl_down_sync(0xffffffffu, b, sft);
      c += __shfl_down_sync(0xffffffffu, c, sft);
    }
    if (lane == 0) {
      red[0][warp] = a;
      red[1][warp] = b;
      red[2][warp] = c;
    }
    __syncthreads();
    if (tid == 0) {
      double o0 = 0.0, o1 = 0.0, o2 = 0.0;
      for (int w = 0; w < 12; ++w) {
        o0 += red[0][w];
        o1 += red[1][w];
        o2 += red[2][w];
      }
      const int fv = atomicExch(th.fail, INT_MAX);
      th.out[0] = o0;
      th.out[1] = o1;
      th.out[2] = (fv != INT_MAX) ? 1.0 : 0.0;
      th.out[3] = o2;
      *th.fail_out = fv;
      *th.tile_counter = 0u;
      *th.ticket = 0u;
      __threadfence_system();  // results before the stamp, as seen by the polling host (and by peers' hosts)
      *reinterpret_cast<volatile double *>(&th.out[kBfOutSeq]) = th.seq;
    }
  }
}

// ---- launcher ---------------------------------------------------------------------------------------
template <int COV>
static cudaError_t tri_launch_one(const BfLaunch &P0, int grid_x, cudaStream_t stream) {
  static int want4_env = -2;  // NNGP_TRI_NC4 = 0 / 1 overrides the per-model choice (A/B measurements)
  if (want4_env == -2) {
    const char *e = getenv("NNGP_TRI_NC4");
    want4_env = e ? atoi(e) : -1;
  }
  static int cw_env = -2;  // NNGP_TRI_CW: consumer weight of the producers' shares, in tenths
  if (cw_env == -2) {
    const char *e = getenv("NNGP_TRI_CW");
    cw_env = e ? atoi(e) : -1;
  }
  BfLaunch P = P0;
  if constexpr (kIsMatern<COV> && NNGP_TRI_SIG_FOLD) {
    // sigma^2 folded into the series coefficients of this launch's copy: the producers' fast path returns sigma^2 rho
    // (one multiplication less per entry; exact for the powers of two of the scaling tests); slow() keeps its own tables
    for (int b = 0; b < P.nbatch; b++)
      for (int k = 0; k < kMatSeriesTerms; k++) {
        P.th[b].ms.sa[k] *= P.th[b].sigmaSq;
        P.th[b].ms.sb[k] *= P.th[b].sigmaSq;
      }
  }
  P.tri_cw = cw_env > 0 ? cw_env : kTriConsumerWeight<COV>;
  // measured (n = 10^6, m = 15, same box): with the quadratic form the loader's gathers pace a four-consumer CTA
  // (exponential 0.498 vs 0.466 ms, spherical 0.446 vs 0.412), without it four consumers win (0.416 vs 0.433, 0.331 vs 0.366)
  P.tri_want4 = want4_env >= 0 ? want4_env : ((kTriWant4<COV> && P.th[0].resid == nullptr) ? 1 : 0);
  const TriGeom G = tri_geometry(P.m, P.tri_want4 != 0);
  const bool wide = (G.NC == 1);
  if (!wide && P.m > kTriNarrowMaxM) return cudaErrorInvalidValue;  // (cannot happen: NC == 1 from m = 20 on)
  // one instantiation per (geometry class, solve): 0 narrow / rolled, 1 narrow / unrolled 10, 2 narrow / unrolled 15, 3 wide
#ifndef NNGP_TRI_ROLLED_ONLY
  const int kind = wide ? 3 : (P.m == 15 ? 2 : (P.m == 10 ? 1 : 0));
#else
  const int kind = wide ? 3 : 0;
#endif
  void (*kern)(BfLaunch) = kind == 3   ? bf_tri_kernel<COV, true, 0>
                           : kind == 2 ? bf_tri_kernel<COV, false, 15>
                           : kind == 1 ? bf_tri_kernel<COV, false, 10>
                                       : bf_tri_kernel<COV, false, 0>;
  static bool configured[4][64] = {};  // the attribute is per device; setting it costs a few microseconds of every launch
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64 || !configured[kind][dev]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - 1024);
    if (e != cudaSuccess) return e;
    if (dev >= 0 && dev < 64) configured[kind][dev] = true;
  }
  dim3 grid(grid_x, P.nbatch, 1);
  if (P.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(kTriThreads, 1, 1);
    cfg.dynamicSmemBytes = (size_t)G.smem_bytes;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, P);
  }
  kern<<<grid, kTriThreads, G.smem_bytes, stream>>>(P);
  return cudaGetLastError();
}

// Where this kernel is the fastest one (n = 5 x 10^5, ms per launch, exponential / Matern: this kernel vs the gather
// kernels): m = 5: 0.148 / 0.163 vs 0.068 / 0.126; 8: 0.179 / 0.209 vs 0.121 / 0.171; 10: 0.174 / 0.215 vs 0.166 / 0.226;
// 12: 0.267 / 0.320 vs 0.239 / 0.297; 13: 0.290 / 0.357 vs 0.292 / 0.380; 14: 0.296 / 0.374 vs 0.309 / 0.424;
// 15: 0.233 / 0.324 vs 0.310 / 0.426; 16: 0.414 / 0.522 vs 0.400 / 0.505; 17: 0.553 / 0.668 vs 0.726 / 0.871;
// 18: 0.564 / 0.694 vs 0.717 / 0.880; 20 (n = 3 x 10^5): 0.50 / 0.56 vs 0.44 / 0.56; 24: 1.13 vs 0.70; 26: 1.17 / 1.27 vs 1.75 / 2.48.
// Small m: per-tile hand-offs outweigh the saved distance arithmetic; 16: the rolled solve loses to bf_ws_kernel's
// unrolled one; from 20 on only one consumer and two or three stages fit (tri_geometry) and bf_ws_kernel's two-stage
// CTA wins at its capacities 20 and 24.  nngp_set_option("tri_force", 1) serves every 2 <= m <= 32 (tests).
// With the team solve (one-consumer geometries, m >= 20; n = 5 x 10^5, exponential / Matern): 20: 0.908 / 1.156 vs 0.718 / 0.915
// (bf_ws_kernel, capacity 20); 22: 0.963 / 1.293 vs 1.139 / 1.377 (capacity 24); 24: 1.165 / 1.516 vs 1.150 / 1.427; 26: 1.226 /
// 1.699 (lone consumer: 1.95 / 2.12); 28: 1.708 / 2.207; 32: 2.217 / 2.853 vs 3.43 / 4.93 (cooperative bf_kernel).
static bool tri_is_fastest(int m) {
  return (m >= 13 && m <= 15) || (m >= 17 && m <= 19) || (m >= 21 && m <= 23) || (m >= 25 && m <= kTriMaxM);
}
static int g_tri_force = 0;
void tri_set_force(int on) { g_tri_force = on; }
int tri_capacity_for_m(int m) {
  if (m < 2 || m > kTriMaxM) return 0;
  return (g_tri_force || tri_is_fastest(m)) ? m : 0;
}

int tri_grid_for(long long nloc, int sms) {
  const long long need = (nloc + 31) / 32;
  long long g = sms;
  if (need < g) g = need;
  return (int)(g < 1 ? 1 : g);
}

cudaError_t tri_launch(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream) {
  if (!tri_capacity_for_m(P.m)) return cudaErrorInvalidValue;
  switch (cov) {
    case NNGP_COV_EXPONENTIAL: return tri_launch_one<NNGP_COV_EXPONENTIAL>(P, grid_x, stream);
    case NNGP_COV_SPHERICAL: return tri_launch_one<NNGP_COV_SPHERICAL>(P, grid_x, stream);
    case NNGP_COV_MATERN:
      if (P.matern_short == 2) return tri_launch_one<NNGP_COV_MATERN_TINY>(P, grid_x, stream);
      if (P.matern_short) return tri_launch_one<NNGP_COV_MATERN_SHORT>(P, grid_x, stream);
      return tri_launch_one<NNGP_COV_MATERN>(P, grid_x, stream);
    case NNGP_COV_GAUSSIAN: return tri_launch_one<NNGP_COV_GAUSSIAN>(P, grid_x, stream);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace nngp

// bf_ws_kernel.cu -- warp-specialised B/F + log-likelihood kernel for 6 <= m <= 24.
//
// Same mathematics as bf_tpl_kernel.cu (one lane owns one location; reference sites: updateBF
// src/rNNGP.cpp:46-101, updateConjBF src/cNNGP.cpp:24-69, Q src/util.cpp:261-280, prediction solves
// src/rNNGPPredict.cpp:154-169 / src/cNNGP.cpp:260-275), different execution structure.  ncu on the
// single-role kernel showed its three phases have opposite needs: the covariance build (70% of the
// instructions) runs at the FP64 pipe's cadence, while the gathers and the Cholesky chain are latency-
// bound, and with ~1 warp per scheduler nothing overlaps them (FP64 pipe 32% busy overall).  Here a CTA
// pipelines TILES of 32 locations through two shared-memory stages:
//
//   producer warps 1..NP : gather neighbour ids -> cp.async sites and residuals into the stage, then
//                          each evaluates its share of the tile's m(m-1)/2 + m correlations
//                          (NP = 3; Matern 5)
//   consumer warp 0      : Cholesky + solves of the previous tile, B/F stores, sum log F / Q terms
//
// hand-off with named barriers (bar.arrive / bar.sync, ids 1..4 = full/empty per stage, id 5 = producers'
// own rendezvous).  Two CTAs per SM (2 x 2 stages x 46 KB; one CTA for the 20 / 24 capacities).
//
// Measured at n = 10^6, m = 15 (ms per launch; exponential / Matern): cooperative kernel 1.11 / 1.79,
// single-role one-lane kernel 0.90 / 2.63, this kernel 0.62 / 0.85.  History: 0.72 / 1.63 with a static tile
// split and gathers inside the producers' critical path; dynamic tiles + the Matern series/quadrature paths
// -> 0.71 / 1.20; gathers software-pipelined one tile ahead -> 0.62 / 1.11; table-driven log/exp -> 0.99;
// launch-level 9-term Matern series -> 0.95; producer loop trimmed (pre-multiplied work list, no clamp, integer
// slow-path test, negated coefficients) -> 0.88; five Matern producer warps -> 0.86; 4-column Cholesky blocks for
// the other models 0.64 -> 0.62.  ncu now: FP64 pipe 45-50% busy, issue slots 46%, shared-memory pipe ~60%; every
// warp of a CTA is busy about equally long -- latency-bound on all sides (stall_wait and stall_short_scoreboard
// ~1.1 per issued instruction each), not bound by one role.
// Variants tried and rejected: one 8-warp CTA with a consumer and a producer on every scheduler (the producer's
// DFMA stream starves the consumer's dependent chain); 2-4 consumers packed on scheduler 0 with 6 producers on
// schedulers 1-3 (the fully unrolled solve, 30 KB of code per consumer, thrashes the instruction cache: ncu
// stall_no_instruction 40-49% in the solve); Matern batches of 6 / 8 chains; batches drawn dynamically from a
// shared counter; unequal static shares; 6 Matern producers (spills at 128 registers); 4 / 5 producers for the
// exponential kernel (0.75 / 0.71 with 4-column blocks, 0.68 / 0.65 with 2-column blocks vs 0.62); the table-driven
// exp for the exponential / gaussian producers; a log table replicated per quarter-warp lane (no bank conflicts, same
// time).  The -D knobs below keep those experiments reproducible (tools/build_variant.sh + tools/ab_bench.py).
#include <stdlib.h>


#include "tpl_device.cuh"

#ifndef NNGP_WS_EXP_TAB
#define NNGP_WS_EXP_TAB 0  // 1: exponential / gaussian producers use the table-driven exp -- measured SLOWER (0.619 -> 0.665 ms,
                           // gaussian 0.569 -> 0.646): the lookups' shared-memory latency outweighs the 8 FP64 instructions saved
#endif
#ifndef NNGP_WS_NP
#define NNGP_WS_NP 3  // producer warps of the exponential / spherical / gaussian kernels
#endif
#ifndef NNGP_WS_MATERN_NP
#define NNGP_WS_MATERN_NP 5  // producer warps of the Matern kernels: 3 / 4 / 5 / 6 -> 0.879 / 0.927 / 0.859 / 0.957 ms (n = 1e6, m = 15)
#endif
#ifndef NNGP_WS_SHARES  // relative shares of the Matern producer warps, one decimal digit each, warp 1 first
#define NNGP_WS_SHARES 11111  // (NNGP_WS_MATERN_NP digits) equal; 58855 / 59935 / 67756 / 68836 measured 0.900 / 0.887 / 0.877 / 0.864 ms
#endif

namespace nngp {

template <int MC, int NP>
struct WsGeo {
  using T = TplGeo<MC>;
  static constexpr int OFF_RR = T::kWarpBytes;                   // double[MC+1][32]: residuals of neighbours, own
  static constexpr int kStageBytes = OFF_RR + (MC + 1) * 32 * 8;
  static constexpr int kStages = 2;
  static constexpr int kRrBytes = (MC + 1) * 32 * 8;
  static constexpr int kSmemBytes = kStages * kStageBytes + kRrBytes;  // + third slot of the residual ring
  static constexpr int kProducerWarps = NP;
  static constexpr int kThreads = 32 * (1 + kProducerWarps);
};

__device__ __forceinline__ void named_bar_sync(int id, int count) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(int id, int count) {
  asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// tile of 32 complete (m-neighbour) rows: its weights are one contiguous run of 32 m doubles in B
__device__ __forceinline__ bool ws_full_tile(long long base, int m, long long loc_end) {
  return base >= m && base + 32 <= loc_end;
}

// residual ring, 3 deep: slots 0 and 1 live inside the two stages, slot 2 behind them
template <int MC, int NP>
__device__ __forceinline__ double *ws_rr(int slot) {
  using Ws = WsGeo<MC, NP>;
  return reinterpret_cast<double *>(tpl_smem + (slot < 2 ? slot * Ws::kStageBytes + Ws::OFF_RR : 2 * Ws::kStageBytes));
}

template <int MC, int COV, int NP>
__global__ void __launch_bounds__(32 * (1 + NP), (MC <= 16) ? 2 : 1) bf_ws_kernel(const __grid_constant__ BfLaunch P) {  // 20, 24: one CTA per SM
  using Geo = TplGeo<MC>;
  using Ws = WsGeo<MC, NP>;
  constexpr int kBatch = kTplBatch<COV>;
  // who stores a solved tile's weights: the producers where the consumer is the critical path (0.62 -> 0.59 ms,
  // spherical), the consumer for Matern, whose producers are the critical path (1.11 vs 1.15 ms)
#ifdef NNGP_WS_PSTORE  // experiment builds: 1 = producers copy B out for every model (Matern with five producers: 0.850 ->
                       // 0.878 ms), 0 = the consumer always does
  constexpr bool kProducersStoreB = (NNGP_WS_PSTORE != 0);
#else
  constexpr bool kProducersStoreB = !kIsMatern<COV>;
#endif
  __shared__ TplEnt<COV> ent[kTplMaxEnt];
  __shared__ int ebeg[NP + 1];
  __shared__ long long next_tile[8];  // ring of claimed tiles, slot j & 7 (>= ntiles: no more work)
  __shared__ double red[2][1 + NP];
  __shared__ int last_flag;
  constexpr int kTabLen = 32 + 128 * kLogTabStride;  // exp table, then the log table (x8 interleaved copies when replicated)
  constexpr bool kExpTab = NNGP_WS_EXP_TAB && (COV == NNGP_COV_EXPONENTIAL || COV == NNGP_COV_GAUSSIAN);  // exp table only
  __shared__ __align__(16) double s_tab[kIsMatern<COV> ? kTabLen : (kExpTab ? 32 : 2)];  // exp / log tables (fasttab.cuh)
  const int tid = threadIdx.x;
  const int warp = tid >> 5, lane = tid & 31;

  const BfTheta &th = P.th[blockIdx.y];
  const int m = P.m;
  const bool site = P.site_mode != 0;
  const double sigmaSq = th.sigmaSq, tauSq = th.tauSq;
  const double diag = sigmaSq + tauSq;
  const double *__restrict__ resid = th.resid;

  if (kExpTab) {
    for (int x = tid; x < 32; x += Ws::kThreads) s_tab[x] = kExpTab32[x];
  }
  if (kIsMatern<COV>) {
    for (int x = tid; x < 32; x += Ws::kThreads) s_tab[x] = kExpTab32[x];
    for (int x = tid; x < 128 * kLogTabStride; x += Ws::kThreads) {  // pair idx, copy c -> slot idx * stride + c
      const int pr = x >> 1, idx = pr / kLogTabStride;
      s_tab[32 + x] = kLogTab64[2 * idx + (x & 1)];
    }
  }
  // ---- one-time setup: work list, zero padding of both stages --------------------------------------
  const int ne = m * (m - 1) / 2;
  const int nent_pad = ((ne + m + kBatch - 1) / kBatch) * kBatch;
  if (tid <= NP) {
    // entry ranges of the producer warps, in whole batches
    const int nb = nent_pad / kBatch;
    int shares[8] = {1, 1, 1, 1, 1, 1, 1, 1};
    if (kIsMatern<COV>) {
      long long d = NNGP_WS_SHARES;
      for (int i = NP - 1; i >= 0; --i, d /= 10) shares[i] = (int)(d % 10);
    }
    int tot = 0, cum = 0;
    for (int i = 0; i < NP; ++i) tot += shares[i];
    for (int i = 0; i < tid; ++i) cum += shares[i];
    ebeg[tid] = ((2 * nb * cum + tot) / (2 * tot)) * kBatch;  // rounded cumulative share, in whole batches
  }
  for (int e = tid; e < nent_pad; e += Ws::kThreads) {
    unsigned w;
    if (e < ne) {
      int k = 1;
      while ((k + 1) * k / 2 <= e) ++k;
      const int l = e - k * (k - 1) / 2;
      w = (unsigned)k | ((unsigned)l << 8) | ((unsigned)e << 16);
    } else if (e < ne + m) {
      const int k = e - ne;
      w = (unsigned)k | ((unsigned)MC << 8) | ((unsigned)((Geo::OFF_CV - Geo::OFF_A) / 256 + k) << 16);
    } else {
      w = 0u | ((unsigned)MC << 8) | ((unsigned)((Geo::OFF_DV - Geo::OFF_A) / 256) << 16);  // padding -> dv row 0
    }
    tpl_pack_ent(w, &ent[e]);
  }
  for (int s = 0; s < Ws::kStages; ++s) {
    unsigned char *st = tpl_smem + s * Ws::kStageBytes;
    double *A = reinterpret_cast<double *>(st + Geo::OFF_A);
    double *cv = reinterpret_cast<double *>(st + Geo::OFF_CV);
    double *rr = reinterpret_cast<double *>(st + Ws::OFF_RR);
    for (int x = tid; x < (Geo::NE - ne) * 32; x += Ws::kThreads) A[ne * 32 + x] = 0.0;  // identity padding, m < MC
    for (int x = tid; x < (MC - m) * 32; x += Ws::kThreads) cv[m * 32 + x] = 0.0;
    for (int x = tid; x < (MC + 1) * 32; x += Ws::kThreads) rr[x] = 0.0;
  }
  for (int x = tid; x < (MC + 1) * 32; x += Ws::kThreads) ws_rr<MC, NP>(2)[x] = 0.0;
  __syncthreads();

  const long long ntiles = ((long long)P.loc_end - P.loc_begin + 31) / 32;
  int failed = INT_MAX;

  if (warp > 0) {
    // ================================ producers ========================================================
    const int pw = warp - 1;          // producer warp index
    const int pl = tid - 32;          // producer lane 0..32 NP - 1
    // Tiles are handed out dynamically (one atomic per tile): the first tiles of an ordering have far-apart
    // neighbours (few predecessors), which sends Matern through its slow branches -- with a static split the
    // CTAs that drew them finished 0.3-0.5 ms after everybody else.
    //
    // Software pipeline, so that no producer ever waits for global memory: while tile j is being built, the
    // sites/residuals of tile j+1 are in flight (cp.async into the other stage's site rows and the next slot
    // of a 3-deep residual ring), the neighbour ids of tile j+2 are in flight into registers, and the atomic
    // that claims tile j+3 is outstanding in thread 0.
    constexpr int IDS = (MC + NP - 1) / NP;  // neighbour slots per producer thread: k = pw + q NP, location = lane
    int ids[IDS];
    auto load_ids = [&](long long tile) {
      long long i = P.loc_begin + tile * 32 + lane;
      if (i >= P.loc_end) i = (long long)P.loc_end - 1;  // tail lanes redo the last location
      const int mi = site ? m : nn_count(i, m);
      const long long off = site ? 0 : nn_offset(i, m);
#pragma unroll
      for (int q = 0; q < IDS; ++q) {
        const int k = pw + q * NP;
        ids[q] = -1;  // missing neighbour (ragged start of the ordering)
        if (k < mi) ids[q] = site ? __ldg(&P.nn0[(size_t)k * P.q + i]) : __ldg(&P.nnIndx[off + k]);
      }
    };
    auto issue_gathers = [&](long long tile, int stage, int ring) {
      double2 *co = reinterpret_cast<double2 *>(tpl_smem + stage * Ws::kStageBytes + Geo::OFF_CO);
      double *rr = ws_rr<MC, NP>(ring);
      long long i = P.loc_begin + tile * 32 + lane;
      if (i >= P.loc_end) i = (long long)P.loc_end - 1;
#pragma unroll
      for (int q = 0; q < IDS; ++q) {
        const int k = pw + q * NP;
        if (k < m) {
          if (ids[q] >= 0) {
            cp_async16(&co[k * 32 + lane], &P.coords[ids[q]]);
            if (resid != nullptr) cp_async8(&rr[k * 32 + lane], &resid[ids[q]]);
          } else {
            // stand in the location's own site, so the masked entries stay on the straight-line correlation path
            cp_async16(&co[k * 32 + lane], &P.coords[i]);
            rr[k * 32 + lane] = 0.0;
          }
        }
      }
      if (pw == 0) {
        cp_async16(&co[MC * 32 + lane], site ? &P.sites[i] : &P.coords[i]);
        if (resid != nullptr && !site) cp_async8(&rr[MC * 32 + lane], &resid[i]);
      }
    };
    if (pl == 0) {  // first three tiles: static round-robin (spreads the slow head of the ordering over the CTAs)
      next_tile[0] = blockIdx.x;
      next_tile[1] = (long long)blockIdx.x + gridDim.x;
      next_tile[2] = (long long)blockIdx.x + 2ll * gridDim.x;
    }
    named_bar_sync(5, 32 * NP);
    if (next_tile[0] < ntiles) {
      load_ids(next_tile[0]);
      issue_gathers(next_tile[0], 0, 0);
    }
    if (next_tile[1] < ntiles) load_ids(next_tile[1]);
    int ring = 0;  // j % 3
    // Weights of a solved tile go to global memory from here, not from the consumer (the CTA's critical path):
    // 32 complete rows of B are contiguous, the producers copy them out of the stage before refilling it.
    auto write_b = [&](long long tile, int stage) {
      if (!kProducersStoreB || th.B == nullptr || site) return;
      const long long base = P.loc_begin + tile * 32;
      if (!ws_full_tile(base, m, P.loc_end)) return;  // ragged / partial tiles: the consumer stored them
      const double *cv = reinterpret_cast<const double *>(tpl_smem + stage * Ws::kStageBytes + Geo::OFF_CV);
      double *dst = th.B + nn_offset(base, m);
      const float rm = 1.0f / (float)m;
      for (int t = pl; t < 32 * m; t += 32 * NP) {
        const int loc = (int)(((float)t + 0.5f) * rm);
        const int k = t - loc * m;
        dst[t] = cv[k * 32 + loc];
      }
    };
    for (int j = 0;; ++j) {
      const int s = j & 1;
      if (j >= Ws::kStages) {
        named_bar_sync(3 + s, Ws::kThreads);  // consumer done with tile j-2: stage s, ring slot j+1
        write_b(next_tile[(j - 2) & 7], s);
      }
      cp_async_wait_all();         // this thread's share of tile j has landed
      named_bar_sync(5, 32 * NP);  // ... everybody's; build j-1 and the copy-out of tile j-2 finished everywhere
      const long long tile = next_tile[j & 7];
      if (tile >= ntiles) {
        if (j >= 1) {
          named_bar_sync(3 + (s ^ 1), Ws::kThreads);  // the CTA's last tile
          write_b(next_tile[(j - 1) & 7], s ^ 1);
        }
        named_bar_arrive(1 + s, Ws::kThreads);  // hands the consumer the end-of-work marker
        break;
      }
      const int ring_next = (ring == 2) ? 0 : ring + 1;
      const long long tile1 = next_tile[(j + 1) & 7], tile2 = next_tile[(j + 2) & 7];
      if (tile1 < ntiles) issue_gathers(tile1, s ^ 1, ring_next);
      if (tile2 < ntiles) load_ids(tile2);
      unsigned claimed = 0u;
      if (pl == 0) claimed = atomicAdd(th.tile_counter, 1u);  // tile 3 gridDim.x + claimed
      // ---- this warp's share of tile j's correlations (lane = location) ---------------------------------
      tpl_build<COV, kBatch, kIsMatern<COV> || kExpTab>(P, lane, s * Ws::kStageBytes + Geo::OFF_CO,
                                                        s * Ws::kStageBytes + Geo::OFF_A, ent, ebeg[pw], kBatch,
                                                        ebeg[pw + 1], s_tab);
      asm volatile("" ::: "memory");
      if (pl == 0) next_tile[(j + 3) & 7] = 3ll * gridDim.x + claimed;  // published by the rendezvous of the next trip
      named_bar_arrive(1 + s, Ws::kThreads);  // stage full
      ring = ring_next;
    }
  } else {
    // ================================ consumer ==========================================================
    for (int j = 0;; ++j) {
      const int s = j & 1;
      unsigned char *st = tpl_smem + s * Ws::kStageBytes;
      double *A = reinterpret_cast<double *>(st + Geo::OFF_A) + lane;
      double *cv = reinterpret_cast<double *>(st + Geo::OFF_CV);
      double *dv = reinterpret_cast<double *>(st + Geo::OFF_DV) + lane;
      const double *rr = ws_rr<MC, NP>(j % 3);
      named_bar_sync(1 + s, Ws::kThreads);  // wait for the producers
      const long long tile = next_tile[j & 7];
      if (tile >= ntiles) break;
      const long long base = P.loc_begin + tile * 32;
      const bool active = base + lane < P.loc_end;
      const long long i = active ? base + lane : (long long)P.loc_end - 1;
      const int mi = site ? m : nn_count(i, m);
      const long long off = site ? 0 : nn_offset(i, m);
      if (mi < m) {  // ragged start of the ordering: identity-pad rows >= m_i
        for (int k = mi; k < m; ++k) {
          cv[k * 32 + lane] = 0.0;
          for (int l = 0; l < k; ++l) A[(k * (k - 1) / 2 + l) * 32] = 0.0;
        }
      }
      double Fsum;
      // Cholesky block width: 4 columns where the consumer is the critical path and registers allow (exponential at
      // m = 15: W = 2 / 3 / 4 -> 0.642 / 0.629 / 0.620 ms; gaussian 0.600 -> 0.571).  Matern (168 registers with five
      // producers) keeps 2, and so do the small capacities, where W = 4 measured slower (m = 10: 0.307 -> 0.343 ms).
#ifdef NNGP_WS_CHOLW  // experiment builds
      constexpr int kCholW = kIsMatern<COV> ? 2 : NNGP_WS_CHOLW;
#else
      constexpr int kCholW = (!kIsMatern<COV> && MC >= 12 && MC <= 16) ? 4 : 2;
#endif
      const bool ok = tpl_solve<MC, kCholW>(A, cv + lane, dv, mi, diag, &Fsum);
      if (!ok && active) failed = min(failed, (int)min(i + 1, (long long)INT_MAX - 1));
      const double Fv = diag - Fsum;
      double part = 0.0;
      if (resid != nullptr) {
#pragma unroll
        for (int k = 0; k < MC; ++k) part = fma(cv[k * 32 + lane], rr[k * 32 + lane], part);
      }
      const double r_own = rr[MC * 32 + lane];
      if (th.B != nullptr) {
        if (site) {
          if (active) {
#pragma unroll
            for (int k = 0; k < MC; ++k)
              if (k < m) th.B[(size_t)k * P.q + i] = cv[k * 32 + lane];
          }
        } else {
          // 32 complete rows are contiguous in B: copied out by the producers (write_b) or, transposed, from here
          const bool full = ws_full_tile(base, m, P.loc_end);
          if (full && !kProducersStoreB) {
            __syncwarp();
            double *dst = th.B + nn_offset(base, m);
            const float rm = 1.0f / (float)m;
            for (int t = lane; t < 32 * m; t += 32) {
              const int loc = (int)(((float)t + 0.5f) * rm);
              const int k = t - loc * m;
              dst[t] = cv[k * 32 + loc];
            }
          } else if (!full && active) {
#pragma unroll
            for (int k = 0; k < MC; ++k)
              if (k < mi) th.B[off + k] = cv[k * 32 + lane];
          }
        }
      }
      __syncwarp();
      named_bar_arrive(3 + s, Ws::kThreads);  // stage free
      if (th.F != nullptr && active) th.F[i] = Fv;
      // per-tile partial sums (fixed lane order), so the grand total does not depend on which CTA drew the tile
      double lt = 0.0, qt = 0.0;
      if (!site && active) {
        lt = (Fv > 0.0) ? nngp_log_pos(Fv) : log(Fv);  // straight-line log on the normal path
        if (resid != nullptr) {
          const double e = r_own - part;
          qt = e * e * nngp_rcp_pos(Fv);
        }
      }
#pragma unroll
      for (int sft = 16; sft >= 1; sft >>= 1) {
        lt += __shfl_down_sync(0xffffffffu, lt, sft);
        qt += __shfl_down_sync(0xffffffffu, qt, sft);
      }
      if (lane == 0) {
        th.partials[2 * tile] = lt;
        th.partials[2 * tile + 1] = qt;
      }
    }
    if (failed != INT_MAX) atomicMin(th.fail, failed);
  }
  // ---- grid reduction: the last CTA to finish adds the per-tile partials in fixed order --------------------
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    const unsigned int tk = atomicAdd(th.ticket, 1u);
    last_flag = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (last_flag) {
    __threadfence();
    double a = 0.0, b = 0.0;
    for (long long t = tid; t < ntiles; t += Ws::kThreads) {  // thread t: tiles t, t + blockDim, ... in order
      a += __ldcg(&th.partials[2 * t]);
      b += __ldcg(&th.partials[2 * t + 1]);
    }
#pragma unroll
    for (int sft = 16; sft >= 1; sft >>= 1) {
      a += __shfl_down_sync(0xffffffffu, a, sft);
      b += __shfl_down_sync(0xffffffffu, b, sft);
    }
    if (lane == 0) {
      red[0][warp] = a;
      red[1][warp] = b;
    }
    __syncthreads();
    if (tid == 0) {
      double o0 = red[0][0], o1 = red[1][0];
#pragma unroll
      for (int w = 1; w <= NP; ++w) {
        o0 += red[0][w];
        o1 += red[1][w];
      }
      th.out[0] = o0;
      th.out[1] = o1;
      const int fv = atomicExch(th.fail, INT_MAX);
      th.out[2] = (fv != INT_MAX) ? 1.0 : 0.0;
      th.out[3] = 0.0;
      *th.fail_out = fv;
      *th.tile_counter = 0u;
      *th.ticket = 0u;
      __threadfence_system();  // results before the stamp, as seen by the polling host (and by peers' hosts)
      *reinterpret_cast<volatile double *>(&th.out[kBfOutSeq]) = th.seq;
    }
  }
}

// ---- launcher ---------------------------------------------------------------------------------------
int tpl_capacity_for_m(int m);

template <int MC, int COV, int NP>
static cudaError_t ws_launch_one(const BfLaunch &P, int grid_x, cudaStream_t stream) {
  using Ws = WsGeo<MC, NP>;
  static bool configured[16] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev >= 16 || !configured[dev]) {
    cudaError_t e = cudaFuncSetAttribute(bf_ws_kernel<MC, COV, NP>, cudaFuncAttributeMaxDynamicSharedMemorySize, Ws::kSmemBytes);
    if (e != cudaSuccess) return e;
    if (dev < 16) configured[dev] = true;
  }
  dim3 grid(grid_x, P.nbatch, 1);
  bf_ws_kernel<MC, COV, NP><<<grid, Ws::kThreads, Ws::kSmemBytes, stream>>>(P);
  return cudaGetLastError();
}

template <int MC>
static cudaError_t ws_launch_cov(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream) {
  switch (cov) {
    case NNGP_COV_EXPONENTIAL: return ws_launch_one<MC, NNGP_COV_EXPONENTIAL, NNGP_WS_NP>(P, grid_x, stream);
    case NNGP_COV_SPHERICAL: return ws_launch_one<MC, NNGP_COV_SPHERICAL, NNGP_WS_NP>(P, grid_x, stream);
    case NNGP_COV_MATERN:
      if (P.matern_short) return ws_launch_one<MC, NNGP_COV_MATERN_SHORT, NNGP_WS_MATERN_NP>(P, grid_x, stream);
      return ws_launch_one<MC, NNGP_COV_MATERN, NNGP_WS_MATERN_NP>(P, grid_x, stream);
    case NNGP_COV_GAUSSIAN: return ws_launch_one<MC, NNGP_COV_GAUSSIAN, NNGP_WS_NP>(P, grid_x, stream);
    default: return cudaErrorInvalidValue;
  }
}

static int ws_stage_bytes(int MC) {
  return (MC + 1) * 32 * 16 + (MC * (MC - 1) / 2) * 32 * 8 + 2 * MC * 32 * 8 + (MC + 1) * 32 * 8;
}

// Tile capacities instantiated for this kernel.  20 and 24 fit one CTA per SM (2 stages of 75 / 102 KB): one consumer per
// SM instead of two, still 2x the cooperative kernel at m = 20 (2.3 vs 5.0 ms per 10^6 locations).
int ws_capacity_for_m(int m) {
  static const int caps[7] = {8, 10, 12, 15, 16, 20, 24};
  for (int c : caps)
    if (m <= c) return c;
  return 0;  // not handled by this kernel
}

int ws_ctas_per_sm(int m) {
  const int MC = ws_capacity_for_m(m);
  int c = (227 * 1024) / (2 * ws_stage_bytes(MC) + (MC + 1) * 32 * 8 + 2048);
  static int override_ctas = -1;
  if (override_ctas < 0) {
    const char *e = getenv("NNGP_WS_CTAS");
    override_ctas = e ? atoi(e) : 0;
  }
  if (override_ctas > 0) c = override_ctas;
  if (c > 4) c = 4;
  return c < 1 ? 1 : c;
}

int ws_grid_for(int m, long long nloc, int sms) {
  long long need = (nloc + 31) / 32;
  long long g = (long long)sms * ws_ctas_per_sm(m);
  if (need < g) g = need;
  return (int)(g < 1 ? 1 : g);
}

cudaError_t ws_launch(const BfLaunch &P, int cov, int grid_x, cudaStream_t stream) {
  switch (ws_capacity_for_m(P.m)) {
#ifdef NNGP_WS_ONLY15  // experiment builds (tools/ab_bench.py): only the m = 13..15 capacity, for short compile times
    case 15: return ws_launch_cov<15>(P, cov, grid_x, stream);
    default: return cudaErrorInvalidValue;
#else
    case 8: return ws_launch_cov<8>(P, cov, grid_x, stream);
    case 10: return ws_launch_cov<10>(P, cov, grid_x, stream);
    case 12: return ws_launch_cov<12>(P, cov, grid_x, stream);
    case 15: return ws_launch_cov<15>(P, cov, grid_x, stream);
    case 16: return ws_launch_cov<16>(P, cov, grid_x, stream);
    case 20: return ws_launch_cov<20>(P, cov, grid_x, stream);
    case 24: return ws_launch_cov<24>(P, cov, grid_x, stream);
    default: return cudaErrorInvalidValue;
#endif
  }
}

}  // namespace nngp

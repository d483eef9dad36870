// tpl_device.cuh -- device building blocks shared by the one-lane-per-location kernels
// (bf_tpl_kernel.cu: every warp does all phases; bf_ws_kernel.cu: warp-specialised producer/consumer).
// All per-location state lives in shared memory laid out [entry][lane] (32 lanes = 32 locations).
#pragma once
#include <limits.h>

#include "cov.cuh"

namespace nngp {

template <int MC>
struct TplGeo {
  static constexpr int NE = MC * (MC - 1) / 2;
  static constexpr int OFF_CO = 0;                          // bytes; row MC holds the location's own site
  static constexpr int OFF_A = OFF_CO + (MC + 1) * 32 * 16;
  static constexpr int OFF_CV = OFF_A + NE * 32 * 8;
  static constexpr int OFF_DV = OFF_CV + MC * 32 * 8;
  static constexpr int kWarpBytes = OFF_DV + MC * 32 * 8;
};

extern __shared__ __align__(16) unsigned char tpl_smem[];

// ---- phase A: correlations ----------------------------------------------------------------------
// Work list shared by all lanes: entry w = k | l << 8 | dst << 16 means "rho between co[k] and co[l]
// goes to row dst (256-byte rows counted from the start of A)".  It holds the m(m-1)/2 pairs of the
// strict lower triangle (dst = packed index), the m entries of c (l = MC: the location's own site,
// dst = a row of cv) and is padded to a multiple of the batch width with harmless duplicates that
// write into dv (overwritten by the factorisation afterwards).
constexpr int kTplMaxEnt = 320;  // m(m-1)/2 + m entries at m = 24, padded to a batch
// Matern keeps the entries pre-multiplied in shared memory: x = (k * 512) | (l * 512) << 16 (byte offsets of the two
// site rows), y = dst * 256 (byte offset of the destination row) -- 3 fewer integer instructions per entry in its
// issue-bound producers (0.95 -> 0.89 ms together with the integer slow-path test).  The other models keep the 32-bit
// word: their kernel is consumer-bound and the wider list only adds shared-memory traffic (0.64 -> 0.69 ms measured).
template <int COV>
struct TplEntSel { typedef unsigned type; };
template <>
struct TplEntSel<NNGP_COV_MATERN> { typedef uint2 type; };
template <>
struct TplEntSel<NNGP_COV_MATERN_SHORT> { typedef uint2 type; };
template <int COV>
using TplEnt = typename TplEntSel<COV>::type;

__device__ __forceinline__ void tpl_pack_ent(unsigned w, unsigned *out) { *out = w; }
__device__ __forceinline__ void tpl_pack_ent(unsigned w, uint2 *out) {
  *out = make_uint2(((w & 0xffu) * 512u) | ((((w >> 8) & 0xffu) * 512u) << 16), (w >> 16) * 256u);
}
// byte offsets of the two site rows and of the destination row
__device__ __forceinline__ void tpl_decode_ent(unsigned w, unsigned &ok, unsigned &ol, unsigned &od) {
  ok = (w & 0xffu) * 512u;
  ol = ((w >> 8) & 0xffu) * 512u;
  od = (w >> 16) * 256u;
}
__device__ __forceinline__ void tpl_decode_ent(uint2 w, unsigned &ok, unsigned &ol, unsigned &od) {
  ok = w.x & 0xffffu;
  ol = w.x >> 16;
  od = w.y;
}
#ifndef NNGP_OTHER_BATCH
#define NNGP_OTHER_BATCH 8  // exponential / spherical / gaussian (4 / 6 / 12 measured 0.628 / 0.629 / 0.659 vs 0.619 ms)
#endif
#ifndef NNGP_MATERN_BATCH
#define NNGP_MATERN_BATCH 4  // 6 and 8 measured slower (0.97 / 0.96 vs 0.95 ms): the register file is full at 4 chains
#endif
template <int COV>
constexpr int kTplBatch = kIsMatern<COV> ? NNGP_MATERN_BATCH : NNGP_OTHER_BATCH;  // entries evaluated together (independent chains)

// Batches of UNR entries: all loads, then UNR independent evaluation chains, then all stores.
// (A plain unrolled loop is serialised by ptxas: it cannot prove that the store of entry e does not
// alias the site loads of entry e+1, so the chains never overlap -- seen in SASS.)
template <int COV, int UNR, bool TAB, typename CorT>
__device__ __forceinline__ void tpl_build_body(const CorT &cor, int lane, int off_co, int off_a, const TplEnt<COV> *ent,
                                               int e_begin, int e_step, int nent_pad, double sigmaSq) {
  const unsigned char *cob = tpl_smem + off_co + lane * 16;
  unsigned char *ab = tpl_smem + off_a + lane * 8;
  for (int e0 = e_begin; e0 < nent_pad; e0 += e_step) {  // e_begin, e_step: multiples of the batch width
    double d2[UNR], v[UNR];
    unsigned dst[UNR];
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      unsigned ok, ol;
      tpl_decode_ent(ent[e0 + u], ok, ol, dst[u]);
      const double2 pk = *reinterpret_cast<const double2 *>(cob + ok);
      const double2 pl = *reinterpret_cast<const double2 *>(cob + ol);
      const double dx = pk.x - pl.x, dy = pk.y - pl.y;
      d2[u] = dx * dx + dy * dy;
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) {
      if constexpr (TAB) v[u] = sigmaSq * cor.fast_tab(d2[u]);
      else v[u] = sigmaSq * cor.fast(d2[u]);
    }
    if constexpr (kIsMatern<COV>) {  // rare arguments outside the straight-line path (phi D > 2, integer nu)
      bool any = false;
#pragma unroll
      for (int u = 0; u < UNR; ++u) any |= cor.maybe_slow(d2[u]);
      if (any) {
#pragma unroll
        for (int u = 0; u < UNR; ++u)
          if (cor.needs_slow(d2[u])) v[u] = sigmaSq * cor.slow(d2[u]);
      }
    }
#pragma unroll
    for (int u = 0; u < UNR; ++u) *reinterpret_cast<double *>(ab + dst[u]) = v[u];
  }
}

// The parameter set of this CTA is P.th[blockIdx.y]; Matern specialises on the index so that its series
// coefficients become constant-bank operands with static addresses (see Cor<NNGP_COV_MATERN>).
// tab: exp / log tables in shared memory (Matern only; NULL: straight-line log / exp)
template <int COV, int UNR, bool TAB = false>
__device__ __forceinline__ void tpl_build(const BfLaunch &P, int lane, int off_co, int off_a, const TplEnt<COV> *ent,
                                          int e_begin, int e_step, int nent_pad, const double *tab = nullptr) {
  if constexpr (kIsMatern<COV>) {
    switch (blockIdx.y) {  // uniform over the CTA
      case 0: tpl_build_body<COV, UNR, TAB>(Cor<COV>(P.th[0], tab), lane, off_co, off_a, ent, e_begin, e_step, nent_pad, P.th[0].sigmaSq); break;
      case 1: tpl_build_body<COV, UNR, TAB>(Cor<COV>(P.th[1], tab), lane, off_co, off_a, ent, e_begin, e_step, nent_pad, P.th[1].sigmaSq); break;
      case 2: tpl_build_body<COV, UNR, TAB>(Cor<COV>(P.th[2], tab), lane, off_co, off_a, ent, e_begin, e_step, nent_pad, P.th[2].sigmaSq); break;
      default: tpl_build_body<COV, UNR, TAB>(Cor<COV>(P.th[3], tab), lane, off_co, off_a, ent, e_begin, e_step, nent_pad, P.th[3].sigmaSq); break;
    }
  } else {
    const BfTheta &th = P.th[blockIdx.y];
    tpl_build_body<COV, UNR, TAB>(Cor<COV>(th, tab), lane, off_co, off_a, ent, e_begin, e_step, nent_pad, th.sigmaSq);
  }
}

// ---- phase B: Cholesky + solves, fully unrolled -----------------------------------------------------
// Left-looking C = V V' W columns at a time; V(i,j) overwrites A, the right-hand side rides along
// as row MC (cv[j] becomes y_j = (V^-1 c)_j), 1/V(j,j) goes to dv[j].  Then V' B = y.
// On return: Bout[0..MC) kriging weights, *Fsum = c' C^-1 c = |y|^2; false if a pivot is not positive.
#define V_(i, j) A[((i) * ((i)-1) / 2 + (j)) * 32]

// columns J0 .. J0+W-1 (compile-time), then recurse; the memory clobber between blocks keeps ptxas from
// hoisting the next block's loads across this block's stores (which would blow up register pressure).
// One load of V(i,k) feeds W FMAs: W = 2 needs 336 loads per location at MC = 15, W = 3 needs 210.
template <int MC, int J0, int W>
__device__ __forceinline__ void tpl_chol_block(double *A, double *cv, double *dv, int mi, double diag, double &Fsum,
                                               bool &ok) {
  constexpr int NW = (J0 + W <= MC) ? W : (MC - J0);  // columns in this block
  double a[NW][MC + 1];  // a[c][i]: rows i = J0+c .. MC of column J0+c (row MC: right-hand side)
#pragma unroll
  for (int c = 0; c < NW; ++c) {
#pragma unroll
    for (int i = J0 + c; i <= MC; ++i) {
      if (i == J0 + c) a[c][i] = (J0 + c < mi) ? diag : 1.0;
      else if (i < MC) a[c][i] = V_(i, J0 + c);
      else a[c][i] = cv[(J0 + c) * 32];
    }
  }
#ifdef NNGP_CHOL_KROLL  // experiment builds: k loop rolled in steps of NNGP_CHOL_KROLL (code 3x smaller)
  constexpr int kKRoll = NNGP_CHOL_KROLL;
#pragma unroll kKRoll
#else
#pragma unroll  // (a rolled k loop, unroll 2 or 4, shrinks the code 3x but measured 3-8% slower: less load hoisting)
#endif
  for (int k = 0; k < J0; ++k) {
    double vc[NW];
#pragma unroll
    for (int c = 0; c < NW; ++c) vc[c] = V_(J0 + c, k);
#pragma unroll
    for (int r = 0; r < NW; ++r) {  // the triangle inside the block
#pragma unroll
      for (int c = 0; c <= r; ++c) a[c][J0 + r] = fma(-vc[r], vc[c], a[c][J0 + r]);
    }
#pragma unroll
    for (int i = J0 + NW; i <= MC; ++i) {  // remaining rows; row MC = right-hand side
      const double v = (i < MC) ? V_(i, k) : cv[k * 32];
#pragma unroll
      for (int c = 0; c < NW; ++c) a[c][i] = fma(-v, vc[c], a[c][i]);
    }
  }
#pragma unroll
  for (int c = 0; c < NW; ++c) {
#pragma unroll
    for (int cc = 0; cc < c; ++cc) {  // columns of this block already finished, still in registers
      const double v = a[cc][J0 + c];
#pragma unroll
      for (int i = J0 + c; i <= MC; ++i) a[c][i] = fma(-a[cc][i], v, a[c][i]);
    }
    const double d = a[c][J0 + c];
    if (J0 + c < mi && !(d > 0.0)) ok = false;
    const double r = nngp_rsqrt_pos(d);
    dv[(J0 + c) * 32] = r;
#pragma unroll
    for (int i = J0 + c + 1; i <= MC; ++i) a[c][i] *= r;
    Fsum = fma(a[c][MC], a[c][MC], Fsum);
    cv[(J0 + c) * 32] = a[c][MC];
#pragma unroll
    for (int i = J0 + c + 1; i < MC; ++i) V_(i, J0 + c) = a[c][i];
  }
  asm volatile("" ::: "memory");
  if constexpr (J0 + NW < MC) tpl_chol_block<MC, J0 + NW, W>(A, cv, dv, mi, diag, Fsum, ok);
}

// On return cv[k] holds the kriging weight B_k (k < MC), *Fsum_out = c' C^-1 c; false if a pivot <= 0.
#ifndef NNGP_CHOL_W
#define NNGP_CHOL_W 2
#endif
template <int MC, int W = NNGP_CHOL_W>
__device__ __forceinline__ bool tpl_solve(double *A, double *cv, double *dv, int mi, double diag, double *Fsum_out) {
  double Fsum = 0.0;
  bool ok = true;
  tpl_chol_block<MC, 0, W>(A, cv, dv, mi, diag, Fsum, ok);
  // back substitution V' B = y: B_k = (y_k - sum_{i>k} V(i,k) B_i) / V(k,k)
  double w[MC];
#pragma unroll
  for (int k = MC - 1; k >= 0; --k) {
    double acc0 = cv[k * 32], acc1 = 0.0;  // two chains: halves the dependent FMA depth of the sweep
#pragma unroll
    for (int i = MC - 1; i > k; --i) {
      if ((MC - 1 - i) & 1) acc1 = fma(-V_(i, k), w[i], acc1);
      else acc0 = fma(-V_(i, k), w[i], acc0);
    }
    w[k] = (acc0 + acc1) * dv[k * 32];
  }
#pragma unroll
  for (int k = 0; k < MC; ++k) cv[k * 32] = w[k];
  *Fsum_out = Fsum;
  return ok;
}
#undef V_

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }


__device__ __forceinline__ void cp_async8(void *smem_dst, const void *gsrc) {
  const unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(s), "l"(gsrc) : "memory");
}

}  // namespace nngp

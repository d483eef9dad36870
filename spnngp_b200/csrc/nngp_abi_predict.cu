// nngp_abi_predict.cu -- C ABI for the kNN + solve prediction paths (include/nngp_b200.h):
//   nngp_rnngp_predict  replaces `rNNGPPredict` (reference src/rNNGPPredict.cpp:25-222)
//   nngp_cnngp          replaces `cNNGP`        (reference src/cNNGP.cpp:73-357)
// The m x m solves run in the B/F kernels' site mode; sites are processed in chunks so that the
// q x nSamples normals and predictions never have to be resident at once.
#include <limits.h>
#include <math.h>
#include <string.h>

#include <algorithm>
#include <map>
#include <thread>
#include <vector>

#include "host_linalg.h"
#include "nngp_ctx.h"
#include "host_stage.h"

namespace nngp {
cudaError_t site_reduce_launch(const double *W, const int *nn0, const double *X, const double *y, const double *X0,
                               int n, int p, int m, int cs, double *a, double *U, cudaStream_t stream);
cudaError_t predict_emit_launch(const double *a, const double *U, const double *F, const double *beta, const double *z,
                                const int *samples, int nsamp, int p, int cs, int S, double *y0c, cudaStream_t stream);
cudaError_t conj_emit_launch(const double *a, const double *U, const double *F, const double *betaHat, const double *V,
                             int p, int cs, double bIG, double aIG, double *y0Hat, double *y0HatVar,
                             cudaStream_t stream);
size_t conj_score_scratch_bytes();
cudaError_t conj_score_launch(const double *y0, const double *y0Hat, const double *y0HatVar, int cs, void *scratch,
                              double *out, cudaStream_t stream);

namespace {

typedef DevBuf Dev;

// device-side description of one chunk of query sites
struct SiteChunk {
  int cs = 0;
  Dev sites, nn0, X0, W, F, a, U;
};

int upload_chunk(nngp_ctx *c, SiteChunk &ch, int q, int i0, int i1, const double *X0, const double *coords0,
                 const int *nnIndx0) {
  const int cs = i1 - i0, m = c->m, p = c->p;
  ch.cs = cs;
  std::vector<double2> xy((size_t)cs);
  for (int i = 0; i < cs; i++) xy[i] = make_double2(coords0[i0 + i], coords0[(size_t)q + i0 + i]);
  std::vector<int> nn((size_t)cs * m);
  for (int k = 0; k < m; k++) {
    const int *src = nnIndx0 + (size_t)k * q + i0;
    for (int i = 0; i < cs; i++) {
      if (src[i] < 0 || src[i] >= c->n) return ctx_fail(c, NNGP_EINVAL, "prediction neighbour index out of range");
      nn[(size_t)k * cs + i] = src[i];
    }
  }
  std::vector<double> x0((size_t)cs * (p > 0 ? p : 1));
  for (int cc = 0; cc < p; cc++) memcpy(&x0[(size_t)cc * cs], X0 + (size_t)cc * q + i0, (size_t)cs * sizeof(double));
  NNGP_CUDA(c, ch.sites.alloc((size_t)cs * sizeof(double2)));
  NNGP_CUDA(c, ch.nn0.alloc((size_t)cs * m * sizeof(int)));
  NNGP_CUDA(c, ch.X0.alloc((size_t)cs * (p > 0 ? p : 1) * sizeof(double)));
  NNGP_CUDA(c, ch.W.alloc((size_t)cs * m * sizeof(double)));
  NNGP_CUDA(c, ch.F.alloc((size_t)cs * sizeof(double)));
  NNGP_CUDA(c, ch.a.alloc((size_t)cs * sizeof(double)));
  NNGP_CUDA(c, ch.U.alloc((size_t)cs * (p > 0 ? p : 1) * sizeof(double)));
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));  // recycled pool blocks: nothing earlier may still be using them
  NNGP_CUDA(c, cudaMemcpy(ch.sites.p, xy.data(), (size_t)cs * sizeof(double2), cudaMemcpyHostToDevice));
  NNGP_CUDA(c, cudaMemcpy(ch.nn0.p, nn.data(), (size_t)cs * m * sizeof(int), cudaMemcpyHostToDevice));
  if (p > 0) NNGP_CUDA(c, cudaMemcpy(ch.X0.p, x0.data(), (size_t)cs * p * sizeof(double), cudaMemcpyHostToDevice));
  return NNGP_OK;
}

// kriging weights W (cs x m column-major) and conditional variances F for theta = (s2, t2, phi, nu),
// followed by the per-site reductions a = w.y_N and U = x0 - X_N'w.  Enqueues only: `mc` / `d_mc` are the Matern tables
// of this theta on the host and (already uploaded) on the device, `fail_dev` the device view of a mapped host int that
// receives the 1-based failing site (INT_MAX: none) when the launch has finished.
int site_solve_enqueue(nngp_ctx *c, int cov, const double *t, SiteChunk &ch, const MaternConsts *mc, const MaternConsts *d_mc,
                       int *fail_dev) {
  BfLaunch P;
  memset(&P, 0, sizeof P);
  P.coords = c->d_coords;
  P.n = c->n;
  P.m = c->m;
  P.loc_begin = 0;
  P.loc_end = ch.cs;
  P.nbatch = 1;
  P.site_mode = 1;
  P.sites = ch.sites.as<double2>();
  P.nn0 = ch.nn0.as<int>();
  P.q = ch.cs;
  BfTheta &th = P.th[0];
  th.sigmaSq = t[0];
  th.tauSq = t[1];
  th.phi = t[2];
  th.mat = nullptr;
  if (cov == NNGP_COV_MATERN) {
    matern_series_params(*mc, &th.ms);
    matern_set_phi(&th.ms, t[2]);
    th.mat = d_mc;
  }
  double *out_dev = nullptr;
  NNGP_CUDA(c, cudaHostGetDevicePointer((void **)&out_dev, c->h_out, 0));
  th.B = ch.W.as<double>();
  th.F = ch.F.as<double>();
  th.resid = nullptr;
  th.out = out_dev;  // 3 doubles, unused in site mode
  th.partials = c->d_partials;
  th.ticket = c->d_ticket;
  th.tile_counter = c->d_ticket + NNGP_MAX_BATCH;
  th.fail = c->d_fail;
  th.fail_out = fail_dev;
  const int grid = bf_grid_any(c->m, cov, ch.cs, c->sms);
  NNGP_CUDA(c, bf_launch(P, cov, grid, c->stream));
  NNGP_CUDA(c, site_reduce_launch(ch.W.as<double>(), ch.nn0.as<int>(), c->d_X, c->d_y, ch.X0.as<double>(), c->n, c->p,
                                  c->m, ch.cs, ch.a.as<double>(), ch.U.as<double>(), c->stream));
  c->launches += 2;
  return NNGP_OK;
}

// The same, synchronous (the conjugate model's one solve per parameter set): tables through the context's slot 0.
int site_solve(nngp_ctx *c, int cov, const double *t, SiteChunk &ch) {
  if (cov == NNGP_COV_MATERN) {
    matern_prepare(t[3], &c->h_mat[0]);
    NNGP_CUDA(c, cudaMemcpyAsync(c->d_mat, c->h_mat, sizeof(MaternConsts), cudaMemcpyHostToDevice, c->stream));
    c->d_mat_nu[0] = t[3];  // what slot 0 of the device tables holds now (prepare_bf skips uploads it believes current)
  }
  int *fail_dev = nullptr;
  NNGP_CUDA(c, cudaHostGetDevicePointer((void **)&fail_dev, c->h_fail, 0));
  int rc = site_solve_enqueue(c, cov, t, ch, c->h_mat, c->d_mat, fail_dev);
  if (rc) return rc;
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
  if (c->h_fail[0] != INT_MAX) return ctx_fail(c, c->h_fail[0], "c++ error: dpotrf failed (prediction neighbour covariance)");
  return NNGP_OK;
}

}  // namespace
}  // namespace nngp

using namespace nngp;

extern "C" {

int nngp_rnngp_predict(nngp_ctx *c, int cov, int q, const double *X0, const double *coords0, const int *nnIndx0,
                       int nSamples, const double *betaSamples, const double *thetaSamples, const double *z,
                       double *y0, int site_begin, int site_end) {
  if (!c) return NNGP_EINVAL;
  if (c->n == 0 || c->p == 0) return ctx_fail(c, NNGP_ESTATE, "nngp_rnngp_predict before nngp_set_data");
  if (q < 1 || nSamples < 1 || !X0 || !coords0 || !nnIndx0 || !betaSamples || !thetaSamples || !z || !y0)
    return ctx_fail(c, NNGP_EINVAL, "nngp_rnngp_predict: bad arguments");
  if (cov < 0 || cov > 3) return ctx_fail(c, NNGP_EINVAL, "cov.model is not correctly specified");
  if (site_begin < 0 || site_end > q || site_begin > site_end) return ctx_fail(c, NNGP_EINVAL, "bad site range");
  NNGP_CUDA(c, cudaSetDevice(c->device));
  ScratchScope scratch(c->stream);
  const int p = c->p, S = nSamples;
  const int nTheta = (cov == NNGP_COV_MATERN) ? 4 : 3;

  // group the samples by distinct theta: rejected Metropolis proposals repeat the previous column
  std::map<std::vector<double>, int> index;
  std::vector<std::vector<int>> groups;
  std::vector<std::vector<double>> thetas;
  for (int s = 0; s < S; s++) {
    std::vector<double> key(thetaSamples + (size_t)s * nTheta, thetaSamples + (size_t)(s + 1) * nTheta);
    auto it = index.find(key);
    if (it == index.end()) {
      index.emplace(key, (int)groups.size());
      groups.emplace_back(1, s);
      key.resize(4, 0.0);
      thetas.push_back(key);
    } else {
      groups[it->second].push_back(s);
    }
  }

  Dev d_beta, d_samples;
  NNGP_CUDA(c, d_beta.alloc((size_t)p * S * sizeof(double)));
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));  // pool blocks are recycled in stream order; the copies below are not
  NNGP_CUDA(c, cudaMemcpy(d_beta.p, betaSamples, (size_t)p * S * sizeof(double), cudaMemcpyHostToDevice));
  std::vector<int> flat;
  std::vector<int> gofs(groups.size() + 1, 0);
  for (size_t g = 0; g < groups.size(); g++) {
    gofs[g + 1] = gofs[g] + (int)groups[g].size();
    flat.insert(flat.end(), groups[g].begin(), groups[g].end());
  }
  NNGP_CUDA(c, d_samples.alloc(flat.size() * sizeof(int)));
  NNGP_CUDA(c, cudaMemcpy(d_samples.p, flat.data(), flat.size() * sizeof(int), cudaMemcpyHostToDevice));

  // chunk of sites: z and y0 of a chunk are 2 x cs x S doubles, both double-buffered (4 x 1 GiB at most)
  long long cs_max = ((long long)1 << 30) / ((long long)8 * S);
  if (cs_max < 1024) cs_max = 1024;
  if (cs_max > 1 << 20) cs_max = 1 << 20;
  if (const char *e = getenv("NNGP_PREDICT_CHUNK")) {  // tests: many small chunks
    const long long v = atoll(e);
    if (v >= 32 && v < cs_max) cs_max = v;
  }
  if (cs_max > site_end - site_begin) cs_max = std::max(1, site_end - site_begin);
  // Full duplex, and nothing but launches between two chunk boundaries: the normals of chunk k + 1 arrive (first pinned
  // ring, or straight from the caller's page-locked array) and the draws of chunk k - 1 leave (second ring) on their own
  // threads while the solves and draws of chunk k run; src/rNNGPPredict.cpp:131-137,176-181 are these two q x S
  // streams.  The Matern tables of every distinct theta are prepared and uploaded once, every launch reports its failing
  // site into its own mapped slot, and the stream is synchronised once per chunk.
  if (!c->stager) {
    c->stager = new HostStager();
    NNGP_CUDA(c, c->stager->init(c->device));
  }
  if (!c->stager2) {
    c->stager2 = new HostStager();
    NNGP_CUDA(c, c->stager2->init(c->device));
  }
  const char *dup_env = getenv("NNGP_PREDICT_DUPLEX");
  const bool duplex = !(dup_env && dup_env[0] == '0');
  const int G = (int)groups.size();
  std::vector<MaternConsts> h_mc;
  Dev d_mc;
  if (cov == NNGP_COV_MATERN) {
    h_mc.resize((size_t)G);
    for (int g = 0; g < G; g++) matern_prepare(thetas[g][3], &h_mc[g]);
    NNGP_CUDA(c, d_mc.alloc((size_t)G * sizeof(MaternConsts)));
    NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
    NNGP_CUDA(c, cudaMemcpy(d_mc.p, h_mc.data(), (size_t)G * sizeof(MaternConsts), cudaMemcpyHostToDevice));
  }
  struct Mapped {  // one fail slot per distinct theta, written by the kernels through the mapped pointer
    int *h = nullptr;
    ~Mapped() { if (h) cudaFreeHost(h); }
  } failg;
  NNGP_CUDA(c, cudaHostAlloc((void **)&failg.h, (size_t)G * sizeof(int), cudaHostAllocMapped));
  int *failg_dev = nullptr;
  NNGP_CUDA(c, cudaHostGetDevicePointer((void **)&failg_dev, failg.h, 0));
  for (int g = 0; g < G; g++) failg.h[g] = INT_MAX;

  Dev d_zb[2], d_y0b[2];
  for (int b = 0; b < 2; b++) {
    NNGP_CUDA(c, d_zb[b].alloc((size_t)cs_max * S * sizeof(double)));
    NNGP_CUDA(c, d_y0b[b].alloc((size_t)cs_max * S * sizeof(double)));
  }
  NNGP_CUDA(c, cudaStreamSynchronize(c->stream));  // pool blocks are recycled in stream order; the staged copies are not
  std::thread dl, ul;
  cudaError_t dl_err = cudaSuccess, ul_err = cudaSuccess;
  struct Joiner {  // never leave a transfer thread running behind an early return
    std::thread &t;
    ~Joiner() { if (t.joinable()) t.join(); }
  } join_dl{dl}, join_ul{ul};
  HostStager *st1 = c->stager, *st2 = c->stager2;
  auto start_upload = [&](int k, int i0, int cs) {
    void *dst = d_zb[k & 1].p;
    ul = std::thread([=, &ul_err] {
      std::vector<StageJob> up{{StageJob::RAW, dst, z + (size_t)i0 * S, (size_t)cs * S * sizeof(double), 0}};
      ul_err = st1->run(up, nullptr);
    });
  };
  const int first_cs = (int)std::min<long long>(cs_max, site_end - site_begin);
  if (site_end > site_begin) start_upload(0, site_begin, first_cs);
  int kchunk = 0;
  for (int i0 = site_begin; i0 < site_end; i0 += (int)cs_max, ++kchunk) {
    const int i1 = (int)std::min<long long>(site_end, (long long)i0 + cs_max);
    const int cs = i1 - i0;
    Dev &d_z = d_zb[kchunk & 1];
    Dev &d_y0 = d_y0b[kchunk & 1];  // its previous content left with the download of chunk k - 2
    SiteChunk ch;
    int rc = upload_chunk(c, ch, q, i0, i1, X0, coords0, nnIndx0);
    if (rc) return rc;
    if (ul.joinable()) ul.join();  // the normals of this chunk
    NNGP_CUDA(c, ul_err);
    if (i1 < site_end) {  // the next chunk's normals travel while this one is solved (its buffer was read by chunk k - 1)
      start_upload(kchunk + 1, i1, (int)std::min<long long>(cs_max, site_end - i1));
      if (!duplex) {
        ul.join();
        NNGP_CUDA(c, ul_err);
      }
    }
    for (int g = 0; g < G; g++) {
      rc = site_solve_enqueue(c, cov, thetas[g].data(), ch, h_mc.empty() ? nullptr : &h_mc[g],
                              h_mc.empty() ? nullptr : d_mc.as<MaternConsts>() + g, failg_dev + g);
      if (rc) return rc;
      NNGP_CUDA(c, predict_emit_launch(ch.a.as<double>(), ch.U.as<double>(), ch.F.as<double>(), d_beta.as<double>(),
                                       d_z.as<double>(), d_samples.as<int>() + gofs[g], (int)groups[g].size(), p, cs, S,
                                       d_y0.as<double>(), c->stream));
      c->launches++;
    }
    NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
    for (int g = 0; g < G; g++)
      if (failg.h[g] != INT_MAX) return ctx_fail(c, failg.h[g], "c++ error: dpotrf failed (prediction neighbour covariance)");
    if (dl.joinable()) dl.join();  // the download of chunk k - 1 ran alongside this chunk's solves
    NNGP_CUDA(c, dl_err);
    // y0 is q x S column-major on the host: column s of the chunk is contiguous on both sides
    const double *src = d_y0.as<double>();
    dl = std::thread([=, &dl_err] {
      std::vector<StageJob> down((size_t)S);
      for (int s = 0; s < S; s++)
        down[s] = {StageJob::RAW, y0 + (size_t)s * q + i0, src + (size_t)s * cs, (size_t)cs * sizeof(double), 0};
      dl_err = st2->run_d2h(down);
    });
    if (!duplex) dl.join();  // NNGP_PREDICT_DUPLEX=0: one direction at a time (A/B switch)
  }
  if (dl.joinable()) dl.join();
  NNGP_CUDA(c, dl_err);
  return NNGP_OK;
}

// Shared body of nngp_cnngp and nngp_cnngp_score.  Score mode (y0Obs != NULL): the predictions stay on the device,
// scoreSums gets 2 doubles per parameter set -- sum of the CRPS terms, sum of squared errors -- and betaHat / ab /
// bBInv / bb / y0Hat / y0HatVar may be NULL.
static int cnngp_impl(nngp_ctx *c, int cov, int g, const double *thetaAlpha, const double sigmaSqIG[2], int n0,
                      const double *X0, const double *coords0, const int *nnIndx0, double *betaHat, double *ab,
                      double *bBInv, double *bb, double *y0Hat, double *y0HatVar, const double *y0Obs,
                      double *scoreSums) {
  if (!c) return NNGP_EINVAL;
  if (c->n == 0 || c->p == 0) return ctx_fail(c, NNGP_ESTATE, "nngp_cnngp before nngp_set_data");
  const bool score = (y0Obs != nullptr);
  if (g < 1 || !thetaAlpha || !sigmaSqIG) return ctx_fail(c, NNGP_EINVAL, "nngp_cnngp: bad arguments");
  if (!score && (!betaHat || !ab || !bBInv || !bb)) return ctx_fail(c, NNGP_EINVAL, "nngp_cnngp: bad arguments");
  if (score && (n0 < 1 || !scoreSums)) return ctx_fail(c, NNGP_EINVAL, "nngp_cnngp_score: bad arguments");
  if (n0 > 0 && (!X0 || !coords0 || !nnIndx0 || (!score && (!y0Hat || !y0HatVar))))
    return ctx_fail(c, NNGP_EINVAL, "nngp_cnngp: prediction inputs missing");
  if (cov < 0 || cov > 3) return ctx_fail(c, NNGP_EINVAL, "cov.model is not correctly specified");
  NNGP_CUDA(c, cudaSetDevice(c->device));
  ScratchScope scratch(c->stream);
  const int p = c->p, K = p + 1, n = c->n;
  const int nTheta = (cov == NNGP_COV_MATERN) ? 3 : 2;
  std::vector<double> S((size_t)K * K), V((size_t)p * p), tmp_p(p);
  Dev d_bh, d_V, d_y0, d_y0v, d_obs, d_sc, d_scratch;
  NNGP_CUDA(c, d_bh.alloc(p * sizeof(double)));
  NNGP_CUDA(c, d_V.alloc((size_t)p * p * sizeof(double)));
  std::vector<double> bh_local(p);
  if (score) {
    NNGP_CUDA(c, d_obs.alloc((size_t)n0 * sizeof(double)));
    NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
    NNGP_CUDA(c, cudaMemcpy(d_obs.p, y0Obs, (size_t)n0 * sizeof(double), cudaMemcpyHostToDevice));
    NNGP_CUDA(c, d_sc.alloc((size_t)2 * g * sizeof(double)));
    NNGP_CUDA(c, cudaMemsetAsync(d_sc.p, 0, (size_t)2 * g * sizeof(double), c->stream));
    NNGP_CUDA(c, d_scratch.alloc(conj_score_scratch_bytes()));
    NNGP_CUDA(c, cudaMemsetAsync(d_scratch.p, 0, conj_score_scratch_bytes(), c->stream));
  }

  const int cs_max = 1 << 20;
  std::vector<SiteChunk> chunks;  // the query sites do not change across parameter sets: upload once
  if (n0 > 0) {
    const int nchunks = (n0 + cs_max - 1) / cs_max;
    chunks.resize(nchunks);
    for (int k = 0; k < nchunks; k++) {
      const int i0 = k * cs_max, i1 = std::min(n0, i0 + cs_max);
      int rc = upload_chunk(c, chunks[k], n0, i0, i1, X0, coords0, nnIndx0);
      if (rc) return rc;
    }
    NNGP_CUDA(c, d_y0.alloc((size_t)std::min(n0, cs_max) * sizeof(double)));
    NNGP_CUDA(c, d_y0v.alloc((size_t)std::min(n0, cs_max) * sizeof(double)));
  }

  for (int k = 0; k < g; k++) {  // src/cNNGP.cpp:215-316
    const double phi = thetaAlpha[(size_t)k * nTheta], alpha = thetaAlpha[(size_t)k * nTheta + 1];
    const double nu = (cov == NNGP_COV_MATERN) ? thetaAlpha[(size_t)k * nTheta + 2] : 0.0;
    const double theta[4] = {1.0, alpha, phi, nu};  // updateConjBF == updateBF with (s2, t2) = (1, alpha)
    double out2[2];
    int slot0 = 0;
    int rc = nngp_bf_loglik(c, cov, 1, theta, nullptr, &slot0, out2);
    if (rc) return rc;
    rc = nngp_q_forms(c, 0, S.data());
    if (rc) return rc;
    for (int i = 0; i < p; i++) {
      tmp_p[i] = S[(size_t)p * K + i];                                   // X'C^-1 y   (:228)
      for (int j = 0; j <= i; j++) V[(size_t)j * p + i] = S[(size_t)j * K + i];  // X'C^-1 X (:230)
    }
    if (hl::chol(V.data(), p)) return ctx_fail(c, NNGP_EINVAL, "c++ error: dpotrf failed (X'C^-1X)");
    hl::chol_inverse(V.data(), p);
    for (int i = 0; i < p; i++)
      for (int j = 0; j < i; j++) V[(size_t)i * p + j] = V[(size_t)j * p + i];  // full symmetric
    double *bh = betaHat ? betaHat + (size_t)k * p : bh_local.data();
    hl::symv(V.data(), p, tmp_p.data(), bh);                               // :239
    const double aIG = sigmaSqIG[0] + 1.0 * n / 2.0;                       // :242
    const double bIG = sigmaSqIG[1] + 0.5 * (S[(size_t)p * K + p] - hl::dot(bh, tmp_p.data(), p));  // :245
    if (ab) {
      ab[(size_t)k * 2] = aIG;  // stride 2 even for Matern (reference quirk, :242-245)
      ab[(size_t)k * 2 + 1] = bIG;
    }
    if (bBInv) memcpy(bBInv + (size_t)k * p * p, V.data(), (size_t)p * p * sizeof(double));
    if (bb) memcpy(bb + (size_t)k * p, tmp_p.data(), p * sizeof(double));

    if (n0 > 0) {  // :252-300
      NNGP_CUDA(c, cudaMemcpyAsync(d_bh.p, bh, p * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      NNGP_CUDA(c, cudaMemcpyAsync(d_V.p, V.data(), (size_t)p * p * sizeof(double), cudaMemcpyHostToDevice, c->stream));
      for (size_t ck = 0; ck < chunks.size(); ck++) {
        SiteChunk &ch = chunks[ck];
        rc = site_solve(c, cov, theta, ch);
        if (rc) return rc;
        NNGP_CUDA(c, conj_emit_launch(ch.a.as<double>(), ch.U.as<double>(), ch.F.as<double>(), d_bh.as<double>(),
                                      d_V.as<double>(), p, ch.cs, bIG, aIG, d_y0.as<double>(), d_y0v.as<double>(),
                                      c->stream));
        c->launches++;
        const size_t i0 = ck * (size_t)cs_max;
        if (score) {  // R/spConjNNGP.R:330-334: the scores are all that leaves the device
          NNGP_CUDA(c, conj_score_launch(d_obs.as<double>() + i0, d_y0.as<double>(), d_y0v.as<double>(), ch.cs,
                                         d_scratch.p, d_sc.as<double>() + (size_t)2 * k, c->stream));
          c->launches++;
        }
        if (y0Hat)
          NNGP_CUDA(c, cudaMemcpyAsync(y0Hat + (size_t)k * n0 + i0, d_y0.p, (size_t)ch.cs * sizeof(double),
                                       cudaMemcpyDeviceToHost, c->stream));
        if (y0HatVar)
          NNGP_CUDA(c, cudaMemcpyAsync(y0HatVar + (size_t)k * n0 + i0, d_y0v.p, (size_t)ch.cs * sizeof(double),
                                       cudaMemcpyDeviceToHost, c->stream));
        NNGP_CUDA(c, cudaStreamSynchronize(c->stream));
      }
    }
  }
  if (score) NNGP_CUDA(c, cudaMemcpy(scoreSums, d_sc.p, (size_t)2 * g * sizeof(double), cudaMemcpyDeviceToHost));
  return NNGP_OK;
}

int nngp_cnngp(nngp_ctx *c, int cov, int g, const double *thetaAlpha, const double sigmaSqIG[2], int n0,
               const double *X0, const double *coords0, const int *nnIndx0, double *betaHat, double *ab, double *bBInv,
               double *bb, double *y0Hat, double *y0HatVar) {
  return cnngp_impl(c, cov, g, thetaAlpha, sigmaSqIG, n0, X0, coords0, nnIndx0, betaHat, ab, bBInv, bb, y0Hat, y0HatVar,
                    nullptr, nullptr);
}

int nngp_cnngp_score(nngp_ctx *c, int cov, int g, const double *thetaAlpha, const double sigmaSqIG[2], int n0,
                     const double *X0, const double *coords0, const int *nnIndx0, const double *y0, double *crps,
                     double *rmspe) {
  if (!c) return NNGP_EINVAL;
  if (!y0 || !crps || !rmspe || g < 1) return ctx_fail(c, NNGP_EINVAL, "nngp_cnngp_score: bad arguments");
  std::vector<double> sums((size_t)2 * g);
  const int rc = cnngp_impl(c, cov, g, thetaAlpha, sigmaSqIG, n0, X0, coords0, nnIndx0, nullptr, nullptr, nullptr,
                            nullptr, nullptr, nullptr, y0, sums.data());
  if (rc) return rc;
  for (int k = 0; k < g; k++) {
    crps[k] = sums[(size_t)2 * k] / n0;               // mean(...)          R/util.R:250
    rmspe[k] = sqrt(sums[(size_t)2 * k + 1] / n0);    // sqrt(mean(e^2))    R/util.R:254
  }
  return NNGP_OK;
}

// The k-fold loop of spConjNNGP (R/spConjNNGP.R:286-349) as one call: per fold the hold-out rows are removed, the
// ordered neighbour search is rebuilt on the remaining sites (mkNNIndx / mkNNIndxCB), the held-out sites get their
// m nearest fitted sites (nn2), the conjugate model is fitted and scored for every parameter set.  The hold-out ids
// come from the caller (R draws them with sample(), :289), so the R RNG stream is untouched.
int nngp_conj_kfold(nngp_ctx *c, int n, int p, int m, const double *y, const double *X, const double *coords, int cov,
                    int g, const double *thetaAlpha, const double sigmaSqIG[2], int search_type, int k_fold, int n_ho,
                    const int *hoIdx, int set_begin, int set_end, double *scores) {
  if (!c) return NNGP_EINVAL;
  if (n < 2 || p < 1 || m < 1 || g < 1 || k_fold < 1 || n_ho < 1 || n_ho >= n || !y || !X || !coords || !thetaAlpha ||
      !sigmaSqIG || !hoIdx || !scores || set_begin < 0 || set_end > g || set_begin > set_end)
    return ctx_fail(c, NNGP_EINVAL, "nngp_conj_kfold: bad arguments");
  if (cov < 0 || cov > 3) return ctx_fail(c, NNGP_EINVAL, "cov.model is not correctly specified");
  const int nTheta = (cov == NNGP_COV_MATERN) ? 3 : 2;
  const int n_mod = n - n_ho, gs = set_end - set_begin;
  if (m >= n_mod) return ctx_fail(c, NNGP_EINVAL, "nngp_conj_kfold: n.neighbors must be smaller than the fold's fit set");
  for (int k = 0; k < 2 * g; k++) scores[k] = 0.0;
  if (gs == 0) return NNGP_OK;
  const long long nIndx = (long long)(1 + m) * m / 2 + (long long)(n_mod - m - 1) * m;
  std::vector<double> y_mod(n_mod), X_mod((size_t)n_mod * p), co_mod((size_t)2 * n_mod);
  std::vector<double> y_ho(n_ho), X_ho((size_t)n_ho * p), co_ho((size_t)2 * n_ho);
  std::vector<int> nnIndx((size_t)nIndx), nnIndxLU((size_t)2 * n_mod), nn0((size_t)n_ho * m);
  std::vector<double> nnDist((size_t)nIndx), crps(gs), rmspe(gs);
  std::vector<char> held(n);
  for (int f = 0; f < k_fold; f++) {
    const int *ho = hoIdx + (size_t)f * n_ho;
    std::fill(held.begin(), held.end(), 0);
    for (int j = 0; j < n_ho; j++) {
      if (ho[j] < 0 || ho[j] >= n || held[ho[j]] || (j > 0 && ho[j] <= ho[j - 1]))
        return ctx_fail(c, NNGP_EINVAL, "nngp_conj_kfold: hold-out ids must be ascending, distinct, in [0, n)");
      held[ho[j]] = 1;
    }
    int a = 0, b = 0;  // X[-ho.id,], X[ho.id,] (:291-299): both keep the row order
    for (int i = 0; i < n; i++) {
      if (held[i]) {
        y_ho[b] = y[i];
        co_ho[b] = coords[i];
        co_ho[(size_t)n_ho + b] = coords[(size_t)n + i];
        for (int cc = 0; cc < p; cc++) X_ho[(size_t)cc * n_ho + b] = X[(size_t)cc * n + i];
        b++;
      } else {
        y_mod[a] = y[i];
        co_mod[a] = coords[i];
        co_mod[(size_t)n_mod + a] = coords[(size_t)n + i];
        for (int cc = 0; cc < p; cc++) X_mod[(size_t)cc * n_mod + a] = X[(size_t)cc * n + i];
        a++;
      }
    }
    int rc = nngp_mk_nn_indx(c, n_mod, m, co_mod.data(), search_type, nnIndx.data(), nnDist.data(), nnIndxLU.data(),
                             nullptr);
    if (rc) return rc;
    rc = nngp_nn_query(c, n_mod, co_mod.data(), n_ho, co_ho.data(), m, nn0.data());
    if (rc) return rc;
    rc = nngp_set_data(c, n_mod, m, p, co_mod.data(), nnIndx.data(), nnIndxLU.data(), X_mod.data(), y_mod.data(), 0,
                       n_mod);
    if (rc) return rc;
    rc = nngp_cnngp_score(c, cov, gs, thetaAlpha + (size_t)set_begin * nTheta, sigmaSqIG, n_ho, X_ho.data(),
                          co_ho.data(), nn0.data(), y_ho.data(), crps.data(), rmspe.data());
    if (rc) return rc;
    for (int k = 0; k < gs; k++) {  // column order of k.fold.scores: rmspe, crps (:271-274)
      scores[set_begin + k] += rmspe[k];
      scores[(size_t)g + set_begin + k] += crps[k];
    }
  }
  for (int k = 0; k < 2 * g; k++) scores[k] /= k_fold;  // :345
  return NNGP_OK;
}

}  // extern "C"

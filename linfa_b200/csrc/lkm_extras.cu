// Second translation unit of liblkm.so: the callers and data formats either side of the K-Means path
// (SURVEY 8(f) "next" rows): silhouette score (src/metrics_clustering.rs:65-135) and .npy files
// (algorithms/linfa-clustering/examples/kmeans.rs:40-42, ndarray-npy `write_npy`).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "lkm_ctx.hpp"
#include "lkm_metrics.cuh"
#include "lkm_gmm.cuh"
#include "lkm_rand.hpp"

#define LKM_EXPORT extern "C" __attribute__((visibility("default")))

using namespace lkm;

namespace lkm {

template <typename F>
static lkm_status silhouette_impl(lkm_ctx* ctx, const F* x, uint64_t n, uint64_t d, int64_t row_stride, const uint64_t* labels,
                                  F* score_out) {
  if (!x || !labels || !score_out) return fail(ctx, LKM_ERR_INVALID_ARG, "null argument");
  if (n == 0 || d == 0) return fail(ctx, LKM_ERR_SHAPE, "empty dataset");
  if (row_stride < (int64_t)d) return fail(ctx, LKM_ERR_INVALID_ARG, "row stride smaller than the row");
  CK(cudaSetDevice(ctx->device));
  // label_count (src/dataset/impl_dataset.rs): distinct labels -> dense ids in order of first appearance is irrelevant for
  // the score (min over the other clusters); ids here follow the sorted label values
  std::vector<uint64_t> order(n);
  std::iota(order.begin(), order.end(), 0ull);
  std::stable_sort(order.begin(), order.end(), [&](uint64_t p, uint64_t q) { return labels[p] < labels[q]; });
  std::vector<unsigned long long> off;
  std::vector<uint32_t> cluster_of(n);
  for (uint64_t i = 0; i < n; ++i) {
    if (i == 0 || labels[order[i]] != labels[order[i - 1]]) off.push_back(i);
    cluster_of[i] = (uint32_t)(off.size() - 1);
  }
  const int n_clusters = (int)off.size();
  off.push_back(n);
  // "Single label dataset, all points are in the same cluster" (:77-80)
  if (n_clusters == 1) { *score_out = (F)1; return LKM_OK; }
  std::vector<F> xs((size_t)n * d);
  for (uint64_t i = 0; i < n; ++i) std::memcpy(&xs[i * d], x + order[i] * (uint64_t)row_stride, d * sizeof(F));
  const size_t xi_bytes = (size_t)kSilThreads * (d | 1) * sizeof(F);
  const size_t limit = ctx->smem_optin - 1024;
  if (xi_bytes + d * sizeof(F) > limit) return fail(ctx, LKM_ERR_UNSUPPORTED, "row too wide for the silhouette tile");
  const int tj = (int)std::min<size_t>(256, (limit - xi_bytes) / (d * sizeof(F)));
  const size_t smem = xi_bytes + (size_t)tj * d * sizeof(F);
  DevBuf dx, doff, dcl, dsc;
  lkm_status st = LKM_OK;
  auto run = [&]() -> lkm_status {
    CK(dx.ensure(xs.size() * sizeof(F)));
    CK(doff.ensure(off.size() * 8));
    CK(dcl.ensure(n * 4));
    CK(dsc.ensure(n * sizeof(F)));
    CK(cudaMemcpyAsync(dx.p, xs.data(), xs.size() * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(doff.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dcl.p, cluster_of.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream));
    SilArgs<F> a{};
    a.X = dx.as<F>(); a.n = n; a.d = (int)d; a.n_clusters = n_clusters; a.off = doff.as<unsigned long long>();
    a.cluster_of = dcl.as<uint32_t>(); a.tj = tj; a.scores = dsc.as<F>();
    CK(cudaFuncSetAttribute(silhouette_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    silhouette_kernel<F><<<(unsigned)((n + kSilThreads - 1) / kSilThreads), kSilThreads, smem, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->launches++;
    std::vector<F> sc(n), unsorted(n);
    CK(cudaMemcpyAsync(sc.data(), dsc.p, n * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // `.sum::<F>()` over sample_iter (:128): sample order, in F
    for (uint64_t i = 0; i < n; ++i) unsorted[order[i]] = sc[i];
    F total = (F)0;
    for (uint64_t i = 0; i < n; ++i) total = total + unsorted[i];
    *score_out = total / (F)n;
    return LKM_OK;
  };
  st = run();
  dx.release(); doff.release(); dcl.release(); dsc.release();
  return st;
}

// ---- Gaussian mixture (gaussian_mixture/algorithm.rs) ---------------------------------------------
template <typename F> struct GmmModel {
  uint64_t k = 0, d = 0;
  std::vector<F> weights, means, covariances, precisions, precisions_chol;
};

struct GmmWork {
  DevBuf means, pchol, logdet, logw, log_resp, part, comb, lpn_part, resp0;
  void release() { for (DevBuf* b : {&means, &pchol, &logdet, &logw, &log_resp, &part, &comb, &lpn_part, &resp0}) b->release(); }
};

// compute_precisions_cholesky_full (:261-276): L = cholesky(cov) (lower), P_chol = (L^-1)^T; false = not positive definite
template <typename F>
static bool gmm_precisions_chol(const F* cov, uint64_t d, F* pchol) {
  std::vector<F> L(d * d, (F)0), inv(d * d, (F)0);
  for (uint64_t j = 0; j < d; ++j) {
    F s = cov[j * d + j];
    for (uint64_t p = 0; p < j; ++p) s -= L[j * d + p] * L[j * d + p];
    if (!(s > (F)0)) return false;   // LinalgError::NotPositiveDefinite
    const F ljj = std::sqrt(s);
    L[j * d + j] = ljj;
    for (uint64_t i = j + 1; i < d; ++i) {
      F v = cov[i * d + j];
      for (uint64_t p = 0; p < j; ++p) v -= L[i * d + p] * L[j * d + p];
      L[i * d + j] = v / ljj;
    }
  }
  // solve L * inv = I by forward substitution, column by column
  for (uint64_t c = 0; c < d; ++c)
    for (uint64_t i = 0; i < d; ++i) {
      F v = (i == c) ? (F)1 : (F)0;
      for (uint64_t p = 0; p < i; ++p) v -= L[i * d + p] * inv[p * d + c];
      inv[i * d + c] = v / L[i * d + i];
    }
  for (uint64_t i = 0; i < d; ++i)
    for (uint64_t j = 0; j < d; ++j) pchol[i * d + j] = inv[j * d + i];
  return true;
}

// compute_precisions_full (:278-289): P_chol . P_chol^T
template <typename F> static void gmm_precisions(const F* pchol, uint64_t d, F* prec) {
  for (uint64_t i = 0; i < d; ++i)
    for (uint64_t j = 0; j < d; ++j) {
      F s = (F)0;
      for (uint64_t p = 0; p < d; ++p) s += pchol[i * d + p] * pchol[j * d + p];
      prec[i * d + j] = s;
    }
}

template <typename F>
static lkm_status gmm_upload_model(lkm_ctx* ctx, GmmWork& w, const GmmModel<F>& m) {
  const uint64_t k = m.k, d = m.d;
  std::vector<F> logdet(k), logw(k);
  for (uint64_t c = 0; c < k; ++c) {
    F s = (F)0;   // compute_log_det_cholesky_full (:392-405)
    for (uint64_t i = 0; i < d; ++i) s += std::log(m.precisions_chol[c * d * d + i * d + i]);
    logdet[c] = s;
    logw[c] = std::log(m.weights[c]);   // estimate_log_weights (:407-409)
  }
  CK(w.means.ensure(k * d * sizeof(F)));
  CK(w.pchol.ensure(k * d * d * sizeof(F)));
  CK(w.logdet.ensure(k * sizeof(F)));
  CK(w.logw.ensure(k * sizeof(F)));
  CK(cudaMemcpyAsync(w.means.p, m.means.data(), k * d * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(w.pchol.p, m.precisions_chol.data(), k * d * d * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(w.logdet.p, logdet.data(), k * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(w.logw.p, logw.data(), k * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));   // the host vectors go out of scope
  return LKM_OK;
}

// e_step (:291-298) on a device matrix: mean of log_prob_norm; log_resp stays in w.log_resp
template <typename F>
static lkm_status gmm_e_step(lkm_ctx* ctx, GmmWork& w, const F* dX, uint64_t n, uint64_t d, uint64_t k, F* lower_bound,
                             uint32_t* d_labels, F* d_proba) {
  CK(w.log_resp.ensure(std::max<uint64_t>(n * k, 1) * sizeof(F)));
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + kGmmThreads - 1) / kGmmThreads, (uint64_t)ctx->sm_count * 4));
  CK(w.lpn_part.ensure((size_t)grid * 8));
  GmmEArgs<F> a{};
  a.X = dX; a.n = n; a.d = (int)d; a.k = (int)k; a.means = w.means.as<F>(); a.pchol = w.pchol.as<F>(); a.log_det = w.logdet.as<F>();
  a.log_w = w.logw.as<F>(); a.log_resp = w.log_resp.as<F>(); a.part_lpn = w.lpn_part.as<double>(); a.labels = d_labels; a.proba = d_proba;
  const size_t tile = (size_t)kGmmThreads * (d | 1) * sizeof(F), par = (k * d + k * d * d) * sizeof(F);
  if (tile + 1024 > ctx->smem_optin) return fail(ctx, LKM_ERR_UNSUPPORTED, "row too wide for the E-step tile");
  a.params_in_smem = tile + par + 2048 <= ctx->smem_optin ? 1 : 0;
  const size_t smem = tile + (a.params_in_smem ? par : 0);
  CK(cudaFuncSetAttribute(gmm_e_step_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gmm_e_step_kernel<F><<<grid, kGmmThreads, smem, ctx->stream>>>(a);
  CK(cudaGetLastError());
  ctx->launches++;
  std::vector<double> part(grid);
  CK(cudaMemcpyAsync(part.data(), w.lpn_part.p, (size_t)grid * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  double tot = 0.0;
  for (int b = 0; b < grid; ++b) tot += part[b];
  if (lower_bound) *lower_bound = (F)(tot / (double)n);   // log_prob_norm.mean()
  return LKM_OK;
}

// estimate_gaussian_parameters (:217-259) from the responsibilities in `resp` (log_resp when is_log), then
// compute_precisions_cholesky_full; statuses: 45 EmptyCluster, 46 NotPositiveDefinite
template <typename F>
static lkm_status gmm_m_step(lkm_ctx* ctx, GmmWork& w, const F* dX, uint64_t n, uint64_t d, uint64_t k, const F* d_resp, int is_log,
                             F reg_covar, GmmModel<F>& m) {
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count));
  const uint64_t rows_per = (n + grid - 1) / grid;
  const size_t smem = (size_t)kGmmTile * (d + k) * sizeof(F);
  if (smem > ctx->smem_optin) return fail(ctx, LKM_ERR_UNSUPPORTED, "too many components / features for the M-step tile");
  CK(cudaFuncSetAttribute(gmm_moments_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  const int E0 = (int)(k + k * d), E1 = (int)(k * d * d);
  CK(w.part.ensure((size_t)grid * std::max(E0, E1) * 8));
  CK(w.comb.ensure((size_t)std::max(E0, E1) * 8));
  GmmMArgs<F> a{};
  a.X = dX; a.n = n; a.d = (int)d; a.k = (int)k; a.log_resp = d_resp; a.resp_is_log = is_log; a.part = w.part.as<double>();
  a.rows_per_cta = rows_per; a.mode = 0;
  gmm_moments_kernel<F><<<grid, 256, smem, ctx->stream>>>(a);
  gmm_combine_kernel<<<(E0 + 255) / 256, 256, 0, ctx->stream>>>(w.part.as<double>(), grid, E0, w.comb.as<double>());
  CK(cudaGetLastError());
  ctx->launches += 2;
  std::vector<double> h0(E0);
  CK(cudaMemcpyAsync(h0.data(), w.comb.p, (size_t)E0 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  // nk.min() < 10 eps -> EmptyCluster (:224-229)
  std::vector<F> nk(k);
  for (uint64_t c = 0; c < k; ++c) nk[c] = (F)h0[c];
  for (uint64_t c = 0; c < k; ++c)
    if (!(nk[c] >= (F)10 * std::numeric_limits<F>::epsilon())) {
      ctx->err = "Cluster #" + std::to_string(c + 1) + " has no more point. Consider decreasing number of clusters or change initialization.";
      return (lkm_status)45;
    }
  for (uint64_t c = 0; c < k; ++c)
    for (uint64_t j = 0; j < d; ++j) m.means[c * d + j] = (F)h0[k + c * d + j] / nk[c];
  CK(w.means.ensure(k * d * sizeof(F)));
  CK(cudaMemcpyAsync(w.means.p, m.means.data(), k * d * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  a.mode = 1; a.means = w.means.as<F>();
  gmm_moments_kernel<F><<<grid, 256, smem, ctx->stream>>>(a);
  gmm_combine_kernel<<<(E1 + 255) / 256, 256, 0, ctx->stream>>>(w.part.as<double>(), grid, E1, w.comb.as<double>());
  CK(cudaGetLastError());
  ctx->launches += 2;
  std::vector<double> h1(E1);
  CK(cudaMemcpyAsync(h1.data(), w.comb.p, (size_t)E1 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  for (uint64_t c = 0; c < k; ++c) {
    for (uint64_t e = 0; e < d * d; ++e) m.covariances[c * d * d + e] = (F)h1[c * d * d + e] / nk[c];
    for (uint64_t i = 0; i < d; ++i) m.covariances[c * d * d + i * d + i] += reg_covar;   // cov_k.diag_mut() + reg_covar (:254)
    m.weights[c] = nk[c] / (F)n;
    if (!gmm_precisions_chol<F>(&m.covariances[c * d * d], d, &m.precisions_chol[c * d * d])) {
      ctx->err = "Matrix is not positive definite";
      return (lkm_status)46;
    }
  }
  return LKM_OK;
}

template <typename F>
static lkm_status gmm_fit_impl(lkm_ctx* ctx, const lkm_gmm_params* p, lkm_rng rng_in, F* weights_out, F* means_out, F* cov_out,
                               F* prec_out, F* pchol_out, uint64_t* n_iter_out, F* lower_bound_out) {
  if (!p || !weights_out || !means_out || !cov_out || !prec_out || !pchol_out) return fail(ctx, LKM_ERR_INVALID_ARG, "null argument");
  // ParamGuard::check_ref (gaussian_mixture/hyperparams.rs:167-193), same order
  if (p->n_clusters == 0) return fail(ctx, (lkm_status)40, "`n_clusters` cannot be 0!");
  if (p->tolerance <= 0.0) return fail(ctx, (lkm_status)41, "`tolerance` must be greater than 0!");
  if (p->reg_covariance < 0.0) return fail(ctx, (lkm_status)42, "`reg_covar` must be positive!");
  if (p->n_runs == 0) return fail(ctx, (lkm_status)43, "`n_runs` cannot be 0!");
  if (p->max_n_iterations == 0) return fail(ctx, (lkm_status)44, "`max_n_iterations` cannot be 0!");
  if (ctx->elem != (int)sizeof(F) || !ctx->X) return fail(ctx, LKM_ERR_NO_DATA, "no resident data of this type");
  if (ctx->world > 1) return fail(ctx, LKM_ERR_UNSUPPORTED, "the Gaussian mixture runs on one GPU");
  CK(cudaSetDevice(ctx->device));
  const uint64_t n = ctx->n, d = ctx->d, k = p->n_clusters;
  const F* dX = reinterpret_cast<const F*>(ctx->X);
  const F tol = (F)p->tolerance, reg = (F)p->reg_covariance;
  GmmWork w;
  GmmModel<F> m;
  m.k = k; m.d = d;
  m.weights.assign(k, 0); m.means.assign(k * d, 0); m.covariances.assign(k * d * d, 0); m.precisions.assign(k * d * d, 0);
  m.precisions_chol.assign(k * d * d, 0);
  GmmModel<F> best;
  auto body = [&]() -> lkm_status {
    // ---- GaussianMixtureModel::new (:124-178): initial responsibilities ----
    CK(w.resp0.ensure(std::max<uint64_t>(n * k, 1) * sizeof(F)));
    std::vector<F> resp(n * k, (F)0);
    if (p->init_method == 0) {
      // KMeans::params_with_rng(k, rng).check().unwrap().fit(dataset)?, then model.predict (:137-146)
      lkm_params kp{};
      kp.n_clusters = k; kp.n_runs = 10; kp.max_n_iterations = 300; kp.tolerance = 1e-4;
      kp.init_method = LKM_INIT_KMEANS_PLUSPLUS; kp.algorithm = LKM_ALGO_LLOYD; kp.pick_mode = p->kmeans_pick_mode;
      std::vector<F> cen(k * d), cnt(k);
      F inertia = 0;
      lkm_status st;
      std::vector<uint64_t> labels(n);
      if constexpr (sizeof(F) == 4) {
        st = lkm_fit_f32(ctx, &kp, rng_in, cen.data(), cnt.data(), &inertia, nullptr);
        if (st == LKM_OK) st = lkm_predict_f32(ctx, cen.data(), k, nullptr, n, d, (int64_t)d, labels.data());
      } else {
        st = lkm_fit_f64(ctx, &kp, rng_in, cen.data(), cnt.data(), &inertia, nullptr);
        if (st == LKM_OK) st = lkm_predict_f64(ctx, cen.data(), k, nullptr, n, d, (int64_t)d, labels.data());
      }
      if (st != LKM_OK) return st;
      for (uint64_t i = 0; i < n; ++i) resp[i * k + labels[i]] = (F)1;
    } else {
      // Array2::<f64>::random_using((n, k), Uniform::new(0., 1.), &mut rng), rows normalised by their sums (:148-157)
      HostRng rng{rng_in};
      std::vector<double> r(n * k);
      for (uint64_t e = 0; e < n * k; ++e) r[e] = unit_f64(rng);
      for (uint64_t i = 0; i < n; ++i) {
        double tot = 0.0;
        for (uint64_t c = 0; c < k; ++c) tot += r[i * k + c];
        for (uint64_t c = 0; c < k; ++c) resp[i * k + c] = (F)(r[i * k + c] / tot);
      }
    }
    CK(cudaMemcpyAsync(w.resp0.p, resp.data(), n * k * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
    CKS(gmm_m_step<F>(ctx, w, dX, n, d, k, w.resp0.as<F>(), 0, reg, m));
    for (uint64_t c = 0; c < k; ++c) gmm_precisions<F>(&m.precisions_chol[c * d * d], d, &m.precisions[c * d * d]);
    // ---- Fit::fit (:411-467): every run continues from the same model, as in the reference ----
    F max_lower_bound = -std::numeric_limits<F>::infinity();
    bool have_best = false, best_converged = false;
    uint64_t best_iter = 0;
    for (uint64_t run = 0; run < p->n_runs; ++run) {
      F lower_bound = -std::numeric_limits<F>::infinity();
      bool converged = false;
      uint64_t conv_iter = 0;
      for (uint64_t it = 0; it < p->max_n_iterations; ++it) {
        const F prev = lower_bound;
        CKS(gmm_upload_model<F>(ctx, w, m));
        F lb = 0;
        CKS(gmm_e_step<F>(ctx, w, dX, n, d, k, &lb, nullptr, nullptr));
        CKS(gmm_m_step<F>(ctx, w, dX, n, d, k, w.log_resp.as<F>(), 1, reg, m));
        lower_bound = lb;
        const F change = lower_bound - prev;
        if (std::fabs(change) < tol) { converged = true; conv_iter = it; break; }
      }
      if (lower_bound > max_lower_bound) {
        max_lower_bound = lower_bound;
        for (uint64_t c = 0; c < k; ++c) gmm_precisions<F>(&m.precisions_chol[c * d * d], d, &m.precisions[c * d * d]);   // refresh_precisions_full
        best = m;
        have_best = true;
        best_converged = converged;
        best_iter = conv_iter;
      }
    }
    if (!best_converged) return fail(ctx, (lkm_status)47, "EM fitting algorithm did not converge. Try different init parameters, or increase max_n_iterations, tolerance or check for degenerate data.");
    if (!have_best) return fail(ctx, (lkm_status)48, "No lower bound improvement (-inf)");
    if (n_iter_out) *n_iter_out = best_iter;
    if (lower_bound_out) *lower_bound_out = max_lower_bound;
    return LKM_OK;
  };
  const lkm_status st = body();
  w.release();
  if (st != LKM_OK) return st;
  std::memcpy(weights_out, best.weights.data(), k * sizeof(F));
  std::memcpy(means_out, best.means.data(), k * d * sizeof(F));
  std::memcpy(cov_out, best.covariances.data(), k * d * d * sizeof(F));
  std::memcpy(prec_out, best.precisions.data(), k * d * d * sizeof(F));
  std::memcpy(pchol_out, best.precisions_chol.data(), k * d * d * sizeof(F));
  return LKM_OK;
}

// estimate_log_prob_resp on host observations with a given model: predict (:470-484), predict_proba (:205-208), e_step
template <typename F>
static lkm_status gmm_predict_impl(lkm_ctx* ctx, uint64_t k, const F* weights, const F* means, const F* pchol, const F* x, uint64_t n,
                                   uint64_t d, int64_t row_stride, uint64_t* labels_out, F* proba_out, F* log_prob_norm_mean_out) {
  if (!weights || !means || !pchol || k == 0) return fail(ctx, LKM_ERR_INVALID_ARG, "null argument");
  if (n == 0 || d == 0) return fail(ctx, LKM_ERR_SHAPE, "empty dataset");
  CK(cudaSetDevice(ctx->device));
  GmmWork w;
  DevBuf dx, dlab, dproba;
  auto body = [&]() -> lkm_status {
    const F* dX = nullptr;
    if (x == nullptr) {
      if (ctx->elem != (int)sizeof(F) || !ctx->X || ctx->n != n || ctx->d != d) return fail(ctx, LKM_ERR_NO_DATA, "no resident data of this shape");
      dX = reinterpret_cast<const F*>(ctx->X);
    } else {
      if (row_stride < (int64_t)d) return fail(ctx, LKM_ERR_INVALID_ARG, "row stride smaller than the row");
      CK(dx.ensure(n * d * sizeof(F)));
      CK(cudaMemcpy2DAsync(dx.p, d * sizeof(F), x, (size_t)row_stride * sizeof(F), d * sizeof(F), n, cudaMemcpyHostToDevice, ctx->stream));
      dX = dx.as<F>();
    }
    GmmModel<F> m;
    m.k = k; m.d = d;
    m.weights.assign(weights, weights + k); m.means.assign(means, means + k * d); m.precisions_chol.assign(pchol, pchol + k * d * d);
    CKS(gmm_upload_model<F>(ctx, w, m));
    if (labels_out) CK(dlab.ensure(n * 4));
    if (proba_out) CK(dproba.ensure(n * k * sizeof(F)));
    F lb = 0;
    CKS(gmm_e_step<F>(ctx, w, dX, n, d, k, &lb, labels_out ? dlab.as<uint32_t>() : nullptr, proba_out ? dproba.as<F>() : nullptr));
    if (log_prob_norm_mean_out) *log_prob_norm_mean_out = lb;
    if (labels_out) {
      std::vector<uint32_t> l32(n);
      CK(cudaMemcpyAsync(l32.data(), dlab.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      for (uint64_t i = 0; i < n; ++i) labels_out[i] = l32[i];
    }
    if (proba_out) {
      CK(cudaMemcpyAsync(proba_out, dproba.p, n * k * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
    }
    return LKM_OK;
  };
  const lkm_status st = body();
  w.release(); dx.release(); dlab.release(); dproba.release();
  return st;
}

// ---- .npy (format 1.0, little endian, C order) ------------------------------------------------
static const char* npy_descr(int dtype) { return dtype == 0 ? "<f4" : dtype == 1 ? "<f8" : dtype == 2 ? "<u8" : nullptr; }
static size_t npy_itemsize(int dtype) { return dtype == 0 ? 4 : 8; }

}  // namespace lkm

LKM_EXPORT lkm_status lkm_silhouette_score_f32(lkm_ctx* ctx, const float* x, uint64_t n, uint64_t d, int64_t row_stride,
                                               const uint64_t* labels, float* score_out) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  return silhouette_impl<float>(ctx, x, n, d, row_stride, labels, score_out);
}
LKM_EXPORT lkm_status lkm_silhouette_score_f64(lkm_ctx* ctx, const double* x, uint64_t n, uint64_t d, int64_t row_stride,
                                               const uint64_t* labels, double* score_out) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  return silhouette_impl<double>(ctx, x, n, d, row_stride, labels, score_out);
}


LKM_EXPORT lkm_status lkm_gmm_fit_f32(lkm_ctx* ctx, const lkm_gmm_params* p, lkm_rng rng, float* weights_out, float* means_out,
                                      float* covariances_out, float* precisions_out, float* precisions_chol_out, uint64_t* n_iter_out,
                                      float* lower_bound_out) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  return gmm_fit_impl<float>(ctx, p, rng, weights_out, means_out, covariances_out, precisions_out, precisions_chol_out, n_iter_out, lower_bound_out);
}
LKM_EXPORT lkm_status lkm_gmm_fit_f64(lkm_ctx* ctx, const lkm_gmm_params* p, lkm_rng rng, double* weights_out, double* means_out,
                                      double* covariances_out, double* precisions_out, double* precisions_chol_out, uint64_t* n_iter_out,
                                      double* lower_bound_out) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  return gmm_fit_impl<double>(ctx, p, rng, weights_out, means_out, covariances_out, precisions_out, precisions_chol_out, n_iter_out, lower_bound_out);
}
LKM_EXPORT lkm_status lkm_gmm_predict_f32(lkm_ctx* ctx, uint64_t k, const float* weights, const float* means, const float* precisions_chol,
                                          const float* x, uint64_t n, uint64_t d, int64_t row_stride, uint64_t* labels_out,
                                          float* proba_out, float* log_prob_norm_mean_out) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  return gmm_predict_impl<float>(ctx, k, weights, means, precisions_chol, x, n, d, row_stride, labels_out, proba_out, log_prob_norm_mean_out);
}
LKM_EXPORT lkm_status lkm_gmm_predict_f64(lkm_ctx* ctx, uint64_t k, const double* weights, const double* means, const double* precisions_chol,
                                          const double* x, uint64_t n, uint64_t d, int64_t row_stride, uint64_t* labels_out,
                                          double* proba_out, double* log_prob_norm_mean_out) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  return gmm_predict_impl<double>(ctx, k, weights, means, precisions_chol, x, n, d, row_stride, labels_out, proba_out, log_prob_norm_mean_out);
}

LKM_EXPORT lkm_status lkm_write_npy(const char* path, int dtype, const void* data, int ndim, const uint64_t* shape) {
  if (!path || !data || ndim < 0 || ndim > 8 || (ndim > 0 && !shape) || !npy_descr(dtype)) return LKM_ERR_INVALID_ARG;
  std::string dict = std::string("{'descr': '") + npy_descr(dtype) + "', 'fortran_order': False, 'shape': (";
  size_t count = 1;
  for (int i = 0; i < ndim; ++i) {
    dict += std::to_string((unsigned long long)shape[i]);
    if (ndim == 1 || i + 1 < ndim) dict += ",";
    if (i + 1 < ndim) dict += " ";
    count *= (size_t)shape[i];
  }
  dict += "), }";
  // magic (6) + version (2) + header length (2) + dict + padding + '\n' is a multiple of 64
  size_t total = 10 + dict.size() + 1;
  const size_t pad = (64 - total % 64) % 64;
  dict.append(pad, ' ');
  dict += "\n";
  if (dict.size() > 65535) return LKM_ERR_INVALID_ARG;
  FILE* f = std::fopen(path, "wb");
  if (!f) return LKM_ERR_INVALID_ARG;
  const unsigned char head[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(dict.size() & 255), (unsigned char)(dict.size() >> 8)};
  bool ok = std::fwrite(head, 1, 10, f) == 10 && std::fwrite(dict.data(), 1, dict.size(), f) == dict.size();
  const size_t bytes = count * npy_itemsize(dtype);
  ok = ok && (bytes == 0 || std::fwrite(data, 1, bytes, f) == bytes);
  ok = (std::fclose(f) == 0) && ok;
  return ok ? LKM_OK : LKM_ERR_INVALID_ARG;
}

// reads the header: dtype code, ndim, shape (up to 8 dims); data_offset = first byte of the array
LKM_EXPORT lkm_status lkm_read_npy_header(const char* path, int* dtype_out, int* ndim_out, uint64_t* shape_out, uint64_t* data_offset_out) {
  if (!path || !dtype_out || !ndim_out || !shape_out || !data_offset_out) return LKM_ERR_INVALID_ARG;
  FILE* f = std::fopen(path, "rb");
  if (!f) return LKM_ERR_INVALID_ARG;
  unsigned char head[12];
  lkm_status st = LKM_ERR_INVALID_ARG;
  do {
    if (std::fread(head, 1, 10, f) != 10 || std::memcmp(head, "\x93NUMPY", 6) != 0) break;
    size_t hlen = head[8] | ((size_t)head[9] << 8), hoff = 10;
    if (head[6] >= 2) {   // versions 2.0 / 3.0: four-byte header length
      if (std::fread(head + 10, 1, 2, f) != 2) break;
      hlen |= ((size_t)head[10] << 16) | ((size_t)head[11] << 24);
      hoff = 12;
    }
    std::string dict(hlen, '\0');
    if (std::fread(&dict[0], 1, hlen, f) != hlen) break;
    const size_t dp = dict.find("'descr'");
    if (dp == std::string::npos) break;
    const size_t q0 = dict.find('\'', dp + 7), q1 = q0 == std::string::npos ? q0 : dict.find('\'', q0 + 1);
    if (q1 == std::string::npos) break;
    const std::string descr = dict.substr(q0 + 1, q1 - q0 - 1);
    int dt = -1;
    if (descr == "<f4") dt = 0; else if (descr == "<f8") dt = 1; else if (descr == "<u8") dt = 2;
    if (dt < 0) { st = LKM_ERR_UNSUPPORTED; break; }
    if (dict.find("'fortran_order': False") == std::string::npos) { st = LKM_ERR_UNSUPPORTED; break; }
    const size_t sp = dict.find("'shape'");
    const size_t p0 = sp == std::string::npos ? sp : dict.find('(', sp), p1 = p0 == std::string::npos ? p0 : dict.find(')', p0);
    if (p1 == std::string::npos) break;
    int nd = 0;
    bool bad = false;
    size_t i = p0 + 1;
    while (i < p1) {
      while (i < p1 && (dict[i] == ' ' || dict[i] == ',')) ++i;
      if (i >= p1) break;
      if (dict[i] < '0' || dict[i] > '9' || nd == 8) { bad = true; break; }
      unsigned long long v = 0;
      while (i < p1 && dict[i] >= '0' && dict[i] <= '9') v = v * 10 + (unsigned)(dict[i++] - '0');
      shape_out[nd++] = v;
    }
    if (bad) break;
    *dtype_out = dt; *ndim_out = nd; *data_offset_out = hoff + hlen;
    st = LKM_OK;
  } while (0);
  std::fclose(f);
  return st;
}

LKM_EXPORT lkm_status lkm_read_npy_data(const char* path, uint64_t data_offset, void* out, uint64_t bytes) {
  if (!path || (!out && bytes)) return LKM_ERR_INVALID_ARG;
  FILE* f = std::fopen(path, "rb");
  if (!f) return LKM_ERR_INVALID_ARG;
  bool ok = std::fseek(f, (long)data_offset, SEEK_SET) == 0 && (bytes == 0 || std::fread(out, 1, bytes, f) == bytes);
  std::fclose(f);
  return ok ? LKM_OK : LKM_ERR_INVALID_ARG;
}

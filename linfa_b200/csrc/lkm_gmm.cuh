// Gaussian mixture model, full covariances (algorithms/linfa-clustering/src/gaussian_mixture/algorithm.rs): the first
// in-repo caller of the K-Means path (GmmInitMethod::KMeans, :137-146) and the E / M steps around it.
//
//   gmm_e_step_kernel   estimate_log_prob_resp (:339-352): thread = sample; per component y = (x - mu) . P_chol,
//                       log N = -0.5 (|y|^2 + d ln 2 pi) + log det P_chol + ln w, then the log-sum-exp over the components
//   gmm_moments_kernel  estimate_gaussian_parameters (:217-259): per-CTA partial sums of resp, resp * x and
//                       resp * (x - mu)(x - mu)^T over a contiguous row range; every output element has one owning thread that
//                       walks the rows in order (f64 accumulators), gmm_combine_kernel adds the CTAs in CTA order: reproducible
//
// The k small Cholesky factorizations / triangular inverses of compute_precisions_cholesky_full (:261-276) run on the host.
#pragma once
#include "lkm_kernels.cuh"

namespace lkm {

constexpr int kGmmThreads = 128;

template <typename F> struct GmmEArgs {
  const F* X;
  unsigned long long n;
  int d, k;
  const F* means;       // k x d
  const F* pchol;       // k x d x d (row-major per component)
  const F* log_det;     // k: sum of ln diag(P_chol)
  const F* log_w;       // k: ln weights
  F* log_resp;          // n x k (workspace and output)
  F* log_prob_norm;     // n (may be null)
  double* part_lpn;     // [grid] partial sums of log_prob_norm
  uint32_t* labels;     // optional: argmax of the responsibilities (predict_inplace :470-484)
  F* proba;             // optional: exp(log_resp), n x k (predict_proba :205-208)
  int params_in_smem;
};

template <typename F>
__global__ void __launch_bounds__(kGmmThreads) gmm_e_step_kernel(const GmmEArgs<F> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ double red[kGmmThreads];
  const int d = a.d, k = a.k, t = threadIdx.x, xs = d | 1;
  F* xt = reinterpret_cast<F*>(smem_raw);                         // [128][xs]
  F* sm_means = xt + (size_t)kGmmThreads * xs;                     // k x d
  F* sm_pchol = sm_means + (size_t)k * d;                          // k x d x d
  const F* means = a.means;
  const F* pchol = a.pchol;
  if (a.params_in_smem) {
    for (int e = t; e < k * d; e += kGmmThreads) sm_means[e] = a.means[e];
    for (int e = t; e < k * d * d; e += kGmmThreads) sm_pchol[e] = a.pchol[e];
    means = sm_means;
    pchol = sm_pchol;
  }
  const F half_d_ln2pi = (F)((double)d * 1.8378770664093453);   // n_features * ln(2 pi)
  double lpn_acc = 0.0;
  const unsigned long long n_tiles = (a.n + kGmmThreads - 1) / kGmmThreads;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long row0 = tile * kGmmThreads;
    const int rows = (int)min((unsigned long long)kGmmThreads, a.n - row0);
    __syncthreads();
    for (int e = t; e < rows * d; e += kGmmThreads) xt[(size_t)(e / d) * xs + (e % d)] = a.X[row0 * d + e];
    __syncthreads();
    if (t >= rows) continue;
    const unsigned long long row = row0 + t;
    const F* x = xt + (size_t)t * xs;
    F* lr = a.log_resp + row * (unsigned long long)k;
    F mx = -(F)INFINITY;
    for (int c = 0; c < k; ++c) {
      const F* mu = means + (size_t)c * d;
      const F* P = pchol + (size_t)c * d * d;
      F q = (F)0;
      for (int j = 0; j < d; ++j) {
        F y = (F)0;
        for (int i = 0; i < d; ++i) y += (x[i] - mu[i]) * P[(size_t)i * d + j];
        q += y * y;
      }
      const F lp = (F)-0.5 * (q + half_d_ln2pi) + a.log_det[c] + a.log_w[c];
      lr[c] = lp;
      mx = fmax(mx, lp);
    }
    F se = (F)0;
    for (int c = 0; c < k; ++c) se += exp(lr[c] - mx);
    const F lpn = log(se) + mx;
    if (a.log_prob_norm != nullptr) a.log_prob_norm[row] = lpn;
    lpn_acc += (double)lpn;
    int am = 0;
    F best = -(F)INFINITY;
    for (int c = 0; c < k; ++c) {
      const F v = lr[c] - lpn;
      lr[c] = v;
      const F p = exp(v);
      if (a.proba != nullptr) a.proba[row * (unsigned long long)k + c] = p;
      if (p > best) { best = p; am = c; }   // argmax: first maximum
    }
    if (a.labels != nullptr) a.labels[row] = (uint32_t)am;
  }
  __syncthreads();
  red[t] = lpn_acc;
  __syncthreads();
  for (int o = kGmmThreads / 2; o > 0; o >>= 1) {
    if (t < o) red[t] += red[t + o];
    __syncthreads();
  }
  if (t == 0 && a.part_lpn != nullptr) a.part_lpn[blockIdx.x] = red[0];
}

template <typename F> struct GmmMArgs {
  const F* X;
  unsigned long long n;
  int d, k;
  const F* log_resp;    // n x k (mode 2: resp itself when resp_is_log == 0)
  int resp_is_log;
  const F* means;       // mode 1 (second moments): k x d
  int mode;             // 0: [k] sums of resp and [k x d] sums of resp * x; 1: [k x d x d] sums of resp * (x - mu)(x - mu)^T
  double* part;         // [grid][E]
  unsigned long long rows_per_cta;
};

constexpr int kGmmTile = 32;   // rows staged per step

template <typename F>
__global__ void __launch_bounds__(256) gmm_moments_kernel(const GmmMArgs<F> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = a.d, k = a.k, t = threadIdx.x;
  F* xt = reinterpret_cast<F*>(smem_raw);            // [32][d]
  F* rt = xt + (size_t)kGmmTile * d;                 // [32][k] responsibilities
  const int E = a.mode == 0 ? k + k * d : k * d * d;
  const unsigned long long r_begin = min(a.n, (unsigned long long)blockIdx.x * a.rows_per_cta);
  const unsigned long long r_end = min(a.n, r_begin + a.rows_per_cta);
  // this thread's elements: e = t, t + 256, ... (at most 8 kept in registers per sweep)
  for (int e0 = 0; e0 < E; e0 += 256 * 8) {
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (unsigned long long r0 = r_begin; r0 < r_end; r0 += kGmmTile) {
      const int rows = (int)min((unsigned long long)kGmmTile, r_end - r0);
      __syncthreads();
      for (int e = t; e < rows * d; e += 256) xt[e] = a.X[r0 * d + e];
      for (int e = t; e < rows * k; e += 256) {
        const F v = a.log_resp[r0 * k + e];
        rt[e] = a.resp_is_log ? exp(v) : v;
      }
      __syncthreads();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int e = e0 + q * 256 + t;
        if (e >= E) break;
        double s = acc[q];
        if (a.mode == 0) {
          if (e < k) {
            for (int r = 0; r < rows; ++r) s += (double)rt[r * k + e];
          } else {
            const int c = (e - k) / d, j = (e - k) % d;
            for (int r = 0; r < rows; ++r) s += (double)rt[r * k + c] * (double)xt[r * d + j];
          }
        } else {
          const int c = e / (d * d), ij = e % (d * d), i = ij / d, j = ij % d;
          const double mi = (double)a.means[c * d + i], mj = (double)a.means[c * d + j];
          for (int r = 0; r < rows; ++r) s += (double)rt[r * k + c] * ((double)xt[r * d + i] - mi) * ((double)xt[r * d + j] - mj);
        }
        acc[q] = s;
      }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const int e = e0 + q * 256 + t;
      if (e < E) a.part[(size_t)blockIdx.x * E + e] = acc[q];
    }
  }
}

// out[e] = sum over the CTAs, in CTA order
static __global__ void gmm_combine_kernel(const double* __restrict__ part, int n_part, int E, double* __restrict__ out) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= E) return;
  double s = 0.0;
  for (int b = 0; b < n_part; ++b) s += part[(size_t)b * E + e];
  out[e] = s;
}

}  // namespace lkm

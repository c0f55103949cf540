// Silhouette score (src/metrics_clustering.rs:65-135): for every sample the mean Euclidean distance to the samples of its
// own cluster (itself excluded) and the smallest mean distance to another cluster; score = mean of (b - a) / max(a, b).
// O(n^2 d): the quality metric the project introduced to test K-Means.
//
// The host sorts the rows stably by label, so the samples of a cluster are contiguous and stay in sample order.  CTA =
// 128 samples i (thread = sample, its row in a padded shared-memory tile); all CTAs stream the n rows j through shared
// memory in tiles.  Every thread walks the j rows in order with ONE running total: at a cluster boundary (uniform over
// the CTA) the total becomes that cluster's mean and a / b are updated.  The arithmetic follows the reference operation
// by operation — (x_i - x_j)^2 summed like ndarray's 8-lane sum() (DistanceCount::add_point :61-63), sqrt and the
// running total in F, in sample order — so the per-sample scores are bit-identical to the reference's; the host adds
// them in sample order in F (Iterator::sum, :128).
#pragma once
#include "lkm_kernels.cuh"

namespace lkm {

template <typename F> struct SilArgs {
  const F* X;                 // n x d, rows sorted by cluster
  unsigned long long n;
  int d;
  int n_clusters;
  const unsigned long long* off;   // [n_clusters + 1] first row of every cluster
  const uint32_t* cluster_of;      // [n] cluster of every (sorted) row
  int tj;                     // rows per j tile
  F* scores;                  // [n] silhouette of every (sorted) row
};

constexpr int kSilThreads = 128;

// sum of squares of (a - b) in ndarray's unrolled_fold order (numeric_util::unrolled_fold: 8 partial sums over the
// full groups of 8, combined as ((p0+p4)+(p1+p5))+((p2+p6)+(p3+p7)) into the accumulator, then the tail in order)
template <typename F>
__device__ __forceinline__ F sil_sq_dist(const F* __restrict__ a, const F* __restrict__ b, int d) {
  F p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  int j = 0;
  for (; j + 8 <= d; j += 8) {
#pragma unroll
    for (int l = 0; l < 8; ++l) {
      const F df = Rn<F>::sub(a[j + l], b[j + l]);
      p[l] = Rn<F>::add(p[l], Rn<F>::mul(df, df));
    }
  }
  F acc = (F)0;
  acc = Rn<F>::add(acc, Rn<F>::add(p[0], p[4]));
  acc = Rn<F>::add(acc, Rn<F>::add(p[1], p[5]));
  acc = Rn<F>::add(acc, Rn<F>::add(p[2], p[6]));
  acc = Rn<F>::add(acc, Rn<F>::add(p[3], p[7]));
  for (; j < d; ++j) {
    const F df = Rn<F>::sub(a[j], b[j]);
    acc = Rn<F>::add(acc, Rn<F>::mul(df, df));
  }
  return acc;
}

template <typename F>
__global__ void __launch_bounds__(kSilThreads) silhouette_kernel(const SilArgs<F> a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = a.d, t = threadIdx.x;
  const int xs = d | 1;                                   // odd row stride: thread = row reads are conflict-free
  F* xi = reinterpret_cast<F*>(smem_raw);                 // [128][xs]
  F* xj = xi + (size_t)kSilThreads * xs;                  // [tj][d]
  const unsigned long long i0 = (unsigned long long)blockIdx.x * kSilThreads;
  const int rows_i = (int)min((unsigned long long)kSilThreads, a.n - i0);
  for (int e = t; e < rows_i * d; e += kSilThreads) xi[(size_t)(e / d) * xs + (e % d)] = a.X[i0 * d + e];
  const bool valid = t < rows_i;
  const uint32_t my_c = valid ? a.cluster_of[i0 + t] : 0u;
  const F* me = xi + (size_t)(valid ? t : 0) * xs;
  F tot = (F)0, a_x = (F)0, b_x = (F)0;
  bool have_b = false;
  int c = 0;
  unsigned long long c_end = a.off[1];
  while (c_end == 0ull && c + 1 < a.n_clusters) { ++c; c_end = a.off[c + 1]; }   // (dense ids: no empty clusters; kept for safety)
  for (unsigned long long j0 = 0; j0 < a.n; j0 += a.tj) {
    const int rows_j = (int)min((unsigned long long)a.tj, a.n - j0);
    __syncthreads();
    for (int e = t; e < rows_j * d; e += kSilThreads) xj[e] = a.X[j0 * d + e];
    __syncthreads();
    for (int r = 0; r < rows_j; ++r) {
      const unsigned long long j = j0 + r;
      // DistanceCount::add_point: total_distance += sqrt(sum((x_i - x_j)^2))
      tot = Rn<F>::add(tot, sqrt(sil_sq_dist<F>(me, xj + (size_t)r * d, d)));
      if (j + 1 == c_end) {
        // cluster c is complete: same_label_mean_distance / mean_distance, then reset
        const unsigned long long cnt = c_end - a.off[c];
        if ((uint32_t)c == my_c) a_x = (cnt == 1ull) ? (F)0 : tot / (F)(cnt - 1ull);
        else {
          const F m = tot / (F)cnt;
          if (!have_b || m < b_x) b_x = m;
          have_b = true;
        }
        tot = (F)0;
        ++c;
        if (c < a.n_clusters) c_end = a.off[c + 1];
      }
    }
  }
  if (valid) a.scores[i0 + t] = (a_x >= b_x) ? (b_x - a_x) / a_x : (b_x - a_x) / b_x;
}

}  // namespace lkm

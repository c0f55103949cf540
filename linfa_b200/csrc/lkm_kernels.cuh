// CUDA-core kernels of the K-Means hot path (sm_100a).
//
//   lloyd_kernel<F, D, MODE>   assignment (bit-exact direct form) and/or the
//                              deterministic per-CTA accumulation of cluster
//                              sums, counts and inertia      
//   finalize_kernel<...>       fixed-order combine of the partials, m_k-means
//                              update, shift, convergence flag   
//   seq_prefix / count_le      WeightedIndex exact pick       
//   block_sums / fast_pick     fixed-order f64 scan pick      (fast)
//   bernoulli_select           k-means|| oversampling         
//   gather_rows, gen_blobs     row gather; synthetic blobs
//
// Reference arithmetic being reproduced (KM = algorithms/linfa-clustering/src/k_means):
//   KM/algorithm.rs:923-946 closest_centroid: strict '<', lowest index wins
//   linfa-nn/src/distance.rs:68-70 + ndarray-stats sq_l2_dist: acc += (c_j-x_j)^2
//     in F, feature order, no FMA  -> __fsub_rn/__fmul_rn/__fadd_rn (never contracted)
//   KM/algorithm.rs:785-810 compute_centroids: (sum + old) / (count + 1)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lkm {

struct LloydState {
  unsigned long long n_iter;
  unsigned long long fallback_rows;
  double inertia;  // sum of min squared distances of the last assignment
  double shift;    // ||C_t - C_{t+1}||_F of the last iteration
  int done;
  unsigned int blocks_done;
  unsigned long long mixed_tiles;   // tiles the fused update of the pair kernel left to the listed update pass
  unsigned long long label_breaks;  // adjacent rows with different labels (label_breaks_kernel, first iteration of a call)
};

enum LloydMode { MODE_ASSIGN = 0, MODE_FUSED = 1, MODE_UPDATE = 2 };

// Programmatic dependent launch: a kernel launched with the stream-serialization attribute may be
// scheduled while its predecessor drains; pdl_wait() blocks until the predecessor grid has completed
// and its writes are visible, pdl_launch_dependents() lets the successor start its own launch early.
// Both are no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

constexpr int kThreads = 256;  // threads per CTA == rows per tile

template <typename F> struct Rn;
template <> struct Rn<float> {
  static __device__ __forceinline__ float sub(float a, float b) { return __fsub_rn(a, b); }
  static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
  static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
};
template <> struct Rn<double> {
  static __device__ __forceinline__ double sub(double a, double b) { return __dsub_rn(a, b); }
  static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
  static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
};

template <typename F> struct LloydArgs {
  const F* X;            // n x d row-major (this rank's shard)
  unsigned long long n;
  int d;
  int k;
  const F* C;            // k x d
  // assignment outputs (any may be null)
  uint32_t* labels;
  F* dists;              // min squared distance; with min_combine: dists = min(dists, new)
  int min_combine;
  const F* weights;      // optional: wdists = dists * weights (KM/init.rs:147)
  F* wdists;
  const uint32_t* labels_in;  // MODE_UPDATE
  // accumulation outputs: per-CTA partials
  F* part_sums;          // CTA b, cluster c, feature j at part_sums[b*k*d + c*d + j]
  uint32_t* part_counts; // CTA b, cluster c at part_counts[b*k + c]
  double* part_inertia;  // [grid]
  const LloydState* state;  // if non-null and state->done: exit at once
  int n_subsets;         // private accumulator copies per CTA
  int x_stride;          // smem row stride of the tile, in elements
  int kc;                // centroid rows resident in smem per chunk
  int chunk_base;        // clusters [chunk_base, chunk_base+chunk_rows) are accumulated
  int chunk_rows;
  int vec;               // rows are 16-byte vectorisable
  int tpr;               // threads cooperating on one row (power of two <= 32); tile = kThreads/tpr rows
  int certified;         // 1: expanded-form FMA scores + certified argmin (D > 0, all centroids resident)
  LloydState* state_rw;  // fallback-row counter of the certified path (may be null)
};

// ---- squared distance, exact reference order ------------------------------
template <typename F, int D>
__device__ __forceinline__ F sqdist_reg(const F* __restrict__ c, const F (&x)[D]) {
  F acc = (F)0;
#pragma unroll
  for (int j = 0; j < D; ++j) {
    const F diff = Rn<F>::sub(c[j], x[j]);
    acc = Rn<F>::add(acc, Rn<F>::mul(diff, diff));
  }
  return acc;
}
template <typename F>
__device__ __forceinline__ F sqdist_mem(const F* __restrict__ c, const F* __restrict__ x, int d) {
  F acc = (F)0;
  for (int j = 0; j < d; ++j) {
    const F diff = Rn<F>::sub(c[j], x[j]);
    acc = Rn<F>::add(acc, Rn<F>::mul(diff, diff));
  }
  return acc;
}

// scan `rows_c` centroids at cs (row stride d) in ascending order, strict '<'
template <typename F, int D>
__device__ __forceinline__ void scan_centroids(const F* __restrict__ cs, int rows_c, int d, int base,
                                               const F* __restrict__ xrow, bool first, F& best, int& bidx) {
  if constexpr (D > 0) {
    F x[D];
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = xrow[j];
    int c = 0;
    if (first) { best = sqdist_reg<F, D>(cs, x); bidx = base; c = 1; }
    for (; c + 4 <= rows_c; c += 4) {
      const F v0 = sqdist_reg<F, D>(cs + (c + 0) * D, x);
      const F v1 = sqdist_reg<F, D>(cs + (c + 1) * D, x);
      const F v2 = sqdist_reg<F, D>(cs + (c + 2) * D, x);
      const F v3 = sqdist_reg<F, D>(cs + (c + 3) * D, x);
      if (v0 < best) { best = v0; bidx = base + c; }
      if (v1 < best) { best = v1; bidx = base + c + 1; }
      if (v2 < best) { best = v2; bidx = base + c + 2; }
      if (v3 < best) { best = v3; bidx = base + c + 3; }
    }
    for (; c < rows_c; ++c) {
      const F v = sqdist_reg<F, D>(cs + c * D, x);
      if (v < best) { best = v; bidx = base + c; }
    }
  } else {
    int c = 0;
    if (first) { best = sqdist_mem<F>(cs, xrow, d); bidx = base; c = 1; }
    for (; c + 2 <= rows_c; c += 2) {
      const F v0 = sqdist_mem<F>(cs + (size_t)(c + 0) * d, xrow, d);
      const F v1 = sqdist_mem<F>(cs + (size_t)(c + 1) * d, xrow, d);
      if (v0 < best) { best = v0; bidx = base + c; }
      if (v1 < best) { best = v1; bidx = base + c + 1; }
    }
    for (; c < rows_c; ++c) {
      const F v = sqdist_mem<F>(cs + (size_t)c * d, xrow, d);
      if (v < best) { best = v; bidx = base + c; }
    }
  }
}

// strided variant for `tpr` threads sharing one row: thread q looks at centroids
// q, q+tpr, ...  Only the seeding thread (q == 0 on the first chunk) starts from
// d(c_0) like the reference; the others start from "no candidate" (+inf, -1).
template <typename F>
__device__ __forceinline__ void scan_centroids_strided(const F* __restrict__ cs, int rows_c, int d, int base,
                                                       const F* __restrict__ xrow, bool seed, int q, int step,
                                                       F& best, int& bidx) {
  int c = q;
  if (seed) { best = sqdist_mem<F>(cs, xrow, d); bidx = base; c = step; }
  for (; c < rows_c; c += step) {
    const F v = sqdist_mem<F>(cs + (size_t)c * d, xrow, d);
    if (v < best) { best = v; bidx = base + c; }
  }
}

// butterfly combine of the per-thread candidates of one row.  A NaN can only be
// held by the seeding thread (sticky NaN of centroid 0, as in the reference).
template <typename F>
__device__ __forceinline__ void combine_candidates(int tpr, F& best, int& bidx) {
  for (int o = tpr >> 1; o > 0; o >>= 1) {
    const F ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bidx, o);
    const bool a_nan = best != best, b_nan = ov != ov;
    const bool b_wins = b_nan || (!a_nan && oi >= 0 && (bidx < 0 || ov < best || (ov == best && oi < bidx)));
    if (b_wins) { best = ov; bidx = oi; }
  }
}

// running (best, second best, index of best): two compares per element.  A NaN score leaves the
// state untouched; a row whose scores are all NaN ends with best = second = +inf and is not certified.
template <typename F> __device__ __forceinline__ void top2_update(F w, int c, F& best, F& second, int& idx) {
  const bool lt_b = w < best, lt_s = w < second;
  second = lt_b ? best : (lt_s ? w : second);
  best = lt_b ? w : best;
  idx = lt_b ? c : idx;
}

// running three lowest scores with the indices of the first two (equal scores: the earlier index stays in front)
template <typename F> __device__ __forceinline__ void top3_update(F w, int c, F& best, F& second, F& third, int& ib, int& is) {
  const bool lt_b = w < best, lt_s = w < second, lt_t = w < third;
  third = lt_s ? second : (lt_t ? w : third);
  second = lt_b ? best : (lt_s ? w : second);
  is = lt_b ? ib : (lt_s ? c : is);
  best = lt_b ? w : best;
  ib = lt_b ? c : ib;
}

// Certified argmin on CUDA cores (expanded form): w_c = ||c||^2 - 2 x.c with FMA chains, running best
// and second best.  With eps = 2^-24 (f32) / 2^-53 (f64), |w_c - W_c| <= eps[(2d+2)|x|C + 2C^2]
// (d FMA roundings on the dot, one on ||c||^2, one on the final fma; C = max|c|) and the
// reference's direct form is within (d+2) eps of the true squared distance <= (|x|+C)^2, so
//   second - best > 1.1 eps [ (4d+4)|x|C + 4C^2 + 2(d+2)(|x|+C)^2 ]
// proves that the reference's strict-'<' scan picks the same centroid.  Otherwise (near ties,
// exact ties, NaN) the row is re-scanned in the reference's exact arithmetic.  The winner's
// distance is always evaluated in the exact direct form.
template <typename F, int D>
__device__ __forceinline__ bool scan_centroids_certified(const F* __restrict__ cs, const F* __restrict__ cn, int k,
                                                         F cmax, const F* __restrict__ xrow, F& best_dist, int& bidx) {
  F x[D];
#pragma unroll
  for (int j = 0; j < D; ++j) x[j] = xrow[j];
  F xx = (F)0;
#pragma unroll
  for (int j = 0; j < D; ++j) xx = fma(x[j], x[j], xx);
  F b0 = (F)INFINITY, s0 = (F)INFINITY, b1 = (F)INFINITY, s1 = (F)INFINITY;
  int i0 = 0, i1 = 1;
  int c = 0;
  for (; c + 2 <= k; c += 2) {
    F d0 = (F)0, d1 = (F)0;
#pragma unroll
    for (int j = 0; j < D; ++j) {
      d0 = fma(cs[c * D + j], x[j], d0);
      d1 = fma(cs[(c + 1) * D + j], x[j], d1);
    }
    const F w0 = fma((F)-2, d0, cn[c]);
    const F w1 = fma((F)-2, d1, cn[c + 1]);
    top2_update<F>(w0, c, b0, s0, i0);
    top2_update<F>(w1, c + 1, b1, s1, i1);
  }
  if (c < k) {
    F d0 = (F)0;
#pragma unroll
    for (int j = 0; j < D; ++j) d0 = fma(cs[c * D + j], x[j], d0);
    const F w0 = fma((F)-2, d0, cn[c]);
    top2_update<F>(w0, c, b0, s0, i0);
  }
  F best, second;
  int bi;
  if (b1 < b0 || (b1 == b0 && i1 < i0)) { best = b1; bi = i1; second = (b0 < s1) ? b0 : s1; }
  else { best = b0; bi = i0; second = (b1 < s0) ? b1 : s0; }
  const F eps = sizeof(F) == 4 ? (F)5.9604645e-8 : (F)1.1102230246251565e-16;
  const F xn = sqrt(xx) * (F)1.000001;
  // + an absolute term for scores in the denormal range, where roundings are multiples of the smallest denormal
  const F dmin = sizeof(F) == 4 ? (F)1.4012985e-45 : (F)4.9406564584124654e-324;
  const F margin = fma((F)1.1 * eps, (F)(4 * D + 4) * xn * cmax + (F)4 * cmax * cmax + (F)(2 * (D + 2)) * (xn + cmax) * (xn + cmax),
                       (F)(8 * D + 16) * dmin * ((F)1 + xn + cmax));
  if ((second - best) > margin) {  // NaN anywhere -> not certified
    bidx = bi;
    best_dist = sqdist_reg<F, D>(cs + bi * D, x);
    return true;
  }
  scan_centroids<F, D>(cs, k, D, 0, xrow, true, best_dist, bidx);
  return false;
}

// coalesced global -> smem copy of `rows` contiguous rows into a padded tile
template <typename F>
__device__ __forceinline__ void load_tile(F* __restrict__ xs, const F* __restrict__ g, int rows, int d, int x_stride,
                                          int vec) {
  const int t = threadIdx.x;
  if (vec) {
    const int cpr = (d * (int)sizeof(F)) / 16;  // 16-byte chunks per row
    const int total = rows * cpr;
    const int4* g4 = reinterpret_cast<const int4*>(g);
    const int step_r = kThreads / cpr, step_c = kThreads % cpr;
    int r = t / cpr, c = t % cpr;
    for (int q = t; q < total; q += kThreads) {
      const int4 v = __ldg(g4 + q);
      *reinterpret_cast<int4*>(reinterpret_cast<char*>(xs) + ((size_t)r * x_stride * sizeof(F) + (size_t)c * 16)) = v;
      r += step_r; c += step_c;
      if (c >= cpr) { c -= cpr; r += 1; }
    }
  } else {
    const int total = rows * d;
    const int step_r = kThreads / d, step_c = kThreads % d;
    int r = t / d, c = t % d;
    for (int q = t; q < total; q += kThreads) {
      xs[(size_t)r * x_stride + c] = __ldg(g + q);
      r += step_r; c += step_c;
      if (c >= d) { c -= d; r += 1; }
    }
  }
}

template <typename F, int D, int MODE>
__global__ void __launch_bounds__(kThreads) lloyd_kernel(const LloydArgs<F> a) {
  pdl_launch_dependents();
  pdl_wait();
  if (a.state != nullptr && a.state->done) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = (D > 0) ? D : a.d;
  const int t = threadIdx.x;
  // carve: centroids | tile | accumulators | counts | labels | reduction scratch
  size_t off = 0;
  F* cs = reinterpret_cast<F*>(smem_raw + off);
  if (MODE != MODE_UPDATE) off += ((size_t)a.kc * d * sizeof(F) + 15) & ~(size_t)15;
  F* cnorm = reinterpret_cast<F*>(smem_raw + off);  // squared norms (+ max norm at [k]) of the certified path
  if (MODE != MODE_UPDATE && a.certified) off += ((size_t)(a.k + 1) * sizeof(F) + 15) & ~(size_t)15;
  const int tpr = (D > 0) ? 1 : a.tpr;
  const int tile_rows = kThreads / tpr;
  F* xs = reinterpret_cast<F*>(smem_raw + off);
  off += max(((size_t)tile_rows * a.x_stride * sizeof(F) + 15) & ~(size_t)15, (size_t)2048);
  F* acc = reinterpret_cast<F*>(smem_raw + off);
  uint32_t* cnt = nullptr;
  uint32_t* labs = nullptr;
  if (MODE != MODE_ASSIGN) {
    off += ((size_t)a.n_subsets * a.chunk_rows * d * sizeof(F) + 15) & ~(size_t)15;
    cnt = reinterpret_cast<uint32_t*>(smem_raw + off);
    off += ((size_t)a.n_subsets * a.chunk_rows * sizeof(uint32_t) + 15) & ~(size_t)15;
    labs = reinterpret_cast<uint32_t*>(smem_raw + off);
    off += kThreads * sizeof(uint32_t);
    for (int e = t; e < a.n_subsets * a.chunk_rows * d; e += kThreads) acc[e] = (F)0;
    for (int e = t; e < a.n_subsets * a.chunk_rows; e += kThreads) cnt[e] = 0u;
  }
  const int n_chunks = (MODE == MODE_UPDATE) ? 1 : (a.k + a.kc - 1) / a.kc;
  if (MODE != MODE_UPDATE && n_chunks == 1) {
    for (int e = t; e < a.k * d; e += kThreads) cs[e] = a.C[e];
  }
  __syncthreads();
  if (MODE != MODE_UPDATE && D > 0 && a.certified) {
    for (int c = t; c < a.k; c += kThreads) {
      double nn = 0.0;
      for (int j = 0; j < d; ++j) nn += (double)cs[c * d + j] * (double)cs[c * d + j];
      cnorm[c] = (F)nn;
    }
    __syncthreads();
    if (t == 0) {
      F m = (F)0;
      for (int c = 0; c < a.k; ++c) m = (cnorm[c] > m) ? cnorm[c] : m;  // NaN norms are skipped: such rows fall back
      cnorm[a.k] = sqrt(m) * (F)1.000001;
    }
    __syncthreads();
  }
  unsigned int my_fallbacks = 0;

  // accumulation role of this thread
  int sub = 0, feat = t;
  bool acc_active = false;
  if (MODE != MODE_ASSIGN) {
    if (d <= kThreads) { sub = t / d; feat = t - sub * d; acc_active = sub < a.n_subsets; }
    else { sub = 0; feat = t; acc_active = true; }
  }
  int cur_lab = -1;
  F cur_acc = (F)0;
  double inert = 0.0;

  const unsigned long long n_tiles = (a.n + tile_rows - 1) / tile_rows;
  const int my_row = t / tpr, my_q = t - my_row * tpr;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long row0 = tile * tile_rows;
    const int rows = (int)min((unsigned long long)tile_rows, a.n - row0);
    load_tile<F>(xs, a.X + row0 * (unsigned long long)d, rows, d, a.x_stride, a.vec);
    if (MODE == MODE_UPDATE) {
      for (int r = t; r < rows; r += kThreads) {
        const uint32_t lab = a.labels_in[row0 + r];
        labs[r] = lab;
        const int lc = (int)lab - a.chunk_base;
        if ((unsigned)lc < (unsigned)a.chunk_rows) atomicAdd(&cnt[lc], 1u);
      }
    }
    __syncthreads();
    if (MODE != MODE_UPDATE) {
      F best = (F)0;
      int bidx = 0;
      const bool row_ok = my_row < rows;
      const F* xrow = xs + (size_t)(row_ok ? my_row : 0) * a.x_stride;
      if (tpr == 1) {
        if (n_chunks == 1) {
          if constexpr (D > 0) {
            if (row_ok) {
              if (a.certified) my_fallbacks += scan_centroids_certified<F, D>(cs, cnorm, a.k, cnorm[a.k], xrow, best, bidx) ? 0u : 1u;
              else scan_centroids<F, D>(cs, a.k, d, 0, xrow, true, best, bidx);
            }
          } else {
            if (row_ok) scan_centroids<F, D>(cs, a.k, d, 0, xrow, true, best, bidx);
          }
        } else {
          for (int ch = 0; ch < n_chunks; ++ch) {
            const int cb = ch * a.kc, rc = min(a.kc, a.k - cb);
            __syncthreads();
            for (int e = t; e < rc * d; e += kThreads) cs[e] = a.C[(size_t)cb * d + e];
            __syncthreads();
            if (row_ok) scan_centroids<F, D>(cs, rc, d, cb, xrow, ch == 0, best, bidx);
          }
        }
      } else {
        best = (F)INFINITY;
        bidx = -1;
        for (int ch = 0; ch < n_chunks; ++ch) {
          const int cb = ch * a.kc, rc = min(a.kc, a.k - cb);
          if (n_chunks > 1) {
            __syncthreads();
            for (int e = t; e < rc * d; e += kThreads) cs[e] = a.C[(size_t)cb * d + e];
            __syncthreads();
          }
          if (row_ok) scan_centroids_strided<F>(cs, rc, d, cb, xrow, ch == 0 && my_q == 0, my_q, tpr, best, bidx);
        }
        combine_candidates<F>(tpr, best, bidx);
      }
      if (row_ok && my_q == 0) {
        const unsigned long long row = row0 + my_row;
        if (a.labels) a.labels[row] = (uint32_t)bidx;
        F outd = best;
        if (a.dists) {
          if (a.min_combine) {
            const F old = a.dists[row];
            if (best < old) a.dists[row] = best; else outd = old;
          } else {
            a.dists[row] = best;
          }
        }
        if (a.wdists) a.wdists[row] = a.weights ? outd * a.weights[row] : outd;
        if (MODE == MODE_FUSED) {
          labs[my_row] = (uint32_t)bidx;
          inert += (double)best;
          atomicAdd(&cnt[bidx], 1u);  // integer atomics: exact, order-independent (copy 0 of the counts)
        }
      }
      if (MODE == MODE_FUSED) __syncthreads();
    }
    if (MODE != MODE_ASSIGN) {
      if (acc_active) {
        if (d <= kThreads) {
          for (int r = sub; r < rows; r += a.n_subsets) {
            const int lab = (int)labs[r] - a.chunk_base;
            if ((unsigned)lab >= (unsigned)a.chunk_rows) continue;
            const F xv = xs[(size_t)r * a.x_stride + feat];
            if (lab != cur_lab) {
              if (cur_lab >= 0) acc[((size_t)sub * a.chunk_rows + cur_lab) * d + feat] += cur_acc;
              cur_lab = lab;
              cur_acc = (F)0;
            }
            cur_acc += xv;
          }
        } else {
          for (int r = 0; r < rows; ++r) {
            const int lab = (int)labs[r] - a.chunk_base;
            if ((unsigned)lab >= (unsigned)a.chunk_rows) continue;
            for (int j = t; j < d; j += kThreads) acc[(size_t)lab * d + j] += xs[(size_t)r * a.x_stride + j];
          }
        }
      }
    }
    __syncthreads();
  }

  if (MODE != MODE_UPDATE && a.certified && a.state_rw != nullptr && my_fallbacks)
    atomicAdd(&a.state_rw->fallback_rows, (unsigned long long)my_fallbacks);
  if (MODE != MODE_ASSIGN) {
    if (acc_active && cur_lab >= 0) acc[((size_t)sub * a.chunk_rows + cur_lab) * d + feat] += cur_acc;
    __syncthreads();
    const int kd = a.chunk_rows * d;
    F* ps = a.part_sums + ((size_t)blockIdx.x * a.k + a.chunk_base) * d;
    for (int e = t; e < kd; e += kThreads) {
      F s = acc[e];
      for (int q = 1; q < a.n_subsets; ++q) s += acc[(size_t)q * kd + e];
      ps[e] = s;
    }
    uint32_t* pc = a.part_counts + (size_t)blockIdx.x * a.k + a.chunk_base;
    for (int e = t; e < a.chunk_rows; e += kThreads) {
      uint32_t s = cnt[e];
      for (int q = 1; q < a.n_subsets; ++q) s += cnt[q * a.chunk_rows + e];
      pc[e] = s;
    }
    if (MODE == MODE_FUSED) {
      double* red = reinterpret_cast<double*>(xs);  // tile is dead now
      __syncthreads();
      red[t] = inert;
      __syncthreads();
      for (int o = kThreads / 2; o > 0; o >>= 1) {
        if (t < o) red[t] += red[t + o];
        __syncthreads();
      }
      if (t == 0) a.part_inertia[blockIdx.x] = red[0];
    }
  }
}

// ---------------------------------------------------------------------------
// finalize: sum `n_src` partial vectors in a fixed order (f64), then either write the
// per-rank partial (NCCL fallback of the multi-GPU path, before the all-gather) or do
// the m_k-means update.  grid = ceil(k*d / 16) CTAs of 256 threads: thread (g, e) sums
// sources b = g, g+16, ... for element e; the 16 groups are then added in order.  The
// last CTA to finish reduces the shift and the inertia.  (Row-sharded runs with NVLink
// peer access use finalize_x_kernel below instead: one launch for the whole exchange.)
// ---------------------------------------------------------------------------
template <typename F, typename S, typename CT> struct FinalizeArgs {
  const S* src_sums;          // source b, element e at src_sums[b*sum_stride + e]
  const CT* src_counts;       // source b, cluster c at src_counts[b*cnt_stride + c]
  const double* src_inertia;  // source b at src_inertia[b*inertia_stride]
  size_t sum_stride, cnt_stride, inertia_stride;
  int n_src;
  int n_src_inertia;   // inertia partials when their number differs from n_src (0 = n_src)
  int k, d;
  const unsigned char* src_touched;   // optional [n_src][k]: which clusters a source wrote (see strided_sum)
  // reduce-only output (per-rank partial): [kd sums | k counts | 1 inertia] as f64
  double* rank_out;
  // update outputs
  const F* C_old;
  F* C_new;
  double* counts_out;    // k, members per cluster of this assignment (without the +1)
  double* shift_part;    // [grid]
  LloydState* state;
  double tol;
  unsigned long long max_iter;
};

// peers may be a whole kernel behind (uneven shards, exact re-scans): waits on their flags are bounded by time, not spins
constexpr unsigned long long kPeerWaitNs = 20ull * 1000ull * 1000ull * 1000ull;
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}

constexpr int kFinElems = 16;    // elements (of the k*d sums) per CTA
constexpr int kFinGroups = 16;   // source groups per element; CTA = kFinElems * kFinGroups threads
constexpr int kFinThreads = kFinElems * kFinGroups;

// fixed-order sum over the CTA: warp butterflies, then the warp sums in warp order
__device__ __forceinline__ double block_sum(double v, double* red) {
  const int t = threadIdx.x;
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((t & 31) == 0) red[t >> 5] = v;
  __syncthreads();
  double tot = 0.0;
#pragma unroll
  for (int w = 0; w < kFinThreads / 32; ++w) tot += red[w];
  return tot;
}

// sum of src[b*stride + e] over b = g, g+G, g+2G, ... : batches of 16 sources per thread are loaded before the first
// addition (one L2 round trip per 256 sources), then added into four chains in a fixed association — reproducible.
// `touched` (optional): one byte per (source, cluster); a source that never accumulated into cluster `c` left its slot
// unwritten, and the slot is read as zero (the fused pair kernel and the streaming update skip the zero-fill and the
// write-back of the clusters a CTA never saw — most of a k x d partial on ordered data)
template <typename S>
__device__ __forceinline__ double strided_sum(const S* __restrict__ src, size_t stride, size_t e, int g, int n_src,
                                              const unsigned char* __restrict__ touched = nullptr, size_t touched_stride = 0, int c = 0) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
  for (int base = g; base < n_src; base += 16 * kFinGroups) {
    S v[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) {
      const int b = base + q * kFinGroups;
      bool on = b < n_src;
      if (on && touched != nullptr) on = touched[(size_t)b * touched_stride + c] != 0;
      v[q] = on ? src[(size_t)b * stride + e] : S(0);
    }
#pragma unroll
    for (int q = 0; q < 16; q += 4) { s0 += (double)v[q]; s1 += (double)v[q + 1]; s2 += (double)v[q + 2]; s3 += (double)v[q + 3]; }
  }
  return (s0 + s1) + (s2 + s3);
}

template <typename F, typename S, typename CT, bool UPDATE>
__global__ void __launch_bounds__(kFinThreads) finalize_kernel(const FinalizeArgs<F, S, CT> a) {
  pdl_launch_dependents();
  pdl_wait();
  if (a.state->done) return;
  __shared__ double sh[kFinGroups][kFinElems];
  __shared__ double shc[kFinGroups][kFinElems];
  __shared__ double red[kFinThreads];
  __shared__ bool is_last;
  const int t = threadIdx.x, g = t / kFinElems, l = t % kFinElems;
  const int kd = a.k * a.d;
  const int e = blockIdx.x * kFinElems + l;
  sh[g][l] = (e < kd) ? strided_sum<S>(a.src_sums, a.sum_stride, (size_t)e, g, a.n_src, a.src_touched, (size_t)a.k, e / a.d) : 0.0;
  // counts of the clusters this CTA's elements belong to
  const int c_lo = (blockIdx.x * kFinElems) / a.d;
  const int c_hi = (min(kd, (int)(blockIdx.x + 1) * kFinElems) - 1) / a.d;
  const int c = c_lo + l;
  shc[g][l] = (c <= c_hi) ? strided_sum<CT>(a.src_counts, a.cnt_stride, (size_t)c, g, a.n_src) : 0.0;
  __syncthreads();
  // every cluster's count is published by the CTA that holds its first element
  if (g == 0 && c <= c_hi && (c * a.d) / kFinElems == (int)blockIdx.x) {
    double cntv = 0.0;
#pragma unroll
    for (int q = 0; q < kFinGroups; ++q) cntv += shc[q][l];
    if (UPDATE) a.counts_out[c] = cntv;
    else a.rank_out[kd + c] = cntv;
  }
  double sq = 0.0;
  if (g == 0 && e < kd) {
    double tot = 0.0;
#pragma unroll
    for (int q = 0; q < kFinGroups; ++q) tot += sh[q][l];
    if (!UPDATE) {
      a.rank_out[e] = tot;
    } else {
      const int ce = e / a.d - c_lo;
      double cntv = 0.0;
#pragma unroll
      for (int q = 0; q < kFinGroups; ++q) cntv += shc[q][ce];
      const F oldv = a.C_old[e];
      const F newv = (F)((tot + (double)oldv) / (cntv + 1.0));
      a.C_new[e] = newv;
      const double df = (double)(oldv - newv);
      sq = df * df;
    }
  }
  if (!UPDATE) {
    if (blockIdx.x == 0) {
      double v = 0.0;
      const int n_in = a.n_src_inertia ? a.n_src_inertia : a.n_src;
      for (int b = t; b < n_in; b += kFinThreads) v += a.src_inertia[(size_t)b * a.inertia_stride];
      const double tot = block_sum(v, red);
      if (t == 0) a.rank_out[kd + a.k] = tot;
    }
    return;
  }
  const double sq_cta = block_sum(sq, red);
  if (t == 0) {
    a.shift_part[blockIdx.x] = sq_cta;
    __threadfence();
    const unsigned int prev = atomicAdd(&a.state->blocks_done, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // last CTA: shift, inertia, loop control (fixed order)
  double v = 0.0;
  for (int b = t; b < (int)gridDim.x; b += kFinThreads) v += ((volatile double*)a.shift_part)[b];
  const double shift_sq = block_sum(v, red);
  v = 0.0;
  const int n_in = a.n_src_inertia ? a.n_src_inertia : a.n_src;
  for (int b = t; b < n_in; b += kFinThreads) v += a.src_inertia[(size_t)b * a.inertia_stride];
  const double inertia = block_sum(v, red);
  if (t == 0) {
    const F shift = (F)sqrt(shift_sq);  // L2Dist::distance: sqrt in f64, cast to F
    a.state->shift = (double)shift;
    a.state->inertia = inertia;
    const unsigned long long it = a.state->n_iter + 1;
    a.state->n_iter = it;
    a.state->blocks_done = 0u;
    if (shift < (F)a.tol || it >= a.max_iter) a.state->done = 1;
  }
}

// ---------------------------------------------------------------------------
// finalize_x: the multi-GPU iteration tail in ONE launch.  CTA b owns a contiguous range of 16-element slices of the
// k*d sums.  Phase A: fixed-order f64 sum of the local per-CTA partials, stored straight into every rank's gather buffer
// (own included) over NVLink, system fence, one flag per (rank, CTA).  Phase B: wait for the same CTA's flags of every
// rank (plus the CTAs that publish the counts of the clusters it touches), combine the `world` partials in rank order —
// identical inputs and order on every rank, hence bit-identical centroids without a broadcast — m_k-means update, shift.
// The last CTA to finish sums the shift and the per-rank inertia and runs the stop rule (KM/algorithm.rs:323-331).
// Every CTA of the grid is resident (grid <= 2 x SM count, 256 threads, no shared memory to speak of), and a CTA waits
// only for its own index on the peers, so the waits cannot deadlock.
// ---------------------------------------------------------------------------
constexpr int kFxMaxCtas = 1024;   // flag slots per (parity, rank)

template <typename F, typename S, typename CT> struct FinalizeXArgs {
  const S* src_sums;
  const CT* src_counts;
  const double* src_inertia;
  size_t sum_stride, cnt_stride, inertia_stride;
  int n_src, n_src_inertia, k, d;
  const unsigned char* src_touched;     // optional [n_src][k] (see strided_sum)
  unsigned long long* prof;             // LKM_STEP_TRACE: CTA 0 stamps %globaltimer at its phase boundaries (8 slots)
  int slices_per_cta;
  double* peer_out[8];                  // rank p's gather slot for THIS rank (this parity): [kd sums | k counts | inertia]
  unsigned long long* peer_flags[8];    // rank p's flags for THIS rank (this parity): [kFxMaxCtas]
  const double* gather;                 // own gather buffer of this parity: rank r at gather + r * cap
  const unsigned long long* my_flags;   // own flags of this parity: rank r, CTA b at my_flags[r * kFxMaxCtas + b]
  uint4* peer_ll[8];                    // finalize_cl, self-validating exchange: rank p's 16-byte slots for THIS rank (this parity), or nullptr
  const uint4* my_ll;                   // own slots of this parity: rank r, element e at my_ll[r * cap + e]
  size_t cap;
  int world;
  unsigned long long epoch;
  const F* C_old;
  F* C_new;
  double* counts_out;
  double* shift_part;   // [grid]
  LloydState* state;
  double tol;
  unsigned long long max_iter;
};

__device__ __forceinline__ bool wait_flag(const unsigned long long* f, unsigned long long want) {
  unsigned long long v;
  const unsigned long long t0 = global_ns();
  for (;;) {
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(f) : "memory");
    if (v == want) return true;
    if (global_ns() - t0 > kPeerWaitNs) return false;
  }
}

template <typename F, typename S, typename CT>
__global__ void __launch_bounds__(kFinThreads) finalize_x_kernel(const FinalizeXArgs<F, S, CT> a) {
  pdl_launch_dependents();
  pdl_wait();
  if (a.state->done) return;
  __shared__ double sh[kFinGroups][kFinElems];
  __shared__ double shc[kFinGroups][kFinElems];
  __shared__ double red[kFinThreads];
  __shared__ bool is_last;
  const int t = threadIdx.x, g = t / kFinElems, l = t % kFinElems;
  const int kd = a.k * a.d;
  const int n_slices = (kd + kFinElems - 1) / kFinElems;
  const int s_lo = blockIdx.x * a.slices_per_cta, s_hi = min(n_slices, s_lo + a.slices_per_cta);
  // ---- phase A: local reduction, published to every rank ----
  for (int slice = s_lo; slice < s_hi; ++slice) {
    const int e = slice * kFinElems + l;
    sh[g][l] = (e < kd) ? strided_sum<S>(a.src_sums, a.sum_stride, (size_t)e, g, a.n_src, a.src_touched, (size_t)a.k, e / a.d) : 0.0;
    const int c_lo = (slice * kFinElems) / a.d;
    const int c_hi = (min(kd, (slice + 1) * kFinElems) - 1) / a.d;
    const int c = c_lo + l;
    // a cluster's count is published with the slice that holds its first element
    const bool own_c = c <= c_hi && (c * a.d) / kFinElems == slice;
    shc[g][l] = own_c ? strided_sum<CT>(a.src_counts, a.cnt_stride, (size_t)c, g, a.n_src) : 0.0;
    __syncthreads();
    if (g == 0 && e < kd) {
      double tot = 0.0;
#pragma unroll
      for (int q = 0; q < kFinGroups; ++q) tot += sh[q][l];
      for (int p = 0; p < a.world; ++p) a.peer_out[p][e] = tot;
    }
    if (g == 1 && own_c) {
      double cntv = 0.0;
#pragma unroll
      for (int q = 0; q < kFinGroups; ++q) cntv += shc[q][l];
      for (int p = 0; p < a.world; ++p) a.peer_out[p][kd + c] = cntv;
    }
    __syncthreads();
  }
  if (blockIdx.x == 0) {
    double v = 0.0;
    const int n_in = a.n_src_inertia ? a.n_src_inertia : a.n_src;
    for (int b = t; b < n_in; b += kFinThreads) v += a.src_inertia[(size_t)b * a.inertia_stride];
    const double tot = block_sum(v, red);
    if (t == 0)
      for (int p = 0; p < a.world; ++p) a.peer_out[p][kd + a.k] = tot;
  }
  // the stores of every thread are ordered by the barrier; the flag-writing threads' fence is cumulative over them
  __syncthreads();
  if (t < a.world) {
    __threadfence_system();
    unsigned long long* f = a.peer_flags[t] + blockIdx.x;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(a.epoch) : "memory");
  }
  // ---- phase B: the peers' copies of this CTA's range (and of the counts of its clusters) have landed ----
  const int e_lo = s_lo * kFinElems, e_hi = min(kd, s_hi * kFinElems);
  bool ok = true;
  if (e_lo < e_hi) {
    const int c_first = e_lo / a.d;
    const int need_lo = min((int)blockIdx.x, ((c_first * a.d) / kFinElems) / a.slices_per_cta);
    const int n_need = ((int)blockIdx.x - need_lo + 1) * a.world;
    for (int i = t; i < n_need; i += kFinThreads)
      ok = wait_flag(a.my_flags + (size_t)(i % a.world) * kFxMaxCtas + need_lo + i / a.world, a.epoch) && ok;
  }
  if (!ok) atomicOr(&a.state->fallback_rows, 1ull << 41);   // timeout marker, reported by the host
  __syncthreads();
  double sq = 0.0;
  for (int e = e_lo + t; e < e_hi; e += kFinThreads) {
    const int c = e / a.d;
    double tot = 0.0, cntv = 0.0;
    for (int r = 0; r < a.world; ++r) {
      const volatile double* gr = a.gather + (size_t)r * a.cap;
      tot += gr[e];
      cntv += gr[kd + c];
    }
    const F oldv = a.C_old[e];
    const F newv = (F)((tot + (double)oldv) / (cntv + 1.0));
    a.C_new[e] = newv;
    if (e == c * a.d) a.counts_out[c] = cntv;
    const double df = (double)(oldv - newv);
    sq += df * df;
  }
  const double sq_cta = block_sum(sq, red);
  if (t == 0) {
    a.shift_part[blockIdx.x] = sq_cta;
    __threadfence();
    const unsigned int prev = atomicAdd(&a.state->blocks_done, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  // last CTA: shift, inertia (published by CTA 0 of every rank), loop control — fixed order, identical on every rank
  if (t < a.world && !wait_flag(a.my_flags + (size_t)t * kFxMaxCtas, a.epoch)) atomicOr(&a.state->fallback_rows, 1ull << 41);
  __syncthreads();
  double v = 0.0;
  for (int b = t; b < (int)gridDim.x; b += kFinThreads) v += ((volatile double*)a.shift_part)[b];
  const double shift_sq = block_sum(v, red);
  if (t == 0) {
    double inertia = 0.0;
    for (int r = 0; r < a.world; ++r) inertia += ((const volatile double*)a.gather)[(size_t)r * a.cap + kd + a.k];
    const F shift = (F)sqrt(shift_sq);  // L2Dist::distance: sqrt in f64, cast to F
    a.state->shift = (double)shift;
    a.state->inertia = inertia;
    const unsigned long long it = a.state->n_iter + 1;
    a.state->n_iter = it;
    a.state->blocks_done = 0u;
    if (shift < (F)a.tol || it >= a.max_iter) a.state->done = 1;
  }
}

// ---------------------------------------------------------------------------
// finalize_cl: the iteration tail for SPARSE partials (src_touched set: fused pair kernel + streaming update), on one
// GPU or row-sharded.  The slice-wise kernels above give every thread ~n_src / 16 strided 4-byte loads in two dependent
// rounds; with k * d = 32768 and 296 sources that is 2048 CTAs of latency chains (60-90 us measured on the cfg3 shape,
// a quarter of a 1.25M-row step).  Here CTA b owns whole clusters: it compacts the sources that touched cluster c into a
// list (ascending source order), warp w sums the sources list[w], list[w + 8], ... (lane = feature, four sources and
// four features per lane in flight: coalesced rows, a handful of L2 round trips per cluster), and the eight warp partials
// are added in warp order — a fixed association for a given list, hence reproducible.  Counts are integers: their f64
// sum is exact in any order.
// Row-sharded: the CTA publishes its clusters' sums and counts to every rank over NVLink, waits for the same clusters
// of every rank and combines in rank order.  With peer_ll the exchange is self-validating (the LL scheme of the
// collective libraries): every value travels as one 16-byte store {lo, epoch, hi, epoch}, each 8-byte half carrying its
// own tag, and the reader polls the slot until both tags are this epoch's — no fence, no flag, nothing waits for an
// NVLink acknowledgement (the system fence of the flag protocol cost 4-17 us per iteration at 8 GPUs,
// profiles/r2_scale_trace_8gpu.txt).  Slots are double-buffered by epoch parity; a writer can be at most one epoch ahead
// of a reader (it needs the reader's previous publication to get there), so a slot is never overwritten while it is read.
// Without peer_ll: plain stores, one system fence per flag, flags (as finalize_x_kernel).
// slices_per_cta is read as "clusters per CTA".
// ---------------------------------------------------------------------------
constexpr int kFclMaxSrc = 2048;
constexpr int kFclCols = 128;   // features per pass of the source sum (4 per lane)

__device__ __forceinline__ void ll_store(uint4* p, double v, uint32_t tag) {
  const unsigned long long b = (unsigned long long)__double_as_longlong(v);
  asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"((uint32_t)b), "r"(tag), "r"((uint32_t)(b >> 32)), "r"(tag)
               : "memory");
}
__device__ __forceinline__ uint4 ll_peek(const uint4* p) {
  uint4 v;
  asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  return v;
}
// sum over the ranks (in rank order) of slot e; all `world` slots are requested before the first is looked at, slots whose
// tags are not yet this epoch's are asked for again.  false: a peer did not publish within kPeerWaitNs.
__device__ __forceinline__ bool ll_gather_sum(const uint4* base, size_t cap, int world, size_t e, uint32_t tag, double& out) {
  uint4 v[8];
  unsigned long long t0 = 0;
  for (;;) {
    bool all = true;
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < world) v[r] = ll_peek(base + (size_t)r * cap + e);
#pragma unroll
    for (int r = 0; r < 8; ++r)
      if (r < world) all = all && v[r].y == tag && v[r].w == tag;
    if (all) break;
    const unsigned long long now = global_ns();
    if (t0 == 0) t0 = now;
    else if (now - t0 > kPeerWaitNs) { out = 0.0; return false; }
  }
  double tot = 0.0;
#pragma unroll
  for (int r = 0; r < 8; ++r)
    if (r < world) tot += __longlong_as_double((long long)(((unsigned long long)v[r].z << 32) | (unsigned long long)v[r].x));
  out = tot;
  return true;
}

template <typename F, typename S, typename CT>
__global__ void __launch_bounds__(kFinThreads) finalize_cl_kernel(const FinalizeXArgs<F, S, CT> a) {
  pdl_launch_dependents();
  pdl_wait();
  if (a.state->done) return;
  static_assert(kFinThreads == 256, "eight warps");
  __shared__ int list[kFclMaxSrc];
  __shared__ int warp_tot[kFinThreads / 32];
  __shared__ int n_list_s;
  __shared__ double red[kFinThreads];
  __shared__ double part[kFinThreads / 32][kFclCols];
  __shared__ bool is_last;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int kd = a.k * a.d;
  const int c_lo = blockIdx.x * a.slices_per_cta, c_hi = min(a.k, c_lo + a.slices_per_cta);
  const bool ll = a.world > 1 && a.peer_ll[0] != nullptr;
  const uint32_t tag = (uint32_t)a.epoch;
  double sq = 0.0;
  const bool stamp = a.prof != nullptr && blockIdx.x == 0 && t == 0;
  if (stamp) a.prof[0] = global_ns();
  for (int c = c_lo; c < c_hi; ++c) {
    // ---- sources that touched cluster c, in ascending order ----
    __syncthreads();
    if (t == 0) n_list_s = 0;
    __syncthreads();
    for (int b0 = 0; b0 < a.n_src; b0 += kFinThreads) {
      const int b = b0 + t;
      const bool on = b < a.n_src && a.src_touched[(size_t)b * a.k + c] != 0;
      const unsigned int m = __ballot_sync(0xffffffffu, on);
      if (lane == 0) warp_tot[warp] = __popc(m);
      __syncthreads();
      int base = n_list_s;
      for (int w = 0; w < warp; ++w) base += warp_tot[w];
      if (on) list[base + __popc(m & ((1u << lane) - 1u))] = b;
      __syncthreads();
      if (t == 0) { int tot = 0; for (int w = 0; w < kFinThreads / 32; ++w) tot += warp_tot[w]; n_list_s += tot; }
      __syncthreads();
    }
    const int nl = n_list_s;
    // ---- counts: integers, exact in f64 whatever the order ----
    double cv = 0.0;
    for (int i = t; i < nl; i += kFinThreads) cv += (double)a.src_counts[(size_t)list[i] * a.cnt_stride + c];
    cv = block_sum(cv, red);
    if (t == 0) {
      if (ll) for (int p = 0; p < a.world; ++p) ll_store(a.peer_ll[p] + kd + c, cv, tag);
      else if (a.world > 1) for (int p = 0; p < a.world; ++p) a.peer_out[p][kd + c] = cv;
      else a.counts_out[c] = cv;
    }
    // ---- sums: warp w takes the sources list[w], list[w + 8], ..., lane l the features j0 + l + 32 q ----
    for (int j0 = 0; j0 < a.d; j0 += kFclCols) {
      double acc[4] = {0.0, 0.0, 0.0, 0.0};
      const S* col = a.src_sums + (size_t)c * a.d + j0 + lane;
      bool in[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) in[q] = j0 + lane + 32 * q < a.d;
      int i = warp;
      for (; i + 24 < nl; i += 32) {
        S v[4][4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const S* row = col + (size_t)list[i + 8 * u] * a.sum_stride;
#pragma unroll
          for (int q = 0; q < 4; ++q) v[u][q] = in[q] ? row[32 * q] : S(0);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
          for (int q = 0; q < 4; ++q) acc[q] += (double)v[u][q];
      }
      for (; i < nl; i += 8) {
        const S* row = col + (size_t)list[i] * a.sum_stride;
        S v[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) v[q] = in[q] ? row[32 * q] : S(0);
#pragma unroll
        for (int q = 0; q < 4; ++q) acc[q] += (double)v[q];
      }
#pragma unroll
      for (int q = 0; q < 4; ++q) part[warp][lane + 32 * q] = acc[q];
      __syncthreads();
      if (t < kFclCols && j0 + t < a.d) {
        const size_t e = (size_t)c * a.d + j0 + t;
        double tot = 0.0;
#pragma unroll
        for (int w = 0; w < kFinThreads / 32; ++w) tot += part[w][t];
        if (ll) {
          for (int p = 0; p < a.world; ++p) ll_store(a.peer_ll[p] + e, tot, tag);
        } else if (a.world > 1) {
          for (int p = 0; p < a.world; ++p) a.peer_out[p][e] = tot;
        } else {
          const F oldv = a.C_old[e];
          const F newv = (F)((tot + (double)oldv) / (cv + 1.0));
          a.C_new[e] = newv;
          const double df = (double)(oldv - newv);
          sq += df * df;
        }
      }
      __syncthreads();
    }
  }
  if (a.world > 1) {
    if (blockIdx.x == 0) {
      double v = 0.0;
      const int n_in = a.n_src_inertia ? a.n_src_inertia : a.n_src;
      for (int b = t; b < n_in; b += kFinThreads) v += a.src_inertia[(size_t)b * a.inertia_stride];
      const double tot = block_sum(v, red);
      if (t == 0) {
        if (ll) for (int p = 0; p < a.world; ++p) ll_store(a.peer_ll[p] + kd + a.k, tot, tag);
        else for (int p = 0; p < a.world; ++p) a.peer_out[p][kd + a.k] = tot;
      }
    }
    if (stamp) a.prof[1] = global_ns();
    bool ok = true;
    if (!ll) {
      // every thread's NVLink stores happen before the barrier; the fence of the flag-writing threads after it is cumulative
      // over them (one system fence per flag instead of one per thread: the fence waits for the peers' acknowledgements)
      __syncthreads();
      if (t < a.world) {
        __threadfence_system();
        unsigned long long* f = a.peer_flags[t] + blockIdx.x;
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(a.epoch) : "memory");
      }
      if (stamp) a.prof[2] = global_ns();
      if (t < a.world) ok = wait_flag(a.my_flags + (size_t)t * kFxMaxCtas + blockIdx.x, a.epoch);
      __syncthreads();
    } else if (stamp) a.prof[2] = global_ns();
    bool first = true;
    for (int c = c_lo; c < c_hi; ++c) {
      double cntv = 0.0;
      if (ll) ok = ll_gather_sum(a.my_ll, a.cap, a.world, (size_t)kd + c, tag, cntv) && ok;
      else for (int r = 0; r < a.world; ++r) cntv += ((const volatile double*)a.gather)[(size_t)r * a.cap + kd + c];
      if (t == 0) a.counts_out[c] = cntv;
      for (int j = t; j < a.d; j += kFinThreads) {
        const size_t e = (size_t)c * a.d + j;
        double tot = 0.0;
        if (ll) ok = ll_gather_sum(a.my_ll, a.cap, a.world, e, tag, tot) && ok;
        else for (int r = 0; r < a.world; ++r) tot += ((const volatile double*)a.gather)[(size_t)r * a.cap + e];
        const F oldv = a.C_old[e];
        const F newv = (F)((tot + (double)oldv) / (cntv + 1.0));
        a.C_new[e] = newv;
        const double df = (double)(oldv - newv);
        sq += df * df;
      }
      if (first && stamp) a.prof[3] = global_ns();
      first = false;
    }
    if (!ok) atomicOr(&a.state->fallback_rows, 1ull << 41);   // timeout marker, reported by the host
  }
  const double sq_cta = block_sum(sq, red);
  if (stamp) a.prof[4] = global_ns();
  if (t == 0) {
    a.shift_part[blockIdx.x] = sq_cta;
    __threadfence();
    const unsigned int prev = atomicAdd(&a.state->blocks_done, 1u);
    is_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  if (a.prof != nullptr && t == 0) a.prof[5] = global_ns();
  __threadfence();
  if (a.world > 1 && !ll && t < a.world && !wait_flag(a.my_flags + (size_t)t * kFxMaxCtas, a.epoch)) atomicOr(&a.state->fallback_rows, 1ull << 41);
  __syncthreads();
  double v = 0.0;
  for (int b = t; b < (int)gridDim.x; b += kFinThreads) v += ((volatile double*)a.shift_part)[b];
  const double shift_sq = block_sum(v, red);
  double inertia = 0.0;
  if (a.world == 1) {
    v = 0.0;
    const int n_in = a.n_src_inertia ? a.n_src_inertia : a.n_src;
    for (int b = t; b < n_in; b += kFinThreads) v += a.src_inertia[(size_t)b * a.inertia_stride];
    inertia = block_sum(v, red);
  }
  if (t == 0) {
    if (ll) {
      if (!ll_gather_sum(a.my_ll, a.cap, a.world, (size_t)kd + a.k, tag, inertia)) atomicOr(&a.state->fallback_rows, 1ull << 41);
    } else if (a.world > 1) {
      for (int r = 0; r < a.world; ++r) inertia += ((const volatile double*)a.gather)[(size_t)r * a.cap + kd + a.k];
    }
    const F shift = (F)sqrt(shift_sq);  // L2Dist::distance: sqrt in f64, cast to F
    a.state->shift = (double)shift;
    a.state->inertia = inertia;
    const unsigned long long it = a.state->n_iter + 1;
    a.state->n_iter = it;
    a.state->blocks_done = 0u;
    if (shift < (F)a.tol || it >= a.max_iter) a.state->done = 1;
    if (a.prof != nullptr) a.prof[6] = global_ns();
  }
}

// ---------------------------------------------------------------------------
// small utilities
// ---------------------------------------------------------------------------
template <typename F>
__global__ void gather_rows_kernel(const F* __restrict__ X, int d, const long long* __restrict__ idx, int k,
                                   F* __restrict__ out) {
  // idx[c] < 0: row not on this rank, leave out[c] untouched
  for (int c = blockIdx.x; c < k; c += gridDim.x) {
    const long long r = idx[c];
    if (r < 0) continue;
    for (int j = threadIdx.x; j < d; j += blockDim.x) out[(size_t)c * d + j] = X[(size_t)r * d + j];
  }
}

__device__ __forceinline__ unsigned long long mix64(unsigned long long z) {
  z += 0x9e3779b97f4a7c15ULL;
  z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
  z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double unit_open(unsigned long long bits) {  // (0,1)
  return ((double)(bits >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// blobs in the sense of datasets/src/generate.rs:28-58; one thread per element,
// N(0,1) by Box-Muller from two counter-keyed hashes (counter = global element index).
template <typename F>
__global__ void gen_blobs_kernel(F* __restrict__ X, unsigned long long n, int d, unsigned long long n_blobs,
                                 const F* __restrict__ centres, F sigma, unsigned long long seed,
                                 unsigned long long row_offset, unsigned long long n_global, int shuffled) {
  // shuffled != 0: every row draws its blob from a hash of its global index instead of the contiguous
  // generate.rs layout (same noise): what real, unordered data looks like to the update path
  const unsigned long long total = n * (unsigned long long)d;
  const unsigned long long m = n_global / n_blobs;
  for (unsigned long long e = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; e < total;
       e += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long r = e / d, j = e - r * d;
    const unsigned long long gr = row_offset + r;
    unsigned long long b = m ? gr / m : 0;
    if (b >= n_blobs) b = n_blobs - 1;
    if (shuffled) b = mix64(seed ^ 0x5bd1e995u ^ mix64(~gr)) % n_blobs;
    const unsigned long long ctr = gr * (unsigned long long)d + j;
    const unsigned long long h1 = mix64(seed ^ mix64(2 * ctr)), h2 = mix64(seed ^ mix64(2 * ctr + 1));
    const double u1 = unit_open(h1), u2 = unit_open(h2);
    const double z = sqrt(-2.0 * log(u1)) * cospi(2.0 * u2);
    X[e] = (F)((double)centres[b * d + j] + (double)sigma * z);
  }
}

// WeightedIndex::new: sequential prefix sum in F on ONE thread (bit-exact with
// rand 0.8.5).  cum[i] = w_0 + ... + w_i for i < n-1 ; out[0] = total,
// out[1] = 1 if any weight is negative/NaN (InvalidWeight).
template <typename F>
__global__ void __launch_bounds__(256) seq_prefix_kernel(const F* __restrict__ w, unsigned long long n,
                                                         F* __restrict__ cum, F* __restrict__ out, F carry_in,
                                                         unsigned long long n_cum) {
  // carry_in / n_cum: row-sharded use — this rank continues the prefix of the ranks before it and
  // writes cum entries for its first n_cum rows (n for inner ranks, n-1 on the last rank)
  constexpr int CH = 2048;
  __shared__ F buf[CH];
  __shared__ F carry;
  __shared__ int bad;
  if (threadIdx.x == 0) { carry = carry_in; bad = 0; }
  __syncthreads();
  for (unsigned long long base = 0; base < n; base += CH) {
    const int m = (int)min((unsigned long long)CH, n - base);
    for (int i = threadIdx.x; i < m; i += blockDim.x) buf[i] = w[base + i];
    __syncthreads();
    if (threadIdx.x == 0) {
      F total = carry;
      int b = 0;
      for (int i = 0; i < m; ++i) {
        const F wi = buf[i];
        if (!(wi >= (F)0)) b = 1;
        total = Rn<F>::add(total, wi);
        buf[i] = total;  // inclusive prefix
      }
      carry = total;
      if (b) bad = 1;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < m; i += blockDim.x)
      if (base + i < n_cum) cum[base + i] = buf[i];
    __syncthreads();
  }
  if (threadIdx.x == 0) { out[0] = carry; out[1] = bad ? (F)1 : (F)0; }
}

// partition point of the non-decreasing cum[0..m): #{j : cum[j] <= chosen}
template <typename F>
__global__ void count_le_kernel(const F* __restrict__ cum, unsigned long long m, F chosen,
                                unsigned long long* __restrict__ out) {
  unsigned long long local = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < m;
       i += (unsigned long long)gridDim.x * blockDim.x)
    local += (cum[i] <= chosen) ? 1ull : 0ull;
  for (int o = 16; o > 0; o >>= 1) local += __shfl_down_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(out, local);
}

// fixed-order f64 block sums of w (blocks of kPickBlock elements) for the fast pick
constexpr int kPickBlock = 2048;
template <typename F>
__global__ void __launch_bounds__(256) block_sums_kernel(const F* __restrict__ w, unsigned long long n,
                                                         double* __restrict__ bsum) {
  __shared__ double red[256];
  const unsigned long long nb = (n + kPickBlock - 1) / kPickBlock;
  for (unsigned long long b = blockIdx.x; b < nb; b += gridDim.x) {
    const unsigned long long base = b * kPickBlock;
    double s = 0.0;
#pragma unroll
    for (int q = 0; q < kPickBlock / 256; ++q) {
      const unsigned long long i = base + threadIdx.x + q * 256;
      if (i < n) s += (double)w[i];
    }
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
    if (threadIdx.x == 0) bsum[b] = red[0];
    __syncthreads();
  }
}

// fast pick inside ONE block of kPickBlock weights (the host has already walked the block sums
// in order and found the block whose running prefix crosses the threshold): 256 threads, 8
// elements each, fixed-association prefix (thread-serial, warp shuffle scan, warp totals), first
// element whose running prefix exceeds thr.  out_idx[0] = row index (the block's last row if
// rounding hides the crossing).
template <typename F>
__global__ void __launch_bounds__(256) fast_pick_kernel(const F* __restrict__ w, unsigned long long n,
                                                        unsigned long long block, double prefix_before, double thr,
                                                        long long* __restrict__ out_idx) {
  constexpr int PER = kPickBlock / 256;
  __shared__ double wsum[8];
  __shared__ unsigned long long found;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const unsigned long long base = block * kPickBlock + (unsigned long long)t * PER;
  const unsigned long long end = min(n, (block + 1) * (unsigned long long)kPickBlock);
  double v[PER];
  double local = 0.0;
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    v[i] = (base + i < end) ? (double)w[base + i] : 0.0;
    local += v[i];
  }
  if (t == 0) found = ~0ull;
  double incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const double up = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += up;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  double woff = 0.0;
  for (int q = 0; q < warp; ++q) woff += wsum[q];
  double run = prefix_before + woff + (incl - local);
#pragma unroll
  for (int i = 0; i < PER; ++i) {
    run += v[i];
    if (run > thr && base + i < end) { atomicMin(&found, base + i); break; }
  }
  __syncthreads();
  if (t == 0) out_idx[0] = (found == ~0ull) ? (long long)end - 1 : (long long)found;
}

// Fixed order of the FAST pick over the block sums, shared by the host walk (weighted_pick) and the device walk
// (fast_pick_device_kernel) so that both pick the same row: groups of kPickGroup consecutive blocks are summed in order,
// the group sums in order give the total, and the threshold is located group first, then block.
constexpr int kPickGroup = 64;
constexpr int kPickMaxGroups = 2048;   // shared-memory group sums of the device walk: up to 2048 * 64 * 2048 = 268M weights
__host__ __device__ inline double pick_group_sum(const double* bsum, unsigned long long nb, unsigned long long g) {
  double s = 0.0;
  const unsigned long long b1 = (g + 1) * kPickGroup < nb ? (g + 1) * kPickGroup : nb;
  for (unsigned long long b = g * kPickGroup; b < b1; ++b) s += bsum[b];
  return s;
}
__host__ __device__ inline void pick_locate(const double* bsum, unsigned long long nb, const double* gsum, unsigned long long ng,
                                            double thr, unsigned long long& fb, double& fp) {
  double run = 0.0;
  unsigned long long g = 0;
  for (; g < ng; ++g) {
    const double nxt = run + gsum[g];
    if (nxt > thr) break;
    run = nxt;
  }
  if (g == ng) { g = ng - 1; run -= gsum[g]; }   // rounding hid the crossing: the last group
  const unsigned long long b0 = g * kPickGroup, b1 = (g + 1) * kPickGroup < nb ? (g + 1) * kPickGroup : nb;
  for (unsigned long long b = b0; b < b1; ++b) {
    const double nxt = run + bsum[b];
    if (nxt > thr) { fb = b; fp = run; return; }
    run = nxt;
  }
  fb = b1 - 1;
  fp = run - bsum[b1 - 1];
}

// The whole FAST pick on the device, for picks 2..k of k-means++ (no host round trip between the picks: the k uniforms are
// drawn up front, the kernels of all picks are queued back to back).  One CTA: thread 0 walks the block sums in order (the
// same two walks the host does in weighted_pick: total, then the block whose running prefix crosses thr = u01 * total), all
// threads finish inside that block like fast_pick_kernel, then copy the chosen row behind the centroids chosen so far.
// No positive total (every remaining distance zero): row 0, the reference's `.unwrap_or(0)` (KM/init.rs:149-152).
template <typename F>
__global__ void __launch_bounds__(256) fast_pick_device_kernel(const F* __restrict__ w, unsigned long long n,
                                                               const double* __restrict__ bsum, unsigned long long nb, F u01,
                                                               const F* __restrict__ X, int d, F* __restrict__ dst,
                                                               long long* __restrict__ out_idx) {
  constexpr int PER = kPickBlock / 256;
  __shared__ double wsum[8];
  __shared__ unsigned long long found, s_block;
  __shared__ double s_prefix, s_thr;
  __shared__ int s_none;
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  __shared__ double gsum[kPickMaxGroups];
  const unsigned long long ng = (nb + kPickGroup - 1) / kPickGroup;
  for (unsigned long long g = t; g < ng; g += 256) gsum[g] = pick_group_sum(bsum, nb, g);
  __syncthreads();
  if (t == 0) {
    double total = 0.0;
    for (unsigned long long g = 0; g < ng; ++g) total += gsum[g];
    s_none = !(total > 0.0);
    const double thr = (double)(u01 * (F)total);   // Uniform<F>::new(0, total).sample: value0_1 * scale, scale = (F)total
    unsigned long long fb = 0;
    double fp = 0.0;
    pick_locate(bsum, nb, gsum, ng, thr, fb, fp);
    s_block = fb; s_prefix = fp; s_thr = thr;
    found = ~0ull;
  }
  __syncthreads();
  long long idx = 0;
  if (!s_none) {
    const unsigned long long block = s_block;
    const unsigned long long base = block * kPickBlock + (unsigned long long)t * PER;
    const unsigned long long end = min(n, (block + 1) * (unsigned long long)kPickBlock);
    double v[PER];
    double local = 0.0;
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      v[i] = (base + i < end) ? (double)w[base + i] : 0.0;
      local += v[i];
    }
    double incl = local;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += up;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    double woff = 0.0;
    for (int q = 0; q < warp; ++q) woff += wsum[q];
    double run = s_prefix + woff + (incl - local);
#pragma unroll
    for (int i = 0; i < PER; ++i) {
      run += v[i];
      if (run > s_thr && base + i < end) { atomicMin(&found, base + i); break; }
    }
    __syncthreads();
    idx = (found == ~0ull) ? (long long)end - 1 : (long long)found;
  }
  if (t == 0 && out_idx != nullptr) out_idx[0] = idx;
  for (int j = t; j < d; j += 256) dst[j] = X[(size_t)idx * d + j];
}

// k-means|| oversampling (KM/init.rs:239-270): keep row i iff U_i < mult * d_i / cost,
// U_i from a counter-based generator keyed by (seed, global row).  Selected global
// row ids are appended through an atomic cursor (the host sorts them ascending).
template <typename F>
__global__ void bernoulli_select_kernel(const F* __restrict__ dists, unsigned long long n, F mult, F cost,
                                        unsigned long long seed, unsigned long long row_offset,
                                        unsigned long long* __restrict__ out, unsigned int cap,
                                        unsigned int* __restrict__ cursor) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long bits = mix64(seed ^ mix64(row_offset + i));
    const F rnd = (F)((double)(bits >> 11) * (1.0 / 9007199254740992.0));
    const F prob = mult * dists[i] / cost;
    if (rnd < prob) {
      const unsigned int p = atomicAdd(cursor, 1u);
      if (p < cap) out[p] = row_offset + i;
    }
  }
}

// histogram of labels (integer atomics: exact, order-independent) -> counts as F
static __global__ void label_hist_kernel(const uint32_t* __restrict__ labels, unsigned long long n,
                                  unsigned int* __restrict__ hist) {
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x)
    atomicAdd(hist + labels[i], 1u);
}
template <typename F>
__global__ void u32_to_float_kernel(const unsigned int* __restrict__ in, int m, F* __restrict__ out) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += gridDim.x * blockDim.x) out[i] = (F)in[i];
}

// compute_centroids_incremental (KM/algorithm.rs:815-835): per-sample running
// mean, order dependent => rows are walked in order by ONE CTA; thread j owns
// feature j (and j+blockDim, ...).  Every warp keeps its own replica of the
// counts (wcounts[warp][k]) so no block barrier is needed per row.
template <typename F>
__global__ void __launch_bounds__(1024) incremental_update_kernel(const F* __restrict__ X,
                                                                  const uint32_t* __restrict__ labels,
                                                                  unsigned long long n, int d, int k, F* __restrict__ C,
                                                                  F* __restrict__ counts, F* __restrict__ wcounts) {
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31, nw = blockDim.x >> 5;
  F* mine = wcounts + (size_t)warp * k;
  for (int c = lane; c < k; c += 32) mine[c] = counts[c];
  __syncwarp();
  for (unsigned long long i = 0; i < n; ++i) {
    const uint32_t c = labels[i];
    F cnt = (F)0;
    if (lane == 0) { cnt = Rn<F>::add(mine[c], (F)1); mine[c] = cnt; }
    cnt = __shfl_sync(0xffffffffu, cnt, 0);
    for (int j = t; j < d; j += blockDim.x) {
      const F cen = C[(size_t)c * d + j];
      const F shift = Rn<F>::sub(X[i * (unsigned long long)d + j], cen) / cnt;
      C[(size_t)c * d + j] = Rn<F>::add(cen, shift);
    }
  }
  __syncwarp();
  if (warp == 0)
    for (int c = lane; c < k; c += 32) counts[c] = mine[c];
  (void)nw;
}

// L2Dist::distance(C_old, C_new) exactly as the reference evaluates it
// (sequential F accumulation over k*d, sqrt in f64, cast to F); one thread.
template <typename F>
__global__ void seq_l2_distance_kernel(const F* __restrict__ a, const F* __restrict__ b, int len, F* __restrict__ out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    F acc = (F)0;
    for (int e = 0; e < len; ++e) {
      const F df = Rn<F>::sub(a[e], b[e]);
      acc = Rn<F>::add(acc, Rn<F>::mul(df, df));
    }
    out[0] = (F)sqrt((double)acc);
  }
}

// fixed-order f64 sum of an F vector: per-block partials then a single-CTA pass
template <typename F>
__global__ void __launch_bounds__(256) sum_f64_kernel(const F* __restrict__ x, unsigned long long n,
                                                      double* __restrict__ part) {
  __shared__ double red[256];
  double s = 0.0;
  const unsigned long long per = (n + gridDim.x - 1) / gridDim.x;
  const unsigned long long lo = blockIdx.x * per, hi = min(n, lo + per);
  for (unsigned long long i = lo + threadIdx.x; i < hi; i += 256) s += (double)x[i];
  red[threadIdx.x] = s;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) { if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o]; __syncthreads(); }
  if (threadIdx.x == 0) part[blockIdx.x] = red[0];
}

}  // namespace lkm

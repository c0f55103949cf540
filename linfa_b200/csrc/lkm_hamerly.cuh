// Hamerly's accelerated K-Means (KM/algorithm.rs:248-291, HamerlyAlgorithm :396-566) on the device.
//
// What the reference does per iteration: nearest_inter_centroid_distances (:462-470), reassign_observations (:472-524:
// skip a row while upper <= max(s[label] / 2, lower); otherwise tighten upper, and if that is not enough run
// two_closest_centroids), move the changed rows between the running sums, recompute_centroids (:527-553, the same
// m_k-means update as Lloyd), update_bounds (:555-568).  The result differs from Lloyd's only in what `inertia` means
// (compute_inertia against the POST-update centroids, :570-577) and in whose memberships give cluster_count (the best run's).
//
// On a GPU the point of the bounds is HBM traffic: a row whose bounds hold costs 12 bytes (label, upper, lower) instead of
// d * sizeof(F).  Layout: labels u32[n], upper F[n], lower F[n]; running sums as 64-bit fixed point [k][d] (exact,
// order-independent integer atomics -> reproducible; the reference's own sums drift with every add / subtract in F).
//
//   hamerly_bounds_kernel    update_bounds of the previous iteration fused with this iteration's bound test; rows that
//                            fail tighten `upper` with one exact distance (warp per row, coalesced); rows that still fail
//                            are appended to a list
//   rows_top2_kernel         (lkm_rows.cuh) two_closest_centroids for the listed rows, new bounds, sums moved
//   hamerly_centroids_kernel new centroids from the fixed-point sums, squared shift per CTA
//   hamerly_moved_kernel     distances_moved[c] and s[c] = distance to the nearest other centroid (one CTA per centroid)
//   hamerly_state_kernel     convergence_dist, stop rule, two_farthest_indices (:585-603)
//   inertia_rows_kernel      compute_inertia: sum of rdistance(x_i, c_label(i)) in the direct form
//
// Bounds are kept conservative by a few ulps (lkm_rows.cuh), so a skipped row is one whose label the reference's exact
// arithmetic would not change either; the labels equal Lloyd's, which is what the reference's own tests assert
// (hamerly_lloyd_equivalence_*, KM/algorithm.rs:1266-1345).  Ties between centroids 0 and 1 go to 0 here (Lloyd's rule) and
// to 1 in two_closest_centroids (:901-903); tie-free data is what parity is stated on.
#pragma once
#include "lkm_rows.cuh"

namespace lkm {

struct HamState {
  int far1, far2;                  // two_farthest_indices of the last distances_moved
  unsigned int list_count;         // rows appended by hamerly_bounds_kernel in this iteration
  unsigned int pad;
  unsigned long long rows_tightened, rows_scanned, rows_changed, rows_exact;   // totals over the run
};

template <typename F> struct HamArgs {
  const F* X;
  unsigned long long n;
  int d, k;
  const F* C;             // current centroids
  const uint32_t* labels;
  F* ub;
  F* lb;
  const F* moved;         // distances_moved of the previous recompute (nullptr: bounds are fresh)
  const F* s;             // distance of every centroid to its nearest other centroid
  uint32_t* list;
  HamState* hs;
  const LloydState* state;
};

template <typename F>
__global__ void __launch_bounds__(256) hamerly_bounds_kernel(const HamArgs<F> a) {
  if (a.state != nullptr && a.state->done) return;
  const int lane = threadIdx.x & 31;
  const int far1 = a.hs->far1, far2 = a.hs->far2;
  const F slack = (F)1 + ((F)a.d + (F)4) * rows_eps<F>();
  unsigned int tightened = 0;
  const unsigned long long n_blk = (a.n + 255) / 256;
  for (unsigned long long blk = blockIdx.x; blk < n_blk; blk += gridDim.x) {
    const unsigned long long row = blk * 256 + threadIdx.x;
    const bool valid = row < a.n;
    uint32_t m = 0;
    F ub = (F)0, lb = (F)0, thr = (F)0;
    if (valid) {
      m = a.labels[row];
      ub = a.ub[row];
      lb = a.lb[row];
      if (a.moved != nullptr) {   // update_bounds (KM/algorithm.rs:555-568)
        ub += a.moved[m];
        lb -= ((int)m == far1) ? a.moved[far2] : a.moved[far1];
      }
      thr = fmax(a.s[m] / (F)2, lb);
    }
    bool fail = valid && ub > thr;
    // tighten: upper = distance(x, c_label), one warp per failing row (features across the lanes)
    uint32_t todo = __ballot_sync(0xffffffffu, fail);
    tightened += (lane == 0) ? (unsigned)__popc(todo) : 0u;
    while (todo) {
      const int src = __ffs(todo) - 1;
      todo &= todo - 1u;
      const unsigned long long r = blk * 256 + (threadIdx.x & ~31) + src;
      const uint32_t mm = __shfl_sync(0xffffffffu, m, src);
      const F* xp = a.X + r * (unsigned long long)a.d;
      const F* cp = a.C + (size_t)mm * a.d;
      F acc = (F)0;
      for (int j = lane; j < a.d; j += 32) {
        const F df = cp[j] - __ldg(xp + j);
        acc = fma(df, df, acc);
      }
      for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
      if (lane == src) ub = (F)sqrt((double)acc) * slack;
    }
    fail = fail && ub > thr;
    if (valid) { a.ub[row] = ub; a.lb[row] = lb; }
    // rows whose bounds still do not hold: two_closest_centroids in rows_top2_kernel
    const uint32_t fm = __ballot_sync(0xffffffffu, fail);
    if (fm) {
      unsigned int base = 0;
      if (lane == 0) base = atomicAdd(&a.hs->list_count, (unsigned)__popc(fm));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (fail) a.list[base + __popc(fm & ((1u << lane) - 1u))] = (uint32_t)row;
    }
  }
  if (lane == 0 && tightened) atomicAdd(&a.hs->rows_tightened, (unsigned long long)tightened);
}

// The running sums of HamerlyAlgorithm::new (KM/algorithm.rs:441-446) in fixed point: sum over the rows of
// round(x * scale), exactly the integers that rows_top2_kernel later moves between clusters — a cluster that loses all
// its rows is back at exactly zero (an f32 / f64 sum would keep a rounding residue, and (residue + old) / (0 + 1) makes
// an empty cluster's centroid creep by that residue in every iteration).  One warp per 32 consecutive rows, lane =
// feature: runs of equal labels are summed in registers, one integer atomic per run and feature.
template <typename F>
__global__ void __launch_bounds__(256) hamerly_sums_kernel(const F* __restrict__ X, unsigned long long n, int d,
                                                           const uint32_t* __restrict__ labels, double scale,
                                                           long long* __restrict__ fx_sums, long long* __restrict__ fx_counts) {
  const int lane = threadIdx.x & 31;
  const unsigned long long n_groups = (n + 31) / 32;
  const unsigned long long warps = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  for (unsigned long long g = (((unsigned long long)blockIdx.x * blockDim.x) + threadIdx.x) >> 5; g < n_groups; g += warps) {
    const unsigned long long r0 = g * 32;
    const int rows = (int)min(32ull, n - r0);
    const uint32_t lab = (lane < rows) ? labels[r0 + lane] : 0xFFFFFFFFu;
    // counts: one atomic per distinct label of the group
    const uint32_t peers = __match_any_sync(0xffffffffu, lab);
    if (lane < rows && lane == __ffs(peers) - 1)
      atomicAdd(reinterpret_cast<unsigned long long*>(fx_counts) + lab, (unsigned long long)__popc(peers));
    for (int j = lane; j - lane < d; j += 32) {
      const bool on = j < d;
      long long acc = 0;
      uint32_t cur = __shfl_sync(0xffffffffu, lab, 0);
      for (int r = 0; r < rows; ++r) {
        const uint32_t L = __shfl_sync(0xffffffffu, lab, r);
        if (L != cur) {
          if (on && acc != 0) atomicAdd(reinterpret_cast<unsigned long long*>(fx_sums) + (size_t)cur * d + j, (unsigned long long)acc);
          acc = 0;
          cur = L;
        }
        if (on) acc += __double2ll_rn((double)X[(r0 + r) * (unsigned long long)d + j] * scale);
      }
      if (on && acc != 0) atomicAdd(reinterpret_cast<unsigned long long*>(fx_sums) + (size_t)cur * d + j, (unsigned long long)acc);
    }
  }
}

template <typename F> struct HamCenArgs {
  const long long* fx_sums;
  const long long* fx_counts;
  double inv_scale;
  int k, d;
  const F* C_old;
  F* C_new;
  double* shift_part;   // [grid]
  double* counts_out;   // k
  const LloydState* state;
};

// recompute_centroids (KM/algorithm.rs:527-553): (sum + old) / (count + 1)
template <typename F>
__global__ void __launch_bounds__(256) hamerly_centroids_kernel(const HamCenArgs<F> a) {
  if (a.state->done) return;
  __shared__ double red[kFinThreads];
  const int e = blockIdx.x * 256 + threadIdx.x;
  double sq = 0.0;
  if (e < a.k * a.d) {
    const int c = e / a.d;
    const F oldv = a.C_old[e];
    const double cnt = (double)a.fx_counts[c];
    const F newv = (F)(((double)a.fx_sums[e] * a.inv_scale + (double)oldv) / (cnt + 1.0));
    a.C_new[e] = newv;
    const double df = (double)(oldv - newv);
    sq = df * df;
    if (e == c * a.d) a.counts_out[c] = cnt;
  }
  const double tot = block_sum(sq, red);
  if (threadIdx.x == 0) a.shift_part[blockIdx.x] = tot;
}

template <typename F> struct HamMovedArgs {
  const F* C_old;
  const F* C_new;
  int k, d;
  F* moved;   // k
  F* s;       // k
  const LloydState* state;
};

// one CTA per centroid: distances_moved[c] = distance(old_c, new_c); s[c] = min over the other centroids of
// distance(new_c, new_c') — the `second` of two_closest_centroids(new, new_c), whose closest is c itself at 0
template <typename F>
__global__ void __launch_bounds__(128) hamerly_moved_kernel(const HamMovedArgs<F> a) {
  if (a.state != nullptr && a.state->done) return;
  __shared__ F red[128];
  const int c = blockIdx.x, t = threadIdx.x;
  const F* me = a.C_new + (size_t)c * a.d;
  F best = (F)INFINITY;
  for (int o = t; o < a.k; o += 128) {
    if (o == c) continue;
    const F v = sqdist_mem<F>(a.C_new + (size_t)o * a.d, me, a.d);
    best = (v < best) ? v : best;
  }
  red[t] = best;
  __syncthreads();
  for (int o = 64; o > 0; o >>= 1) {
    if (t < o) red[t] = (red[t + o] < red[t]) ? red[t + o] : red[t];
    __syncthreads();
  }
  if (t == 0) {
    // with a single centroid two_closest_centroids returns (0, 0, 0)
    a.s[c] = (a.k == 1) ? (F)0 : (F)sqrt((double)red[0]);
    if (a.C_old != nullptr) a.moved[c] = (F)sqrt((double)sqdist_mem<F>(a.C_old + (size_t)c * a.d, me, a.d));
  }
}

template <typename F> struct HamStateArgs {
  const double* shift_part;
  int n_part;
  const F* moved;
  int k;
  HamState* hs;
  LloydState* state;   // nullptr: only the farthest pair
  double tol;
  unsigned long long max_iter;
};

// convergence_dist and the stop rule of fit_hamerly (KM/algorithm.rs:272-276); two_farthest_indices (:585-603)
template <typename F>
__global__ void hamerly_state_kernel(const HamStateArgs<F> a) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  if (a.state != nullptr && a.state->done) return;
  a.hs->rows_scanned += a.hs->list_count;
  a.hs->list_count = 0u;
  int f1 = 0, f2 = 0;
  if (a.k >= 2) {
    if (a.moved[1] >= a.moved[0]) { f1 = 1; f2 = 0; } else { f1 = 0; f2 = 1; }
    for (int i = 2; i < a.k; ++i) {
      const F v = a.moved[i];
      if (v >= a.moved[f1]) { f2 = f1; f1 = i; }
      else if (v > a.moved[f2]) f2 = i;
    }
  }
  a.hs->far1 = f1;
  a.hs->far2 = f2;
  if (a.state == nullptr) return;
  double sq = 0.0;
  for (int b = 0; b < a.n_part; ++b) sq += a.shift_part[b];
  const F shift = (F)sqrt(sq);
  a.state->shift = (double)shift;
  const unsigned long long it = a.state->n_iter + 1;
  a.state->n_iter = it;
  if (shift < (F)a.tol || it >= a.max_iter) a.state->done = 1;
}

// compute_inertia (KM/algorithm.rs:608-620): sum of rdistance(x_i, c_label(i)).  HBM-bound: `lpr` lanes share a row (16-byte
// loads when VEC > 1), 32 / lpr rows per warp step; every lane sums the squares of its features of a row in F, the rows in
// f64, then the CTA tree — a fixed order, the documented accumulation rule of this library (thread = row with serial feature
// order read X at a sixth of the HBM rate: 4.8 ms for the 5.12 GB of cfg3).
template <typename F, int VEC>
__global__ void __launch_bounds__(256) inertia_rows_kernel(const F* __restrict__ X, unsigned long long n, int d, const F* __restrict__ C,
                                                           const uint32_t* __restrict__ labels, double* __restrict__ part, int lpr) {
  __shared__ double red[kFinThreads];
  const int lane = threadIdx.x & 31;
  const int rps = 32 / lpr, sub_row = lane / lpr, sub_lane = lane % lpr;
  const unsigned long long warp = ((unsigned long long)blockIdx.x * 256 + threadIdx.x) >> 5, n_warps = (unsigned long long)gridDim.x * 8;
  double acc = 0.0;
  for (unsigned long long row = warp * rps + sub_row; row < n; row += n_warps * rps) {
    const F* xr = X + row * (unsigned long long)d;
    const F* cr = C + (size_t)labels[row] * d;
    F e = (F)0;
    for (int j = sub_lane * VEC; j < d; j += lpr * VEC) {
      if constexpr (VEC == 4) {
        const float4 x = __ldg(reinterpret_cast<const float4*>(xr + j)), c = __ldg(reinterpret_cast<const float4*>(cr + j));
        const float d0 = c.x - x.x, d1 = c.y - x.y, d2 = c.z - x.z, d3 = c.w - x.w;
        e = fmaf(d3, d3, fmaf(d2, d2, fmaf(d1, d1, fmaf(d0, d0, e))));
      } else if constexpr (VEC == 2) {
        const double2 x = __ldg(reinterpret_cast<const double2*>(xr + j)), c = __ldg(reinterpret_cast<const double2*>(cr + j));
        const double d0 = c.x - x.x, d1 = c.y - x.y;
        e = fma(d1, d1, fma(d0, d0, e));
      } else {
        const F df = __ldg(cr + j) - __ldg(xr + j);
        e = fma(df, df, e);
      }
    }
    acc += (double)e;
  }
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) part[blockIdx.x] = tot;
}

template <typename F>
__global__ void fill_kernel(F* __restrict__ p, unsigned long long n, F v) {
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (unsigned long long)gridDim.x * blockDim.x) p[i] = v;
}

// max |x_ij| of the resident matrix (scale of the fixed-point sums)
template <typename F>
__global__ void __launch_bounds__(256) abs_max_kernel(const F* __restrict__ X, unsigned long long count, unsigned long long* __restrict__ out_bits) {
  double m = 0.0;
  for (unsigned long long i = (unsigned long long)blockIdx.x * 256 + threadIdx.x; i < count; i += (unsigned long long)gridDim.x * 256) {
    const double v = fabs((double)X[i]);
    m = (v > m) ? v : m;   // NaN is skipped
  }
  for (int o = 16; o > 0; o >>= 1) { const double ov = __shfl_xor_sync(0xffffffffu, m, o); m = (ov > m) ? ov : m; }
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, (unsigned long long)__double_as_longlong(m));   // non-negative doubles order like their bits
}

}  // namespace lkm

// rows_top2_kernel: the reference's closest centroid (and the runner-up) for a LIST of rows — the shared second stage of
// every screened path.  The tensor-core assignment kernels append the rows their certificate cannot decide to a list
// instead of re-scanning them inline (one warp per row inside a 10-warp CTA was the tail of every iteration from a
// k-means++ start: 34 000 listed rows cost 3 ms); Hamerly's bound test lists the rows whose bounds do not hold
// (two_closest_centroids, KM/algorithm.rs:889-919).  Here they are processed thread = row at full occupancy:
//
//   1. the row sits in registers (D a template parameter), the centroids in shared memory (chunks of kc rows);
//   2. expanded-form FMA scores w_c = ||c||^2 - 2 x.c with a running (best, runner-up) and the certificate of
//      scan_centroids_certified (lkm_kernels.cuh): a gap above the rounding bound proves that the reference's strict-'<'
//      scan of the direct form picks the same centroid;
//   3. rows that miss it (near ties, ties, NaN) are scanned again in the reference's exact arithmetic (direct form in feature
//      order, never contracted), lowest index on ties;
//   4. the winner's squared distance is always evaluated in the exact direct form.
//
// Optional outputs: Hamerly's bounds (upper = distance to the winner, lower = a proven lower bound on the distance to the
// runner-up), and the move of the row between Hamerly's running centroid sums (KM/algorithm.rs:507-523) as 64-bit
// fixed-point atomics — integer adds are exact and order-independent, so the sums are reproducible run to run.
#pragma once
#include "lkm_kernels.cuh"

namespace lkm {

template <typename F> struct RowsArgs {
  const F* X;
  unsigned long long n;
  int d, k;
  const F* C;                       // k x d
  const uint32_t* list;             // rows to visit; nullptr: rows [0, all_rows)
  const unsigned int* list_count;   // device count of list entries (nullptr: all_rows entries)
  unsigned long long all_rows;
  int kc;                           // centroids per shared-memory chunk
  uint32_t* labels;                 // out: label of every visited row (in: its previous label when fx_sums is set)
  F* best_dist;                     // out (optional): min squared distance, direct form, per visited row
  int min_combine;                  // best_dist = min(best_dist, new) (running minimum of the seeding rounds, KM/init.rs:138-143)
  F* ub;                            // out (optional): Hamerly upper bound (distance to the winner)
  F* lb;                            // out (optional): Hamerly lower bound (distance to the runner-up)
  long long* fx_sums;               // optional: [k][d] fixed-point running sums, moved from the old to the new label
  long long* fx_counts;             // [k]
  double fx_scale;
  const LloydState* state;
  unsigned long long* n_exact;      // rows that needed step 3
  unsigned long long* n_changed;    // rows whose label changed (fx_sums set)
};

constexpr int kRowsThreads = 256;

template <typename F> __device__ __forceinline__ F rows_eps() { return sizeof(F) == 4 ? (F)5.9604645e-8 : (F)1.1102230246251565e-16; }

// load one row into registers (16-byte loads when the row size allows it)
template <typename F, int D>
__device__ __forceinline__ void rows_load(const F* __restrict__ p, F (&x)[D]) {
  constexpr int PER = 16 / (int)sizeof(F);
  if constexpr ((D % PER) == 0) {
    const int4* p4 = reinterpret_cast<const int4*>(p);
#pragma unroll
    for (int q = 0; q < D / PER; ++q) {
      const int4 v = __ldg(p4 + q);
      if constexpr (sizeof(F) == 4) {
        x[4 * q] = __int_as_float(v.x); x[4 * q + 1] = __int_as_float(v.y);
        x[4 * q + 2] = __int_as_float(v.z); x[4 * q + 3] = __int_as_float(v.w);
      } else {
        x[2 * q] = __hiloint2double(v.y, v.x);
        x[2 * q + 1] = __hiloint2double(v.w, v.z);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < D; ++j) x[j] = __ldg(p + j);
  }
}

template <typename F, int D>
__global__ void __launch_bounds__(kRowsThreads) rows_top2_kernel(const RowsArgs<F> a) {
  pdl_launch_dependents();
  pdl_wait();
  if (a.state != nullptr && a.state->done) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = (D > 0) ? D : a.d;
  F* cs = reinterpret_cast<F*>(smem_raw);            // [kc][d]
  F* cn = cs + (size_t)a.kc * d;                     // [kc] squared norms
  const int t = threadIdx.x;
  const unsigned long long total = (a.list != nullptr && a.list_count != nullptr) ? (unsigned long long)*a.list_count : a.all_rows;
  const unsigned long long n_tiles = (total + kRowsThreads - 1) / kRowsThreads;
  const int n_chunks = (a.k + a.kc - 1) / a.kc;
  const F eps = rows_eps<F>();
  const F dmin = sizeof(F) == 4 ? (F)1.4012985e-45 : (F)4.9406564584124654e-324;
  unsigned int my_exact = 0, my_changed = 0;
  bool staged = false;

  auto stage = [&](int ch) {
    const int cb = ch * a.kc, rc = min(a.kc, a.k - cb);
    __syncthreads();
    for (int e = t; e < rc * d; e += kRowsThreads) cs[e] = a.C[(size_t)cb * d + e];
    __syncthreads();
    for (int c = t; c < rc; c += kRowsThreads) {
      double nn = 0.0;
      for (int j = 0; j < d; ++j) nn += (double)cs[(size_t)c * d + j] * (double)cs[(size_t)c * d + j];
      cn[c] = (F)nn;
    }
    __syncthreads();
    return rc;
  };

  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long idx = tile * kRowsThreads + t;
    const bool valid = idx < total;
    const unsigned long long row = valid ? (a.list != nullptr ? (unsigned long long)a.list[idx] : idx) : 0ull;
    const F* xp = a.X + row * (unsigned long long)d;
    F x[D > 0 ? D : 1];
    F xx = (F)0;
    if constexpr (D > 0) {
      rows_load<F, D>(xp, x);
#pragma unroll
      for (int j = 0; j < D; ++j) xx = fma(x[j], x[j], xx);
    } else {
      for (int j = 0; j < d; ++j) { const F v = __ldg(xp + j); xx = fma(v, v, xx); }
    }
    // ---- step 2: FMA scores, running best / runner-up ----
    F b0 = (F)INFINITY, s0 = (F)INFINITY, t0 = (F)INFINITY, cm = (F)0;   // the three lowest scores, indices of the first two
    int i0 = 0, j0 = 0;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int cb = ch * a.kc;
      int rc = min(a.kc, a.k - cb);
      if (n_chunks > 1 || !staged) { rc = stage(ch); staged = true; }
      int c = 0;
      for (; c + 2 <= rc; c += 2) {
        F d0 = (F)0, d1 = (F)0;
        if constexpr (D > 0) {
#pragma unroll
          for (int j = 0; j < D; ++j) {
            d0 = fma(cs[c * D + j], x[j], d0);
            d1 = fma(cs[(c + 1) * D + j], x[j], d1);
          }
        } else {
          for (int j = 0; j < d; ++j) {
            const F v = __ldg(xp + j);
            d0 = fma(cs[(size_t)c * d + j], v, d0);
            d1 = fma(cs[(size_t)(c + 1) * d + j], v, d1);
          }
        }
        const F n0 = cn[c], n1 = cn[c + 1];
        cm = fmax(cm, fmax(n0, n1));
        top3_update<F>(fma((F)-2, d0, n0), cb + c, b0, s0, t0, i0, j0);
        top3_update<F>(fma((F)-2, d1, n1), cb + c + 1, b0, s0, t0, i0, j0);
      }
      if (c < rc) {
        F d0 = (F)0;
        if constexpr (D > 0) {
#pragma unroll
          for (int j = 0; j < D; ++j) d0 = fma(cs[c * D + j], x[j], d0);
        } else {
          for (int j = 0; j < d; ++j) d0 = fma(cs[(size_t)c * d + j], __ldg(xp + j), d0);
        }
        const F n0 = cn[c];
        cm = fmax(cm, n0);
        top3_update<F>(fma((F)-2, d0, n0), cb + c, b0, s0, t0, i0, j0);
      }
    }
    const F best = b0, second = s0, third = t0;
    int bi = i0;
    const F xn = sqrt(xx) * (F)1.000001, cmax = sqrt(cm) * (F)1.000001;
    const F dd = (F)d;
    // Certificate.  |w_c - W_c| <= errdot = eps [(2d+2)|x|C + 2C^2] for the FMA-form scores; the reference's direct form is
    // within (d+2) eps of the TRUE squared distance (a sum of non-negative terms), which the scores bound from above as
    // D_c = w_c + |x|^2 + errdot + (d+1) eps |x|^2.  So the reference's strict-'<' scan picks `best` when
    //   second - best > 2 errdot + (d+2) eps (D_best + D_second).
    // (scan_centroids_certified in lkm_kernels.cuh bounds the distances by (|x|+C)^2 instead: simpler, and looser for rows
    // far from the origin.  Measured and rejected here: a second scoring pass in f64 for the rows this misses — slower than
    // the exact scan it saves — and a CTA-per-row exact scan — tiles from a k-means++ start hold dozens of open rows, which
    // thread = row handles side by side.)
    const F errdot = eps * (((F)2 * dd + (F)2) * xn * cmax + (F)2 * cmax * cmax);
    const F tiny = ((F)8 * dd + (F)16) * dmin * ((F)1 + xn + cmax);
    auto margin_to = [&](F other) {
      const F dsum = fmax(best + xx, (F)0) + fmax(other + xx, (F)0) + (F)2 * errdot + (F)2 * (dd + (F)1) * eps * xx;
      return (F)1.1 * ((F)2 * errdot + (dd + (F)2) * eps * dsum) + tiny;
    };
    const bool cert = (second - best) > margin_to(second);   // NaN / Inf anywhere -> false
    // The same certificate against the THIRD lowest score: every centroid but the first two is proven farther than `best`
    // in the reference's arithmetic, so its scan is decided between those two — two exact distances instead of k.  From a
    // k-means++ start most open rows are of this kind (a blob shared by two centroids); a warp with one such row used to
    // walk all k centroids in step 3.
    const bool cert3 = (third - best) > margin_to(third);
    const bool need_pair = valid && !cert && cert3;
    const bool need_exact = valid && !cert && !cert3;
    // ---- step 3: the reference's scan for the rows without a certificate ----
    F eb = (F)INFINITY, es = (F)INFINITY;
    int ei = 0;
    if (__syncthreads_or(need_exact ? 1 : 0)) {
      for (int ch = 0; ch < n_chunks; ++ch) {
        const int cb = ch * a.kc;
        int rc = min(a.kc, a.k - cb);
        if (n_chunks > 1) rc = stage(ch);
        if (need_exact) {
          // centroid 0 seeds the scan (a NaN there is sticky, like in the reference); strict '<' keeps the lowest index
          auto take = [&](F v, int gc) {
            if (gc == 0) { eb = v; ei = 0; }
            else if (v < eb) { es = eb; eb = v; ei = gc; }
            else if (v < es) es = v;
          };
          int c = 0;
          if constexpr (D > 0) {
            // four centroids at a time: each distance is a chain of D dependent additions, four chains fill the pipe
            for (; c + 4 <= rc; c += 4) {
              const F v0 = sqdist_reg<F, D>(cs + (c + 0) * D, x), v1 = sqdist_reg<F, D>(cs + (c + 1) * D, x);
              const F v2 = sqdist_reg<F, D>(cs + (c + 2) * D, x), v3 = sqdist_reg<F, D>(cs + (c + 3) * D, x);
              take(v0, cb + c); take(v1, cb + c + 1); take(v2, cb + c + 2); take(v3, cb + c + 3);
            }
          }
          for (; c < rc; ++c) {
            F v;
            if constexpr (D > 0) v = sqdist_reg<F, D>(cs + c * D, x);
            else v = sqdist_mem<F>(cs + (size_t)c * d, xp, d);
            take(v, cb + c);
          }
        }
      }
      if (need_exact) { bi = ei; ++my_exact; }
    }
    if (!valid) continue;
    F bd = eb;
    if (need_pair) {
      F e0, e1;
      if constexpr (D > 0) { e0 = sqdist_reg<F, D>(a.C + (size_t)i0 * D, x); e1 = sqdist_reg<F, D>(a.C + (size_t)j0 * D, x); }
      else { e0 = sqdist_mem<F>(a.C + (size_t)i0 * d, xp, d); e1 = sqdist_mem<F>(a.C + (size_t)j0 * d, xp, d); }
      // the strict-'<' scan in index order keeps the lower index on equal distances
      const bool second_wins = e1 < e0 || (e1 == e0 && j0 < i0);
      bi = second_wins ? j0 : i0;
      bd = second_wins ? e1 : e0;
      es = second_wins ? e0 : e1;
      ++my_exact;
    } else if (!need_exact) {
      if constexpr (D > 0) bd = sqdist_reg<F, D>(a.C + (size_t)bi * D, x);
      else bd = sqdist_mem<F>(a.C + (size_t)bi * d, xp, d);
    }
    const uint32_t old = a.labels != nullptr ? a.labels[row] : 0u;
    if (a.labels != nullptr) a.labels[row] = (uint32_t)bi;
    if (a.best_dist != nullptr) {
      if (!a.min_combine || bd < a.best_dist[row]) a.best_dist[row] = bd;
    }
    if (a.ub != nullptr) {
      // bounds must bound the TRUE distances: the direct form is within (d + 2) eps of them
      const F slack = (dd + (F)4) * eps;
      a.ub[row] = (F)sqrt((double)bd) * ((F)1 + slack);
      // a proven lower bound on the true squared distance behind an FMA-form score
      auto below = [&](F w) {
        return (w + xx) - (F)1.1 * eps * (((F)2 * dd + (F)2) * xn * cmax + (F)2 * cmax * cmax + (dd + (F)2) * xx) - tiny;
      };
      F lsq;
      if (need_exact) lsq = es * ((F)1 - (F)2 * slack);
      else if (need_pair) lsq = fmin(es * ((F)1 - (F)2 * slack), below(third));   // the runner-up is the loser or a third centroid
      else lsq = below(second);
      a.lb[row] = lsq > (F)0 ? (F)sqrt((double)lsq) * ((F)1 - slack) : (F)0;
    }
    if (a.fx_sums != nullptr && old != (uint32_t)bi) {
      ++my_changed;
      unsigned long long* so = reinterpret_cast<unsigned long long*>(a.fx_sums) + (size_t)old * d;
      unsigned long long* sn = reinterpret_cast<unsigned long long*>(a.fx_sums) + (size_t)bi * d;
      if constexpr (D > 0) {
#pragma unroll
        for (int j = 0; j < D; ++j) {
          const long long q = __double2ll_rn((double)x[j] * a.fx_scale);
          atomicAdd(so + j, (unsigned long long)(-q));
          atomicAdd(sn + j, (unsigned long long)q);
        }
      } else {
        for (int j = 0; j < d; ++j) {
          const long long q = __double2ll_rn((double)__ldg(xp + j) * a.fx_scale);
          atomicAdd(so + j, (unsigned long long)(-q));
          atomicAdd(sn + j, (unsigned long long)q);
        }
      }
      atomicAdd(reinterpret_cast<unsigned long long*>(a.fx_counts) + old, (unsigned long long)(-1ll));
      atomicAdd(reinterpret_cast<unsigned long long*>(a.fx_counts) + bi, 1ull);
    }
  }
  if (a.n_exact != nullptr && my_exact) atomicAdd(a.n_exact, (unsigned long long)my_exact);
  if (a.n_changed != nullptr && my_changed) atomicAdd(a.n_changed, (unsigned long long)my_changed);
}

// ---- pivot pruning for listed rows (f32, d = 64 / 128) ------------------------------------------------------------------
// The rows the screening kernel lists carry a provisional label b (its best TF32 score).  Whatever b is, a centroid c with
// ||c - c_b|| > 2 ||x - c_b|| cannot be closer to x than c_b (triangle inequality), so the reference's scan is decided among
// the few centroids inside that ball — from a k-means++ start the two or three that share the row's blob.  With the
// roundings covered by the 2.001 / 1.0001 / 0.9999 factors below a pruned centroid is STRICTLY farther than c_b in the
// reference's own arithmetic ((d + 2) eps relative on a squared distance, three orders of magnitude below the slack), so it
// cannot win even on index order.  Rows with at most kPivMax survivors are decided here by the reference's scan (direct form,
// strict '<', index order) over the survivors; the others (overlapping clusters, NaN / Inf / denormal-range data: nothing is
// pruned) are passed on to rows_top2_kernel.  cfg3 from a k-means++ start, 211 000 listed rows per iteration: 116 us here plus
// 45 us for the 2 % passed on, against 603 us for the FMA pass over all k centroids.  HAM = true (Hamerly, the pivot is the
// row's current centroid) also writes the bounds and moves the row between the running sums, as rows_top2_kernel does.
constexpr int kPivMax = 8;
constexpr int kPivThreads = 128;

struct PivotArgs {
  const float* X;
  int k;
  const float* C;
  float* CC;                     // [k][k] lower bounds on the centroid-to-centroid distances (centroid_dist_kernel)
  const uint32_t* list;
  const unsigned int* count;
  uint32_t* labels;              // in: provisional label of a listed row; out: final label of the rows decided here
  uint32_t* list2;               // rows left for rows_top2_kernel
  unsigned int* count2;
  const LloydState* state;
  // Hamerly (HAM = true): bounds and the move between the running sums, as in RowsArgs
  float* ub;
  float* lb;
  long long* fx_sums;
  long long* fx_counts;
  double fx_scale;
  unsigned long long* n_exact;
  unsigned long long* n_changed;
};

// CTA b: lower bounds on ||c_b - c||, c = 0 .. k-1
template <int D>
__global__ void __launch_bounds__(128) centroid_dist_kernel(const PivotArgs a) {
  if (a.state != nullptr && a.state->done) return;
  if (*a.count == 0u) return;
  __shared__ float cb[D];
  const int b = blockIdx.x;
  for (int j = threadIdx.x; j < D; j += 128) cb[j] = a.C[(size_t)b * D + j];
  __syncthreads();
  for (int c = threadIdx.x; c < a.k; c += 128) {
    const float4* cp = reinterpret_cast<const float4*>(a.C + (size_t)c * D);
    float s = 0.f;
#pragma unroll 8
    for (int q = 0; q < D / 4; ++q) {
      const float4 v = __ldg(cp + q);
      const float d0 = v.x - cb[4 * q], d1 = v.y - cb[4 * q + 1], d2 = v.z - cb[4 * q + 2], d3 = v.w - cb[4 * q + 3];
      s = fmaf(d0, d0, fmaf(d1, d1, fmaf(d2, d2, fmaf(d3, d3, s))));
    }
    a.CC[(size_t)b * a.k + c] = sqrtf(s) * 0.9999f;
  }
}

template <int D, bool HAM>
__global__ void __launch_bounds__(kPivThreads) rows_pivot_kernel(const PivotArgs a) {
  if (a.state != nullptr && a.state->done) return;
  __shared__ uint16_t surv[kPivThreads][kPivMax];
  const int t = threadIdx.x, lane = t & 31;
  const unsigned int total = *a.count;
  const unsigned int n_tiles = (total + kPivThreads - 1) / kPivThreads;
  const unsigned lt = (1u << lane) - 1u;
  unsigned int my_exact = 0, my_changed = 0;
  for (unsigned int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned int idx = tile * kPivThreads + t;
    const bool valid = idx < total;
    const uint32_t row = valid ? a.list[idx] : 0u;
    float x[D];
    rows_load<float, D>(a.X + (size_t)row * D, x);
    const uint32_t old = valid ? a.labels[row] : 0u;
    const uint32_t b = old < (uint32_t)a.k ? old : 0u;
    const float bd = sqdist_reg<float, D>(a.C + (size_t)b * D, x);
    // nothing is pruned (threshold +inf) when the distance is not a normal number of moderate size
    const bool normal = valid && bd > 1e-30f && bd < 1e30f;
    const float ub_b = sqrtf(bd) * 1.0001f;   // >= the true distance to the pivot
    const float thr = normal ? 2.001f * ub_b : INFINITY;
    int my_cnt = 0;
    float my_ccmin = INFINITY;   // HAM: the nearest pruned centroid to the pivot
    for (int r = 0; r < 32; ++r) {
      const uint32_t b_r = __shfl_sync(0xffffffffu, b, r);
      const float thr_r = __shfl_sync(0xffffffffu, thr, r);
      const float* ccr = a.CC + (size_t)b_r * a.k;
      int cnt_r = 0;
      float ccmin = INFINITY;
      for (int c0 = 0; c0 < a.k; c0 += 32) {
        const int c = c0 + lane;
        const float cc = c < a.k ? ccr[c] : INFINITY;
        const bool keep = c < a.k && !(cc > thr_r);   // NaN distances are kept
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (keep) {
          const int pos = cnt_r + __popc(m & lt);
          if (pos < kPivMax) surv[(t & ~31) + r][pos] = (uint16_t)c;
        } else if (HAM) {
          ccmin = fminf(ccmin, cc);
        }
        cnt_r += __popc(m);
      }
      if (HAM) {
        for (int o = 16; o > 0; o >>= 1) ccmin = fminf(ccmin, __shfl_xor_sync(0xffffffffu, ccmin, o));
        if (lane == r) my_ccmin = ccmin;
      }
      if (lane == r) my_cnt = cnt_r;
    }
    __syncwarp();
    const bool open = valid && my_cnt > kPivMax;
    if (valid && !open) {
      float eb = INFINITY, es = INFINITY;
      int ei = 0;
      for (int i = 0; i < my_cnt; ++i) {
        const int c = surv[t][i];
        const float v = sqdist_reg<float, D>(a.C + (size_t)c * D, x);
        if (c == 0) { eb = v; ei = 0; }   // centroid 0 seeds the reference's scan (a NaN there is sticky)
        else if (v < eb) { es = eb; eb = v; ei = c; }
        else if (v < es) es = v;
      }
      a.labels[row] = (uint32_t)ei;
      ++my_exact;
      if (HAM) {
        // bounds on the TRUE distances: the direct form is within (d + 2) eps of them.  The runner-up is a survivor (exact
        // distance es) or a pruned centroid, which is at least CC[b][c] - ||x - c_b|| away.
        const float slack = ((float)D + 4.f) * rows_eps<float>();
        a.ub[row] = (float)sqrt((double)eb) * (1.f + slack);
        const float lsq = es * (1.f - 2.f * slack);
        const float lb_surv = lsq > 0.f ? (float)sqrt((double)lsq) * (1.f - slack) : 0.f;
        const float lb_pruned = (my_ccmin - ub_b) * 0.9999f;
        a.lb[row] = fmaxf(fminf(lb_surv, lb_pruned), 0.f);
        if (old != (uint32_t)ei) {
          ++my_changed;
          unsigned long long* so = reinterpret_cast<unsigned long long*>(a.fx_sums) + (size_t)old * D;
          unsigned long long* sn = reinterpret_cast<unsigned long long*>(a.fx_sums) + (size_t)ei * D;
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const long long q = __double2ll_rn((double)x[j] * a.fx_scale);
            atomicAdd(so + j, (unsigned long long)(-q));
            atomicAdd(sn + j, (unsigned long long)q);
          }
          atomicAdd(reinterpret_cast<unsigned long long*>(a.fx_counts) + old, (unsigned long long)(-1ll));
          atomicAdd(reinterpret_cast<unsigned long long*>(a.fx_counts) + ei, 1ull);
        }
      }
    }
    const unsigned fm = __ballot_sync(0xffffffffu, open);
    if (fm != 0u) {
      unsigned int base = 0;
      if (lane == 0) base = atomicAdd(a.count2, (unsigned int)__popc(fm));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (open) a.list2[base + __popc(fm & lt)] = row;
    }
    __syncwarp();
  }
  if (HAM) {
    if (a.n_exact != nullptr && my_exact) atomicAdd(a.n_exact, (unsigned long long)my_exact);
    if (a.n_changed != nullptr && my_changed) atomicAdd(a.n_changed, (unsigned long long)my_changed);
  }
}

}  // namespace lkm

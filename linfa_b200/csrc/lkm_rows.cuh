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
    F b0 = (F)INFINITY, s0 = (F)INFINITY, b1 = (F)INFINITY, s1 = (F)INFINITY, cm = (F)0;
    int i0 = 0, i1 = 1;
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
        top2_update<F>(fma((F)-2, d0, n0), cb + c, b0, s0, i0);
        top2_update<F>(fma((F)-2, d1, n1), cb + c + 1, b1, s1, i1);
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
        top2_update<F>(fma((F)-2, d0, n0), cb + c, b0, s0, i0);
      }
    }
    F best, second;
    int bi;
    if (b1 < b0 || (b1 == b0 && i1 < i0)) { best = b1; bi = i1; second = (b0 < s1) ? b0 : s1; }
    else { best = b0; bi = i0; second = (b1 < s0) ? b1 : s0; }
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
    const F dsum = fmax(best + xx, (F)0) + fmax(second + xx, (F)0) + (F)2 * errdot + (F)2 * (dd + (F)1) * eps * xx;
    const F margin = (F)1.1 * ((F)2 * errdot + (dd + (F)2) * eps * dsum) + ((F)8 * dd + (F)16) * dmin * ((F)1 + xn + cmax);
    const bool cert = (second - best) > margin;   // NaN / Inf anywhere -> false
    const bool need_exact = valid && !cert;
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
    if (!need_exact) {
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
      F lsq;
      if (need_exact) lsq = es * ((F)1 - (F)2 * slack);
      else lsq = (second + xx) - (F)1.1 * eps * (((F)2 * dd + (F)2) * xn * cmax + (F)2 * cmax * cmax + (dd + (F)2) * xx) -
                 ((F)8 * dd + (F)16) * dmin * ((F)1 + xn + cmax);
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

}  // namespace lkm

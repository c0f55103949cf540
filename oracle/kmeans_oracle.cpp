// ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product.
//
// CPU restatement (C++17, g++ -O2 -ffp-contract=off, no fast-math) of the
// K-Means hot path of rust-ml/linfa v0.8.1 (linfa-clustering).  Only tests/,
// __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this library, and only as the checker or the timed CPU baseline.
// The product (linfa_b200/liblkm.so) never links, loads or calls it.
//
// Parity status: the reference is Rust and cannot be built here (no cargo /
// rustc, dependencies not vendored), so this oracle is pinned against the
// reference's OWN test literals (tests/test_oracle_kats.py replays them, see
// SURVEY.md Appendix B) and against the published xoshiro256+ / SplitMix64
// vectors.  Two things the reference's tests do NOT pin are restated from the
// crates.io sources of the pinned dependency versions and are therefore
// "parity unpinned":
//   * floating-point summation ORDER (ndarray 0.16 `sum()`, ndarray-stats 0.6
//     `sq_l2_dist`, the row-order accumulation in `compute_centroids`);
//   * the RNG STREAM consumed by rand 0.8.5 (`WeightedIndex`, `Uniform`,
//     `gen_range`, `seq::index::sample`) on top of rand_xoshiro 0.6.
//
// Every function cites the reference file:line it follows.  Paths are relative
// to /root/reference/ ; "KM" = algorithms/linfa-clustering/src/k_means.

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <thread>
#include <unordered_set>
#include <vector>

namespace {

// ---------------------------------------------------------------------------
// RNG: rand_xoshiro 0.6 Xoshiro256Plus + SplitMix64 seeding.
// Default rng of KMeans::params is Xoshiro256Plus::seed_from_u64(42)
// (KM/algorithm.rs:211).
// ---------------------------------------------------------------------------
struct Xoshiro256Plus {
  uint64_t s[4];
  static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next_u64() {
    const uint64_t r = s[0] + s[3];
    const uint64_t t = s[1] << 17;
    s[2] ^= s[0];
    s[3] ^= s[1];
    s[1] ^= s[2];
    s[0] ^= s[3];
    s[2] ^= t;
    s[3] = rotl(s[3], 45);
    return r;
  }
  uint32_t next_u32() { return (uint32_t)(next_u64() >> 32); }
  static Xoshiro256Plus seed_from_u64(uint64_t x) {
    Xoshiro256Plus g;
    for (int i = 0; i < 4; ++i) {
      x += 0x9e3779b97f4a7c15ULL;
      uint64_t z = x;
      z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ULL;
      z = (z ^ (z >> 27)) * 0x94d049bb133111ebULL;
      g.s[i] = z ^ (z >> 31);
    }
    if ((g.s[0] | g.s[1] | g.s[2] | g.s[3]) == 0) return seed_from_u64(0);
    return g;
  }
};

// ---------------------------------------------------------------------------
// rand 0.8.5 distributions, restated.
// ---------------------------------------------------------------------------
inline void wmul64(uint64_t a, uint64_t b, uint64_t& hi, uint64_t& lo) {
  unsigned __int128 p = (unsigned __int128)a * b;
  hi = (uint64_t)(p >> 64);
  lo = (uint64_t)p;
}
inline int clz32(uint32_t x) { return x ? __builtin_clz(x) : 32; }
inline int clz64(uint64_t x) { return x ? __builtin_clzll(x) : 64; }

// UniformInt::<u64|usize>::sample_single_inclusive (rand 0.8.5 uniform.rs)
uint64_t gen_range_incl_u64(Xoshiro256Plus& rng, uint64_t low, uint64_t high) {
  const uint64_t range = high - low + 1;  // wrapping
  if (range == 0) return rng.next_u64();
  const uint64_t zone = (range << clz64(range)) - 1;
  for (;;) {
    const uint64_t v = rng.next_u64();
    uint64_t hi, lo;
    wmul64(v, range, hi, lo);
    if (lo <= zone) return low + hi;
  }
}
// UniformInt::<u32>::sample_single_inclusive
uint32_t gen_range_incl_u32(Xoshiro256Plus& rng, uint32_t low, uint32_t high) {
  const uint32_t range = high - low + 1;
  if (range == 0) return rng.next_u32();
  const uint32_t zone = (range << clz32(range)) - 1;
  for (;;) {
    const uint32_t v = rng.next_u32();
    const uint64_t p = (uint64_t)v * range;
    const uint32_t hi = (uint32_t)(p >> 32), lo = (uint32_t)p;
    if (lo <= zone) return low + hi;
  }
}
// Uniform::<u32>::new(0, length).sample — the *distribution object* path used
// by index::sample_rejection; zone is computed with a modulus here.
struct UniformU32 {
  uint32_t low, range, z;
  UniformU32(uint32_t lo_, uint32_t hi_excl) {
    low = lo_;
    const uint32_t high = hi_excl - 1;
    range = high - low + 1;
    z = range > 0 ? (uint32_t)((0xFFFFFFFFu - range + 1) % range) : 0;
  }
  uint32_t sample(Xoshiro256Plus& rng) const {
    if (range == 0) return rng.next_u32();
    const uint32_t zone = 0xFFFFFFFFu - z;
    for (;;) {
      const uint32_t v = rng.next_u32();
      const uint64_t p = (uint64_t)v * range;
      if ((uint32_t)p <= zone) return low + (uint32_t)(p >> 32);
    }
  }
};
struct UniformU64 {
  uint64_t low, range, z;
  UniformU64(uint64_t lo_, uint64_t hi_excl) {
    low = lo_;
    const uint64_t high = hi_excl - 1;
    range = high - low + 1;
    z = range > 0 ? (uint64_t)((0xFFFFFFFFFFFFFFFFULL - range + 1) % range) : 0;
  }
  uint64_t sample(Xoshiro256Plus& rng) const {
    if (range == 0) return rng.next_u64();
    const uint64_t zone = 0xFFFFFFFFFFFFFFFFULL - z;
    for (;;) {
      uint64_t hi, lo;
      wmul64(rng.next_u64(), range, hi, lo);
      if (lo <= zone) return low + hi;
    }
  }
};

// value in [0,1) exactly as UniformFloat does it (rand 0.8.5 uniform.rs):
// f64: (next_u64 >> 12) into mantissa of [1,2) minus 1; f32: (next_u32 >> 9).
inline double u01_f64(Xoshiro256Plus& rng) {
  const uint64_t bits = (rng.next_u64() >> 12) | (0x3FFULL << 52);
  double v;
  std::memcpy(&v, &bits, 8);
  return v - 1.0;
}
inline float u01_f32(Xoshiro256Plus& rng) {
  const uint32_t bits = (rng.next_u32() >> 9) | (0x7Fu << 23);
  float v;
  std::memcpy(&v, &bits, 4);
  return v - 1.0f;
}
template <class F> struct FloatTraits;
template <> struct FloatTraits<float> {
  static float u01(Xoshiro256Plus& r) { return u01_f32(r); }
  static float max_rand() { return 1.0f - 1.1920929e-7f; }  // (u32::MAX>>9) in [1,2) minus 1
};
template <> struct FloatTraits<double> {
  static double u01(Xoshiro256Plus& r) { return u01_f64(r); }
  static double max_rand() { return 1.0 - 2.220446049250313e-16; }
};

// UniformFloat::new(low, high) + sample  (used by WeightedIndex with low = 0)
template <class F> struct UniformFloat {
  F low, scale;
  UniformFloat(F lo, F hi) {
    low = lo;
    scale = hi - lo;
    const F max_rand = FloatTraits<F>::max_rand();
    for (;;) {
      volatile F t = scale * max_rand;  // mul then add, never fused
      if (!(t + low >= hi)) break;
      scale = std::nextafter(scale, (F)0);  // decr(): one ulp towards zero
    }
  }
  F sample(Xoshiro256Plus& rng) const {
    const F v01 = FloatTraits<F>::u01(rng);
    volatile F t = v01 * scale;
    return t + low;
  }
};

// Rng::gen_range(0.0..1.0) for f64: UniformFloat::sample_single (KM/init.rs:263)
inline double gen_range_unit_f64(Xoshiro256Plus& rng) {
  for (;;) {
    const double res = u01_f64(rng) * 1.0 + 0.0;
    if (res < 1.0) return res;
  }
}

// WeightedIndex::new(weights).sample(rng)  (rand 0.8.5 weighted_index.rs),
// SURVEY A.6.  Returns -1 on any WeightedError (the caller maps that to the
// reference's `.expect(..)` panic or `.unwrap_or(0)`).
template <class F>
int64_t weighted_index_sample(const F* w, uint64_t n, Xoshiro256Plus& rng, std::vector<F>& cum) {
  if (n == 0) return -1;  // NoItem
  F total = w[0];
  if (!(total >= (F)0)) return -1;  // InvalidWeight
  cum.clear();
  cum.reserve(n ? n - 1 : 0);
  for (uint64_t i = 1; i < n; ++i) {
    if (!(w[i] >= (F)0)) return -1;
    cum.push_back(total);
    total += w[i];  // sequential, in F
  }
  if (total == (F)0) return -1;  // AllWeightsZero
  UniformFloat<F> distr((F)0, total);
  const F chosen = distr.sample(rng);
  // binary_search_by(|w| if *w <= chosen {Less} else {Greater}).unwrap_err()
  // == number of leading entries with cum[j] <= chosen (cum is non-decreasing).
  uint64_t lo = 0, hi = cum.size();
  while (lo < hi) {
    const uint64_t mid = lo + (hi - lo) / 2;
    if (cum[mid] <= chosen) lo = mid + 1; else hi = mid;
  }
  return (int64_t)lo;
}

// rand::seq::index::sample(rng, length, amount)  (rand 0.8.5 seq/index.rs),
// SURVEY A.7 — constants restated from memory, low confidence; only
// KMeansInit::Random depends on them (KM/init.rs:111).
std::vector<uint64_t> index_sample(Xoshiro256Plus& rng, uint64_t length, uint64_t amount, bool* ok) {
  std::vector<uint64_t> out;
  *ok = true;
  if (amount > length) { *ok = false; return out; }  // reference panics
  if (length > 0xFFFFFFFFULL) {  // sample_rejection::<usize>
    UniformU64 distr(0, length);
    std::unordered_set<uint64_t> cache;
    for (uint64_t i = 0; i < amount; ++i) {
      uint64_t pos = distr.sample(rng);
      while (!cache.insert(pos).second) pos = distr.sample(rng);
      out.push_back(pos);
    }
    return out;
  }
  const uint32_t len = (uint32_t)length, amt = (uint32_t)amount;
  auto inplace = [&]() {
    std::vector<uint32_t> idx(len);
    for (uint32_t i = 0; i < len; ++i) idx[i] = i;
    for (uint32_t i = 0; i < amt; ++i) {
      const uint32_t j = gen_range_incl_u32(rng, i, len - 1);
      std::swap(idx[i], idx[j]);
    }
    for (uint32_t i = 0; i < amt; ++i) out.push_back(idx[i]);
  };
  auto floyd = [&]() {
    const bool floyd_shuffle = amt < 50;
    std::vector<uint32_t> idx;
    idx.reserve(amt);
    for (uint32_t j = len - amt; j < len; ++j) {
      const uint32_t t = gen_range_incl_u32(rng, 0, j);
      auto it = std::find(idx.begin(), idx.end(), t);
      if (floyd_shuffle) {
        if (it != idx.end()) { idx.insert(it, j); continue; }
      } else if (it != idx.end()) {
        idx.push_back(j);
        continue;
      }
      idx.push_back(t);
    }
    if (!floyd_shuffle) {
      for (uint32_t i = amt - 1; i >= 1; --i) {
        const uint32_t r = gen_range_incl_u32(rng, 0, i);
        std::swap(idx[i], idx[r]);
      }
    }
    for (uint32_t v : idx) out.push_back(v);
  };
  auto rejection = [&]() {
    UniformU32 distr(0, len);
    std::unordered_set<uint32_t> cache;
    for (uint32_t i = 0; i < amt; ++i) {
      uint32_t pos = distr.sample(rng);
      while (!cache.insert(pos).second) pos = distr.sample(rng);
      out.push_back(pos);
    }
  };
  const int j = len < 500000u ? 0 : 1;
  if (amt < 163) {
    const float C[2][2] = {{1.6f, 8.0f / 45.0f}, {10.0f, 70.0f / 9.0f}};
    const float amount_fp = (float)amt;
    const float m4 = C[0][j] * amount_fp;
    if (amt > 11 && (float)len < (C[1][j] + m4) * amount_fp) inplace(); else floyd();
  } else {
    const float C[2] = {270.0f, 330.0f / 9.0f};
    if ((float)len < C[j] * (float)amt) inplace(); else rejection();
  }
  return out;
}

// ---------------------------------------------------------------------------
// Arithmetic kernels of the path
// ---------------------------------------------------------------------------

// ndarray-stats 0.6 DeviationExt::sq_l2_dist as called by L2Dist::rdistance
// (algorithms/linfa-nn/src/distance.rs:68-70): acc += (a_j-b_j)*(a_j-b_j) in F,
// feature order, never fused.  a = centroid, b = observation
// (KM/algorithm.rs:935,939).
template <class F> inline F sq_l2_dist(const F* a, const F* b, uint64_t d) {
  F acc = (F)0;
  for (uint64_t j = 0; j < d; ++j) {
    const F diff = a[j] - b[j];
    const F sq = diff * diff;
    acc = acc + sq;
  }
  return acc;
}

// closest_centroid (KM/algorithm.rs:923-946): strict '<', centroid 0 evaluated
// as the seed and again inside the loop; NaN never wins.
template <class F>
inline void closest_centroid(const F* C, uint64_t k, uint64_t d, const F* x, uint64_t* idx, F* dist) {
  uint64_t best = 0;
  F bd = sq_l2_dist(C, x, d);
  for (uint64_t c = 0; c < k; ++c) {
    const F v = sq_l2_dist(C + c * d, x, d);
    if (v < bd) { best = c; bd = v; }
  }
  *idx = best;
  *dist = bd;
}

// ndarray 0.16 ArrayBase::sum() on a contiguous array = numeric_util::
// unrolled_fold.  Used for inertia (KM/algorithm.rs:329), the
// k-means|| cost (KM/init.rs:245) and weights.sum() (KM/init.rs:127).
template <class F> F ndarray_sum(const F* x, uint64_t n) {
  F p[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  uint64_t i = 0;
  for (; i + 8 <= n; i += 8)
    for (int l = 0; l < 8; ++l) p[l] = p[l] + x[i + l];
  F acc = (F)0;
  acc = acc + (p[0] + p[4]);
  acc = acc + (p[1] + p[5]);
  acc = acc + (p[2] + p[6]);
  acc = acc + (p[3] + p[7]);
  for (; i < n; ++i) acc = acc + x[i];
  return acc;
}

// L2Dist::distance on the two (k x d) centroid matrices
// (KM/algorithm.rs:323-325, linfa-nn distance.rs:63-65): sq_l2_dist in F over
// all k*d elements in logical order, sqrt taken in f64, cast back to F.
template <class F> F l2_distance(const F* a, const F* b, uint64_t len) {
  const F sq = sq_l2_dist(a, b, len);
  return (F)std::sqrt((double)sq);
}

// parallel row loop standing in for Zip::par_for_each
// (KM/algorithm.rs:844-880): rows are independent, so the thread partition
// does not change any value.
template <class Fn> void par_rows(uint64_t n, int n_threads, Fn fn) {
  if (n_threads <= 1 || n < 4096) { fn((uint64_t)0, n); return; }
  std::vector<std::thread> th;
  const uint64_t chunk = (n + n_threads - 1) / n_threads;
  for (int t = 0; t < n_threads; ++t) {
    const uint64_t lo = std::min(n, (uint64_t)t * chunk), hi = std::min(n, lo + chunk);
    if (lo < hi) th.emplace_back([=]() { fn(lo, hi); });
  }
  for (auto& t : th) t.join();
}

// update_memberships_and_dists / update_cluster_memberships / update_min_dists
// (KM/algorithm.rs:838-881).  labels or dists may be null.
template <class F>
void assign(const F* X, uint64_t n, uint64_t d, const F* C, uint64_t k, uint64_t* labels, F* dists, int n_threads) {
  par_rows(n, n_threads, [=](uint64_t lo, uint64_t hi) {
    for (uint64_t i = lo; i < hi; ++i) {
      uint64_t l;
      F dd;
      closest_centroid(C, k, d, X + i * d, &l, &dd);
      if (labels) labels[i] = l;
      if (dists) dists[i] = dd;
    }
  });
}

// compute_centroids (KM/algorithm.rs:785-810): counts start at 1, rows added
// sequentially in row order, old centroid added as an extra member, true
// division.  Acc = F is the faithful version; Acc = double is the
// "f64-accumulated oracle" BASELINE.md section 4 asks parity to be stated against.
template <class F, class Acc>
void compute_centroids(const F* C_old, const F* X, const uint64_t* labels, uint64_t n, uint64_t d, uint64_t k, F* C_new) {
  std::vector<uint64_t> counts(k, 1);
  std::vector<Acc> S(k * d, (Acc)0);
  for (uint64_t i = 0; i < n; ++i) {
    Acc* row = S.data() + labels[i] * d;
    const F* x = X + i * d;
    for (uint64_t j = 0; j < d; ++j) row[j] = row[j] + (Acc)x[j];
    counts[labels[i]] += 1;
  }
  for (uint64_t c = 0; c < k; ++c)
    for (uint64_t j = 0; j < d; ++j) {
      const Acc s = S[c * d + j] + (Acc)C_old[c * d + j];
      C_new[c * d + j] = (F)(s / (Acc)counts[c]);
    }
}

// compute_centroids_incremental (KM/algorithm.rs:815-835), mini-batch update.
template <class F>
void compute_centroids_incremental(const F* X, const uint64_t* labels, uint64_t n, uint64_t d, uint64_t k,
                                   const F* C_old, F* counts, F* C_new) {
  std::memcpy(C_new, C_old, sizeof(F) * k * d);
  for (uint64_t i = 0; i < n; ++i) {
    const uint64_t c = labels[i];
    counts[c] = counts[c] + (F)1;
    F* cen = C_new + c * d;
    const F* x = X + i * d;
    for (uint64_t j = 0; j < d; ++j) {
      const F shift = (x[j] - cen[j]) / counts[c];
      cen[j] = cen[j] + shift;
    }
  }
}

// ---------------------------------------------------------------------------
// Seeding (KM/init.rs)
// ---------------------------------------------------------------------------
enum InitMethod { INIT_RANDOM = 0, INIT_PRECOMPUTED = 1, INIT_KMEANSPP = 2, INIT_KMEANSPARA = 3 };

// random_init (KM/init.rs:105-113)
template <class F>
int random_init(uint64_t k, const F* X, uint64_t n, uint64_t d, Xoshiro256Plus& rng, F* C, uint64_t* picked) {
  bool ok;
  std::vector<uint64_t> idx = index_sample(rng, n, k, &ok);
  if (!ok) return 3;  // reference panics: amount > length
  for (uint64_t c = 0; c < k; ++c) {
    std::memcpy(C + c * d, X + idx[c] * d, sizeof(F) * d);
    if (picked) picked[c] = idx[c];
  }
  return 0;
}

// weighted_k_means_plusplus (KM/init.rs:118-158)
template <class F>
int weighted_kmeanspp(uint64_t k, const F* X, uint64_t n, uint64_t d, const F* weights, Xoshiro256Plus& rng,
                      F* C, uint64_t* picked, int n_threads) {
  if (ndarray_sum(weights, n) == (F)0) return 4;  // assert_ne!(weights.sum(), 0)
  std::vector<F> cum;
  const int64_t first = weighted_index_sample(weights, n, rng, cum);
  if (first < 0) return 5;  // .expect("invalid weights")
  std::memcpy(C, X + (uint64_t)first * d, sizeof(F) * d);
  if (picked) picked[0] = (uint64_t)first;
  std::vector<F> dists(n, (F)0);
  for (uint64_t c_cnt = 1; c_cnt < k; ++c_cnt) {
    assign<F>(X, n, d, C, c_cnt, nullptr, dists.data(), n_threads);  // vs ALL chosen so far (:138-143)
    for (uint64_t i = 0; i < n; ++i) dists[i] = dists[i] * weights[i];  // :147
    int64_t idx = weighted_index_sample(dists.data(), n, rng, cum);
    if (idx < 0) idx = 0;  // .unwrap_or(0) (:149-152)
    std::memcpy(C + c_cnt * d, X + (uint64_t)idx * d, sizeof(F) * d);
    if (picked) picked[c_cnt] = (uint64_t)idx;
  }
  return 0;
}

// sample_subsequent_candidates (KM/init.rs:239-270).  The reference seeds one
// Xoshiro256Plus per rayon work split from an atomic counter, so which RNG a
// row sees depends on rayon's dynamic splitting (non-deterministic; the
// reference's own tests say so, KM/algorithm.rs:1309-1312).  `rayon_threads`
// picks one legal schedule: rayon's adaptive splitter without work stealing
// (splits = T; at every node with len/2 >= 1 and splits > 0: splits /= 2 and
// cut at len/2), leaves visited left to right, leaf s seeded with seed+s.
// T = 1 is what RAYON_NUM_THREADS=1 does (two halves); T = 0 means no split.
inline void rayon_leaves(uint64_t lo, uint64_t hi, uint64_t splits,
                         std::vector<std::pair<uint64_t, uint64_t>>& leaves) {
  const uint64_t len = hi - lo;
  if (len / 2 >= 1 && splits > 0) {
    const uint64_t mid = lo + len / 2;
    rayon_leaves(lo, mid, splits / 2, leaves);
    rayon_leaves(mid, hi, splits / 2, leaves);
  } else {
    leaves.push_back({lo, hi});
  }
}
template <class F>
void sample_subsequent_candidates(const F* dists, uint64_t n, F multiplier, uint64_t seed, int rayon_threads,
                                  std::vector<uint64_t>& out) {
  const F cost = ndarray_sum(dists, n);
  out.clear();
  std::vector<std::pair<uint64_t, uint64_t>> pieces;
  rayon_leaves(0, n, (uint64_t)std::max(0, rayon_threads), pieces);
  uint64_t s = seed;
  for (auto& p : pieces) {
    if (p.first == p.second) continue;
    Xoshiro256Plus rng = Xoshiro256Plus::seed_from_u64(s++);
    for (uint64_t i = p.first; i < p.second; ++i) {
      const F rnd = (F)gen_range_unit_f64(rng);
      const F prob = multiplier * dists[i] / cost;
      if (rnd < prob) out.push_back(i);
    }
  }
}

// cluster_membership_counts (KM/init.rs:273-285)
template <class F>
void cluster_membership_counts(const F* C, uint64_t k, const F* X, uint64_t n, uint64_t d, F* counts, int n_threads) {
  std::vector<uint64_t> lab(n);
  assign<F>(X, n, d, C, k, lab.data(), nullptr, n_threads);
  for (uint64_t c = 0; c < k; ++c) counts[c] = (F)0;
  for (uint64_t i = 0; i < n; ++i) counts[lab[i]] = counts[lab[i]] + (F)1;
}

// k_means_para (KM/init.rs:181-234)
template <class F>
int kmeans_para(uint64_t k, const F* X, uint64_t n, uint64_t d, Xoshiro256Plus& rng, F* C, int n_threads, int n_splits) {
  const uint64_t n_rounds = 8, cap = k * n_rounds;
  std::vector<F> cand(cap * d, (F)0);
  const uint64_t first = gen_range_incl_u64(rng, 0, n - 1);  // rng.gen_range(0..n_samples)
  std::memcpy(cand.data(), X + first * d, sizeof(F) * d);
  uint64_t n_cand = 1;
  std::vector<F> dists(n, (F)0);
  std::vector<uint64_t> next;
  bool full = false;
  for (uint64_t r = 0; r < n_rounds && !full; ++r) {
    assign<F>(X, n, d, cand.data(), n_cand, nullptr, dists.data(), n_threads);
    const uint64_t seed = gen_range_incl_u64(rng, 0, 0xFFFFFFFFFFFFFFFEULL);  // gen_range(0..u64::MAX)
    sample_subsequent_candidates<F>(dists.data(), n, (F)k, seed, n_splits, next);
    for (uint64_t idx : next) {
      std::memcpy(cand.data() + n_cand * d, X + idx * d, sizeof(F) * d);
      n_cand += 1;
      if (n_cand >= cap) { full = true; break; }
    }
  }
  std::vector<F> w(n_cand);
  cluster_membership_counts<F>(cand.data(), n_cand, X, n, d, w.data(), n_threads);
  return weighted_kmeanspp<F>(k, cand.data(), n_cand, d, w.data(), rng, C, nullptr, n_threads);
}

// KMeansInit::run (KM/init.rs:83-101)
template <class F>
int run_init(int method, uint64_t k, const F* X, uint64_t n, uint64_t d, const F* precomputed, Xoshiro256Plus& rng,
             F* C, int n_threads, int n_splits) {
  switch (method) {
    case INIT_RANDOM: return random_init<F>(k, X, n, d, rng, C, nullptr);
    case INIT_PRECOMPUTED: std::memcpy(C, precomputed, sizeof(F) * k * d); return 0;
    case INIT_KMEANSPP: {
      std::vector<F> ones(n, (F)1);
      return weighted_kmeanspp<F>(k, X, n, d, ones.data(), rng, C, nullptr, n_threads);
    }
    case INIT_KMEANSPARA: return kmeans_para<F>(k, X, n, d, rng, C, n_threads, n_splits);
  }
  return 2;
}

// fit_lloyd + get_kmeans_result (KM/algorithm.rs:294-367), SURVEY A.1/A.10.
// status: 0 ok, 1 InertiaError, 10..13 the four KMeansParamsError (hyperparams.rs:124-136).
template <class F, class Acc>
int fit_lloyd(const F* X, uint64_t n, uint64_t d, uint64_t k, uint64_t n_runs, uint64_t max_iter, F tol, int method,
              const F* precomputed, Xoshiro256Plus& rng_in, F* C_out, F* counts_out, F* inertia_out,
              uint64_t* n_iter_out, uint64_t* labels_out, int n_threads, int n_splits) {
  if (k == 0) return 10;
  if (n_runs == 0) return 11;
  if (tol <= (F)0) return 12;  // `tolerance <= 0` is the error; NaN passes, as in the reference
  if (max_iter == 0) return 13;
  Xoshiro256Plus rng = rng_in;  // self.rng().clone() (:298)
  std::vector<uint64_t> memberships(n, 0);
  std::vector<F> dists(n, (F)0);
  std::vector<F> C(k * d), Cn(k * d), best(k * d);
  F min_inertia = std::numeric_limits<F>::infinity();
  bool have_best = false;
  uint64_t last_iters = 0;
  for (uint64_t run = 0; run < n_runs; ++run) {
    const int st = run_init<F>(method, k, X, n, d, precomputed, rng, C.data(), n_threads, n_splits);
    if (st) return st;
    uint64_t n_iter = 0;
    F inertia;
    for (;;) {
      assign<F>(X, n, d, C.data(), k, memberships.data(), dists.data(), n_threads);
      compute_centroids<F, Acc>(C.data(), X, memberships.data(), n, d, k, Cn.data());
      const F shift = l2_distance<F>(C.data(), Cn.data(), k * d);
      C.swap(Cn);
      n_iter += 1;
      if (shift < tol || n_iter == max_iter) {
        if (sizeof(Acc) == sizeof(F)) {
          inertia = ndarray_sum<F>(dists.data(), n);
        } else {
          double s = 0;
          for (uint64_t i = 0; i < n; ++i) s += (double)dists[i];
          inertia = (F)s;
        }
        break;
      }
    }
    last_iters = n_iter;
    if (inertia < min_inertia) {
      min_inertia = inertia;
      best = C;
      have_best = true;
    }
  }
  if (n_iter_out) *n_iter_out = last_iters;
  if (labels_out) std::memcpy(labels_out, memberships.data(), sizeof(uint64_t) * n);
  if (!have_best) return 1;
  for (uint64_t c = 0; c < k; ++c) counts_out[c] = (F)0;
  for (uint64_t i = 0; i < n; ++i) counts_out[memberships[i]] = counts_out[memberships[i]] + (F)1;
  std::memcpy(C_out, best.data(), sizeof(F) * k * d);
  *inertia_out = min_inertia / (F)n;
  return 0;
}

// ---------------------------------------------------------------------------
// Hamerly (KM/algorithm.rs:248-291, 396-603, 889-919)
// ---------------------------------------------------------------------------

// L2Dist::distance on two rows (linfa-nn distance.rs:63-65): sq_l2_dist in F, sqrt in f64, cast to F
template <class F> inline F l2_dist_row(const F* a, const F* b, uint64_t d) {
  return (F)std::sqrt((double)sq_l2_dist<F>(a, b, d));
}

// two_closest_centroids (KM/algorithm.rs:889-919): argument order distance(observation, centroid); centroids 0 and 1
// seed the scan (a tie between them goes to 1), then `closest <= dist < second` / `dist < closest`
template <class F>
inline void two_closest_centroids(const F* C, uint64_t k, uint64_t d, const F* x, uint64_t* idx, F* closest, F* second) {
  if (k == 1) { *idx = 0; *closest = (F)0; *second = (F)0; return; }
  const F dist1 = l2_dist_row<F>(x, C, d);
  const F dist2 = l2_dist_row<F>(x, C + d, d);
  uint64_t ci = dist1 < dist2 ? 0 : 1;
  F cd = dist1 < dist2 ? dist1 : dist2;
  F sd = dist1 < dist2 ? dist2 : dist1;
  for (uint64_t c = 2; c < k; ++c) {
    const F v = l2_dist_row<F>(x, C + c * d, d);
    if (cd <= v && v < sd) sd = v;
    else if (v < cd) { sd = cd; ci = c; cd = v; }
  }
  *idx = ci; *closest = cd; *second = sd;
}

// two_farthest_indices (KM/algorithm.rs:585-603)
template <class F> inline void two_farthest_indices(const F* m, uint64_t k, uint64_t* f1, uint64_t* f2) {
  if (k < 2) { *f1 = 0; *f2 = 0; return; }
  uint64_t a, b;
  if (m[1] >= m[0]) { a = 1; b = 0; } else { a = 0; b = 1; }
  for (uint64_t i = 2; i < k; ++i) {
    if (m[i] >= m[a]) { b = a; a = i; }
    else if (m[i] > m[b]) b = i;
  }
  *f1 = a; *f2 = b;
}

// compute_inertia (KM/algorithm.rs:608-620): fold in F, row order
template <class F, class Acc>
F compute_inertia(const F* X, uint64_t n, uint64_t d, const uint64_t* labels, const F* C) {
  Acc acc = (Acc)0;
  for (uint64_t i = 0; i < n; ++i) acc = acc + (Acc)sq_l2_dist<F>(X + i * d, C + labels[i] * d, d);
  return (F)acc;
}

// HamerlyAlgorithm (KM/algorithm.rs:396-583).  Acc = F is the reference's arithmetic (running sums in F, moved by
// += / -= per changed row, :507-523); Acc = double keeps the running sums in f64 (the accumulation rule the GPU
// path is compared against).
template <class F, class Acc> struct Hamerly {
  const F* X; uint64_t n, d, k;
  std::vector<F> C;
  std::vector<uint64_t> memberships, prev;
  std::vector<F> upper, lower;
  std::vector<uint64_t> counts;
  std::vector<Acc> sums;
  int n_threads;
  uint64_t scans = 0, tightened = 0;

  Hamerly(const F* X_, uint64_t n_, uint64_t d_, uint64_t k_, const F* C0, int threads)   // new (:418-456)
      : X(X_), n(n_), d(d_), k(k_), C(C0, C0 + k_ * d_), memberships(n_, 0), prev(n_, 0), upper(n_, (F)0), lower(n_, (F)0),
        counts(k_, 0), sums(k_ * d_, (Acc)0), n_threads(threads) {
    par_rows(n, n_threads, [&](uint64_t lo, uint64_t hi) {
      for (uint64_t i = lo; i < hi; ++i) two_closest_centroids<F>(C.data(), k, d, X + i * d, &memberships[i], &upper[i], &lower[i]);
    });
    for (uint64_t i = 0; i < n; ++i) {
      counts[memberships[i]] += 1;
      for (uint64_t j = 0; j < d; ++j) sums[memberships[i] * d + j] = sums[memberships[i] * d + j] + (Acc)X[i * d + j];
    }
  }
  void reassign() {   // reassign_observations (:472-524)
    std::vector<F> near(k);
    for (uint64_t c = 0; c < k; ++c) {   // nearest_inter_centroid_distances (:462-470)
      uint64_t ii; F cl;
      two_closest_centroids<F>(C.data(), k, d, C.data() + c * d, &ii, &cl, &near[c]);
    }
    std::atomic<uint64_t> sc{0}, ti{0};
    par_rows(n, n_threads, [&](uint64_t lo, uint64_t hi) {
      uint64_t my_sc = 0, my_ti = 0;
      for (uint64_t i = lo; i < hi; ++i) {
        const uint64_t cur = memberships[i];
        prev[i] = cur;
        const F half = near[cur] / (F)2;
        const F threshold = half > lower[i] ? half : lower[i];   // F::max
        if (upper[i] > threshold) {
          upper[i] = l2_dist_row<F>(X + i * d, C.data() + cur * d, d);
          ++my_ti;
          if (upper[i] > threshold) {
            two_closest_centroids<F>(C.data(), k, d, X + i * d, &memberships[i], &upper[i], &lower[i]);
            ++my_sc;
          }
        }
      }
      sc += my_sc; ti += my_ti;
    });
    scans += sc; tightened += ti;
    for (uint64_t i = 0; i < n; ++i) {
      if (prev[i] != memberships[i]) {
        counts[prev[i]] -= 1;
        counts[memberships[i]] += 1;
        for (uint64_t j = 0; j < d; ++j) sums[prev[i] * d + j] = sums[prev[i] * d + j] - (Acc)X[i * d + j];
        for (uint64_t j = 0; j < d; ++j) sums[memberships[i] * d + j] = sums[memberships[i] * d + j] + (Acc)X[i * d + j];
      }
    }
  }
  // recompute_centroids (:527-553): returns convergence_dist, fills distances_moved
  F recompute(std::vector<F>& moved) {
    std::vector<F> Cn(k * d);
    for (uint64_t c = 0; c < k; ++c)
      for (uint64_t j = 0; j < d; ++j) {
        const Acc v = sums[c * d + j] + (Acc)C[c * d + j];
        Cn[c * d + j] = (F)(v / (Acc)(F)(counts[c] + 1));
      }
    moved.assign(k, (F)0);
    for (uint64_t c = 0; c < k; ++c) moved[c] = l2_dist_row<F>(C.data() + c * d, Cn.data() + c * d, d);
    const F conv = l2_distance<F>(C.data(), Cn.data(), k * d);
    C.swap(Cn);
    return conv;
  }
  void update_bounds(const std::vector<F>& moved) {   // (:555-568)
    uint64_t f1, f2;
    two_farthest_indices<F>(moved.data(), k, &f1, &f2);
    for (uint64_t i = 0; i < n; ++i) {
      upper[i] = upper[i] + moved[memberships[i]];
      lower[i] = lower[i] - (memberships[i] == f1 ? moved[f2] : moved[f1]);
    }
  }
};

// fit_hamerly + get_kmeans_result (KM/algorithm.rs:248-291, 345-367)
template <class F, class Acc>
int fit_hamerly(const F* X, uint64_t n, uint64_t d, uint64_t k, uint64_t n_runs, uint64_t max_iter, F tol, int method,
                const F* precomputed, Xoshiro256Plus& rng_in, F* C_out, F* counts_out, F* inertia_out,
                uint64_t* n_iter_out, uint64_t* labels_out, uint64_t* scans_out, int n_threads, int n_splits) {
  if (k == 0) return 10;
  if (n_runs == 0) return 11;
  if (tol <= (F)0) return 12;
  if (max_iter == 0) return 13;
  Xoshiro256Plus rng = rng_in;
  F min_inertia = std::numeric_limits<F>::infinity();
  bool have_best = false;
  std::vector<F> best(k * d), C0(k * d);
  std::vector<uint64_t> best_members(n, 0);
  uint64_t last_iters = 0, scans = 0;
  for (uint64_t run = 0; run < n_runs; ++run) {
    const int st = run_init<F>(method, k, X, n, d, precomputed, rng, C0.data(), n_threads, n_splits);
    if (st) return st;
    Hamerly<F, Acc> h(X, n, d, k, C0.data(), n_threads);
    uint64_t n_iter = 0;
    std::vector<F> moved;
    F inertia;
    for (;;) {
      if (n_iter > 0) h.reassign();
      n_iter += 1;
      const F conv = h.recompute(moved);
      if (conv < tol || n_iter == max_iter) { inertia = compute_inertia<F, Acc>(X, n, d, h.memberships.data(), h.C.data()); break; }
      h.update_bounds(moved);
    }
    last_iters = n_iter;
    scans += h.scans;
    if (inertia < min_inertia) {
      min_inertia = inertia;
      best = h.C;
      best_members = h.memberships;
      have_best = true;
    }
  }
  if (n_iter_out) *n_iter_out = last_iters;
  if (scans_out) *scans_out = scans;
  if (labels_out) std::memcpy(labels_out, best_members.data(), sizeof(uint64_t) * n);
  if (!have_best) return 1;
  for (uint64_t c = 0; c < k; ++c) counts_out[c] = (F)0;
  for (uint64_t i = 0; i < n; ++i) counts_out[best_members[i]] = counts_out[best_members[i]] + (F)1;
  std::memcpy(C_out, best.data(), sizeof(F) * k * d);
  *inertia_out = min_inertia / (F)n;
  return 0;
}

// ---------------------------------------------------------------------------
// SilhouetteScore::silhouette_score (src/metrics_clustering.rs:65-135), labels as u64
// ---------------------------------------------------------------------------
template <class F> F silhouette_score(const F* X, uint64_t n, uint64_t d, const uint64_t* labels, F* per_sample) {
  // label_count: distinct labels with their counts
  std::vector<uint64_t> uniq(labels, labels + n);
  std::sort(uniq.begin(), uniq.end());
  uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
  const size_t L = uniq.size();
  if (L == 1) return (F)1;   // :77-80
  std::vector<uint64_t> id(n), count(L, 0);
  for (uint64_t i = 0; i < n; ++i) {
    id[i] = (uint64_t)(std::lower_bound(uniq.begin(), uniq.end(), labels[i]) - uniq.begin());
    count[id[i]] += 1;
  }
  std::vector<F> total(L, (F)0), diff(d);
  F score = (F)0;
  for (uint64_t i = 0; i < n; ++i) {
    for (uint64_t j = 0; j < n; ++j) {
      // add_point (:61-63): eval_sample.sub(&other_sample).mapv(|x| x * x).sum().sqrt(); sum() = ndarray's unrolled fold
      for (uint64_t f = 0; f < d; ++f) { const F v = X[i * d + f] - X[j * d + f]; diff[f] = v * v; }
      total[id[j]] = total[id[j]] + std::sqrt(ndarray_sum<F>(diff.data(), d));
    }
    F a_x = (F)0, b_x = (F)0;
    bool have_b = false;
    for (size_t l = 0; l < L; ++l) {
      if (l == id[i]) a_x = count[l] == 1 ? (F)0 : total[l] / (F)(count[l] - 1);
      else {
        const F m = total[l] / (F)count[l];
        if (!have_b || m < b_x) b_x = m;
        have_b = true;
      }
      total[l] = (F)0;
    }
    const F s = a_x >= b_x ? (b_x - a_x) / a_x : (b_x - a_x) / b_x;
    if (per_sample) per_sample[i] = s;
    score = score + s;
  }
  return score / (F)n;
}

// FitWith::fit_with (KM/algorithm.rs:635-715), one mini-batch step.
// model_in: has_model != 0 -> (C_inout, counts_inout) hold the previous model.
// returns 0 converged (Ok), 20 NotConverged (model still written), else error.
template <class F>
int fit_with(const F* X, uint64_t n, uint64_t d, uint64_t k, uint64_t n_runs, F tol, int method, const F* precomputed,
             Xoshiro256Plus& rng_in, int has_model, F* C_inout, F* counts_inout, F* inertia_out, int n_threads,
             int n_splits) {
  if (k == 0) return 10;
  if (n_runs == 0) return 11;
  if (tol <= (F)0) return 12;
  std::vector<F> dists(n, (F)0);
  if (!has_model) {
    if (method == INIT_PRECOMPUTED) {
      std::memcpy(C_inout, precomputed, sizeof(F) * k * d);
    } else {
      Xoshiro256Plus rng = rng_in;
      std::vector<F> C(k * d);
      bool have = false;
      F best = 0;
      // (0..n_runs).map(|_| (init, dists.sum())).min_by(|d1, d2| d1 < d2 ? Less : Greater)
      for (uint64_t r = 0; r < n_runs; ++r) {
        const int st = run_init<F>(method, k, X, n, d, precomputed, rng, C.data(), n_threads, n_splits);
        if (st) return st;
        assign<F>(X, n, d, C.data(), k, nullptr, dists.data(), n_threads);
        const F cost = ndarray_sum<F>(dists.data(), n);
        // Iterator::min_by folds with `match compare(&acc, &x) { Greater => x, _ => acc }`,
        // compare(acc, x) = acc < x ? Less : Greater  =>  x replaces acc unless acc < x.
        if (!have || !(best < cost)) {
          best = cost;
          std::memcpy(C_inout, C.data(), sizeof(F) * k * d);
          have = true;
        }
      }
    }
    for (uint64_t c = 0; c < k; ++c) counts_inout[c] = (F)0;
  }
  std::vector<uint64_t> lab(n);
  assign<F>(X, n, d, C_inout, k, lab.data(), dists.data(), n_threads);
  std::vector<F> Cn(k * d);
  compute_centroids_incremental<F>(X, lab.data(), n, d, k, C_inout, counts_inout, Cn.data());
  *inertia_out = ndarray_sum<F>(dists.data(), n) / (F)n;
  const F shift = l2_distance<F>(C_inout, Cn.data(), k * d);
  std::memcpy(C_inout, Cn.data(), sizeof(F) * k * d);
  return shift < tol ? 0 : 20;
}

}  // namespace

// ---------------------------------------------------------------------------
// C interface for ctypes (tests/, bench.py cpu_baseline, smoke()).
// ---------------------------------------------------------------------------
#define ORC_EXPORT extern "C" __attribute__((visibility("default")))

ORC_EXPORT void orc_xoshiro_seed(uint64_t seed, uint64_t* state) {
  Xoshiro256Plus g = Xoshiro256Plus::seed_from_u64(seed);
  std::memcpy(state, g.s, 32);
}
ORC_EXPORT uint64_t orc_xoshiro_next_u64(uint64_t* state) {
  Xoshiro256Plus g;
  std::memcpy(g.s, state, 32);
  const uint64_t r = g.next_u64();
  std::memcpy(state, g.s, 32);
  return r;
}
ORC_EXPORT uint32_t orc_xoshiro_next_u32(uint64_t* state) { return (uint32_t)(orc_xoshiro_next_u64(state) >> 32); }
ORC_EXPORT double orc_u01_f64(uint64_t* state) {
  Xoshiro256Plus g;
  std::memcpy(g.s, state, 32);
  const double r = u01_f64(g);
  std::memcpy(state, g.s, 32);
  return r;
}
ORC_EXPORT float orc_u01_f32(uint64_t* state) {
  Xoshiro256Plus g;
  std::memcpy(g.s, state, 32);
  const float r = u01_f32(g);
  std::memcpy(state, g.s, 32);
  return r;
}
ORC_EXPORT uint64_t orc_gen_range_u64(uint64_t* state, uint64_t low, uint64_t high_excl) {
  Xoshiro256Plus g;
  std::memcpy(g.s, state, 32);
  const uint64_t r = gen_range_incl_u64(g, low, high_excl - 1);
  std::memcpy(state, g.s, 32);
  return r;
}
ORC_EXPORT int orc_index_sample(uint64_t* state, uint64_t length, uint64_t amount, uint64_t* out) {
  Xoshiro256Plus g;
  std::memcpy(g.s, state, 32);
  bool ok;
  std::vector<uint64_t> v = index_sample(g, length, amount, &ok);
  std::memcpy(state, g.s, 32);
  if (!ok) return 3;
  std::memcpy(out, v.data(), 8 * v.size());
  return 0;
}

#define ORC_DEFINE(F, SFX)                                                                                            \
  ORC_EXPORT F orc_sq_l2_dist_##SFX(const F* a, const F* b, uint64_t d) { return sq_l2_dist<F>(a, b, d); }           \
  ORC_EXPORT F orc_l2_distance_##SFX(const F* a, const F* b, uint64_t len) { return l2_distance<F>(a, b, len); }     \
  ORC_EXPORT F orc_sum_##SFX(const F* x, uint64_t n) { return ndarray_sum<F>(x, n); }                                \
  ORC_EXPORT void orc_assign_##SFX(const F* X, uint64_t n, uint64_t d, const F* C, uint64_t k, uint64_t* labels,     \
                                   F* dists, int n_threads) {                                                        \
    assign<F>(X, n, d, C, k, labels, dists, n_threads);                                                              \
  }                                                                                                                   \
  ORC_EXPORT void orc_compute_centroids_##SFX(const F* C_old, const F* X, const uint64_t* labels, uint64_t n,        \
                                              uint64_t d, uint64_t k, F* C_new, int acc_f64) {                       \
    if (acc_f64) compute_centroids<F, double>(C_old, X, labels, n, d, k, C_new);                                     \
    else compute_centroids<F, F>(C_old, X, labels, n, d, k, C_new);                                                  \
  }                                                                                                                   \
  ORC_EXPORT void orc_compute_centroids_incremental_##SFX(const F* X, const uint64_t* labels, uint64_t n,            \
                                                          uint64_t d, uint64_t k, const F* C_old, F* counts,         \
                                                          F* C_new) {                                                 \
    compute_centroids_incremental<F>(X, labels, n, d, k, C_old, counts, C_new);                                      \
  }                                                                                                                   \
  ORC_EXPORT int64_t orc_weighted_index_##SFX(const F* w, uint64_t n, uint64_t* state) {                             \
    Xoshiro256Plus g;                                                                                                 \
    std::memcpy(g.s, state, 32);                                                                                      \
    std::vector<F> cum;                                                                                               \
    const int64_t r = weighted_index_sample<F>(w, n, g, cum);                                                         \
    std::memcpy(state, g.s, 32);                                                                                      \
    return r;                                                                                                         \
  }                                                                                                                   \
  ORC_EXPORT int orc_init_##SFX(int method, uint64_t k, const F* X, uint64_t n, uint64_t d, const F* precomputed,    \
                                uint64_t* state, F* C, int n_threads, int n_splits) {                                \
    Xoshiro256Plus g;                                                                                                 \
    std::memcpy(g.s, state, 32);                                                                                      \
    const int st = run_init<F>(method, k, X, n, d, precomputed, g, C, n_threads, n_splits);                          \
    std::memcpy(state, g.s, 32);                                                                                      \
    return st;                                                                                                        \
  }                                                                                                                   \
  ORC_EXPORT int orc_weighted_kmeanspp_##SFX(uint64_t k, const F* X, uint64_t n, uint64_t d, const F* w,             \
                                             uint64_t* state, F* C, uint64_t* picked, int n_threads) {               \
    Xoshiro256Plus g;                                                                                                 \
    std::memcpy(g.s, state, 32);                                                                                      \
    const int st = weighted_kmeanspp<F>(k, X, n, d, w, g, C, picked, n_threads);                                     \
    std::memcpy(state, g.s, 32);                                                                                      \
    return st;                                                                                                        \
  }                                                                                                                   \
  ORC_EXPORT uint64_t orc_sample_subsequent_##SFX(const F* dists, uint64_t n, F multiplier, uint64_t seed,           \
                                                  int n_splits, uint64_t* out) {                                     \
    std::vector<uint64_t> v;                                                                                          \
    sample_subsequent_candidates<F>(dists, n, multiplier, seed, n_splits, v);                                        \
    std::memcpy(out, v.data(), 8 * v.size());                                                                         \
    return v.size();                                                                                                  \
  }                                                                                                                   \
  ORC_EXPORT void orc_membership_counts_##SFX(const F* C, uint64_t k, const F* X, uint64_t n, uint64_t d, F* counts, \
                                              int n_threads) {                                                       \
    cluster_membership_counts<F>(C, k, X, n, d, counts, n_threads);                                                  \
  }                                                                                                                   \
  ORC_EXPORT int orc_fit_##SFX(const F* X, uint64_t n, uint64_t d, uint64_t k, uint64_t n_runs, uint64_t max_iter,   \
                               F tol, int method, const F* precomputed, const uint64_t* state, F* C_out,             \
                               F* counts_out, F* inertia_out, uint64_t* n_iter_out, uint64_t* labels_out,            \
                               int n_threads, int n_splits, int acc_f64) {                                           \
    Xoshiro256Plus g;                                                                                                 \
    std::memcpy(g.s, state, 32);                                                                                      \
    if (acc_f64)                                                                                                      \
      return fit_lloyd<F, double>(X, n, d, k, n_runs, max_iter, tol, method, precomputed, g, C_out, counts_out,      \
                                  inertia_out, n_iter_out, labels_out, n_threads, n_splits);                         \
    return fit_lloyd<F, F>(X, n, d, k, n_runs, max_iter, tol, method, precomputed, g, C_out, counts_out,             \
                           inertia_out, n_iter_out, labels_out, n_threads, n_splits);                                \
  }                                                                                                                   \
  ORC_EXPORT int orc_fit_hamerly_##SFX(const F* X, uint64_t n, uint64_t d, uint64_t k, uint64_t n_runs,              \
                                       uint64_t max_iter, F tol, int method, const F* precomputed,                   \
                                       const uint64_t* state, F* C_out, F* counts_out, F* inertia_out,               \
                                       uint64_t* n_iter_out, uint64_t* labels_out, uint64_t* scans_out,              \
                                       int n_threads, int n_splits, int acc_f64) {                                   \
    Xoshiro256Plus g;                                                                                                 \
    std::memcpy(g.s, state, 32);                                                                                      \
    if (acc_f64)                                                                                                      \
      return fit_hamerly<F, double>(X, n, d, k, n_runs, max_iter, tol, method, precomputed, g, C_out, counts_out,    \
                                    inertia_out, n_iter_out, labels_out, scans_out, n_threads, n_splits);            \
    return fit_hamerly<F, F>(X, n, d, k, n_runs, max_iter, tol, method, precomputed, g, C_out, counts_out,           \
                             inertia_out, n_iter_out, labels_out, scans_out, n_threads, n_splits);                   \
  }                                                                                                                   \
  ORC_EXPORT void orc_two_closest_##SFX(const F* C, uint64_t k, uint64_t d, const F* x, uint64_t* idx, F* closest,   \
                                        F* second) {                                                                  \
    two_closest_centroids<F>(C, k, d, x, idx, closest, second);                                                      \
  }                                                                                                                   \
  ORC_EXPORT void orc_two_farthest_##SFX(const F* m, uint64_t k, uint64_t* f1, uint64_t* f2) {                       \
    two_farthest_indices<F>(m, k, f1, f2);                                                                            \
  }                                                                                                                   \
  ORC_EXPORT F orc_compute_inertia_##SFX(const F* X, uint64_t n, uint64_t d, const uint64_t* labels, const F* C) {   \
    return compute_inertia<F, F>(X, n, d, labels, C);                                                                 \
  }                                                                                                                   \
  ORC_EXPORT F orc_silhouette_##SFX(const F* X, uint64_t n, uint64_t d, const uint64_t* labels, F* per_sample) {     \
    return silhouette_score<F>(X, n, d, labels, per_sample);                                                         \
  }                                                                                                                   \
  ORC_EXPORT int orc_fit_with_##SFX(const F* X, uint64_t n, uint64_t d, uint64_t k, uint64_t n_runs, F tol,          \
                                    int method, const F* precomputed, const uint64_t* state, int has_model,          \
                                    F* C_inout, F* counts_inout, F* inertia_out, int n_threads, int n_splits) {      \
    Xoshiro256Plus g;                                                                                                 \
    std::memcpy(g.s, state, 32);                                                                                      \
    return fit_with<F>(X, n, d, k, n_runs, tol, method, precomputed, g, has_model, C_inout, counts_inout,            \
                       inertia_out, n_threads, n_splits);                                                            \
  }

ORC_DEFINE(float, f32)
ORC_DEFINE(double, f64)

ORC_EXPORT int orc_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

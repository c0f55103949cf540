// Host-side random choices of the seeding path, drawn from the caller's RNG
// through the lkm_rng trampolines.  The sequence of next_u32 / next_u64 calls
// and the arithmetic on their results follow rand 0.8.5 (the version
// ndarray-rand 0.15 re-exports, algorithms/linfa-clustering/Cargo.toml), so the
// same `R: Rng` picks the same rows as linfa does:
//   KM/init.rs:111       rand::seq::index::sample(rng, n, k)
//   KM/init.rs:131,148   WeightedIndex::new(w).sample(rng)  (the Uniform<F> draw)
//   KM/init.rs:197,211   rng.gen_range(0..n), rng.gen_range(0..u64::MAX)
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <unordered_set>
#include <vector>

#include "../../include/lkm.h"

namespace lkm {

struct HostRng {
  lkm_rng r;
  uint32_t u32() { return r.next_u32(r.user); }
  uint64_t u64() { return r.next_u64(r.user); }
};

// UniformInt::sample_single_inclusive, u64 / usize flavour (draws next_u64)
inline uint64_t range_incl_u64(HostRng& g, uint64_t low, uint64_t high) {
  const uint64_t range = high - low + 1;
  if (range == 0) return g.u64();
  const int lz = __builtin_clzll(range);
  const uint64_t zone = (range << lz) - 1;
  for (;;) {
    const unsigned __int128 m = (unsigned __int128)g.u64() * range;
    if ((uint64_t)m <= zone) return low + (uint64_t)(m >> 64);
  }
}
// u32 flavour (draws next_u32)
inline uint32_t range_incl_u32(HostRng& g, uint32_t low, uint32_t high) {
  const uint32_t range = high - low + 1;
  if (range == 0) return g.u32();
  const int lz = __builtin_clz(range);
  const uint32_t zone = (range << lz) - 1;
  for (;;) {
    const uint64_t m = (uint64_t)g.u32() * range;
    if ((uint32_t)m <= zone) return low + (uint32_t)(m >> 32);
  }
}

// Uniform::new(0, len) as a distribution object (modulus zone), used only by
// the rejection branch of index::sample.
template <class U> struct UniformObj;
template <> struct UniformObj<uint32_t> {
  uint32_t range, zone;
  explicit UniformObj(uint32_t len) : range(len), zone(0xFFFFFFFFu - (uint32_t)((0xFFFFFFFFu - len + 1) % len)) {}
  uint32_t draw(HostRng& g) const {
    for (;;) {
      const uint64_t m = (uint64_t)g.u32() * range;
      if ((uint32_t)m <= zone) return (uint32_t)(m >> 32);
    }
  }
};
template <> struct UniformObj<uint64_t> {
  uint64_t range, zone;
  explicit UniformObj(uint64_t len) : range(len), zone(~0ULL - ((~0ULL - len + 1) % len)) {}
  uint64_t draw(HostRng& g) const {
    for (;;) {
      const unsigned __int128 m = (unsigned __int128)g.u64() * range;
      if ((uint64_t)m <= zone) return (uint64_t)(m >> 64);
    }
  }
};

// rand::seq::index::sample — returns false where the reference panics (amount > length)
inline bool index_sample(HostRng& g, uint64_t length, uint64_t amount, std::vector<uint64_t>& out) {
  out.clear();
  if (amount > length) return false;
  if (length > 0xFFFFFFFFULL) {
    UniformObj<uint64_t> u(length);
    std::unordered_set<uint64_t> seen;
    while (out.size() < amount) {
      uint64_t p = u.draw(g);
      while (!seen.insert(p).second) p = u.draw(g);
      out.push_back(p);
    }
    return true;
  }
  const uint32_t len = (uint32_t)length, amt = (uint32_t)amount;
  const int big = len >= 500000u;
  enum { FLOYD, INPLACE, REJECT } algo;
  if (amt < 163) {
    const float c0[2] = {1.6f, 8.0f / 45.0f}, c1[2] = {10.0f, 70.0f / 9.0f};
    const float a = (float)amt, m4 = c0[big] * a;
    algo = (amt > 11 && (float)len < (c1[big] + m4) * a) ? INPLACE : FLOYD;
  } else {
    const float c[2] = {270.0f, 330.0f / 9.0f};
    algo = ((float)len < c[big] * (float)amt) ? INPLACE : REJECT;
  }
  if (algo == INPLACE) {
    std::vector<uint32_t> idx(len);
    for (uint32_t i = 0; i < len; ++i) idx[i] = i;
    for (uint32_t i = 0; i < amt; ++i) std::swap(idx[i], idx[range_incl_u32(g, i, len - 1)]);
    out.assign(idx.begin(), idx.begin() + amt);
  } else if (algo == FLOYD) {
    const bool shuffled_variant = amt < 50;
    std::vector<uint32_t> idx;
    idx.reserve(amt);
    for (uint32_t j = len - amt; j < len; ++j) {
      const uint32_t t = range_incl_u32(g, 0, j);
      const auto hit = std::find(idx.begin(), idx.end(), t);
      if (hit == idx.end()) idx.push_back(t);
      else if (shuffled_variant) idx.insert(hit, j);
      else idx.push_back(j);
    }
    if (!shuffled_variant)
      for (uint32_t i = amt - 1; i >= 1; --i) std::swap(idx[i], idx[range_incl_u32(g, 0, i)]);
    out.assign(idx.begin(), idx.end());
  } else {
    UniformObj<uint32_t> u(len);
    std::unordered_set<uint32_t> seen;
    while (out.size() < amt) {
      uint32_t p = u.draw(g);
      while (!seen.insert(p).second) p = u.draw(g);
      out.push_back(p);
    }
  }
  return true;
}

// The [0,1) value UniformFloat builds from the top mantissa bits.
inline float unit_f32(HostRng& g) {
  const uint32_t b = (g.u32() >> 9) | 0x3F800000u;
  float v;
  std::memcpy(&v, &b, 4);
  return v - 1.0f;
}
inline double unit_f64(HostRng& g) {
  const uint64_t b = (g.u64() >> 12) | 0x3FF0000000000000ULL;
  double v;
  std::memcpy(&v, &b, 8);
  return v - 1.0;
}
template <class F> inline F unit(HostRng& g);
template <> inline float unit<float>(HostRng& g) { return unit_f32(g); }
template <> inline double unit<double>(HostRng& g) { return unit_f64(g); }

// Uniform<F>::new(0, total).sample(rng): value0_1 * scale + 0 with the scale
// shrunk until scale * max_rand < total (a no-op for low = 0, kept for fidelity).
template <class F> inline F weighted_threshold(HostRng& g, F total) {
  const F max_rand = sizeof(F) == 4 ? (F)(1.0f - 1.1920929e-7f) : (F)(1.0 - 2.220446049250313e-16);
  F scale = total;
  while (scale * max_rand + (F)0 >= total) scale = std::nextafter(scale, (F)0);
  const F v = unit<F>(g);
  return v * scale + (F)0;
}

// rng.gen_range(0.0..1.0) (f64), KM/init.rs:263
inline double range_unit_f64(HostRng& g) {
  for (;;) {
    const double v = unit_f64(g) * 1.0 + 0.0;
    if (v < 1.0) return v;
  }
}

}  // namespace lkm

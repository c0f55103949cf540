// The distance pass of k-means++ (KM/init.rs:138-147): after every pick, d_i = min(d_i, ||c_new - x_i||^2) for ALL rows and
// wd_i = w_i d_i for the next weighted draw.  k - 1 such passes read the whole matrix once each — for cfg3 that is 255 x 5.12 GB,
// ten times the traffic of the Lloyd iterations that follow — so the pass has to run at the HBM rate while keeping the
// reference's arithmetic: the squared distance in the direct form, feature by feature, never contracted (the min distances
// are compared bit for bit with the oracle's).  The generic assignment kernel (one centroid = no reuse, runtime d) reached
// 3.1 TB/s.  Here: 64-row tiles stream through a 3-stage cp.async ring into shared memory with a row stride of d + 4 floats
// (16-byte chunks land conflict-free, and so do the 128-bit reads of thread = row), the centroid sits in registers, each
// thread walks its row in feature order.
#pragma once
#include "lkm_kernels.cuh"

namespace lkm {

constexpr int kSpRows = 64;      // rows per tile = threads per CTA (A/B on cfg3 / cfg4: 64 x 3 stages 237 / 577 ms of seeding,
                                 // 128 x 3: 263 / 602, 128 x 2: 262 / 574, 64 x 4: 355 / 635)
constexpr int kSpStages = 3;

struct SeedPassArgs {
  const float* X;
  unsigned long long n;
  const float* c;          // the centroid picked last (d floats)
  float* dists;            // running minimum (in / out)
  int min_combine;         // 0: first pass, dists = distance
  const float* weights;    // optional
  float* wdists;           // optional: (weights ? weights[i] : 1) * dists[i]
};

template <int D> __host__ __device__ constexpr size_t seedpass_smem_bytes() { return (size_t)kSpStages * kSpRows * (D + 4) * 4; }

__device__ __forceinline__ void sp_cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void sp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void sp_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int D>
__global__ void __launch_bounds__(kSpRows) seed_pass_kernel(const SeedPassArgs a) {
  static_assert(D % 4 == 0 && D >= 16 && D <= 128, "row width");
  constexpr int RS = D + 4;          // shared-memory row stride in floats
  constexpr int CPR = D / 4;         // 16-byte chunks per row
  constexpr int CPT = CPR;           // chunks per thread and tile (kSpRows * CPR / kSpRows)
  extern __shared__ __align__(16) float sp_smem[];
  const int t = threadIdx.x;
  const unsigned long long n_tiles = (a.n + kSpRows - 1) / kSpRows;
  float c[D];
#pragma unroll
  for (int j = 0; j < D; j += 4) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(a.c + j));
    c[j] = v.x; c[j + 1] = v.y; c[j + 2] = v.z; c[j + 3] = v.w;
  }
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(sp_smem);
  auto issue = [&](unsigned long long tile, int st) {
    const unsigned long long row0 = tile * kSpRows;
    const float* src = a.X + row0 * D;
    const uint32_t dst = s_base + (uint32_t)st * (uint32_t)(kSpRows * RS * 4);
    const bool whole = row0 + kSpRows <= a.n;
#pragma unroll
    for (int i = 0; i < CPT; ++i) {
      const int q = t + i * kSpRows;            // chunk of the tile: consecutive threads, consecutive 16 bytes of X
      const int r = q / CPR, cc = q % CPR;
      const bool ok = whole || row0 + (unsigned)r < a.n;
      sp_cp_async16(dst + (uint32_t)(r * RS + cc * 4) * 4u, ok ? (const void*)(src + (size_t)q * 4) : (const void*)a.X, ok ? 16 : 0);
    }
  };
  unsigned long long tile = blockIdx.x;
#pragma unroll
  for (int s = 0; s < kSpStages - 1; ++s) {
    if (tile + (unsigned long long)s * gridDim.x < n_tiles) issue(tile + (unsigned long long)s * gridDim.x, s);
    sp_commit();
  }
  int st = 0;
  for (; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long ahead = tile + (unsigned long long)(kSpStages - 1) * gridDim.x;
    int st_ahead = st + kSpStages - 1;
    if (st_ahead >= kSpStages) st_ahead -= kSpStages;
    if (ahead < n_tiles) issue(ahead, st_ahead);   // the stage read in the previous iteration (all threads passed its barrier)
    sp_commit();
    sp_wait<kSpStages - 1>();
    __syncthreads();
    const float* xr = sp_smem + (size_t)st * (kSpRows * RS) + (size_t)t * RS;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < D; j += 4) {
      const float4 v = *reinterpret_cast<const float4*>(xr + j);
      const float d0 = Rn<float>::sub(c[j], v.x), d1 = Rn<float>::sub(c[j + 1], v.y);
      const float d2 = Rn<float>::sub(c[j + 2], v.z), d3 = Rn<float>::sub(c[j + 3], v.w);
      acc = Rn<float>::add(acc, Rn<float>::mul(d0, d0));
      acc = Rn<float>::add(acc, Rn<float>::mul(d1, d1));
      acc = Rn<float>::add(acc, Rn<float>::mul(d2, d2));
      acc = Rn<float>::add(acc, Rn<float>::mul(d3, d3));
    }
    const unsigned long long row = tile * kSpRows + t;
    if (row < a.n) {
      float outd = acc;
      if (a.min_combine) {
        const float old = a.dists[row];
        if (acc < old) a.dists[row] = acc; else outd = old;
      } else {
        a.dists[row] = acc;
      }
      if (a.wdists != nullptr) a.wdists[row] = a.weights != nullptr ? outd * a.weights[row] : outd;
    }
    __syncthreads();   // the stage may be refilled by the next iteration's issue
    if (++st == kSpStages) st = 0;
  }
  sp_wait<0>();
}

// ---- the same pass with pruning --------------------------------------------------------------------------------------------
// Row i keeps, next to its running minimum d_i, the seed s_i that realises it.  A new seed c with ||c - s_i|| > 2 ||x_i - s_i||
// cannot be closer to x_i than s_i (triangle inequality), so the minimum does not change and the row need not be READ: 8 bytes
// (d_i, s_i) instead of d * 4.  As the picks proceed most rows sit next to a seed and far from the new one; on blobs with one
// seed per blob at the end that halves the traffic of the whole seeding, and the late passes cost little more than the
// 80 MB of row state.  The factors below (2.001 on the radius, 1.0001 / 0.9999 on the two distances) are three orders of
// magnitude above the (d + 2) eps relative error of a squared distance, so a skipped row's new distance would have compared
// `>= d_i` in the reference's arithmetic too: the running minima stay bit-identical to the unpruned pass (and to the oracle).
struct SeedPruneArgs {
  const float* X;
  unsigned long long n;
  const float* seeds;      // the seeds chosen so far, row-major; the last one (index n_seeds - 1) is the new seed
  int n_seeds;
  float* seed_dist;        // [n_seeds - 1] lower bounds on ||new - seed_j|| (seed_dist_kernel)
  float* dists;
  uint32_t* nearest;       // seed index behind dists[i]
  const float* weights;
  float* wdists;
};

// warp j: lower bound on the distance between the new seed and seed j
template <int D>
__global__ void __launch_bounds__(256) seed_dist_kernel(const SeedPruneArgs a) {
  const int lane = threadIdx.x & 31;
  const int j = (blockIdx.x * 256 + threadIdx.x) >> 5;
  if (j >= a.n_seeds - 1) return;
  const float* cn = a.seeds + (size_t)(a.n_seeds - 1) * D;
  const float* cj = a.seeds + (size_t)j * D;
  float s = 0.f;
  for (int f = lane; f < D; f += 32) { const float df = cn[f] - cj[f]; s = fmaf(df, df, s); }
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) a.seed_dist[j] = sqrtf(s) * 0.9999f;   // NaN stays NaN: such a seed prunes nothing
}

constexpr int kSqRows = 64;   // rows per tile = threads per CTA

template <int D> __host__ __device__ constexpr size_t seedprune_smem_bytes() { return (size_t)kSqRows * (D + 4) * 4; }

template <int D>
__global__ void __launch_bounds__(kSqRows) seed_pass_pruned_kernel(const SeedPruneArgs a) {
  static_assert(D % 4 == 0 && D >= 16 && D <= 128, "row width");
  constexpr int RS = D + 4, CPR = D / 4;
  extern __shared__ __align__(16) float sp_smem[];
  __shared__ uint16_t list[kSqRows];
  __shared__ float old_s[kSqRows];
  __shared__ int warp_cnt[2];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const unsigned long long n_tiles = (a.n + kSqRows - 1) / kSqRows;
  const bool first = a.n_seeds == 1;
  const uint32_t me = (uint32_t)(a.n_seeds - 1);
  float c[D];
  {
    const float* cn = a.seeds + (size_t)me * D;
#pragma unroll
    for (int j = 0; j < D; j += 4) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(cn + j));
      c[j] = v.x; c[j + 1] = v.y; c[j + 2] = v.z; c[j + 3] = v.w;
    }
  }
  const uint32_t s_base = (uint32_t)__cvta_generic_to_shared(sp_smem);
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long row0 = tile * kSqRows, row = row0 + t;
    // ---- which rows of the tile can the new seed reach? ----
    bool need = row < a.n;
    float old = INFINITY;
    if (need && !first) {
      old = a.dists[row];
      const float sd = a.seed_dist[a.nearest[row]];
      // only a normal, moderate distance is trusted for pruning
      if (old > 1e-30f && old < 1e30f && sd > 2.001f * (sqrtf(old) * 1.0001f)) need = false;
    }
    const unsigned m = __ballot_sync(0xffffffffu, need);
    if (lane == 0) warp_cnt[warp] = __popc(m);
    __syncthreads();
    const int base = warp == 0 ? 0 : warp_cnt[0], cnt = warp_cnt[0] + warp_cnt[1];
    if (need) {
      const int slot = base + __popc(m & ((1u << lane) - 1u));
      list[slot] = (uint16_t)t;
      old_s[slot] = old;
    }
    __syncthreads();
    // ---- bring the surviving rows in (16-byte chunks, consecutive threads = consecutive chunks of a row) ----
    for (int q = t; q < cnt * CPR; q += kSqRows) {
      const int slot = q / CPR, cc = q % CPR;
      sp_cp_async16(s_base + (uint32_t)(slot * RS + cc * 4) * 4u, a.X + (row0 + list[slot]) * D + (size_t)cc * 4, 16);
    }
    sp_commit();
    sp_wait<0>();
    __syncthreads();
    if (t < cnt) {
      const float* xr = sp_smem + (size_t)t * RS;
      float acc = 0.f;
#pragma unroll
      for (int j = 0; j < D; j += 4) {
        const float4 v = *reinterpret_cast<const float4*>(xr + j);
        const float d0 = Rn<float>::sub(c[j], v.x), d1 = Rn<float>::sub(c[j + 1], v.y);
        const float d2 = Rn<float>::sub(c[j + 2], v.z), d3 = Rn<float>::sub(c[j + 3], v.w);
        acc = Rn<float>::add(acc, Rn<float>::mul(d0, d0));
        acc = Rn<float>::add(acc, Rn<float>::mul(d1, d1));
        acc = Rn<float>::add(acc, Rn<float>::mul(d2, d2));
        acc = Rn<float>::add(acc, Rn<float>::mul(d3, d3));
      }
      const unsigned long long r = row0 + list[t];
      if (first || acc < old_s[t]) {
        a.dists[r] = acc;
        a.nearest[r] = me;
        if (a.wdists != nullptr) a.wdists[r] = a.weights != nullptr ? acc * a.weights[r] : acc;
      }
    }
    __syncthreads();   // list / old_s / the tile are rewritten by the next iteration
  }
}

}  // namespace lkm

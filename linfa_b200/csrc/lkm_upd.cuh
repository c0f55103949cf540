// Centroid-update pass of the two-pass Lloyd iteration for wide rows / many clusters (f32, d = 32, 64 or 128):
// given the labels of the assignment pass, per-CTA partial sums, counts and inertia for finalize_kernel
// (compute_centroids, KM/algorithm.rs:785-810, and dists.sum(), :329).  HBM-bound: X is streamed once per
// cluster chunk through a TMA ring; the accumulators of the whole chunk (chunk_rows x d f32) stay in shared memory.
//
// Deterministic without atomics: every accumulator element has ONE owning thread.  A 32-row sub-tile is handled
// by 16 consumer warps; warp w = (class, feature slice) owns the clusters c with c % NCLS == class and the 32
// features of its slice (lane = feature), walks the sub-tiles of the CTA in ascending order and, per sub-tile,
// sums the rows of each of its clusters in registers (row order) before ONE read-modify-write of the shared
// accumulator.  Blob-ordered data gives one segment of 32 rows per sub-tile; shuffled data gives many one-row
// segments, taken four at a time so that their centroid loads and read-modify-writes overlap.
//
//   P  warp 16      one thread keeps the ring full: one TMA box (32 full rows, row-major) and the
//                   128 label bytes of the sub-tile per stage
//   R  warp 17      waits for the stage's `full` mbarrier and hands the stage to the consumers through a named
//                   barrier: the 16 consumer warps block in the barrier unit instead of polling the mbarrier
//                   (with all of them polling, two thirds of the executed instructions were wait loops)
//   C  warps 0-15   as above; the stage is released when all 16 warps have arrived on its `empty` barrier
#pragma once
#include "lkm_tma.cuh"

namespace lkm {

constexpr int kUpdConsumers = 16;
constexpr int kUpdThreads = (kUpdConsumers + 2) * 32;
constexpr int kUpdSubRows = 32;
constexpr int kUpdMaxStages = 8;
constexpr unsigned long long kUpdPrefetch = 12;   // L2 prefetch distance in sub-tiles beyond the ring's newest request

struct UpdArgs {
  const float* X;              // for L2 prefetches (the loads themselves go through the tensor map)
  unsigned long long n;
  int k;
  int chunk_rows;              // clusters per chunk; blockIdx.y = chunk, chunk c owns [c * chunk_rows, ...)
  int stages;
  const uint32_t* labels;      // n labels; the allocation is readable up to the next multiple of 128 rows
  const float* C;              // k x d, the centroids the labels were computed against (inertia)
  float* part_sums;            // [gridDim.x][k * d]
  uint32_t* part_counts;       // [gridDim.x][k]
  unsigned char* part_touched; // [gridDim.x][k] or nullptr: with it only the rows of clusters this CTA met are written back
  double* part_inertia;        // [gridDim.y][gridDim.x]
  const LloydState* state;
  // list mode (behind the fused screening kernel): CTA b only visits the 128-row tiles tile_list[b * list_cap + i],
  // i < tile_count[b], in that order
  const uint32_t* tile_list;
  const uint32_t* tile_count;
  int list_cap;
};

__host__ __device__ inline int upd_stage_rows(int d) { return kUpdSubRows * (128 / d); }   // 16 KB of rows per TMA operation
__host__ __device__ inline size_t upd_stage_bytes(int d) { return (size_t)upd_stage_rows(d) * d * 4 + 512; }
__host__ __device__ inline size_t upd_smem_bytes(int d, int chunk_rows, int stages) {
  return (size_t)stages * upd_stage_bytes(d) + (size_t)chunk_rows * d * 4 + (size_t)chunk_rows * 4 + kUpdThreads * 8 + 256;
}

namespace tc {
// 1-D bulk copy global -> shared, completion on an mbarrier (bytes and both addresses multiples of 16)
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
}  // namespace tc

// one or two consecutive features per lane
template <int V> struct UpdVec;
template <> struct UpdVec<1> { using type = float; };
template <> struct UpdVec<2> { using type = float2; };
__device__ __forceinline__ float vadd(float a, float b) { return a + b; }
__device__ __forceinline__ float vsqacc(float x, float c, float e) { const float d = x - c; return fmaf(d, d, e); }
__device__ __forceinline__ float2 vadd(float2 a, float2 b) { return make_float2(a.x + b.x, a.y + b.y); }
__device__ __forceinline__ float vsqacc(float2 x, float2 c, float e) {
  const float d0 = x.x - c.x, d1 = x.y - c.y;
  return fmaf(d1, d1, fmaf(d0, d0, e));
}
template <typename Vf> __device__ __forceinline__ Vf upd_zero();
template <> __device__ __forceinline__ float upd_zero<float>() { return 0.f; }
template <> __device__ __forceinline__ float2 upd_zero<float2>() { return make_float2(0.f, 0.f); }
template <typename Vf> __device__ __forceinline__ Vf upd_ld(const float* p) { return *reinterpret_cast<const Vf*>(p); }
template <typename Vf> __device__ __forceinline__ Vf upd_ldg(const float* p) { return __ldg(reinterpret_cast<const Vf*>(p)); }
template <typename Vf> __device__ __forceinline__ void upd_st(float* p, Vf v) { *reinterpret_cast<Vf*>(p) = v; }

template <int D>
__global__ void __launch_bounds__(kUpdThreads, 1) update_rows_kernel(const UpdArgs a, const __grid_constant__ CUtensorMap tmap) {
  static_assert(D == 32 || D == 64 || D == 128, "row width");
  constexpr int V = D >= 64 ? 2 : 1;         // features per lane
  constexpr int QF = D / (32 * V);           // feature slices
  constexpr int NCLS = kUpdConsumers / QF;   // cluster classes
  using Vf = typename UpdVec<V>::type;
  constexpr int SUBS = 128 / D;                            // 32-row sub-tiles per stage (a TMA operation moves 16 KB:
                                                           // smaller ones are bound by the per-operation cost of the copy engine)
  constexpr int kStageRows = kUpdSubRows * SUBS;
  constexpr uint32_t kStageX = (uint32_t)kStageRows * D * 4;
  constexpr uint32_t kStage = kStageX + 512;
  if (a.state != nullptr && a.state->done) return;
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((128u - (tc::smem_u32(smem_unaligned) & 127u)) & 127u);
  const int t = threadIdx.x, lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int R = a.stages;
  const int chunk_base = (int)blockIdx.y * a.chunk_rows;
  const int chunk_rows = min(a.chunk_rows, a.k - chunk_base);
  unsigned char* ring = smem;                                                // [R][32 x D f32 | 32 labels]
  float* acc = reinterpret_cast<float*>(ring + (size_t)R * kStage);          // [chunk_rows][D]
  uint32_t* cnt = reinterpret_cast<uint32_t*>(acc + (size_t)a.chunk_rows * D);
  double* red = reinterpret_cast<double*>(cnt + ((a.chunk_rows + 1) & ~1));
  __shared__ __align__(8) uint64_t bar_full[kUpdMaxStages], bar_empty[kUpdMaxStages];

  // contiguous range of sub-tiles per CTA column (ascending row order inside every accumulator)
  const unsigned long long n_sub = (a.n + kStageRows - 1) / kStageRows;   // stages
  const unsigned long long per = (n_sub + gridDim.x - 1) / gridDim.x;
  constexpr int SPT = 128 / kStageRows;                                    // stages per listed tile
  const bool listed = a.tile_list != nullptr;
  const uint32_t* my_list = listed ? a.tile_list + (size_t)blockIdx.x * a.list_cap : nullptr;
  const unsigned long long j0 = listed ? 0ull : min(n_sub, per * blockIdx.x);
  const unsigned long long j1 = listed ? (unsigned long long)a.tile_count[blockIdx.x] * SPT : min(n_sub, j0 + per);
  // stage index (units of kStageRows rows) of loop position j
  auto stage_of = [&](unsigned long long j) -> unsigned long long {
    return listed ? (unsigned long long)my_list[j / SPT] * SPT + (j % SPT) : j;
  };

  if (listed && j1 == 0) {
    // nothing listed for this CTA (the rule behind the fused pair kernel on ordered rows): publish an empty partial
    uint32_t* pc0 = a.part_counts + (size_t)blockIdx.x * a.k + chunk_base;
    for (int e = t; e < chunk_rows; e += kUpdThreads) {
      pc0[e] = 0u;
      if (a.part_touched != nullptr) a.part_touched[(size_t)blockIdx.x * a.k + chunk_base + e] = 0;
      else for (int j = 0; j < D; ++j) a.part_sums[(size_t)blockIdx.x * a.k * D + (size_t)(chunk_base + e) * D + j] = 0.f;
    }
    if (t == 0) a.part_inertia[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = 0.0;
    return;
  }
  if (t == 0) {
    for (int s = 0; s < kUpdMaxStages; ++s) {
      tc::mbar_init(&bar_full[s], 1);
      tc::mbar_init(&bar_empty[s], kUpdConsumers);
    }
    tc::fence_mbar_init();
  }
  for (int e = t; e < chunk_rows * D; e += kUpdThreads) acc[e] = 0.f;
  for (int e = t; e < chunk_rows; e += kUpdThreads) cnt[e] = 0u;
  __syncthreads();

  double inert = 0.0;
  if (warp == kUpdConsumers) {
    if (lane == 0) {
      int s = 0;
      uint32_t use = 0;
      for (unsigned long long j = j0; j < j1; ++j) {
        if (use > 0) tc::mbar_wait_spin(&bar_empty[s], (use - 1u) & 1u);
        unsigned char* st = ring + (size_t)s * kStage;
        tc::mbar_expect_tx(&bar_full[s], kStageX + 128u * SUBS);
        // one box of 32 full rows: the TMA unit works through a box one inner line at a time (measured ~8 cycles per
        // line whatever its length), so 512-byte lines move four times as many bytes per cycle as 128-byte lines
        const unsigned long long sj = stage_of(j);
        tc::tma_load_2d(st, &tmap, 0, (int)(sj * kStageRows), &bar_full[s]);
        if (!listed && j + kUpdPrefetch < j1) {
          const unsigned long long r0 = (j + kUpdPrefetch) * kStageRows;
          tc::l2_prefetch(a.X + r0 * (unsigned long long)D, (uint32_t)min((unsigned long long)kStageRows, a.n - r0) * (uint32_t)(D * 4));
        }
        tc::bulk_load_1d(st + kStageX, a.labels + sj * kStageRows, 128u * SUBS, &bar_full[s]);
        if (++s == R) { s = 0; ++use; }
      }
    }
  } else if (warp == kUpdConsumers + 1) {
    int s = 0;
    uint32_t ph = 0;
    for (unsigned long long j = j0; j < j1; ++j) {
      tc::mbar_wait_spin(&bar_full[s], ph);
      tc::nb_arrive(1 + s, (kUpdConsumers + 1) * 32);
      if (++s == R) { s = 0; ph ^= 1u; }
    }
  } else {
    // lane = V consecutive features (V = 2 for D >= 64: 64-bit shared-memory accesses, half the instructions per row)
    const int fs = warp % QF, cls = warp / QF;
    const int fo = (fs * 32 + lane) * V;       // first feature owned by this lane
    const float* Cf = a.C + (size_t)chunk_base * D + fo;
    float* accf = acc + fo;
    int s = 0;
    uint32_t last_L = 0xFFFFFFFFu;
    Vf last_cv = upd_zero<Vf>();
    const uint32_t lt_mask = (1u << lane) - 1u;
    Vf pend_x[4], pend_c[4];
    uint32_t pend_on = 0u;
#pragma unroll
    for (int q = 0; q < 4; ++q) { pend_x[q] = upd_zero<Vf>(); pend_c[q] = upd_zero<Vf>(); }
    for (unsigned long long j = j0; j < j1; ++j) {
      const unsigned char* st = ring + (size_t)s * kStage;
      const unsigned long long sj = stage_of(j);
      tc::nb_sync(1 + s, (kUpdConsumers + 1) * 32);
#pragma unroll 1
      for (int sub = 0; sub < SUBS; ++sub) {
      const float* xs = reinterpret_cast<const float*>(st) + (size_t)sub * kUpdSubRows * D + fo;     // row r at xs[r * D]
      uint32_t lab = reinterpret_cast<const uint32_t*>(st + kStageX)[sub * kUpdSubRows + lane];
      if (sj * kStageRows + (unsigned)(sub * kUpdSubRows + lane) >= a.n) lab = 0xFFFFFFFFu;
      const uint32_t c = lab - (uint32_t)chunk_base;
      const bool mine = c < (uint32_t)chunk_rows && (int)(c % (uint32_t)NCLS) == cls;
      uint32_t m = __ballot_sync(0xffffffffu, mine);
      if (m == 0u) continue;
      if (m == 0xffffffffu && __all_sync(0xffffffffu, c == __shfl_sync(0xffffffffu, c, 0))) {
        // the whole sub-tile belongs to one cluster of this warp (the rule for ordered rows): no grouping needed; two
        // interleaved chains (even / odd rows), fixed order
        const uint32_t L0 = __shfl_sync(0xffffffffu, c, 0);
        if (L0 != last_L) { last_cv = upd_ldg<Vf>(Cf + (size_t)L0 * D); last_L = L0; }
        const Vf a0 = upd_ld<Vf>(accf + (size_t)L0 * D);
        Vf s0 = upd_zero<Vf>(), s1 = upd_zero<Vf>();
        float e0 = 0.f, e1 = 0.f;
#pragma unroll
        for (int r = 0; r < kUpdSubRows; r += 2) {
          const Vf x0 = upd_ld<Vf>(xs + r * D), x1 = upd_ld<Vf>(xs + (r + 1) * D);
          s0 = vadd(s0, x0); s1 = vadd(s1, x1);
          e0 = vsqacc(x0, last_cv, e0); e1 = vsqacc(x1, last_cv, e1);
        }
        upd_st<Vf>(accf + (size_t)L0 * D, vadd(a0, vadd(s0, s1)));
        inert += (double)(e0 + e1);
        if (fs == 0 && lane == 0) cnt[L0] += (uint32_t)kUpdSubRows;
        continue;
      }
      // rows of this warp with the same cluster; rank = position of the row among them
      // (one key for all the other lanes: MATCH.ANY walks the distinct keys one by one)
      const uint32_t peers = __match_any_sync(0xffffffffu, mine ? c : 0xFFFFFFFFu);
      const int rank = __popc(peers & lt_mask);
      const uint32_t leaders = __ballot_sync(0xffffffffu, mine && rank == 0);
      if (__popc(leaders) > 2) {
        // Unordered rows: many clusters, one or two rows each.  Round t takes the t-th row of every cluster, so the rows
        // of a round touch distinct accumulators and their loads / read-modify-writes overlap (four in flight); a
        // cluster's rows still reach its accumulator in row order.  The centroid row a row needs for the inertia comes
        // from L2: its use is DEFERRED by one batch (the row and the centroid slice wait in registers), so that the
        // round trip overlaps the next batch instead of stalling this one.
        float e_sub = 0.f;
        for (int round = 0;; ++round) {
          uint32_t p = __ballot_sync(0xffffffffu, mine && rank == round);
          if (p == 0u) break;
          while (p) {
            int r[4];
            uint32_t L[4];
            bool on[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              on[q] = p != 0u;
              r[q] = on[q] ? __ffs(p) - 1 : 0;
              p &= p - 1u;
              L[q] = __shfl_sync(0xffffffffu, c, r[q]);
            }
            Vf x[4], av[4], cv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (on[q]) {
                cv[q] = upd_ldg<Vf>(Cf + (size_t)L[q] * D);
                x[q] = upd_ld<Vf>(xs + r[q] * D);
                av[q] = upd_ld<Vf>(accf + (size_t)L[q] * D);
              }
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (on[q]) {
                upd_st<Vf>(accf + (size_t)L[q] * D, vadd(av[q], x[q]));
                if (fs == 0 && lane == 0) cnt[L[q]] += 1u;
              }
            }
            // the previous batch's centroid slices have had a whole batch to arrive
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              if (pend_on & (1u << q)) e_sub = vsqacc(pend_x[q], pend_c[q], e_sub);
              pend_x[q] = x[q];
              pend_c[q] = cv[q];
            }
            pend_on = (on[0] ? 1u : 0u) | (on[1] ? 2u : 0u) | (on[2] ? 4u : 0u) | (on[3] ? 8u : 0u);
          }
        }
        inert += (double)e_sub;
        continue;
      }
      while (m) {
        // up to four segments (clusters of this warp present in the sub-tile) at a time
        uint32_t pm[4], L[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          pm[q] = 0u; L[q] = 0u;
          if (m) {
            L[q] = __shfl_sync(0xffffffffu, c, __ffs(m) - 1);
            pm[q] = __ballot_sync(0xffffffffu, mine && c == L[q]);
            m &= ~pm[q];
          }
        }
        Vf cv[4], av[4], sv[4];
        float ev[4];
        // the centroid coordinate comes from L2 (hundreds of cycles under load) and sits in front of the first
        // subtraction: keep the last one, which is the one needed again while the rows of a cluster are contiguous
        cv[0] = last_cv; av[0] = upd_zero<Vf>();
        if (pm[0]) {
          if (L[0] != last_L) { cv[0] = upd_ldg<Vf>(Cf + (size_t)L[0] * D); last_L = L[0]; last_cv = cv[0]; }
          av[0] = upd_ld<Vf>(accf + (size_t)L[0] * D);
        }
#pragma unroll
        for (int q = 1; q < 4; ++q) {
          cv[q] = upd_zero<Vf>(); av[q] = upd_zero<Vf>();
          if (pm[q]) { cv[q] = upd_ldg<Vf>(Cf + (size_t)L[q] * D); av[q] = upd_ld<Vf>(accf + (size_t)L[q] * D); }
        }
        if (pm[0] == 0xffffffffu) {
          // the whole sub-tile belongs to one cluster: two interleaved chains (even / odd rows), fixed order
          Vf s0 = upd_zero<Vf>(), s1 = upd_zero<Vf>();
          float e0 = 0.f, e1 = 0.f;
#pragma unroll
          for (int r = 0; r < kUpdSubRows; r += 2) {
            const Vf x0 = upd_ld<Vf>(xs + r * D), x1 = upd_ld<Vf>(xs + (r + 1) * D);
            s0 = vadd(s0, x0); s1 = vadd(s1, x1);
            e0 = vsqacc(x0, cv[0], e0); e1 = vsqacc(x1, cv[0], e1);
          }
          sv[0] = vadd(s0, s1); ev[0] = e0 + e1;
          sv[1] = sv[2] = sv[3] = upd_zero<Vf>(); ev[1] = ev[2] = ev[3] = 0.f;
        } else {
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            Vf sq = upd_zero<Vf>();
            float eq = 0.f;
            uint32_t p = pm[q];
            while (p) {
              const int r = __ffs(p) - 1;
              p &= p - 1u;
              const Vf x = upd_ld<Vf>(xs + r * D);
              sq = vadd(sq, x);
              eq = vsqacc(x, cv[q], eq);
            }
            sv[q] = sq; ev[q] = eq;
          }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          if (pm[q]) {
            upd_st<Vf>(accf + (size_t)L[q] * D, vadd(av[q], sv[q]));
            inert += (double)ev[q];
            if (fs == 0 && lane == 0) cnt[L[q]] += (uint32_t)__popc(pm[q]);
          }
        }
      }
      }
      __syncwarp();
      if (lane == 0) tc::mbar_arrive(&bar_empty[s]);
      if (++s == R) s = 0;
    }
    {
      float e_tail = 0.f;
#pragma unroll
      for (int q = 0; q < 4; ++q)
        if (pend_on & (1u << q)) e_tail = vsqacc(pend_x[q], pend_c[q], e_tail);
      inert += (double)e_tail;
    }
  }
  // ---- per-CTA partials: accumulators as they are, inertia by warp butterflies and then in warp order ----
  for (int o = 16; o > 0; o >>= 1) inert += __shfl_xor_sync(0xffffffffu, inert, o);
  if (lane == 0) red[warp] = inert;
  __syncthreads();
  const size_t kd = (size_t)a.k * D;
  float* ps = a.part_sums + (size_t)blockIdx.x * kd + (size_t)chunk_base * D;
  for (int e = t; e < chunk_rows * D / 4; e += kUpdThreads)
    if (a.part_touched == nullptr || cnt[e / (D / 4)] != 0u) reinterpret_cast<float4*>(ps)[e] = reinterpret_cast<const float4*>(acc)[e];
  uint32_t* pc = a.part_counts + (size_t)blockIdx.x * a.k + chunk_base;
  for (int e = t; e < chunk_rows; e += kUpdThreads) pc[e] = cnt[e];
  if (a.part_touched != nullptr)
    for (int e = t; e < chunk_rows; e += kUpdThreads) a.part_touched[(size_t)blockIdx.x * a.k + chunk_base + e] = cnt[e] != 0u ? 1 : 0;
  if (t == 0) {
    double tot = 0.0;
    for (int e = 0; e < kUpdConsumers; ++e) tot += red[e];
    a.part_inertia[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = tot;
  }
}

}  // namespace lkm

// Centroid update for rows in NO particular order (f32, d = 32 / 64 / 128): compute_centroids (KM/algorithm.rs:785-810)
// and dists.sum() (:329) from the labels of the assignment pass, when neighbouring rows rarely share a cluster.
// update_rows_kernel streams X in row order and pays per row for finding the owner of its accumulator (deterministic
// without atomics = one owning thread per accumulator element); on hashed rows that bookkeeping, not HBM, is the bound.
// Here the rows are first put in (label, row) order by a stable counting sort of the 4-byte labels — 12 bytes of
// traffic per row against d * 4 for the row itself — and the update is then a segmented sum over a gather of whole
// rows (every row is 128 .. 512 contiguous bytes, so the gather runs at streaming efficiency):
//
//   sort_hist_kernel     a warp per segment of 2048 rows: label histogram in warp-private shared memory
//   sort_scan_kernel     per label the exclusive prefix over the segments (in place); the last CTA scans the k totals
//                        into seg_start[0..k]
//   sort_scatter_kernel  the same warps walk their segment again: position = seg_start[label] + prefix[segment][label]
//                        + rank among the equal labels met so far (MATCH.ANY) — stable, hence reproducible
//   update_sorted_kernel warp w owns positions [w * per_warp, ..) of the sorted order, sums the rows of one cluster in
//                        registers (lane = 1 / 2 / 4 features, 8 row loads in flight) and closes one ENTRY per
//                        (warp, cluster) it met: entry index = w + c, unique because both only grow along the order
//   sorted_combine_kernel one CTA per cluster adds the entries of the warps that met it, in warp order, into ONE
//                        partial for the usual finalize kernels (m_k-means update, NVLink exchange, stop rule)
//
// Every sum has a fixed order: same input, same bits.  Counts are the segment lengths (exact).
#pragma once
#include "lkm_upd.cuh"

namespace lkm {

constexpr int kSrtSegRows = 2048;   // rows per warp segment of the counting sort
constexpr int kSrtWarps = 8;
constexpr int kSrtThreads = kSrtWarps * 32;
constexpr int kSrtMaxK = 2048;      // kSrtWarps * k counters of 4 bytes in shared memory
constexpr int kSuWarps = 16;
constexpr int kSuThreads = kSuWarps * 32;

struct SortArgs {
  const uint32_t* labels;
  unsigned long long n;
  int k;
  uint32_t* hist;            // [n_seg][k]: counts, after sort_scan_kernel the exclusive prefix over the segments
  uint32_t* seg_start;       // [k + 1]
  uint32_t* perm;            // [n]: rows in (label, row) order
  unsigned int* scan_done;   // zero before the launch; the last CTA of sort_scan_kernel resets it
  const LloydState* state;
};

// adjacent rows with different labels (decides between the streaming and the sorted update)
static __global__ void label_breaks_kernel(const uint32_t* __restrict__ labels, unsigned long long n, LloydState* state) {
  if (state->done) return;
  unsigned int mine = 0;
  for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x + 1; i < n;
       i += (unsigned long long)gridDim.x * blockDim.x)
    mine += labels[i] != labels[i - 1];
  for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
  if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&state->label_breaks, (unsigned long long)mine);   // integer: exact
}

static __global__ void __launch_bounds__(kSrtThreads) sort_hist_kernel(const SortArgs a) {
  if (a.state != nullptr && a.state->done) return;
  extern __shared__ uint32_t srt_ctr[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* ctr = srt_ctr + (size_t)warp * a.k;
  const unsigned long long n_seg = (a.n + kSrtSegRows - 1) / kSrtSegRows;
  const unsigned long long g = (unsigned long long)blockIdx.x * kSrtWarps + warp;
  if (g >= n_seg) return;
  for (int c = lane; c < a.k; c += 32) ctr[c] = 0u;
  __syncwarp();
  const unsigned long long r0 = g * kSrtSegRows, r1 = min(a.n, r0 + kSrtSegRows);
  for (unsigned long long r = r0 + lane; r < r1; r += 32) atomicAdd(&ctr[a.labels[r]], 1u);
  __syncwarp();
  for (int c = lane; c < a.k; c += 32) a.hist[g * a.k + c] = ctr[c];
}

// block (32, 32): x = label within the CTA's group of 32 (coalesced 128-byte accesses), y = slice of the segments; every
// thread walks its slice in batches of 8 independent loads (one load per step is a chain of L2 latencies: 52 us for the
// 4883 x 256 counters of cfg3; column-wise layouts — a warp per label, or 4 labels per CTA — were slower still, 90 - 122 us).
// The last CTA turns the k totals into seg_start.
static __global__ void __launch_bounds__(1024) sort_scan_kernel(const SortArgs a) {
  if (a.state != nullptr && a.state->done) return;
  __shared__ uint32_t part[32][33];
  __shared__ uint32_t wsum[32];
  __shared__ bool last;
  const int x = threadIdx.x, y = threadIdx.y;
  const int c = blockIdx.x * 32 + x;
  const unsigned long long n_seg = (a.n + kSrtSegRows - 1) / kSrtSegRows;
  const unsigned long long per = (n_seg + 31) / 32;
  const unsigned long long g0 = min(n_seg, per * y), g1 = min(n_seg, g0 + per);
  constexpr int B = 8;
  uint32_t s = 0;
  if (c < a.k) {
    for (unsigned long long g = g0; g < g1; g += B) {
      uint32_t h[B];
#pragma unroll
      for (int i = 0; i < B; ++i) h[i] = g + i < g1 ? a.hist[(g + i) * a.k + c] : 0u;
#pragma unroll
      for (int i = 0; i < B; ++i) s += h[i];
    }
  }
  part[y][x] = s;
  __syncthreads();
  uint32_t run = 0;
  for (int yy = 0; yy < y; ++yy) run += part[yy][x];
  if (c < a.k) {
    for (unsigned long long g = g0; g < g1; g += B) {
      uint32_t h[B];
#pragma unroll
      for (int i = 0; i < B; ++i) h[i] = g + i < g1 ? a.hist[(g + i) * a.k + c] : 0u;
#pragma unroll
      for (int i = 0; i < B; ++i) {
        if (g + i < g1) a.hist[(g + i) * a.k + c] = run;
        run += h[i];
      }
    }
    if (y == 31) a.seg_start[c + 1] = run;   // the label's total, until the last CTA turns the totals into starts
  }
  __threadfence();
  __syncthreads();
  const int t = y * 32 + x;
  if (t == 0) last = atomicAdd(a.scan_done, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  // inclusive scan of the k <= 2048 totals, two per thread
  const uint32_t v0 = 2 * t < a.k ? __ldcg(a.seg_start + 2 * t + 1) : 0u;
  const uint32_t v1 = 2 * t + 1 < a.k ? __ldcg(a.seg_start + 2 * t + 2) : 0u;
  uint32_t inc = v0 + v1;
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t up = __shfl_up_sync(0xffffffffu, inc, o);
    if (x >= o) inc += up;
  }
  if (x == 31) wsum[y] = inc;
  __syncthreads();
  uint32_t base = 0;
  for (int yy = 0; yy < y; ++yy) base += wsum[yy];
  const uint32_t excl = base + inc - (v0 + v1);
  if (2 * t < a.k) a.seg_start[2 * t + 1] = excl + v0;
  if (2 * t + 1 < a.k) a.seg_start[2 * t + 2] = excl + v0 + v1;
  if (t == 0) {
    a.seg_start[0] = 0u;
    *a.scan_done = 0u;
  }
}

static __global__ void __launch_bounds__(kSrtThreads) sort_scatter_kernel(const SortArgs a) {
  if (a.state != nullptr && a.state->done) return;
  extern __shared__ uint32_t srt_ctr[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint32_t* ctr = srt_ctr + (size_t)warp * a.k;
  const unsigned long long n_seg = (a.n + kSrtSegRows - 1) / kSrtSegRows;
  const unsigned long long g = (unsigned long long)blockIdx.x * kSrtWarps + warp;
  if (g >= n_seg) return;
  for (int c = lane; c < a.k; c += 32) ctr[c] = a.seg_start[c] + a.hist[g * a.k + c];
  __syncwarp();
  const unsigned long long r0 = g * kSrtSegRows, r1 = min(a.n, r0 + kSrtSegRows);
  const unsigned lt = (1u << lane) - 1u;
  for (unsigned long long rb = r0; rb < r1; rb += 32) {
    const unsigned long long r = rb + lane;
    const bool on = r < r1;
    const uint32_t lab = on ? a.labels[r] : 0xffffffffu;
    const unsigned peers = __match_any_sync(0xffffffffu, lab);
    if (on) a.perm[ctr[lab] + __popc(peers & lt)] = (uint32_t)r;
    __syncwarp();
    if (on && (peers & lt) == 0u) ctr[lab] += (uint32_t)__popc(peers);
    __syncwarp();
  }
}

struct SortUpdArgs {
  const float* X;
  unsigned long long n;
  int k;
  const uint32_t* perm;
  const uint32_t* seg_start;
  const float* C;                // the centroids the labels were computed against (inertia)
  float* ent_sums;               // [n_warps + k][D]
  double* ent_inertia;           // [n_warps]
  unsigned long long per_warp;   // positions per warp
  unsigned long long n_warps;
  // outputs of sorted_combine_kernel: one partial source
  float* part_sums;              // [k * D]
  uint32_t* part_counts;         // [k]
  unsigned char* part_touched;   // [k] or nullptr
  double* part_inertia;          // [1]
  const LloydState* state;
};

template <int V> struct SuVec;
template <> struct SuVec<1> { using type = float; };
template <> struct SuVec<2> { using type = float2; };
template <> struct SuVec<4> { using type = float4; };
__device__ __forceinline__ float4 vadd(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float vsqacc(float4 x, float4 c, float e) {
  const float d0 = x.x - c.x, d1 = x.y - c.y, d2 = x.z - c.z, d3 = x.w - c.w;
  return fmaf(d3, d3, fmaf(d2, d2, fmaf(d1, d1, fmaf(d0, d0, e))));
}
template <> __device__ __forceinline__ float4 upd_zero<float4>() { return make_float4(0.f, 0.f, 0.f, 0.f); }

template <int D>
__global__ void __launch_bounds__(kSuThreads, 2) update_sorted_kernel(const SortUpdArgs a) {
  static_assert(D == 32 || D == 64 || D == 128, "row width");
  constexpr int V = D / 32;
  using Vf = typename SuVec<V>::type;
  constexpr int U = D == 128 ? 4 : 8;   // row loads in flight per warp (2 KB / 2 KB / 1 KB)
  if (a.state != nullptr && a.state->done) return;
  const int lane = threadIdx.x & 31;
  const unsigned long long w = (unsigned long long)blockIdx.x * kSuWarps + (threadIdx.x >> 5);
  if (w >= a.n_warps) return;
  const unsigned long long p0 = min(a.n, w * a.per_warp), p1 = min(a.n, p0 + a.per_warp);
  double inert = 0.0;
  if (p0 < p1) {
    int lo = 0, hi = a.k;   // seg_start[lo] <= p0 < seg_start[hi]
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (a.seg_start[mid] <= p0) lo = mid; else hi = mid;
    }
    int c = lo;
    unsigned long long seg_end = a.seg_start[c + 1];
    Vf cen = upd_ldg<Vf>(a.C + (size_t)c * D + lane * V);
    Vf acc = upd_zero<Vf>();
    for (unsigned long long p = p0; p < p1; p += 32) {
      const unsigned long long q = p + lane;
      const uint32_t idx = q < p1 ? a.perm[q] : 0u;
      const int m = (int)min(32ull, p1 - p);
      Vf bsum = upd_zero<Vf>();
      float e = 0.f;
      for (int j0 = 0; j0 < m; j0 += U) {
        Vf x[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t r = __shfl_sync(0xffffffffu, idx, (j0 + u) & 31);
          x[u] = upd_ldg<Vf>(a.X + (size_t)r * D + lane * V);   // lanes past m hold row 0: loaded, never added
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
          if (j0 + u < m) {
            const unsigned long long pos = p + (unsigned)(j0 + u);
            if (pos >= seg_end) {   // the cluster's last row is behind us: close its entry
              acc = vadd(acc, bsum);
              bsum = upd_zero<Vf>();
              upd_st<Vf>(a.ent_sums + (size_t)(w + (unsigned)c) * D + lane * V, acc);
              acc = upd_zero<Vf>();
              do { ++c; seg_end = a.seg_start[c + 1]; } while (pos >= seg_end);
              cen = upd_ldg<Vf>(a.C + (size_t)c * D + lane * V);
            }
            bsum = vadd(bsum, x[u]);
            e = vsqacc(x[u], cen, e);
          }
        }
      }
      acc = vadd(acc, bsum);
      inert += (double)e;
    }
    upd_st<Vf>(a.ent_sums + (size_t)(w + (unsigned)c) * D + lane * V, acc);
  }
  for (int o = 16; o > 0; o >>= 1) inert += __shfl_xor_sync(0xffffffffu, inert, o);
  if (lane == 0) a.ent_inertia[w] = inert;
}

// grid k + 1, D threads: CTA c < k combines cluster c, CTA k the inertia
template <int D>
__global__ void __launch_bounds__(D) sorted_combine_kernel(const SortUpdArgs a) {
  if (a.state != nullptr && a.state->done) return;
  const int c = blockIdx.x, t = threadIdx.x;
  if (c < a.k) {
    const unsigned long long s0 = a.seg_start[c], s1 = a.seg_start[c + 1];
    if (s1 == s0) {
      if (t == 0) a.part_counts[c] = 0u;
      if (a.part_touched != nullptr) { if (t == 0) a.part_touched[c] = 0; }
      else a.part_sums[(size_t)c * D + t] = 0.f;
      return;
    }
    const unsigned long long w_lo = s0 / a.per_warp, w_hi = (s1 - 1) / a.per_warp;
    double s = 0.0;
    for (unsigned long long w = w_lo; w <= w_hi; ++w) s += (double)a.ent_sums[(size_t)(w + (unsigned)c) * D + t];
    a.part_sums[(size_t)c * D + t] = (float)s;
    if (t == 0) {
      a.part_counts[c] = (uint32_t)(s1 - s0);
      if (a.part_touched != nullptr) a.part_touched[c] = 1;
    }
  } else {
    __shared__ double red[D];
    double s = 0.0;
    for (unsigned long long w = t; w < a.n_warps; w += D) s += a.ent_inertia[w];
    red[t] = s;
    __syncthreads();
    for (int o = D / 2; o > 0; o >>= 1) {
      if (t < o) red[t] += red[t + o];
      __syncthreads();
    }
    if (t == 0) a.part_inertia[0] = red[0];
  }
}

}  // namespace lkm

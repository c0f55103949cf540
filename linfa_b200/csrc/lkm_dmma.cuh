// f64 assignment on the FP64 tensor pipe (DMMA): update_memberships_and_dists (KM/algorithm.rs:866-881) for f64 data whose
// contraction is too large for the small-d kernel of lkm_f64s.cuh (north_star: "DMMA for f64" in the large k*d regime).
//
// Scores w_c = ||c||^2 - 2 x.c with the x.c products on `mma.sync.aligned.m8n8k4.row.col.f64` (DMMA in SASS): one
// instruction = 8 rows x 8 centroids x 4 features = 256 FMAs, against 32 for a warp-wide DFMA — the FP64 rate of the SM is
// the same, but the scan leaves the issue slots, the register ports and the shared-memory port almost idle (one LDS.64 per
// 16 DMMA per operand), and the row tile is read from HBM once, coalesced.
//
//   CTA = 8 warps, tile = 128 rows in shared memory (row stride = 4 mod 16 doubles: the fragment loads — lane (g, t) reads
//   element [g][t] of an 8 x 4 block — are conflict-free per half warp); centroids in chunks of 64 with the same stride.
//   Warp w owns rows [16 w, 16 w + 16) x the 64 centroids of the chunk: 2 x 8 accumulator fragments (32 doubles per lane).
//   Epilogue per chunk: every lane folds its 2 x 16 scores into a running (best, runner-up, index) per row; after the last
//   chunk the four lanes that share a row combine theirs by shuffles.
//
// Certificate (same bound as scan_centroids_certified, lkm_kernels.cuh — it holds for any summation order of the dot
// product): runner-up - best > 1.1 eps [(4d+4)|x|C + 4C^2 + 2(d+2)(|x|+C)^2] proves that the reference's direct-form
// strict-'<' scan picks the same centroid.  Other rows go to the list of rows_top2_kernel (lkm_rows.cuh).  The winner's
// distance is evaluated in the reference's direct form (bit-identical `dists`).
#pragma once
#include "lkm_kernels.cuh"

namespace lkm {

struct DmmaArgs {
  const double* X;
  unsigned long long n;
  int d, k;
  const double* C;
  int xs;                       // shared-memory row stride in doubles (>= d, = 4 mod 16)
  uint32_t* labels;
  double* dists;                // min squared distance per row (direct form)
  uint32_t* rescan_list;
  unsigned int* rescan_count;
  const LloydState* state;
  LloydState* state_rw;         // fallback_rows counter
};

constexpr int kDmThreads = 256;
constexpr int kDmRows = 128;    // rows per tile
constexpr int kDmKc = 64;       // centroids per chunk

__host__ __device__ inline int dmma_row_stride(int d) { return ((d - 4 + 15) / 16) * 16 + 4; }   // smallest >= d that is 4 mod 16 (d >= 4)
__host__ __device__ inline size_t dmma_smem_bytes(int d) {
  const size_t xs = (size_t)dmma_row_stride(d);
  return ((size_t)kDmRows * xs + (size_t)kDmKc * xs + kDmKc + kDmRows) * 8;
}

__device__ __forceinline__ void dmma_884(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void __launch_bounds__(kDmThreads) assign_dmma_kernel(const DmmaArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  if (a.state != nullptr && a.state->done) return;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int d = a.d, xs = a.xs, t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int g = lane >> 2, tq = lane & 3;
  double* xt = reinterpret_cast<double*>(smem_raw);       // [128][xs]
  double* ct = xt + (size_t)kDmRows * xs;                  // [64][xs]
  double* cn = ct + (size_t)kDmKc * xs;                    // [64] squared norms of the chunk
  double* xn2 = cn + kDmKc;                                // [128] squared norms of the rows
  const unsigned long long n_tiles = (a.n + kDmRows - 1) / kDmRows;
  const int n_chunks = (a.k + kDmKc - 1) / kDmKc;
  const double eps = 1.1102230246251565e-16, dmin = 4.9406564584124654e-324;
  unsigned int my_fallbacks = 0;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long row0 = tile * kDmRows;
    const int rows = (int)min((unsigned long long)kDmRows, a.n - row0);
    __syncthreads();
    // tile: coalesced 16-byte loads, rows past the end as zeros
    {
      const int vpr = d / 2;   // double2 per row
      for (int e = t; e < kDmRows * vpr; e += kDmThreads) {
        const int r = e / vpr, c2 = e - r * vpr;
        double2 v = make_double2(0.0, 0.0);
        if (r < rows) v = __ldg(reinterpret_cast<const double2*>(a.X + (row0 + r) * (unsigned long long)d) + c2);
        *reinterpret_cast<double2*>(xt + (size_t)r * xs + 2 * c2) = v;
      }
    }
    __syncthreads();
    if (t < kDmRows) {
      double s = 0.0;
      for (int j = 0; j < d; ++j) s = fma(xt[(size_t)t * xs + j], xt[(size_t)t * xs + j], s);
      xn2[t] = s;
    }
    double best[2] = {INFINITY, INFINITY}, second[2] = {INFINITY, INFINITY}, cm = 0.0;
    int bidx[2] = {0, 0};
    const double* xa0 = xt + (size_t)(warp * 16 + g) * xs + tq;
    const double* xa1 = xa0 + (size_t)8 * xs;
    for (int ch = 0; ch < n_chunks; ++ch) {
      const int cb = ch * kDmKc, rc = min(kDmKc, a.k - cb);
      __syncthreads();
      {
        const int vpr = d / 2;
        for (int e = t; e < kDmKc * vpr; e += kDmThreads) {
          const int r = e / vpr, c2 = e - r * vpr;
          double2 v = make_double2(0.0, 0.0);
          if (r < rc) v = __ldg(reinterpret_cast<const double2*>(a.C + (size_t)(cb + r) * d) + c2);
          *reinterpret_cast<double2*>(ct + (size_t)r * xs + 2 * c2) = v;
        }
      }
      __syncthreads();
      if (t < kDmKc) {
        double s = 0.0;
        for (int j = 0; j < d; ++j) s = fma(ct[(size_t)t * xs + j], ct[(size_t)t * xs + j], s);
        cn[t] = s;
      }
      __syncthreads();
      double acc[2][8][2];
#pragma unroll
      for (int j = 0; j < 8; ++j) { acc[0][j][0] = acc[0][j][1] = acc[1][j][0] = acc[1][j][1] = 0.0; }
      const double* cbp = ct + (size_t)g * xs + tq;
#pragma unroll 2
      for (int k0 = 0; k0 < d; k0 += 4) {
        const double a0 = xa0[k0], a1 = xa1[k0];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const double b = cbp[(size_t)(8 * j) * xs + k0];
          dmma_884(acc[0][j][0], acc[0][j][1], a0, b);
          dmma_884(acc[1][j][0], acc[1][j][1], a1, b);
        }
      }
      // accumulator fragment: row g (block 0) / g + 8 (block 1), centroids 8 j + 2 tq + {0, 1}
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int cl = 8 * j + 2 * tq + e;
          if (cl < rc) {
            const double nn = cn[cl];
            cm = fmax(cm, nn);
            top2_update<double>(fma(-2.0, acc[0][j][e], nn), cb + cl, best[0], second[0], bidx[0]);
            top2_update<double>(fma(-2.0, acc[1][j][e], nn), cb + cl, best[1], second[1], bidx[1]);
          }
        }
      }
    }
    // the four lanes of a row (tq = 0..3) combine: best by (score, lowest index), runner-up = min of the rest
#pragma unroll
    for (int o = 1; o <= 2; o <<= 1) {
      cm = fmax(cm, __shfl_xor_sync(0xffffffffu, cm, o));
#pragma unroll
      for (int rb = 0; rb < 2; ++rb) {
        const double ob = __shfl_xor_sync(0xffffffffu, best[rb], o), os = __shfl_xor_sync(0xffffffffu, second[rb], o);
        const int oi = __shfl_xor_sync(0xffffffffu, bidx[rb], o);
        const bool other = ob < best[rb] || (ob == best[rb] && oi < bidx[rb]);
        const double lose = other ? best[rb] : ob;
        second[rb] = fmin(fmin(second[rb], os), lose);
        if (other) { best[rb] = ob; bidx[rb] = oi; }
      }
    }
    // lanes tq = 0 / 1 finish row block 0 / 1
    const int rb = tq & 1;
    const int r_local = warp * 16 + rb * 8 + g;
    const bool mine = tq < 2 && r_local < rows;
    const double bb = rb ? best[1] : best[0], ss = rb ? second[1] : second[0];
    int bi = rb ? bidx[1] : bidx[0];
    bool cert = false;
    if (mine) {
      const double xn = sqrt(xn2[r_local]) * 1.000001, cmax = sqrt(cm) * 1.000001, dd = (double)d;
      const double margin = fma(1.1 * eps, (4.0 * dd + 4.0) * xn * cmax + 4.0 * cmax * cmax + (2.0 * (dd + 2.0)) * (xn + cmax) * (xn + cmax),
                                (8.0 * dd + 16.0) * dmin * (1.0 + xn + cmax));
      cert = (ss - bb) > margin;   // NaN anywhere -> false
      const unsigned long long row = row0 + r_local;
      a.labels[row] = (uint32_t)bi;   // provisional when not certified
      if (cert && a.dists != nullptr) a.dists[row] = sqdist_mem<double>(a.C + (size_t)bi * d, xt + (size_t)r_local * xs, d);
    }
    const uint32_t fm = __ballot_sync(0xffffffffu, mine && !cert);
    if (fm) {
      unsigned int base = 0;
      if (lane == 0) base = atomicAdd(a.rescan_count, (unsigned int)__popc(fm));
      base = __shfl_sync(0xffffffffu, base, 0);
      if (mine && !cert) a.rescan_list[base + __popc(fm & ((1u << lane) - 1u))] = (uint32_t)(row0 + r_local);
      if (lane == 0) my_fallbacks += (unsigned int)__popc(fm);
    }
  }
  if (a.state_rw != nullptr && my_fallbacks) atomicAdd(&a.state_rw->fallback_rows, (unsigned long long)my_fallbacks);
}

}  // namespace lkm

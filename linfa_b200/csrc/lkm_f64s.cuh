// Fused Lloyd iteration for f64 rows of 2 / 4 / 8 / 16 features and few centroids (BASELINE cfg5: 125M x 8 per GPU, k = 16).
//
// Replaces, for that shape, lloyd_kernel<double, D, MODE_FUSED> (lkm_kernels.cuh), whose 128 DFMA + 32 DSETP per row keep
// the FP64 pipe (64 lanes / clk / SM) 4x over the HBM budget of 2.8 cycles per row and SM.  Reference semantics:
// KM/algorithm.rs:866-881, 923-946 (assignment, strict '<'), :785-810 (update), :329 (inertia).
//
//   * one warp = one 64-row tile (two rows per lane), no CTA-wide barrier in the loop; tiles are dealt round-robin to
//     the warps of the grid (static schedule => bit-reproducible sums)
//   * rows arrive by cp.async (16-byte chunks, fully coalesced) into a per-warp double-buffered stage whose layout has
//     one 16-byte pad per 128 bytes: the row-per-lane reads (LDS.128) and the feature-per-lane reads of the update walk
//     (LDS.64) are both bank-conflict-free
//   * SCREENING IN F32: w_c = |c|^2 - 2 x.c with x, -2c, |c|^2 rounded to f32 (FFMA2: two centroids per instruction),
//     running best / runner-up by FMNMX on scores that carry the centroid index in their low mantissa bits.  The row is
//     certified when runner-up - best exceeds a rigorous bound on everything that separates this arithmetic from the
//     reference's f64 direct form (derivation at f6_margin); then the reference's strict-'<' scan picks the same centroid.
//     Uncertified rows (near ties, ties, NaN / Inf, f32 overflow / underflow) go through scan_centroids_certified<double>
//     (f64 FMA scores, own certificate, exact re-scan in the reference's arithmetic) exactly like the kernel this one replaces.
//   * winner's distance in f64 (direct form), inertia per lane in f64
//   * update: lane = (feature f, row group h) walks the tile's rows h, h + H, ... straight from the stage, keeps the
//     running sum of the current cluster in a register and touches the warp's accumulator copy in shared memory only when
//     the label changes (blob-ordered data: once per blob); concurrent flushes of one cluster by different row groups are
//     detected (match.any) and serialised.  Per-CTA partials in warp order -> finalize_kernel as for every other path.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "lkm_kernels.cuh"

namespace lkm {
namespace f6 {

constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
// RPT rows per lane: a warp's tile is 32 RPT rows.  RPT = 2 halves the centroid operand loads per row (2 CTAs / SM);
// RPT = 1 needs fewer registers and a smaller stage (3 CTAs / SM).
constexpr int kStages = 2;

struct Args {
  const double* X;
  unsigned long long n;
  int k;
  const double* C;
  double* part_sums;        // [grid][k * D]
  uint32_t* part_counts;    // [grid][k]
  double* part_inertia;     // [grid]
  const LloydState* state;
  LloydState* state_rw;     // fallback_rows += rows that needed the exact re-scan
  uint32_t* labels_out;     // optional (lkm_set_keep_labels): label of every row, for the parity tests
};

template <int D, int RPT> __host__ __device__ constexpr int stage_bytes() { return 32 * RPT * D * 8 + (32 * RPT * D * 8 / 128) * 16; }
// byte offset of a row inside a stage: 16 bytes of padding after every 128 bytes of rows
template <int D> __device__ __forceinline__ uint32_t row_off(int row) {
  constexpr int RP = 16 / D;   // rows per 128 bytes
  return (uint32_t)(row * (D * 8) + (row / RP) * 16);
}

__host__ __device__ inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }
// dynamic shared memory: exact centroids | norms + max | f32 pair operands | f32 pair norms | per warp {stages, sums, counts}
template <int D, int RPT> __host__ __device__ inline size_t smem_bytes(int k) {
  const int kp = (k + 1) & ~1;
  return al16((size_t)k * D * 8) + al16((size_t)(k + 1) * 8) + al16((size_t)kp * D * 4) + al16((size_t)kp * 4) + 16 +
         (size_t)kWarps * (kStages * stage_bytes<D, RPT>() + al16((size_t)(32 / D) * k * D * 8) + al16((size_t)(32 / D) * k * 4));
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src, int src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
  unsigned long long d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ void unpack2(unsigned long long v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// score with the centroid index in the low mantissa bits; then running best / runner-up by min / max
__device__ __forceinline__ void top2_packed(float w, uint32_t idx, uint32_t keep, float& best, float& second) {
  const float q = __uint_as_float((__float_as_uint(w) & keep) | idx);
  second = fminf(second, fmaxf(best, q));
  best = fminf(best, q);
}

// Certificate of the f32 screening.  u = 2^-24.  Inputs x^ = fl32(x), c^ = fl32(-2c), n^ = fl32(|c|^2): relative error u
// each.  The FMA chain s_j = fl(s_{j-1} + x^_j c^_j) rounds d partial sums of magnitude <= M = C^2 + 2XC (X >= |x|,
// C >= max|c|), so |w - W| <= u [ d M + C^2 + 4XC ] (1 + O(u)) against W = |c|^2 - 2x.c.  Packing the index changes a
// score by < (idx_mask + 1) 2^-23 M.  The reference's f64 direct form is within (d+2) 2^-53 (X+C)^2 of the true distance.
// Denormal inputs / results add at most (4d+8) 2^-149 (1 + X + C) in absolute terms.  Both candidates:
//   runner-up - best > 1.1 * 2 * { u [ (d+1) C^2 + (2d+4) XC ] + (idx_mask+1) 2^-23 (C^2 + 2XC) + (d+2) 2^-53 (X+C)^2 }
//                      + (8d+16) 2^-149 (1 + X + C)
// proves that the reference picks `best`.  NaN / Inf anywhere (x, c, overflow) makes the comparison false.
// The bound is a quadratic in X; its coefficients are computed once per thread (rounded up by 1.001):
//   margin(xx) = m0 + m1 sqrt(xx) + m2 xx     with X = 1.0001 sqrt(xx) covering the f32 rounding of xx and the approximate root
struct Margin { float m0, m1, m2; };
template <int D>
__device__ __forceinline__ Margin f6_margin_coeffs(float cmax, float idx_span) {
  const float u = 5.9604645e-8f, pk = idx_span * 1.1920929e-7f, rf = (float)(D + 2) * 1.1102230e-16f;
  const float ab = (float)(8 * D + 16) * 1.4012985e-45f;
  const float c2 = cmax * cmax;
  Margin m;
  m.m0 = 1.001f * (2.2f * (u * (float)(D + 1) + pk + rf) * c2 + ab * (1.f + cmax));
  m.m1 = 1.001f * 1.0001f * (2.2f * (u * (float)(2 * D + 4) + 2.f * pk + 2.f * rf) * cmax + ab);
  m.m2 = 1.001f * 1.0002001f * 2.2f * rf;
  return m;
}
__device__ __forceinline__ float f6_margin(const Margin& m, float xx) {
  float r;
  asm("sqrt.approx.f32 %0, %1;" : "=f"(r) : "f"(xx));
  return fmaf(m.m2, xx, fmaf(m.m1, r, m.m0));
}

template <int D, bool PACKED, int RPT>
__global__ void __launch_bounds__(kThreads, RPT == 1 ? 3 : 2) lloyd_f64s_kernel(const Args a) {
  static_assert(RPT == 1 || RPT == 2, "rows per lane");
  constexpr int kTileRows = 32 * RPT;
  static_assert(D == 2 || D == 4 || D == 8 || D == 16, "row width");
  if (a.state != nullptr && a.state->done) return;
  constexpr int CPR = D / 2;            // 16-byte chunks per row
  constexpr int H = 32 / D;             // rows per step of the update walk
  constexpr int STEPS = kTileRows / H;
  constexpr int SB = stage_bytes<D, RPT>();
  const unsigned full = 0xffffffffu;
  extern __shared__ __align__(16) unsigned char smem[];
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  const int k = a.k, kp = (k + 1) & ~1;
  size_t off = 0;
  double* cs = reinterpret_cast<double*>(smem + off); off += al16((size_t)k * D * 8);
  double* cn = reinterpret_cast<double*>(smem + off); off += al16((size_t)(k + 1) * 8);
  float* cfp = reinterpret_cast<float*>(smem + off); off += al16((size_t)kp * D * 4);   // [pair][feature][2] = -2c
  float* cnp = reinterpret_cast<float*>(smem + off); off += al16((size_t)kp * 4);       // [pair][2] = |c|^2
  float* cmaxf_s = reinterpret_cast<float*>(smem + off); off += 16;
  const size_t acc_bytes = al16((size_t)H * k * D * 8);   // one accumulator copy per row group of the warp
  const size_t per_warp = (size_t)kStages * SB + acc_bytes + al16((size_t)H * k * 4);
  unsigned char* wbase = smem + off + (size_t)warp * per_warp;
  unsigned char* stage0 = wbase;
  double* acc = reinterpret_cast<double*>(wbase + (size_t)kStages * SB);
  uint32_t* cnt = reinterpret_cast<uint32_t*>(wbase + (size_t)kStages * SB + acc_bytes);

  // ---- tile schedule and the first load (before the prologue, so that it overlaps it) ----
  const unsigned long long n_tiles = (a.n + kTileRows - 1) / kTileRows;
  const unsigned long long G = (unsigned long long)gridDim.x * kWarps;
  unsigned long long tile = (unsigned long long)blockIdx.x * kWarps + warp;
  // one cp.async instruction of the warp covers 32 chunks = 512 contiguous bytes of X = 576 bytes of the padded stage
  const uint32_t lane_soff = row_off<D>(lane / CPR) + (uint32_t)(lane % CPR) * 16u;
  const uint32_t stage_s = smem_u32(stage0);
  const unsigned char* xbytes = reinterpret_cast<const unsigned char*>(a.X);
  auto issue = [&](unsigned long long tl, int st) {
    const unsigned long long row0 = tl * kTileRows;
    const uint32_t dst = stage_s + (uint32_t)st * (uint32_t)SB + lane_soff;
    const unsigned char* src = xbytes + row0 * (unsigned long long)(D * 8) + (unsigned)lane * 16u;
    if (row0 + kTileRows <= a.n) {
#pragma unroll
      for (int q = 0; q < kTileRows * CPR / 32; ++q) cp_async16(dst + q * 576, src + q * 512, 16);
    } else {   // ragged last tile: rows beyond n are zero-filled (and ignored through their label -1)
#pragma unroll
      for (int q = 0; q < kTileRows * CPR / 32; ++q) {
        const bool ok = row0 + (unsigned)(q * (32 / CPR) + lane / CPR) < a.n;
        cp_async16(dst + q * 576, ok ? src + q * 512 : xbytes, ok ? 16 : 0);
      }
    }
  };
  if (tile < n_tiles) issue(tile, 0);
  cp_async_commit();

  // ---- prologue: centroids in f64 (exact paths) and as f32 pair operands (screening) ----
  for (int e = t; e < k * D; e += kThreads) cs[e] = a.C[e];
  for (int e = lane; e < H * k * D; e += 32) acc[e] = 0.0;
  for (int e = lane; e < H * k; e += 32) cnt[e] = 0u;
  __syncthreads();
  for (int c = t; c < kp; c += kThreads) {
    double nn = 0.0;
    if (c < k)
      for (int j = 0; j < D; ++j) nn += cs[c * D + j] * cs[c * D + j];
    if (c < k) cn[c] = nn;
    cnp[c] = c < k ? (float)nn : 3.0e38f;
    for (int j = 0; j < D; ++j) cfp[((c >> 1) * D + j) * 2 + (c & 1)] = c < k ? (float)(-2.0 * cs[c * D + j]) : 0.f;
  }
  __syncthreads();
  if (t == 0) {
    double m = 0.0;
    bool bad = false;
    for (int c = 0; c < k; ++c) {
      m = (cn[c] > m) ? cn[c] : m;   // NaN norms are skipped here (the f64 certified scan sends such rows to the exact path)
      bad |= !(cn[c] < 1.0e300);     // NaN / Inf / huge: no f32 certificate at all
    }
    const double cm = sqrt(m) * 1.000001;
    cn[k] = cm;
    cmaxf_s[0] = bad ? __int_as_float(0x7fc00000) : __double2float_ru(cm);
  }
  __syncthreads();
  const double cmax64 = cn[k];
  const float cmaxf = cmaxf_s[0];
  int idx_bits = 1;
  while ((1 << idx_bits) < kp) ++idx_bits;
  const uint32_t idx_mask = (1u << idx_bits) - 1u, keep = ~idx_mask;
  const Margin mg = f6_margin_coeffs<D>(cmaxf, (float)(idx_mask + 1u));

  const int f = lane % D, h = lane / D;
  double* acc_h = acc + (size_t)h * k * D;
  uint32_t* cnt_h = cnt + (size_t)h * k;
  int cur_lab = -1;
  double cur_acc = 0.0;
  uint32_t cur_cnt = 0u;
  double inert = 0.0;
  unsigned int my_fallbacks = 0;

  for (int it = 0; tile < n_tiles; ++it, tile += G) {
    const int st = it & 1;
    if (tile + G < n_tiles) issue(tile + G, st ^ 1);
    cp_async_commit();
    cp_async_wait<1>();
    __syncwarp();
    const unsigned char* sp = stage0 + (size_t)st * SB;
    const unsigned long long row0 = tile * kTileRows;

    // ---- own rows: lane (+ 32) ----
    double x[RPT][D];
    float xf[RPT][D], xx[RPT], bq[RPT], sq[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
#pragma unroll
      for (int c = 0; c < CPR; ++c) {
        const double2 v = *reinterpret_cast<const double2*>(sp + row_off<D>(lane + 32 * r) + c * 16);
        x[r][2 * c] = v.x; x[r][2 * c + 1] = v.y;
      }
      xx[r] = 0.f;
#pragma unroll
      for (int j = 0; j < D; ++j) {
        xf[r][j] = (float)x[r][j];
        xx[r] = fmaf(xf[r][j], xf[r][j], xx[r]);
      }
      bq[r] = 3.0e38f; sq[r] = 3.0e38f;
    }
    if constexpr (PACKED) {
      unsigned long long xd[RPT][D];
#pragma unroll
      for (int r = 0; r < RPT; ++r)
#pragma unroll
        for (int j = 0; j < D; ++j) xd[r][j] = pack2(xf[r][j], xf[r][j]);
#pragma unroll 2
      for (int p = 0; p < kp / 2; ++p) {
        unsigned long long w[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) w[r] = *reinterpret_cast<const unsigned long long*>(cnp + 2 * p);
#pragma unroll
        for (int j = 0; j < D; j += 2) {
          const ulonglong2 c2 = *reinterpret_cast<const ulonglong2*>(cfp + (p * D + j) * 2);
#pragma unroll
          for (int r = 0; r < RPT; ++r) w[r] = f2_fma(xd[r][j], c2.x, w[r]);
#pragma unroll
          for (int r = 0; r < RPT; ++r) w[r] = f2_fma(xd[r][j + 1], c2.y, w[r]);
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          float wa, wb;
          unpack2(w[r], wa, wb);
          top2_packed(wa, 2u * p, keep, bq[r], sq[r]);
          top2_packed(wb, 2u * p + 1u, keep, bq[r], sq[r]);
        }
      }
    } else {
#pragma unroll 2
      for (int p = 0; p < kp / 2; ++p) {
        const float2 nn = *reinterpret_cast<const float2*>(cnp + 2 * p);
        float wa[RPT], wb[RPT];
#pragma unroll
        for (int r = 0; r < RPT; ++r) { wa[r] = nn.x; wb[r] = nn.y; }
#pragma unroll
        for (int j = 0; j < D; j += 2) {
          const float4 c4 = *reinterpret_cast<const float4*>(cfp + (p * D + j) * 2);
#pragma unroll
          for (int r = 0; r < RPT; ++r) {
            wa[r] = fmaf(xf[r][j], c4.x, wa[r]); wb[r] = fmaf(xf[r][j], c4.y, wb[r]);
            wa[r] = fmaf(xf[r][j + 1], c4.z, wa[r]); wb[r] = fmaf(xf[r][j + 1], c4.w, wb[r]);
          }
        }
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
          top2_packed(wa[r], 2u * p, keep, bq[r], sq[r]);
          top2_packed(wb[r], 2u * p + 1u, keep, bq[r], sq[r]);
        }
      }
    }
    // ---- certificate, winner's distance ----
    int labs[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) {
      const bool ok = row0 + (unsigned)(r * 32 + lane) < a.n;
      int bi = (int)(__float_as_uint(bq[r]) & idx_mask);
      const bool cert = ((sq[r] - bq[r]) > f6_margin(mg, xx[r])) && bi < k;
      double dist = 0.0;
      if (ok) {
        if (cert) {
#pragma unroll
          for (int j = 0; j < D; ++j) {
            const double df = cs[bi * D + j] - x[r][j];
            dist = fma(df, df, dist);
          }
        } else {
          my_fallbacks += scan_centroids_certified<double, D>(cs, cn, k, cmax64, x[r], dist, bi) ? 0u : 1u;
        }
        inert += dist;
      }
      labs[r] = ok ? bi : -1;
      if (a.labels_out != nullptr && ok) a.labels_out[row0 + (unsigned)(r * 32 + lane)] = (uint32_t)bi;
    }
    // ---- update walk: lane = (feature f, row group h), rows h, h + H, ... of the tile ----
    const int lab_t = __shfl_sync(full, labs[0], 0);
    bool same = labs[0] == lab_t;
#pragma unroll
    for (int r = 1; r < RPT; ++r) same = same && labs[r] == lab_t;
    if (__all_sync(full, same) && lab_t >= 0) {
      // one cluster for the whole tile (the rule for blob-ordered rows): sums stay in registers; cur_lab is warp-uniform
      if (lab_t != cur_lab) {
        if (cur_lab >= 0) {
          acc_h[cur_lab * D + f] += cur_acc;
          if (f == 0) cnt_h[cur_lab] += cur_cnt;
        }
        cur_lab = lab_t; cur_acc = 0.0; cur_cnt = 0u;
      }
      double e0 = 0.0, e1 = 0.0;
#pragma unroll
      for (int m = 0; m < STEPS; m += 2) {
        e0 += *reinterpret_cast<const double*>(sp + row_off<D>(m * H + h) + 8 * f);
        e1 += *reinterpret_cast<const double*>(sp + row_off<D>((m + 1) * H + h) + 8 * f);
      }
      cur_acc += e0 + e1;
      cur_cnt += (uint32_t)STEPS;
    } else {
      // mixed tile: every row goes straight to this row group's private accumulator copy (no conflicts, no votes)
      if (cur_lab >= 0) {
        acc_h[cur_lab * D + f] += cur_acc;
        if (f == 0) cnt_h[cur_lab] += cur_cnt;
        cur_lab = -1; cur_acc = 0.0; cur_cnt = 0u;
      }
#pragma unroll
      for (int m = 0; m < STEPS; ++m) {
        const int row = m * H + h;
        const int lab = __shfl_sync(full, labs[(m * H) >> 5], row & 31);
        const double xv = *reinterpret_cast<const double*>(sp + row_off<D>(row) + 8 * f);
        if (lab >= 0) {
          acc_h[lab * D + f] += xv;
          if (f == 0) cnt_h[lab] += 1u;
        }
      }
    }
    __syncwarp();   // every lane is done with the stage before the next iteration's cp.async overwrites it
  }
  cp_async_wait<0>();
  if (cur_lab >= 0) {
    acc_h[cur_lab * D + f] += cur_acc;
    if (f == 0) cnt_h[cur_lab] += cur_cnt;
  }
  if (a.state_rw != nullptr && my_fallbacks) atomicAdd(&a.state_rw->fallback_rows, (unsigned long long)my_fallbacks);

  // ---- per-CTA partials: the warps' copies in warp order ----
  __syncthreads();
  const unsigned char* w0base = smem + off;
  for (int e = t; e < k * D; e += kThreads) {
    double s = 0.0;
    for (int w = 0; w < kWarps; ++w) {
      const double* aw = reinterpret_cast<const double*>(w0base + (size_t)w * per_warp + (size_t)kStages * SB);
      for (int g = 0; g < H; ++g) s += aw[(size_t)g * k * D + e];
    }
    a.part_sums[(size_t)blockIdx.x * k * D + e] = s;
  }
  for (int e = t; e < k; e += kThreads) {
    uint32_t s = 0u;
    for (int w = 0; w < kWarps; ++w) {
      const uint32_t* cw = reinterpret_cast<const uint32_t*>(w0base + (size_t)w * per_warp + (size_t)kStages * SB + acc_bytes);
      for (int g = 0; g < H; ++g) s += cw[g * k + e];
    }
    a.part_counts[(size_t)blockIdx.x * k + e] = s;
  }
  for (int o = 16; o > 0; o >>= 1) inert += __shfl_xor_sync(full, inert, o);
  __shared__ double red[kWarps];
  if (lane == 0) red[warp] = inert;
  __syncthreads();
  if (t == 0) {
    double s = 0.0;
    for (int w = 0; w < kWarps; ++w) s += red[w];
    a.part_inertia[blockIdx.x] = s;
  }
}

}  // namespace f6
}  // namespace lkm

// Fused Lloyd iteration for f32, d = 32, k <= 64 — third-generation tcgen05 kernel (cfg2 shape).
//
// One persistent CTA per SM, three independent 128-thread worker groups.  Group w owns the tiles
// j = w (mod 3) of the CTA and walks each one through
//   S  build the TF32 low part A_lo = x - trunc(x) of the tile (the raw tile, loaded by TMA with
//      SWIZZLE_128B, IS the high part: kind::tf32 ignores the 13 low mantissa bits — probed, see
//      profiles/umma_layout_probes.md)
//   M  13 tcgen05 MMAs into the group's 64 TMEM columns:  x_lo.c_hi + x_hi.c_lo + x_hi.c_hi  and one
//      last K-step  ones . (K0 - |c|^2 / 2)  that adds the per-centroid bias, so that the accumulator
//      holds v_c = x.c - |c|^2/2 + K0  (argmax v = argmin distance)
//   E  thread = row: K0 is chosen so that every v lies in one 2^8-wide binade window; then
//      q_c = (bits(v_c) << 6) | c is monotone in v and carries the index, best = max3-tree(q),
//      runner-up = best - 1 - min3-tree(best - 1 - q): three integer instructions per score.
//      The row is certified when v_best - v_second exceeds the rigorous error bound of DESIGN.md;
//      other rows are re-scanned in the reference's exact arithmetic (one warp per row).
//   U  thread = (row subset, feature): sums and inertia straight from the raw tile, thread-owned
//      shared-memory accumulators.  The U phases of the three groups are serialised in tile order
//      by an mbarrier token, so every accumulator sees its rows in ascending order (deterministic)
//      and one accumulator set serves all groups.
// While a group waits (TMA, MMA, token) the other two issue.  Tile j+2 of a group is requested as
// soon as U(j) has released its buffer, so each group always has 16 KB in flight.
//
// Same outputs as lloyd_tc_kernel: per-CTA partial sums / counts / inertia for finalize_kernel.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "lkm_kernels.cuh"
#include "lkm_tc.cuh"

namespace lkm {
namespace tc {
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// 2-D tiled TMA load global -> shared (coordinates: c0 = innermost), completion on an mbarrier
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* tmap, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(smem_u32(dst)), "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// hint: bring `bytes` (multiple of 16) at `src` into L2; the later TMA load of the same lines then sees L2 latency, so a
// shared-memory ring of a few stages covers it (the ring alone would need HBM latency x bandwidth bytes in flight)
__device__ __forceinline__ void l2_prefetch(const void* src, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void group_sync(int id) { asm volatile("bar.sync %0, 128;" ::"r"(id) : "memory"); }
// K-major SW128 descriptor whose 8-row groups all alias the same 1 KB atom (SBO = 0): a constant tile
__device__ __forceinline__ uint64_t smem_desc_sw128_alias(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// bounded wait; a barrier that never completes is a protocol bug: kill the kernel instead of hanging the GPU.
// The suspend-time hint lets the hardware park the warp until the phase completes instead of re-polling every
// few cycles (polling warps steal issue slots and mbarrier bandwidth from the working roles).
#ifndef LKM_WAIT_HINT_NS
#define LKM_WAIT_HINT_NS 10000
#endif
__device__ __forceinline__ void mbar_wait_or_trap(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1
  for (uint32_t i = 0; i < (LKM_WAIT_HINT_NS ? (1u << 18) : (1u << 28)); ++i) {
    uint32_t ok;
#if LKM_WAIT_HINT_NS
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity), "r"((uint32_t)LKM_WAIT_HINT_NS)
        : "memory");
#else
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
#endif
    if (ok) return;
  }
  __trap();
}
// the same without a suspend-time hint, for single-warp roles on the critical path (producer, relay, MMA issuer):
// they re-test as soon as the try_wait returns instead of sleeping in 10 us slices woken by unrelated barrier events
__device__ __forceinline__ void mbar_wait_spin(uint64_t* bar, uint32_t parity) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1
  for (uint32_t i = 0; i < (1u << 28); ++i) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return;
  }
  __trap();
}
__device__ __forceinline__ float trunc_tf32(float x) { return __uint_as_float(__float_as_uint(x) & 0xFFFFE000u); }
}  // namespace tc

// max_i |x_i|^2 over the resident rows (d = 32): one warp per 4 rows, 8 lanes per row
__global__ void __launch_bounds__(256) row_norm2_max_kernel(const float* __restrict__ X, unsigned long long n,
                                                            unsigned int* __restrict__ out_bits) {
  const unsigned long long n_chunks = n * 8;
  float m = 0.f;
  bool bad = false;
  for (unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; i < ((n_chunks + 31) & ~31ull);
       i += (unsigned long long)gridDim.x * blockDim.x) {
    float s = 0.f;
    if (i < n_chunks) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(X) + i);
      s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
    }
    s += __shfl_xor_sync(0xffffffffu, s, 1);
    s += __shfl_xor_sync(0xffffffffu, s, 2);
    s += __shfl_xor_sync(0xffffffffu, s, 4);
    bad |= !(s <= 3.0e38f);
    m = fmaxf(m, s);
  }
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  bad = __any_sync(0xffffffffu, bad);
  // non-negative floats order like their bit patterns; a non-finite norm is published as NaN bits
  if ((threadIdx.x & 31) == 0) atomicMax(out_bits, bad ? 0x7FC00000u : __float_as_uint(m));
}

struct Tc3Args {
  const float* X;
  unsigned long long n;
  int k;
  const float* C;
  const unsigned int* xmax2_bits;  // row_norm2_max_kernel's result
  float* part_sums;
  uint32_t* part_counts;
  double* part_inertia;
  LloydState* state;
};

constexpr int kT3Groups = 3;
constexpr int kT3Threads = 128 * kT3Groups;
constexpr int kT3TmemCols = 256;

__host__ __device__ inline size_t tc3_smem_bytes(int k, int kp) {
  size_t b = 0;
  b += (size_t)kT3Groups * 2 * 16384;            // raw tiles (TMA destinations, A_hi operands)
  b += (size_t)kT3Groups * 16384;                // A_lo
  b += (size_t)kp * 128 * 3;                     // B_hi, B_lo, B_bias
  b += 1024;                                     // ones atom
  b += (size_t)k * kTcCStride * 4;               // raw centroids
  b += (size_t)kTcSubsets * (k + 1) * 32 * 4;    // sums
  b += (((size_t)k * 4) + 15) & ~(size_t)15;     // counts
  b += (size_t)kp * 8;                           // centroid squared norms (f64)
  b += (size_t)kT3Groups * kTcRows * 4 * 2;      // per group: subset-major labels, flag list
  b += (size_t)kT3Groups * 16;                   // per group: warp-uniform labels
  return b + 1024;
}

template <int KP>
__global__ void __launch_bounds__(kT3Threads, 1) lloyd_tc3_kernel(const Tc3Args a, const __grid_constant__ CUtensorMap tmap) {
  static_assert(KP == 32 || KP == 64, "score columns per group");
  pdl_launch_dependents();
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (tc::smem_u32(smem_unaligned) & 1023u)) & 1023u);
  const int t = threadIdx.x, lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);   // warp-uniform by construction: lets ptxas use uniform registers
  const int w = warp >> 2;       // worker group
  const int wt = t & 127;        // thread within the group = row of the tile in E
  const int wwarp = warp & 3;    // TMEM lane quadrant
  // the one thread of the group that issues TMA / MMA / token arrivals: lane 0 of warp w, so that the
  // three groups' issuers sit on three different schedulers
  const bool issuer = (wwarp == w) && (lane == 0);
  const int k = a.k;
  unsigned char* raw = smem;                                   // [group][2][16 KB]
  unsigned char* alo = raw + (size_t)kT3Groups * 2 * 16384;    // [group][16 KB]
  unsigned char* b_hi = alo + (size_t)kT3Groups * 16384;
  unsigned char* b_lo = b_hi + KP * 128;
  unsigned char* b_bias = b_lo + KP * 128;
  unsigned char* ones = b_bias + KP * 128;
  float* rawc = reinterpret_cast<float*>(ones + 1024);
  float* acc = rawc + (size_t)k * kTcCStride;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(acc + (size_t)kTcSubsets * (k + 1) * 32);
  double* nn_s = reinterpret_cast<double*>(cnt + ((k + 3) & ~3));
  uint32_t* labs2_all = reinterpret_cast<uint32_t*>(nn_s + KP);
  uint32_t* flags_all = labs2_all + kT3Groups * kTcRows;
  uint32_t* wuni_all = flags_all + kT3Groups * kTcRows;
  __shared__ __align__(8) uint64_t bar_full[kT3Groups][2], bar_mma[kT3Groups], bar_tok[kT3Groups];
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t n_flag[kT3Groups];
  __shared__ uint32_t cmax_bits;
  __shared__ uint32_t abort_flag;

  const unsigned long long n_tiles = (a.n + kTcRows - 1) / kTcRows;
  const unsigned long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const unsigned long long my_cnt = my_tiles > (unsigned long long)w ? (my_tiles - w + kT3Groups - 1) / kT3Groups : 0;
  auto tile_of = [&](unsigned long long i) { return blockIdx.x + ((unsigned long long)w + kT3Groups * i) * gridDim.x; };

  if (warp == 0) tc::tmem_alloc(&tmem_slot, kT3TmemCols);
  if (t == 0) {
    for (int g = 0; g < kT3Groups; ++g) {
      tc::mbar_init(&bar_full[g][0], 1);
      tc::mbar_init(&bar_full[g][1], 1);
      tc::mbar_init(&bar_mma[g], 1);
      tc::mbar_init(&bar_tok[g], 1);
      n_flag[g] = 0;
    }
    tc::fence_mbar_init();
    cmax_bits = 0;
    abort_flag = 0;
  }
  __syncthreads();
  // The observations do not depend on the previous kernel: request the first two tiles of every
  // group before waiting for it.
  if (issuer) {
    for (int b = 0; b < 2; ++b)
      if ((unsigned long long)b < my_cnt) {
        tc::mbar_expect_tx(&bar_full[w][b], 16384u);
        tc::tma_load_2d(raw + ((size_t)w * 2 + b) * 16384, &tmap, 0, (int)(tile_of(b) * kTcRows), &bar_full[w][b]);
      }
  }
  if (t == 0) tc::mbar_arrive(&bar_tok[0]);  // group 0 holds the accumulation token first
  pdl_wait();
  const bool done = a.state != nullptr && a.state->done;

  // ---- centroids: TF32 hi/lo operand tiles, padded raw copy, squared norms (8 threads per row) ----
  for (int item = t; item < KP * 8; item += kT3Threads) {
    const int r = item >> 3, c = item & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < k) v = __ldg(reinterpret_cast<const float4*>(a.C + (size_t)r * 32) + c);
    float4 h, l;
    tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
    tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
    const uint32_t o = tc::sw128_offset(r, c);
    *reinterpret_cast<float4*>(b_hi + o) = h;
    *reinterpret_cast<float4*>(b_lo + o) = l;
    *reinterpret_cast<float4*>(b_bias + o) = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < k) *reinterpret_cast<float4*>(rawc + r * kTcCStride + 4 * c) = v;
    double nn = (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
    nn += __shfl_xor_sync(0xffffffffu, nn, 1);
    nn += __shfl_xor_sync(0xffffffffu, nn, 2);
    nn += __shfl_xor_sync(0xffffffffu, nn, 4);
    if (c == 0) {
      nn_s[r] = nn;
      if (r < k) atomicMax(&cmax_bits, __float_as_uint(__double2float_ru(nn)));  // a NaN norm wins and voids every certificate
    }
  }
  for (int e = t; e < 256; e += kT3Threads) reinterpret_cast<float*>(ones)[e] = 0.f;
  for (int e = t; e < kTcSubsets * (k + 1) * 32; e += kT3Threads) acc[e] = 0.f;
  for (int e = t; e < k; e += kT3Threads) cnt[e] = 0u;
  __syncthreads();
  // ---- window, bias, margin (identical in every thread) ----
  const float cmax2 = __uint_as_float(cmax_bits);
  const float cmax = sqrtf(cmax2) * 1.000001f;
  const float xmax2 = __uint_as_float(__ldg(a.xmax2_bits));
  const float xmax = sqrtf(xmax2) * 1.000001f;
  // |x.c - |c|^2/2| <= R for every row and centroid
  const float R = 1.01f * (xmax * cmax + 0.5f * cmax2) + 1e-30f;
  const bool params_ok = R < 1e30f;  // false for NaN / Inf: every row then takes the exact scan
  // binade window [2^lo, 2^(lo+8)), biased exponent of 2^lo a multiple of 8, with 127 * 2^lo > R
  uint32_t e_lo = ((__float_as_uint(R * (1.0f / 127.0f)) >> 23) + 1u + 7u) & ~7u;
  e_lo = min(max(e_lo, 8u), 240u);
  const float two_lo = __uint_as_float(e_lo << 23);
  const float K0 = R + two_lo;                 // v = x.c - |c|^2/2 + K0 in [2^lo, 255 * 2^lo)
  const uint32_t top_bits = (e_lo << 23) & 0xFC000000u;
  const float eta = (3.f * 32.f + 12.f) * 2.3841858e-7f;  // (3d + 12) * 2^-22, 3xTF32 dot error / (|x||c|)
  const float gamma = (32.f + 2.f) * 5.9604645e-8f;       // (d + 2) * 2^-24, reference's own rounding
  const float m_w = 1.1f * (2.3841858e-7f * cmax2 + xmax * (4.f * eta + 4.7683716e-7f) * cmax +
                            2.f * gamma * (xmax + cmax) * (xmax + cmax));
  // v is half a squared distance; the bias K-step rounds at most 3 ulp(v) on each of the two candidates
  const float m_v = 0.5f * m_w + 6.f * 1.1920929e-7f * (K0 + R);
  if (t < KP) {
    // bias row t: K0 - |c|^2/2 as three TF32 pieces (exact to ~2^-33), K-step 0 of the tile
    const double b = (t < k) ? ((double)K0 - 0.5 * nn_s[t]) : (double)two_lo;
    const float h = tc::trunc_tf32((float)b);
    const double r1 = b - (double)h;
    const float l = tc::trunc_tf32((float)r1);
    const float l2 = tc::trunc_tf32((float)(r1 - (double)l));
    float* p = reinterpret_cast<float*>(b_bias + tc::sw128_offset(t, 0));
    p[0] = h; p[1] = l; p[2] = l2;
  } else if (t >= 64 && t < 72) {
    float* p = reinterpret_cast<float*>(ones + tc::sw128_offset(t - 64, 0));
    p[0] = 1.f; p[1] = 1.f; p[2] = 1.f;
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;

  unsigned char* my_raw = raw + (size_t)w * 2 * 16384;
  unsigned char* my_alo = alo + (size_t)w * 16384;
  uint32_t* labs2 = labs2_all + w * kTcRows;   // labs2[(row & 3) * 32 + (row >> 2)]
  uint32_t* flags = flags_all + w * kTcRows;
  uint32_t* wuni = wuni_all + w * 4;
  const int bar_id = 1 + w;
  // U role: feature `feat` of rows sub, sub + 4, ...
  const int sub = wwarp, feat = lane;
  const uint32_t acc_off_e = (uint32_t)sub * 128u + ((((uint32_t)feat >> 2) ^ (uint32_t)sub) << 4) + ((uint32_t)feat & 3u) * 4u;
  const uint32_t acc_off_o = (uint32_t)(sub + 4) * 128u + ((((uint32_t)feat >> 2) ^ (uint32_t)(sub + 4)) << 4) + ((uint32_t)feat & 3u) * 4u;
  float* my_acc = acc + (size_t)sub * (k + 1) * 32 + feat;
  double inert = 0.0;
  float inert_f = 0.f;   // per-tile partials, flushed into the f64 sum every 8 tiles
  unsigned int my_fallbacks = 0;
  // operand descriptors of this group (start address in the low 14 bits, 16-byte units)
  constexpr uint32_t idesc = tc::idesc_tf32(128, (uint32_t)KP);
  const uint64_t desc_raw = tc::smem_desc_sw128(tc::smem_u32(my_raw));
  const uint64_t desc_alo = tc::smem_desc_sw128(tc::smem_u32(my_alo));
  const uint64_t desc_bhi = tc::smem_desc_sw128(tc::smem_u32(b_hi));
  const uint64_t desc_blo = tc::smem_desc_sw128(tc::smem_u32(b_lo));
  const uint64_t desc_bias = tc::smem_desc_sw128(tc::smem_u32(b_bias));
  const uint64_t desc_ones = tc::smem_desc_sw128_alias(tc::smem_u32(ones));
  const uint32_t dcol = tmem + (uint32_t)w * 64u;

  if (done) {
    // nothing to do, but the requested tiles must land before the CTA may exit
    if (issuer)
      for (int b = 0; b < 2; ++b)
        if ((unsigned long long)b < my_cnt) tc::mbar_wait_or_trap(&bar_full[w][b], 0);
  } else {
    for (unsigned long long i = 0; i < my_cnt; ++i) {
      const int b = (int)(i & 1);
      const uint32_t ph = (uint32_t)(i >> 1) & 1u;
      const unsigned long long row0 = tile_of(i) * kTcRows;
      const int rows = (int)min((unsigned long long)kTcRows, a.n - row0);
      unsigned char* rw = my_raw + (size_t)b * 16384;
      tc::mbar_wait_or_trap(&bar_full[w][b], ph);
      // ---- S: A_lo = x - trunc_tf32(x), chunk by chunk at the same (swizzled) position ----
      {
        float4 v[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) v[q] = *reinterpret_cast<const float4*>(rw + (uint32_t)wt * 16u + 2048u * q);
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float4 l;
          l.x = v[q].x - tc::trunc_tf32(v[q].x); l.y = v[q].y - tc::trunc_tf32(v[q].y);
          l.z = v[q].z - tc::trunc_tf32(v[q].z); l.w = v[q].w - tc::trunc_tf32(v[q].w);
          *reinterpret_cast<float4*>(my_alo + (uint32_t)wt * 16u + 2048u * q) = l;
        }
      }
      tc::fence_async_smem();
      tc::tc_fence_before();
      tc::group_sync(bar_id);
      // ---- M ----
      if (issuer) {
        tc::tc_fence_after();
        const uint64_t da_hi = desc_raw + (uint64_t)(b ? 1024u : 0u);   // 16 KB further, in 16-byte units
        // x_lo.c_hi, x_hi.c_lo, x_hi.c_hi; one K-step of 8 TF32 = 32 bytes = 2 descriptor units
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) tc::mma_tf32(dcol, desc_alo + 2u * ks, desc_bhi + 2u * ks, idesc, ks ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) tc::mma_tf32(dcol, da_hi + 2u * ks, desc_blo + 2u * ks, idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) tc::mma_tf32(dcol, da_hi + 2u * ks, desc_bhi + 2u * ks, idesc, 1u);
        tc::mma_tf32(dcol, desc_ones, desc_bias, idesc, 1u);
        tc::mma_commit(&bar_mma[w]);
      }
      tc::mbar_wait_or_trap(&bar_mma[w], (uint32_t)i & 1u);
      tc::tc_fence_after();
      // ---- E: packed scores, best and runner-up by 3-input integer max / min ----
      uint32_t q[KP];
      {
        const uint32_t taddr = tmem + ((uint32_t)(wwarp * 32) << 16) + (uint32_t)w * 64u;
        uint32_t r0[32];
        tc::tmem_ld32(taddr, r0);
        if constexpr (KP == 64) {
          uint32_t r1[32];
          tc::tmem_ld32(taddr + 32u, r1);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) q[32 + c] = (r1[c] << 6) | (uint32_t)(32 + c);
        } else {
          tc::tmem_ld_wait();
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) q[c] = (r0[c] << 6) | (uint32_t)c;
      }
      tc::tc_fence_before();
      uint32_t m0 = q[0], m1 = q[1], m2 = q[2], m3 = q[3];
#pragma unroll
      for (int c = 4; c + 7 < KP; c += 8) {
        m0 = __vimax3_u32(m0, q[c], q[c + 1]);
        m1 = __vimax3_u32(m1, q[c + 2], q[c + 3]);
        m2 = __vimax3_u32(m2, q[c + 4], q[c + 5]);
        m3 = __vimax3_u32(m3, q[c + 6], q[c + 7]);
      }
      // KP - 4 is a multiple of 4 but not of 8: four scores left
      m0 = __vimax3_u32(m0, q[KP - 4], q[KP - 3]);
      m1 = __vimax3_u32(m1, q[KP - 2], q[KP - 1]);
      const uint32_t bestq = max(__vimax3_u32(m0, m1, m2), m3);
      const uint32_t bm1 = bestq - 1u;
#pragma unroll
      for (int c = 0; c < KP; ++c) q[c] = bm1 - q[c];   // the best wraps to 0xFFFFFFFF, the others are gap - 1
      m0 = q[0]; m1 = q[1]; m2 = q[2]; m3 = q[3];
#pragma unroll
      for (int c = 4; c + 7 < KP; c += 8) {
        m0 = __vimin3_u32(m0, q[c], q[c + 1]);
        m1 = __vimin3_u32(m1, q[c + 2], q[c + 3]);
        m2 = __vimin3_u32(m2, q[c + 4], q[c + 5]);
        m3 = __vimin3_u32(m3, q[c + 6], q[c + 7]);
      }
      m0 = __vimin3_u32(m0, q[KP - 4], q[KP - 3]);
      m1 = __vimin3_u32(m1, q[KP - 2], q[KP - 1]);
      const uint32_t tmin = min(__vimin3_u32(m0, m1, m2), m3);
      const uint32_t secondq = bm1 - tmin;
      const float v_best = __uint_as_float((bestq >> 6) | top_bits);
      const float v_second = __uint_as_float((secondq >> 6) | top_bits);
      const uint32_t bidx = bestq & 63u;
      const bool row_ok = wt < rows;
      const bool certified = params_ok && ((v_best - v_second) > m_v) && bidx < (uint32_t)k;
      uint32_t lab = certified ? bidx : 0xFFFFFFFFu;
      if (!row_ok) lab = 0xFFFFFFFEu;
      {
        const uint32_t lab0 = __shfl_sync(0xffffffffu, lab, 0);
        const bool uni = __all_sync(0xffffffffu, lab == lab0) && lab0 < 0xFFFFFFFEu;
        if (uni) {
          if (lane == 0) atomicAdd(&cnt[lab0], 32u);
        } else if (lab < 0xFFFFFFFEu) {
          atomicAdd(&cnt[lab], 1u);
        }
        if (lane == 0) wuni[wwarp] = uni ? lab0 : 0xFFFFFFFFu;
      }
      labs2[(wt & 3) * 32 + (wt >> 2)] = lab;
      if (lab == 0xFFFFFFFFu) {
        const uint32_t p = atomicAdd(&n_flag[w], 1u);
        flags[p] = (uint32_t)wt;
      }
      tc::group_sync(bar_id);
      // ---- uncertified rows: exact scan, one warp per row, lanes split the centroids ----
      const uint32_t nf = n_flag[w];
      if (nf) {
        for (uint32_t f = wwarp; f < nf; f += 4) {
          const uint32_t frow = flags[f];
          float x2[32];
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const float4 v = *reinterpret_cast<const float4*>(rw + tc::sw128_offset(frow, (uint32_t)c));
            x2[4 * c] = v.x; x2[4 * c + 1] = v.y; x2[4 * c + 2] = v.z; x2[4 * c + 3] = v.w;
          }
          float fb = INFINITY;
          int fi = -1;
          int c = lane;
          if (lane == 0) { fb = sqdist_reg<float, 32>(rawc, x2); fi = 0; c = 32; }
          for (; c < k; c += 32) {
            const float v = sqdist_reg<float, 32>(rawc + c * kTcCStride, x2);
            if (v < fb) { fb = v; fi = c; }
          }
          combine_candidates<float>(32, fb, fi);
          if (lane == 0) {
            labs2[(frow & 3) * 32 + (frow >> 2)] = (uint32_t)fi;
            atomicAdd(&cnt[fi], 1u);
          }
        }
        my_fallbacks += (wt == 0) ? nf : 0u;
        tc::group_sync(bar_id);
        if (wt == 0) n_flag[w] = 0;
      }
      // ---- U: wait for the token, then sums and inertia from the raw tile ----
      tc::mbar_wait_or_trap(&bar_tok[w], (uint32_t)i & 1u);
      {
        const uint4 wu = *reinterpret_cast<const uint4*>(wuni);
        float e0 = 0.f, e1 = 0.f;
        if (wu.x != 0xFFFFFFFFu && wu.x == wu.y && wu.x == wu.z && wu.x == wu.w) {
          // the whole tile goes to one cluster: register sums in row order
          const uint32_t l1 = wu.x;
          const float cf = rawc[l1 * kTcCStride + feat];
          float s0 = 0.f, s1 = 0.f;
#pragma unroll
          for (int ii = 0; ii < 16; ++ii) {
            const float x0 = *reinterpret_cast<const float*>(rw + 1024u * ii + acc_off_e);
            const float x1 = *reinterpret_cast<const float*>(rw + 1024u * ii + acc_off_o);
            const float d0 = x0 - cf, d1 = x1 - cf;
            s0 += x0; s1 += x1;
            e0 = fmaf(d0, d0, e0); e1 = fmaf(d1, d1, e1);
          }
          my_acc[l1 * 32] += s0 + s1;
        } else {
          // mixed labels (or a tail tile): runs of equal labels are summed in registers, then flushed
          uint32_t cur_lab = (uint32_t)k;   // dummy accumulator row
          float cur = 0.f, cf = 0.f;
#pragma unroll
          for (int qq = 0; qq < 8; ++qq) {
            const uint4 lv = *reinterpret_cast<const uint4*>(labs2 + sub * 32 + 4 * qq);
            const uint32_t l4[4] = {lv.x, lv.y, lv.z, lv.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int ii = 4 * qq + u;  // row sub + 4 ii
              if (sub + 4 * ii < rows) {
                const float xv = *reinterpret_cast<const float*>(rw + 1024u * (ii >> 1) + ((ii & 1) ? acc_off_o : acc_off_e));
                if (l4[u] != cur_lab) {
                  my_acc[cur_lab * 32] += cur;
                  cur_lab = l4[u];
                  cur = 0.f;
                  cf = rawc[cur_lab * kTcCStride + feat];
                }
                const float dv = xv - cf;
                cur += xv;
                e0 = fmaf(dv, dv, e0);
              }
            }
          }
          my_acc[cur_lab * 32] += cur;
        }
        inert_f += e0 + e1;
        if ((i & 7) == 7) { inert += (double)inert_f; inert_f = 0.f; }
      }
      tc::group_sync(bar_id);
      if (issuer) {
        tc::mbar_arrive(&bar_tok[(w + 1) % kT3Groups]);
        if (i + 2 < my_cnt) {
          tc::mbar_expect_tx(&bar_full[w][b], 16384u);
          tc::tma_load_2d(rw, &tmap, 0, (int)(tile_of(i + 2) * kTcRows), &bar_full[w][b]);
        }
      }
    }
  }
  __syncthreads();
  if (!done) {
    // ---- per-CTA partials (same layout as the other fused kernels) ----
    const int kd = k * 32, sstride = (k + 1) * 32;
    float* ps = a.part_sums + (size_t)blockIdx.x * kd;
    for (int e = t; e < kd; e += kT3Threads) {
      float s2 = acc[e];
#pragma unroll
      for (int qq = 1; qq < kTcSubsets; ++qq) s2 += acc[(size_t)qq * sstride + e];
      ps[e] = s2;
    }
    uint32_t* pc = a.part_counts + (size_t)blockIdx.x * k;
    for (int e = t; e < k; e += kT3Threads) pc[e] = cnt[e];
    double* red = reinterpret_cast<double*>(raw);
    red[t] = inert + (double)inert_f;
    __syncthreads();
    if (t < 128) red[t] += red[t + 128] + red[t + 256];
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if (t < o) red[t] += red[t + o];
      __syncthreads();
    }
    if (t == 0) a.part_inertia[blockIdx.x] = red[0];
    if (a.state) {
      if ((t & 127) == 0 && my_fallbacks) atomicAdd(&a.state->fallback_rows, (unsigned long long)my_fallbacks);
      if (t == 0 && abort_flag) atomicOr(&a.state->fallback_rows, 1ull << 40);
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, kT3TmemCols);
}

}  // namespace lkm

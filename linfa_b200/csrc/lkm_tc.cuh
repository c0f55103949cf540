// tcgen05 (5th-gen tensor core) building blocks for the f32 assignment path:
// scores S = X * C^T by 3xTF32 (hi*hi + hi*lo + lo*hi) with the accumulator in
// TMEM.  Raw PTX only; operands are staged in shared memory in the canonical
// K-major SWIZZLE_128B layout (8-row x 128-byte atoms, 16-byte chunk c of row
// r stored at chunk position c ^ (r & 7)).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace lkm {
namespace tc {

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- mbarrier ----------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// bounded wait: returns false if the phase did not complete within `spins` polls
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, uint32_t spins = 1u << 22) {
  const uint32_t addr = smem_u32(bar);
#pragma unroll 1
  for (uint32_t i = 0; i < spins; ++i) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return true;
  }
  return false;
}

// ---- proxies / fences ------------------------------------------------------------
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---- TMEM ---------------------------------------------------------------------------
// one full warp allocates `cols` (power of two >= 32) columns; the base address lands in *slot
__device__ __forceinline__ void tmem_alloc(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}

// ---- descriptors ---------------------------------------------------------------------
// shared-memory matrix descriptor, K-major, SWIZZLE_128B, dense 8-row groups (SBO = 1024 B)
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);         // start address, 16-byte units
  d |= (uint64_t)0 << 16;                            // leading byte offset: unused for K-major swizzled
  d |= (uint64_t)(1024u >> 4) << 32;                 // stride byte offset between 8-row groups
  d |= (uint64_t)1 << 46;                            // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                            // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor: kind::tf32, A and B K-major, f32 accumulate, M x N
__host__ __device__ constexpr uint32_t idesc_tf32(uint32_t m, uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread
__device__ __forceinline__ void mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// arrive on an mbarrier once every MMA issued so far by this thread has completed
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}

// 32 lanes x 32 consecutive columns -> 32 registers per thread (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
// the same wait, with the destination registers of an earlier tmem_ld32 as read-write operands: their uses cannot be
// scheduled above it (for loads that stay in flight across other work)
__device__ __forceinline__ void tmem_ld_fence(uint32_t (&r)[32]) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]), "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]), "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]), "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
               :
               : "memory");
}

// ---- operand staging -------------------------------------------------------------------
// byte offset of (row, 16-byte chunk) inside a K-major SW128 tile whose rows are 128 bytes
__device__ __forceinline__ uint32_t sw128_offset(uint32_t row, uint32_t chunk) {
  return (row >> 3) * 1024u + (row & 7u) * 128u + ((chunk ^ (row & 7u)) << 4);
}
// 3xTF32 split: hi = x with the low 13 mantissa bits cleared (exact TF32), lo = x - hi (exact in
// f32) truncated to TF32 as well, so both operands are exactly representable whatever the
// tensor core does with the low bits
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
  hi = __uint_as_float(__float_as_uint(x) & 0xFFFFE000u);
  lo = __uint_as_float(__float_as_uint(x - hi) & 0xFFFFE000u);
}

}  // namespace tc
// =============================================================================================
// Fused Lloyd iteration on tensor cores, f32, d = 32 (one 128-byte SW128 atom per row).
//
// Per 128-row tile: the scores S = X * C^T come from tcgen05 3xTF32 MMAs (accumulator in
// TMEM); each thread then owns one row: w_c = ||c||^2 - 2 S_c, running best / second best,
// and a CERTIFIED argmin — if (second - best) exceeds a rigorous bound on the rounding error of
// both this path and the reference's direct-form evaluation, the winner is the reference's
// winner; otherwise the row is queued and re-scanned in the reference's exact arithmetic
// (direct form, feature order, no FMA, strict '<').  The winner's distance is always
// re-evaluated in that exact form, so labels and min-distances are bit-identical to
// KM/algorithm.rs:923-946 + linfa-nn/src/distance.rs:68-70.  Sums / counts / inertia are then
// accumulated exactly as in lloyd_kernel (thread-owned shared-memory accumulators, fixed order).
// =============================================================================================
struct TcArgs {
  const float* X;
  unsigned long long n;
  int k;       // clusters
  int kp;      // k rounded up to a multiple of 32 (<= 256): UMMA N and TMEM columns used
  int tcols;   // TMEM columns to allocate (power of two >= kp)
  const float* C;
  float* part_sums;
  uint32_t* part_counts;
  double* part_inertia;
  LloydState* state;
  uint32_t* labels;  // optional global outputs (assign-only use)
  float* dists;
  int accumulate;    // 0: assignment only
};

constexpr int kTcRows = 128;
constexpr int kTcSubsets = 4;
constexpr int kTcCStride = 36;  // raw centroid row stride in words (16-byte aligned rows)

__host__ __device__ inline size_t tc_smem_bytes(int k, int kp, int accumulate) {
  size_t b = 0;
  b += 16384 * 2;                 // A_hi, A_lo
  b += (size_t)kp * 128 * 2;      // B_hi, B_lo
  b += 16384;                     // raw tile
  b += (size_t)k * kTcCStride * 4;  // raw centroids
  b += (size_t)kp * 4;            // ||c||^2
  if (accumulate) {
    b += (size_t)kTcSubsets * (k + 1) * 32 * 4;   // sums (+1 dummy row per subset)
    b += (((size_t)k * 4) + 15) & ~(size_t)15;    // counts
  }
  b += kTcRows * 4 * 4 + 16;      // labels, subset-major labels, dists, flag list
  return b + 1024;                // slack for the 1024-byte alignment of the operand tiles
}

__global__ void __launch_bounds__(128, 2) lloyd_tc_kernel(const TcArgs a) {
  pdl_launch_dependents();
  pdl_wait();
  if (a.state != nullptr && a.state->done) return;
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  // keep the shared address space visible to the compiler: offset, do not re-cast
  unsigned char* smem = smem_unaligned + ((1024u - (tc::smem_u32(smem_unaligned) & 1023u)) & 1023u);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int k = a.k, kp = a.kp;
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + 16384;
  unsigned char* b_hi = a_lo + 16384;
  unsigned char* b_lo = b_hi + (size_t)kp * 128;
  unsigned char* raw = b_lo + (size_t)kp * 128;
  float* rawc = reinterpret_cast<float*>(raw + 16384);
  float* cn = rawc + (size_t)k * kTcCStride;
  float* acc = cn + kp;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(acc + (a.accumulate ? (size_t)kTcSubsets * (k + 1) * 32 : 0));
  uint32_t* labs = cnt + (a.accumulate ? ((k + 3) & ~3) : 0);
  uint32_t* labs2 = labs + kTcRows;  // labs2[sub * 32 + i] = label of row sub + 4 i
  float* dist_s = reinterpret_cast<float*>(labs2 + kTcRows);
  uint32_t* flags = reinterpret_cast<uint32_t*>(dist_s + kTcRows);
  double* red = reinterpret_cast<double*>(a_hi);  // only used after the last tile
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t n_flag;
  __shared__ uint32_t cmax_bits;

  if (warp == 0) tc::tmem_alloc(&tmem_slot, (uint32_t)a.tcols);
  if (t == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); n_flag = 0; cmax_bits = 0; }
  __syncthreads();
  // ---- centroids: split into TF32 hi/lo operand tiles, padded raw copy, squared norms ----
  for (int r = t; r < kp; r += 128) {
    if (r < k) {
      double nn = 0.0;
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(a.C + (size_t)r * 32) + c);
        float4 h, l;
        tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
        tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
        const uint32_t o = tc::sw128_offset(r, c);
        *reinterpret_cast<float4*>(b_hi + o) = h;
        *reinterpret_cast<float4*>(b_lo + o) = l;
        *reinterpret_cast<float4*>(rawc + r * kTcCStride + 4 * c) = v;
        nn += (double)v.x * v.x + (double)v.y * v.y + (double)v.z * v.z + (double)v.w * v.w;
      }
      const float nf = (float)nn;
      cn[r] = nf;
      atomicMax(&cmax_bits, __float_as_uint(nf));  // non-negative floats order like their bits; a NaN norm wins and voids every certificate
    } else {
      const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t o = tc::sw128_offset(r, c);
        *reinterpret_cast<float4*>(b_hi + o) = z;
        *reinterpret_cast<float4*>(b_lo + o) = z;
      }
      cn[r] = INFINITY;  // padded centroid never wins
    }
  }
  if (a.accumulate) {
    for (int e = t; e < kTcSubsets * (k + 1) * 32; e += 128) acc[e] = 0.f;
    for (int e = t; e < k; e += 128) cnt[e] = 0u;
  }
  // ---- first tile into registers ----
  const unsigned long long n_tiles = (a.n + kTcRows - 1) / kTcRows;
  int4 pre[8];
  auto prefetch = [&](unsigned long long tile) {
    const unsigned long long row0 = tile * kTcRows;
    const int rows = (int)min((unsigned long long)kTcRows, a.n - row0);
    const int4* g = reinterpret_cast<const int4*>(a.X + row0 * 32) + t;
    const int my_row = t >> 3;
#pragma unroll
    for (int i = 0; i < 8; ++i)  // 16-byte chunk t + 128 i of the tile: row t/8 + 16 i, chunk t%8
      pre[i] = (my_row + 16 * i) < rows ? __ldg(g + 128 * i) : make_int4(0, 0, 0, 0);
  };
  if (blockIdx.x < n_tiles) prefetch(blockIdx.x);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const float cmax2 = __uint_as_float(cmax_bits);
  const float cmax = sqrtf(cmax2) * 1.000001f;
  // rigorous margin (see DESIGN.md "certified argmin"): eta bounds the 3xTF32 dot error relative to
  // |x||c|, gamma the reference's own direct-form rounding relative to the true squared distance
  const float eta = (3.f * 32.f + 12.f) * 2.3841858e-7f;  // (3d + 12) * 2^-22
  const float gamma = (32.f + 2.f) * 5.9604645e-8f;       // (d + 2) * 2^-24
  const float m_a0 = 2.3841858e-7f * cmax2;
  const float m_a1 = (4.f * eta + 4.7683716e-7f) * cmax;
  const float m_g2 = 2.f * gamma;

  // staging position of this thread's chunks: chunk t + 128 i -> st_off + 2048 i
  const uint32_t st_off = tc::sw128_offset((uint32_t)(t >> 3), (uint32_t)(t & 7));
  // this thread's row in the swizzled tile
  const uint32_t row_off = (uint32_t)(t >> 3) * 1024u + (uint32_t)(t & 7) * 128u;
  const uint32_t row_x = (uint32_t)(t & 7);
  // accumulation role: feature `feat` of rows sub, sub+4, ... ; two precomputed offsets (row&7 = sub / sub+4)
  const int sub = t >> 5, feat = t & 31;
  const uint32_t acc_off_e = (uint32_t)sub * 128u + ((((uint32_t)feat >> 2) ^ (uint32_t)sub) << 4) + ((uint32_t)feat & 3u) * 4u;
  const uint32_t acc_off_o = (uint32_t)(sub + 4) * 128u + ((((uint32_t)feat >> 2) ^ (uint32_t)(sub + 4)) << 4) + ((uint32_t)feat & 3u) * 4u;
  float* my_acc = acc + (size_t)sub * (k + 1) * 32 + feat;
  int cur_lab = k;  // dummy row: the first flush adds 0 to it
  float cur_acc = 0.f;
  double inert = 0.0;
  unsigned int my_fallbacks = 0;
  uint32_t parity = 0;

  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long row0 = tile * kTcRows;
    const int rows = (int)min((unsigned long long)kTcRows, a.n - row0);
    // ---- stage the tile: raw copy + TF32 hi/lo operand tiles, same swizzled position ----
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const uint32_t o = st_off + 2048u * i;
      const float4 v = *reinterpret_cast<const float4*>(&pre[i]);
      float4 h, l;
      h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
      h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
      h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
      h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
      *reinterpret_cast<float4*>(raw + o) = v;
      *reinterpret_cast<float4*>(a_hi + o) = h;
      *reinterpret_cast<float4*>(a_lo + o) = l;   // low bits beyond TF32 are dropped by the tensor core
    }
    if (tile + gridDim.x < n_tiles) prefetch(tile + gridDim.x);
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (t == 0) {
      tc::tc_fence_after();
      const uint32_t idesc = tc::idesc_tf32(128, (uint32_t)kp);
      const uint32_t sa_hi = tc::smem_u32(a_hi), sa_lo = tc::smem_u32(a_lo), sb_hi = tc::smem_u32(b_hi), sb_lo = tc::smem_u32(b_lo);
      uint32_t accf = 0;
#pragma unroll
      for (int pass = 0; pass < 3; ++pass) {
        const uint32_t sa = pass == 2 ? sa_lo : sa_hi;
        const uint32_t sb = pass == 1 ? sb_lo : sb_hi;
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          tc::mma_tf32(tmem, tc::smem_desc_sw128(sa + ks * 32), tc::smem_desc_sw128(sb + ks * 32), idesc, accf);
          accf = 1;
        }
      }
      tc::mma_commit(&bar);
    }
    // ---- this thread's row (exact data) while the tensor core works ----
    float xr[32];
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      const float4 v = *reinterpret_cast<const float4*>(raw + row_off + (((uint32_t)c ^ row_x) << 4));
      xr[4 * c] = v.x; xr[4 * c + 1] = v.y; xr[4 * c + 2] = v.z; xr[4 * c + 3] = v.w;
    }
    float xx0 = 0.f, xx1 = 0.f, xx2 = 0.f, xx3 = 0.f;
#pragma unroll
    for (int j = 0; j < 32; j += 4) {
      xx0 = fmaf(xr[j], xr[j], xx0); xx1 = fmaf(xr[j + 1], xr[j + 1], xx1);
      xx2 = fmaf(xr[j + 2], xr[j + 2], xx2); xx3 = fmaf(xr[j + 3], xr[j + 3], xx3);
    }
    const float xn = sqrtf((xx0 + xx1) + (xx2 + xx3)) * 1.000001f;
    const float margin = 1.1f * (m_a0 + xn * m_a1 + m_g2 * (xn + cmax) * (xn + cmax));

    const bool ok = tc::mbar_wait(&bar, parity);
    parity ^= 1u;
    tc::tc_fence_after();
    if (!ok) {  // never expected; leave a trace instead of hanging
      if (t == 0 && a.state) atomicAdd(&a.state->fallback_rows, 1ull << 40);
    }
    float best0 = INFINITY, second0 = INFINITY, best1 = INFINITY, second1 = INFINITY;
    int idx0 = 0, idx1 = 1;
    for (int c0 = 0; c0 < kp; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        const float4 cn4 = *reinterpret_cast<const float4*>(cn + c0 + j);
        const float w0 = fmaf(-2.f, __uint_as_float(r[j]), cn4.x);
        const float w1 = fmaf(-2.f, __uint_as_float(r[j + 1]), cn4.y);
        const float w2 = fmaf(-2.f, __uint_as_float(r[j + 2]), cn4.z);
        const float w3 = fmaf(-2.f, __uint_as_float(r[j + 3]), cn4.w);
        second0 = fminf(second0, fmaxf(w0, best0));
        idx0 = (w0 < best0) ? (c0 + j) : idx0;
        best0 = fminf(best0, w0);
        second1 = fminf(second1, fmaxf(w1, best1));
        idx1 = (w1 < best1) ? (c0 + j + 1) : idx1;
        best1 = fminf(best1, w1);
        second0 = fminf(second0, fmaxf(w2, best0));
        idx0 = (w2 < best0) ? (c0 + j + 2) : idx0;
        best0 = fminf(best0, w2);
        second1 = fminf(second1, fmaxf(w3, best1));
        idx1 = (w3 < best1) ? (c0 + j + 3) : idx1;
        best1 = fminf(best1, w3);
      }
    }
    // merge the two trackers (ties: lowest index)
    float best, second;
    int bidx;
    if (best1 < best0 || (best1 == best0 && idx1 < idx0)) {
      best = best1; bidx = idx1; second = fminf(second1, best0);
    } else {
      best = best0; bidx = idx0; second = fminf(second0, best1);
    }
    const bool row_ok = t < rows;
    const bool certified = (second - best) > margin;  // NaN anywhere -> not certified
    if (row_ok) {
      if (certified) {
        // the winner's distance in the reference's exact arithmetic
        const float* cw = rawc + bidx * kTcCStride;
        float dacc = 0.f;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 cv = *reinterpret_cast<const float4*>(cw + 4 * c);
          float df;
          df = __fsub_rn(cv.x, xr[4 * c]);     dacc = __fadd_rn(dacc, __fmul_rn(df, df));
          df = __fsub_rn(cv.y, xr[4 * c + 1]); dacc = __fadd_rn(dacc, __fmul_rn(df, df));
          df = __fsub_rn(cv.z, xr[4 * c + 2]); dacc = __fadd_rn(dacc, __fmul_rn(df, df));
          df = __fsub_rn(cv.w, xr[4 * c + 3]); dacc = __fadd_rn(dacc, __fmul_rn(df, df));
        }
        labs[t] = (uint32_t)bidx;
        dist_s[t] = dacc;
      } else {
        const uint32_t p = atomicAdd(&n_flag, 1u);
        flags[p] = (uint32_t)t;
      }
    }
    tc::tc_fence_before();
    __syncthreads();
    // ---- uncertified rows: exact scan, one warp per row, lanes split the centroids ----
    const uint32_t nf = n_flag;
    if (nf) {
      for (uint32_t f = warp; f < nf; f += 4) {
        const uint32_t row = flags[f];
        float x2[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const float4 v = *reinterpret_cast<const float4*>(raw + tc::sw128_offset(row, (uint32_t)c));
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
        if (lane == 0) { labs[row] = (uint32_t)fi; dist_s[row] = fb; }
      }
      my_fallbacks += (t == 0) ? nf : 0u;
      __syncthreads();
      if (t == 0) n_flag = 0;
    }
    if (row_ok) {
      const uint32_t lab = labs[t];
      labs2[(t & 3) * 32 + (t >> 2)] = lab;
      inert += (double)dist_s[t];
      if (a.accumulate) atomicAdd(&cnt[lab], 1u);  // integer atomics: exact, order-independent
      if (a.labels) a.labels[row0 + t] = lab;
      if (a.dists) a.dists[row0 + t] = dist_s[t];
    }
    // ---- deterministic accumulation: thread (sub, feat) walks rows sub, sub+4, ... in order ----
    if (a.accumulate) {
      __syncthreads();  // labs2 complete
      if (rows == kTcRows) {
        // the 32 labels of this subset (identical for the 32 lanes of the warp -> uniform branches)
        uint4 lv[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) lv[q] = *reinterpret_cast<const uint4*>(labs2 + sub * 32 + 4 * q);
        const uint32_t first = lv[0].x;
        uint32_t diff = 0;
#pragma unroll
        for (int q = 0; q < 8; ++q) diff |= (lv[q].x ^ first) | (lv[q].y ^ first) | (lv[q].z ^ first) | (lv[q].w ^ first);
        if (diff == 0) {
          // one label for the whole subset: sum its 32 rows in registers, in row order
          if ((int)first != cur_lab) { my_acc[cur_lab * 32] += cur_acc; cur_lab = (int)first; cur_acc = 0.f; }
#pragma unroll
          for (int ii = 0; ii < 16; ++ii) {
            cur_acc += *reinterpret_cast<const float*>(raw + 1024u * ii + acc_off_e);
            cur_acc += *reinterpret_cast<const float*>(raw + 1024u * ii + acc_off_o);
          }
        } else {
          // mixed labels: flush the run and add row by row into the thread-owned accumulators
          my_acc[cur_lab * 32] += cur_acc;
          cur_lab = k;
          cur_acc = 0.f;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const uint32_t l4[4] = {lv[q].x, lv[q].y, lv[q].z, lv[q].w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const int i = 4 * q + u;  // row sub + 4 i
              const float xv = *reinterpret_cast<const float*>(raw + 1024u * (i >> 1) + ((i & 1) ? acc_off_o : acc_off_e));
              my_acc[l4[u] * 32] += xv;
            }
          }
        }
      } else {
        for (int r = sub; r < rows; r += kTcSubsets) {
          const int lab = (int)labs[r];
          const float xv = *reinterpret_cast<const float*>(raw + tc::sw128_offset((uint32_t)r, (uint32_t)(feat >> 2)) + (feat & 3) * 4);
          if (lab != cur_lab) { my_acc[cur_lab * 32] += cur_acc; cur_lab = lab; cur_acc = 0.f; }
          cur_acc += xv;
        }
      }
    }
    __syncthreads();
  }

  if (a.accumulate) {
    my_acc[cur_lab * 32] += cur_acc;
    __syncthreads();
    const int kd = k * 32, sstride = (k + 1) * 32;
    float* ps = a.part_sums + (size_t)blockIdx.x * kd;
    for (int e = t; e < kd; e += 128) {
      float s = acc[e];
#pragma unroll
      for (int q = 1; q < kTcSubsets; ++q) s += acc[(size_t)q * sstride + e];
      ps[e] = s;
    }
    uint32_t* pc = a.part_counts + (size_t)blockIdx.x * k;
    for (int e = t; e < k; e += 128) pc[e] = cnt[e];
    __syncthreads();
    red[t] = inert;
    __syncthreads();
    for (int o = 64; o > 0; o >>= 1) {
      if (t < o) red[t] += red[t + o];
      __syncthreads();
    }
    if (t == 0) a.part_inertia[blockIdx.x] = red[0];
  }
  if (t == 0 && a.state && my_fallbacks) atomicAdd(&a.state->fallback_rows, (unsigned long long)my_fallbacks);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, (uint32_t)a.tcols);
}

// =============================================================================================
// Tensor-core ASSIGNMENT for wide rows / many clusters (f32, d = 32 * ka with ka in {1, 2, 4}, any k):
// labels and exact min distances to global memory; the centroid update then runs as the
// cluster-chunked MODE_UPDATE pass of lloyd_kernel.  Centroids are pre-split once per iteration
// into TF32 hi / lo operand tiles (tc_prep_centroids_kernel) and streamed from L2 in chunks of 64;
// the 128-row X tile stays resident as hi / lo (x = hi + lo exactly, lo is kept unmasked).
// =============================================================================================
__device__ __forceinline__ void cp_async16_zfill(void* smem_dst, const void* gsrc, bool valid) {
  const uint32_t d = tc::smem_u32(smem_dst);
  const uint32_t sz = valid ? 16u : 0u;
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(sz) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int kAsgChunk = 64;

struct TcAssignArgs {
  const float* X;
  unsigned long long n;
  int k, d, ka, n_chunks;
  const unsigned char* bsplit;  // [chunk][atom][hi 8 KB | lo 8 KB], rows in SW128 K-major order
  const float* cn;              // n_chunks * 64 squared norms (+inf for padding)
  const unsigned int* cmax_bits;
  const float* C;               // k x d, for the exact winner distance and the exact fallback
  uint32_t* labels;
  float* dists;
  LloydState* state;
};

// grid = n_chunks * ka CTAs of 64 threads: CTA (chunk, atom) writes one hi and one lo tile
__global__ void __launch_bounds__(64) tc_prep_centroids_kernel(const float* __restrict__ C, int k, int d, int ka,
                                                               unsigned char* __restrict__ bsplit, float* __restrict__ cn,
                                                               unsigned int* __restrict__ cmax_bits) {
  const int chunk = blockIdx.x / ka, atom = blockIdx.x % ka, r = threadIdx.x;
  const int c = chunk * kAsgChunk + r;
  unsigned char* hi = bsplit + ((size_t)chunk * ka + atom) * 16384;
  unsigned char* lo = hi + 8192;
  if (c < k) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(C + (size_t)c * d + atom * 32) + q);
      float4 h, l;
      tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
      tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
      const uint32_t o = tc::sw128_offset(r, q);
      *reinterpret_cast<float4*>(hi + o) = h;
      *reinterpret_cast<float4*>(lo + o) = l;
    }
  } else {
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const uint32_t o = tc::sw128_offset(r, q);
      *reinterpret_cast<float4*>(hi + o) = z;
      *reinterpret_cast<float4*>(lo + o) = z;
    }
  }
  if (atom == 0) {
    if (c < k) {
      double nn = 0.0;
      for (int j = 0; j < d; ++j) { const double v = (double)C[(size_t)c * d + j]; nn += v * v; }
      const float nf = (float)nn;
      cn[c] = nf;
      atomicMax(cmax_bits, __float_as_uint(nf));  // a NaN norm wins the max and voids every certificate
    } else {
      cn[c] = INFINITY;
    }
  }
}

__host__ __device__ inline size_t tc_assign_smem_bytes(int ka) {
  return (size_t)ka * 16384 * 2 /* A_hi, A_lo */ + (size_t)ka * 16384 /* B chunk */ + 256 /* cn chunk */ + 640 /* flags */ + 1024;
}

__global__ void __launch_bounds__(128) tc_assign_kernel(const TcAssignArgs a) {
  if (a.state != nullptr && a.state->done) return;
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (tc::smem_u32(smem_unaligned) & 1023u)) & 1023u);
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  const int ka = a.ka, d = a.d, k = a.k;
  unsigned char* a_hi = smem;
  unsigned char* a_lo = a_hi + (size_t)ka * 16384;
  unsigned char* bb = a_lo + (size_t)ka * 16384;   // [atom][hi 8 KB | lo 8 KB]
  float* cnb = reinterpret_cast<float*>(bb + (size_t)ka * 16384);
  uint32_t* flags = reinterpret_cast<uint32_t*>(cnb + kAsgChunk);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t n_flag;

  if (warp == 0) tc::tmem_alloc(&tmem_slot, 64);
  if (t == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); n_flag = 0; }
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const float cmax2 = __uint_as_float(*a.cmax_bits);
  const float cmax = sqrtf(cmax2) * 1.000001f;
  const float eta = (3.f * (float)d + 12.f) * 2.3841858e-7f;
  const float gamma = ((float)d + 2.f) * 5.9604645e-8f;
  const float m_a0 = 2.3841858e-7f * cmax2;
  const float m_a1 = (4.f * eta + 4.7683716e-7f) * cmax;
  const float m_g2 = 2.f * gamma;
  const uint32_t st_off = tc::sw128_offset((uint32_t)(t >> 3), (uint32_t)(t & 7));
  const uint32_t row_off = (uint32_t)(t >> 3) * 1024u + (uint32_t)(t & 7) * 128u;
  const uint32_t row_x = (uint32_t)(t & 7);
  const uint32_t idesc = tc::idesc_tf32(128, kAsgChunk);
  uint32_t parity = 0;
  unsigned int my_fallbacks = 0;

  const unsigned long long n_tiles = (a.n + 127) / 128;
  for (unsigned long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const unsigned long long row0 = tile * 128;
    const int rows = (int)min(128ull, a.n - row0);
    // ---- stage the X tile as TF32 hi / lo atoms (x = hi + lo exactly) ----
    for (int atom = 0; atom < ka; ++atom) {
      const float* g = a.X + row0 * (unsigned long long)d + atom * 32 + (t & 7) * 4;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int row = (t >> 3) + 16 * i;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (row < rows) v = __ldg(reinterpret_cast<const float4*>(g + (size_t)row * d));
        float4 h, l;
        h.x = __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u); l.x = v.x - h.x;
        h.y = __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u); l.y = v.y - h.y;
        h.z = __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u); l.z = v.z - h.z;
        h.w = __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u); l.w = v.w - h.w;
        const uint32_t o = (uint32_t)atom * 16384u + st_off + 2048u * i;
        *reinterpret_cast<float4*>(a_hi + o) = h;
        *reinterpret_cast<float4*>(a_lo + o) = l;
      }
    }
    float best0 = INFINITY, second0 = INFINITY, best1 = INFINITY, second1 = INFINITY;
    int idx0 = 0, idx1 = 1;
    for (int ch = 0; ch < a.n_chunks; ++ch) {
      // ---- centroid chunk: ka atoms of (hi | lo) tiles + 64 squared norms, flat copy from L2 ----
      const int4* src = reinterpret_cast<const int4*>(a.bsplit + (size_t)ch * ka * 16384);
      for (int q = t; q < ka * 1024; q += 128) cp_async16_zfill(bb + (size_t)q * 16, src + q, true);
      if (t < 16) cp_async16_zfill(cnb + t * 4, a.cn + (size_t)ch * kAsgChunk + t * 4, true);
      cp_async_commit();
      cp_async_wait_all();
      tc::fence_async_smem();
      tc::tc_fence_before();
      __syncthreads();
      if (t == 0) {
        tc::tc_fence_after();
        uint32_t accf = 0;
        for (int pass = 0; pass < 3; ++pass) {
          const uint32_t sa = tc::smem_u32(pass == 2 ? a_lo : a_hi);
          const uint32_t sb = tc::smem_u32(bb) + (pass == 1 ? 8192u : 0u);
          for (int atom = 0; atom < ka; ++atom) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) {
              tc::mma_tf32(tmem, tc::smem_desc_sw128(sa + atom * 16384 + ks * 32), tc::smem_desc_sw128(sb + atom * 16384 + ks * 32),
                           idesc, accf);
              accf = 1;
            }
          }
        }
        tc::mma_commit(&bar);
      }
      const bool ok = tc::mbar_wait(&bar, parity);
      parity ^= 1u;
      tc::tc_fence_after();
      if (!ok && t == 0 && a.state) atomicAdd(&a.state->fallback_rows, 1ull << 40);
      const int cbase = ch * kAsgChunk;
#pragma unroll
      for (int c0 = 0; c0 < kAsgChunk; c0 += 32) {
        uint32_t r[32];
        tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
        tc::tmem_ld_wait();
#pragma unroll
        for (int j = 0; j < 32; j += 4) {
          const float4 cn4 = *reinterpret_cast<const float4*>(cnb + c0 + j);
          const float w0 = fmaf(-2.f, __uint_as_float(r[j]), cn4.x);
          const float w1 = fmaf(-2.f, __uint_as_float(r[j + 1]), cn4.y);
          const float w2 = fmaf(-2.f, __uint_as_float(r[j + 2]), cn4.z);
          const float w3 = fmaf(-2.f, __uint_as_float(r[j + 3]), cn4.w);
          second0 = fminf(second0, fmaxf(w0, best0));
          idx0 = (w0 < best0) ? (cbase + c0 + j) : idx0;
          best0 = fminf(best0, w0);
          second1 = fminf(second1, fmaxf(w1, best1));
          idx1 = (w1 < best1) ? (cbase + c0 + j + 1) : idx1;
          best1 = fminf(best1, w1);
          second0 = fminf(second0, fmaxf(w2, best0));
          idx0 = (w2 < best0) ? (cbase + c0 + j + 2) : idx0;
          best0 = fminf(best0, w2);
          second1 = fminf(second1, fmaxf(w3, best1));
          idx1 = (w3 < best1) ? (cbase + c0 + j + 3) : idx1;
          best1 = fminf(best1, w3);
        }
      }
      tc::tc_fence_before();
      __syncthreads();  // chunk buffer and accumulator are free again
    }
    float best, second;
    int bidx;
    if (best1 < best0 || (best1 == best0 && idx1 < idx0)) { best = best1; bidx = idx1; second = fminf(second1, best0); }
    else { best = best0; bidx = idx0; second = fminf(second0, best1); }
    // ---- margin from ||x||, exact winner distance (direct form, feature order) ----
    float xx = 0.f;
    for (int atom = 0; atom < ka; ++atom) {
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint32_t o = (uint32_t)atom * 16384u + row_off + (((uint32_t)c ^ row_x) << 4);
        const float4 h = *reinterpret_cast<const float4*>(a_hi + o);
        const float4 l = *reinterpret_cast<const float4*>(a_lo + o);
        const float x0 = h.x + l.x, x1 = h.y + l.y, x2 = h.z + l.z, x3 = h.w + l.w;
        xx = fmaf(x0, x0, xx); xx = fmaf(x1, x1, xx); xx = fmaf(x2, x2, xx); xx = fmaf(x3, x3, xx);
      }
    }
    const float xn = sqrtf(xx) * 1.000001f;
    const float margin = 1.1f * (m_a0 + xn * m_a1 + m_g2 * (xn + cmax) * (xn + cmax));
    const bool row_ok = t < rows;
    const bool certified = (second - best) > margin;
    if (row_ok) {
      if (certified) {
        const float* cw = a.C + (size_t)bidx * d;
        float dacc = 0.f;
        for (int atom = 0; atom < ka; ++atom) {
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            const uint32_t o = (uint32_t)atom * 16384u + row_off + (((uint32_t)c ^ row_x) << 4);
            const float4 h = *reinterpret_cast<const float4*>(a_hi + o);
            const float4 l = *reinterpret_cast<const float4*>(a_lo + o);
            const float4 cv = __ldg(reinterpret_cast<const float4*>(cw + atom * 32 + 4 * c));
            float df;
            df = __fsub_rn(cv.x, h.x + l.x); dacc = __fadd_rn(dacc, __fmul_rn(df, df));
            df = __fsub_rn(cv.y, h.y + l.y); dacc = __fadd_rn(dacc, __fmul_rn(df, df));
            df = __fsub_rn(cv.z, h.z + l.z); dacc = __fadd_rn(dacc, __fmul_rn(df, df));
            df = __fsub_rn(cv.w, h.w + l.w); dacc = __fadd_rn(dacc, __fmul_rn(df, df));
          }
        }
        a.labels[row0 + t] = (uint32_t)bidx;
        a.dists[row0 + t] = dacc;
      } else {
        const uint32_t p = atomicAdd(&n_flag, 1u);
        flags[p] = (uint32_t)t;
      }
    }
    __syncthreads();
    // ---- uncertified rows: exact scan over all k centroids (from L2), one warp per row ----
    const uint32_t nf = n_flag;
    if (nf) {
      for (uint32_t f = warp; f < nf; f += 4) {
        const uint32_t row = flags[f];
        const uint32_t ro = (row >> 3) * 1024u + (row & 7u) * 128u, rx = row & 7u;
        float fb = INFINITY;
        int fi = -1;
        for (int c = lane; c < k; c += 32) {
          const float* cw = a.C + (size_t)c * d;
          float dacc = 0.f;
          for (int atom = 0; atom < ka; ++atom) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
              const uint32_t o = (uint32_t)atom * 16384u + ro + (((uint32_t)q ^ rx) << 4);
              const float4 h = *reinterpret_cast<const float4*>(a_hi + o);
              const float4 l = *reinterpret_cast<const float4*>(a_lo + o);
              const float4 cv = __ldg(reinterpret_cast<const float4*>(cw + atom * 32 + 4 * q));
              float df;
              df = __fsub_rn(cv.x, h.x + l.x); dacc = __fadd_rn(dacc, __fmul_rn(df, df));
              df = __fsub_rn(cv.y, h.y + l.y); dacc = __fadd_rn(dacc, __fmul_rn(df, df));
              df = __fsub_rn(cv.z, h.z + l.z); dacc = __fadd_rn(dacc, __fmul_rn(df, df));
              df = __fsub_rn(cv.w, h.w + l.w); dacc = __fadd_rn(dacc, __fmul_rn(df, df));
            }
          }
          // lane 0 starts with centroid 0 (sticky NaN like the reference); the others only take v < best
          if (c == 0) { fb = dacc; fi = 0; }
          else if (dacc < fb) { fb = dacc; fi = c; }
        }
        combine_candidates<float>(32, fb, fi);
        if (lane == 0) { a.labels[row0 + row] = (uint32_t)fi; a.dists[row0 + row] = fb; }
      }
      my_fallbacks += (t == 0) ? nf : 0u;
    }
    __syncthreads();
    if (t == 0) n_flag = 0;
  }
  if (t == 0 && a.state && my_fallbacks) atomicAdd(&a.state->fallback_rows, (unsigned long long)my_fallbacks);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 64);
}

namespace tc {
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
}  // namespace tc

}  // namespace lkm

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

// ---------------------------------------------------------------------------------------------
// Debug / validation kernel: S[128 x N] = X[128 x 32] * C[N x 32]^T by 3xTF32 on tcgen05, one
// CTA of 128 threads.  status[0] = 0 ok, 1 = mbarrier wait timed out.
// ---------------------------------------------------------------------------------------------
template <int N>
__global__ void __launch_bounds__(128) tc_scores_debug_kernel(const float* __restrict__ X, const float* __restrict__ Cn,
                                                              float* __restrict__ S, int* __restrict__ status) {
  static_assert(N % 16 == 0 && N >= 16 && N <= 256, "UMMA N for M = 128");
  extern __shared__ __align__(1024) unsigned char smem[];
  // [A_hi 16K][A_lo 16K][B_hi N*128][B_lo N*128]
  unsigned char* a_hi = smem;
  unsigned char* a_lo = smem + 16384;
  unsigned char* b_hi = smem + 32768;
  unsigned char* b_lo = b_hi + N * 128;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  constexpr uint32_t TCOLS = N < 32 ? 32 : (N <= 32 ? 32 : (N <= 64 ? 64 : (N <= 128 ? 128 : 256)));

  if (warp == 0) tc::tmem_alloc(&tmem_slot, TCOLS);
  if (t == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  // stage operands: thread t owns row t of X (and row t of C if t < N)
  for (int c = 0; c < 8; ++c) {
    const float4 v = reinterpret_cast<const float4*>(X + (size_t)t * 32)[c];
    float4 h, l;
    tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
    tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
    const uint32_t o = tc::sw128_offset(t, c);
    *reinterpret_cast<float4*>(a_hi + o) = h;
    *reinterpret_cast<float4*>(a_lo + o) = l;
  }
  for (int r = t; r < N; r += 128) {
    for (int c = 0; c < 8; ++c) {
      const float4 v = reinterpret_cast<const float4*>(Cn + (size_t)r * 32)[c];
      float4 h, l;
      tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
      tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
      const uint32_t o = tc::sw128_offset(r, c);
      *reinterpret_cast<float4*>(b_hi + o) = h;
      *reinterpret_cast<float4*>(b_lo + o) = l;
    }
  }
  tc::fence_async_smem();   // generic-proxy stores -> visible to the tensor core (async proxy)
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    constexpr uint32_t idesc = tc::idesc_tf32(128, N);
    const uint32_t sa_hi = tc::smem_u32(a_hi), sa_lo = tc::smem_u32(a_lo), sb_hi = tc::smem_u32(b_hi), sb_lo = tc::smem_u32(b_lo);
    uint32_t acc = 0;
    for (int pass = 0; pass < 3; ++pass) {
      const uint32_t sa = pass == 2 ? sa_lo : sa_hi;
      const uint32_t sb = pass == 1 ? sb_lo : sb_hi;
      for (int ks = 0; ks < 4; ++ks) {  // K = 8 TF32 (32 bytes) per instruction
        tc::mma_tf32(tmem, tc::smem_desc_sw128(sa + ks * 32), tc::smem_desc_sw128(sb + ks * 32), idesc, acc);
        acc = 1;
      }
    }
    tc::mma_commit(&bar);
  }
  const bool ok = tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  if (!ok) {
    if (t == 0) status[0] = 1;
  } else {
    // warp w reads TMEM lanes 32w..32w+31 (= rows), 32 columns at a time
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j)
        if (c0 + j < N) S[(size_t)t * N + c0 + j] = __uint_as_float(r[j]);
    }
    if (t == 0) status[0] = 0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, TCOLS);
}

}  // namespace lkm

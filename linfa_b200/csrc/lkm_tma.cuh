// TMA / mbarrier helpers shared by the tcgen05 kernels (lkm_tc4.cuh, lkm_tc5.cuh, lkm_upd.cuh) and the
// max-row-norm pass that sizes their score window.
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

}  // namespace lkm

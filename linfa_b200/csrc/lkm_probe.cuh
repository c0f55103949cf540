// Development probes and micro-benchmarks (tcgen05 operand layouts, rounding behaviour, instruction rates).
// NOT part of the product library: compiled only into liblkm_probe.so (-DLKM_WITH_PROBES), which
// tests/test_gpu_tc_probe.py and tools/*.py load.  Findings: profiles/umma_layout_probes.md.
#pragma once
#include "lkm_tc4.cuh"

namespace lkm {

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

// ---------------------------------------------------------------------------------------------
// Probe for the tensor-core centroid accumulation: D2[M x 32] = onehot^T[M x 128] * X[128 x 32]
// (M = 64 or 128 clusters).  A = one-hot, MN-major (cluster contiguous per data row), groups of 32
// clusters 16 KB apart; B = the X tile as staged for the assignment MMA, read MN-major (feature
// contiguous per data row).  Dumps all 128 TMEM lanes x 32 columns so that the host can learn the
// accumulator layout.  status: 0 ok, 1 timeout.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t smem_desc_mn_sw128(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16;   // stride between 128-byte MN groups
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32;   // stride between 8-row K groups
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__host__ __device__ constexpr uint32_t idesc_tf32_mn(uint32_t m, uint32_t n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (1u << 15) | (1u << 16) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

template <int M>
__global__ void __launch_bounds__(128) tc_accum_debug_kernel(const float* __restrict__ X, const int* __restrict__ labels,
                                                             float* __restrict__ dump, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (tc::smem_u32(smem_unaligned) & 1023u)) & 1023u);
  unsigned char* xt = smem;              // [128 rows][128 B] SW128
  unsigned char* oh = smem + 16384;      // M/32 groups of [128 rows][128 B] SW128
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 32);
  if (t == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  for (int i = t; i < (M / 32) * 16384 / 16; i += 128) reinterpret_cast<int4*>(oh)[i] = make_int4(0, 0, 0, 0);
  for (int c = 0; c < 8; ++c)
    *reinterpret_cast<float4*>(xt + tc::sw128_offset(t, c)) = reinterpret_cast<const float4*>(X + (size_t)t * 32)[c];
  __syncthreads();
  {
    const int lab = labels[t];
    const uint32_t g = (uint32_t)lab >> 5, w = (uint32_t)lab & 31u;
    *reinterpret_cast<float*>(oh + g * 16384u + tc::sw128_offset(t, w >> 2) + (w & 3u) * 4u) = 1.0f;
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    constexpr uint32_t idesc = idesc_tf32_mn(M, 32);
    uint32_t acc = 0;
    for (int ks = 0; ks < 16; ++ks) {  // 8 data rows per MMA
      tc::mma_tf32(tmem, smem_desc_mn_sw128(tc::smem_u32(oh) + ks * 1024, 16384, 1024),
                   smem_desc_mn_sw128(tc::smem_u32(xt) + ks * 1024, 16384, 1024), idesc, acc);
      acc = 1;
    }
    tc::mma_commit(&bar);
  }
  const bool ok = tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  if (ok) {
    uint32_t r[32];
    tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16), r);
    tc::tmem_ld_wait();
#pragma unroll
    for (int j = 0; j < 32; ++j) dump[(size_t)t * 32 + j] = __uint_as_float(r[j]);
  }
  if (t == 0) status[0] = ok ? 0 : 1;
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 32);
}

// Raw UMMA probe: the host supplies byte images of the A and B operand regions, the two
// descriptors (without start address), the per-k-step byte advance and the instruction
// descriptor; the kernel copies the images to 1024-aligned shared memory, issues `ksteps` MMAs and
// dumps 128 TMEM lanes x `ncols` columns.  Used to pin down operand layouts from Python.
__global__ void __launch_bounds__(128) umma_raw_probe_kernel(const unsigned char* __restrict__ a_img, int a_bytes,
                                                             const unsigned char* __restrict__ b_img, int b_bytes,
                                                             unsigned long long a_desc_hi, unsigned long long b_desc_hi,
                                                             int a_step, int b_step, unsigned int idesc, int ksteps,
                                                             int ncols, float* __restrict__ dump, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (tc::smem_u32(smem_unaligned) & 1023u)) & 1023u);
  unsigned char* sa = smem;
  unsigned char* sb = smem + ((a_bytes + 1023) & ~1023);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
  if (t == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  for (int i = t; i < a_bytes / 16; i += 128) reinterpret_cast<int4*>(sa)[i] = reinterpret_cast<const int4*>(a_img)[i];
  for (int i = t; i < b_bytes / 16; i += 128) reinterpret_cast<int4*>(sb)[i] = reinterpret_cast<const int4*>(b_img)[i];
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    uint32_t acc = 0;
    for (int ks = 0; ks < ksteps; ++ks) {
      const uint64_t da = a_desc_hi | (uint64_t)(((tc::smem_u32(sa) + ks * a_step) & 0x3FFFFu) >> 4);
      const uint64_t db = b_desc_hi | (uint64_t)(((tc::smem_u32(sb) + ks * b_step) & 0x3FFFFu) >> 4);
      tc::mma_tf32(tmem, da, db, idesc, acc);
      acc = 1;
    }
    tc::mma_commit(&bar);
  }
  const bool ok = tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  if (ok) {
    for (int c0 = 0; c0 < ncols; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) dump[(size_t)t * ncols + c0 + j] = __uint_as_float(r[j]);
    }
  }
  if (t == 0) status[0] = ok ? 0 : 1;
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}

// Micro-benchmark: one thread per CTA issues `reps` batches of `per_batch` tcgen05 TF32 MMAs (M = 128, N = n_cols,
// K = 8) on zeroed shared-memory operands, one commit + wait per batch; out[block] = cycles.  mode bit 0: all
// MMAs of a batch read the same A K-step (operand fetch hits one 4 KB region); bit 1: use the SBO = 0 alias for A.
__global__ void __launch_bounds__(128) mma_bench_kernel(int n_cols, int reps, int per_batch, int mode, long long* out) {
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (tc::smem_u32(smem_unaligned) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  unsigned char* a_t = smem;            // 2 x 16 KB
  unsigned char* b_t = smem + 32768;    // n_cols x 128 B
  for (int i = t; i < (32768 + n_cols * 128) / 16; i += 128) reinterpret_cast<int4*>(smem)[i] = make_int4(0, 0, 0, 0);
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 256);
  if (t == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  if (t == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, (uint32_t)n_cols);
    const uint64_t da = (mode & 2) ? tc::smem_desc_sw128_alias(tc::smem_u32(a_t)) : tc::smem_desc_sw128(tc::smem_u32(a_t));
    const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(b_t));
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      for (int i = 0; i < per_batch; ++i) {
        const uint32_t ks = (mode & 1) ? 0u : (uint32_t)(i & 3);
        const uint64_t a_off = ((i >> 2) & 1) ? 1024u : 0u;
        if (mode & 8) tc::mma_f16(tmem, da + a_off + 2u * ks, db + 2u * ks, tc::idesc_f16(128, (uint32_t)n_cols, 0), i ? 1u : 0u);
        else if (mode & 4) tc::mma_tf32_ts(tmem, tmem + 128u + 8u * ks + (((i >> 2) & 1) ? 32u : 0u), db + 2u * ks, idesc, i ? 1u : 0u);
        else tc::mma_tf32(tmem, da + a_off + 2u * ks, db + 2u * ks, idesc, i ? 1u : 0u);
      }
      tc::mma_commit(&bar);
      tc::mbar_wait_or_trap(&bar, (uint32_t)r & 1u);
    }
    out[blockIdx.x] = clock64() - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 256);
}


// Same measurement with the issue pattern of a production MMA warp: warp 0 runs the loop uniformly, 13 unrolled
// MMAs per batch with constant descriptor offsets, elect inside the instruction wrapper.
template <int MODE>  // 0 tf32 SS, 1 tf32 TS, 2 f16 SS, 3 f16 TS
__global__ void __launch_bounds__(128) mma_bench2_kernel(int n_cols, int reps, long long* out) {
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (tc::smem_u32(smem_unaligned) & 1023u)) & 1023u);
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  for (int i = t; i < (32768 + n_cols * 128) / 16; i += 128) reinterpret_cast<int4*>(smem)[i] = make_int4(0, 0, 0, 0);
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  if (t == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, tmem_slot, 0);
  if (warp == 0) {
    const uint32_t idesc = tc::idesc_tf32(128, (uint32_t)n_cols);
    const uint64_t da = tc::smem_desc_sw128(tc::smem_u32(smem));
    const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(smem + 32768));
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int i = 0; i < 13; ++i) {
        if (MODE == 0) tc::mma_tf32_elect(tmem, da + (uint64_t)(((i >> 2) & 1) * 1024 + 2 * (i & 3)), db + (uint64_t)(2 * (i & 3)), idesc, i ? 1u : 0u);
        if (MODE == 1) tc::mma_tf32_ts_elect(tmem, tmem + 256u + 32u * ((i >> 2) & 1) + 8u * (i & 3), db + (uint64_t)(2 * (i & 3)), idesc, i ? 1u : 0u);
        if (MODE == 2) tc::mma_f16_elect(tmem, da + (uint64_t)(((i >> 2) & 1) * 1024 + 2 * (i & 3)), db + (uint64_t)(2 * (i & 3)), tc::idesc_f16(128, (uint32_t)n_cols, 0), i ? 1u : 0u);
        if (MODE == 3) tc::mma_f16_ts_elect(tmem, tmem + 256u + 32u * ((i >> 2) & 1) + 8u * (i & 3), db + (uint64_t)(2 * (i & 3)), tc::idesc_f16(128, (uint32_t)n_cols, 0), i ? 1u : 0u);
      }
      tc::mma_commit_elect(&bar);
      tc::mbar_wait_or_trap(&bar, (uint32_t)r & 1u);
    }
    if (t == 0) out[blockIdx.x] = clock64() - t0;
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// TMEM load throughput: `nwarps` warps each issue `reps` x 8 tcgen05.ld.32x32b.x32 (4 KB each) back to back.
__global__ void __launch_bounds__(256) ldtm_bench_kernel(int nwarps, int reps, long long* out) {
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t sink = 0;
  if (warp < nwarps) {
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        uint32_t v[32];
        tc::tmem_ld32(base + 32u * (uint32_t)i, v);
        tc::tmem_ld_wait();
        sink += v[0] ^ v[31];
      }
    }
    const long long t1 = clock64();
    if ((t & 31) == 0) out[blockIdx.x * 8 + warp] = t1 - t0 + (sink == 0x12345u ? 1 : 0);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// TMEM store throughput (what = 1): `nwarps` warps each issue `reps` x 8 tcgen05.st.32x32b.x32 back to back, one
// tcgen05.wait::st per 8; what = 2: warps 0-3 store while warps 4-7 load (the S / E mix of the pair kernel).
__global__ void __launch_bounds__(256) sttm_bench_kernel(int nwarps, int reps, int what, long long* out) {
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 512);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  uint32_t sink = 0;
  if (warp < nwarps) {
    const bool storing = what == 1 || warp < 4;
    const uint32_t base = tmem + ((uint32_t)((warp & 3) * 32) << 16) + (storing ? 256u : 0u);
    uint32_t v[32];
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = (uint32_t)(t * 32 + i);
    const long long t0 = clock64();
    for (int r = 0; r < reps; ++r) {
      if (storing) {
#pragma unroll
        for (int i = 0; i < 8; ++i) tc::tmem_st32(base + 32u * (uint32_t)i, v);
        tc::tmem_st_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint32_t w[32];
          tc::tmem_ld32(base + 32u * (uint32_t)i, w);
          tc::tmem_ld_wait();
          sink += w[0] ^ w[31];
        }
      }
    }
    const long long t1 = clock64();
    if ((t & 31) == 0) out[blockIdx.x * 8 + warp] = t1 - t0 + (sink == 0x12345u ? 1 : 0);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 512);
}

// Probe for A-from-TMEM MMAs: thread t stores row t of X (32 f32, raw bits) into TMEM columns 64..95 of its lane,
// then S[128 x N] = X * Cn^T with four TS-mode MMAs (A K-step ks = columns 64 + 8 ks), B = Cn from shared memory.
template <int N>
__global__ void __launch_bounds__(128) tc_ts_probe_kernel(const float* __restrict__ X, const float* __restrict__ Cn,
                                                          float* __restrict__ S, int* __restrict__ status) {
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (tc::smem_u32(smem_unaligned) & 1023u)) & 1023u);
  unsigned char* b_t = smem;
  __shared__ __align__(8) uint64_t bar;
  __shared__ uint32_t tmem_slot;
  const int t = threadIdx.x, warp = t >> 5;
  if (warp == 0) tc::tmem_alloc(&tmem_slot, 128);
  if (t == 0) { tc::mbar_init(&bar, 1); tc::fence_mbar_init(); }
  for (int r = t; r < N; r += 128)
    for (int c = 0; c < 8; ++c)
      *reinterpret_cast<float4*>(b_t + tc::sw128_offset(r, c)) = reinterpret_cast<const float4*>(Cn + (size_t)r * 32)[c];
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  {
    uint32_t r[32];
#pragma unroll
    for (int j = 0; j < 32; ++j) r[j] = __float_as_uint(X[(size_t)t * 32 + j]);
    tc::tmem_st32(tmem + ((uint32_t)(warp * 32) << 16) + 64u, r);
    tc::tmem_st_wait();
  }
  tc::tc_fence_before();
  __syncthreads();
  if (t == 0) {
    tc::tc_fence_after();
    constexpr uint32_t idesc = tc::idesc_tf32(128, N);
    const uint64_t db = tc::smem_desc_sw128(tc::smem_u32(b_t));
    for (int ks = 0; ks < 4; ++ks) tc::mma_tf32_ts(tmem, tmem + 64u + 8u * ks, db + 2u * ks, idesc, ks ? 1u : 0u);
    tc::mma_commit(&bar);
  }
  const bool ok = tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  if (ok) {
    for (int c0 = 0; c0 < N; c0 += 32) {
      uint32_t r[32];
      tc::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, r);
      tc::tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) S[(size_t)t * N + c0 + j] = __uint_as_float(r[j]);
    }
  }
  if (t == 0) status[0] = ok ? 0 : 1;
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, 128);
}


}  // namespace lkm

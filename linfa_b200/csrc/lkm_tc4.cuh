// Fused Lloyd iteration for f32, d = 32, k <= 64 — warp-specialised tcgen05 pipeline (cfg2 shape).
//
// One persistent CTA per SM, 18 warps in five roles that work on different tiles at the same time
// and hand 128-row tiles to each other through mbarriers:
//
//   P  warp 16   one thread keeps the ring of raw tiles full with TMA (SWIZZLE_128B, so that row-wise and
//                feature-wise shared-memory reads are both conflict-free)
//   S  warps 0-3 thread = row: copies the row into TMEM as the A operand of tile j — once as raw bits (kind::tf32
//                ignores the 13 low mantissa bits, so raw = x_hi) and once as x_lo = x - trunc_tf32(x)
//   M  warp 17   issues the 13 tcgen05 MMAs of tile j into TMEM accumulator j & 1, A from TMEM, B (centroid
//                hi / lo / bias tiles) from shared memory:
//                x_lo.c_hi + x_hi.c_lo + x_hi.c_hi + ones.(K0 - |c|^2/2)  ->  v_c = x.c - |c|^2/2 + K0
//                (A in TMEM because a 128 x 8 TF32 A operand read from shared memory costs 4 KB of its
//                128 B/clk per MMA: measured 47 cycles per N = 64 MMA from shared memory, 32 from TMEM)
//   E  warps 4-7 (even tiles) and 8-11 (odd tiles), thread = row: packed integer top-2 of the 64
//                scores (q_c = (bits(v_c) << 6) | c is monotone in v_c and carries the index), certificate, label / "uncertified" into a label slot
//   U  warps 12-15, thread = (row subset, feature): exact re-scan of uncertified rows, then sums
//                and inertia straight from the raw tile into thread-owned shared-memory
//                accumulators, tiles in ascending order (deterministic); frees the raw stage
//
// Tiles j+3.. are in flight from HBM while S works on j+2, the tensor core on j+1, E on j and U on
// j-1.  Outputs: per-CTA partial sums / counts / inertia for finalize_kernel, as the other kernels.
#pragma once
#include "lkm_tma.cuh"

namespace lkm {

constexpr int kT4Threads = 18 * 32;
constexpr int kT4TmemCols = 512;   // D 2 x 64 | A_hi 2 x 32 | A_lo 2 x 32 | ones 8
constexpr uint32_t kT4ColAhi = 128, kT4ColAlo = 192, kT4ColOnes = 256;
constexpr int kT4MaxStages = 12;
#ifndef LKM_T4_PREFETCH
#define LKM_T4_PREFETCH 0
#endif
constexpr unsigned long long kT4Prefetch = LKM_T4_PREFETCH;   // L2 prefetch distance of the P role, in tiles beyond the ring's newest request
constexpr int kNbLab = 2, kNbLfree = 4, kNbAlo = 6, kNbTfree = 8;   // named barrier ids (+ parity); 1 = U-internal
constexpr int kT4Subsets = 8;      // U role: thread = (row subset of 8, feature pair)

__host__ __device__ inline size_t tc4_smem_bytes(int k, int kp, int stages, int passes = 3) {
  size_t b = 0;
  b += (size_t)stages * 16384;                   // raw tiles (TMA destinations)
  b += (size_t)kp * 128 * (passes == 1 ? 2 : 3); // B_hi, B_lo (not in the one-product variant: one more ring stage), B_bias
  b += (size_t)k * kTcCStride * 4;               // raw centroids
  b += (size_t)kT4Subsets * (k + 1) * 32 * 4;    // sums
  b += (((size_t)k * 4) + 15) & ~(size_t)15;     // counts
  b += (size_t)kp * 8;                           // centroid squared norms (f64)
  b += (size_t)2 * kTcRows * 4 * 2;              // per label slot: subset-major labels, flag list
  b += (size_t)2 * 16;                           // per label slot: warp-uniform labels
  return b + 1024;
}
__host__ __device__ inline int tc4_stages(int k, int kp, size_t smem_limit, int passes = 3) {
  int s = kT4MaxStages;
  while (s > 3 && tc4_smem_bytes(k, kp, s, passes) > smem_limit) --s;
  return s;
}

struct Tc4Args {
  const float* X;
  unsigned long long n;
  int k;
  int stages;
  const float* C;
  const unsigned int* xmax2_bits;  // row_norm2_max_kernel's result
  float* part_sums;
  uint32_t* part_counts;
  double* part_inertia;
  LloydState* state;
  long long* prof;   // LKM_T4_PROF builds only
  uint32_t pack_mult;   // 64, passed as an argument so that the score pack of the E role stays an IMAD (FMA pipe)
  // in-kernel finalize (single GPU; fin_sync == nullptr: finalize_kernel follows as a separate launch)
  unsigned int* fin_sync;   // [0] grid rendezvous, [1] last-CTA detection; zero between launches
  float* C_new;
  double* counts_out;       // k, members per cluster of this assignment (without the +1)
  double* shift_part;       // [grid]
  double tol;
  unsigned long long max_iter;
  uint32_t* labels_out;     // optional (lkm_set_keep_labels): the label every row was accumulated under, for the parity tests
};

// Optional phase profile (compile with -DLKM_T4_PROF): cycles spent in each wait of each role, written
// by one lane per role for CTA 0 into a.prof[role * 8 + slot]; slot 7 = the role's total loop time.
#ifdef LKM_T4_PROF
#define T4WAIT(slot, bar, ph) do { const long long t0__ = clock64(); tc::mbar_wait_or_trap(bar, ph); prof_acc[slot] += clock64() - t0__; } while (0)
#else
#define T4WAIT(slot, bar, ph) tc::mbar_wait_or_trap(bar, ph)
#endif
// LKM_T4_TRACE: CTA 0 stamps event `ev` of tile `j` (cycles since the roles started) into a.prof[64 + j * 16 + ev]
#ifdef LKM_T4_TRACE
#define T4TRACE(j, ev) do { if (blockIdx.x == 0 && lane == 0 && (j) < 40 && a.prof) a.prof[64 + (j) * 16 + (ev)] = clock64() - trace_t0; } while (0)
#else
#define T4TRACE(j, ev)
#endif
#ifdef LKM_T4_PROF
#define T4MARK(var) const long long var = clock64()
#define T4ADD(slot, a, b) prof_acc[slot] += (b) - (a)
#else
#define T4MARK(var)
#define T4ADD(slot, a, b)
#endif

namespace tc {
// packed pairs of f32 (Blackwell FADD2 / FFMA2): two IEEE operations per lane and instruction
__device__ __forceinline__ unsigned long long f2_add(unsigned long long a, unsigned long long b) {
  unsigned long long d;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
  return d;
}
__device__ __forceinline__ unsigned long long f2_fma(unsigned long long a, unsigned long long b, unsigned long long c) {
  unsigned long long d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
// (d0, d1) = (a0, a1) + (b0, b1) as one FADD2 on explicit 32-bit halves (keeps ptxas from doing 64-bit integer
// arithmetic on the packed value)
__device__ __forceinline__ void f2_add_parts(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, uint32_t& d0, uint32_t& d1) {
  asm("{\n\t.reg .b64 a, b, d;\n\t"
      "mov.b64 a, {%2, %3};\n\t"
      "mov.b64 b, {%4, %5};\n\t"
      "add.rn.f32x2 d, a, b;\n\t"
      "mov.b64 {%0, %1}, d;\n\t}"
      : "=r"(d0), "=r"(d1)
      : "r"(a0), "r"(a1), "r"(b0), "r"(b1));
}
// -(x with the 13 low mantissa bits cleared): (x & mask) ^ sign in one LOP3
__device__ __forceinline__ uint32_t neg_trunc_tf32_bits(uint32_t x, uint32_t mask, uint32_t sign) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0x6A;" : "=r"(d) : "r"(x), "r"(mask), "r"(sign));
  return d;
}
__device__ __forceinline__ unsigned long long f2_pack(float lo, float hi) {
  return (unsigned long long)__float_as_uint(lo) | ((unsigned long long)__float_as_uint(hi) << 32);
}
__device__ __forceinline__ float f2_lo(unsigned long long v) { return __uint_as_float((uint32_t)v); }
__device__ __forceinline__ float f2_hi(unsigned long long v) { return __uint_as_float((uint32_t)(v >> 32)); }
// Named-barrier hand-offs between roles: the consumer blocks in the barrier unit (no polling, no issue
// slots), the producer only arrives.  `threads` = producer threads + consumer threads.
__device__ __forceinline__ void nb_sync(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
__device__ __forceinline__ void nb_arrive(int id, int threads) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(threads) : "memory"); }
// one arrival per warp: every lane has finished its part (memory ordered by __syncwarp)
__device__ __forceinline__ void warp_arrive(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
// D[tmem] (+)= A[tmem] * B[smem]^T : A = M x 8 TF32 in 8 consecutive TMEM columns (lane = row); one thread
__device__ __forceinline__ void mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// kind::f16 (A, B = f16 or bf16 in shared memory, K = 16 per instruction, f32 accumulate)
__host__ __device__ constexpr uint32_t idesc_f16(uint32_t m, uint32_t n, uint32_t ab_format /* 0 = f16, 1 = bf16 */) {
  return (1u << 4) | (ab_format << 7) | (ab_format << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}
__device__ __forceinline__ void mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// Warp-uniform issue: the whole warp executes this, one elected lane issues the MMA.  Keeps the operands in
// uniform registers (no per-instruction R2UR / ELECT loop as under `if (lane == 0)`).
__device__ __forceinline__ void mma_tf32_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void mma_tf32_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void mma_f16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void mma_f16_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void mma_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
// 32 registers per thread -> 32 lanes x 32 consecutive columns (thread i <-> lane base+i)
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};"
      ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]),
        "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]),
        "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]),
        "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
}  // namespace tc

// PASSES = 3: x_lo.c_hi + x_hi.c_lo + x_hi.c_hi (3xTF32, A operands staged in TMEM by the S role).
// PASSES = 1: x_hi.c_hi only (screening: 5 MMAs per tile instead of 13; same structure as PASSES = 2).
// PASSES = 2: x_hi.c_lo + x_hi.c_hi only — the raw tile in shared memory is the A operand, the S role is idle, and the
// dropped x_lo.c term (|x_lo.c| <= 2^-10 |x||c|) widens the certificate; rows it no longer certifies take the exact
// re-scan, so results are identical.  The host starts with 2 passes and switches to 3 when too many rows fall back.
template <int KP, int PASSES>
__global__ void __launch_bounds__(kT4Threads, 1) lloyd_tc4_kernel(const Tc4Args a, const __grid_constant__ CUtensorMap tmap) {
  static_assert(KP == 32 || KP == 64, "score columns");
#ifdef LKM_T4_TRACE
  const long long k_t0 = clock64();
#define T4K(ev) do { if (threadIdx.x == 0 && a.prof) { if (blockIdx.x == 0) a.prof[ev] = clock64() - k_t0; \
    if ((ev) == 0 || (ev) == 4 || (ev) == 6 || (ev) == 8) { long long gt__; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt__)); \
      a.prof[64 + 40 * 16 + blockIdx.x * 4 + ((ev) == 0 ? 0 : (ev) == 4 ? 1 : (ev) == 6 ? 2 : 3)] = gt__; } } } while (0)
#else
#define T4K(ev)
#endif
  pdl_launch_dependents();
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (tc::smem_u32(smem_unaligned) & 1023u)) & 1023u);
  const int t = threadIdx.x, lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);   // warp-uniform by construction
  const int k = a.k;
  const int R = a.stages;
  unsigned char* raw = smem;                                   // [R][16 KB]
  unsigned char* b_hi = raw + (size_t)R * 16384;
  unsigned char* b_lo = b_hi + (PASSES == 1 ? 0 : KP * 128);   // the one-product variant has no B_lo tile
  unsigned char* b_bias = b_lo + KP * 128;
  float* rawc = reinterpret_cast<float*>(b_bias + KP * 128);
  float* acc = rawc + (size_t)k * kTcCStride;
  uint32_t* cnt = reinterpret_cast<uint32_t*>(acc + (size_t)kT4Subsets * (k + 1) * 32);
  float* nn_s = reinterpret_cast<float*>(cnt + ((k + 3) & ~3));   // [KP] (the f64-sized slot of the layout is kept)
  uint32_t* labs2_all = reinterpret_cast<uint32_t*>(nn_s + 2 * KP);  // [slot][128], index (row & 3) * 32 + (row >> 2)
  uint32_t* flags_all = labs2_all + 2 * kTcRows;                // [slot][128]
  uint32_t* wuni_all = flags_all + 2 * kTcRows;                 // [slot][4]
  __shared__ __align__(8) uint64_t bar_full[kT4MaxStages], bar_free[kT4MaxStages], bar_alo[2], bar_mma[2], bar_tfree[2], bar_lab[2], bar_lfree[2];
  __shared__ uint32_t tmem_slot;
  __shared__ uint32_t n_flag[2];
  __shared__ uint32_t cmax_bits;

  const unsigned long long n_tiles = (a.n + kTcRows - 1) / kTcRows;
  const unsigned long long my_tiles = blockIdx.x < n_tiles ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto tile_of = [&](unsigned long long j) { return blockIdx.x + j * gridDim.x; };

  if (warp == 0) tc::tmem_alloc(&tmem_slot, kT4TmemCols);
  if (t == 32) {
    for (int s = 0; s < kT4MaxStages; ++s) {
      tc::mbar_init(&bar_full[s], 1);
      tc::mbar_init(&bar_free[s], 2);
    }
    for (int s = 0; s < 2; ++s) {
      tc::mbar_init(&bar_alo[s], 4);
      tc::mbar_init(&bar_mma[s], 1);
      tc::mbar_init(&bar_tfree[s], 4);
      tc::mbar_init(&bar_lab[s], 4);
      tc::mbar_init(&bar_lfree[s], 4);
      n_flag[s] = 0;
    }
    tc::fence_mbar_init();
    cmax_bits = 0;
  }
  __syncthreads();
  T4K(0);
  // The observations do not depend on the previous kernel: fill the ring before waiting for it.
  // Only three tiles up front: a full ring (16 KB x stages x 148 CTAs) ahead of the centroid loads would delay
  // those by the whole burst; the P role requests the rest as soon as it starts.
  unsigned long long produced = 0;
  if (warp == 16 && lane == 0) {
    for (; produced < my_tiles && produced < 3ull; ++produced) {
      tc::mbar_expect_tx(&bar_full[produced], 16384u);
      tc::tma_load_2d(raw + (size_t)produced * 16384, &tmap, 0, (int)(tile_of(produced) * kTcRows), &bar_full[produced]);
    }
    for (unsigned long long j = produced; j < my_tiles && j < produced + kT4Prefetch; ++j) {
      const unsigned long long r0 = tile_of(j) * kTcRows;
      tc::l2_prefetch(a.X + r0 * 32ull, (uint32_t)min((unsigned long long)kTcRows, a.n - r0) * 128u);
    }
  }
  // everything that does not depend on the previous kernel's centroids happens before the dependency wait
  for (int e = t; e < kT4Subsets * (k + 1) * 32; e += kT4Threads) acc[e] = 0.f;
  for (int e = t; e < k; e += kT4Threads) cnt[e] = 0u;
  for (int e = t; e < KP * 8; e += kT4Threads) reinterpret_cast<float4*>(b_bias)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  pdl_wait();
  T4K(1);
  const bool done = a.state != nullptr && a.state->done;
  const float xmax2 = __uint_as_float(__ldg(a.xmax2_bits));   // issued together with the centroid loads

  // ---- centroids: TF32 hi/lo operand tiles, padded raw copy, squared norms (8 threads per row, f32) ----
  for (int item = t; item < KP * 8; item += kT4Threads) {
    const int r = item >> 3, c = item & 7;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < k) v = __ldg(reinterpret_cast<const float4*>(a.C + (size_t)r * 32) + c);
    float4 h, l;
    tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
    tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
    const uint32_t o = tc::sw128_offset(r, c);
    *reinterpret_cast<float4*>(b_hi + o) = h;
    if constexpr (PASSES != 1) *reinterpret_cast<float4*>(b_lo + o) = l;
    if (r < k) *reinterpret_cast<float4*>(rawc + r * kTcCStride + 4 * c) = v;
    float nn = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, v.w * v.w)));
    nn += __shfl_xor_sync(0xffffffffu, nn, 1);
    nn += __shfl_xor_sync(0xffffffffu, nn, 2);
    nn += __shfl_xor_sync(0xffffffffu, nn, 4);
    if (c == 0) {
      nn_s[r] = nn;   // |nn - |c|^2| <= (d + 1) 2^-24 |c|^2 whatever the summation order
      if (r < k) atomicMax(&cmax_bits, __float_as_uint(nn));  // a NaN norm wins and voids every certificate
    }
  }
  T4K(2);
  __syncthreads();
  T4K(3);
  // ---- window, bias, margin (identical in every thread; see lkm_tc3.cuh) ----
  const float cmax2 = __uint_as_float(cmax_bits) * 1.00001f;
  const float cmax = sqrtf(cmax2) * 1.000001f;
  const float xmax = sqrtf(xmax2) * 1.000001f;
  const float Rb = 1.01f * (xmax * cmax + 0.5f * cmax2) + 1e-30f;   // |x.c - |c|^2/2| <= Rb
  const bool params_ok = Rb < 1e30f;  // false for NaN / Inf: every row then takes the exact scan
  uint32_t e_lo = ((__float_as_uint(Rb * (1.0f / 127.0f)) >> 23) + 1u + 7u) & ~7u;
  e_lo = min(max(e_lo, 8u), 240u);
  const float two_lo = __uint_as_float(e_lo << 23);
  const float K0 = Rb + two_lo;                // v = x.c - |c|^2/2 + K0 in [2^lo, 255 * 2^lo)
  const uint32_t top_bits = (e_lo << 23) & 0xFC000000u;
  const float eta = (3.f * 32.f + 12.f) * 2.3841858e-7f;  // (3d + 12) * 2^-22
  const float gamma = (32.f + 2.f) * 5.9604645e-8f;       // (d + 2) * 2^-24
  const float m_w = 1.1f * (2.3841858e-7f * cmax2 + xmax * (4.f * eta + 4.7683716e-7f) * cmax +
                            2.f * gamma * (xmax + cmax) * (xmax + cmax));
  // v is half a squared distance.  On each of the two candidates: the bias K-step rounds at most 3 ulp(v), the
  // f32 bias K0 - nn/2 is off by at most 2^-24 K0 + (d + 1) 2^-25 |c|^2 (rounding of the norm and of the subtraction)
  const float m_v = 0.5f * m_w + 7.f * 1.1920929e-7f * (K0 + Rb) + 33.f * 5.9604645e-8f * cmax2 +
                    (PASSES == 2 ? 1.001f * 1.953125e-3f * xmax * cmax : 0.f) +   // + 2 * 2^-10 |x||c| without the x_lo pass (|x_lo| < 2^-10 |x| element-wise)
                    (PASSES == 1 ? 1.001f * 3.90625e-3f * xmax * cmax : 0.f);     // + 2 (2^-9 + 2^-20) |x||c| with x_hi.c_hi alone
  if (t < KP) {
    // bias row t: K0 - |c|^2/2 as three TF32 pieces that add up to the f32 value exactly, K-step 0 of the tile
    const float b = (t < k) ? (K0 - 0.5f * nn_s[t]) : two_lo;
    const float h = tc::trunc_tf32(b);
    const float l = b - h;
    const float l1 = tc::trunc_tf32(l);
    const float l2 = l - l1;
    float* p = reinterpret_cast<float*>(b_bias + tc::sw128_offset(t, 0));
    p[0] = h; p[1] = l1; p[2] = l2;
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  T4K(4);
  if (warp < 4) {
    // the constant A operand of the bias K-step: every lane holds [1, 1, 1, 0, 0, 0, 0, 0]
    const uint32_t one = __float_as_uint(1.f);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(tmem + ((uint32_t)(warp * 32) << 16) + kT4ColOnes), "r"(one), "r"(one), "r"(one), "r"(0u), "r"(0u), "r"(0u),
                   "r"(0u), "r"(0u)
                 : "memory");
    tc::tmem_st_wait();
    tc::tc_fence_before();
  }
  double inert = 0.0;            // U role
#ifdef LKM_T4_TRACE
  const long long trace_t0 = clock64();
#endif
#ifdef LKM_T4_PROF
  long long prof_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  const long long prof_t0 = clock64();
#endif
  unsigned int my_fallbacks = 0;

  if (done) {
    // nothing to do, but the requested tiles must land before the CTA may exit
    if (warp == 16 && lane == 0)
      for (unsigned long long j = 0; j < produced; ++j) tc::mbar_wait_or_trap(&bar_full[j], 0);
  } else if (warp == 16) {
    // =============================== P: refill the ring ===============================
    if (lane == 0) {
      for (unsigned long long j = produced; j < my_tiles; ++j) {
        const int s = (int)(j % (unsigned long long)R);
        const uint32_t use = (uint32_t)(j / (unsigned long long)R);
        if (use > 0) T4WAIT(0, &bar_free[s], (use - 1u) & 1u);
        tc::mbar_expect_tx(&bar_full[s], 16384u);
        tc::tma_load_2d(raw + (size_t)s * 16384, &tmap, 0, (int)(tile_of(j) * kTcRows), &bar_full[s]);
        T4TRACE(j, 0);
        if (j + kT4Prefetch < my_tiles) {
          const unsigned long long r0 = tile_of(j + kT4Prefetch) * kTcRows;
          tc::l2_prefetch(a.X + r0 * 32ull, (uint32_t)min((unsigned long long)kTcRows, a.n - r0) * 128u);
        }
      }
    }
  } else if (warp == 17) {
    // =============================== M: issue the MMAs ===============================
    // The whole warp runs the loop (uniform control flow and uniform-register operands); one elected lane
    // issues each instruction.
    {
      constexpr uint32_t idesc = tc::idesc_tf32(128, (uint32_t)KP);
      const uint64_t desc_bhi = tc::smem_desc_sw128(tc::smem_u32(b_hi));
      const uint64_t desc_blo = tc::smem_desc_sw128(tc::smem_u32(b_lo));
      const uint64_t desc_bias = tc::smem_desc_sw128(tc::smem_u32(b_bias));
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      // x_lo.c_hi, x_hi.c_lo, x_hi.c_hi, bias; A K-step ks = 8 TMEM columns, B K-step = 32 bytes = 2 descriptor units
      const uint64_t desc_raw = tc::smem_desc_sw128(tc::smem_u32(raw));
      auto issue = [&](const uint32_t p2, const int stage) {
        const uint32_t dcol = tmem_u + p2 * 64u;
        if constexpr (PASSES == 3) {
          const uint32_t a_hi = tmem_u + kT4ColAhi + p2 * 32u, a_lo = tmem_u + kT4ColAlo + p2 * 32u;
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc::mma_tf32_ts_elect(dcol, a_lo + 8u * ks, desc_bhi + 2u * ks, idesc, ks ? 1u : 0u);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc::mma_tf32_ts_elect(dcol, a_hi + 8u * ks, desc_blo + 2u * ks, idesc, 1u);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc::mma_tf32_ts_elect(dcol, a_hi + 8u * ks, desc_bhi + 2u * ks, idesc, 1u);
        } else {
          const uint64_t da = desc_raw + (uint64_t)stage * 1024u;   // 16 KB per stage, 16-byte units
          if constexpr (PASSES == 2) {
#pragma unroll
            for (int ks = 0; ks < 4; ++ks) tc::mma_tf32_elect(dcol, da + 2u * ks, desc_blo + 2u * ks, idesc, ks ? 1u : 0u);
          }
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc::mma_tf32_elect(dcol, da + 2u * ks, desc_bhi + 2u * ks, idesc, (PASSES == 2 || ks) ? 1u : 0u);
        }
        tc::mma_tf32_ts_elect(dcol, tmem_u + kT4ColOnes, desc_bias, idesc, 1u);
      };
      int s = 0;
      uint32_t full_ph = 0;
      for (unsigned long long j = 0; j < my_tiles; ++j) {
        const uint32_t p2 = (uint32_t)j & 1u, u = (uint32_t)(j >> 1);
        if constexpr (PASSES == 3) tc::nb_sync(kNbAlo + (int)p2, 160);
        else tc::mbar_wait_or_trap(&bar_full[s], full_ph);      // the raw tile has landed
        T4TRACE(j, 12);
        if (u > 0) tc::nb_sync(kNbTfree + (int)p2, 160);   // accumulator drained by E
        T4TRACE(j, 13);
        tc::tc_fence_after();
        T4TRACE(j, 3);
        if (p2) issue(1u, s); else issue(0u, s);
        tc::mma_commit_elect(&bar_mma[p2]);
        T4TRACE(j, 4);
        if (++s == R) { s = 0; full_ph ^= 1u; }
      }
    }
  } else if (warp < 4) {
    // =============================== S: the A operands of tile j into TMEM ===============================
    if constexpr (PASSES == 3) {
    int s = 0;
    uint32_t full_ph = 0;
    const uint32_t row_off = (uint32_t)(t >> 3) * 1024u + (uint32_t)(t & 7) * 128u;   // row t of a SW128 tile
    const uint32_t row_x = (uint32_t)(t & 7);
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t tf32_mask = 0xFFFFE000u, sign_bit = 0x80000000u;
    asm volatile("" : "+r"(tf32_mask), "+r"(sign_bit));   // keep both in registers: one LOP3 per element
    for (unsigned long long j = 0; j < my_tiles; ++j) {
      const uint32_t p2 = (uint32_t)j & 1u, u = (uint32_t)(j >> 1);
      const unsigned char* rw = raw + (size_t)s * 16384;
      T4WAIT(0, &bar_full[s], full_ph);
      if (warp == 0) T4TRACE(j, 1);
      uint32_t xr[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const uint4 v = *reinterpret_cast<const uint4*>(rw + row_off + (((uint32_t)c ^ row_x) << 4));
        xr[4 * c] = v.x; xr[4 * c + 1] = v.y; xr[4 * c + 2] = v.z; xr[4 * c + 3] = v.w;
      }
      if (u > 0) T4WAIT(1, &bar_mma[p2], (u - 1u) & 1u);   // the MMAs of tile j-2 have read these TMEM columns
      tc::tc_fence_after();
      tc::tmem_st32(lane_base + kT4ColAhi + p2 * 32u, xr);
#pragma unroll
      for (int q = 0; q < 32; q += 2) {
        // x_lo = x + (-trunc_tf32(x)), two features per FADD2
        const uint32_t n0 = tc::neg_trunc_tf32_bits(xr[q], tf32_mask, sign_bit);
        const uint32_t n1 = tc::neg_trunc_tf32_bits(xr[q + 1], tf32_mask, sign_bit);
        tc::f2_add_parts(xr[q], xr[q + 1], n0, n1, xr[q], xr[q + 1]);
      }
      tc::tmem_st32(lane_base + kT4ColAlo + p2 * 32u, xr);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      tc::nb_arrive(kNbAlo + (int)p2, 160);
      if (warp == 0) T4TRACE(j, 2);
      if (++s == R) { s = 0; full_ph ^= 1u; }
    }
    }
  } else if (warp < 12) {
    // =============================== E: score the rows of tiles j = g (mod 2) ===============================
    const int g = (warp - 4) >> 2;            // tile parity = TMEM accumulator = label slot
    const int wq = warp & 3;                  // TMEM lane quadrant
    const int row = (t - 128) & 127;          // row of the tile owned by this thread
    uint32_t* labs2 = labs2_all + g * kTcRows;
    uint32_t* flags = flags_all + g * kTcRows;
    uint32_t* wuni = wuni_all + g * 4;
    const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)g * 64u;
    for (unsigned long long j = g; j < my_tiles; j += 2) {
      const uint32_t u = (uint32_t)(j >> 1);
      const unsigned long long row0 = tile_of(j) * kTcRows;
      const int rows = (int)min((unsigned long long)kTcRows, a.n - row0);
      T4WAIT(0, &bar_mma[g], u & 1u);
      if (wq == 0) T4TRACE(j, 5);
      T4MARK(te_a);
      tc::tc_fence_after();
      uint32_t q[KP];
      {
        uint32_t r0[32];
        tc::tmem_ld32(taddr, r0);
        if constexpr (KP == 64) {
          uint32_t r1[32];
          tc::tmem_ld32(taddr + 32u, r1);
          tc::tmem_ld_wait();
#pragma unroll
          for (int c = 0; c < 32; ++c) q[32 + c] = r1[c] * a.pack_mult + (uint32_t)(32 + c);
        } else {
          tc::tmem_ld_wait();
        }
#pragma unroll
        for (int c = 0; c < 32; ++c) q[c] = r0[c] * a.pack_mult + (uint32_t)c;
      }
      T4MARK(te0);
      T4ADD(3, te_a, te0);
      tc::tc_fence_before();
      tc::nb_arrive(kNbTfree + g, 160);        // the accumulator may be overwritten by tile j + 2
      if (wq == 0) T4TRACE(j, 6);
      // integer min / max / add share the 16-lane ALU pipe (two cycles per warp instruction), the busiest unit of this
      // loop: the pack above is an IMAD (FMA pipe) and the runner-up pass is chains of min(m, best - 1 - q), one fused
      // VIADDMNMX per score (the best itself wraps to 0xFFFFFFFF, the others become gap - 1)
      uint32_t m0 = q[0], m1 = q[1], m2 = q[2], m3 = q[3];
#pragma unroll
      for (int c = 4; c + 7 < KP; c += 8) {
        m0 = __vimax3_u32(m0, q[c], q[c + 1]);
        m1 = __vimax3_u32(m1, q[c + 2], q[c + 3]);
        m2 = __vimax3_u32(m2, q[c + 4], q[c + 5]);
        m3 = __vimax3_u32(m3, q[c + 6], q[c + 7]);
      }
      m0 = __vimax3_u32(m0, q[KP - 4], q[KP - 3]);
      m1 = __vimax3_u32(m1, q[KP - 2], q[KP - 1]);
      const uint32_t bestq = max(__vimax3_u32(m0, m1, m2), m3);
      const uint32_t bm1 = bestq - 1u;
      m0 = 0xFFFFFFFFu; m1 = 0xFFFFFFFFu; m2 = 0xFFFFFFFFu; m3 = 0xFFFFFFFFu;
#pragma unroll
      for (int c = 0; c < KP; c += 4) {
        m0 = min(m0, bm1 - q[c]);
        m1 = min(m1, bm1 - q[c + 1]);
        m2 = min(m2, bm1 - q[c + 2]);
        m3 = min(m3, bm1 - q[c + 3]);
      }
      const uint32_t tmin = min(min(m0, m1), min(m2, m3));
      const uint32_t secondq = bm1 - tmin;
      const float v_best = __uint_as_float((bestq >> 6) | top_bits);
      const float v_second = __uint_as_float((secondq >> 6) | top_bits);
      const uint32_t bidx = bestq & 63u;
      const bool certified = params_ok && ((v_best - v_second) > m_v) && bidx < (uint32_t)k;
      uint32_t lab = certified ? bidx : 0xFFFFFFFFu;
      if (row >= rows) lab = 0xFFFFFFFEu;
      const uint32_t lab0 = __shfl_sync(0xffffffffu, lab, 0);
      const bool uni = __all_sync(0xffffffffu, lab == lab0) && lab0 < 0xFFFFFFFEu;
      T4MARK(te1);
      T4ADD(2, te0, te1);
      if (wq == 0) T4TRACE(j, 7);
      if (u > 0) tc::nb_sync(kNbLfree + g, 192);   // U group g is done with this label slot
      T4MARK(te2);
      if (uni) {
        if (lane == 0) atomicAdd(&cnt[lab0], 32u);
      } else if (lab < 0xFFFFFFFEu) {
        atomicAdd(&cnt[lab], 1u);
      }
      if (lane == 0) wuni[wq] = uni ? lab0 : 0xFFFFFFFFu;
      labs2[(row & 3) * 32 + (row >> 2)] = lab;
      if (lab == 0xFFFFFFFFu) {
        const uint32_t p = atomicAdd(&n_flag[g], 1u);
        flags[p] = (uint32_t)row;
      }
      tc::nb_arrive(kNbLab + g, 192);
      if (wq == 0) T4TRACE(j, 8);
      T4MARK(te3);
      T4ADD(4, te2, te3);
    }
  } else {
    // =============================== U: exact re-scans, sums, inertia ===============================
    // Two independent U groups of two warps: group g takes the tiles j = g (mod 2) scored by E group g and owns
    // its own accumulator subsets, so the even and the odd tiles form two pipelines that only share P, S and M.
    const int g = (warp - 12) >> 1;           // tile parity
    const int wl = (warp - 12) & 1;           // warp within the group
    const int tg = wl * 32 + lane;            // 0..63
    // thread = (row subset sub4 of 4, feature pair fp): features 2 fp, 2 fp + 1 of rows sub4, sub4 + 4, ...
    const int sub4 = wl * 2 + (lane >> 4), fp = lane & 15;
    // byte offsets of (row, feature 2 fp) for row & 7 = sub4 (even i) and sub4 + 4 (odd i); + 1024 (i >> 1)
    const uint32_t off_e = (uint32_t)sub4 * 128u + ((((uint32_t)fp >> 1) ^ (uint32_t)sub4) << 4) + ((uint32_t)fp & 1u) * 8u;
    const uint32_t off_o = (uint32_t)(sub4 + 4) * 128u + ((((uint32_t)fp >> 1) ^ (uint32_t)(sub4 + 4)) << 4) + ((uint32_t)fp & 1u) * 8u;
    float* my_acc = acc + (size_t)(g * 4 + sub4) * (k + 1) * 32 + 2 * fp;
    uint32_t* labs2 = labs2_all + g * kTcRows;
    const uint32_t* flags = flags_all + g * kTcRows;
    const int ubar = g ? 10 : 1;              // group-internal named barrier
    float inert_f = 0.f, inert_g = 0.f, inert_h = 0.f;   // per-tile partials -> 8-tile -> 64-tile partials -> f64 at the end
    int s = g;
    for (unsigned long long j = g; j < my_tiles; j += 2) {
      const unsigned long long row0 = tile_of(j) * kTcRows;
      const int rows = (int)min((unsigned long long)kTcRows, a.n - row0);
      const unsigned char* rw = raw + (size_t)s * 16384;
      tc::nb_sync(kNbLab + g, 192);
      if (wl == 0) T4TRACE(j, 9);
      T4MARK(tu_a);
      const uint32_t nf = n_flag[g];
      const uint4 wu = *reinterpret_cast<const uint4*>(wuni_all + g * 4);
      if (nf) {
        // uncertified rows: exact scan, one warp per row, lanes split the centroids
        for (uint32_t f = wl; f < nf; f += 2) {
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
        my_fallbacks += (tg == 0) ? nf : 0u;
        tc::nb_sync(ubar, 64);
        if (tg == 0) n_flag[g] = 0;
      }
      T4MARK(tu0);
      T4ADD(4, tu_a, tu0);
      if (a.labels_out != nullptr) {
        // the labels this tile is accumulated under (certified by E or re-scanned above), row order
        for (int r = tg; r < rows; r += 64) a.labels_out[row0 + r] = labs2[(r & 3) * 32 + (r >> 2)];
      }
      unsigned long long e2 = 0ull;
      if (nf == 0 && wu.x != 0xFFFFFFFFu && wu.x == wu.y && wu.x == wu.z && wu.x == wu.w) {
        // the whole tile goes to one cluster: register sums in row order, two features per instruction
        const uint32_t l1 = wu.x;
        const float2 cf = *reinterpret_cast<const float2*>(rawc + l1 * kTcCStride + 2 * fp);
        const unsigned long long ncf = tc::f2_pack(-cf.x, -cf.y);
        unsigned long long s0 = 0ull, s1 = 0ull, e3 = 0ull;
#pragma unroll
        for (int ii = 0; ii < 16; ++ii) {
          const unsigned long long x0 = *reinterpret_cast<const unsigned long long*>(rw + 1024u * ii + off_e);
          const unsigned long long x1 = *reinterpret_cast<const unsigned long long*>(rw + 1024u * ii + off_o);
          const unsigned long long d0 = tc::f2_add(x0, ncf), d1 = tc::f2_add(x1, ncf);
          s0 = tc::f2_add(s0, x0); s1 = tc::f2_add(s1, x1);
          e2 = tc::f2_fma(d0, d0, e2); e3 = tc::f2_fma(d1, d1, e3);
        }
        s0 = tc::f2_add(s0, s1);
        e2 = tc::f2_add(e2, e3);
        float2* ap = reinterpret_cast<float2*>(my_acc + l1 * 32);
        float2 av = *ap;
        av.x += tc::f2_lo(s0); av.y += tc::f2_hi(s0);
        *ap = av;
      } else {
        // mixed labels (or a tail tile), four rows at a time.  Four different labels touch four different
        // accumulators, so their read-modify-writes are independent and overlap; batches with repeated labels
        // (sorted data at a cluster boundary) sum runs of equal labels in registers and flush on a change.
#pragma unroll 2
        for (int qq = 0; qq < 8; ++qq) {
          const uint4 lv = *reinterpret_cast<const uint4*>(labs2 + sub4 * 32 + 4 * qq);
          const uint32_t l4[4] = {lv.x, lv.y, lv.z, lv.w};
          const unsigned char* rq = rw + 2048u * qq;       // rows sub4 + 16 qq + {0, 4, 8, 12}
          const bool full4 = sub4 + 16 * qq + 12 < rows;
          const bool distinct = l4[0] != l4[1] && l4[0] != l4[2] && l4[0] != l4[3] && l4[1] != l4[2] && l4[1] != l4[3] && l4[2] != l4[3];
          if (full4 && distinct) {
            unsigned long long xv[4];
            float2 av[4], cv[4];
            xv[0] = *reinterpret_cast<const unsigned long long*>(rq + off_e);
            xv[1] = *reinterpret_cast<const unsigned long long*>(rq + off_o);
            xv[2] = *reinterpret_cast<const unsigned long long*>(rq + 1024u + off_e);
            xv[3] = *reinterpret_cast<const unsigned long long*>(rq + 1024u + off_o);
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              av[uu] = *reinterpret_cast<const float2*>(my_acc + l4[uu] * 32);
              cv[uu] = *reinterpret_cast<const float2*>(rawc + l4[uu] * kTcCStride + 2 * fp);
            }
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              av[uu].x += tc::f2_lo(xv[uu]); av[uu].y += tc::f2_hi(xv[uu]);
              *reinterpret_cast<float2*>(my_acc + l4[uu] * 32) = av[uu];
              const unsigned long long dv = tc::f2_add(xv[uu], tc::f2_pack(-cv[uu].x, -cv[uu].y));
              e2 = tc::f2_fma(dv, dv, e2);
            }
          } else {
            uint32_t cur_lab = (uint32_t)k;   // dummy accumulator row
            unsigned long long cur = 0ull, ncf = 0ull;
#pragma unroll
            for (int uu = 0; uu < 4; ++uu) {
              if (sub4 + 16 * qq + 4 * uu < rows) {
                const unsigned long long xv = *reinterpret_cast<const unsigned long long*>(rq + 1024u * (uu >> 1) + ((uu & 1) ? off_o : off_e));
                if (l4[uu] != cur_lab) {
                  float2* ap = reinterpret_cast<float2*>(my_acc + cur_lab * 32);
                  float2 a2 = *ap;
                  a2.x += tc::f2_lo(cur); a2.y += tc::f2_hi(cur);
                  *ap = a2;
                  cur_lab = l4[uu];
                  cur = 0ull;
                  const float2 cf = *reinterpret_cast<const float2*>(rawc + cur_lab * kTcCStride + 2 * fp);
                  ncf = tc::f2_pack(-cf.x, -cf.y);
                }
                const unsigned long long dv = tc::f2_add(xv, ncf);
                cur = tc::f2_add(cur, xv);
                e2 = tc::f2_fma(dv, dv, e2);
              }
            }
            float2* ap = reinterpret_cast<float2*>(my_acc + cur_lab * 32);
            float2 a2 = *ap;
            a2.x += tc::f2_lo(cur); a2.y += tc::f2_hi(cur);
            *ap = a2;
          }
        }
      }
      T4MARK(tu1);
      T4ADD(2, tu0, tu1);
      if (wl == 0) T4TRACE(j, 10);
      inert_f += tc::f2_lo(e2) + tc::f2_hi(e2);
      if ((j & 14) == 14) { inert_g += inert_f; inert_f = 0.f; }
      if ((j & 126) == 126) { inert_h += inert_g; inert_g = 0.f; }
      tc::warp_arrive(&bar_free[s], lane);   // the raw stage may be refilled
      tc::nb_arrive(kNbLfree + g, 192);      // the label slot may be rewritten
      if (wl == 0) T4TRACE(j, 11);
      T4MARK(tu2);
      T4ADD(3, tu1, tu2);
      s += 2;
      if (s >= R) s -= R;
    }
    inert = (double)inert_h + (double)inert_g + (double)inert_f;
  }
#ifdef LKM_T4_PROF
  prof_acc[7] = clock64() - prof_t0;
  if (blockIdx.x == 0 && lane == 0 && a.prof) {
    const int role = warp == 16 ? 0 : warp == 17 ? 1 : warp == 0 ? 2 : warp == 4 ? 3 : warp == 8 ? 4 : warp == 12 ? 5 : -1;
    if (role >= 0)
      for (int q = 0; q < 8; ++q) a.prof[role * 8 + q] = prof_acc[q];
  }
#endif
  T4K(5);
  __syncthreads();
  T4K(6);
  if (!done) {
    // ---- per-CTA partials (same layout as the other fused kernels) ----
    const int kd = k * 32, sstride = (k + 1) * 32;
    float* ps = a.part_sums + (size_t)blockIdx.x * kd;
    for (int e = t; e < kd; e += kT4Threads) {
      float s2 = acc[e];
#pragma unroll
      for (int qq = 1; qq < kT4Subsets; ++qq) s2 += acc[(size_t)qq * sstride + e];
      ps[e] = s2;
    }
    uint32_t* pc = a.part_counts + (size_t)blockIdx.x * k;
    for (int e = t; e < k; e += kT4Threads) pc[e] = cnt[e];
    // inertia: the 128 U threads' partials, fixed order (warp butterfly, then the four warp sums)
    double* red = reinterpret_cast<double*>(raw);
    if (t >= 384 && t < 512) {
      float v = (float)inert;
      for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
      if (lane == 0) red[(t - 384) >> 5] = (double)v;
    }
    __syncthreads();
    if (t == 0) a.part_inertia[blockIdx.x] = (red[0] + red[1]) + (red[2] + red[3]);
    if (a.state && (t == 384 || t == 448) && my_fallbacks) atomicAdd(&a.state->fallback_rows, (unsigned long long)my_fallbacks);
  }
  T4K(7);
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc(tmem, kT4TmemCols);
  T4K(8);
  if (a.fin_sync != nullptr && !done) {
    // ---- in-kernel finalize (KM/algorithm.rs:785-810, 314-331; same arithmetic as finalize_kernel<UPDATE>) ----
    // All CTAs of the grid are resident (grid <= SM count, one CTA per SM), so a counter rendezvous is safe.  Then CTA b
    // combines elements [b epb, (b + 1) epb) of the k*d sums over the per-CTA partials: lane = source b, b + 32, ...
    // (from L2: the partials were written by other SMs), warp butterfly, one warp per element: a fixed order.
    __shared__ bool fin_last;
    double* redf = reinterpret_cast<double*>(raw);
    const unsigned int G = gridDim.x;
    __threadfence();
    __syncthreads();
    if (t == 0) {
      atomicAdd(&a.fin_sync[0], 1u);
      unsigned int spins = 0;
      while (*((volatile unsigned int*)&a.fin_sync[0]) < G && ++spins < (1u << 25)) {}
      if (*((volatile unsigned int*)&a.fin_sync[0]) < G) atomicOr(&a.state->fallback_rows, 1ull << 41);   // timeout marker, reported by the host
      __threadfence();
    }
    __syncthreads();
    const int kd = k * 32;
    const int epb = (kd + (int)G - 1) / (int)G;
    double sq_w = 0.0;
    for (int el = warp; el < epb; el += kT4Threads / 32) {
      const int e = (int)blockIdx.x * epb + el;
      if (e < kd) {
        const int c = e >> 5;
        double sv = 0.0;
        unsigned int cn = 0u;
        for (unsigned int b = lane; b < G; b += 32u) {
          sv += (double)__ldcg(a.part_sums + (size_t)b * kd + e);
          cn += __ldcg(a.part_counts + (size_t)b * k + c);
        }
        for (int o = 16; o > 0; o >>= 1) {
          sv += __shfl_xor_sync(0xffffffffu, sv, o);
          cn += __shfl_xor_sync(0xffffffffu, cn, o);
        }
        if (lane == 0) {
          const float oldv = a.C[e];
          const float newv = (float)((sv + (double)oldv) / ((double)cn + 1.0));
          a.C_new[e] = newv;
          const double df = (double)(oldv - newv);
          sq_w += df * df;
          if ((e & 31) == 0) a.counts_out[c] = (double)cn;
        }
      }
    }
    if (lane == 0) redf[warp] = sq_w;
    __syncthreads();
    if (t == 0) {
      double v = 0.0;
      for (int w = 0; w < kT4Threads / 32; ++w) v += redf[w];
      a.shift_part[blockIdx.x] = v;
      __threadfence();
      fin_last = atomicAdd(&a.fin_sync[1], 1u) == G - 1u;
    }
    __syncthreads();
    if (fin_last) {
      __threadfence();
      if (warp < 2) {
        const double* src = warp == 0 ? a.shift_part : a.part_inertia;
        double v = 0.0;
        for (unsigned int b = lane; b < G; b += 32u) v += __ldcg(src + b);
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) redf[32 + warp] = v;
      }
      __syncthreads();
      if (t == 0) {
        const float shift = (float)sqrt(redf[32]);   // L2Dist::distance: sqrt in f64, cast to F
        a.state->shift = (double)shift;
        a.state->inertia = redf[33];
        const unsigned long long itn = a.state->n_iter + 1;
        a.state->n_iter = itn;
        if (shift < (float)a.tol || itn >= a.max_iter) a.state->done = 1;
        a.fin_sync[0] = 0u;
        a.fin_sync[1] = 0u;
      }
    }
  }
}

}  // namespace lkm

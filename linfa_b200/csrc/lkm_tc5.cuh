// Tensor-core ASSIGNMENT for wide rows (f32, d = 64 or 128, k <= 256) on CTA PAIRS — cfg3 shape.
//
// update_memberships_and_dists / closest_centroid (KM/algorithm.rs:866-881, 923-946) as one 256 x 256 x d
// 3xTF32 contraction per pair of 128-row tiles: the two CTAs of a cluster run tcgen05.mma.cta_group::2, each
// supplying its own 128 rows of X (A operand, in its TMEM) and HALF of the centroids (B operand, 128 rows of
// hi / lo / bias tiles resident in its shared memory for the whole launch), so the 2 x k x d x 4 bytes of
// centroid operands fit on chip once and no operand is re-staged per tile.  Per CTA, 10 warps:
//
//   P  warp 9     one thread streams the tile through a ring of four 32-feature slabs (TMA, SWIZZLE_128B)
//   S  warps 0-3  thread = row: slab -> TMEM, once as raw bits (= x_hi: kind::tf32 ignores the 13 low mantissa
//                 bits) and once as x_lo = x - trunc(x)
//   M  warp 8     (leader CTA only) 3 d/8 + 1 MMAs of M = 256, N = 256 into the 256 accumulator columns of both
//                 CTAs:  x_lo.c_hi + x_hi.c_lo + x_hi.c_hi + ones.(K0 - |c|^2/2)  ->  v_c = x.c - |c|^2/2 + K0;
//                 one multicast commit signals both CTAs
//   E  warps 4-7  thread = row: packed integer top-2 of the 256 scores (q = bits(v) << 8 | c, every v inside
//                 one 4x window so that the shifted bits stay monotone), certificate (v-form), label;
//                 uncertified rows are re-scanned in the reference's exact arithmetic, one warp per row
//
// TMEM (512 columns): accumulator 256 | A_hi d | A_lo d.  The roles of the two CTAs meet at three barriers of
// the leader (A operands ready: 8 S warps; accumulator drained: 8 E warps) and at the multicast MMA commit.
// Labels only: the update pass (lkm_upd.cuh) reads X again and produces sums, counts and inertia.
#pragma once
#include "lkm_tc4.cuh"

namespace lkm {

constexpr int kT5Threads = 10 * 32;
__host__ __device__ constexpr int t5_fuse_threads(int d) { return (10 + d / 32) * 32; }   // screening kernel with the fused update: + one U warp per slab
constexpr int kT5Stages = 4;
constexpr uint32_t kT5ColAcc = 0, kT5ColA = 256;
constexpr unsigned long long kT5Prefetch = 3;   // L2 prefetch distance of the P roles, in tiles
constexpr int kT5K = 256;   // score columns = centroid rows of one MMA (128 per CTA)

struct Tc5Args {
  const float* X;
  unsigned long long n;
  int k;                            // all centroids
  int chunk_idx, n_chunks;          // this launch scores centroids [chunk_n chunk_idx, min(k, chunk_n (chunk_idx + 1)))
  int chunk_n;                      // score columns per launch: 256 (64 for the d = 32, k <= 64 instantiation)
  uint4* merge;                     // n_chunks > 1: per row {best, runner-up, chunk of best, -} carried between the launches
  const float* C;
  const unsigned int* xmax2_bits;   // max_i |x_i|^2 of the resident matrix (row_norm2_max_wide_kernel)
  uint32_t* labels;
  LloydState* state;
  long long* prof;   // LKM_T5_TRACE builds only
  // fused update of the screening kernel (FUSE = true; with several chunks: the launch of the last one): per-CTA partials for finalize_kernel and the list of the tiles
  // whose rows did not all certify to one cluster (those go through update_rows_kernel)
  float* part_sums;         // [gridDim.x][k * d], zeroed by the kernel
  uint32_t* part_counts;    // [gridDim.x][k]
  unsigned char* part_touched;   // [gridDim.x][k] or nullptr: clusters this CTA flushed; with it (k <= 1024) the sums are not zero-filled
  double* part_inertia;     // [gridDim.x]
  uint32_t* mixed_list;     // [gridDim.x][list_cap] tile indices
  uint32_t* mixed_count;    // [gridDim.x]
  int list_cap;
  uint32_t pack_mult;   // 256, passed as an argument so that the score pack stays an IMAD (see the E role)
  // rows the certificate cannot decide: appended here for rows_top2_kernel (lkm_rows.cuh) instead of the inline exact scan
  // (nullptr: scanned inline, one warp per row)
  uint32_t* rescan_list;
  unsigned int* rescan_count;
};

// appends the rows of this warp whose verdict is 2 to the global list (one atomic per warp)
__device__ __forceinline__ unsigned int t5_defer(const Tc5Args& a, bool undecided, unsigned long long grow, int lane) {
  const uint32_t fm = __ballot_sync(0xffffffffu, undecided);
  if (fm == 0u) return 0u;
  unsigned int base = 0;
  if (lane == 0) base = atomicAdd(a.rescan_count, (unsigned int)__popc(fm));
  base = __shfl_sync(0xffffffffu, base, 0);
  if (undecided) a.rescan_list[base + __popc(fm & ((1u << lane) - 1u))] = (uint32_t)grow;
  return lane == 0 ? (unsigned int)__popc(fm) : 0u;
}

// LKM_T5_TRACE: CTA 0 / 1 stamp event `ev` of tile `it` (cycles since the roles started) into a.prof[(cta * 16 + it) * 16 + ev]
#ifdef LKM_T5_TRACE
#define T5TRACE(it, ev) do { if (blockIdx.x < 2 && lane == 0 && (it) >= 32 && (it) < 48 && a.prof) a.prof[((blockIdx.x * 16) + ((it) - 32)) * 16 + (ev)] = clock64() - trace_t0; } while (0)   /* tiles 32-47: steady state */
#else
#define T5TRACE(it, ev)
#endif

__host__ __device__ inline size_t tc5_smem_bytes(int d) {
  return (size_t)kT5Stages * 16384 + (size_t)(d / 32) * 16384 * 2 + 16384 + 1024 + kT5K * 4 + 2 * 128 * 4 + 1024;
}

// max_i |x_i|^2 for any row width (d % 4 == 0): one warp per row; a non-finite norm is published as NaN bits
__global__ void __launch_bounds__(256) row_norm2_max_wide_kernel(const float* __restrict__ X, unsigned long long n, int d,
                                                                 unsigned int* __restrict__ out_bits) {
  const int lane = threadIdx.x & 31;
  const unsigned long long w0 = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const unsigned long long nw = ((unsigned long long)gridDim.x * blockDim.x) >> 5;
  float m = 0.f;
  bool bad = false;
  for (unsigned long long r = w0; r < n; r += nw) {
    const float4* p = reinterpret_cast<const float4*>(X + r * (unsigned long long)d);
    float s = 0.f;
    for (int q = lane; q < d / 4; q += 32) {
      const float4 v = __ldg(p + q);
      s = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, s))));
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    bad |= !(s <= 3.0e38f);
    m = fmaxf(m, s);
  }
  if (lane == 0) atomicMax(out_bits, bad ? 0x7FC00000u : __float_as_uint(m));
}

namespace tc {
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
// the same with cluster-scope release: orders this thread's earlier observations (a landed TMA slab) before the arrival
__device__ __forceinline__ void mbar_arrive_cluster_release(uint64_t* bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.release.cluster.shared::cluster.b64 _, [ra];\n\t}"
      ::"r"(smem_u32(bar)), "r"(cta)
      : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* slot, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(slot)), "r"(cols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// M = 256 over the CTA pair; the whole (leader) warp executes these, one elected lane issues
__device__ __forceinline__ void mma2_tf32_ts_elect(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
      ::"r"(d_tmem), "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
__device__ __forceinline__ void mma2_tf32_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accum) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accum)
      : "memory");
}
// one arrival on the mbarrier at this offset in BOTH CTAs once every MMA issued so far has completed
__device__ __forceinline__ void mma2_commit_elect(uint64_t* bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t.reg .b16 m;\n\t"
      "mov.b16 m, 3;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], m;\n\t}"
      ::"r"(smem_u32(bar))
      : "memory");
}
}  // namespace tc


// ---- pieces shared by the pair kernels -------------------------------------------------------------------------------
struct T5Win {
  float K0, pad_bias, m_v;
  uint32_t top_bits;
  bool ok;
};
// Score window and certificate margin (identical in every thread of both CTAs).  `dropped` = bound, in units of
// max|x| max|c|, on the product terms the scoring leaves out for BOTH candidates together (0 for 3xTF32).
template <int D>
__device__ __forceinline__ T5Win t5_window(float xmax2, float cmax2_raw, float dropped) {
  T5Win w;
  const float cmax2 = cmax2_raw * 1.00001f;
  const float cmax = sqrtf(cmax2) * 1.000001f;
  const float xmax = sqrtf(xmax2) * 1.000001f;
  const float Rb = 1.01f * (xmax * cmax + 0.5f * cmax2) + 1e-30f;   // |x.c - |c|^2/2| <= Rb
  w.ok = Rb < 1e30f;   // false for NaN / Inf: every row then takes the exact scan
  // smallest even exponent field e with 2^e >= Rb; K0 = 2^(e+1), so v = x.c - |c|^2/2 + K0 lies in [2^e, 3 * 2^e):
  // exponent field e or e + 1, the top 8 bits of every score are the same and (bits << 8) stays monotone
  uint32_t e_w = ((__float_as_uint(Rb) >> 23) + 2u) & ~1u;
  e_w = min(max(e_w, 2u), 250u);
  w.K0 = __uint_as_float((e_w + 1u) << 23);
  w.pad_bias = __uint_as_float(e_w << 23);
  w.top_bits = (e_w << 23) & 0xFF000000u;
  const float eta = (3.f * (float)D + 12.f) * 2.3841858e-7f;   // (3d + 12) * 2^-22
  const float gamma = ((float)D + 2.f) * 5.9604645e-8f;        // (d + 2) * 2^-24
  const float m_w = 1.1f * (2.3841858e-7f * cmax2 + xmax * (4.f * eta + 4.7683716e-7f) * cmax +
                            2.f * gamma * (xmax + cmax) * (xmax + cmax));
  w.m_v = 0.5f * m_w + 7.f * 1.1920929e-7f * (w.K0 + Rb) + ((float)D + 1.f) * 5.9604645e-8f * cmax2 + dropped * xmax * cmax;
  return w;
}

// squared norms (8 threads per row): their maximum over ALL k centroids (bits) into *cmax_bits — every chunk launch
// must use the same score window —, the norms of this launch's chunk into nn_s[256]; all threads of the CTA call it
template <int D>
__device__ __forceinline__ void t5_centroid_norms(const float* __restrict__ C, int k, int chunk_base, int chunk_n, float* nn_s,
                                                  uint32_t* cmax_bits, int t) {
  for (int base = 0; base < k * 8; base += (int)blockDim.x) {
    const int item = base + t;
    const int r = item >> 3, c8 = item & 7;
    float nn = 0.f;
    if (r < k) {
#pragma unroll
      for (int q = 0; q < D / 32; ++q) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(C + (size_t)r * D) + c8 + 8 * q);
        nn = fmaf(v.x, v.x, fmaf(v.y, v.y, fmaf(v.z, v.z, fmaf(v.w, v.w, nn))));
      }
    }
    nn += __shfl_xor_sync(0xffffffffu, nn, 1);
    nn += __shfl_xor_sync(0xffffffffu, nn, 2);
    nn += __shfl_xor_sync(0xffffffffu, nn, 4);
    if (r < k && c8 == 0) {
      if (r >= chunk_base && r < chunk_base + chunk_n) nn_s[r - chunk_base] = nn;   // |nn - |c|^2| <= (d + 1) 2^-24 |c|^2
      atomicMax(cmax_bits, __float_as_uint(nn));   // a NaN norm wins and voids every certificate
    }
  }
}

// bias row t (K0 - |c|^2/2 as three TF32 pieces that add up to the f32 value exactly) and the constant ones atom
__device__ __forceinline__ void t5_write_bias(unsigned char* b_bias, unsigned char* ones, const float* nn_s, const T5Win& w,
                                              int c, int k, int t) {
  const float b = (c < k) ? (w.K0 - 0.5f * nn_s[c]) : w.pad_bias;
  const float h = tc::trunc_tf32(b);
  const float l = b - h;
  const float l1 = tc::trunc_tf32(l);
  const float l2 = l - l1;
  float* p = reinterpret_cast<float*>(b_bias + tc::sw128_offset((uint32_t)t, 0));
  p[0] = h; p[1] = l1; p[2] = l2;
  if (t < 8) {
    float* o1 = reinterpret_cast<float*>(ones + tc::sw128_offset((uint32_t)t, 0));
    o1[0] = 1.f; o1[1] = 1.f; o1[2] = 1.f;
  }
}

// Packed top-2 of the 256 scores of this thread's row: q = bits(v) << 8 | c.  Integer min / max / add run on the
// 16-lane ALU pipe (two cycles per warp instruction), which this role saturates: the pack is kept an IMAD (FMA
// pipe) by passing the multiplier (256) as a kernel argument, and the runner-up pass is written as chains of
// min(m, cb - 1 - q) so that each step is one fused VIADDMNMX.
//
// Both pipes take one warp instruction every two cycles (fma: IMAD; alu: IADD3, VIMNMX3).  ptxas turns the runner-up
// pass into 32 IADD3 + one chain of 16 VIMNMX3 per 32 columns, which with the 16 VIMNMX3 of the first pass puts 72
// instructions on the alu pipe against the 32 IMADs of the pack.  Five of every eight subtractions are therefore written
// as q * neg1 + (best - 1) with neg1 = 0xFFFFFFFF derived from a kernel argument (an IMAD the compiler cannot fold
// back): 52 instructions on each pipe (cfg4 shard step 4.20 -> 3.87 ms).
//
// The 32-column TMEM loads are software-pipelined: the load of chunk i + 1 is in flight while chunk i is scored (under
// the MMAs of the other accumulator a tcgen05.ld takes several hundred cycles, and a warp that waits for each load before
// it computes spends most of a tile waiting: ~540 cycles per chunk for ~140 cycles of pipe work).  tmem_ld_fence ties
// the destination registers to the wait so that no use is scheduled above it.
template <int N>
__device__ __forceinline__ void t5_score_chunk(uint32_t (&q)[32], int c0, uint32_t mult, uint32_t neg1, uint32_t& m1, uint32_t& m2);

// `drained()` runs as soon as the last chunk sits in registers, before it is scored: the accumulator goes back to the MMA
// role one chunk earlier.
template <int N, typename Drained>
__device__ __forceinline__ void t5_top2(uint32_t taddr, uint32_t mult, uint32_t neg1, uint32_t& m1, uint32_t& m2, Drained drained) {
  static_assert(N % 64 == 0, "two chunks per round");
  m1 = 0u; m2 = 0u;   // real scores are >= 2^e > 0
  uint32_t qa[32], qb[32];
  tc::tmem_ld32(taddr, qa);
#pragma unroll 1
  for (int c0 = 0; c0 < N; c0 += 64) {
    tc::tmem_ld_fence(qa);
    tc::tmem_ld32(taddr + (uint32_t)c0 + 32u, qb);
    t5_score_chunk<N>(qa, c0, mult, neg1, m1, m2);
    tc::tmem_ld_fence(qb);
    if (c0 + 64 < N) tc::tmem_ld32(taddr + (uint32_t)c0 + 64u, qa);
    else drained();
    t5_score_chunk<N>(qb, c0 + 32, mult, neg1, m1, m2);
  }
}

template <int N>
__device__ __forceinline__ void t5_score_chunk(uint32_t (&q)[32], int c0, uint32_t mult, uint32_t neg1, uint32_t& m1, uint32_t& m2) {
  {
#pragma unroll
    for (int c = 0; c < 32; ++c) q[c] = q[c] * mult + (uint32_t)c;   // index inside the chunk; c0 is added to the winners
    uint32_t b0 = __vimax3_u32(q[0], q[1], q[2]), b1 = __vimax3_u32(q[3], q[4], q[5]);
    uint32_t b2 = __vimax3_u32(q[6], q[7], q[8]), b3 = __vimax3_u32(q[9], q[10], q[11]);
#pragma unroll
    for (int c = 12; c + 7 < 32; c += 8) {
      b0 = __vimax3_u32(b0, q[c], q[c + 1]);
      b1 = __vimax3_u32(b1, q[c + 2], q[c + 3]);
      b2 = __vimax3_u32(b2, q[c + 4], q[c + 5]);
      b3 = __vimax3_u32(b3, q[c + 6], q[c + 7]);
    }
    b0 = __vimax3_u32(b0, q[28], q[29]);
    b1 = __vimax3_u32(b1, q[30], q[31]);
    const uint32_t cb = max(__vimax3_u32(b0, b1, b2), b3);
    const uint32_t cm1 = cb - 1u;
    // the chunk's best wraps to 0xFFFFFFFF, the others become gap - 1.  Four independent chains of three-input minima
    // (ptxas folds two-input chains into ONE chain of 16 dependent VIMNMX3: 16 x 4.5 cycles of latency per chunk)
    uint32_t gp[32];
#pragma unroll
    for (int c = 0; c < 32; ++c) gp[c] = ((c & 7) < 5) ? q[c] * neg1 + cm1 : cm1 - q[c];
    uint32_t t0 = __vimin3_u32(gp[0], gp[4], gp[8]), t1 = __vimin3_u32(gp[1], gp[5], gp[9]);
    uint32_t t2 = __vimin3_u32(gp[2], gp[6], gp[10]), t3 = __vimin3_u32(gp[3], gp[7], gp[11]);
    t0 = __vimin3_u32(t0, gp[12], gp[16]); t1 = __vimin3_u32(t1, gp[13], gp[17]);
    t2 = __vimin3_u32(t2, gp[14], gp[18]); t3 = __vimin3_u32(t3, gp[15], gp[19]);
    t0 = __vimin3_u32(t0, gp[20], gp[24]); t1 = __vimin3_u32(t1, gp[21], gp[25]);
    t2 = __vimin3_u32(t2, gp[22], gp[26]); t3 = __vimin3_u32(t3, gp[23], gp[27]);
    t0 = min(t0, gp[28]); t1 = min(t1, gp[29]); t2 = min(t2, gp[30]); t3 = min(t3, gp[31]);
    const uint32_t cs = cm1 - min(__vimin3_u32(t0, t1, t2), t3) + (uint32_t)c0;   // runner-up of the chunk
    const uint32_t cbg = cb + (uint32_t)c0;
    m2 = max(max(m2, cs), min(m1, cbg));
    m1 = max(m1, cbg);
  }
}

// The reference's scan of row fr (direct form in feature order, strict <) by one warp: each lane walks the centroids
// c = lane, lane + 32, ... (256 at a time) with the row read from global memory 32 features at a time.  Returns the label.
template <int D>
__device__ __forceinline__ int t5_exact_label(const float* __restrict__ X, const float* __restrict__ C, int k, unsigned long long fr,
                                              int lane) {
  const float* xrow = X + fr * (unsigned long long)D;
  float fb = INFINITY;
  int fi = -1;
#pragma unroll 1
  for (int cb = 0; cb < k; cb += kT5K) {
    float dacc[kT5K / 32];
#pragma unroll
    for (int i = 0; i < kT5K / 32; ++i) dacc[i] = 0.f;
#pragma unroll 1
    for (int j0 = 0; j0 < D; j0 += 32) {
      float x2[32];
#pragma unroll
      for (int c = 0; c < 8; ++c) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(xrow + j0) + c);
        x2[4 * c] = v.x; x2[4 * c + 1] = v.y; x2[4 * c + 2] = v.z; x2[4 * c + 3] = v.w;
      }
#pragma unroll
      for (int i = 0; i < kT5K / 32; ++i) {
        const int c = cb + lane + 32 * i;
        if (c < k) {
          const float* cw = C + (size_t)c * D + j0;
          float dd = dacc[i];
#pragma unroll
          for (int q4 = 0; q4 < 8; ++q4) {
            const float4 cv = __ldg(reinterpret_cast<const float4*>(cw) + q4);
            float df;
            df = __fsub_rn(cv.x, x2[4 * q4]); dd = __fadd_rn(dd, __fmul_rn(df, df));
            df = __fsub_rn(cv.y, x2[4 * q4 + 1]); dd = __fadd_rn(dd, __fmul_rn(df, df));
            df = __fsub_rn(cv.z, x2[4 * q4 + 2]); dd = __fadd_rn(dd, __fmul_rn(df, df));
            df = __fsub_rn(cv.w, x2[4 * q4 + 3]); dd = __fadd_rn(dd, __fmul_rn(df, df));
          }
          dacc[i] = dd;
        }
      }
    }
#pragma unroll
    for (int i = 0; i < kT5K / 32; ++i) {
      const int c = cb + lane + 32 * i;
      if (c < k) {
        // lane 0 starts with centroid 0 (sticky NaN like the reference); the others only take v < best
        if (c == 0) { fb = dacc[i]; fi = 0; }
        else if (dacc[i] < fb) { fb = dacc[i]; fi = c; }
      }
    }
  }
  combine_candidates<float>(32, fb, fi);
  return fi;
}

// Folds this launch's packed (best, runner-up) of row `grow` into the state of the earlier chunks; on the last chunk
// decides: 1 = certified (label set), 2 = needs the exact re-scan, 0 = nothing to do yet / row past the end.
__device__ __forceinline__ int t5_merge_certify(const Tc5Args& a, const T5Win& win, unsigned long long grow, uint32_t m1, uint32_t m2,
                                                const uint4 p, uint32_t& label) {
  if (grow >= a.n) return 0;
  uint32_t ch1 = (uint32_t)a.chunk_idx;
  if (a.chunk_idx > 0) {
    // packed values order by score first; equal scores in different chunks end as best / runner-up with gap 0: uncertified
    if (m1 > p.x) { m2 = max(p.x, m2); }
    else { m2 = max(p.y, m1); m1 = p.x; ch1 = p.z; }
  }
  if (a.chunk_idx + 1 < a.n_chunks) {
    a.merge[grow] = make_uint4(m1, m2, ch1, 0u);
    return 0;
  }
  const float v_best = __uint_as_float((m1 >> 8) | win.top_bits);
  const float v_second = __uint_as_float((m2 >> 8) | win.top_bits);
  label = ch1 * (uint32_t)a.chunk_n + (m1 & 255u);
  return (win.ok && ((v_best - v_second) > win.m_v) && label < (uint32_t)a.k) ? 1 : 2;
}

template <int D>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kT5Threads, 1)
    assign_tc5_kernel(const Tc5Args a, const __grid_constant__ CUtensorMap tmap) {
  static_assert(D == 64 || D == 128, "row width");
  constexpr int NS = D / 32;    // 32-feature slabs per tile
  constexpr int KS = D / 8;     // K-steps per product
  if (a.state != nullptr && a.state->done) return;
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (tc::smem_u32(smem_unaligned) & 1023u)) & 1023u);
  const int t = threadIdx.x, lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int chunk_base = a.chunk_idx * kT5K;
  const int k = min(kT5K, a.k - chunk_base);   // centroids of this launch's chunk
  const float* Cc = a.C + (size_t)chunk_base * D;
  const uint32_t rank = tc::cluster_ctarank();
  unsigned char* ring = smem;                                   // [kT5Stages][16 KB] raw slabs
  unsigned char* b_hi = ring + (size_t)kT5Stages * 16384;       // [NS][16 KB]: this CTA's 128 centroids, K-major SW128
  unsigned char* b_lo = b_hi + (size_t)NS * 16384;
  unsigned char* b_bias = b_lo + (size_t)NS * 16384;            // K-step 0 only
  unsigned char* ones = b_bias + 16384;                         // constant A operand of the bias K-step (1 KB atom)
  float* nn_s = reinterpret_cast<float*>(ones + 1024);          // [256] squared norms of all centroids
  uint32_t* flags_all = reinterpret_cast<uint32_t*>(nn_s + kT5K);   // [2][128] uncertified rows of the current / previous tile
  __shared__ __align__(8) uint64_t bar_full[kT5Stages], bar_empty[kT5Stages], bar_aready, bar_mma, bar_accfree;
  __shared__ uint32_t tmem_slot, n_flag[2], cmax_bits;

  // pair-tile pt = two consecutive 128-row tiles; CTA `rank` of the pair takes tile 2 pt + rank
  const unsigned long long n_tiles = (a.n + 127) / 128;
  const unsigned long long n_pt = (n_tiles + 1) / 2;
  const unsigned long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const unsigned long long my_pt = pair < n_pt ? (n_pt - pair + n_pairs - 1) / n_pairs : 0;
  auto tile_of = [&](unsigned long long it) { return 2ull * (pair + it * n_pairs) + rank; };

  if (warp == 8) tc::tmem_alloc_pair(&tmem_slot, 512);
  if (t == 0) {
    for (int s = 0; s < kT5Stages; ++s) {
      tc::mbar_init(&bar_full[s], 1);
      tc::mbar_init(&bar_empty[s], 4);
    }
    tc::mbar_init(&bar_aready, 8);
    tc::mbar_init(&bar_mma, 1);
    tc::mbar_init(&bar_accfree, 8);
    tc::fence_mbar_init();
    n_flag[0] = 0;
    n_flag[1] = 0;
    cmax_bits = 0;
  }
  for (int e = t; e < 1024 + 64; e += kT5Threads) reinterpret_cast<float4*>(b_bias)[e] = make_float4(0.f, 0.f, 0.f, 0.f);   // bias tile + ones atom
  __syncthreads();
  t5_centroid_norms<D>(a.C, a.k, chunk_base, kT5K, nn_s, &cmax_bits, t);
  // ---- this CTA's half of the centroids as TF32 hi / lo operand slabs ----
  for (int item = t; item < 128 * NS * 8; item += kT5Threads) {
    const int q = item & 7, s = (item >> 3) % NS, r = item / (8 * NS);
    const int c = (int)rank * 128 + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (c < k) v = __ldg(reinterpret_cast<const float4*>(Cc + (size_t)c * D + s * 32) + q);
    float4 h, l;
    tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y);
    tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
    const uint32_t o = (uint32_t)s * 16384u + tc::sw128_offset((uint32_t)r, (uint32_t)q);
    *reinterpret_cast<float4*>(b_hi + o) = h;
    *reinterpret_cast<float4*>(b_lo + o) = l;
  }
  __syncthreads();
  const T5Win win = t5_window<D>(__uint_as_float(__ldg(a.xmax2_bits)), __uint_as_float(cmax_bits), 0.f);
  if (t < 128) t5_write_bias(b_bias, ones, nn_s, win, (int)rank * 128 + t, k, t);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();   // both CTAs: barriers initialised, operand tiles written, TMEM allocated
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  unsigned int my_fallbacks = 0;
#ifdef LKM_T5_TRACE
  const long long trace_t0 = clock64();
#endif

  if (warp == 9) {
    // =============================== P: stream the slabs ===============================
    if (lane == 0) {
      int s = 0;
      uint32_t use = 0;
      for (unsigned long long it = 0; it < my_pt; ++it) {
        const unsigned long long tile = min(tile_of(it), n_tiles - 1);   // a pair-tile past the end re-reads the last tile
        for (int sl = 0; sl < NS; ++sl) {
          if (use > 0) tc::mbar_wait_or_trap(&bar_empty[s], (use - 1u) & 1u);
          tc::mbar_expect_tx(&bar_full[s], 16384u);
          tc::tma_load_2d(ring + (size_t)s * 16384, &tmap, sl * 32, (int)(tile * 128), &bar_full[s]);
          if (sl == 0 && it + kT5Prefetch < my_pt) {
            const unsigned long long r0 = tile_of(it + kT5Prefetch) * 128ull;   // 128 full rows are contiguous
            if (r0 < a.n) tc::l2_prefetch(a.X + r0 * (unsigned long long)D, (uint32_t)min(128ull, a.n - r0) * (uint32_t)(D * 4));
          }
          if (++s == kT5Stages) { s = 0; ++use; }
        }
      }
    }
  } else if (warp == 8) {
    // =============================== M: issue the MMAs (leader CTA) ===============================
    if (rank == 0) {
      constexpr uint32_t idesc = tc::idesc_tf32(256, (uint32_t)kT5K);
      const uint64_t desc_bhi = tc::smem_desc_sw128(tc::smem_u32(b_hi));
      const uint64_t desc_blo = tc::smem_desc_sw128(tc::smem_u32(b_lo));
      const uint64_t desc_bias = tc::smem_desc_sw128(tc::smem_u32(b_bias));
      const uint64_t desc_ones = tc::smem_desc_sw128_alias(tc::smem_u32(ones));
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      const uint32_t dcol = tmem_u + kT5ColAcc, a_hi = tmem_u + kT5ColA, a_lo = tmem_u + kT5ColA + (uint32_t)D;
      for (unsigned long long it = 0; it < my_pt; ++it) {
        tc::mbar_wait_or_trap(&bar_aready, (uint32_t)it & 1u);
        T5TRACE(it, 0);
        if (it > 0) tc::mbar_wait_or_trap(&bar_accfree, (uint32_t)(it - 1) & 1u);
        T5TRACE(it, 1);
        tc::tc_fence_after();
        // B K-step ks: slab ks / 4 (16 KB = 1024 descriptor units), 32 bytes = 2 units inside the slab
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
          tc::mma2_tf32_ts_elect(dcol, a_lo + 8u * ks, desc_bhi + (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2), idesc, ks ? 1u : 0u);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
          tc::mma2_tf32_ts_elect(dcol, a_hi + 8u * ks, desc_blo + (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2), idesc, 1u);
#pragma unroll
        for (int ks = 0; ks < KS; ++ks)
          tc::mma2_tf32_ts_elect(dcol, a_hi + 8u * ks, desc_bhi + (uint64_t)((ks >> 2) * 1024 + (ks & 3) * 2), idesc, 1u);
        tc::mma2_tf32_elect(dcol, desc_ones, desc_bias, idesc, 1u);
        tc::mma2_commit_elect(&bar_mma);
        T5TRACE(it, 2);
      }
    }
  } else if (warp < 4) {
    // =============================== S: the A operands of the tile into TMEM ===============================
    int s = 0;
    uint32_t full_ph = 0;
    const uint32_t row_off = (uint32_t)(t >> 3) * 1024u + (uint32_t)(t & 7) * 128u;   // row t of a SW128 slab
    const uint32_t row_x = (uint32_t)(t & 7);
    const uint32_t lane_base = tmem + ((uint32_t)(warp * 32) << 16);
    uint32_t tf32_mask = 0xFFFFE000u, sign_bit = 0x80000000u;
    asm volatile("" : "+r"(tf32_mask), "+r"(sign_bit));
    for (unsigned long long it = 0; it < my_pt; ++it) {
      if (it > 0) tc::mbar_wait_or_trap(&bar_mma, (uint32_t)(it - 1) & 1u);   // the MMAs of the previous tile have read A
      if (warp == 0) T5TRACE(it, 3);
      tc::tc_fence_after();
#pragma unroll 1
      for (int sl = 0; sl < NS; ++sl) {
        const unsigned char* rw = ring + (size_t)s * 16384;
        tc::mbar_wait_or_trap(&bar_full[s], full_ph);
        uint32_t xr[32];
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          const uint4 v = *reinterpret_cast<const uint4*>(rw + row_off + (((uint32_t)c ^ row_x) << 4));
          xr[4 * c] = v.x; xr[4 * c + 1] = v.y; xr[4 * c + 2] = v.z; xr[4 * c + 3] = v.w;
        }
        tc::tmem_st32(lane_base + kT5ColA + 32u * sl, xr);
#pragma unroll
        for (int q = 0; q < 32; q += 2) {
          const uint32_t n0 = tc::neg_trunc_tf32_bits(xr[q], tf32_mask, sign_bit);
          const uint32_t n1 = tc::neg_trunc_tf32_bits(xr[q + 1], tf32_mask, sign_bit);
          tc::f2_add_parts(xr[q], xr[q + 1], n0, n1, xr[q], xr[q + 1]);
        }
        tc::tmem_st32(lane_base + kT5ColA + (uint32_t)D + 32u * sl, xr);
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bar_empty[s]);
        if (++s == kT5Stages) { s = 0; full_ph ^= 1u; }
      }
      if (warp == 0) T5TRACE(it, 4);
      tc::tmem_st_wait();
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) tc::mbar_arrive_cluster(&bar_aready, 0u);
      if (warp == 0) T5TRACE(it, 5);
    }
  } else {
    // =============================== E: score the rows ===============================
    const int wq = warp & 3;
    const int row = t - 128;
    const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + kT5ColAcc;
    for (unsigned long long it = 0; it < my_pt; ++it) {
      const unsigned long long row0 = tile_of(it) * 128ull;
      // state of the earlier chunks: requested before the wait so that its latency hides behind the MMAs
      uint4 prev = make_uint4(0u, 0u, 0u, 0u);
      if (a.chunk_idx > 0 && row0 + (unsigned long long)row < a.n) prev = a.merge[row0 + (unsigned long long)row];
      tc::mbar_wait_or_trap(&bar_mma, (uint32_t)it & 1u);
      if (wq == 0) T5TRACE(it, 6);
      tc::tc_fence_after();
      uint32_t m1, m2;   // best / runner-up (packed)
      t5_top2<kT5K>(taddr, a.pack_mult, 0u - (a.pack_mult >> 8), m1, m2, [&]() {
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_cluster(&bar_accfree, 0u);   // the accumulator may be overwritten
      });
      if (wq == 0) T5TRACE(it, 7);
      const unsigned long long grow = row0 + (unsigned long long)row;
      uint32_t bidx = 0u;
      const int verdict = t5_merge_certify(a, win, grow, m1, m2, prev, bidx);
      // flag lists are double-buffered by tile parity: the other list's counter is cleared before this tile's
      // barrier, so that clearing and the next tile's appends are ordered by it
      uint32_t* flags = flags_all + (it & 1) * 128;
      if (a.rescan_list != nullptr) {
        if (verdict) a.labels[grow] = bidx;   // provisional for verdict 2: rows_top2_kernel overwrites it
        my_fallbacks += t5_defer(a, verdict == 2, grow, lane);
        if (wq == 0) T5TRACE(it, 8);
        continue;
      }
      if (verdict) {
        if (verdict == 1) a.labels[grow] = bidx;
        else flags[atomicAdd(&n_flag[it & 1], 1u)] = (uint32_t)row;
      }
      if (row == 0) n_flag[(it + 1) & 1] = 0u;
      tc::nb_sync(1, 128);
      if (wq == 0) T5TRACE(it, 8);
      const uint32_t nf = n_flag[it & 1];
      if (nf) {
        // uncertified rows: the reference's exact scan, one warp per row
        for (uint32_t f = (uint32_t)wq; f < nf; f += 4) {
          const unsigned long long fr = row0 + flags[f];
          const int fi = t5_exact_label<D>(a.X, a.C, a.k, fr, lane);
          if (lane == 0) a.labels[fr] = (uint32_t)fi;
        }
        my_fallbacks += (row == 0) ? nf : 0u;
      }
    }
    if (a.state && my_fallbacks) atomicAdd(&a.state->fallback_rows, (unsigned long long)my_fallbacks);
  }
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();   // the peer may still be reading this CTA's operands / arriving on its barriers
  if (warp == 8) tc::tmem_dealloc_pair(tmem, 512);
}


// =============================================================================================================
// SCREENING variant of the pair kernel: ONE TF32 product (x_hi . c_hi + bias), the raw TMA slabs are the A operand.
// The terms it leaves out (x_lo.c_hi, x_hi.c_lo, x_lo.c_lo <= (2^-9 + 2^-20) |x||c| per candidate) widen the
// certificate; rows it cannot certify take the exact re-scan, so the labels are the same as with 3xTF32.  On data whose
// clusters are separated by more than ~1 % of max|x| max|c| every row certifies and the kernel is HBM-bound
// (17 MMAs per pair of tiles instead of 49, no operand staging by threads); the host probes one iteration and
// falls back to the 3xTF32 kernel when more than 2 % of the rows had to be re-scanned.
//
//   P  warp 9     TMA ring of eight 32-feature slabs (two tiles in flight)
//   M  warp 8     leader: per slab, waits for its own and (relayed by the peer's warp 8) the peer's copy, issues the
//                 slab's 4 K-steps, and commits the slab back to both rings; after the tile's bias K-step a second
//                 commit hands accumulator it & 1 to the E group of that parity.  Peer: relays its `full` barriers.
//   E  warps 0-3 (even tiles) / 4-7 (odd tiles): as in the 3xTF32 kernel, double-buffered accumulators
// =============================================================================================================
constexpr int kT5sStages = 8;

__host__ __device__ inline size_t tc5s_smem_bytes(int d) {
  return (size_t)kT5sStages * 16384 + (size_t)(d / 32) * 16384 + 16384 + 1024 + kT5K * 4 + 4 * 128 * 4 + 1024;
}

template <int D, bool FUSE, int N>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(FUSE ? t5_fuse_threads(D) : kT5Threads, 1)
    assign_tc5s_kernel(const Tc5Args a, const __grid_constant__ CUtensorMap tmap) {
  static_assert(D == 32 || D == 64 || D == 128, "row width");
  static_assert(N == 64 || N == 128 || N == 256, "score columns per launch (N / 2 centroid rows per CTA)");
  constexpr int NS = D / 32;
  constexpr int NT = FUSE ? t5_fuse_threads(D) : kT5Threads;
  if (a.state != nullptr && a.state->done) return;
  extern __shared__ __align__(16) unsigned char smem_unaligned[];
  unsigned char* smem = smem_unaligned + ((1024u - (tc::smem_u32(smem_unaligned) & 1023u)) & 1023u);
  const int t = threadIdx.x, lane = t & 31;
  const int warp = __shfl_sync(0xffffffffu, t >> 5, 0);
  const int chunk_base = a.chunk_idx * N;
  const int k = min(N, a.k - chunk_base);   // centroids of this launch's chunk
  const float* Cc = a.C + (size_t)chunk_base * D;
  const uint32_t rank = tc::cluster_ctarank();
  unsigned char* ring = smem;                                    // [kT5sStages][16 KB] raw slabs = A operands
  unsigned char* b_hi = ring + (size_t)kT5sStages * 16384;       // [NS][16 KB]
  unsigned char* b_bias = b_hi + (size_t)NS * 16384;
  unsigned char* ones = b_bias + 16384;
  float* nn_s = reinterpret_cast<float*>(ones + 1024);
  uint32_t* flags_all = reinterpret_cast<uint32_t*>(nn_s + kT5K);   // [group][2][128]
  __shared__ __align__(8) uint64_t bar_full[kT5sStages], bar_pfull[kT5sStages], bar_empty[kT5sStages], bar_mma[2], bar_accfree[2];
  __shared__ uint32_t tmem_slot, n_flag[2][2], cmax_bits;
  __shared__ __align__(8) uint64_t bar_verdict[4];   // FUSE: E -> U, one per tile slot (it & 3)
  __shared__ __align__(8) uint64_t bar_vfree[4];     // FUSE: U -> E, the slot's previous verdict has been consumed
  __shared__ __align__(16) uint32_t warp_verdict[4][4];            // [tile slot][E warp]: common label of the warp's 32 rows, or a marker
  __shared__ double u_red[4];

  const unsigned long long n_tiles = (a.n + 127) / 128;
  const unsigned long long n_pt = (n_tiles + 1) / 2;
  const unsigned long long pair = blockIdx.x >> 1, n_pairs = gridDim.x >> 1;
  const unsigned long long my_pt = pair < n_pt ? (n_pt - pair + n_pairs - 1) / n_pairs : 0;
  auto tile_of = [&](unsigned long long it) { return 2ull * (pair + it * n_pairs) + rank; };

  if (warp == 8) tc::tmem_alloc_pair(&tmem_slot, 512);
  if (t == 0) {
    for (int s = 0; s < kT5sStages; ++s) {
      tc::mbar_init(&bar_full[s], 1);
      tc::mbar_init(&bar_pfull[s], 1);
      tc::mbar_init(&bar_empty[s], FUSE ? 2 : 1);   // MMA commit (+ the U warp that sums the slab)
    }
    for (int q = 0; q < 4; ++q) { tc::mbar_init(&bar_verdict[q], 4); tc::mbar_init(&bar_vfree[q], NS); }
    for (int g = 0; g < 2; ++g) {
      tc::mbar_init(&bar_mma[g], 1);
      tc::mbar_init(&bar_accfree[g], 8);
      n_flag[g][0] = 0;
      n_flag[g][1] = 0;
    }
    tc::fence_mbar_init();
    cmax_bits = 0;
  }
  for (int e = t; e < 1024 + 64; e += NT) reinterpret_cast<float4*>(b_bias)[e] = make_float4(0.f, 0.f, 0.f, 0.f);
  // fused update: with a `touched` map (one bit per cluster in the lanes of every U warp, k <= 1024) the k x d partial is
  // neither zero-filled here nor read where it was never written — on ordered rows a CTA sees two or three clusters
  const bool sparse = FUSE && a.part_touched != nullptr && a.k <= 1024;
  if constexpr (FUSE) {
    if (!sparse) {
      float4* ps4 = reinterpret_cast<float4*>(a.part_sums + (size_t)blockIdx.x * a.k * D);
      for (int e = t; e < a.k * D / 4; e += NT) ps4[e] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (a.part_touched != nullptr)
        for (int e = t; e < a.k; e += NT) a.part_touched[(size_t)blockIdx.x * a.k + e] = 1;
    }
    for (int e = t; e < a.k; e += NT) a.part_counts[(size_t)blockIdx.x * a.k + e] = 0u;
  }
  __syncthreads();
  t5_centroid_norms<D>(a.C, a.k, chunk_base, N, nn_s, &cmax_bits, t);
  // this CTA's half of the centroids, raw: kind::tf32 reads them as c_hi
  for (int item = t; item < 128 * NS * 8; item += NT) {
    const int q = item & 7, s = (item >> 3) % NS, r = item / (8 * NS);
    const int c = (int)rank * (N / 2) + r;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (r < N / 2 && c < k) v = __ldg(reinterpret_cast<const float4*>(Cc + (size_t)c * D + s * 32) + q);
    *reinterpret_cast<float4*>(b_hi + (uint32_t)s * 16384u + tc::sw128_offset((uint32_t)r, (uint32_t)q)) = v;
  }
  __syncthreads();
  // both candidates together: 2 (2^-10 + 2^-10 + 2^-20) = 2^-8 (1 + 2^-11)
  const T5Win win = t5_window<D>(__uint_as_float(__ldg(a.xmax2_bits)), __uint_as_float(cmax_bits), 1.001f * 3.90625e-3f);
  if (t < 128) t5_write_bias(b_bias, ones, nn_s, win, (int)rank * (N / 2) + t, (t < N / 2) ? k : 0, t);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::cluster_sync_all();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_slot;
  unsigned int my_fallbacks = 0;
#ifdef LKM_T5_TRACE
  const long long trace_t0 = clock64();
#endif

  if (warp == 9) {
    // =============================== P ===============================
    if (lane == 0) {
      int s = 0;
      uint32_t use = 0;
      for (unsigned long long it = 0; it < my_pt; ++it) {
        const unsigned long long tile = min(tile_of(it), n_tiles - 1);
        for (int sl = 0; sl < NS; ++sl) {
          if (use > 0) tc::mbar_wait_spin(&bar_empty[s], (use - 1u) & 1u);
          if (sl == 0) T5TRACE(it, 0);
          if (sl == NS - 1) T5TRACE(it, 1);
          tc::mbar_expect_tx(&bar_full[s], 16384u);
          tc::tma_load_2d(ring + (size_t)s * 16384, &tmap, sl * 32, (int)(tile * 128), &bar_full[s]);
          if (sl == 0 && it + kT5Prefetch < my_pt) {
            const unsigned long long r0 = tile_of(it + kT5Prefetch) * 128ull;   // 128 full rows are contiguous
            if (r0 < a.n) tc::l2_prefetch(a.X + r0 * (unsigned long long)D, (uint32_t)min(128ull, a.n - r0) * (uint32_t)(D * 4));
          }
          if (++s == kT5sStages) { s = 0; ++use; }
        }
      }
    }
  } else if (warp == 8) {
    if (rank != 0) {
      // =============================== relay: this CTA's slab has landed -> leader ===============================
      int s = 0;
      uint32_t ph = 0;
      for (unsigned long long j = 0; j < my_pt * NS; ++j) {
        tc::mbar_wait_spin(&bar_full[s], ph);
        if (lane == 0) tc::mbar_arrive_cluster(&bar_pfull[s], 0u);   // plain arrive: a cluster-scope release costs a MEMBAR.GPU per slab (measured: 1000+ cycles, the bottleneck)
        __syncwarp();
        if (++s == kT5sStages) { s = 0; ph ^= 1u; }
      }
    } else {
      // =============================== M ===============================
      constexpr uint32_t idesc = tc::idesc_tf32(256, (uint32_t)N);
      const uint64_t desc_ring = tc::smem_desc_sw128(tc::smem_u32(ring));
      const uint64_t desc_bhi = tc::smem_desc_sw128(tc::smem_u32(b_hi));
      const uint64_t desc_bias = tc::smem_desc_sw128(tc::smem_u32(b_bias));
      const uint64_t desc_ones = tc::smem_desc_sw128_alias(tc::smem_u32(ones));
      const uint32_t tmem_u = __shfl_sync(0xffffffffu, tmem, 0);
      int s = 0;
      uint32_t ph = 0;
      for (unsigned long long it = 0; it < my_pt; ++it) {
        const uint32_t g = (uint32_t)it & 1u, u = (uint32_t)(it >> 1);
        const uint32_t dcol = tmem_u + g * 256u;
        T5TRACE(it, 15);
        if (u > 0) tc::mbar_wait_spin(&bar_accfree[g], (u - 1u) & 1u);
        T5TRACE(it, 2);
#pragma unroll 1
        for (int sl = 0; sl < NS; ++sl) {
          tc::mbar_wait_spin(&bar_full[s], ph);
          if (sl == 0) T5TRACE(it, 3);
          if (sl == NS - 1) T5TRACE(it, 5);
          tc::mbar_wait_spin(&bar_pfull[s], ph);
          if (sl == 0) T5TRACE(it, 4);
          if (sl == NS - 1) T5TRACE(it, 6);
          tc::tc_fence_after();
          const uint64_t da = desc_ring + (uint64_t)(s * 1024), db = desc_bhi + (uint64_t)(sl * 1024);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) tc::mma2_tf32_elect(dcol, da + 2u * ks, db + 2u * ks, idesc, (sl | ks) ? 1u : 0u);
          tc::mma2_commit_elect(&bar_empty[s]);   // both CTAs: the slab may be refilled
          if (++s == kT5sStages) { s = 0; ph ^= 1u; }
        }
        tc::mma2_tf32_elect(dcol, desc_ones, desc_bias, idesc, 1u);
        tc::mma2_commit_elect(&bar_mma[g]);
        T5TRACE(it, 7);
      }
    }
  } else if (warp >= 10) {
    // =============================== U: column sums of the raw slabs (FUSE) ===============================
    // U warp u owns slab u (features 32 u + lane).  For every tile it keeps, relative to the tile's first row x0,
    // S' = sum (x - x0) and Q' = sum (x - x0)^2 over the 128 rows (no cancellation: rows of one cluster differ by O(sigma)).
    // Two tiles later the E group has published whether the tile's rows all certified to ONE cluster L; then
    // sum += S' + 128 x0, count += 128 and inertia += Q' + d (128 d - 2 S') with d = c_L - x0 go to the register
    // accumulators of the current cluster (flushed to the CTA's partial sums when the cluster changes); any other tile is
    // listed for update_rows_kernel.  Tiles and clusters are visited in a fixed order: reproducible.
    if constexpr (FUSE) {
      const int u = warp - 10;
      const int feat = 32 * u + lane;
      uint32_t off[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) off[q] = (uint32_t)q * 128u + ((((uint32_t)lane >> 2) ^ (uint32_t)q) << 4) + ((uint32_t)lane & 3u) * 4u;
      float xa0 = 0.f, xa1 = 0.f, xa2 = 0.f, xa3 = 0.f, Sa0 = 0.f, Sa1 = 0.f, Sa2 = 0.f, Sa3 = 0.f, Qa0 = 0.f, Qa1 = 0.f, Qa2 = 0.f, Qa3 = 0.f;
      uint32_t cur_L = 0xFFFFFFFFu, cur_cnt = 0u, n_mixed = 0u;
      float cur_sum = 0.f, cur_c = 0.f, inert_f = 0.f;
      double inert_d = 0.0;
      float* my_sums = a.part_sums + (size_t)blockIdx.x * a.k * D + feat;
      uint32_t* my_counts = a.part_counts + (size_t)blockIdx.x * a.k;
      uint32_t* my_list = a.mixed_list + (size_t)blockIdx.x * a.list_cap;
      uint32_t seen_bits = 0u;   // lane l: clusters 32 l .. 32 l + 31 this CTA has flushed (identical in every U warp)
      auto flush = [&]() {
        if (cur_L != 0xFFFFFFFFu) {
          bool seen = true;
          if (sparse) {
            seen = (__shfl_sync(0xffffffffu, seen_bits, (int)(cur_L >> 5)) >> (cur_L & 31u)) & 1u;
            if ((uint32_t)lane == (cur_L >> 5)) seen_bits |= 1u << (cur_L & 31u);
          }
          if (seen) my_sums[(size_t)cur_L * D] += cur_sum;
          else my_sums[(size_t)cur_L * D] = cur_sum;
          if (u == 0 && lane == 0) my_counts[cur_L] += cur_cnt;
        }
      };
      auto consume = [&](unsigned long long v) {
        // verdict of tile v (slot v & 3), published by the four E warps of group v & 1
        tc::mbar_wait_or_trap(&bar_verdict[v & 3], (uint32_t)(v >> 2) & 1u);
        const uint4 wv = *reinterpret_cast<const uint4*>(warp_verdict[v & 3]);
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bar_vfree[v & 3]);
        const int vs = (int)(v & 3);
        const float x0 = vs == 0 ? xa0 : (vs == 1 ? xa1 : (vs == 2 ? xa2 : xa3));
        const float S = vs == 0 ? Sa0 : (vs == 1 ? Sa1 : (vs == 2 ? Sa2 : Sa3));
        const float Q = vs == 0 ? Qa0 : (vs == 1 ? Qa1 : (vs == 2 ? Qa2 : Qa3));
        if (wv.x < 0xFFFFFFFEu && wv.x == wv.y && wv.x == wv.z && wv.x == wv.w) {
          const uint32_t L = wv.x;
          if (L != cur_L) {
            flush();
            cur_L = L; cur_sum = 0.f; cur_cnt = 0u;
            cur_c = __ldg(a.C + (size_t)L * D + feat);
          }
          cur_sum += fmaf(128.f, x0, S);
          cur_cnt += 128u;
          const float dl = cur_c - x0;
          inert_f += fmaf(dl, fmaf(128.f, dl, -2.f * S), Q);
          if ((v & 15) == 15) { inert_d += (double)inert_f; inert_f = 0.f; }
        } else if (wv.x != 0xFFFFFFFEu || wv.y != 0xFFFFFFFEu || wv.z != 0xFFFFFFFEu || wv.w != 0xFFFFFFFEu) {
          if (u == 0 && lane == 0) my_list[n_mixed] = (uint32_t)tile_of(v);
          ++n_mixed;
        }
      };
      int s = u;   // ring stage of slab u of tile 0
      uint32_t ph = 0;
      for (unsigned long long it = 0; it < my_pt; ++it) {
        const unsigned char* sl = ring + (size_t)s * 16384;
        tc::mbar_wait_or_trap(&bar_full[s], ph);
        const float x0 = *reinterpret_cast<const float*>(sl + off[0]);
        float S0 = 0.f, S1 = 0.f, Q0 = 0.f, Q1 = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
#pragma unroll
          for (int q = 0; q < 8; q += 2) {
            const float xa = *reinterpret_cast<const float*>(sl + i * 1024 + off[q]);
            const float xb = *reinterpret_cast<const float*>(sl + i * 1024 + off[q + 1]);
            const float da = xa - x0, db = xb - x0;
            S0 += da; S1 += db;
            Q0 = fmaf(da, da, Q0); Q1 = fmaf(db, db, Q1);
          }
        }
        if (it >= 2) consume(it - 2);   // before the slab is released: keeps verdict slot (it + 2) & 3 from being rewritten early
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bar_empty[s]);
        {
          const int is = (int)(it & 3);
          const float St = S0 + S1, Qt = Q0 + Q1;
          if (is == 0) { xa0 = x0; Sa0 = St; Qa0 = Qt; }
          else if (is == 1) { xa1 = x0; Sa1 = St; Qa1 = Qt; }
          else if (is == 2) { xa2 = x0; Sa2 = St; Qa2 = Qt; }
          else { xa3 = x0; Sa3 = St; Qa3 = Qt; }
        }
        s += NS;
        if (s >= kT5sStages) { s -= kT5sStages; ph ^= 1u; }
      }
      for (unsigned long long v = my_pt >= 2 ? my_pt - 2 : 0; v < my_pt; ++v) consume(v);
      flush();
      if (sparse && u == 0) {
        unsigned char* tb = a.part_touched + (size_t)blockIdx.x * a.k;
        for (int b = 0; b < 32; ++b) {
          const int c = lane * 32 + b;
          if (c < a.k) tb[c] = (unsigned char)((seen_bits >> b) & 1u);
        }
      }
      if (u == 0 && lane == 0) {
        a.mixed_count[blockIdx.x] = n_mixed;
        if (a.state != nullptr && n_mixed) atomicAdd(&a.state->mixed_tiles, (unsigned long long)n_mixed);
      }
      inert_d += (double)inert_f;
      for (int o = 16; o > 0; o >>= 1) inert_d += __shfl_xor_sync(0xffffffffu, inert_d, o);
      if (lane == 0) u_red[u] = inert_d;
    }
  } else {
    // =============================== E: group g scores the tiles it = g (mod 2) ===============================
    const int g = warp >> 2, wq = warp & 3;
    const int row = t & 127;
    const uint32_t taddr = tmem + ((uint32_t)(wq * 32) << 16) + (uint32_t)g * 256u;
    for (unsigned long long it = g; it < my_pt; it += 2) {
      const uint32_t u = (uint32_t)(it >> 1);
      const unsigned long long row0 = tile_of(it) * 128ull;
      uint4 prev = make_uint4(0u, 0u, 0u, 0u);
      if (a.chunk_idx > 0 && row0 + (unsigned long long)row < a.n) prev = a.merge[row0 + (unsigned long long)row];
      tc::mbar_wait_spin(&bar_mma[g], u & 1u);   // no suspend-time hint: a parked warp wakes ~400 cycles after the commit, and this wait is on the accumulator's critical path
      if (wq == 0) T5TRACE(it, 8);
      tc::tc_fence_after();
      uint32_t m1, m2;
      auto drained = [&]() {
        tc::tc_fence_before();
        __syncwarp();
        if (lane == 0) tc::mbar_arrive_cluster(&bar_accfree[g], 0u);
      };
      t5_top2<N>(taddr, a.pack_mult, 0u - (a.pack_mult >> 8), m1, m2, drained);
      if (wq == 0) T5TRACE(it, 9);
      T5TRACE(it, 11 + wq);
      const unsigned long long grow = row0 + (unsigned long long)row;
      uint32_t bidx = 0u;
      const int verdict = t5_merge_certify(a, win, grow, m1, m2, prev, bidx);
      uint32_t* flags = flags_all + (g * 2 + (u & 1u)) * 128;
      const bool defer = a.rescan_list != nullptr;
      if (defer) {
        if (verdict) a.labels[grow] = bidx;   // provisional for verdict 2: rows_top2_kernel overwrites it
        my_fallbacks += t5_defer(a, verdict == 2, grow, lane);
      } else if (verdict) {
        if (verdict == 1) a.labels[grow] = bidx;
        else flags[atomicAdd(&n_flag[g][u & 1u], 1u)] = (uint32_t)row;
      }
      if constexpr (FUSE) {
        // one label for the warp's 32 rows (all certified, full tile) -> U adds the tile's column sums to that cluster
        const uint32_t lab0 = __shfl_sync(0xffffffffu, bidx, 0);
        const bool uni = __all_sync(0xffffffffu, verdict == 1 && bidx == lab0) && row0 + 128ull <= a.n;
        if (it >= 4) tc::mbar_wait_or_trap(&bar_vfree[it & 3], (uint32_t)((it >> 2) - 1) & 1u);   // U is done with the slot
        if (lane == 0) warp_verdict[it & 3][wq] = uni ? lab0 : (row0 < a.n ? 0xFFFFFFFFu : 0xFFFFFFFEu);
        __syncwarp();
        if (lane == 0) tc::mbar_arrive(&bar_verdict[it & 3]);
      }
      if (!defer) {
        if (row == 0) n_flag[g][(u + 1u) & 1u] = 0u;
        tc::nb_sync(1 + g, 128);
        const uint32_t nf = n_flag[g][u & 1u];
        if (nf) {
          for (uint32_t f = (uint32_t)wq; f < nf; f += 4) {
            const unsigned long long fr = row0 + flags[f];
            const int fi = t5_exact_label<D>(a.X, a.C, a.k, fr, lane);
            if (lane == 0) a.labels[fr] = (uint32_t)fi;
          }
          my_fallbacks += (row == 0) ? nf : 0u;
        }
      }
      if (wq == 0) T5TRACE(it, 10);
    }
    if (a.state && my_fallbacks) atomicAdd(&a.state->fallback_rows, (unsigned long long)my_fallbacks);
  }
  tc::tc_fence_before();
  __syncthreads();
  if constexpr (FUSE) {
    if (t == 0) {
      double tot = 0.0;
      for (int q = 0; q < NS; ++q) tot += u_red[q];
      a.part_inertia[blockIdx.x] = tot;
    }
  }
  tc::cluster_sync_all();
  if (warp == 8) tc::tmem_dealloc_pair(tmem, 512);
}

}  // namespace lkm

// C ABI (include/lkm.h) and host side of the B200-native K-Means path.
//
// Host logic mirrors fit_lloyd / get_kmeans_result (KM/algorithm.rs:294-367),
// KMeansInit::run (KM/init.rs:83-285) and the ParamGuard checks
// (KM/hyperparams.rs:124-136); every floating-point operation on the data runs
// in the CUDA kernels of lkm_kernels.cuh.  There is no CPU fallback: without a
// CUDA device lkm_create fails with LKM_ERR_CUDA.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#include "../../include/lkm.h"
#include "lkm_kernels.cuh"
#include "lkm_rand.hpp"
#include "lkm_tc.cuh"
#include "lkm_tma.cuh"
#include "lkm_tc4.cuh"
#include "lkm_upd.cuh"
#include "lkm_sortupd.cuh"
#include "lkm_seedpass.cuh"
#include "lkm_tc5.cuh"
#include "lkm_f64s.cuh"
#include "lkm_rows.cuh"
#include "lkm_hamerly.cuh"
#include "lkm_dmma.cuh"
#ifdef LKM_WITH_PROBES
#include "lkm_probe.cuh"
#endif

#define LKM_EXPORT extern "C" __attribute__((visibility("default")))

#include "lkm_ctx.hpp"

using namespace lkm;

namespace lkm {

static NcclApi g_nccl;
#define CKN(call)                                                                                     \
  do {                                                                                                \
    int r__ = (call);                                                                                 \
    if (r__ != 0) {                                                                                   \
      ctx->err = std::string(#call) + ": " + (g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "nccl error"); \
      return LKM_ERR_NCCL;                                                                            \
    }                                                                                                 \
  } while (0)

// Benchmark hygiene: after the scratch buffer has been overwritten (which evicts the resident matrix from L2 but
// leaves the L2 full of DIRTY lines, whose write-back would be charged to the next timed kernel), a read-only sweep
// over the first half of the same buffer replaces them by clean lines.  LKM_FLUSH_SWEEP=0 disables it.
constexpr int kT5MaxChunks = 64;   // CTA-pair assignment kernel: k <= 256 * 64 centroids, one launch per chunk of 256
static int g_flush_sweep = 1;
__global__ void __launch_bounds__(512) l2_read_sweep_kernel(const uint4* __restrict__ p, size_t n16, unsigned int* sink) {
  unsigned int acc = 0;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) {
    const uint4 v = __ldg(p + i);
    acc ^= v.x ^ v.w;
  }
  if (acc == 0x9e3779b9u) *sink = acc;   // never true for a memset pattern; keeps the loads alive
}

// Benchmark hygiene, multi-GPU: the per-step spans of the flushed-L2 mode start behind each rank's own scratch writes, so
// without a rendezvous the skew between the ranks' flushes is charged to the exchange step of every span.  One warp:
// lane r raises this rank's slot in rank r's barrier flags (NVLink peer store) and waits for rank r's slot in its own.
struct RankBarArgs {
  unsigned long long* peer[8];          // rank r's barrier flags + my rank
  const unsigned long long* mine;       // my barrier flags [world]
  unsigned long long epoch;
  int world;
};
__global__ void rank_barrier_kernel(const RankBarArgs a) {
  const int r = threadIdx.x;
  if (r < a.world) {
    __threadfence_system();
    *((volatile unsigned long long*)a.peer[r]) = a.epoch;
    const volatile unsigned long long* f = a.mine + r;
    unsigned int spins = 0;
    while (*f < a.epoch && ++spins < (1u << 24)) {}
  }
}

// ---- launch helpers ----------------------------------------------------------
struct Plan {
  bool ok = false;
  int certified = 0;
  int vec = 0, x_stride = 0, tpr = 1;
  size_t tile_bytes = 0;
  // fused
  bool fused = false;
  int n_subsets = 1, kc = 0;
  size_t smem = 0;
  int grid = 0;
};

static bool has_static_d(uint64_t d) { return d == 2 || d == 3 || d == 4 || d == 8 || d == 16 || d == 32; }

template <typename F> static void tile_layout(uint64_t d, const void* X, Plan& p) {
  const size_t rowbytes = d * sizeof(F);
  p.vec = (rowbytes % 16 == 0) && ((reinterpret_cast<uintptr_t>(X) & 15) == 0);
  size_t stride_bytes;
  if (p.vec) stride_bytes = ((rowbytes / 16) % 2 == 1) ? rowbytes : rowbytes + 16;
  else stride_bytes = (size_t)(d | 1) * sizeof(F);
  p.x_stride = (int)(stride_bytes / sizeof(F));
  // wide rows: several threads share a row so that the tile stays near 64 KB
  p.tpr = 1;
  if (!has_static_d(d))
    while (p.tpr < 32 && (size_t)(kThreads / p.tpr) * stride_bytes > 64 * 1024) p.tpr *= 2;
  p.tile_bytes = std::max<size_t>((size_t)(kThreads / p.tpr) * stride_bytes, 2048);
  p.tile_bytes = (p.tile_bytes + 15) & ~(size_t)15;
}

static size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

static int g_pdl = 1;   // LKM_PDL=0: no programmatic dependent launches (A/B knob)
// launch with programmatic stream serialization (see pdl_wait / pdl_launch_dependents)
template <typename Kern, typename Arg>
static cudaError_t launch_pdl(Kern kern, int grid, int block, size_t smem, cudaStream_t s, const Arg& arg, bool pdl) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = (pdl && g_pdl) ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, arg);
}

template <typename F, int D, int MODE>
static cudaError_t launch_lloyd_t(const LloydArgs<F>& a, int grid, size_t smem, cudaStream_t s) {
  // per launch, not cached: the attribute belongs to the current device's context and one process may drive several devices
  // (lkm_create_multi), each from its own thread
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(lloyd_kernel<F, D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  return launch_pdl(lloyd_kernel<F, D, MODE>, grid, kThreads, smem, s, a, MODE == MODE_FUSED);
}
template <typename F, int D, int MODE> static int occupancy_t(size_t smem) {
  int nb = 0;
  cudaFuncSetAttribute(lloyd_kernel<F, D, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, lloyd_kernel<F, D, MODE>, kThreads, smem);
  return nb;
}
#define LKM_DISPATCH_D(d, CALL)          \
  switch (d) {                           \
    case 2: { constexpr int DD = 2; CALL; } break;   \
    case 3: { constexpr int DD = 3; CALL; } break;   \
    case 4: { constexpr int DD = 4; CALL; } break;   \
    case 8: { constexpr int DD = 8; CALL; } break;   \
    case 16: { constexpr int DD = 16; CALL; } break; \
    case 32: { constexpr int DD = 32; CALL; } break; \
    default: { constexpr int DD = 0; CALL; } break;  \
  }

template <typename F, int MODE>
static cudaError_t launch_lloyd(const LloydArgs<F>& a, int grid, size_t smem, cudaStream_t s) {
  cudaError_t e = cudaSuccess;
  LKM_DISPATCH_D(a.d, (e = launch_lloyd_t<F, DD, MODE>(a, grid, smem, s)));
  return e;
}
template <typename F, int MODE> static int occupancy(int d, size_t smem) {
  int nb = 0;
  LKM_DISPATCH_D(d, (nb = occupancy_t<F, DD, MODE>(smem)));
  return nb;
}

// smem plan for MODE_ASSIGN (assignment only; centroids chunked over k if needed)
template <typename F> static lkm_status plan_assign(lkm_ctx* ctx, const void* X, uint64_t n, uint64_t d, uint64_t k, Plan& p) {
  tile_layout<F>(d, X, p);
  const size_t limit = ctx->smem_optin - 1024;
  if (p.tile_bytes + al16(d * sizeof(F)) > limit) return fail(ctx, LKM_ERR_UNSUPPORTED, "row too wide for the shared-memory tile");
  const size_t c_budget = std::min<size_t>(limit - p.tile_bytes, 96 * 1024);
  uint64_t kc = std::min<uint64_t>(k, c_budget / (d * sizeof(F)));
  if (kc == 0) return fail(ctx, LKM_ERR_UNSUPPORTED, "row too wide for the shared-memory tile");
  p.kc = (int)kc;
  // certified FMA path: compile-time d, every centroid resident, not forced to the exact path
  p.certified = (has_static_d(d) && kc == k && ctx->force_path != 0) ? 1 : 0;
  p.smem = al16(kc * d * sizeof(F)) + p.tile_bytes + (p.certified ? al16((k + 1) * sizeof(F)) : 0);
  const uint64_t n_tiles = (n + kThreads / p.tpr - 1) / (kThreads / p.tpr);
  const int occ = std::max(1, occupancy<F, MODE_ASSIGN>((int)d, p.smem));
  p.grid = (int)std::min<uint64_t>(std::max<uint64_t>(n_tiles, 1), (uint64_t)ctx->sm_count * occ);
  p.ok = true;
  return LKM_OK;
}

// smem plan for the fused iteration (assignment + accumulation in one pass);
// p.fused = false means "does not fit: use assign + chunked update passes".
template <typename F> static lkm_status plan_fused(lkm_ctx* ctx, const void* X, uint64_t n, uint64_t d, uint64_t k, Plan& p) {
  tile_layout<F>(d, X, p);
  const size_t limit = ctx->smem_optin - 1024;
  const size_t kdF = k * d * sizeof(F);
  const size_t fixed = p.tile_bytes + al16(kdF) + al16(k * 4) + kThreads * 4 + 64 + al16((k + 1) * sizeof(F));
  p.fused = fixed + al16(kdF) <= limit;
  if (!p.fused) return LKM_OK;
  int subsets = d <= (uint64_t)kThreads ? (int)(kThreads / d) : 1;
  // keep the accumulator copies within ~48 KB (or one copy if a single one is larger)
  const size_t per = al16(kdF) + al16(k * 4);
  subsets = (int)std::max<size_t>(1, std::min<size_t>(subsets, (48 * 1024) / std::max<size_t>(per, 1)));
  while (subsets > 1 && fixed + (size_t)subsets * per > limit) --subsets;
  p.n_subsets = subsets;
  p.kc = (int)k;
  p.certified = (has_static_d(d) && ctx->force_path != 0) ? 1 : 0;
  p.smem = al16(kdF) + p.tile_bytes + al16((size_t)subsets * kdF) + al16((size_t)subsets * k * 4) + kThreads * 4 +
           (p.certified ? al16((k + 1) * sizeof(F)) : 0);
  const uint64_t n_tiles = (n + kThreads / p.tpr - 1) / (kThreads / p.tpr);
  const int occ = std::max(1, occupancy<F, MODE_FUSED>((int)d, p.smem));
  p.grid = (int)std::min<uint64_t>(std::max<uint64_t>(n_tiles, 1), (uint64_t)ctx->sm_count * occ);
  p.ok = true;
  return LKM_OK;
}

template <typename F>
static lkm_status launch_rows(lkm_ctx* ctx, RowsArgs<F>& a, uint64_t max_rows);

template <typename F>
static lkm_status run_assign(lkm_ctx* ctx, const F* dX, uint64_t n, uint64_t d, const F* dC, uint64_t k, uint32_t* labels,
                             F* dists, int min_combine, const F* weights, F* wdists) {
  if (n == 0) return LKM_OK;
  // wide rows (no compile-time-d instance of lloyd_kernel): thread = row with the row in registers, FMA scores + certificate,
  // exact scan where it fails, winner's distance in the direct form — the same labels and the same bits as the exact scan
  // (rows_top2_kernel over all rows); predict / transform / assign and the seeding passes over several centroids take this
  // (not for the one-centroid passes of k-means++: measured slower there, 197 against 176 ms per 4M x 128 seeding — one CTA
  // of 8 warps per SM is too little occupancy for a pass that only streams)
  if (ctx->rows_assign && ctx->force_path != 0 && weights == nullptr && wdists == nullptr && k >= 8 && n < (1ull << 32) &&
      (d == 64 || (d == 128 && sizeof(F) == 4)) && (reinterpret_cast<uintptr_t>(dX) & 15) == 0) {
    RowsArgs<F> ra{};
    ra.X = dX; ra.n = n; ra.d = (int)d; ra.k = (int)k; ra.C = dC; ra.all_rows = n;
    ra.labels = labels; ra.best_dist = dists; ra.min_combine = min_combine;
    return launch_rows<F>(ctx, ra, n);
  }
  // one centroid, distances only: the per-pick pass of k-means++ at the HBM rate (lkm_seedpass.cuh)
  if constexpr (std::is_same<F, float>::value) {
    if (ctx->seed_pass && ctx->force_path != 0 && k == 1 && labels == nullptr && dists != nullptr && (d == 32 || d == 64 || d == 128) &&
        (reinterpret_cast<uintptr_t>(dX) & 15) == 0 && (reinterpret_cast<uintptr_t>(dC) & 15) == 0) {
      SeedPassArgs sp{};
      sp.X = dX; sp.n = n; sp.c = dC; sp.dists = dists; sp.min_combine = min_combine; sp.weights = weights; sp.wdists = wdists;
      const size_t smem = d == 32 ? seedpass_smem_bytes<32>() : (d == 64 ? seedpass_smem_bytes<64>() : seedpass_smem_bytes<128>());
      const void* fn = d == 32 ? (const void*)seed_pass_kernel<32> : (d == 64 ? (const void*)seed_pass_kernel<64> : (const void*)seed_pass_kernel<128>);
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per launch: see launch_lloyd_t
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
      const int occ = std::max(1, (int)((ctx->smem_optin + 1024) / (smem + 1024)));
      const uint64_t tiles = (n + kSpRows - 1) / kSpRows;
      const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)ctx->sm_count * std::min(occ, 4)));
      if (d == 32) seed_pass_kernel<32><<<grid, kSpRows, smem, ctx->stream>>>(sp);
      else if (d == 64) seed_pass_kernel<64><<<grid, kSpRows, smem, ctx->stream>>>(sp);
      else seed_pass_kernel<128><<<grid, kSpRows, smem, ctx->stream>>>(sp);
      CK(cudaGetLastError());
      ctx->launches++;
      return LKM_OK;
    }
  }
  Plan p;
  CKS(plan_assign<F>(ctx, dX, n, d, k, p));
  LloydArgs<F> a{};
  a.X = dX; a.n = n; a.d = (int)d; a.k = (int)k; a.C = dC;
  a.labels = labels; a.dists = dists; a.min_combine = min_combine; a.weights = weights; a.wdists = wdists;
  a.n_subsets = 1; a.x_stride = p.x_stride; a.kc = p.kc; a.chunk_base = 0; a.chunk_rows = (int)k; a.vec = p.vec;
  a.tpr = p.tpr; a.certified = p.certified;
  CK((launch_lloyd<F, MODE_ASSIGN>(a, p.grid, p.smem, ctx->stream)));
  ctx->launches++;
  return LKM_OK;
}

template <typename F, typename S, typename CT, bool UPDATE>
static cudaError_t launch_finalize(const FinalizeArgs<F, S, CT>& a, cudaStream_t s) {
  const int grid = (a.k * a.d + kFinElems - 1) / kFinElems;
  return launch_pdl(finalize_kernel<F, S, CT, UPDATE>, grid, kFinThreads, 0, s, a, true);
}

static lkm_status comm_allgather_host(lkm_ctx* ctx, const void* src, size_t bytes, std::vector<unsigned char>& out);

// (re)creates the IPC-shared gather buffers for partial vectors of up to L doubles.  Collective:
// every rank calls it with the same L.  On any failure the run falls back to ncclAllGather.
static lkm_status p2p_setup(lkm_ctx* ctx, size_t L) {
  if (!ctx->p2p_enabled || ctx->world > 8) return LKM_OK;
  if (ctx->p2p_local && L <= ctx->p2p_cap) return LKM_OK;
  const int world = ctx->world;
  const size_t cap = std::max<size_t>(L, 4096);
  // [2 parities][world][cap] partials | [2][world][kFxMaxCtas] exchange flags | [world] barrier flags | (16-byte aligned)
  // [2][world][cap] 16-byte self-validating slots of finalize_cl_kernel
  const size_t ll_off = (((2 * (size_t)world * cap + 2 * (size_t)world * kFxMaxCtas + (size_t)world) * 8) + 15) & ~(size_t)15;
  const size_t bytes = ll_off + 2 * (size_t)world * cap * 16;
  ctx->p2p_ll_off = ll_off;
  if (ctx->group != nullptr) {
    // one process: every rank allocates its buffer, the pointers and device numbers travel through the group, and the
    // peers are reached by plain peer access (no IPC handles)
    {
      std::vector<unsigned char> all0;
      int dummy = 0;
      CKS(comm_allgather_host(ctx, &dummy, sizeof(int), all0));   // nobody still uses the old buffers
    }
    if (ctx->p2p_local) { cudaFree(ctx->p2p_local); ctx->p2p_local = nullptr; }
    ctx->p2p_peer.clear();
    struct Rec { void* p; int device; int ok; } rec{nullptr, ctx->device, 1};
    if (cudaMalloc(&ctx->p2p_local, bytes) != cudaSuccess) { rec.ok = 0; ctx->p2p_local = nullptr; cudaGetLastError(); }
    else { cudaMemset(ctx->p2p_local, 0, bytes); cudaDeviceSynchronize(); }   // zeroed before any peer learns the address (tags and flags start at 0)
    rec.p = ctx->p2p_local;
    std::vector<unsigned char> all;
    CKS(comm_allgather_host(ctx, &rec, sizeof(rec), all));
    int mine_ok = rec.ok;
    ctx->p2p_peer.assign(world, nullptr);
    for (int r = 0; r < world; ++r) {
      Rec q;
      std::memcpy(&q, all.data() + (size_t)r * sizeof(Rec), sizeof(Rec));
      if (!q.ok) mine_ok = 0;
      ctx->p2p_peer[r] = q.p;
      if (r != ctx->rank && q.device != ctx->device) {
        int can = 0;
        if (cudaDeviceCanAccessPeer(&can, ctx->device, q.device) != cudaSuccess || !can) mine_ok = 0;
        else {
          const cudaError_t e = cudaDeviceEnablePeerAccess(q.device, 0);
          if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) mine_ok = 0;
          cudaGetLastError();
        }
      }
    }
    CKS(comm_allgather_host(ctx, &mine_ok, sizeof(int), all));
    for (int r = 0; r < world; ++r) { int v; std::memcpy(&v, all.data() + (size_t)r * 4, 4); if (!v) return fail(ctx, LKM_ERR_UNSUPPORTED, "the devices of the group cannot reach each other over NVLink / PCIe peer access"); }
    ctx->p2p_cap = cap;
    CK(cudaStreamSynchronize(ctx->stream));
    return LKM_OK;
  }
  // release the old mapping
  for (int r = 0; r < (int)ctx->p2p_peer.size(); ++r)
    if (r != ctx->rank && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
  ctx->p2p_peer.clear();
  if (ctx->p2p_local) { cudaFree(ctx->p2p_local); ctx->p2p_local = nullptr; }
  int ok = 1;
  cudaIpcMemHandle_t mine;
  std::memset(&mine, 0, sizeof(mine));
  if (cudaMalloc(&ctx->p2p_local, bytes) != cudaSuccess) { ok = 0; ctx->p2p_local = nullptr; cudaGetLastError(); }
  if (ok) {
    cudaMemset(ctx->p2p_local, 0, bytes);
    cudaDeviceSynchronize();   // zeroed before any peer learns the handle (tags and flags start at 0)
    if (cudaIpcGetMemHandle(&mine, ctx->p2p_local) != cudaSuccess) { ok = 0; cudaGetLastError(); }
  }
  std::vector<unsigned char> all;
  struct Rec { cudaIpcMemHandle_t h; int ok; int pad; } rec{mine, ok, 0};
  CKS(comm_allgather_host(ctx, &rec, sizeof(rec), all));
  bool all_ok = true;
  for (int r = 0; r < world; ++r) { Rec q; std::memcpy(&q, all.data() + (size_t)r * sizeof(Rec), sizeof(Rec)); all_ok = all_ok && q.ok; }
  int open_ok = 1;
  if (all_ok) {
    ctx->p2p_peer.assign(world, nullptr);
    for (int r = 0; r < world; ++r) {
      if (r == ctx->rank) { ctx->p2p_peer[r] = ctx->p2p_local; continue; }
      Rec q;
      std::memcpy(&q, all.data() + (size_t)r * sizeof(Rec), sizeof(Rec));
      if (cudaIpcOpenMemHandle(&ctx->p2p_peer[r], q.h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { open_ok = 0; cudaGetLastError(); ctx->p2p_peer[r] = nullptr; }
    }
  }
  // agree on the outcome so that every rank takes the same path
  int mine_ok = (all_ok && open_ok) ? 1 : 0;
  CKS(comm_allgather_host(ctx, &mine_ok, sizeof(int), all));
  bool everyone = true;
  for (int r = 0; r < world; ++r) { int v; std::memcpy(&v, all.data() + (size_t)r * 4, 4); everyone = everyone && v; }
  if (!everyone) {
    for (int r = 0; r < (int)ctx->p2p_peer.size(); ++r)
      if (r != ctx->rank && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
    ctx->p2p_peer.clear();
    if (ctx->p2p_local) { cudaFree(ctx->p2p_local); ctx->p2p_local = nullptr; }
    ctx->p2p_enabled = 0;
    return LKM_OK;
  }
  ctx->p2p_cap = cap;
  CK(cudaStreamSynchronize(ctx->stream));
  return LKM_OK;
}

// ---- TMA descriptor of the resident f32 matrix (d = 32): 128-row x 128-byte boxes, SWIZZLE_128B ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static lkm_status make_row_tile_map(lkm_ctx* ctx, const float* dX, uint64_t n, CUtensorMap& map, uint64_t d = 32) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(ctx, LKM_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)n};
  const cuuint64_t strides[1] = {(cuuint64_t)d * 4};
  const cuuint32_t box[2] = {32, (cuuint32_t)kTcRows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(dX), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, LKM_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return LKM_OK;
}

// TMA descriptor of the resident f32 matrix for any row width: box_cols x box_rows boxes, no swizzle
static lkm_status make_plain_tile_map(lkm_ctx* ctx, const float* dX, uint64_t n, uint64_t d, uint32_t box_cols, uint32_t box_rows,
                                      CUtensorMap& map) {
  static EncodeTiledFn encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    if (!fn || qres != cudaDriverEntryPointSuccess) return fail(ctx, LKM_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    encode = reinterpret_cast<EncodeTiledFn>(fn);
  }
  const cuuint64_t dims[2] = {(cuuint64_t)d, (cuuint64_t)n};
  const cuuint64_t strides[1] = {(cuuint64_t)d * 4};
  const cuuint32_t box[2] = {box_cols, box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(dX), dims, strides, box, estr,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(ctx, LKM_ERR_CUDA, "cuTensorMapEncodeTiled failed");
  return LKM_OK;
}

static int g_coop = 1;   // LKM_COOP=0: the in-kernel finalize relies on grid <= SM count alone (no cooperative-launch attribute)
static cudaError_t launch_tc4(int kp, int passes, int grid, size_t smem, cudaStream_t s, const Tc4Args& arg, const CUtensorMap& map) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3((unsigned)kT4Threads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  int na = 0;
  if (arg.fin_sync != nullptr && g_coop) {
    // the kernel's tail waits for every CTA of the grid: a cooperative launch guarantees that they are co-resident even
    // when other streams compete for the SMs
    at[na].id = cudaLaunchAttributeCooperative;
    at[na].val.cooperative = 1;
    ++na;
  } else if (g_pdl) {
    at[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = at;
  cfg.numAttrs = na;
  if (passes == 2) {
    if (kp == 32) return cudaLaunchKernelEx(&cfg, lloyd_tc4_kernel<32, 2>, arg, map);
    return cudaLaunchKernelEx(&cfg, lloyd_tc4_kernel<64, 2>, arg, map);
  }
  if (passes == 1) {
    if (kp == 32) return cudaLaunchKernelEx(&cfg, lloyd_tc4_kernel<32, 1>, arg, map);
    return cudaLaunchKernelEx(&cfg, lloyd_tc4_kernel<64, 1>, arg, map);
  }
  if (kp == 32) return cudaLaunchKernelEx(&cfg, lloyd_tc4_kernel<32, 3>, arg, map);
  return cudaLaunchKernelEx(&cfg, lloyd_tc4_kernel<64, 3>, arg, map);
}

// ---- rows_top2_kernel (lkm_rows.cuh): the second stage of the screened paths and Hamerly's two_closest_centroids ----
template <typename F, int D>
static cudaError_t launch_rows_t(const RowsArgs<F>& a, int grid, size_t smem, cudaStream_t s) {
  if (smem > 48 * 1024) {   // per launch: see launch_lloyd_t
    cudaError_t e = cudaFuncSetAttribute(rows_top2_kernel<F, D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  rows_top2_kernel<F, D><<<grid, kRowsThreads, smem, s>>>(a);
  return cudaGetLastError();
}
// max_rows bounds the grid (the kernel reads the real count on the device when a.list_count is set)
template <typename F>
static lkm_status launch_rows(lkm_ctx* ctx, RowsArgs<F>& a, uint64_t max_rows) {
  const uint64_t d = (uint64_t)a.d, k = (uint64_t)a.k;
  const size_t per = (d + 1) * sizeof(F);
  const size_t limit = ctx->smem_optin - 2048;
  // every centroid resident when that fits, else chunks of at most 96 KB (two CTAs per SM)
  uint64_t kc = k;
  if (k * per > limit) kc = std::max<uint64_t>(1, std::min<uint64_t>(k, (96 * 1024) / per));
  if (kc * per > limit) return fail(ctx, LKM_ERR_UNSUPPORTED, "row too wide for the shared-memory centroid chunk");
  a.kc = (int)kc;
  const size_t smem = al16(kc * per);
  const int occ = smem > 100 * 1024 ? 1 : 2;
  const uint64_t tiles = (max_rows + kRowsThreads - 1) / kRowsThreads;
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)ctx->sm_count * occ));
  const bool aligned = (reinterpret_cast<uintptr_t>(a.X) & 15) == 0;
  cudaError_t e = cudaSuccess;
#define LKM_ROWS_CASE(DD) case DD: e = launch_rows_t<F, DD>(a, grid, smem, ctx->stream); break;
  if (!aligned) e = launch_rows_t<F, 0>(a, grid, smem, ctx->stream);
  else if constexpr (std::is_same<F, float>::value) {
    switch (d) {
      LKM_ROWS_CASE(2) LKM_ROWS_CASE(3) LKM_ROWS_CASE(4) LKM_ROWS_CASE(8) LKM_ROWS_CASE(16) LKM_ROWS_CASE(32) LKM_ROWS_CASE(64) LKM_ROWS_CASE(128)
      default: e = launch_rows_t<F, 0>(a, grid, smem, ctx->stream); break;
    }
  } else {
    switch (d) {
      LKM_ROWS_CASE(2) LKM_ROWS_CASE(3) LKM_ROWS_CASE(4) LKM_ROWS_CASE(8) LKM_ROWS_CASE(16) LKM_ROWS_CASE(32) LKM_ROWS_CASE(64)
      default: e = launch_rows_t<F, 0>(a, grid, smem, ctx->stream); break;
    }
  }
#undef LKM_ROWS_CASE
  CK(e);
  ctx->launches++;
  return LKM_OK;
}

static void trace_mark(lkm_ctx* ctx, const char* name) {
  if (!ctx->step_trace) return;
  cudaEvent_t e;
  if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, ctx->stream);
  ctx->trace_ev.push_back(e);
  ctx->trace_name.push_back(name);
}
static void trace_report(lkm_ctx* ctx) {
  if (!ctx->step_trace || ctx->trace_ev.size() < 2) return;
  cudaEventSynchronize(ctx->trace_ev.back());
  std::vector<std::pair<std::string, std::pair<double, int>>> agg;
  for (size_t i = 1; i < ctx->trace_ev.size(); ++i) {
    float ms = 0;
    cudaEventElapsedTime(&ms, ctx->trace_ev[i - 1], ctx->trace_ev[i]);
    const std::string nm = ctx->trace_name[i];
    size_t q = 0;
    for (; q < agg.size(); ++q) if (agg[q].first == nm) break;
    if (q == agg.size()) agg.push_back({nm, {0.0, 0}});
    agg[q].second.first += ms;
    agg[q].second.second += 1;
  }
  for (auto& a : agg) std::fprintf(stderr, "step_trace %-22s n=%4d mean %9.2f us\n", a.first.c_str(), a.second.second, 1e3 * a.second.first / a.second.second);
  for (cudaEvent_t e : ctx->trace_ev) cudaEventDestroy(e);
  ctx->trace_ev.clear();
  ctx->trace_name.clear();
}

struct LloydOut {
  uint64_t n_iter = 0, fallback_rows = 0;
  double inertia = 0, shift = 0;
  float ms = 0, main_ms = 0;
  int final_buf = 0;
};

// The loop of fit_lloyd (KM/algorithm.rs:314-331) on the device.  dC2 is a
// double buffer [2][k*d]; iteration i reads buffer i&1 and writes (i+1)&1.
// Kernels become no-ops once the device-side `done` flag is set; the host
// polls the flag once per batch of iterations.
template <typename F>
static lkm_status run_lloyd(lkm_ctx* ctx, F* dC2, uint64_t k, uint64_t max_iter, double tol, LloydOut& out) {
  const uint64_t n = ctx->n, d = ctx->d, kd = k * d;
  const F* dX = reinterpret_cast<const F*>(ctx->X);
  if (kd > (uint64_t)std::numeric_limits<int>::max() / 4) return fail(ctx, LKM_ERR_UNSUPPORTED, "k*d too large");
  // tensor-core path: f32, d = 32, k <= 256, everything resident in shared memory
  bool use_tc = false, pair32 = false;
  int tc_grid = 0, tc_kp = 0, tc_cols = 0, tc_threads = 128;
  size_t tc_smem = 0;
  int tc4_nstages = 0, tc4_nstages1 = 0;   // ring depth of the 2- / 3-product variants and of the one-product variant
  size_t tc_smem1 = 0;
  CUtensorMap tc4_map;
  std::memset(&tc4_map, 0, sizeof(tc4_map));
  if constexpr (std::is_same<F, float>::value) {
    pair32 = ctx->pair32 && ctx->tc5_enabled && ctx->upd_variant != 0 && (ctx->force_path == -1 || ctx->force_path == 2) && d == 32 &&
             k <= 64 && kd >= 512 && n >= 1 && (reinterpret_cast<uintptr_t>(dX) & 15) == 0;
    if (!pair32 && (ctx->force_path == -1 || ctx->force_path == 2) && d == 32 && k <= 256 && n >= 1 &&
        (reinterpret_cast<uintptr_t>(dX) & 15) == 0) {
      tc_kp = (int)((k + 31) / 32 * 32);
      tc_cols = 32;
      while (tc_cols < tc_kp) tc_cols *= 2;
      const bool v4 = ctx->tc_variant == 4 && k <= 64 && n + kTcRows < (1ull << 31) &&
                      tc4_smem_bytes((int)k, tc_kp, 3) <= ctx->smem_optin - 1024;
      if (v4) {
        tc4_nstages = tc4_stages((int)k, tc_kp, ctx->smem_optin - 1024);
        tc4_nstages1 = tc4_stages((int)k, tc_kp, ctx->smem_optin - 1024, 1);
        if (const char* ev = std::getenv("LKM_TC4_STAGES")) {   // tuning aid
          tc4_nstages = std::max(3, std::min(tc4_nstages, std::atoi(ev)));
          tc4_nstages1 = std::max(3, std::min(tc4_nstages1, std::atoi(ev)));
        }
        tc_smem1 = tc4_smem_bytes((int)k, tc_kp, tc4_nstages1, 1);
        tc_smem = tc4_smem_bytes((int)k, tc_kp, tc4_nstages);
        const void* fn = tc_kp == 32 ? (const void*)lloyd_tc4_kernel<32, 3> : (const void*)lloyd_tc4_kernel<64, 3>;
        const void* fn2 = tc_kp == 32 ? (const void*)lloyd_tc4_kernel<32, 2> : (const void*)lloyd_tc4_kernel<64, 2>;
        CK(cudaFuncSetAttribute(fn2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem));
        CK(cudaFuncSetAttribute(fn2, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        const void* fn1 = tc_kp == 32 ? (const void*)lloyd_tc4_kernel<32, 1> : (const void*)lloyd_tc4_kernel<64, 1>;
        CK(cudaFuncSetAttribute(fn1, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem1));
        CK(cudaFuncSetAttribute(fn1, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem));
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        CKS(make_row_tile_map(ctx, dX, n, tc4_map));
        // max row norm of the resident matrix (cached for owned data; external pointers are rescanned)
        if (!ctx->own_X || ctx->xmax2_epoch != ctx->data_epoch) {
          CK(ctx->xmax2.ensure(16));
          CK(cudaMemsetAsync(ctx->xmax2.p, 0, 4, ctx->stream));
          row_norm2_max_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(dX, n, ctx->xmax2.as<unsigned int>());
          CK(cudaGetLastError());
          ctx->launches++;
          ctx->xmax2_epoch = ctx->data_epoch;
        }
        const uint64_t n_tiles = (n + kTcRows - 1) / kTcRows;
        tc_grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count);
        use_tc = true;
        tc_threads = kT4Threads;
      } else {
        // 64 < k <= 256: tcgen05 3xTF32 from shared-memory operands, 128 threads, two CTAs per SM
        tc_smem = tc_smem_bytes((int)k, tc_kp, 1);
        if (tc_smem <= ctx->smem_optin) {
          CK(cudaFuncSetAttribute(lloyd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc_smem));
          CK(cudaFuncSetAttribute(lloyd_tc_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
          // with the carve-out at its maximum, 228 KB per SM hold floor(228K / (dynamic + static + 1K reserved)) CTAs
          int occ = (int)((228 * 1024) / (tc_smem + 2048));
          occ = std::min(occ, 2);                // register file: two CTAs per SM
          occ = std::min(occ, 512 / tc_cols);    // TMEM columns per SM
          if (occ >= 1) {
            const uint64_t n_tiles = (n + kTcRows - 1) / kTcRows;
            tc_grid = (int)std::min<uint64_t>(n_tiles, (uint64_t)ctx->sm_count * occ);
            use_tc = true;
            tc_threads = 128;
          }
        }
      }
    }
    if (ctx->force_path == 2 && !use_tc && !pair32) return fail(ctx, LKM_ERR_UNSUPPORTED, "tcgen05 path needs f32, d = 32, k <= 256");
  } else {
    if (ctx->force_path == 2) return fail(ctx, LKM_ERR_UNSUPPORTED, "tcgen05 path needs f32, d = 32, k <= 256");
  }
  Plan pf;
  CKS(plan_fused<F>(ctx, dX, n, d, k, pf));
  if (use_tc) { pf.fused = true; pf.grid = tc_grid; }
  if constexpr (std::is_same<F, float>::value) {
    // shapes of the CTA-pair assignment kernel: two-pass (tensor-core assignment + streaming update) even when the
    // centroids would fit the fused CUDA-core kernel, unless the contraction is tiny
    if (!use_tc && (ctx->force_path == -1 || ctx->force_path == 2) && ctx->tc5_enabled && ctx->upd_variant != 0 &&
        (d == 64 || d == 128) && k <= (uint64_t)kT5K * kT5MaxChunks && kd >= 2048 && n >= 1 && (reinterpret_cast<uintptr_t>(dX) & 15) == 0)
      pf.fused = false;
    if (pair32) pf.fused = false;
  }
  ctx->last_path = use_tc ? 2 : ((pf.fused && pf.certified) ? 1 : 0);
  // f64, 2 / 4 / 8 / 16 features, few centroids: f32-screened warp-per-tile kernel (lkm_f64s.cuh)
  bool use_f6 = false;
  size_t f6_smem = 0;
  const void* f6_fn = nullptr;
  if constexpr (std::is_same<F, double>::value) {
    const bool shape_ok = (d == 2 || d == 4 || d == 8 || d == 16) && k <= 64 && kd <= 512 && n >= 1 &&
                          (reinterpret_cast<uintptr_t>(dX) & 15) == 0;
    if (shape_ok && ((ctx->force_path == -1 && ctx->f64s_enabled && n >= 32768) || ctx->force_path == 3)) {
      const void* fn = nullptr;
      const bool pk = ctx->f64s_packed != 0;
#define LKM_F6_PICK(DD, RR) { f6_smem = f6::smem_bytes<DD, RR>((int)k); fn = pk ? (const void*)f6::lloyd_f64s_kernel<DD, true, RR> : (const void*)f6::lloyd_f64s_kernel<DD, false, RR>; }
      if (ctx->f64s_rpt == 1) {
        switch (d) { case 2: LKM_F6_PICK(2, 1) break; case 4: LKM_F6_PICK(4, 1) break; case 8: LKM_F6_PICK(8, 1) break; default: LKM_F6_PICK(16, 1) break; }
      } else {
        switch (d) { case 2: LKM_F6_PICK(2, 2) break; case 4: LKM_F6_PICK(4, 2) break; case 8: LKM_F6_PICK(8, 2) break; default: LKM_F6_PICK(16, 2) break; }
      }
#undef LKM_F6_PICK
      f6_fn = fn;
      if (f6_smem <= ctx->smem_optin) {
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)f6_smem));
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        int occ = 0;
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, f6::kThreads, f6_smem));
        occ = std::max(1, occ);
        const uint64_t n_tiles = (n + 32 * ctx->f64s_rpt - 1) / (32 * ctx->f64s_rpt);
        const uint64_t want = (n_tiles + f6::kWarps - 1) / f6::kWarps;
        pf.fused = true;
        pf.grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(want, (uint64_t)ctx->sm_count * occ));
        use_f6 = true;
        ctx->last_path = 3;
      }
    }
    if (ctx->force_path == 3 && !use_f6) return fail(ctx, LKM_ERR_UNSUPPORTED, "f32-screened f64 path needs d in {2,4,8,16}, k <= 64, k*d <= 512");
  } else {
    if (ctx->force_path == 3) return fail(ctx, LKM_ERR_UNSUPPORTED, "f32-screened f64 path needs f64 data");
  }
  // f64 with a large contraction: scores on the FP64 tensor pipe (lkm_dmma.cuh), two-pass
  bool use_dm = false;
  if constexpr (std::is_same<F, double>::value) {
    const bool shape_ok = d % 4 == 0 && d >= 8 && d <= 128 && k >= 8 && n >= 1 && (reinterpret_cast<uintptr_t>(dX) & 15) == 0 &&
                          n + 256 < (1ull << 32) && dmma_smem_bytes((int)d) <= ctx->smem_optin;
    // (measured: 2M x 64, k = 256: 6.2 ms against 36.7 ms on the CUDA cores; 1M x 128: 6.5 against 66.8; at d = 16 the fused
    // certified FMA kernel is as fast, 1.95 against 2.11 ms per 4M rows, and keeps the single pass)
    const bool fused_small_d = pf.fused && pf.certified && d <= 16;
    if (!use_f6 && shape_ok && ((ctx->force_path == -1 && ctx->dmma_enabled && kd >= 1024 && n >= 8192 && !fused_small_d) || ctx->force_path == 4)) {
      use_dm = true;
      pf.fused = false;
      ctx->last_path = 4;
    }
    if (ctx->force_path == 4 && !use_dm) return fail(ctx, LKM_ERR_UNSUPPORTED, "DMMA path needs f64 data, d a multiple of 4 in [8, 128], k >= 8");
  } else {
    if (ctx->force_path == 4) return fail(ctx, LKM_ERR_UNSUPPORTED, "DMMA path needs f64 data");
  }
  Plan pa, pu;
  int n_chunks = 1, chunk_rows = (int)k;
  if (!pf.fused) {
    CKS(plan_assign<F>(ctx, dX, n, d, k, pa));
    // update pass: tile + one accumulator copy for a cluster chunk
    pu = pa;
    const size_t limit = ctx->smem_optin - 1024;
    const size_t avail = limit - pu.tile_bytes - kThreads * 4 - 64;
    uint64_t cr = std::min<uint64_t>(k, avail / (d * sizeof(F) + 4 + 1));
    if (cr == 0) return fail(ctx, LKM_ERR_UNSUPPORTED, "row too wide for the update pass");
    chunk_rows = (int)cr;
    n_chunks = (int)((k + cr - 1) / cr);
    pu.n_subsets = 1;
    pu.smem = pu.tile_bytes + al16(cr * d * sizeof(F)) + al16(cr * 4) + kThreads * 4;
    const uint64_t n_tiles = (n + kThreads / pu.tpr - 1) / (kThreads / pu.tpr);
    const int occ = std::max(1, occupancy<F, MODE_UPDATE>((int)d, pu.smem));
    pu.grid = (int)std::min<uint64_t>(std::max<uint64_t>(n_tiles, 1), (uint64_t)ctx->sm_count * occ);
    CK(ctx->labels.ensure(std::max<uint64_t>(n, 1) * 4));
    CK(ctx->dists.ensure(std::max<uint64_t>(n, 1) * sizeof(F)));
  }
  // two-pass shapes: tensor-core assignment when f32 and d is 32, 64 or 128
  bool use_tca = false;
  int tca_ka = 0, tca_chunks = 0, tca_grid = 0;
  size_t tca_smem = 0;
  if constexpr (std::is_same<F, float>::value) {
    if (!pf.fused && (ctx->force_path == -1 || ctx->force_path == 2) && (d == 32 || d == 64 || d == 128) &&
        (reinterpret_cast<uintptr_t>(dX) & 15) == 0) {
      tca_ka = (int)(d / 32);
      tca_chunks = (int)((k + kAsgChunk - 1) / kAsgChunk);
      tca_smem = tc_assign_smem_bytes(tca_ka);
      if (tca_smem <= ctx->smem_optin) {
        CK(cudaFuncSetAttribute(tc_assign_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tca_smem));
        CK(cudaFuncSetAttribute(tc_assign_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        int occ = (int)((228 * 1024) / (tca_smem + 2048));
        occ = std::max(1, std::min(occ, 3));
        const uint64_t n_tiles = (n + 127) / 128;
        tca_grid = (int)std::min<uint64_t>(std::max<uint64_t>(n_tiles, 1), (uint64_t)ctx->sm_count * occ);
        CK(ctx->tca_bsplit.ensure((size_t)tca_chunks * tca_ka * 16384));
        CK(ctx->tca_cn.ensure((size_t)tca_chunks * kAsgChunk * 4));
        CK(ctx->tca_cmax.ensure(16));
        use_tca = true;
        ctx->last_path = 2;
      }
    }
  }
  // streaming update pass (lkm_upd.cuh) behind the tensor-core assignment: the accumulators of a cluster chunk stay in
  // shared memory, blockIdx.y walks the chunks (X is re-read once per chunk)
  bool use_upd = false;
  int upd_ny = 1, upd_chunk = (int)k, upd_stages = 0, upd_gx = 0;
  size_t upd_smem = 0;
  CUtensorMap upd_map;
  std::memset(&upd_map, 0, sizeof(upd_map));
  if constexpr (std::is_same<F, float>::value) {
    if (use_tca && ctx->upd_variant != 0 && n + 64 < (1ull << 31)) {
      const size_t limit = ctx->smem_optin - 1024;
      int ny = 1;
      for (; ny <= ctx->sm_count; ++ny)
        if (upd_smem_bytes((int)d, (int)((k + ny - 1) / ny), 3) <= limit) break;
      if (ny <= ctx->sm_count) {
        upd_ny = ny;
        upd_chunk = (int)((k + ny - 1) / ny);
        upd_ny = (int)((k + upd_chunk - 1) / upd_chunk);
        upd_stages = 3;
        while (upd_stages < kUpdMaxStages && upd_smem_bytes((int)d, upd_chunk, upd_stages + 1) <= limit) ++upd_stages;
        upd_smem = upd_smem_bytes((int)d, upd_chunk, upd_stages);
        const uint64_t n_sub = (n + upd_stage_rows((int)d) - 1) / upd_stage_rows((int)d);
        upd_gx = (int)std::max<uint64_t>(1, std::min<uint64_t>(n_sub, (uint64_t)(ctx->sm_count / upd_ny)));
        const void* fn = d == 32 ? (const void*)update_rows_kernel<32> : (d == 64 ? (const void*)update_rows_kernel<64> : (const void*)update_rows_kernel<128>);
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)upd_smem));
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        CKS(make_plain_tile_map(ctx, dX, n, d, (uint32_t)d, (uint32_t)upd_stage_rows((int)d), upd_map));
        CK(ctx->labels.ensure((n + 256) * 4));
        use_upd = true;
      }
    }
  }
  // CTA-pair assignment kernel (lkm_tc5.cuh): all centroid operand tiles resident across the two CTAs of a cluster
  bool use_tc5 = false;
  int tc5_grid = 0, tc5_chunks = 1, tc5_list_cap = 0, tc5_chunk_n = kT5K;
  bool tc5_can_fuse = false;
  size_t tc5_smem = 0;
  CUtensorMap tc5_map;
  std::memset(&tc5_map, 0, sizeof(tc5_map));
  bool piv_ok = false;
  if constexpr (std::is_same<F, float>::value) {
    if (use_upd && ctx->tc5_enabled && (d == 64 || d == 128 || pair32) && k <= (uint64_t)kT5K * kT5MaxChunks && n + 512 < (1ull << 31) &&
        tc5_smem_bytes((int)d) <= ctx->smem_optin && tc5s_smem_bytes((int)d) <= ctx->smem_optin) {
      tc5_smem = tc5_smem_bytes((int)d);
      const void* fn = d == 64 ? (const void*)assign_tc5_kernel<64> : (const void*)assign_tc5_kernel<128>;
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc5_smem));
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
      const void* fns = d == 64 ? (const void*)assign_tc5s_kernel<64, false, 256> : (const void*)assign_tc5s_kernel<128, false, 256>;
      CK(cudaFuncSetAttribute(fns, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc5s_smem_bytes((int)d)));
      CK(cudaFuncSetAttribute(fns, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
      if (pair32) {
        CK(cudaFuncSetAttribute(assign_tc5s_kernel<32, true, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc5s_smem_bytes(32)));
        CK(cudaFuncSetAttribute(assign_tc5s_kernel<32, true, 64>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        CK(cudaFuncSetAttribute(assign_tc5s_kernel<32, false, 64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc5s_smem_bytes(32)));
        CK(cudaFuncSetAttribute(assign_tc5s_kernel<32, false, 64>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
      }
      const void* fnf = d == 64 ? (const void*)assign_tc5s_kernel<64, true, 256> : (const void*)assign_tc5s_kernel<128, true, 256>;
      CK(cudaFuncSetAttribute(fnf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tc5s_smem_bytes((int)d)));
      CK(cudaFuncSetAttribute(fnf, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
      CKS(make_row_tile_map(ctx, dX, n, tc5_map, d));
      if (!ctx->own_X || ctx->xmax2_epoch != ctx->data_epoch) {
        CK(ctx->xmax2.ensure(16));
        CK(cudaMemsetAsync(ctx->xmax2.p, 0, 4, ctx->stream));
        row_norm2_max_wide_kernel<<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(dX, n, (int)d, ctx->xmax2.as<unsigned int>());
        CK(cudaGetLastError());
        ctx->launches++;
        ctx->xmax2_epoch = ctx->data_epoch;
      }
      const uint64_t n_pt = ((n + 127) / 128 + 1) / 2;
      tc5_grid = 2 * (int)std::max<uint64_t>(1, std::min<uint64_t>(n_pt, (uint64_t)(ctx->sm_count / 2)));
      tc5_chunk_n = pair32 ? 64 : kT5K;
      tc5_chunks = (int)((k + tc5_chunk_n - 1) / tc5_chunk_n);
      if (tc5_chunks > 1) CK(ctx->t5merge.ensure((n + 256) * 16));
      if (ctx->tc5_defer) CK(ctx->ham_list.ensure((n + 256) * 4 + 16));   // rows for rows_top2_kernel; the count sits behind them
      // fused update of uniform tiles inside the screening kernel
      tc5_can_fuse = ctx->tc5_fuse != 0;   // in the launch of the last chunk, where the merged verdicts are known
      if (tc5_can_fuse) {
        tc5_list_cap = (int)((n_pt + (uint64_t)(tc5_grid / 2) - 1) / (uint64_t)(tc5_grid / 2)) + 1;
        CK(ctx->t5list.ensure(((size_t)tc5_grid * tc5_list_cap + (size_t)tc5_grid + 64) * 4));
      }
      use_tc5 = true;
      // pivot pruning of the rows the screening kernel lists (lkm_rows.cuh)
      if (ctx->tc5_defer && ctx->rows_pivot != 0 && d != 32 && k <= 4096) {
        CK(ctx->piv_cc.ensure(k * k * 4));
        CK(ctx->piv_list2.ensure((n + 256) * 4 + 64));
        piv_ok = true;
      }
    }
  }
  const int grid = pf.fused ? pf.grid : (use_upd ? (tc5_can_fuse ? std::max(upd_gx, 2 * tc5_grid) : upd_gx) : pu.grid);
  if (tc5_can_fuse) CK(ctx->part_inertia.ensure((size_t)std::max(1024, tc5_grid * (1 + upd_ny)) * 8));
  CK(ctx->part_sums.ensure((size_t)grid * kd * sizeof(F)));
  CK(ctx->part_counts.ensure((size_t)grid * k * 4));
  if (use_upd && ctx->sparse_partials) CK(ctx->part_touched.ensure((size_t)grid * k));
  unsigned char* touched = (use_upd && ctx->sparse_partials) ? ctx->part_touched.as<unsigned char>() : nullptr;
  // sorted update (lkm_sortupd.cuh) for unordered rows: decided after the first iteration of the call from the number of
  // label changes between neighbouring rows (a one-iteration first batch, like the other probes)
  const bool sort_ok = use_upd && ctx->sort_upd != 0 && k <= (uint64_t)kSrtMaxK && n < (1ull << 32) - 64 &&
                       (ctx->sort_upd > 0 || n * d * sizeof(F) >= (32ull << 20));
  bool use_sorted = sort_ok && ctx->sort_upd > 0;
  bool sort_probe = sort_ok && !use_sorted;
  if (use_sorted) tc5_can_fuse = false;
  const unsigned long long srt_segs = (n + kSrtSegRows - 1) / kSrtSegRows;
  const unsigned long long su_warps = (unsigned long long)ctx->sm_count * 2 * kSuWarps;
  const unsigned long long su_per = std::max<unsigned long long>(32, ((n + su_warps - 1) / su_warps + 31) / 32 * 32);
  if (sort_ok) {
    CK(ctx->srt_hist.ensure((size_t)srt_segs * k * 4));
    CK(ctx->srt_seg.ensure((k + 2) * 4 + 64));
    CK(ctx->srt_perm.ensure((n + 64) * 4));
    CK(ctx->srt_ent.ensure((size_t)(su_warps + k) * d * 4));
    CK(ctx->srt_inert.ensure((size_t)su_warps * 8));
    CK(cudaMemsetAsync(ctx->srt_seg.p, 0, (k + 2) * 4 + 64, ctx->stream));
    if (kSrtWarps * k * 4 > 48 * 1024) {
      CK(cudaFuncSetAttribute(sort_hist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSrtWarps * k * 4)));
      CK(cudaFuncSetAttribute(sort_scatter_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(kSrtWarps * k * 4)));
    }
  }
  CK(ctx->part_inertia.ensure((size_t)std::max(grid, 1024) * 8));
  const int fin_grid = (int)((kd + kFinElems - 1) / kFinElems);
  CK(ctx->shift_part.ensure((size_t)std::max(std::max(fin_grid, grid), kFxMaxCtas) * 8));
  // pipeline kernel on one GPU: optionally the finalize step folded into the kernel tail (LKM_FUSED_FINALIZE=1; off by default)
  const bool fin_fused = use_tc && tc_threads == kT4Threads && ctx->world == 1 && ctx->fused_finalize && tc_grid <= ctx->sm_count;
  if (fin_fused) {
    CK(ctx->fin_sync.ensure(16));
    CK(cudaMemsetAsync(ctx->fin_sync.p, 0, 16, ctx->stream));
  }
  CK(ctx->counts_out.ensure(k * 8));
  CK(ctx->state.ensure(sizeof(LloydState)));
  // lkm_set_keep_labels: the fused kernels also write the label every row was accumulated under (the two-pass shapes
  // always do: their update pass reads them)
  uint32_t* dbg_labels = nullptr;
  if (ctx->keep_labels) {
    CK(ctx->labels.ensure((n + 256) * 4));
    dbg_labels = ctx->labels.as<uint32_t>();
  }
  ctx->kept_labels_n = (ctx->keep_labels || !pf.fused) ? n : 0;
  const size_t L = kd + k + 1;
  if (ctx->world > 1) {
    CK(ctx->gather.ensure((size_t)ctx->world * L * 8));
    CKS(p2p_setup(ctx, L));
  }
  const bool use_p2p = ctx->world > 1 && ctx->p2p_enabled && ctx->p2p_local != nullptr && L <= ctx->p2p_cap;
  LloydState* dstate = ctx->state.as<LloydState>();
  ctx->last_update_path = 0;
  CK(cudaMemsetAsync(dstate, 0, sizeof(LloydState), ctx->stream));
  const bool flush = ctx->flush_bytes > 0;
  if (flush) {
    if (max_iter > 100000) return fail(ctx, LKM_ERR_INVALID_ARG, "l2 flush mode is for short benchmark runs");
    CK(ctx->flushbuf.ensure(ctx->flush_bytes));
    while (ctx->iter_events.size() < 3 * max_iter) {
      cudaEvent_t e;
      CK(cudaEventCreate(&e));
      ctx->iter_events.push_back(e);
    }
  }
  CK(cudaEventRecord(ctx->ev0, ctx->stream));

  uint64_t launched = 0;
  const uint64_t batch = 16;
  // Scoring passes of the tensor-core kernels.  One pass (x_hi.c_hi) certifies well-separated data at the lowest cost; the
  // first batch of every call is ONE iteration, after which the host looks at the rows that needed the exact re-scan and at
  // the tiles the fused update had to list, and moves to 3xTF32 / to the separate update pass when they are too many.  The
  // choice is made afresh in every call from what this call's first iteration saw — nothing is remembered between calls, so
  // the same call on the same data gives the same bits (labels never depend on the variant, the order of the sums does).
  const bool is_tc4 = use_tc && tc_threads == kT4Threads;
  int tc4_passes = is_tc4 ? (ctx->tc4_passes_forced ? ctx->tc4_passes_forced : 1) : 3;
  bool tc4_probe = is_tc4 && !ctx->tc4_passes_forced;
  int tc5_passes = use_tc5 ? (ctx->tc5_passes_forced ? ctx->tc5_passes_forced : 1) : 3;
  if (use_tc5 && d == 32) tc5_passes = 1;   // the d = 32 instantiation only exists as the screening variant
  bool tc5_probe = use_tc5 && ((!ctx->tc5_passes_forced && tc5_passes == 1) || tc5_can_fuse);
  // an exact re-scan reads every centroid from L2 (k * d * 4 bytes per row): past ~0.4 % of the rows the 3xTF32 variant is
  // cheaper for the pair kernel (measured on the cfg3 shape from a k-means++ start); the pipeline kernel re-scans from
  // shared memory and tolerates 2 %
  const double tc5_rescan_limit = ctx->tc5_defer ? ctx->tc5_rescan_limit : 0.004, tc4_rescan_limit = 0.02, fuse_mixed_limit = 0.25;
  unsigned long long mixed_seen = 0;
  unsigned long long fb_seen = 0;
  // the listed rows go through the pivot pruning first: always in the first batch of a call (one iteration when the screening
  // kernel probes), afterwards when the previous batch listed more than piv_limit rows per iteration
  bool piv_now = piv_ok;
  const double piv_limit = 2048.0;
  for (;;) {
    const uint64_t todo = std::min<uint64_t>((tc4_probe || tc5_probe) ? 1 : batch, max_iter - launched);
    for (uint64_t q = 0; q < todo; ++q) {
      const uint64_t it = launched + q;
      const F* Ccur = dC2 + (it & 1) * kd;
      F* Cnext = dC2 + ((it + 1) & 1) * kd;
      if (flush) {
        CK(cudaMemsetAsync(ctx->flushbuf.p, (int)(it & 0xff), ctx->flush_bytes, ctx->stream));
        if (g_flush_sweep && ctx->flush_bytes >= 64) {
          l2_read_sweep_kernel<<<ctx->sm_count * 4, 512, 0, ctx->stream>>>(reinterpret_cast<const uint4*>(ctx->flushbuf.p),
                                                                            (size_t)(ctx->flush_bytes / 2 / 16),
                                                                            reinterpret_cast<unsigned int*>(ctx->flushbuf.p));
          CK(cudaGetLastError());
        }
        if (use_p2p) {   // rendezvous of the ranks behind their flushes (see rank_barrier_kernel)
          RankBarArgs rb{};
          const size_t cap = ctx->p2p_cap, w = (size_t)ctx->world;
          rb.world = ctx->world; rb.epoch = ++ctx->bar_epoch;
          for (int pr = 0; pr < ctx->world; ++pr)
            rb.peer[pr] = reinterpret_cast<unsigned long long*>(reinterpret_cast<double*>(ctx->p2p_peer[pr]) + 2 * w * cap) + 2 * w * kFxMaxCtas + (size_t)ctx->rank;
          rb.mine = reinterpret_cast<const unsigned long long*>(reinterpret_cast<double*>(ctx->p2p_local) + 2 * w * cap) + 2 * w * kFxMaxCtas;
          rank_barrier_kernel<<<1, 32, 0, ctx->stream>>>(rb);
          CK(cudaGetLastError());
        }
        CK(cudaEventRecord(ctx->iter_events[3 * it], ctx->stream));
      }
      trace_mark(ctx, "start");
      int fin_src = grid, fin_src_inertia = use_upd ? upd_gx * upd_ny : 0;   // partial sources of this iteration's finalize
      bool upd_partials = false;   // the partials come from update_rows_kernel (and the fused pair kernel): `touched` map is valid
      if (use_tc) {
        if constexpr (std::is_same<F, float>::value) {
          TcArgs ta{};
          ta.X = dX; ta.n = n; ta.k = (int)k; ta.kp = tc_kp; ta.tcols = tc_cols; ta.C = Ccur;
          ta.part_sums = ctx->part_sums.as<float>(); ta.part_counts = ctx->part_counts.as<uint32_t>();
          ta.part_inertia = ctx->part_inertia.as<double>(); ta.state = dstate; ta.accumulate = 1; ta.labels = dbg_labels;
          if (tc_threads == kT4Threads) {
            Tc4Args t4{};
            t4.X = dX; t4.n = n; t4.k = (int)k; t4.pack_mult = 64u; t4.stages = tc4_passes == 1 ? tc4_nstages1 : tc4_nstages; t4.C = Ccur; t4.xmax2_bits = ctx->xmax2.as<unsigned int>();
#if defined(LKM_T4_PROF) || defined(LKM_T4_TRACE)
            CK(ctx->t4prof.ensure((64 + 40 * 16 + 4 * 160) * 8));
            t4.prof = ctx->t4prof.as<long long>();
#endif
            t4.part_sums = ta.part_sums; t4.part_counts = ta.part_counts; t4.part_inertia = ta.part_inertia; t4.state = dstate;
            t4.labels_out = dbg_labels;
            if (fin_fused) {   // single GPU: the kernel's tail does the finalize step (no finalize_kernel launch)
              t4.fin_sync = ctx->fin_sync.as<unsigned int>(); t4.C_new = Cnext; t4.counts_out = ctx->counts_out.as<double>();
              t4.shift_part = ctx->shift_part.as<double>(); t4.tol = tol; t4.max_iter = max_iter;
            }
            CK(launch_tc4(tc_kp, tc4_passes, tc_grid, tc4_passes == 1 ? tc_smem1 : tc_smem, ctx->stream, t4, tc4_map));
          } else CK(launch_pdl(lloyd_tc_kernel, tc_grid, 128, tc_smem, ctx->stream, ta, true));
          ctx->launches++;
        }
      } else if (use_f6) {
        if constexpr (std::is_same<F, double>::value) {
          f6::Args fa6{};
          fa6.X = dX; fa6.n = n; fa6.k = (int)k; fa6.C = Ccur;
          fa6.part_sums = ctx->part_sums.as<double>(); fa6.part_counts = ctx->part_counts.as<uint32_t>();
          fa6.part_inertia = ctx->part_inertia.as<double>(); fa6.state = dstate; fa6.state_rw = dstate; fa6.labels_out = dbg_labels;
          void* kargs[] = {(void*)&fa6};
          CK(cudaLaunchKernel(f6_fn, dim3((unsigned)pf.grid), dim3(f6::kThreads), kargs, f6_smem, ctx->stream));
          CK(cudaGetLastError());
          ctx->launches++;
        }
      } else if (pf.fused) {
        LloydArgs<F> a{};
        a.X = dX; a.n = n; a.d = (int)d; a.k = (int)k; a.C = Ccur;
        a.part_sums = ctx->part_sums.as<F>(); a.part_counts = ctx->part_counts.as<uint32_t>();
        a.part_inertia = ctx->part_inertia.as<double>(); a.state = dstate; a.labels = dbg_labels;
        a.n_subsets = pf.n_subsets; a.x_stride = pf.x_stride; a.kc = pf.kc; a.chunk_base = 0; a.chunk_rows = (int)k;
        a.vec = pf.vec; a.tpr = pf.tpr; a.certified = pf.certified; a.state_rw = dstate;
        CK((launch_lloyd<F, MODE_FUSED>(a, pf.grid, pf.smem, ctx->stream)));
        ctx->launches++;
      } else {
        bool assigned = false;
        bool fused_now = false;   // this iteration's screening kernel also accumulated the uniform tiles
        if constexpr (std::is_same<F, float>::value) {
          if (use_tc5) {
            Tc5Args t5{};
            t5.X = dX; t5.n = n; t5.k = (int)k; t5.C = Ccur; t5.xmax2_bits = ctx->xmax2.as<unsigned int>();
            t5.labels = ctx->labels.as<uint32_t>(); t5.state = dstate; t5.pack_mult = 256u;
            t5.n_chunks = tc5_chunks; t5.chunk_n = tc5_chunk_n; t5.merge = tc5_chunks > 1 ? ctx->t5merge.as<uint4>() : nullptr;
#ifdef LKM_T5_TRACE
            CK(ctx->t4prof.ensure(2 * 16 * 16 * 8));
            t5.prof = ctx->t4prof.as<long long>();
#endif
            unsigned int* rescan_count = nullptr;
            if (ctx->tc5_defer) {
              t5.rescan_list = ctx->ham_list.as<uint32_t>();
              rescan_count = reinterpret_cast<unsigned int*>(ctx->ham_list.as<uint32_t>() + ((n + 256) & ~(uint64_t)3));
              t5.rescan_count = rescan_count;
              CK(cudaMemsetAsync(rescan_count, 0, 4, ctx->stream));
            }
            // one launch per chunk of 256 centroids; the per-row (best, runner-up) state travels through t5merge
            for (int ch = 0; ch < tc5_chunks; ++ch) {
              t5.chunk_idx = ch;
              if (tc5_passes == 1 && tc5_can_fuse && ch == tc5_chunks - 1) {
                t5.part_sums = ctx->part_sums.as<float>(); t5.part_counts = ctx->part_counts.as<uint32_t>();
                t5.part_inertia = ctx->part_inertia.as<double>(); t5.part_touched = touched;
                t5.mixed_list = ctx->t5list.as<uint32_t>() + tc5_grid; t5.mixed_count = ctx->t5list.as<uint32_t>();
                t5.list_cap = tc5_list_cap;
                if (d == 32) assign_tc5s_kernel<32, true, 64><<<tc5_grid, t5_fuse_threads(32), tc5s_smem_bytes(32), ctx->stream>>>(t5, tc5_map);
                else if (d == 64) assign_tc5s_kernel<64, true, 256><<<tc5_grid, t5_fuse_threads(64), tc5s_smem_bytes(64), ctx->stream>>>(t5, tc5_map);
                else assign_tc5s_kernel<128, true, 256><<<tc5_grid, t5_fuse_threads(128), tc5s_smem_bytes(128), ctx->stream>>>(t5, tc5_map);
                fused_now = true;
              } else if (tc5_passes == 1) {
                if (d == 32) assign_tc5s_kernel<32, false, 64><<<tc5_grid, kT5Threads, tc5s_smem_bytes(32), ctx->stream>>>(t5, tc5_map);
                else if (d == 64) assign_tc5s_kernel<64, false, 256><<<tc5_grid, kT5Threads, tc5s_smem_bytes(64), ctx->stream>>>(t5, tc5_map);
                else assign_tc5s_kernel<128, false, 256><<<tc5_grid, kT5Threads, tc5s_smem_bytes(128), ctx->stream>>>(t5, tc5_map);
              } else if (d == 64) assign_tc5_kernel<64><<<tc5_grid, kT5Threads, tc5_smem, ctx->stream>>>(t5, tc5_map);
              else assign_tc5_kernel<128><<<tc5_grid, kT5Threads, tc5_smem, ctx->stream>>>(t5, tc5_map);
              CK(cudaGetLastError());
              ctx->launches++;
            }
            if (ctx->tc5_defer) {
              // second stage: the rows the tensor-core certificate could not decide (fp32 FMA certificate, then the exact scan)
              RowsArgs<float> ra{};
              ra.X = dX; ra.n = n; ra.d = (int)d; ra.k = (int)k; ra.C = Ccur; ra.list = t5.rescan_list; ra.list_count = rescan_count;
              ra.labels = ctx->labels.as<uint32_t>(); ra.state = dstate;
              trace_mark(ctx, "assign_tc5");
              if (piv_now && d != 32) {
                // decide the listed rows among the centroids near their provisional label; what stays open goes on to rows_top2_kernel
                PivotArgs pv{};
                pv.X = dX; pv.k = (int)k; pv.C = Ccur; pv.CC = ctx->piv_cc.as<float>(); pv.list = t5.rescan_list; pv.count = rescan_count;
                pv.labels = ctx->labels.as<uint32_t>(); pv.list2 = ctx->piv_list2.as<uint32_t>();
                pv.count2 = reinterpret_cast<unsigned int*>(ctx->piv_list2.as<uint32_t>() + ((n + 256) & ~(uint64_t)3));
                pv.state = dstate;
                CK(cudaMemsetAsync(pv.count2, 0, 4, ctx->stream));
                if (d == 64) {
                  centroid_dist_kernel<64><<<(unsigned)k, 128, 0, ctx->stream>>>(pv);
                  rows_pivot_kernel<64, false><<<ctx->sm_count * 4, kPivThreads, 0, ctx->stream>>>(pv);
                } else {
                  centroid_dist_kernel<128><<<(unsigned)k, 128, 0, ctx->stream>>>(pv);
                  rows_pivot_kernel<128, false><<<ctx->sm_count * 3, kPivThreads, 0, ctx->stream>>>(pv);
                }
                CK(cudaGetLastError());
                ctx->launches += 2;
                ra.list = pv.list2; ra.list_count = pv.count2;
                trace_mark(ctx, "rows_pivot");
              }
              CKS(launch_rows<float>(ctx, ra, n));
              trace_mark(ctx, "rows_top2");
            }
            assigned = true;
          } else if (use_tca) {
            // tensor-core assignment: split the centroids into TF32 operand tiles, then score
            CK(cudaMemsetAsync(ctx->tca_cmax.p, 0, 4, ctx->stream));
            tc_prep_centroids_kernel<<<tca_chunks * tca_ka, 64, 0, ctx->stream>>>(Ccur, (int)k, (int)d, tca_ka, ctx->tca_bsplit.as<unsigned char>(),
                                                                              ctx->tca_cn.as<float>(), ctx->tca_cmax.as<unsigned int>());
            CK(cudaGetLastError());
            TcAssignArgs ta{};
            ta.X = dX; ta.n = n; ta.k = (int)k; ta.d = (int)d; ta.ka = tca_ka; ta.n_chunks = tca_chunks;
            ta.bsplit = ctx->tca_bsplit.as<unsigned char>(); ta.cn = ctx->tca_cn.as<float>();
            ta.cmax_bits = ctx->tca_cmax.as<unsigned int>(); ta.C = Ccur;
            ta.labels = ctx->labels.as<uint32_t>(); ta.dists = ctx->dists.as<float>(); ta.state = dstate;
            tc_assign_kernel<<<tca_grid, 128, tca_smem, ctx->stream>>>(ta);
            CK(cudaGetLastError());
            ctx->launches += 2;
            assigned = true;
          }
        }
        if constexpr (std::is_same<F, double>::value) {
          if (use_dm) {
            CK(ctx->ham_list.ensure((n + 256) * 4 + 16));
            unsigned int* rescan_count = reinterpret_cast<unsigned int*>(ctx->ham_list.as<uint32_t>() + ((n + 256) & ~(uint64_t)3));
            CK(cudaMemsetAsync(rescan_count, 0, 4, ctx->stream));
            DmmaArgs da{};
            da.X = dX; da.n = n; da.d = (int)d; da.k = (int)k; da.C = Ccur; da.xs = dmma_row_stride((int)d);
            da.labels = ctx->labels.as<uint32_t>(); da.dists = ctx->dists.as<double>();
            da.rescan_list = ctx->ham_list.as<uint32_t>(); da.rescan_count = rescan_count; da.state = dstate; da.state_rw = dstate;
            const size_t dsm = dmma_smem_bytes((int)d);
            if (dsm > 48 * 1024) CK(cudaFuncSetAttribute(assign_dmma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dsm));
            const int occ = dsm > 100 * 1024 ? 1 : 2;
            const int dgrid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + kDmRows - 1) / kDmRows, (uint64_t)ctx->sm_count * occ));
            assign_dmma_kernel<<<dgrid, kDmThreads, dsm, ctx->stream>>>(da);
            CK(cudaGetLastError());
            ctx->launches++;
            RowsArgs<double> ra{};
            ra.X = dX; ra.n = n; ra.d = (int)d; ra.k = (int)k; ra.C = Ccur; ra.list = da.rescan_list; ra.list_count = rescan_count;
            ra.labels = ctx->labels.as<uint32_t>(); ra.best_dist = ctx->dists.as<double>(); ra.state = dstate;
            CKS(launch_rows<double>(ctx, ra, n));
            assigned = true;
          }
        }
        if (!assigned) {
        LloydArgs<F> a{};
        a.X = dX; a.n = n; a.d = (int)d; a.k = (int)k; a.C = Ccur;
        a.labels = ctx->labels.as<uint32_t>(); a.dists = ctx->dists.as<F>(); a.state = dstate;
        a.n_subsets = 1; a.x_stride = pa.x_stride; a.kc = pa.kc; a.chunk_rows = (int)k; a.vec = pa.vec; a.tpr = pa.tpr; a.certified = pa.certified; a.state_rw = dstate;
        CK((launch_lloyd<F, MODE_ASSIGN>(a, pa.grid, pa.smem, ctx->stream)));
        ctx->launches++;
        }
        bool updated = false;
        if constexpr (std::is_same<F, float>::value) {
          if (use_upd && sort_probe) {
            // decided here, between the assignment and the update of the call's first iteration: one host round trip (~20 us)
            // against 1.6 ms of streaming update on the unordered rows of cfg3.  More than one label change per four rows:
            // the streaming update is bound by its per-row bookkeeping, sort instead.  What the fused screening kernel may
            // already have accumulated is simply not read (the sorted update sums every row).
            label_breaks_kernel<<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->labels.as<uint32_t>(), n, dstate);
            CK(cudaGetLastError());
            ctx->launches++;
            CK(cudaMemcpyAsync(ctx->h_state, dstate, sizeof(LloydState), cudaMemcpyDeviceToHost, ctx->stream));
            CK(cudaStreamSynchronize(ctx->stream));
            use_sorted = (double)ctx->h_state->label_breaks > 0.25 * (double)n;
            if (use_sorted) tc5_can_fuse = false;
            sort_probe = false;
          }
          if (use_upd && use_sorted) {
            SortArgs sa{};
            sa.labels = ctx->labels.as<uint32_t>(); sa.n = n; sa.k = (int)k; sa.hist = ctx->srt_hist.as<uint32_t>();
            sa.seg_start = ctx->srt_seg.as<uint32_t>(); sa.perm = ctx->srt_perm.as<uint32_t>();
            sa.scan_done = ctx->srt_seg.as<unsigned int>() + (k + 2); sa.state = dstate;
            const unsigned sgrid = (unsigned)((srt_segs + kSrtWarps - 1) / kSrtWarps);
            const size_t ssm = (size_t)kSrtWarps * k * 4;
            sort_hist_kernel<<<sgrid, kSrtThreads, ssm, ctx->stream>>>(sa);
            sort_scan_kernel<<<(unsigned)((k + 31) / 32), dim3(32, 32), 0, ctx->stream>>>(sa);
            sort_scatter_kernel<<<sgrid, kSrtThreads, ssm, ctx->stream>>>(sa);
            CK(cudaGetLastError());
            trace_mark(ctx, "sort_labels");
            SortUpdArgs su{};
            su.X = dX; su.n = n; su.k = (int)k; su.perm = sa.perm; su.seg_start = sa.seg_start; su.C = Ccur;
            su.ent_sums = ctx->srt_ent.as<float>(); su.ent_inertia = ctx->srt_inert.as<double>();
            su.per_warp = su_per; su.n_warps = su_warps;
            su.part_sums = ctx->part_sums.as<float>(); su.part_counts = ctx->part_counts.as<uint32_t>();
            su.part_touched = touched; su.part_inertia = ctx->part_inertia.as<double>(); su.state = dstate;
            const unsigned ugrid = (unsigned)(su_warps / kSuWarps);
            if (d == 32) {
              update_sorted_kernel<32><<<ugrid, kSuThreads, 0, ctx->stream>>>(su);
              sorted_combine_kernel<32><<<(unsigned)k + 1, 32, 0, ctx->stream>>>(su);
            } else if (d == 64) {
              update_sorted_kernel<64><<<ugrid, kSuThreads, 0, ctx->stream>>>(su);
              sorted_combine_kernel<64><<<(unsigned)k + 1, 64, 0, ctx->stream>>>(su);
            } else {
              update_sorted_kernel<128><<<ugrid, kSuThreads, 0, ctx->stream>>>(su);
              sorted_combine_kernel<128><<<(unsigned)k + 1, 128, 0, ctx->stream>>>(su);
            }
            CK(cudaGetLastError());
            ctx->launches += 5;
            trace_mark(ctx, "update_sorted");
            ctx->last_update_path = 2;
            fin_src = 1;
            fin_src_inertia = 1;
            updated = true;
            upd_partials = true;
          } else if (use_upd) {
            UpdArgs ua{};
            ua.X = dX; ua.n = n; ua.k = (int)k; ua.chunk_rows = upd_chunk; ua.stages = upd_stages;
            ua.labels = ctx->labels.as<uint32_t>(); ua.C = Ccur;
            ua.part_sums = ctx->part_sums.as<float>(); ua.part_counts = ctx->part_counts.as<uint32_t>();
            ua.part_inertia = ctx->part_inertia.as<double>(); ua.state = dstate; ua.part_touched = touched;
            int ugx = upd_gx;
            if (fused_now) {
              // only the tiles the screening kernel listed; its own partials are sources [0, tc5_grid), these follow
              ua.tile_list = ctx->t5list.as<uint32_t>() + tc5_grid; ua.tile_count = ctx->t5list.as<uint32_t>();
              ua.list_cap = tc5_list_cap;
              ua.part_sums += (size_t)tc5_grid * kd; ua.part_counts += (size_t)tc5_grid * k; ua.part_inertia += tc5_grid;
              if (touched) ua.part_touched += (size_t)tc5_grid * k;
              ugx = tc5_grid;
            }
            fin_src = fused_now ? 2 * tc5_grid : upd_gx;
            fin_src_inertia = fused_now ? tc5_grid * (1 + upd_ny) : upd_gx * upd_ny;
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3((unsigned)ugx, (unsigned)upd_ny);
            cfg.blockDim = dim3((unsigned)kUpdThreads);
            cfg.dynamicSmemBytes = upd_smem;
            cfg.stream = ctx->stream;
            if (d == 32) CK(cudaLaunchKernelEx(&cfg, update_rows_kernel<32>, ua, upd_map));
            else if (d == 64) CK(cudaLaunchKernelEx(&cfg, update_rows_kernel<64>, ua, upd_map));
            else CK(cudaLaunchKernelEx(&cfg, update_rows_kernel<128>, ua, upd_map));
            ctx->launches++;
            trace_mark(ctx, "update_rows");
            ctx->last_update_path = 1;
            updated = true;
            upd_partials = true;
          }
        }
        if (!updated) {
        // inertia: fixed-order f64 partial sums of the min distances, one per source slot
        sum_f64_kernel<F><<<grid, 256, 0, ctx->stream>>>(ctx->dists.as<F>(), n, ctx->part_inertia.as<double>());
        CK(cudaGetLastError());
        ctx->launches += 1;
        }
        for (int ch = 0; ch < (updated ? 0 : n_chunks); ++ch) {
          LloydArgs<F> u{};
          u.X = dX; u.n = n; u.d = (int)d; u.k = (int)k; u.labels_in = ctx->labels.as<uint32_t>();
          u.chunk_base = ch * chunk_rows;
          u.chunk_rows = (int)std::min<uint64_t>(chunk_rows, k - (uint64_t)ch * chunk_rows);
          u.part_sums = ctx->part_sums.as<F>(); u.part_counts = ctx->part_counts.as<uint32_t>();
          u.state = dstate; u.n_subsets = 1; u.x_stride = pu.x_stride; u.vec = pu.vec; u.tpr = pu.tpr;
          CK((launch_lloyd<F, MODE_UPDATE>(u, pu.grid, pu.smem, ctx->stream)));
          ctx->launches++;
        }
      }
      if (flush && ctx->kernel_timing) CK(cudaEventRecord(ctx->iter_events[3 * it + 1], ctx->stream));
      ctx->last_fin_src = fin_fused ? -1 : fin_src;
      ctx->last_fin_src_inertia = fin_src_inertia;
      // combine the per-CTA partials in fixed order, then the m_k-means update
      if (fin_fused) {
        // done inside lloyd_tc4_kernel
      } else if (upd_partials && touched != nullptr && ctx->fin_cl && fin_src <= kFclMaxSrc && (ctx->world == 1 || use_p2p)) {
        // sparse partials: cluster-wise combine (finalize_cl_kernel), with the NVLink exchange when row-sharded
        FinalizeXArgs<F, F, uint32_t> xa{};
        xa.src_sums = ctx->part_sums.as<F>(); xa.src_counts = ctx->part_counts.as<uint32_t>();
        xa.src_inertia = ctx->part_inertia.as<double>();
        xa.sum_stride = kd; xa.cnt_stride = k; xa.inertia_stride = 1; xa.n_src = fin_src; xa.n_src_inertia = fin_src_inertia;
        xa.k = (int)k; xa.d = (int)d; xa.src_touched = touched;
        const int g_max = std::min(kFxMaxCtas, 2 * ctx->sm_count);
        xa.slices_per_cta = (int)((k + g_max - 1) / g_max);
        const int xgrid = (int)((k + xa.slices_per_cta - 1) / xa.slices_per_cta);
        xa.world = ctx->world;
        if (ctx->world > 1) {
          const unsigned long long epoch = ++ctx->p2p_epoch;
          const size_t par = (size_t)(epoch & 1);
          const size_t cap = ctx->p2p_cap, w = (size_t)ctx->world;
          for (int pr = 0; pr < ctx->world; ++pr) {
            double* base = reinterpret_cast<double*>(ctx->p2p_peer[pr]);
            xa.peer_out[pr] = base + (par * w + (size_t)ctx->rank) * cap;
            xa.peer_flags[pr] = reinterpret_cast<unsigned long long*>(base + 2 * w * cap) + (par * w + (size_t)ctx->rank) * kFxMaxCtas;
          }
          double* mine = reinterpret_cast<double*>(ctx->p2p_local);
          xa.gather = mine + par * w * cap;
          xa.my_flags = reinterpret_cast<const unsigned long long*>(mine + 2 * w * cap) + par * w * kFxMaxCtas;
          xa.cap = cap; xa.epoch = epoch;
          if (ctx->p2p_ll) {
            for (int pr = 0; pr < ctx->world; ++pr)
              xa.peer_ll[pr] = reinterpret_cast<uint4*>(static_cast<unsigned char*>(ctx->p2p_peer[pr]) + ctx->p2p_ll_off) + (par * w + (size_t)ctx->rank) * cap;
            xa.my_ll = reinterpret_cast<const uint4*>(static_cast<unsigned char*>(ctx->p2p_local) + ctx->p2p_ll_off) + par * w * cap;
          }
        }
        xa.C_old = Ccur; xa.C_new = Cnext; xa.counts_out = ctx->counts_out.as<double>();
        xa.shift_part = ctx->shift_part.as<double>(); xa.state = dstate; xa.tol = tol; xa.max_iter = max_iter;
        if (ctx->step_trace) { CK(ctx->t4prof.ensure(64)); xa.prof = ctx->t4prof.as<unsigned long long>(); }
        CK(launch_pdl(finalize_cl_kernel<F, F, uint32_t>, xgrid, kFinThreads, 0, ctx->stream, xa, true));
        ctx->launches += 1;
      } else if (ctx->world == 1) {
        FinalizeArgs<F, F, uint32_t> fa{};
        fa.src_sums = ctx->part_sums.as<F>(); fa.src_counts = ctx->part_counts.as<uint32_t>();
        fa.src_inertia = ctx->part_inertia.as<double>();
        fa.sum_stride = kd; fa.cnt_stride = k; fa.inertia_stride = 1; fa.n_src = fin_src; fa.k = (int)k; fa.d = (int)d;
        fa.n_src_inertia = fin_src_inertia;
        fa.src_touched = upd_partials ? touched : nullptr;
        fa.C_old = Ccur; fa.C_new = Cnext; fa.counts_out = ctx->counts_out.as<double>();
        fa.shift_part = ctx->shift_part.as<double>(); fa.state = dstate; fa.tol = tol; fa.max_iter = max_iter;
        CK((launch_finalize<F, F, uint32_t, true>(fa, ctx->stream)));
        ctx->launches++;
      } else if (use_p2p) {
        // one launch: local reduction -> NVLink stores into every rank's gather buffer + per-CTA flags -> wait -> combine
        // in rank order -> update (finalize_x_kernel)
        FinalizeXArgs<F, F, uint32_t> xa{};
        xa.src_sums = ctx->part_sums.as<F>(); xa.src_counts = ctx->part_counts.as<uint32_t>();
        xa.src_inertia = ctx->part_inertia.as<double>();
        xa.sum_stride = kd; xa.cnt_stride = k; xa.inertia_stride = 1; xa.n_src = fin_src; xa.n_src_inertia = fin_src_inertia;
        xa.k = (int)k; xa.d = (int)d; xa.src_touched = upd_partials ? touched : nullptr;
        const int n_slices = (int)((kd + kFinElems - 1) / kFinElems);
        const int g_max = std::min(kFxMaxCtas, 2 * ctx->sm_count);
        xa.slices_per_cta = (n_slices + g_max - 1) / g_max;
        const int xgrid = (n_slices + xa.slices_per_cta - 1) / xa.slices_per_cta;
        const unsigned long long epoch = ++ctx->p2p_epoch;
        const size_t par = (size_t)(epoch & 1);
        const size_t cap = ctx->p2p_cap, w = (size_t)ctx->world;
        for (int pr = 0; pr < ctx->world; ++pr) {
          double* base = reinterpret_cast<double*>(ctx->p2p_peer[pr]);
          xa.peer_out[pr] = base + (par * w + (size_t)ctx->rank) * cap;
          xa.peer_flags[pr] = reinterpret_cast<unsigned long long*>(base + 2 * w * cap) + (par * w + (size_t)ctx->rank) * kFxMaxCtas;
        }
        double* mine = reinterpret_cast<double*>(ctx->p2p_local);
        xa.gather = mine + par * w * cap;
        xa.my_flags = reinterpret_cast<const unsigned long long*>(mine + 2 * w * cap) + par * w * kFxMaxCtas;
        xa.cap = cap; xa.world = ctx->world; xa.epoch = epoch;
        xa.C_old = Ccur; xa.C_new = Cnext; xa.counts_out = ctx->counts_out.as<double>();
        xa.shift_part = ctx->shift_part.as<double>(); xa.state = dstate; xa.tol = tol; xa.max_iter = max_iter;
        CK(launch_pdl(finalize_x_kernel<F, F, uint32_t>, xgrid, kFinThreads, 0, ctx->stream, xa, true));
        ctx->launches += 1;
      } else if (ctx->group != nullptr) {
        return fail(ctx, LKM_ERR_UNSUPPORTED, "device group without peer access");
      } else {
        // NCCL fallback (LKM_P2P=0 or no peer access): reduce -> ncclAllGather -> combine, three launches
        FinalizeArgs<F, F, uint32_t> ra{};
        ra.src_sums = ctx->part_sums.as<F>(); ra.src_counts = ctx->part_counts.as<uint32_t>();
        ra.src_inertia = ctx->part_inertia.as<double>();
        ra.sum_stride = kd; ra.cnt_stride = k; ra.inertia_stride = 1; ra.n_src = fin_src; ra.k = (int)k; ra.d = (int)d;
        ra.n_src_inertia = fin_src_inertia;
        ra.src_touched = upd_partials ? touched : nullptr;
        ra.rank_out = ctx->gather.as<double>() + (size_t)ctx->rank * L;
        ra.state = dstate;
        FinalizeArgs<F, double, double> fa{};
        CK((launch_finalize<F, F, uint32_t, false>(ra, ctx->stream)));
        CKN(g_nccl.AllGather(ctx->gather.as<double>() + (size_t)ctx->rank * L, ctx->gather.as<double>(), L, kNcclFloat64,
                             ctx->comm, ctx->stream));
        const double* gbase = ctx->gather.as<double>();
        fa.src_sums = gbase; fa.src_counts = gbase + kd;
        fa.src_inertia = gbase + kd + k;
        fa.sum_stride = L; fa.cnt_stride = L; fa.inertia_stride = L; fa.n_src = ctx->world; fa.k = (int)k; fa.d = (int)d;
        fa.C_old = Ccur; fa.C_new = Cnext; fa.counts_out = ctx->counts_out.as<double>();
        fa.shift_part = ctx->shift_part.as<double>(); fa.state = dstate; fa.tol = tol; fa.max_iter = max_iter;
        CK((launch_finalize<F, double, double, true>(fa, ctx->stream)));
        ctx->launches += 2;  // reduce + update (the all-gather kernel is NCCL's, not counted)
      }
      trace_mark(ctx, "finalize");
      if (flush) CK(cudaEventRecord(ctx->iter_events[3 * it + 2], ctx->stream));
    }
    launched += todo;
    const uint64_t todo_done = todo;
    CK(cudaMemcpyAsync(ctx->h_state, dstate, sizeof(LloydState), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    {
      const unsigned long long fb = (ctx->h_state->fallback_rows & ((1ull << 40) - 1)) - fb_seen;
      if (is_tc4 && !ctx->tc4_passes_forced && tc4_passes < 3 && (double)fb > tc4_rescan_limit * (double)n * (double)todo_done) tc4_passes = 3;
      if (use_tc5 && !ctx->tc5_passes_forced && tc5_passes == 1 && d != 32 && (double)fb > tc5_rescan_limit * (double)n * (double)todo_done)
        tc5_passes = 3;
      // unordered rows: (nearly) every tile of the fused update ends up on the list, and the listed pass is slower than the
      // plain streaming update — stop fusing for the rest of the call
      const unsigned long long mixed = ctx->h_state->mixed_tiles - mixed_seen;
      if (tc5_can_fuse && (double)mixed > fuse_mixed_limit * (double)((n + 127) / 128) * (double)todo_done) tc5_can_fuse = false;
      mixed_seen = ctx->h_state->mixed_tiles;
      if (piv_ok && ctx->rows_pivot < 0) piv_now = (double)fb > piv_limit * (double)todo_done;
      tc4_probe = false;
      tc5_probe = false;
    }
    fb_seen = ctx->h_state->fallback_rows & ((1ull << 40) - 1);
    if (ctx->h_state->done || launched >= max_iter) break;
  }
  CK(cudaEventRecord(ctx->ev1, ctx->stream));
  CK(cudaEventSynchronize(ctx->ev1));
  CK(cudaEventElapsedTime(&out.ms, ctx->ev0, ctx->ev1));
  out.n_iter = ctx->h_state->n_iter;
  out.fallback_rows = ctx->h_state->fallback_rows;
  if (out.fallback_rows & (3ull << 40)) return fail(ctx, LKM_ERR_CUDA, "device-side wait timed out (tensor-core barrier or NVLink exchange flag)");
  if (flush) {
    // per-iteration spans only (the flush writes sit between them)
    out.ms = 0;
    for (uint64_t it = 0; it < launched; ++it) {
      float a = 0, b = 0;
      CK(cudaEventElapsedTime(&a, ctx->iter_events[3 * it], ctx->iter_events[3 * it + 2]));
      if (ctx->kernel_timing) CK(cudaEventElapsedTime(&b, ctx->iter_events[3 * it], ctx->iter_events[3 * it + 1]));
      out.ms += a;
      out.main_ms += b;
    }
  }
#ifdef LKM_T5_TRACE
  if (use_tc5) {
    static long long ht[2 * 16 * 16];
    CK(cudaMemcpy(ht, ctx->t4prof.p, sizeof(ht), cudaMemcpyDeviceToHost));
    std::fprintf(stderr, "t5trace cta tile  (3xTF32) Maready Maccfree Mcommit  Sbeg Sissued Sarrived  Ewake Edrained Elabels | (screening) Ptma0 Ptma3 Maccfree Mfull0 Mpfull0 Mfull3 Mpfull3 Mcommit Ewake Edrained Edone Edrained(warp 0..3) Mloop\n");
    for (int c = 0; c < 2; ++c)
      for (int j = 0; j < 16; ++j) {
        std::fprintf(stderr, "t5trace %d %3d", c, j);
        for (int e = 0; e < 16; ++e) std::fprintf(stderr, " %7lld", ht[(c * 16 + j) * 16 + e]);
        std::fprintf(stderr, "\n");
      }
  }
#endif
#ifdef LKM_T4_TRACE
  if (use_tc && tc_threads == kT4Threads) {
    static long long ht[64 + 40 * 16 + 4 * 160];
    CK(cudaMemcpy(ht, ctx->t4prof.p, sizeof(ht), cudaMemcpyDeviceToHost));
    {
      const long long* g = ht + 64 + 40 * 16;
      long long t0 = g[0];
      for (int b = 0; b < tc_grid; ++b) t0 = std::min(t0, g[b * 4]);
      for (int b = 0; b < tc_grid; ++b)
        std::fprintf(stderr, "t4cta %3d start %6lld loop %6lld loopend %6lld end %6lld ns\n", b, g[b * 4] - t0, g[b * 4 + 1] - t0, g[b * 4 + 2] - t0, g[b * 4 + 3] - t0);
    }
    std::fprintf(stderr, "t4kern init %lld pdl %lld staged %lld sync %lld ready %lld loopend(t0) %lld allend %lld reduced %lld dealloc %lld\n",
                 ht[0], ht[1], ht[2], ht[3], ht[4], ht[5], ht[6], ht[7], ht[8]);
    std::fprintf(stderr, "t4trace tile  tma  Sbeg Send  Mbeg Mend  Ewake Eld Ecomp Epub  Ubeg Uacc Uend  Malo Mtfree\n");
    for (int j = 0; j < 40; ++j) {
      std::fprintf(stderr, "t4trace %3d", j);
      for (int e = 0; e < 14; ++e) std::fprintf(stderr, " %6lld", ht[64 + j * 16 + e]);
      std::fprintf(stderr, "\n");
    }
  }
#endif
#ifdef LKM_T4_PROF
  if (use_tc && tc_threads == kT4Threads) {
    long long hp[48];
    CK(cudaMemcpy(hp, ctx->t4prof.p, sizeof(hp), cudaMemcpyDeviceToHost));
    const char* names[6] = {"P", "M", "S", "E0", "E1", "U"};
    for (int r = 0; r < 6; ++r)
      std::fprintf(stderr, "t4prof %-2s wait0 %9lld wait1 %9lld s2 %9lld s3 %9lld s4 %9lld total %9lld\n", names[r], hp[r * 8], hp[r * 8 + 1],
                   hp[r * 8 + 2], hp[r * 8 + 3], hp[r * 8 + 4], hp[r * 8 + 7]);
  }
#endif
  if (ctx->step_trace && ctx->t4prof.p != nullptr && ctx->t4prof.cap >= 64) {
    unsigned long long hp[8] = {0};
    cudaMemcpy(hp, ctx->t4prof.p, 56, cudaMemcpyDeviceToHost);
    if (hp[0] && hp[6] >= hp[0])
      std::fprintf(stderr, "step_trace finalize_cl (last iteration, CTA 0, ns from its start): reduced+published %lld fenced %lld peers' flags seen %lld combined %lld | last CTA at %lld done %lld\n",
                   (long long)(hp[1] - hp[0]), (long long)(hp[2] - hp[0]), (long long)(hp[3] - hp[0]), (long long)(hp[4] - hp[0]),
                   (long long)(hp[5] - hp[0]), (long long)(hp[6] - hp[0]));
  }
  trace_report(ctx);
  out.inertia = ctx->h_state->inertia;
  out.shift = ctx->h_state->shift;
  out.final_buf = (int)(out.n_iter & 1);
  return LKM_OK;
}

// compute_inertia (KM/algorithm.rs:608-620) of the resident rows under ctx->labels against dC; row-sharded runs add the
// per-rank totals in rank order
template <typename F>
static lkm_status inertia_pass(lkm_ctx* ctx, const F* dC, double& total) {
  const uint64_t n = ctx->n, d = ctx->d;
  CK(ctx->ham_red.ensure(1024 * 8));
  const F* dX = reinterpret_cast<const F*>(ctx->X);
  constexpr int kVec = 16 / (int)sizeof(F);
  const bool vec = d % kVec == 0 && (reinterpret_cast<uintptr_t>(dX) & 15) == 0 && (reinterpret_cast<uintptr_t>(dC) & 15) == 0;
  const uint64_t per_row = vec ? d / kVec : d;   // loads per row
  int lpr = 1;
  while (lpr < 32 && (uint64_t)lpr < per_row) lpr *= 2;
  const uint64_t rows_per_cta = (uint64_t)8 * (32 / lpr);
  const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + rows_per_cta - 1) / rows_per_cta, std::min(1024, ctx->sm_count * 8)));
  if (vec) inertia_rows_kernel<F, kVec><<<grid, 256, 0, ctx->stream>>>(dX, n, (int)d, dC, ctx->labels.as<uint32_t>(), ctx->ham_red.as<double>(), lpr);
  else inertia_rows_kernel<F, 1><<<grid, 256, 0, ctx->stream>>>(dX, n, (int)d, dC, ctx->labels.as<uint32_t>(), ctx->ham_red.as<double>(), lpr);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaMemcpyAsync(ctx->h_pin, ctx->ham_red.p, (size_t)grid * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  double tot = 0.0;
  for (int b = 0; b < grid; ++b) tot += ctx->h_pin[b];
  if (ctx->world > 1) {
    std::vector<unsigned char> all;
    CKS(comm_allgather_host(ctx, &tot, 8, all));
    tot = 0.0;
    for (int r = 0; r < ctx->world; ++r) { double v; std::memcpy(&v, all.data() + (size_t)r * 8, 8); tot += v; }
  }
  total = tot;
  return LKM_OK;
}

// One run of fit_hamerly (KM/algorithm.rs:258-279) from the centroids in buffer 0 of dC2.  Iteration 1 is Lloyd's
// (HamerlyAlgorithm::new assigns every row, :418-456); from iteration 2 on the bounds decide which rows are looked at.
// out.inertia = compute_inertia against the final centroids (:275,570-577); ctx->counts_out = the final memberships' counts.
template <typename F>
static lkm_status run_hamerly(lkm_ctx* ctx, F* dC2, uint64_t k, uint64_t max_iter, double tol, LloydOut& out) {
  const uint64_t n = ctx->n, d = ctx->d, kd = k * d;
  const F* dX = reinterpret_cast<const F*>(ctx->X);
  const int keep = ctx->keep_labels;
  ctx->keep_labels = 1;
  ctx->ham_rows_scanned = ctx->ham_rows_tightened = 0;
  // row-sharded runs (the running sums would need their own exchange) and k = 1 run Lloyd's iterations: same labels and
  // centroids, and the result below is still Hamerly's (post-update inertia, this run's counts)
  bool bounds = ctx->world == 1 && k >= 2 && n >= 1 && max_iter >= 2;
  if (bounds) {
    // scale of the fixed-point sums: |sum| <= n * max|x| must stay below 2^62
    if (ctx->absmax_epoch != ctx->data_epoch || !ctx->own_X) {
      CK(ctx->absmax.ensure(8));
      CK(cudaMemsetAsync(ctx->absmax.p, 0, 8, ctx->stream));
      abs_max_kernel<F><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(dX, n * d, ctx->absmax.as<unsigned long long>());
      CK(cudaGetLastError());
      ctx->launches++;
      CK(cudaMemcpyAsync(ctx->h_pin, ctx->absmax.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      ctx->absmax_value = ctx->h_pin[0];
      ctx->absmax_epoch = ctx->data_epoch;
    }
    if (!std::isfinite(ctx->absmax_value)) bounds = false;
  }
  lkm_status st = run_lloyd<F>(ctx, dC2, k, bounds ? 1 : max_iter, tol, out);
  ctx->keep_labels = keep;
  if (st != LKM_OK) return st;
  const bool finished = !bounds || out.shift < tol || out.n_iter >= max_iter;
  if (!finished) {
    int ex = 0;
    std::frexp(std::max(ctx->absmax_value, 1.0) * (double)(n + 2), &ex);   // n * max|x| < 2^ex
    const int S = std::max(-300, std::min(60, 61 - ex));
    const double scale = std::ldexp(1.0, S), inv_scale = std::ldexp(1.0, -S);
    CK(ctx->ham_ub.ensure((n + 1) * sizeof(F)));
    CK(ctx->ham_lb.ensure((n + 1) * sizeof(F)));
    CK(ctx->ham_list.ensure((n + 256) * 4));
    CK(ctx->ham_fx.ensure((kd + k) * 8));
    CK(ctx->ham_moved.ensure(k * sizeof(F)));
    CK(ctx->ham_s.ensure(k * sizeof(F)));
    CK(ctx->ham_state.ensure(sizeof(HamState)));
    const int cen_grid = (int)((kd + 255) / 256);
    CK(ctx->shift_part.ensure((size_t)std::max(cen_grid, 1024) * 8));
    LloydState* dstate = ctx->state.as<LloydState>();
    HamState* hs = ctx->ham_state.as<HamState>();
    long long* fx_sums = ctx->ham_fx.as<long long>();
    long long* fx_counts = fx_sums + kd;
    CK(cudaEventRecord(ctx->ev0, ctx->stream));
    // the loop continues where iteration 1 stopped
    std::memset(ctx->h_state, 0, sizeof(LloydState));
    ctx->h_state->n_iter = 1;
    CK(cudaMemcpyAsync(dstate, ctx->h_state, sizeof(LloydState), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(hs, 0, sizeof(HamState), ctx->stream));
    // running sums and counts of iteration 1's assignment, as exact integers (see hamerly_sums_kernel)
    CK(cudaMemsetAsync(fx_sums, 0, (kd + k) * 8, ctx->stream));
    hamerly_sums_kernel<F><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(dX, n, (int)d, ctx->labels.as<uint32_t>(), scale, fx_sums, fx_counts);
    CK(cudaGetLastError());
    ctx->launches++;
    F* ub = ctx->ham_ub.as<F>();
    F* lb = ctx->ham_lb.as<F>();
    fill_kernel<F><<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ub, n, std::numeric_limits<F>::infinity());
    fill_kernel<F><<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(lb, n, (F)0);
    CK(cudaGetLastError());
    ctx->launches += 2;
    auto moved_and_state = [&](const F* c_old, const F* c_new, bool with_state) -> lkm_status {
      HamMovedArgs<F> ma{};
      ma.C_old = c_old; ma.C_new = c_new; ma.k = (int)k; ma.d = (int)d; ma.moved = ctx->ham_moved.as<F>(); ma.s = ctx->ham_s.as<F>();
      ma.state = with_state ? dstate : nullptr;
      hamerly_moved_kernel<F><<<(int)k, 128, 0, ctx->stream>>>(ma);
      HamStateArgs<F> sa{};
      sa.shift_part = ctx->shift_part.as<double>(); sa.n_part = cen_grid; sa.moved = ctx->ham_moved.as<F>(); sa.k = (int)k;
      sa.hs = hs; sa.state = with_state ? dstate : nullptr; sa.tol = tol; sa.max_iter = max_iter;
      hamerly_state_kernel<F><<<1, 32, 0, ctx->stream>>>(sa);
      CK(cudaGetLastError());
      ctx->launches += 2;
      return LKM_OK;
    };
    CKS(moved_and_state(dC2, dC2 + kd, false));   // iteration 1 moved the centroids from buffer 0 to buffer 1
    const bool ham_pivot = std::is_same<F, float>::value && ctx->rows_pivot != 0 && (d == 64 || d == 128) && k <= 4096 && n < (1ull << 32) - 512 &&
                           (reinterpret_cast<uintptr_t>(dX) & 15) == 0;
    if (ham_pivot) {
      CK(ctx->piv_cc.ensure(k * k * 4));
      CK(ctx->piv_list2.ensure((n + 256) * 4 + 64));
    }
    uint64_t it = 1;
    const int bgrid = (int)std::max<uint64_t>(1, std::min<uint64_t>((n + 255) / 256, (uint64_t)ctx->sm_count * 8));
    for (;;) {
      const uint64_t todo = std::min<uint64_t>(8, max_iter - it);
      for (uint64_t q = 0; q < todo; ++q, ++it) {
        const F* cur = dC2 + (it & 1) * kd;
        F* nxt = dC2 + ((it + 1) & 1) * kd;
        HamArgs<F> ha{};
        ha.X = dX; ha.n = n; ha.d = (int)d; ha.k = (int)k; ha.C = cur; ha.labels = ctx->labels.as<uint32_t>(); ha.ub = ub; ha.lb = lb;
        ha.moved = ctx->ham_moved.as<F>(); ha.s = ctx->ham_s.as<F>(); ha.list = ctx->ham_list.as<uint32_t>(); ha.hs = hs; ha.state = dstate;
        trace_mark(ctx, "start");
        hamerly_bounds_kernel<F><<<bgrid, 256, 0, ctx->stream>>>(ha);
        CK(cudaGetLastError());
        ctx->launches++;
        trace_mark(ctx, "hamerly_bounds");
        RowsArgs<F> ra{};
        ra.X = dX; ra.n = n; ra.d = (int)d; ra.k = (int)k; ra.C = cur; ra.list = ctx->ham_list.as<uint32_t>(); ra.list_count = &hs->list_count;
        ra.labels = ctx->labels.as<uint32_t>(); ra.ub = ub; ra.lb = lb; ra.fx_sums = fx_sums; ra.fx_counts = fx_counts; ra.fx_scale = scale;
        ra.state = dstate; ra.n_exact = &hs->rows_exact; ra.n_changed = &hs->rows_changed;
        if constexpr (std::is_same<F, float>::value) {
          if (ham_pivot) {
            // rows whose bounds failed: decided among the centroids near their current one when those are few (lkm_rows.cuh)
            PivotArgs pv{};
            pv.X = dX; pv.k = (int)k; pv.C = cur; pv.CC = ctx->piv_cc.as<float>(); pv.list = ra.list; pv.count = ra.list_count;
            pv.labels = ra.labels; pv.list2 = ctx->piv_list2.as<uint32_t>();
            pv.count2 = reinterpret_cast<unsigned int*>(ctx->piv_list2.as<uint32_t>() + ((n + 256) & ~(uint64_t)3));
            pv.state = dstate; pv.ub = ub; pv.lb = lb; pv.fx_sums = fx_sums; pv.fx_counts = fx_counts; pv.fx_scale = scale;
            pv.n_exact = &hs->rows_exact; pv.n_changed = &hs->rows_changed;
            CK(cudaMemsetAsync(pv.count2, 0, 4, ctx->stream));
            if (d == 64) {
              centroid_dist_kernel<64><<<(unsigned)k, 128, 0, ctx->stream>>>(pv);
              rows_pivot_kernel<64, true><<<ctx->sm_count * 4, kPivThreads, 0, ctx->stream>>>(pv);
            } else {
              centroid_dist_kernel<128><<<(unsigned)k, 128, 0, ctx->stream>>>(pv);
              rows_pivot_kernel<128, true><<<ctx->sm_count * 3, kPivThreads, 0, ctx->stream>>>(pv);
            }
            CK(cudaGetLastError());
            ctx->launches += 2;
            ra.list = pv.list2; ra.list_count = pv.count2;
            trace_mark(ctx, "hamerly_rows_pivot");
          }
        }
        CKS(launch_rows<F>(ctx, ra, n));
        trace_mark(ctx, "hamerly_rows_top2");
        HamCenArgs<F> ca{};
        ca.fx_sums = fx_sums; ca.fx_counts = fx_counts; ca.inv_scale = inv_scale; ca.k = (int)k; ca.d = (int)d; ca.C_old = cur; ca.C_new = nxt;
        ca.shift_part = ctx->shift_part.as<double>(); ca.counts_out = ctx->counts_out.as<double>(); ca.state = dstate;
        hamerly_centroids_kernel<F><<<cen_grid, 256, 0, ctx->stream>>>(ca);
        CK(cudaGetLastError());
        ctx->launches++;
        CKS(moved_and_state(cur, nxt, true));
        trace_mark(ctx, "hamerly_centroids_moved");
      }
      CK(cudaMemcpyAsync(ctx->h_state, dstate, sizeof(LloydState), cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      if (ctx->h_state->done || it >= max_iter) break;
    }
    CK(cudaEventRecord(ctx->ev1, ctx->stream));
    HamState hh;
    CK(cudaMemcpyAsync(&hh, hs, sizeof(HamState), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaEventSynchronize(ctx->ev1));
    CK(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    out.ms += ms;
    out.n_iter = ctx->h_state->n_iter;
    out.shift = ctx->h_state->shift;
    out.final_buf = (int)(out.n_iter & 1);
    out.fallback_rows += hh.rows_exact;
    ctx->ham_rows_scanned = hh.rows_scanned;
    ctx->ham_rows_tightened = hh.rows_tightened;
  }
  CKS(inertia_pass<F>(ctx, dC2 + (size_t)out.final_buf * kd, out.inertia));
  trace_mark(ctx, "inertia_pass");
  trace_report(ctx);
  ctx->kept_labels_n = n;
  return LKM_OK;
}

// ---- seeding ---------------------------------------------------------------------
template <typename F>
static lkm_status gather_rows(lkm_ctx* ctx, const F* dX, uint64_t d, const std::vector<long long>& idx, F* dOut) {
  const size_t k = idx.size();
  CK(ctx->idxbuf.ensure(k * 8));
  CK(cudaMemcpyAsync(ctx->idxbuf.p, idx.data(), k * 8, cudaMemcpyHostToDevice, ctx->stream));
  gather_rows_kernel<F><<<(int)std::min<size_t>(k, 1024), 128, 0, ctx->stream>>>(dX, (int)d, ctx->idxbuf.as<long long>(),
                                                                                  (int)k, dOut);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(ctx->stream));  // idx (host vector) must outlive the copy
  ctx->launches++;
  return LKM_OK;
}

// WeightedIndex::new(w).sample(rng) on device-resident weights (rand 0.8.5):
// returns the picked index, or -1 where WeightedIndex::new errors.
template <typename F>
static lkm_status weighted_pick(lkm_ctx* ctx, const F* dW, uint64_t n, HostRng& rng, int pick_mode, long long& picked) {
  // AUTO: the bit-exact sequential prefix (one GPU thread walking n weights per pick) up to 2^22 weights, the parallel scan above
  if (pick_mode == LKM_PICK_AUTO) pick_mode = n <= (1ull << 22) ? LKM_PICK_EXACT : LKM_PICK_FAST;
  if (n == 0) { picked = -1; return LKM_OK; }
  if (pick_mode == LKM_PICK_EXACT) {
    CK(ctx->cum.ensure(std::max<uint64_t>(n, 2) * sizeof(F)));
    CK(ctx->scratch.ensure(64));
    seq_prefix_kernel<F><<<1, 256, 0, ctx->stream>>>(dW, n, ctx->cum.as<F>(), ctx->scratch.as<F>(), (F)0, n - 1);
    CK(cudaGetLastError());
    F tot[2];
    CK(cudaMemcpyAsync(tot, ctx->scratch.p, 2 * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->launches++;
    if (tot[1] != (F)0 || !(tot[0] >= (F)0)) { picked = -1; return LKM_OK; }  // InvalidWeight
    if (tot[0] == (F)0) { picked = -1; return LKM_OK; }                       // AllWeightsZero
    const F chosen = weighted_threshold<F>(rng, tot[0]);
    unsigned long long* dcnt = reinterpret_cast<unsigned long long*>(ctx->scratch.as<char>() + 32);
    CK(cudaMemsetAsync(dcnt, 0, 8, ctx->stream));
    if (n > 1) {
      count_le_kernel<F><<<(int)std::min<uint64_t>((n + 255) / 256, 2048), 256, 0, ctx->stream>>>(ctx->cum.as<F>(), n - 1,
                                                                                                   chosen, dcnt);
      CK(cudaGetLastError());
      ctx->launches++;
    }
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, dcnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    picked = (long long)cnt;
    return LKM_OK;
  }
  // FAST: fixed-order f64 block sums + single-CTA walk, same uniform draw
  const uint64_t nb = (n + kPickBlock - 1) / kPickBlock;
  CK(ctx->bsum.ensure(nb * 8));
  CK(ctx->scratch.ensure(64));
  block_sums_kernel<F><<<(int)std::min<uint64_t>(nb, 4096), 256, 0, ctx->stream>>>(dW, n, ctx->bsum.as<double>());
  CK(cudaGetLastError());
  // total (host walks the block sums: nb doubles, tiny next to n)
  std::vector<double> hb(nb);
  CK(cudaMemcpyAsync(hb.data(), ctx->bsum.p, nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  // fixed order shared with fast_pick_device_kernel: group sums, then the groups (pick_group_sum / pick_locate)
  const uint64_t ng = (nb + kPickGroup - 1) / kPickGroup;
  std::vector<double> hg(ng);
  for (uint64_t g = 0; g < ng; ++g) hg[g] = pick_group_sum(hb.data(), nb, g);
  double total = 0.0;
  for (uint64_t g = 0; g < ng; ++g) total += hg[g];
  ctx->launches++;
  if (!(total > 0.0)) { picked = -1; return LKM_OK; }
  // Uniform<F>::new(0, total).sample: value0_1 * scale with scale = (F)total
  const F u01 = unit<F>(rng);
  const double thr = (double)(u01 * (F)total);
  unsigned long long fb = 0;
  double fp = 0.0;
  pick_locate(hb.data(), nb, hg.data(), ng, thr, fb, fp);
  long long* didx = reinterpret_cast<long long*>(ctx->scratch.as<char>() + 48);
  fast_pick_kernel<F><<<1, 256, 0, ctx->stream>>>(dW, n, fb, fp, thr, didx);
  CK(cudaGetLastError());
  ctx->launches++;
  CK(cudaMemcpyAsync(&picked, didx, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (picked < 0) picked = (long long)n - 1;
  return LKM_OK;
}

// One distance pass of k-means++ over the local rows: dists = min(dists, ||x - seeds[c_index]||^2), wdists = weights * dists.
// f32 rows of 32 / 64 / 128 features take the pruned pass of lkm_seedpass.cuh (rows the new seed cannot reach are not read);
// everything else the assignment kernels.  `seeds` holds the c_index + 1 seeds chosen so far.
template <typename F>
static lkm_status seed_distance_pass(lkm_ctx* ctx, const F* dX, uint64_t n, uint64_t d, const F* seeds, uint64_t c_index, F* dD,
                                     const F* dW, F* dWD) {
  if (n == 0) return LKM_OK;
  if constexpr (std::is_same<F, float>::value) {
    if (ctx->seed_pass && ctx->seed_prune && ctx->force_path != 0 && (d == 32 || d == 64 || d == 128) && c_index < (1u << 30) &&
        (reinterpret_cast<uintptr_t>(dX) & 15) == 0 && (reinterpret_cast<uintptr_t>(seeds) & 15) == 0) {
      CK(ctx->seed_nearest.ensure(n * 4));
      CK(ctx->seed_sd.ensure((c_index + 1) * 4));
      SeedPruneArgs sp{};
      sp.X = dX; sp.n = n; sp.seeds = seeds; sp.n_seeds = (int)(c_index + 1); sp.seed_dist = ctx->seed_sd.as<float>();
      sp.dists = dD; sp.nearest = ctx->seed_nearest.as<uint32_t>(); sp.weights = dW; sp.wdists = dWD;
      const size_t smem = d == 32 ? seedprune_smem_bytes<32>() : (d == 64 ? seedprune_smem_bytes<64>() : seedprune_smem_bytes<128>());
      const void* fn = d == 32 ? (const void*)seed_pass_pruned_kernel<32>
                               : (d == 64 ? (const void*)seed_pass_pruned_kernel<64> : (const void*)seed_pass_pruned_kernel<128>);
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   // per launch: see launch_lloyd_t
      CK(cudaFuncSetAttribute(fn, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
      const int occ = std::max(1, std::min(8, (int)((ctx->smem_optin + 1024) / (smem + 2048))));
      const uint64_t tiles = (n + kSqRows - 1) / kSqRows;
      const int grid = (int)std::max<uint64_t>(1, std::min<uint64_t>(tiles, (uint64_t)ctx->sm_count * occ));
      if (c_index > 0) {
        const unsigned sgrid = (unsigned)((c_index * 32 + 255) / 256);
        if (d == 32) seed_dist_kernel<32><<<sgrid, 256, 0, ctx->stream>>>(sp);
        else if (d == 64) seed_dist_kernel<64><<<sgrid, 256, 0, ctx->stream>>>(sp);
        else seed_dist_kernel<128><<<sgrid, 256, 0, ctx->stream>>>(sp);
        ctx->launches++;
      }
      if (d == 32) seed_pass_pruned_kernel<32><<<grid, kSqRows, smem, ctx->stream>>>(sp);
      else if (d == 64) seed_pass_pruned_kernel<64><<<grid, kSqRows, smem, ctx->stream>>>(sp);
      else seed_pass_pruned_kernel<128><<<grid, kSqRows, smem, ctx->stream>>>(sp);
      CK(cudaGetLastError());
      ctx->launches++;
      return LKM_OK;
    }
  }
  return run_assign<F>(ctx, dX, n, d, seeds + c_index * d, 1, nullptr, dD, c_index > 0 ? 1 : 0, dW, dWD);
}

// weighted_k_means_plusplus (KM/init.rs:118-158) over device rows dX (n x d)
// with optional device weights (null = ones).  The min-distance array is kept
// as a running minimum over the centroids chosen so far — the same value the
// reference recomputes from scratch each round (:138-143), since a minimum
// does not depend on evaluation order.
template <typename F>
static lkm_status weighted_kmeanspp(lkm_ctx* ctx, const F* dX, uint64_t n, uint64_t d, const F* dW, uint64_t k,
                                    HostRng& rng, int pick_mode, F* dC) {
  CK(ctx->dists.ensure(std::max<uint64_t>(n, 1) * sizeof(F)));
  CK(ctx->wdists.ensure(std::max<uint64_t>(n, 1) * sizeof(F)));
  F* dD = ctx->dists.as<F>();
  F* dWD = ctx->wdists.as<F>();
  const F* first_w = dW;
  if (!dW) {
    // weights = ones: materialise them once so that the first pick walks the
    // same F prefix sum 1, 2, 3, ... WeightedIndex::new builds
    CK(ctx->weights.ensure(std::max<uint64_t>(n, 1) * sizeof(F)));
    fill_kernel<F><<<ctx->sm_count * 4, 256, 0, ctx->stream>>>(ctx->weights.as<F>(), n, (F)1);
    CK(cudaGetLastError());
    ctx->launches++;
    first_w = ctx->weights.as<F>();
  } else {
    // assert_ne!(weights.sum(), 0) (KM/init.rs:127): all weights are >= 0 here, so
    // the sum is zero iff every weight is zero; the pick below reports that case.
  }
  long long idx = -1;
  CKS(weighted_pick<F>(ctx, first_w, n, rng, pick_mode, idx));
  if (idx < 0) return fail(ctx, dW ? LKM_ERR_ZERO_WEIGHTS : LKM_ERR_BAD_WEIGHTS, "invalid weights for k-means++");
  CK(cudaMemcpyAsync(dC, dX + (size_t)idx * d, d * sizeof(F), cudaMemcpyDeviceToDevice, ctx->stream));
  int mode = pick_mode;
  if (mode == LKM_PICK_AUTO) mode = n <= (1ull << 22) ? LKM_PICK_EXACT : LKM_PICK_FAST;
  if (mode == LKM_PICK_FAST && ctx->device_picks && n > 0 && (n + kPickBlock - 1) / kPickBlock <= (uint64_t)kPickGroup * kPickMaxGroups) {
    // FAST: nothing the host draws depends on the device (one uniform per pick), so the k - 1 picks are queued without a
    // host round trip: distance update -> block sums -> fast_pick_device_kernel (total, threshold, walk, row copy)
    const uint64_t nb = (n + kPickBlock - 1) / kPickBlock;
    CK(ctx->bsum.ensure(nb * 8));
    CK(ctx->scratch.ensure(64));
    long long* didx = reinterpret_cast<long long*>(ctx->scratch.as<char>() + 48);
    for (uint64_t c = 1; c < k; ++c) {
      CKS(seed_distance_pass<F>(ctx, dX, n, d, dC, c - 1, dD, dW, dWD));
      block_sums_kernel<F><<<(int)std::min<uint64_t>(nb, 4096), 256, 0, ctx->stream>>>(dWD, n, ctx->bsum.as<double>());
      const F u01 = unit<F>(rng);
      fast_pick_device_kernel<F><<<1, 256, 0, ctx->stream>>>(dWD, n, ctx->bsum.as<double>(), nb, u01, dX, (int)d, dC + c * d, didx);
      CK(cudaGetLastError());
      ctx->launches += 2;
    }
    CK(cudaStreamSynchronize(ctx->stream));
    return LKM_OK;
  }
  for (uint64_t c = 1; c < k; ++c) {
    CKS(seed_distance_pass<F>(ctx, dX, n, d, dC, c - 1, dD, dW, dWD));
    CKS(weighted_pick<F>(ctx, dWD, n, rng, pick_mode, idx));
    if (idx < 0) idx = 0;  // .unwrap_or(0) (KM/init.rs:149-152)
    CK(cudaMemcpyAsync(dC + c * d, dX + (size_t)idx * d, d * sizeof(F), cudaMemcpyDeviceToDevice, ctx->stream));
  }
  return LKM_OK;
}

// k_means_para (KM/init.rs:181-234).  The reference's per-row Bernoulli draws
// come from rayon-split Xoshiro streams and are not reproducible even there;
// here each row draws from a counter-based generator keyed by (round seed, row).
template <typename F>
static lkm_status kmeans_para(lkm_ctx* ctx, const F* dX, uint64_t n, uint64_t d, uint64_t k, HostRng& rng, int pick_mode,
                              F* dC) {
  const uint64_t n_rounds = 8, cap = k * n_rounds;
  CK(ctx->tmpc.ensure(cap * d * sizeof(F)));
  F* dCand = ctx->tmpc.as<F>();
  CK(ctx->dists.ensure(std::max<uint64_t>(n, 1) * sizeof(F)));
  F* dD = ctx->dists.as<F>();
  const uint64_t first = range_incl_u64(rng, 0, n - 1);
  CK(cudaMemcpyAsync(dCand, dX + first * d, d * sizeof(F), cudaMemcpyDeviceToDevice, ctx->stream));
  uint64_t n_cand = 1, scanned = 0;
  const unsigned int sel_cap = (unsigned int)std::min<uint64_t>(std::max<uint64_t>(16 * k + 1024, 4096), 1u << 24);
  CK(ctx->idxbuf.ensure((size_t)sel_cap * 8 + 64));
  CK(ctx->part_inertia.ensure(1024 * 8));
  bool full = false;
  std::vector<unsigned long long> sel;
  std::vector<double> parts(1024);
  for (uint64_t r = 0; r < n_rounds && !full; ++r) {
    // running min against the candidates added since the last round
    if (n_cand > scanned) {
      const uint64_t fresh = n_cand - scanned;
      if (ctx->para_rows && fresh >= 8 && ctx->force_path != 0) {
        // thread = row, the row in registers, FMA scores with the certificate, exact distance of the winner: the same bits
        // as the exact scan at a fraction of its cost (rows_top2_kernel over ALL rows)
        RowsArgs<F> ra{};
        ra.X = dX; ra.n = n; ra.d = (int)d; ra.k = (int)fresh; ra.C = dCand + scanned * d; ra.all_rows = n;
        ra.best_dist = dD; ra.min_combine = scanned > 0 ? 1 : 0;
        CKS(launch_rows<F>(ctx, ra, n));
      } else {
        CKS(run_assign<F>(ctx, dX, n, d, dCand + scanned * d, fresh, nullptr, dD, scanned > 0 ? 1 : 0, nullptr, nullptr));
      }
    }
    scanned = n_cand;
    const uint64_t seed = range_incl_u64(rng, 0, 0xFFFFFFFFFFFFFFFEULL);
    sum_f64_kernel<F><<<1024, 256, 0, ctx->stream>>>(dD, n, ctx->part_inertia.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(parts.data(), ctx->part_inertia.p, 1024 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    double cost = 0.0;
    for (double v : parts) cost += v;
    unsigned int* dcur = reinterpret_cast<unsigned int*>(ctx->idxbuf.as<char>() + (size_t)sel_cap * 8);
    CK(cudaMemsetAsync(dcur, 0, 4, ctx->stream));
    bernoulli_select_kernel<F><<<(int)std::min<uint64_t>((n + 255) / 256, 4096), 256, 0, ctx->stream>>>(
        dD, n, (F)k, (F)cost, seed, 0ULL, ctx->idxbuf.as<unsigned long long>(), sel_cap, dcur);
    CK(cudaGetLastError());
    ctx->launches += 2;
    unsigned int got = 0;
    CK(cudaMemcpyAsync(&got, dcur, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    got = std::min(got, sel_cap);
    sel.resize(got);
    if (got) CK(cudaMemcpy(sel.data(), ctx->idxbuf.p, (size_t)got * 8, cudaMemcpyDeviceToHost));
    std::sort(sel.begin(), sel.end());  // the reference collects in ascending row order
    {
      // one gather launch for the round's candidates (not one copy per row)
      std::vector<long long> take;
      for (unsigned long long idx : sel) {
        take.push_back((long long)idx);
        if (n_cand + take.size() >= cap) { full = true; break; }
      }
      if (!take.empty()) CKS(gather_rows<F>(ctx, dX, d, take, dCand + n_cand * d));
      n_cand += take.size();
    }
  }
  // weights = cluster_membership_counts (KM/init.rs:273-285): histogram of the
  // nearest candidate over all rows
  bool counted = false;
  if constexpr (std::is_same<F, float>::value) {
    // f32 rows of 64 / 128 features on the resident matrix: the histogram is the cluster-count output of ONE Lloyd iteration
    // with the candidates as centroids, i.e. the CTA-pair tensor-core assignment (labels certified against the exact scan)
    // instead of n x n_cand exact CUDA-core distances (cfg4 shard: 12.5M x 8192 x 64)
    if (ctx->para_tc && ctx->tc5_enabled && ctx->upd_variant != 0 && (ctx->force_path == -1 || ctx->force_path == 2) &&
        (d == 64 || d == 128) && n_cand * d >= 2048 && n_cand <= (uint64_t)kT5K * kT5MaxChunks &&
        dX == reinterpret_cast<const F*>(ctx->X) && n == ctx->n && ctx->world == 1) {
      CK(ctx->para_c2.ensure(2 * n_cand * d * sizeof(F)));
      CK(cudaMemcpyAsync(ctx->para_c2.p, dCand, n_cand * d * sizeof(F), cudaMemcpyDeviceToDevice, ctx->stream));
      LloydOut lo;
      // candidates are data rows, several per blob: their score gaps are far below the one-product screening margin, so
      // the 3xTF32 variant is pinned for this call (the screening variant would re-scan nearly every row exactly)
      const int passes_saved = ctx->tc5_passes_forced;
      ctx->tc5_passes_forced = 3;
      const lkm_status rs = run_lloyd<F>(ctx, ctx->para_c2.as<F>(), n_cand, 1, 0.0, lo);
      ctx->tc5_passes_forced = passes_saved;
      if (rs != LKM_OK) return rs;
      std::vector<double> hc(n_cand);
      CK(cudaMemcpyAsync(hc.data(), ctx->counts_out.p, n_cand * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      std::vector<F> hw(n_cand);
      for (uint64_t c = 0; c < n_cand; ++c) hw[c] = (F)hc[c];
      CK(ctx->tmpx.ensure(n_cand * sizeof(F) + n_cand * 4 + 64));
      CK(cudaMemcpyAsync(ctx->tmpx.p, hw.data(), n_cand * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      counted = true;
    }
  }
  if (!counted) {
    CK(ctx->labels.ensure(std::max<uint64_t>(n, 1) * 4));
    CKS(run_assign<F>(ctx, dX, n, d, dCand, n_cand, ctx->labels.as<uint32_t>(), nullptr, 0, nullptr, nullptr));
    CK(ctx->tmpx.ensure(n_cand * sizeof(F) + n_cand * 4 + 64));
    unsigned int* dhist = reinterpret_cast<unsigned int*>(ctx->tmpx.as<char>() + ((n_cand * sizeof(F) + 15) & ~(size_t)15));
    CK(cudaMemsetAsync(dhist, 0, n_cand * 4, ctx->stream));
    label_hist_kernel<<<(int)std::min<uint64_t>((n + 255) / 256, 4096), 256, 0, ctx->stream>>>(ctx->labels.as<uint32_t>(), n, dhist);
    CK(cudaGetLastError());
    u32_to_float_kernel<F><<<(int)((n_cand + 255) / 256), 256, 0, ctx->stream>>>(dhist, (int)n_cand, ctx->tmpx.as<F>());
    CK(cudaGetLastError());
    ctx->launches += 2;
  }
  // weighted k-means++ over the candidates; its scratch arrays are sized for
  // n_cand <= n rows, and the candidate weights live in tmpx
  return weighted_kmeanspp<F>(ctx, dCand, n_cand, d, ctx->tmpx.as<F>(), k, rng, pick_mode, dC);
}

// ---- row-sharded seeding (world > 1): every rank draws the same random numbers from its copy of
// the rng; row indices are GLOBAL; the owner of a picked row publishes it through an all-gather ----
static lkm_status comm_allgather_host(lkm_ctx* ctx, const void* src, size_t bytes, std::vector<unsigned char>& out) {
  const size_t w = (size_t)ctx->world;
  out.resize(bytes * w);
  if (ctx->group != nullptr) {
    LocalGroup* G = ctx->group;
    G->slot[ctx->rank].assign(static_cast<const unsigned char*>(src), static_cast<const unsigned char*>(src) + bytes);
    if (!G->barrier()) return fail(ctx, LKM_ERR_NCCL, "another rank of the device group failed");
    for (size_t r = 0; r < w; ++r) {
      if (G->slot[r].size() != bytes) { G->abort(); return fail(ctx, LKM_ERR_NCCL, "ranks of the device group disagree on a message size"); }
      std::memcpy(out.data() + r * bytes, G->slot[r].data(), bytes);
    }
    if (!G->barrier()) return fail(ctx, LKM_ERR_NCCL, "another rank of the device group failed");   // the slots may be reused
    return LKM_OK;
  }
  CK(ctx->gather2.ensure(bytes * w));
  unsigned char* g = ctx->gather2.as<unsigned char>();
  CK(cudaMemcpyAsync(g + (size_t)ctx->rank * bytes, src, bytes, cudaMemcpyHostToDevice, ctx->stream));
  CKN(g_nccl.AllGather(g + (size_t)ctx->rank * bytes, g, bytes, kNcclInt8, ctx->comm, ctx->stream));
  CK(cudaMemcpyAsync(out.data(), g, bytes * w, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LKM_OK;
}

struct ShardMap {
  std::vector<uint64_t> lo, n;
  int owner(uint64_t g) const {
    for (size_t r = 0; r < lo.size(); ++r)
      if (g >= lo[r] && g < lo[r] + n[r]) return (int)r;
    return -1;
  }
};
static lkm_status shard_map(lkm_ctx* ctx, ShardMap& m) {
  uint64_t mine[2] = {ctx->row_offset, ctx->n};
  std::vector<unsigned char> all;
  CKS(comm_allgather_host(ctx, mine, sizeof(mine), all));
  m.lo.resize(ctx->world); m.n.resize(ctx->world);
  for (int r = 0; r < ctx->world; ++r) {
    uint64_t v[2];
    std::memcpy(v, all.data() + (size_t)r * 16, 16);
    m.lo[r] = v[0]; m.n[r] = v[1];
  }
  return LKM_OK;
}

// every rank ends with row `g` (global index) of the sharded matrix in dDst (device, d values)
template <typename F>
static lkm_status fetch_global_row(lkm_ctx* ctx, const ShardMap& sm, uint64_t g, F* dDst) {
  const uint64_t d = ctx->d;
  const int own = sm.owner(g);
  if (own < 0) return fail(ctx, LKM_ERR_INVALID_ARG, "picked row outside every shard");
  std::vector<F> row(d, (F)0);
  if (own == ctx->rank) {
    CK(cudaMemcpyAsync(row.data(), reinterpret_cast<const F*>(ctx->X) + (g - ctx->row_offset) * d, d * sizeof(F),
                       cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  std::vector<unsigned char> all;
  CKS(comm_allgather_host(ctx, row.data(), d * sizeof(F), all));
  CK(cudaMemcpyAsync(dDst, all.data() + (size_t)own * d * sizeof(F), d * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LKM_OK;
}

// WeightedIndex::new(w).sample(rng) over the row-sharded weight vector; picked = GLOBAL index or -1
template <typename F>
static lkm_status weighted_pick_sharded(lkm_ctx* ctx, const ShardMap& sm, const F* dW, HostRng& rng, int pick_mode,
                                        long long& picked) {
  {
    uint64_t n_all = 0;
    for (uint64_t v : sm.n) n_all += v;
    if (pick_mode == LKM_PICK_AUTO) pick_mode = n_all <= (1ull << 22) ? LKM_PICK_EXACT : LKM_PICK_FAST;
  }
  const uint64_t n = ctx->n;
  const int world = ctx->world, rank = ctx->rank;
  std::vector<unsigned char> all;
  if (pick_mode == LKM_PICK_EXACT) {
    // the sequential F prefix runs rank after rank, each continuing from the carry of the one before
    CK(ctx->cum.ensure(std::max<uint64_t>(n, 2) * sizeof(F)));
    CK(ctx->scratch.ensure(64));
    F carry = (F)0;
    bool bad = false;
    for (int s = 0; s < world; ++s) {
      F tot[2] = {carry, (F)0};
      if (s == rank && n > 0) {
        const uint64_t n_cum = (rank == world - 1) ? n - 1 : n;
        seq_prefix_kernel<F><<<1, 256, 0, ctx->stream>>>(dW, n, ctx->cum.as<F>(), ctx->scratch.as<F>(), carry, n_cum);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(tot, ctx->scratch.p, 2 * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        ctx->launches++;
      }
      CKS(comm_allgather_host(ctx, tot, sizeof(tot), all));
      F got[2];
      std::memcpy(got, all.data() + (size_t)s * sizeof(tot), sizeof(tot));
      carry = got[0];
      bad = bad || got[1] != (F)0;
    }
    if (bad || !(carry >= (F)0) || carry == (F)0) { picked = -1; return LKM_OK; }
    const F chosen = weighted_threshold<F>(rng, carry);
    unsigned long long* dcnt = reinterpret_cast<unsigned long long*>(ctx->scratch.as<char>() + 32);
    CK(cudaMemsetAsync(dcnt, 0, 8, ctx->stream));
    const uint64_t n_cum = (rank == world - 1) ? (n ? n - 1 : 0) : n;
    if (n_cum) {
      count_le_kernel<F><<<(int)std::min<uint64_t>((n_cum + 255) / 256, 2048), 256, 0, ctx->stream>>>(ctx->cum.as<F>(), n_cum, chosen, dcnt);
      CK(cudaGetLastError());
      ctx->launches++;
    }
    unsigned long long cnt = 0;
    CK(cudaMemcpyAsync(&cnt, dcnt, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CKS(comm_allgather_host(ctx, &cnt, 8, all));
    unsigned long long total_cnt = 0;
    for (int r = 0; r < world; ++r) { unsigned long long v; std::memcpy(&v, all.data() + (size_t)r * 8, 8); total_cnt += v; }
    picked = (long long)total_cnt;
    return LKM_OK;
  }
  // FAST: per-rank fixed-order f64 block sums; rank totals combined in rank order
  const uint64_t nb = (n + kPickBlock - 1) / kPickBlock;
  std::vector<double> hb(std::max<uint64_t>(nb, 1), 0.0);
  CK(ctx->scratch.ensure(64));
  if (nb) {
    CK(ctx->bsum.ensure(nb * 8));
    block_sums_kernel<F><<<(int)std::min<uint64_t>(nb, 4096), 256, 0, ctx->stream>>>(dW, n, ctx->bsum.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(hb.data(), ctx->bsum.p, nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->launches++;
  }
  double local = 0.0;
  for (uint64_t b = 0; b < nb; ++b) local += hb[b];
  CKS(comm_allgather_host(ctx, &local, 8, all));
  std::vector<double> tot(world);
  for (int r = 0; r < world; ++r) std::memcpy(&tot[r], all.data() + (size_t)r * 8, 8);
  double total = 0.0, before = 0.0;
  for (int r = 0; r < world; ++r) { if (r == rank) before = total; total += tot[r]; }
  if (!(total > 0.0)) { picked = -1; return LKM_OK; }
  const F u01 = unit<F>(rng);
  const double thr = (double)(u01 * (F)total);
  long long mine = -1;
  const bool last_nonempty = [&] { for (int r = world - 1; r >= 0; --r) if (sm.n[r]) return r == rank; return false; }();
  if (n && ((thr >= before && thr < before + local) || (last_nonempty && thr >= before + local))) {
    double run = before;
    uint64_t fb = nb - 1;
    double fp = 0.0;
    bool hit = false;
    for (uint64_t b = 0; b < nb; ++b) {
      const double nxt = run + hb[b];
      if (nxt > thr) { fb = b; fp = run; hit = true; break; }
      run = nxt;
    }
    if (!hit) fp = run - hb[nb - 1];
    long long* didx = reinterpret_cast<long long*>(ctx->scratch.as<char>() + 48);
    fast_pick_kernel<F><<<1, 256, 0, ctx->stream>>>(dW, n, fb, fp, thr, didx);
    CK(cudaGetLastError());
    ctx->launches++;
    CK(cudaMemcpyAsync(&mine, didx, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    mine += (long long)ctx->row_offset;
  }
  CKS(comm_allgather_host(ctx, &mine, 8, all));
  picked = -1;
  for (int r = 0; r < world; ++r) { long long v; std::memcpy(&v, all.data() + (size_t)r * 8, 8); if (v >= 0 && picked < 0) picked = v; }
  return LKM_OK;
}

// k_means_plusplus (weights = ones) over the sharded matrix
template <typename F>
static lkm_status kmeanspp_sharded(lkm_ctx* ctx, const ShardMap& sm, uint64_t k, HostRng& rng, int pick_mode, F* dC) {
  const uint64_t n = ctx->n, d = ctx->d;
  const F* dX = reinterpret_cast<const F*>(ctx->X);
  CK(ctx->dists.ensure(std::max<uint64_t>(n, 1) * sizeof(F)));
  CK(ctx->wdists.ensure(std::max<uint64_t>(n, 1) * sizeof(F)));
  CK(ctx->weights.ensure(std::max<uint64_t>(n, 1) * sizeof(F)));
  std::vector<F> ones(std::min<uint64_t>(std::max<uint64_t>(n, 1), 1 << 20), (F)1);
  for (uint64_t o = 0; o < n; o += ones.size())
    CK(cudaMemcpyAsync(ctx->weights.as<F>() + o, ones.data(), std::min<uint64_t>(ones.size(), n - o) * sizeof(F),
                       cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  long long idx = -1;
  CKS(weighted_pick_sharded<F>(ctx, sm, ctx->weights.as<F>(), rng, pick_mode, idx));
  if (idx < 0) return fail(ctx, LKM_ERR_BAD_WEIGHTS, "invalid weights for k-means++");
  CKS(fetch_global_row<F>(ctx, sm, (uint64_t)idx, dC));
  for (uint64_t c = 1; c < k; ++c) {
    CKS(seed_distance_pass<F>(ctx, dX, n, d, dC, c - 1, ctx->dists.as<F>(), nullptr, ctx->wdists.as<F>()));
    CKS(weighted_pick_sharded<F>(ctx, sm, ctx->wdists.as<F>(), rng, pick_mode, idx));
    if (idx < 0) idx = 0;
    CKS(fetch_global_row<F>(ctx, sm, (uint64_t)idx, dC + c * d));
  }
  return LKM_OK;
}

template <typename F>
static lkm_status random_init_sharded(lkm_ctx* ctx, const ShardMap& sm, uint64_t k, HostRng& rng, F* dC) {
  const uint64_t n_glob = ctx->n_global;
  std::vector<uint64_t> idx;
  if (!index_sample(rng, n_glob, k, idx)) return fail(ctx, LKM_ERR_SAMPLE_AMOUNT, "Random init: n_clusters > n_samples");
  for (uint64_t c = 0; c < k; ++c) CKS(fetch_global_row<F>(ctx, sm, idx[c], dC + c * ctx->d));
  return LKM_OK;
}

template <typename F>
static lkm_status kmeans_para_sharded(lkm_ctx* ctx, const ShardMap& sm, uint64_t k, HostRng& rng, int pick_mode, F* dC) {
  const uint64_t n = ctx->n, d = ctx->d, n_glob = ctx->n_global;
  const int world = ctx->world;
  const F* dX = reinterpret_cast<const F*>(ctx->X);
  const uint64_t n_rounds = 8, cap = k * n_rounds;
  CK(ctx->tmpc.ensure(cap * d * sizeof(F)));
  F* dCand = ctx->tmpc.as<F>();
  CK(ctx->dists.ensure(std::max<uint64_t>(n, 1) * sizeof(F)));
  F* dD = ctx->dists.as<F>();
  const uint64_t first = range_incl_u64(rng, 0, n_glob - 1);
  CKS(fetch_global_row<F>(ctx, sm, first, dCand));
  uint64_t n_cand = 1, scanned = 0;
  const unsigned int sel_cap = (unsigned int)std::min<uint64_t>(std::max<uint64_t>(16 * k + 1024, 4096), 1u << 22);
  CK(ctx->idxbuf.ensure((size_t)sel_cap * 8 + 64));
  CK(ctx->part_inertia.ensure(1024 * 8));
  std::vector<double> parts(1024);
  std::vector<unsigned char> all;
  bool full = false;
  for (uint64_t r = 0; r < n_rounds && !full; ++r) {
    if (n && n_cand > scanned)
      CKS(run_assign<F>(ctx, dX, n, d, dCand + scanned * d, n_cand - scanned, nullptr, dD, scanned > 0 ? 1 : 0, nullptr, nullptr));
    scanned = n_cand;
    const uint64_t seed = range_incl_u64(rng, 0, 0xFFFFFFFFFFFFFFFEULL);
    double local_cost = 0.0;
    if (n) {
      sum_f64_kernel<F><<<1024, 256, 0, ctx->stream>>>(dD, n, ctx->part_inertia.as<double>());
      CK(cudaGetLastError());
      CK(cudaMemcpyAsync(parts.data(), ctx->part_inertia.p, 1024 * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      for (double v : parts) local_cost += v;
      ctx->launches++;
    }
    CKS(comm_allgather_host(ctx, &local_cost, 8, all));
    double cost = 0.0;
    for (int q = 0; q < world; ++q) { double v; std::memcpy(&v, all.data() + (size_t)q * 8, 8); cost += v; }
    std::vector<unsigned long long> sel;
    if (n) {
      unsigned int* dcur = reinterpret_cast<unsigned int*>(ctx->idxbuf.as<char>() + (size_t)sel_cap * 8);
      CK(cudaMemsetAsync(dcur, 0, 4, ctx->stream));
      bernoulli_select_kernel<F><<<(int)std::min<uint64_t>((n + 255) / 256, 4096), 256, 0, ctx->stream>>>(
          dD, n, (F)k, (F)cost, seed, ctx->row_offset, ctx->idxbuf.as<unsigned long long>(), sel_cap, dcur);
      CK(cudaGetLastError());
      ctx->launches++;
      unsigned int got = 0;
      CK(cudaMemcpyAsync(&got, dcur, 4, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      got = std::min(got, sel_cap);
      sel.resize(got);
      if (got) CK(cudaMemcpy(sel.data(), ctx->idxbuf.p, (size_t)got * 8, cudaMemcpyDeviceToHost));
      std::sort(sel.begin(), sel.end());
    }
    // publish the selected rows: counts first, then padded (index, row) records, merged in rank order
    // (shards are contiguous and ordered by rank, so the concatenation is globally ascending)
    uint64_t my_cnt = sel.size();
    CKS(comm_allgather_host(ctx, &my_cnt, 8, all));
    std::vector<uint64_t> cnts(world);
    uint64_t max_cnt = 0;
    for (int q = 0; q < world; ++q) { std::memcpy(&cnts[q], all.data() + (size_t)q * 8, 8); max_cnt = std::max(max_cnt, cnts[q]); }
    if (max_cnt == 0) continue;
    const size_t rec = d * sizeof(F);
    std::vector<unsigned char> mine(max_cnt * rec, 0);
    for (uint64_t i = 0; i < my_cnt; ++i)
      CK(cudaMemcpyAsync(mine.data() + i * rec, dX + (sel[i] - ctx->row_offset) * d, rec, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CKS(comm_allgather_host(ctx, mine.data(), mine.size(), all));
    for (int q = 0; q < world && !full; ++q)
      for (uint64_t i = 0; i < cnts[q]; ++i) {
        CK(cudaMemcpyAsync(dCand + n_cand * d, all.data() + ((size_t)q * max_cnt + i) * rec, rec, cudaMemcpyHostToDevice, ctx->stream));
        n_cand += 1;
        if (n_cand >= cap) { full = true; break; }
      }
    CK(cudaStreamSynchronize(ctx->stream));
  }
  // weights = global histogram of the nearest candidate
  std::vector<double> hist(n_cand, 0.0);
  if (n) {
    CK(ctx->labels.ensure(n * 4));
    CKS(run_assign<F>(ctx, dX, n, d, dCand, n_cand, ctx->labels.as<uint32_t>(), nullptr, 0, nullptr, nullptr));
    CK(ctx->tmpx.ensure(n_cand * 4 + 64));
    unsigned int* dhist = ctx->tmpx.as<unsigned int>();
    CK(cudaMemsetAsync(dhist, 0, n_cand * 4, ctx->stream));
    label_hist_kernel<<<(int)std::min<uint64_t>((n + 255) / 256, 4096), 256, 0, ctx->stream>>>(ctx->labels.as<uint32_t>(), n, dhist);
    CK(cudaGetLastError());
    ctx->launches++;
    std::vector<unsigned int> hh(n_cand);
    CK(cudaMemcpyAsync(hh.data(), dhist, n_cand * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint64_t c = 0; c < n_cand; ++c) hist[c] = (double)hh[c];
  }
  CKS(comm_allgather_host(ctx, hist.data(), n_cand * 8, all));
  std::vector<F> w(n_cand, (F)0);
  for (uint64_t c = 0; c < n_cand; ++c) {
    double s = 0.0;
    for (int q = 0; q < world; ++q) { double v; std::memcpy(&v, all.data() + ((size_t)q * n_cand + c) * 8, 8); s += v; }
    w[c] = (F)s;
  }
  CK(ctx->tmpx.ensure(n_cand * sizeof(F) + 64));
  CK(cudaMemcpyAsync(ctx->tmpx.p, w.data(), n_cand * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  // the candidates are replicated on every rank: weighted k-means++ over them is rank-local and,
  // with identical rng streams, identical everywhere
  return weighted_kmeanspp<F>(ctx, dCand, n_cand, d, ctx->tmpx.as<F>(), k, rng, pick_mode, dC);
}

// KMeansInit::run (KM/init.rs:83-101) -> device centroids dC (k x d)
template <typename F>
static lkm_status run_init(lkm_ctx* ctx, const lkm_params* p, HostRng& rng, F* dC) {
  const uint64_t k = p->n_clusters, n = ctx->n, d = ctx->d;
  const F* dX = reinterpret_cast<const F*>(ctx->X);
  if (ctx->world > 1 && p->init_method != LKM_INIT_PRECOMPUTED) {
    if (ctx->n_global == 0) return fail(ctx, LKM_ERR_BAD_WEIGHTS, "seeding on an empty dataset");
    ShardMap sm;
    CKS(shard_map(ctx, sm));
    switch (p->init_method) {
      case LKM_INIT_RANDOM: return random_init_sharded<F>(ctx, sm, k, rng, dC);
      case LKM_INIT_KMEANS_PLUSPLUS: return kmeanspp_sharded<F>(ctx, sm, k, rng, p->pick_mode, dC);
      case LKM_INIT_KMEANS_PARA: return kmeans_para_sharded<F>(ctx, sm, k, rng, p->pick_mode, dC);
      default: return fail(ctx, LKM_ERR_INVALID_ARG, "unknown init_method");
    }
  }
  switch (p->init_method) {
    case LKM_INIT_PRECOMPUTED:
      if (!p->precomputed || p->precomputed_rows != k || p->precomputed_cols != d)
        return fail(ctx, LKM_ERR_SHAPE, "Precomputed centroids must be n_clusters x n_features");
      CK(cudaMemcpyAsync(dC, p->precomputed, k * d * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      return LKM_OK;
    case LKM_INIT_RANDOM: {
      std::vector<uint64_t> idx;
      if (!index_sample(rng, n, k, idx)) return fail(ctx, LKM_ERR_SAMPLE_AMOUNT, "Random init: n_clusters > n_samples");
      std::vector<long long> li(idx.begin(), idx.end());
      return gather_rows<F>(ctx, dX, d, li, dC);
    }
    case LKM_INIT_KMEANS_PLUSPLUS:
      if (n == 0) return fail(ctx, LKM_ERR_BAD_WEIGHTS, "k-means++ on an empty dataset");
      return weighted_kmeanspp<F>(ctx, dX, n, d, nullptr, k, rng, p->pick_mode, dC);
    case LKM_INIT_KMEANS_PARA:
      if (n == 0) return fail(ctx, LKM_ERR_BAD_WEIGHTS, "k-means|| on an empty dataset");
      return kmeans_para<F>(ctx, dX, n, d, k, rng, p->pick_mode, dC);
  }
  return fail(ctx, LKM_ERR_INVALID_ARG, "unknown init_method");
}

static lkm_status check_params(lkm_ctx* ctx, const lkm_params* p, bool need_max_iter) {
  // ParamGuard::check_ref (KM/hyperparams.rs:124-136), same order
  if (p->n_clusters == 0) return fail(ctx, LKM_ERR_N_CLUSTERS, "n_clusters cannot be 0");
  if (p->n_runs == 0) return fail(ctx, LKM_ERR_N_RUNS, "n_runs cannot be 0");
  if (p->tolerance <= 0.0) return fail(ctx, LKM_ERR_TOLERANCE, "tolerance must be greater than 0");
  if (need_max_iter && p->max_n_iterations == 0) return fail(ctx, LKM_ERR_MAX_ITERATIONS, "max_n_iterations cannot be 0");
  return LKM_OK;
}

template <typename F>
static lkm_status fit_impl(lkm_ctx* ctx, const lkm_params* p, lkm_rng rng_in, F* centroids_out, F* counts_out,
                           F* inertia_out, lkm_stats* stats) {
  if (!p || !centroids_out || !counts_out || !inertia_out) return fail(ctx, LKM_ERR_INVALID_ARG, "null argument");
  CKS(check_params(ctx, p, true));
  if (ctx->elem != (int)sizeof(F) || !ctx->X) return fail(ctx, LKM_ERR_NO_DATA, "no resident data of this type");
  CK(cudaSetDevice(ctx->device));
  const uint64_t k = p->n_clusters, d = ctx->d, kd = k * d;
  const uint64_t n_glob = ctx->n_global ? ctx->n_global : ctx->n;
  // an empty SHARD is legal (its partials are zeros; every rank still takes part in each exchange); an empty data set is not
  if (n_glob == 0 || d == 0) return fail(ctx, LKM_ERR_SHAPE, "empty dataset");
  HostRng rng{rng_in};
  CK(ctx->cbuf.ensure(2 * kd * sizeof(F)));
  F* dC2 = ctx->cbuf.as<F>();
  ctx->launches = 0;
  // F tolerance exactly as the reference holds it (tolerance: F)
  const double tol = (double)(F)p->tolerance;
  F min_inertia = std::numeric_limits<F>::infinity();
  bool have = false;
  std::vector<F> best(kd);
  std::vector<double> best_cnt(k);
  const bool hamerly = p->algorithm == LKM_ALGO_HAMERLY;
  LloydOut lo;
  float ms_total = 0;
  for (uint64_t run = 0; run < p->n_runs; ++run) {
    CKS(run_init<F>(ctx, p, rng, dC2));
    if (hamerly) CKS(run_hamerly<F>(ctx, dC2, k, p->max_n_iterations, tol, lo));
    else CKS(run_lloyd<F>(ctx, dC2, k, p->max_n_iterations, tol, lo));
    ms_total += lo.ms;
    const F inertia = (F)lo.inertia;
    if (inertia < min_inertia) {  // strict '<' from +inf: NaN never wins (KM/algorithm.rs:336 / :281)
      min_inertia = inertia;
      CK(cudaMemcpyAsync(best.data(), dC2 + (size_t)lo.final_buf * kd, kd * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
      // fit_hamerly keeps the memberships of the BEST run (KM/algorithm.rs:281-286); fit_lloyd's come from the last run
      if (hamerly) CK(cudaMemcpyAsync(best_cnt.data(), ctx->counts_out.p, k * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CK(cudaStreamSynchronize(ctx->stream));
      have = true;
    }
  }
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->n_iter = lo.n_iter; stats->last_shift = lo.shift; stats->inertia_total = lo.inertia;
    stats->device_ms = ms_total; stats->kernel_launches = ctx->launches; stats->assign_path = (uint32_t)ctx->last_path; stats->update_path = (uint32_t)ctx->last_update_path;
    stats->fallback_rows = lo.fallback_rows;
  }
  if (!have) return fail(ctx, LKM_ERR_INERTIA, "Fitting failed: No inertia improvement (-inf)");
  // cluster_count from the LAST run's final assignment (KM/algorithm.rs:304,342,355-357)
  std::vector<double> cnt(k);
  if (hamerly) cnt = best_cnt;
  else {
    CK(cudaMemcpyAsync(cnt.data(), ctx->counts_out.p, k * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
  }
  for (uint64_t c = 0; c < k; ++c) counts_out[c] = (F)cnt[c];
  std::memcpy(centroids_out, best.data(), kd * sizeof(F));
  *inertia_out = min_inertia / (F)n_glob;
  return LKM_OK;
}

template <typename F>
static lkm_status lloyd_iters_impl(lkm_ctx* ctx, F* centroids_inout, uint64_t k, uint64_t max_iters, double tolerance,
                                   F* counts_out, F* inertia_total_out, lkm_stats* stats, bool hamerly = false) {
  if (!centroids_inout) return fail(ctx, LKM_ERR_INVALID_ARG, "null argument");
  if (k == 0) return fail(ctx, LKM_ERR_N_CLUSTERS, "n_clusters cannot be 0");
  if (tolerance < 0.0) return fail(ctx, LKM_ERR_TOLERANCE, "tolerance must not be negative");
  if (max_iters == 0) return fail(ctx, LKM_ERR_MAX_ITERATIONS, "max_n_iterations cannot be 0");
  if (ctx->elem != (int)sizeof(F) || !ctx->X) return fail(ctx, LKM_ERR_NO_DATA, "no resident data of this type");
  CK(cudaSetDevice(ctx->device));
  const uint64_t kd = k * ctx->d;
  CK(ctx->cbuf.ensure(2 * kd * sizeof(F)));
  F* dC2 = ctx->cbuf.as<F>();
  ctx->launches = 0;
  CK(cudaMemcpyAsync(dC2, centroids_inout, kd * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  LloydOut lo;
  if (hamerly) CKS(run_hamerly<F>(ctx, dC2, k, max_iters, (double)(F)tolerance, lo));
  else CKS(run_lloyd<F>(ctx, dC2, k, max_iters, (double)(F)tolerance, lo));
  CK(cudaMemcpyAsync(centroids_inout, dC2 + (size_t)lo.final_buf * kd, kd * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<double> cnt(k);
  CK(cudaMemcpyAsync(cnt.data(), ctx->counts_out.p, k * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (counts_out) for (uint64_t c = 0; c < k; ++c) counts_out[c] = (F)cnt[c];
  if (inertia_total_out) *inertia_total_out = (F)lo.inertia;
  if (stats) {
    std::memset(stats, 0, sizeof(*stats));
    stats->n_iter = lo.n_iter; stats->last_shift = lo.shift; stats->inertia_total = lo.inertia;
    stats->device_ms = lo.ms; stats->kernel_launches = ctx->launches; stats->assign_path = (uint32_t)ctx->last_path; stats->update_path = (uint32_t)ctx->last_update_path;
    stats->main_kernel_ms = lo.main_ms; stats->fallback_rows = lo.fallback_rows;
    if (hamerly) { stats->rows_scanned = ctx->ham_rows_scanned; stats->rows_tightened = ctx->ham_rows_tightened; }
  }
  return LKM_OK;
}

// stage a host matrix (row stride in elements) as a contiguous device matrix
template <typename F>
static lkm_status upload(lkm_ctx* ctx, DevBuf& buf, const F* x, uint64_t n, uint64_t d, int64_t row_stride) {
  CK(buf.ensure(std::max<uint64_t>(n * d, 1) * sizeof(F)));
  if (n == 0 || d == 0) return LKM_OK;
  if (row_stride == (int64_t)d) {
    CK(cudaMemcpyAsync(buf.p, x, n * d * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  } else {
    CK(cudaMemcpy2DAsync(buf.p, d * sizeof(F), x, (size_t)row_stride * sizeof(F), d * sizeof(F), n, cudaMemcpyHostToDevice,
                         ctx->stream));
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return LKM_OK;
}

template <typename F>
static lkm_status assign_impl(lkm_ctx* ctx, const F* centroids, uint64_t k, const F* x, uint64_t n, uint64_t d,
                              int64_t row_stride, uint64_t* labels_out, F* dists_out) {
  if (!centroids) return fail(ctx, LKM_ERR_INVALID_ARG, "null centroids");
  if (k == 0) return fail(ctx, LKM_ERR_N_CLUSTERS, "n_clusters cannot be 0");
  if (d == 0) return fail(ctx, LKM_ERR_SHAPE, "n_features cannot be 0");
  CK(cudaSetDevice(ctx->device));
  const F* dX;
  if (x) {
    if (row_stride < (int64_t)d) return fail(ctx, LKM_ERR_SHAPE, "row_stride < n_features");
    CKS(upload<F>(ctx, ctx->tmpx, x, n, d, row_stride));
    dX = ctx->tmpx.as<F>();
  } else {
    if (ctx->elem != (int)sizeof(F) || !ctx->X) return fail(ctx, LKM_ERR_NO_DATA, "no resident data of this type");
    if (n != ctx->n || d != ctx->d) return fail(ctx, LKM_ERR_SHAPE, "shape differs from the resident matrix");
    dX = reinterpret_cast<const F*>(ctx->X);
  }
  if (n == 0) return LKM_OK;
  CK(ctx->tmpc.ensure(k * d * sizeof(F)));
  CK(cudaMemcpyAsync(ctx->tmpc.p, centroids, k * d * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  uint32_t* dl = nullptr;
  F* dd = nullptr;
  if (labels_out) { CK(ctx->labels.ensure(n * 4)); dl = ctx->labels.as<uint32_t>(); }
  if (dists_out) { CK(ctx->dists.ensure(n * sizeof(F))); dd = ctx->dists.as<F>(); }
  ctx->launches = 0;
  CKS(run_assign<F>(ctx, dX, n, d, ctx->tmpc.as<F>(), k, dl, dd, 0, nullptr, nullptr));
  if (dists_out) CK(cudaMemcpyAsync(dists_out, dd, n * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
  if (labels_out) {
    // usize labels on the host (Array1<usize>); u32 on the device
    std::vector<uint32_t> tmp(n);
    CK(cudaMemcpyAsync(tmp.data(), dl, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    for (uint64_t i = 0; i < n; ++i) labels_out[i] = tmp[i];
  }
  CK(cudaStreamSynchronize(ctx->stream));
  return LKM_OK;
}

template <typename F>
static lkm_status set_data_impl(lkm_ctx* ctx, const F* x, uint64_t n, uint64_t d, int64_t row_stride) {
  if ((!x && n) || d == 0) return fail(ctx, LKM_ERR_INVALID_ARG, "null data or zero features");
  if (row_stride < (int64_t)d) return fail(ctx, LKM_ERR_SHAPE, "row_stride < n_features");
  CK(cudaSetDevice(ctx->device));
  CKS(upload<F>(ctx, ctx->xbuf, x, n, d, row_stride));
  ctx->X = ctx->xbuf.p; ctx->own_X = true; ctx->n = n; ctx->d = d; ctx->elem = (int)sizeof(F);
  ctx->data_epoch++;
  if (ctx->world == 1) { ctx->row_offset = 0; ctx->n_global = n; }
  return LKM_OK;
}

template <typename F>
static lkm_status gen_blobs_impl(lkm_ctx* ctx, uint64_t n, uint64_t d, uint64_t n_blobs, const F* centres, F sigma,
                                 uint64_t seed, uint64_t row_offset, uint64_t n_global) {
  if (!centres || n_blobs == 0 || d == 0) return fail(ctx, LKM_ERR_INVALID_ARG, "bad blob description");
  CK(cudaSetDevice(ctx->device));
  CK(ctx->xbuf.ensure(std::max<uint64_t>(n * d, 1) * sizeof(F)));
  CK(ctx->tmpc.ensure(n_blobs * d * sizeof(F)));
  CK(cudaMemcpyAsync(ctx->tmpc.p, centres, n_blobs * d * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  if (n) {
    gen_blobs_kernel<F><<<ctx->sm_count * 8, 256, 0, ctx->stream>>>(ctx->xbuf.as<F>(), n, (int)d, n_blobs, ctx->tmpc.as<F>(),
                                                                   sigma, seed, row_offset, n_global ? n_global : n, ctx->blob_shuffled);
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(ctx->stream));
  ctx->X = ctx->xbuf.p; ctx->own_X = true; ctx->n = n; ctx->d = d; ctx->elem = (int)sizeof(F);
  ctx->data_epoch++;
  ctx->row_offset = row_offset; ctx->n_global = n_global ? n_global : n;
  return LKM_OK;
}

template <typename F> static lkm_status get_data_impl(lkm_ctx* ctx, uint64_t row0, uint64_t rows, F* out) {
  if (ctx->elem != (int)sizeof(F) || !ctx->X) return fail(ctx, LKM_ERR_NO_DATA, "no resident data of this type");
  if (row0 + rows > ctx->n) return fail(ctx, LKM_ERR_SHAPE, "row range outside the resident matrix");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(out, reinterpret_cast<const F*>(ctx->X) + row0 * ctx->d, rows * ctx->d * sizeof(F),
                     cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LKM_OK;
}

// FitWith::fit_with (KM/algorithm.rs:635-715)
template <typename F>
static lkm_status fit_with_impl(lkm_ctx* ctx, const lkm_params* p, lkm_rng rng_in, int has_model, uint64_t model_rows,
                                uint64_t model_cols, F* C_inout, F* counts_inout, F* inertia_out) {
  if (!p || !C_inout || !counts_inout || !inertia_out) return fail(ctx, LKM_ERR_INVALID_ARG, "null argument");
  // the ParamGuard blanket impl (src/param_guard.rs:75-86) checks the parameters before fit_with runs its own test
  CKS(check_params(ctx, p, true));
  if (p->algorithm == LKM_ALGO_HAMERLY)
    return fail(ctx, LKM_ERR_INCREMENTAL_HAMERLY, "only KMeansAlgorithm::Lloyd is supported by fit_with");
  if (ctx->elem != (int)sizeof(F) || !ctx->X) return fail(ctx, LKM_ERR_NO_DATA, "no resident data of this type");
  if (ctx->world > 1) return fail(ctx, LKM_ERR_UNSUPPORTED, "fit_with is single-GPU");
  CK(cudaSetDevice(ctx->device));
  // with a model the reference never looks at n_clusters: the model's own shape rules (KM/algorithm.rs:680-690)
  if (has_model && (model_rows == 0 || model_cols != ctx->d))
    return fail(ctx, LKM_ERR_SHAPE, "model centroids must be (rows > 0) x n_features of the batch");
  const uint64_t k = has_model ? model_rows : p->n_clusters, n = ctx->n, d = ctx->d, kd = k * d;
  const F* dX = reinterpret_cast<const F*>(ctx->X);
  HostRng rng{rng_in};
  CK(ctx->cbuf.ensure(2 * kd * sizeof(F)));
  F* dC = ctx->cbuf.as<F>();
  CK(ctx->dists.ensure(std::max<uint64_t>(n, 1) * sizeof(F)));
  CK(ctx->labels.ensure(std::max<uint64_t>(n, 1) * 4));
  CK(ctx->part_inertia.ensure(1024 * 8));
  std::vector<double> parts(1024);
  auto dist_sum = [&](double& s) -> lkm_status {
    sum_f64_kernel<F><<<1024, 256, 0, ctx->stream>>>(ctx->dists.as<F>(), n, ctx->part_inertia.as<double>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(parts.data(), ctx->part_inertia.p, 1024 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    s = 0.0;
    for (double v : parts) s += v;
    return LKM_OK;
  };
  std::vector<F> hC(kd);
  if (!has_model) {
    if (p->init_method == LKM_INIT_PRECOMPUTED) {
      if (!p->precomputed || p->precomputed_rows != k || p->precomputed_cols != d)
        return fail(ctx, LKM_ERR_SHAPE, "Precomputed centroids must be n_clusters x n_features");
      std::memcpy(hC.data(), p->precomputed, kd * sizeof(F));
    } else {
      bool have = false;
      F best = 0;
      std::vector<F> cand(kd);
      for (uint64_t r = 0; r < p->n_runs; ++r) {
        CKS(run_init<F>(ctx, p, rng, dC));
        CKS(run_assign<F>(ctx, dX, n, d, dC, k, nullptr, ctx->dists.as<F>(), 0, nullptr, nullptr));
        double s;
        CKS(dist_sum(s));
        const F cost = (F)s;
        if (!have || !(best < cost)) {  // Iterator::min_by with `d1 < d2 ? Less : Greater`
          best = cost; have = true;
          CK(cudaMemcpy(hC.data(), dC, kd * sizeof(F), cudaMemcpyDeviceToHost));
        }
      }
    }
    for (uint64_t c = 0; c < k; ++c) counts_inout[c] = (F)0;
  } else {
    std::memcpy(hC.data(), C_inout, kd * sizeof(F));
  }
  CK(cudaMemcpyAsync(dC, hC.data(), kd * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  CKS(run_assign<F>(ctx, dX, n, d, dC, k, ctx->labels.as<uint32_t>(), ctx->dists.as<F>(), 0, nullptr, nullptr));
  double s;
  CKS(dist_sum(s));
  // compute_centroids_incremental (KM/algorithm.rs:815-835): order-dependent running
  // mean, walked in row order by one CTA (incremental_update_kernel)
  F* dCn = dC + kd;
  CK(cudaMemcpyAsync(dCn, dC, kd * sizeof(F), cudaMemcpyDeviceToDevice, ctx->stream));
  const int threads = (int)std::min<uint64_t>(1024, ((d + 31) / 32) * 32);
  CK(ctx->tmpc.ensure((k + (size_t)(threads / 32) * k + 8) * sizeof(F)));
  F* dcounts = ctx->tmpc.as<F>();
  CK(cudaMemcpyAsync(dcounts, counts_inout, k * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
  incremental_update_kernel<F><<<1, threads, 0, ctx->stream>>>(dX, ctx->labels.as<uint32_t>(), n, (int)d, (int)k, dCn, dcounts,
                                                               dcounts + k + 4);
  CK(cudaGetLastError());
  CK(ctx->scratch.ensure(64));
  seq_l2_distance_kernel<F><<<1, 32, 0, ctx->stream>>>(dC, dCn, (int)kd, ctx->scratch.as<F>());
  CK(cudaGetLastError());
  ctx->launches += 2;
  F shift = (F)0;
  CK(cudaMemcpyAsync(&shift, ctx->scratch.p, sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(C_inout, dCn, kd * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(counts_inout, dcounts, k * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  *inertia_out = (F)s / (F)n;
  return shift < (F)p->tolerance ? LKM_OK : LKM_NOT_CONVERGED;
}

}  // namespace lkm

// ------------------------------------------------------------------------------------
// C ABI
// ------------------------------------------------------------------------------------
LKM_EXPORT const char* lkm_version(void) { return "linfa_b200 lkm 0.1 (sm_100a)"; }

LKM_EXPORT lkm_status lkm_create(lkm_ctx** out, int device) {
  if (!out) return LKM_ERR_INVALID_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return LKM_ERR_CUDA;
  lkm_ctx* ctx = new lkm_ctx();
  ctx->device = device;
  auto bail = [&](lkm_status s) { delete ctx; return s; };
  if (cudaSetDevice(device) != cudaSuccess) return bail(LKM_ERR_CUDA);
  if (cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking) != cudaSuccess) return bail(LKM_ERR_CUDA);
  cudaEventCreate(&ctx->ev0);
  cudaEventCreate(&ctx->ev1);
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return bail(LKM_ERR_CUDA);
  if (const char* v = std::getenv("LKM_TC_VARIANT")) ctx->tc_variant = std::atoi(v);
  if (const char* v = std::getenv("LKM_UPD")) ctx->upd_variant = std::atoi(v);
  if (const char* v = std::getenv("LKM_TC5")) ctx->tc5_enabled = std::atoi(v);
  if (const char* v = std::getenv("LKM_TC5_FUSE")) ctx->tc5_fuse = std::atoi(v);
  if (const char* v = std::getenv("LKM_DMMA")) ctx->dmma_enabled = std::atoi(v);
  if (const char* v = std::getenv("LKM_DEVICE_PICKS")) ctx->device_picks = std::atoi(v);
  if (const char* v = std::getenv("LKM_ROWS_ASSIGN")) ctx->rows_assign = std::atoi(v);
  if (const char* v = std::getenv("LKM_FIN_CL")) ctx->fin_cl = std::atoi(v);
  if (const char* v = std::getenv("LKM_P2P_LL")) ctx->p2p_ll = std::atoi(v);
  if (const char* v = std::getenv("LKM_SORT_UPD")) ctx->sort_upd = std::atoi(v);
  if (const char* v = std::getenv("LKM_SEED_PASS")) ctx->seed_pass = std::atoi(v);
  if (const char* v = std::getenv("LKM_SEED_PRUNE")) ctx->seed_prune = std::atoi(v);
  if (const char* v = std::getenv("LKM_ROWS_PIVOT")) ctx->rows_pivot = std::atoi(v);
  if (const char* v = std::getenv("LKM_STEP_TRACE")) ctx->step_trace = std::atoi(v);
  if (const char* v = std::getenv("LKM_SPARSE_PARTIALS")) ctx->sparse_partials = std::atoi(v);
  if (const char* v = std::getenv("LKM_TC5_DEFER")) ctx->tc5_defer = std::atoi(v);
  if (const char* v = std::getenv("LKM_TC5_RESCAN_LIMIT")) ctx->tc5_rescan_limit = std::atof(v);
  if (const char* v = std::getenv("LKM_PAIR32")) ctx->pair32 = std::atoi(v);
  if (const char* v = std::getenv("LKM_PDL")) g_pdl = std::atoi(v);
  if (const char* v = std::getenv("LKM_COOP")) g_coop = std::atoi(v);
  if (const char* v = std::getenv("LKM_FLUSH_SWEEP")) g_flush_sweep = std::atoi(v);
  if (const char* v = std::getenv("LKM_F64S")) ctx->f64s_enabled = std::atoi(v);
  if (const char* v = std::getenv("LKM_PARA_TC")) ctx->para_tc = std::atoi(v);
  if (const char* v = std::getenv("LKM_PARA_ROWS")) ctx->para_rows = std::atoi(v);
  if (const char* v = std::getenv("LKM_FUSED_FINALIZE")) ctx->fused_finalize = std::atoi(v);
  if (const char* v = std::getenv("LKM_F64S_PACKED")) ctx->f64s_packed = std::atoi(v);
  if (const char* v = std::getenv("LKM_F64S_RPT")) ctx->f64s_rpt = std::atoi(v) == 1 ? 1 : 2;
  if (const char* v = std::getenv("LKM_TC5_PASSES")) ctx->tc5_passes_forced = std::atoi(v);
  if (const char* v = std::getenv("LKM_P2P")) ctx->p2p_enabled = std::atoi(v);
  if (const char* v = std::getenv("LKM_TC4_PASSES")) ctx->tc4_passes_forced = std::atoi(v);
  ctx->sm_count = prop.multiProcessorCount;
  ctx->smem_optin = prop.sharedMemPerBlockOptin;
  if (cudaMallocHost(&ctx->h_state, sizeof(LloydState)) != cudaSuccess) return bail(LKM_ERR_CUDA);
  if (cudaMallocHost(&ctx->h_pin, 4096 * sizeof(double)) != cudaSuccess) return bail(LKM_ERR_CUDA);
  *out = ctx;
  return LKM_OK;
}

LKM_EXPORT lkm_status lkm_destroy(lkm_ctx* ctx) {
  if (!ctx) return LKM_OK;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  for (int r = 0; r < (int)ctx->p2p_peer.size(); ++r)
    if (ctx->group == nullptr && r != ctx->rank && ctx->p2p_peer[r]) cudaIpcCloseMemHandle(ctx->p2p_peer[r]);
  if (ctx->p2p_local) cudaFree(ctx->p2p_local);
  if (ctx->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
  DevBuf* bufs[] = {&ctx->xbuf, &ctx->cbuf, &ctx->part_sums, &ctx->part_counts, &ctx->part_inertia, &ctx->shift_part,
                    &ctx->counts_out, &ctx->state, &ctx->gather, &ctx->gather2, &ctx->labels, &ctx->dists, &ctx->wdists, &ctx->cum,
                    &ctx->scratch, &ctx->tmpx, &ctx->tmpc, &ctx->idxbuf, &ctx->bsum, &ctx->weights, &ctx->flushbuf, &ctx->fin_sync, &ctx->para_c2, &ctx->tca_bsplit, &ctx->tca_cn, &ctx->tca_cmax,
                    &ctx->xmax2, &ctx->t4prof, &ctx->t5merge, &ctx->t5list, &ctx->ham_ub, &ctx->ham_lb, &ctx->ham_list, &ctx->ham_fx,
                    &ctx->ham_moved, &ctx->ham_s, &ctx->ham_state, &ctx->ham_red, &ctx->absmax, &ctx->part_touched};
  for (DevBuf* b : bufs) b->release();
  for (cudaEvent_t e : ctx->iter_events) cudaEventDestroy(e);
  if (ctx->h_state) cudaFreeHost(ctx->h_state);
  if (ctx->h_pin) cudaFreeHost(ctx->h_pin);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
  return LKM_OK;
}

LKM_EXPORT const char* lkm_last_error(const lkm_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }
LKM_EXPORT void* lkm_stream(lkm_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }
LKM_EXPORT lkm_status lkm_synchronize(lkm_ctx* ctx) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  CK(cudaStreamSynchronize(ctx->stream));
  return LKM_OK;
}
LKM_EXPORT lkm_status lkm_set_assign_path(lkm_ctx* ctx, int path) {
  if (!ctx || path < -1 || path > 4) return LKM_ERR_INVALID_ARG;
  ctx->force_path = path;
  return LKM_OK;
}

LKM_EXPORT lkm_status lkm_set_scoring_passes(lkm_ctx* ctx, int passes) {
  if (!ctx || passes < 0 || passes > 3) return LKM_ERR_INVALID_ARG;
  ctx->tc5_passes_forced = passes == 1 ? 1 : (passes >= 2 ? 3 : 0);
  ctx->tc4_passes_forced = passes;
  return LKM_OK;
}

LKM_EXPORT lkm_status lkm_set_keep_labels(lkm_ctx* ctx, int on) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  ctx->keep_labels = on ? 1 : 0;
  return LKM_OK;
}

LKM_EXPORT lkm_status lkm_get_labels(lkm_ctx* ctx, uint32_t* labels_out, uint64_t n) {
  if (!ctx || !labels_out) return LKM_ERR_INVALID_ARG;
  if (n == 0) return LKM_OK;
  if (n > ctx->kept_labels_n || !ctx->labels.p) return fail(ctx, LKM_ERR_NO_DATA, "no labels kept: call lkm_set_keep_labels before the fit");
  CK(cudaSetDevice(ctx->device));
  CK(cudaMemcpyAsync(labels_out, ctx->labels.p, n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LKM_OK;
}

LKM_EXPORT lkm_status lkm_set_blob_layout(lkm_ctx* ctx, int layout) {
  if (!ctx || layout < 0 || layout > 1) return LKM_ERR_INVALID_ARG;
  ctx->blob_shuffled = layout;
  return LKM_OK;
}

LKM_EXPORT lkm_status lkm_set_l2_flush(lkm_ctx* ctx, uint64_t bytes) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  ctx->flush_bytes = bytes;
  return LKM_OK;
}

LKM_EXPORT lkm_status lkm_set_kernel_timing(lkm_ctx* ctx, int on) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  ctx->kernel_timing = on ? 1 : 0;
  return LKM_OK;
}

LKM_EXPORT lkm_status lkm_comm_unique_id(void* id128) {
  std::string err;
  if (!id128 || !g_nccl.load(err)) return LKM_ERR_NCCL;
  return g_nccl.GetUniqueId(reinterpret_cast<NcclUid*>(id128)) == 0 ? LKM_OK : LKM_ERR_NCCL;
}
LKM_EXPORT lkm_status lkm_comm_init(lkm_ctx* ctx, const void* id128, int rank, int world_size) {
  if (!ctx || !id128 || world_size < 1 || rank < 0 || rank >= world_size) return LKM_ERR_INVALID_ARG;
  if (world_size == 1) { ctx->rank = 0; ctx->world = 1; return LKM_OK; }
  if (!g_nccl.load(ctx->err)) return LKM_ERR_NCCL;
  CK(cudaSetDevice(ctx->device));
  NcclUid uid;
  std::memcpy(&uid, id128, sizeof(uid));
  CKN(g_nccl.CommInitRank(&ctx->comm, world_size, uid, rank));
  ctx->rank = rank;
  ctx->world = world_size;
  return LKM_OK;
}
LKM_EXPORT lkm_status lkm_set_shard(lkm_ctx* ctx, uint64_t row_offset, uint64_t n_global) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  ctx->row_offset = row_offset;
  ctx->n_global = n_global;
  return LKM_OK;
}

// ---- several devices from ONE process (lkm_create(device_ids, n_devices)) ----------------------------
// `fit(&dataset)` of the Rust binding is one call on one thread; to use the whole box it hands the matrix to a device
// group: rows are split into contiguous shards, one context + one host thread per device, the per-iteration partials
// travel over NVLink exactly as in the one-process-per-GPU mode (finalize_x_kernel), the host-side steps of the seeding
// rendezvous through LocalGroup.  Every rank draws the SAME random numbers: the caller's rng is read once per draw and
// replayed to the other ranks.
struct lkm_multi {
  std::vector<lkm_ctx*> ctx;
  LocalGroup group;
  std::string err;
};

namespace lkm {
struct SharedRng {
  std::mutex mu;
  lkm_rng base;
  std::vector<std::pair<int, uint64_t>> log;   // (32 / 64, value) in draw order
};
struct RankRng {
  SharedRng* shared;
  size_t cursor;
};
static uint64_t shared_draw(RankRng* r, int kind) {
  SharedRng* S = r->shared;
  std::lock_guard<std::mutex> lk(S->mu);
  if (r->cursor == S->log.size())
    S->log.emplace_back(kind, kind == 32 ? (uint64_t)S->base.next_u32(S->base.user) : S->base.next_u64(S->base.user));
  // every rank runs the same host code, so the kinds line up; a mismatch would mean diverged control flow
  const uint64_t v = S->log[r->cursor].second;
  ++r->cursor;
  return v;
}
static uint32_t rank_next_u32(void* u) { return (uint32_t)shared_draw(static_cast<RankRng*>(u), 32); }
static uint64_t rank_next_u64(void* u) { return shared_draw(static_cast<RankRng*>(u), 64); }

template <typename Fn>
static lkm_status multi_run(lkm_multi* m, Fn fn) {
  const int w = (int)m->ctx.size();
  std::vector<lkm_status> st(w, LKM_OK);
  std::vector<std::thread> th;
  {
    std::lock_guard<std::mutex> lk(m->group.mu);
    m->group.aborted = false;
    m->group.arrived = 0;
  }
  for (int r = 0; r < w; ++r)
    th.emplace_back([&, r] {
      st[r] = fn(r, m->ctx[r]);
      if (st[r] != LKM_OK && st[r] != LKM_NOT_CONVERGED) m->group.abort();   // release the ranks that wait for this one
    });
  for (auto& t : th) t.join();
  for (int r = 0; r < w; ++r)
    if (st[r] != LKM_OK) { m->err = "rank " + std::to_string(r) + ": " + m->ctx[r]->err; return st[r]; }
  return LKM_OK;
}

template <typename F>
static lkm_status multi_set_data(lkm_multi* m, const F* x, uint64_t n, uint64_t d, int64_t row_stride) {
  if (!x || d == 0 || row_stride < (int64_t)d) { m->err = "bad matrix"; return LKM_ERR_INVALID_ARG; }
  const uint64_t w = m->ctx.size();
  return multi_run(m, [&](int r, lkm_ctx* c) -> lkm_status {
    const uint64_t lo = n * (uint64_t)r / w, hi = n * (uint64_t)(r + 1) / w;
    lkm_status st = set_data_impl<F>(c, x + lo * (uint64_t)row_stride, hi - lo, d, row_stride);
    if (st != LKM_OK) return st;
    c->row_offset = lo;
    c->n_global = n;
    return LKM_OK;
  });
}

template <typename F>
static lkm_status multi_fit(lkm_multi* m, const lkm_params* p, lkm_rng rng, F* centroids_out, F* counts_out, F* inertia_out,
                            lkm_stats* stats_out) {
  if (!p || !centroids_out || !counts_out || !inertia_out) { m->err = "null argument"; return LKM_ERR_INVALID_ARG; }
  const uint64_t k = p->n_clusters, d = m->ctx[0]->d;
  SharedRng shared;
  shared.base = rng;
  const int w = (int)m->ctx.size();
  std::vector<RankRng> rr(w);
  std::vector<std::vector<F>> cen(w, std::vector<F>(std::max<uint64_t>(k * d, 1))), cnt(w, std::vector<F>(std::max<uint64_t>(k, 1)));
  std::vector<F> inertia(w, (F)0);
  std::vector<lkm_stats> stv(w);
  const lkm_status st = multi_run(m, [&](int r, lkm_ctx* c) -> lkm_status {
    rr[r] = RankRng{&shared, 0};
    lkm_rng mine{&rr[r], rank_next_u32, rank_next_u64};
    return fit_impl<F>(c, p, mine, cen[r].data(), cnt[r].data(), &inertia[r], &stv[r]);
  });
  if (st != LKM_OK) return st;
  // identical on every rank (same partials combined in the same order): rank 0's copy is the result
  std::memcpy(centroids_out, cen[0].data(), k * d * sizeof(F));
  std::memcpy(counts_out, cnt[0].data(), k * sizeof(F));
  *inertia_out = inertia[0];
  if (stats_out) *stats_out = stv[0];
  return LKM_OK;
}

template <typename F>
static lkm_status multi_assign(lkm_multi* m, const F* centroids, uint64_t k, const F* x, uint64_t n, uint64_t d, int64_t row_stride,
                               uint64_t* labels_out, F* dist_out) {
  if (!x || row_stride < (int64_t)d) { m->err = "bad matrix"; return LKM_ERR_INVALID_ARG; }
  const uint64_t w = m->ctx.size();
  // rows are independent: every device takes a contiguous block, no exchange (KM/algorithm.rs:838-881)
  return multi_run(m, [&](int r, lkm_ctx* c) -> lkm_status {
    const uint64_t lo = n * (uint64_t)r / w, hi = n * (uint64_t)(r + 1) / w;
    if (hi == lo) return LKM_OK;
    return assign_impl<F>(c, centroids, k, x + lo * (uint64_t)row_stride, hi - lo, d, row_stride, labels_out ? labels_out + lo : nullptr,
                          dist_out ? dist_out + lo : nullptr);
  });
}
}  // namespace lkm

LKM_EXPORT lkm_status lkm_create_multi(lkm_multi** out, const int* device_ids, int n_devices) {
  if (!out || !device_ids || n_devices < 1 || n_devices > 8) return LKM_ERR_INVALID_ARG;
  lkm_multi* m = new (std::nothrow) lkm_multi();
  if (!m) return LKM_ERR_CUDA;
  m->group.world = n_devices;
  m->group.slot.resize(n_devices);
  for (int r = 0; r < n_devices; ++r) {
    lkm_ctx* c = nullptr;
    const lkm_status st = lkm_create(&c, device_ids[r]);
    if (st != LKM_OK) {
      for (lkm_ctx* q : m->ctx) lkm_destroy(q);
      delete m;
      return st;
    }
    c->rank = r;
    c->world = n_devices;
    if (n_devices > 1) c->group = &m->group;
    m->ctx.push_back(c);
  }
  *out = m;
  return LKM_OK;
}
LKM_EXPORT lkm_status lkm_destroy_multi(lkm_multi* m) {
  if (!m) return LKM_OK;
  for (lkm_ctx* c : m->ctx) lkm_destroy(c);
  delete m;
  return LKM_OK;
}
LKM_EXPORT const char* lkm_multi_last_error(const lkm_multi* m) { return m ? m->err.c_str() : "null group"; }
LKM_EXPORT int lkm_multi_size(const lkm_multi* m) { return m ? (int)m->ctx.size() : 0; }
LKM_EXPORT lkm_ctx* lkm_multi_ctx(lkm_multi* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }

#define LKM_MULTI_TYPED(F, SFX)                                                                                        \
  LKM_EXPORT lkm_status lkm_multi_set_data_##SFX(lkm_multi* m, const F* x, uint64_t n, uint64_t d, int64_t row_stride) {\
    if (!m) return LKM_ERR_INVALID_ARG;                                                                                \
    return multi_set_data<F>(m, x, n, d, row_stride);                                                                  \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_multi_fit_##SFX(lkm_multi* m, const lkm_params* p, lkm_rng rng, F* centroids_out,          \
                                            F* cluster_count_out, F* inertia_out, lkm_stats* stats_out) {              \
    if (!m) return LKM_ERR_INVALID_ARG;                                                                                \
    return multi_fit<F>(m, p, rng, centroids_out, cluster_count_out, inertia_out, stats_out);                          \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_multi_assign_##SFX(lkm_multi* m, const F* centroids, uint64_t k, const F* x, uint64_t n,   \
                                               uint64_t d, int64_t row_stride, uint64_t* labels_out,                   \
                                               F* min_sq_dist_out) {                                                   \
    if (!m) return LKM_ERR_INVALID_ARG;                                                                                \
    return multi_assign<F>(m, centroids, k, x, n, d, row_stride, labels_out, min_sq_dist_out);                         \
  }

#define LKM_TYPED(F, SFX)                                                                                              \
  LKM_EXPORT lkm_status lkm_set_data_##SFX(lkm_ctx* ctx, const F* x, uint64_t n, uint64_t d, int64_t row_stride) {     \
    if (!ctx) return LKM_ERR_INVALID_ARG;                                                                              \
    return set_data_impl<F>(ctx, x, n, d, row_stride);                                                                 \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_set_data_device_##SFX(lkm_ctx* ctx, const void* x_device, uint64_t n, uint64_t d) {        \
    if (!ctx || (!x_device && n) || d == 0) return LKM_ERR_INVALID_ARG;                                                \
    ctx->X = const_cast<void*>(x_device); ctx->own_X = false; ctx->n = n; ctx->d = d; ctx->elem = (int)sizeof(F);      \
    ctx->data_epoch++;                                                                                                 \
    if (ctx->world == 1) { ctx->row_offset = 0; ctx->n_global = n; }                                                   \
    return LKM_OK;                                                                                                     \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_gen_blobs_##SFX(lkm_ctx* ctx, uint64_t n, uint64_t d, uint64_t n_blobs, const F* centres,  \
                                            F sigma, uint64_t seed, uint64_t row_offset, uint64_t n_global) {          \
    if (!ctx) return LKM_ERR_INVALID_ARG;                                                                              \
    return gen_blobs_impl<F>(ctx, n, d, n_blobs, centres, sigma, seed, row_offset, n_global);                          \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_get_data_##SFX(lkm_ctx* ctx, uint64_t row0, uint64_t rows, F* out) {                       \
    if (!ctx || !out) return LKM_ERR_INVALID_ARG;                                                                      \
    return get_data_impl<F>(ctx, row0, rows, out);                                                                     \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_fit_##SFX(lkm_ctx* ctx, const lkm_params* p, lkm_rng rng, F* centroids_out,                \
                                      F* cluster_count_out, F* inertia_out, lkm_stats* stats_out) {                    \
    if (!ctx) return LKM_ERR_INVALID_ARG;                                                                              \
    return fit_impl<F>(ctx, p, rng, centroids_out, cluster_count_out, inertia_out, stats_out);                         \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_init_##SFX(lkm_ctx* ctx, const lkm_params* p, lkm_rng rng, F* centroids_out) {             \
    if (!ctx || !p || !centroids_out) return LKM_ERR_INVALID_ARG;                                                      \
    if (p->n_clusters == 0) return fail(ctx, LKM_ERR_N_CLUSTERS, "n_clusters cannot be 0");                            \
    if (ctx->elem != (int)sizeof(F) || !ctx->X) return fail(ctx, LKM_ERR_NO_DATA, "no resident data of this type");    \
    CK(cudaSetDevice(ctx->device));                                                                                    \
    const uint64_t kd = p->n_clusters * ctx->d;                                                                        \
    CK(ctx->cbuf.ensure(2 * kd * sizeof(F)));                                                                          \
    HostRng rng_{rng};                                                                                                 \
    ctx->launches = 0;                                                                                                 \
    CKS(run_init<F>(ctx, p, rng_, ctx->cbuf.as<F>()));                                                                 \
    CK(cudaMemcpyAsync(centroids_out, ctx->cbuf.p, kd * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));              \
    CK(cudaStreamSynchronize(ctx->stream));                                                                            \
    return LKM_OK;                                                                                                     \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_lloyd_iters_##SFX(lkm_ctx* ctx, F* centroids_inout, uint64_t k, uint64_t max_iters,        \
                                              double tolerance, F* cluster_count_out, F* inertia_total_out,            \
                                              lkm_stats* stats_out) {                                                  \
    if (!ctx) return LKM_ERR_INVALID_ARG;                                                                              \
    return lloyd_iters_impl<F>(ctx, centroids_inout, k, max_iters, tolerance, cluster_count_out, inertia_total_out,    \
                               stats_out);                                                                             \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_hamerly_iters_##SFX(lkm_ctx* ctx, F* centroids_inout, uint64_t k, uint64_t max_iters,      \
                                                double tolerance, F* cluster_count_out, F* inertia_total_out,          \
                                                lkm_stats* stats_out) {                                                \
    if (!ctx) return LKM_ERR_INVALID_ARG;                                                                              \
    return lloyd_iters_impl<F>(ctx, centroids_inout, k, max_iters, tolerance, cluster_count_out, inertia_total_out,    \
                               stats_out, true);                                                                       \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_predict_##SFX(lkm_ctx* ctx, const F* centroids, uint64_t k, const F* x, uint64_t n,        \
                                          uint64_t d, int64_t row_stride, uint64_t* labels_out) {                      \
    if (!ctx || !labels_out) return LKM_ERR_INVALID_ARG;                                                               \
    return assign_impl<F>(ctx, centroids, k, x, n, d, row_stride, labels_out, nullptr);                                \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_transform_##SFX(lkm_ctx* ctx, const F* centroids, uint64_t k, const F* x, uint64_t n,      \
                                            uint64_t d, int64_t row_stride, F* min_sq_dist_out) {                      \
    if (!ctx || !min_sq_dist_out) return LKM_ERR_INVALID_ARG;                                                          \
    return assign_impl<F>(ctx, centroids, k, x, n, d, row_stride, nullptr, min_sq_dist_out);                           \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_assign_##SFX(lkm_ctx* ctx, const F* centroids, uint64_t k, const F* x, uint64_t n,         \
                                         uint64_t d, int64_t row_stride, uint64_t* labels_out, F* min_sq_dist_out) {   \
    if (!ctx) return LKM_ERR_INVALID_ARG;                                                                              \
    return assign_impl<F>(ctx, centroids, k, x, n, d, row_stride, labels_out, min_sq_dist_out);                        \
  }                                                                                                                    \
  LKM_EXPORT lkm_status lkm_fit_with_##SFX(lkm_ctx* ctx, const lkm_params* p, lkm_rng rng, int has_model,              \
                                           uint64_t model_rows, uint64_t model_cols, F* centroids_inout,               \
                                           F* cluster_count_inout, F* inertia_out) {                                   \
    if (!ctx) return LKM_ERR_INVALID_ARG;                                                                              \
    return fit_with_impl<F>(ctx, p, rng, has_model, model_rows, model_cols, centroids_inout, cluster_count_inout,     \
                            inertia_out);                                                                              \
  }

LKM_TYPED(float, f32)
LKM_TYPED(double, f64)
LKM_MULTI_TYPED(float, f32)
LKM_MULTI_TYPED(double, f64)

#ifdef LKM_WITH_PROBES
// ---- development probes (not part of include/lkm.h) ---------------------------------------
// S[128 x n_c] = X[128 x 32] * C[n_c x 32]^T through the tcgen05 3xTF32 building blocks.
LKM_EXPORT int lkm_debug_tc_scores(lkm_ctx* ctx, const float* x, const float* c, int n_c, float* s_out) {
  if (!ctx || !x || !c || !s_out) return LKM_ERR_INVALID_ARG;
  if (n_c != 64 && n_c != 16 && n_c != 128 && n_c != 256) return LKM_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(ctx->tmpx.ensure(128 * 32 * 4));
  CK(ctx->tmpc.ensure((size_t)n_c * 32 * 4));
  CK(ctx->dists.ensure((size_t)128 * n_c * 4));
  CK(ctx->scratch.ensure(64));
  CK(cudaMemcpyAsync(ctx->tmpx.p, x, 128 * 32 * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tmpc.p, c, (size_t)n_c * 32 * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->scratch.p, 0xff, 4, ctx->stream));
  const size_t smem = 32768 + (size_t)n_c * 256 + 1024;
  int* st = ctx->scratch.as<int>();
#define LKM_DBG_LAUNCH(NN)                                                                                          \
  {                                                                                                                 \
    CK(cudaFuncSetAttribute(tc_scores_debug_kernel<NN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));   \
    tc_scores_debug_kernel<NN><<<1, 128, smem, ctx->stream>>>(ctx->tmpx.as<float>(), ctx->tmpc.as<float>(),        \
                                                               ctx->dists.as<float>(), st);                         \
  }
  if (n_c == 16) LKM_DBG_LAUNCH(16) else if (n_c == 64) LKM_DBG_LAUNCH(64) else if (n_c == 128) LKM_DBG_LAUNCH(128) else LKM_DBG_LAUNCH(256)
#undef LKM_DBG_LAUNCH
  CK(cudaGetLastError());
  int hst = -1;
  CK(cudaMemcpyAsync(&hst, st, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(s_out, ctx->dists.p, (size_t)128 * n_c * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (hst != 0) return fail(ctx, LKM_ERR_CUDA, hst == 1 ? "tcgen05 probe: mbarrier wait timed out" : "tcgen05 probe: kernel did not report");
  return LKM_OK;
}

// D2 = onehot(labels)^T * X through the MN-major tcgen05 accumulation probe; dumps 128 lanes x 32 columns
LKM_EXPORT int lkm_debug_tc_accum(lkm_ctx* ctx, const float* x, const int* labels, int m, float* dump_out) {
  if (!ctx || !x || !labels || !dump_out || (m != 64 && m != 128)) return LKM_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(ctx->tmpx.ensure(128 * 32 * 4));
  CK(ctx->tmpc.ensure(128 * 4));
  CK(ctx->dists.ensure(128 * 32 * 4));
  CK(ctx->scratch.ensure(64));
  CK(cudaMemcpyAsync(ctx->tmpx.p, x, 128 * 32 * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tmpc.p, labels, 128 * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->scratch.p, 0xff, 4, ctx->stream));
  CK(cudaMemsetAsync(ctx->dists.p, 0, 128 * 32 * 4, ctx->stream));
  const size_t smem = 16384 + (size_t)(m / 32) * 16384 + 1024;
  if (m == 64) {
    CK(cudaFuncSetAttribute(tc_accum_debug_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_accum_debug_kernel<64><<<1, 128, smem, ctx->stream>>>(ctx->tmpx.as<float>(), ctx->tmpc.as<int>(), ctx->dists.as<float>(), ctx->scratch.as<int>());
  } else {
    CK(cudaFuncSetAttribute(tc_accum_debug_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    tc_accum_debug_kernel<128><<<1, 128, smem, ctx->stream>>>(ctx->tmpx.as<float>(), ctx->tmpc.as<int>(), ctx->dists.as<float>(), ctx->scratch.as<int>());
  }
  CK(cudaGetLastError());
  int hst = -1;
  CK(cudaMemcpyAsync(&hst, ctx->scratch.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(dump_out, ctx->dists.p, 128 * 32 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (hst != 0) return fail(ctx, LKM_ERR_CUDA, hst == 1 ? "tcgen05 accumulation probe: timeout" : "tcgen05 accumulation probe: no report");
  return LKM_OK;
}

LKM_EXPORT int lkm_debug_umma_raw(lkm_ctx* ctx, const void* a_img, int a_bytes, const void* b_img, int b_bytes,
                                  unsigned long long a_desc_hi, unsigned long long b_desc_hi, int a_step, int b_step,
                                  unsigned int idesc, int ksteps, int ncols, float* dump_out) {
  if (!ctx || !a_img || !b_img || !dump_out || a_bytes % 16 || b_bytes % 16 || ncols % 32 || ncols > 256) return LKM_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(ctx->tmpx.ensure(a_bytes));
  CK(ctx->tmpc.ensure(b_bytes));
  CK(ctx->dists.ensure((size_t)128 * ncols * 4));
  CK(ctx->scratch.ensure(64));
  CK(cudaMemcpyAsync(ctx->tmpx.p, a_img, a_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tmpc.p, b_img, b_bytes, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->scratch.p, 0xff, 4, ctx->stream));
  CK(cudaMemsetAsync(ctx->dists.p, 0, (size_t)128 * ncols * 4, ctx->stream));
  const size_t smem = (size_t)((a_bytes + 1023) & ~1023) + b_bytes + 2048;
  CK(cudaFuncSetAttribute(umma_raw_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  umma_raw_probe_kernel<<<1, 128, smem, ctx->stream>>>(ctx->tmpx.as<unsigned char>(), a_bytes, ctx->tmpc.as<unsigned char>(), b_bytes,
                                                       a_desc_hi, b_desc_hi, a_step, b_step, idesc, ksteps, ncols,
                                                       ctx->dists.as<float>(), ctx->scratch.as<int>());
  CK(cudaGetLastError());
  int hst = -1;
  CK(cudaMemcpyAsync(&hst, ctx->scratch.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(dump_out, ctx->dists.p, (size_t)128 * ncols * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (hst != 0) return fail(ctx, LKM_ERR_CUDA, "umma raw probe: timeout or no report");
  return LKM_OK;
}

LKM_EXPORT int lkm_debug_mma_bench(lkm_ctx* ctx, int n_cols, int reps, int per_batch, int mode, int grid, long long* cycles_out) {
  if (!ctx || !cycles_out || n_cols % 16 || n_cols < 16 || n_cols > 256 || grid < 1 || grid > 1024) return LKM_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(ctx->scratch.ensure((size_t)grid * 8));
  const size_t smem = 32768 + (size_t)n_cols * 128 + 2048;
  CK(cudaFuncSetAttribute(mma_bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (mode == 33 || mode == 34) {   // TMEM store throughput / stores and loads at the same time: n_cols / 16 = warps
    CK(ctx->dists.ensure((size_t)grid * 64));
    CK(cudaMemsetAsync(ctx->dists.p, 0, (size_t)grid * 64, ctx->stream));
    sttm_bench_kernel<<<grid, 256, 0, ctx->stream>>>(n_cols / 16, reps, mode - 32, ctx->dists.as<long long>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cycles_out, ctx->dists.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LKM_OK;
  }
  if (mode == 32) {   // TMEM load throughput: n_cols = warps, per_batch unused
    CK(cudaMemsetAsync(ctx->scratch.p, 0, (size_t)grid * 8, ctx->stream));
    CK(ctx->dists.ensure((size_t)grid * 64));
    ldtm_bench_kernel<<<grid, 256, 0, ctx->stream>>>(n_cols / 16, reps, ctx->dists.as<long long>());
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(cycles_out, ctx->dists.p, (size_t)std::min(grid, 1) * 64, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return LKM_OK;
  }
  if (mode >= 16 && mode < 20) {
    auto kern = mode == 16 ? mma_bench2_kernel<0> : mode == 17 ? mma_bench2_kernel<1> : mode == 18 ? mma_bench2_kernel<2> : mma_bench2_kernel<3>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kern<<<grid, 128, smem, ctx->stream>>>(n_cols, reps, ctx->scratch.as<long long>());
  } else
  mma_bench_kernel<<<grid, 128, smem, ctx->stream>>>(n_cols, reps, per_batch, mode, ctx->scratch.as<long long>());
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(cycles_out, ctx->scratch.p, (size_t)grid * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  return LKM_OK;
}

LKM_EXPORT int lkm_debug_tc_ts(lkm_ctx* ctx, const float* x, const float* c, float* s_out) {
  if (!ctx || !x || !c || !s_out) return LKM_ERR_INVALID_ARG;
  CK(cudaSetDevice(ctx->device));
  CK(ctx->tmpx.ensure(128 * 32 * 4));
  CK(ctx->tmpc.ensure(64 * 32 * 4));
  CK(ctx->dists.ensure(128 * 64 * 4));
  CK(ctx->scratch.ensure(64));
  CK(cudaMemcpyAsync(ctx->tmpx.p, x, 128 * 32 * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemcpyAsync(ctx->tmpc.p, c, 64 * 32 * 4, cudaMemcpyHostToDevice, ctx->stream));
  CK(cudaMemsetAsync(ctx->scratch.p, 0xff, 4, ctx->stream));
  const size_t smem = 64 * 128 + 2048;
  tc_ts_probe_kernel<64><<<1, 128, smem, ctx->stream>>>(ctx->tmpx.as<float>(), ctx->tmpc.as<float>(), ctx->dists.as<float>(),
                                                       ctx->scratch.as<int>());
  CK(cudaGetLastError());
  int hst = -1;
  CK(cudaMemcpyAsync(&hst, ctx->scratch.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaMemcpyAsync(s_out, ctx->dists.p, 128 * 64 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CK(cudaStreamSynchronize(ctx->stream));
  if (hst != 0) return fail(ctx, LKM_ERR_CUDA, "TS probe: timeout");
  return LKM_OK;
}
#endif  // LKM_WITH_PROBES

// Context of the C ABI (include/lkm.h): device buffers, knobs and workspaces shared by the translation units of liblkm.so.
#pragma once
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <condition_variable>
#include <cstdint>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/lkm.h"
#include "lkm_kernels.cuh"

namespace lkm {

// ---- NCCL through dlopen (only needed when world_size > 1) -----------------
struct NcclUid { char internal[128]; };
struct NcclApi {
  void* handle = nullptr;
  int (*GetUniqueId)(NcclUid*) = nullptr;
  int (*CommInitRank)(void**, int, NcclUid, int) = nullptr;
  int (*CommDestroy)(void*) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err) {
    if (handle) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
      handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
      if (handle) break;
    }
    if (!handle) { err = std::string("dlopen libnccl failed: ") + dlerror(); return false; }
    GetUniqueId = (decltype(GetUniqueId))dlsym(handle, "ncclGetUniqueId");
    CommInitRank = (decltype(CommInitRank))dlsym(handle, "ncclCommInitRank");
    CommDestroy = (decltype(CommDestroy))dlsym(handle, "ncclCommDestroy");
    AllGather = (decltype(AllGather))dlsym(handle, "ncclAllGather");
    GetErrorString = (decltype(GetErrorString))dlsym(handle, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !CommDestroy || !AllGather) { err = "libnccl lacks required symbols"; return false; }
    return true;
  }
};
constexpr int kNcclFloat64 = 8, kNcclInt8 = 0;

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr; cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e == cudaSuccess) cap = bytes;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
  template <class T> T* as() { return reinterpret_cast<T*>(p); }
};

// Ranks that live in ONE process (lkm_create_multi: one context per device, one host thread per rank): the host-side
// all-gather of the seeding / set-up steps goes through this rendezvous instead of NCCL, and the NVLink exchange buffers
// are mapped by plain peer access instead of CUDA IPC.
struct LocalGroup {
  int world = 1;
  std::mutex mu;
  std::condition_variable cv;
  int arrived = 0;
  unsigned long long gen = 0;
  bool aborted = false;
  std::vector<std::vector<unsigned char>> slot;
  // false: another rank failed and left
  bool barrier() {
    std::unique_lock<std::mutex> lk(mu);
    if (aborted) return false;
    const unsigned long long g = gen;
    if (++arrived == world) { arrived = 0; ++gen; cv.notify_all(); return true; }
    cv.wait(lk, [&] { return gen != g || aborted; });
    return !aborted;
  }
  void abort() {
    std::lock_guard<std::mutex> lk(mu);
    aborted = true;
    cv.notify_all();
  }
};

}  // namespace lkm

using namespace lkm;

struct lkm_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  std::string err;
  int sm_count = 148;
  size_t smem_optin = 227 * 1024;
  // resident observations
  int elem = 0;  // 0 none, 4 f32, 8 f64
  void* X = nullptr;
  bool own_X = false;
  DevBuf xbuf;
  uint64_t n = 0, d = 0;
  uint64_t row_offset = 0, n_global = 0;
  int force_path = -1;
  int last_path = 0;
  // fused tcgen05 kernel for d = 32: 4 = TMA + warp-specialised pipeline (k <= 64), 128 = the shared-memory-operand
  // kernel that also serves 64 < k <= 256 (LKM_TC_VARIANT=128 forces it for every k, an A/B knob)
  int tc_variant = 4;
  int tc5_passes_forced = 0;            // LKM_TC5_PASSES = 1 / 3 pins the screening / 3xTF32 variant of the pair kernel
  int fused_finalize = 0;   // LKM_FUSED_FINALIZE=1: finalize folded into the pipeline kernel's tail (measured: no gain, 47.1 vs 46.3 us per cfg2 step)
  int f64s_enabled = 1;  // LKM_F64S=0: lloyd_kernel<double> instead of the f32-screened small-d f64 kernel (lkm_f64s.cuh)
  int f64s_packed = 1;   // LKM_F64S_PACKED=0: scalar FFMA instead of FFMA2 in that kernel
  int f64s_rpt = 2;      // LKM_F64S_RPT=1: one row per lane (32-row tiles, 3 CTAs / SM)
  int pair32 = 0;        // LKM_PAIR32=1: d = 32, k <= 64 through the CTA-pair screening kernel instead of the pipeline kernel
  int tc5_fuse = 1;      // LKM_TC5_FUSE=0: screening kernel without the fused update of uniform tiles
  // LKM_STEP_TRACE=1: an event after every stage of a Lloyd iteration (tc5 / update path); run_lloyd prints the mean
  // time per stage to stderr.  The events serialise the stages (no programmatic dependent launch overlap): a tuning aid.
  int step_trace = 0;
  std::vector<cudaEvent_t> trace_ev;
  std::vector<const char*> trace_name;
  int device_picks = 1;      // LKM_DEVICE_PICKS=0: FAST k-means++ picks with a host round trip each (the first always has one)
  int rows_assign = 1;       // LKM_ROWS_ASSIGN=0: predict / transform of wide rows through lloyd_kernel<MODE_ASSIGN> (exact scan)
  int fin_cl = 1;            // LKM_FIN_CL=0: slice-wise finalize kernels even when the partials are sparse
  int sparse_partials = 1;   // LKM_SPARSE_PARTIALS=0: the pair / update kernels zero-fill and write back whole k x d partials
  DevBuf part_touched;
  int dmma_enabled = 1;  // LKM_DMMA=0: f64 shapes with a large contraction stay on the CUDA-core kernels
  int tc5_enabled = 1;   // LKM_TC5=0: tc_assign_kernel instead of the CTA-pair assignment kernel
  double tc5_rescan_limit = 0.08;   // LKM_TC5_RESCAN_LIMIT: share of listed rows above which the pair kernel moves to 3xTF32
  int tc5_defer = 1;     // LKM_TC5_DEFER=0: undecided rows are scanned inline by the pair kernel instead of by rows_top2_kernel
  int last_update_path = 0;
  int seed_prune = 1;    // LKM_SEED_PRUNE=0: k-means++ distance passes read every row (no triangle-inequality pruning)
  DevBuf seed_nearest, seed_sd;
  int seed_pass = 1;     // LKM_SEED_PASS=0: the per-pick distance pass of k-means++ through the generic assignment kernel
  int rows_pivot = -1;   // LKM_ROWS_PIVOT: -1 = pivot pruning (lkm_rows.cuh) of the listed rows once an iteration listed many, 0 never, 1 always
  DevBuf piv_cc, piv_list2;
  int sort_upd = -1;     // LKM_SORT_UPD: -1 = sorted update (lkm_sortupd.cuh) when the first iteration finds the rows unordered, 0 never, 1 always
  DevBuf srt_hist, srt_seg, srt_perm, srt_ent, srt_inert;
  int upd_variant = 1;   // LKM_UPD=0: cluster-chunked lloyd_kernel<MODE_UPDATE> instead of update_rows_kernel
  uint64_t data_epoch = 1, xmax2_epoch = 0;  // resident-data version / version the cached max row norm belongs to
  DevBuf xmax2, t4prof, t5merge, t5list;
  int tc4_passes_forced = 0;              // LKM_TC4_PASSES = 2 / 3 pins the scoring passes of the pipeline kernel
  uint64_t flush_bytes = 0;
  unsigned long long bar_epoch = 0;   // rank_barrier_kernel (flushed-L2 benchmark mode, multi-GPU)
  int kernel_timing = 0;
  int keep_labels = 0;     // lkm_set_keep_labels: the Lloyd kernels also write the per-row labels of the last iteration
  uint64_t kept_labels_n = 0;
  int blob_shuffled = 0;   // lkm_set_blob_layout
  DevBuf flushbuf, fin_sync, para_c2;
  int para_tc = 1;       // LKM_PARA_TC=0: k-means|| candidate histogram through the exact CUDA-core assignment instead of one tensor-core
                         // Lloyd iteration (same seeds; with the deferred re-scans the tensor-core route is faster: 4M x 128: 1147 -> 730 ms)
  int para_rows = 1;     // LKM_PARA_ROWS=0: k-means|| running minimum through lloyd_kernel<MODE_ASSIGN> instead of rows_top2_kernel
  std::vector<cudaEvent_t> iter_events;
  // workspaces
  DevBuf tca_bsplit, tca_cn, tca_cmax;
  DevBuf cbuf, part_sums, part_counts, part_inertia, shift_part, counts_out, state, gather, gather2, labels, dists, wdists,
      cum, scratch, tmpx, tmpc, idxbuf, bsum, weights;
  LloydState* h_state = nullptr;  // pinned
  double* h_pin = nullptr;        // pinned scratch (>= 4096 doubles)
  // comm
  void* comm = nullptr;
  int rank = 0, world = 1;
  lkm::LocalGroup* group = nullptr;   // ranks of one process (lkm_create_multi); nullptr: one process per rank (NCCL)
  // NVLink one-shot exchange: this rank's gather buffer [2 parities][world][p2p_cap doubles], [2][world][kFxMaxCtas] exchange
  // flags, [world] barrier flags and, from p2p_ll_off on, [2][world][p2p_cap] 16-byte tagged slots (p2p_setup); mapped into every
  // peer through CUDA IPC, or reached by plain peer access inside a device group
  void* p2p_local = nullptr;
  std::vector<void*> p2p_peer;
  size_t p2p_cap = 0;
  size_t p2p_ll_off = 0;   // byte offset of the 16-byte self-validating slots [2 parities][world][p2p_cap] (finalize_cl_kernel)
  int p2p_ll = 1;          // LKM_P2P_LL=0: finalize_cl_kernel exchanges with plain stores + system fence + flags
  unsigned long long p2p_epoch = 0;
  int p2p_enabled = 1;  // LKM_P2P=0 falls back to ncclAllGather
  uint32_t launches = 0;
  // Hamerly (lkm_hamerly.cuh): per-row bounds, list of rows to scan, fixed-point running sums, per-centroid moves
  DevBuf ham_ub, ham_lb, ham_list, ham_fx, ham_moved, ham_s, ham_state, ham_red, absmax;
  uint64_t absmax_epoch = 0;
  double absmax_value = 0;
  int last_fin_src = -1, last_fin_src_inertia = 0;   // partial sources of the last Lloyd iteration's finalize
  unsigned long long ham_rows_scanned = 0, ham_rows_tightened = 0;
};

namespace lkm {

#define CK(call)                                                                                      \
  do {                                                                                                \
    cudaError_t e__ = (call);                                                                         \
    if (e__ != cudaSuccess) {                                                                         \
      ctx->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                 \
      return LKM_ERR_CUDA;                                                                            \
    }                                                                                                 \
  } while (0)
#define CKS(call)                                                                                     \
  do {                                                                                                \
    lkm_status s__ = (call);                                                                          \
    if (s__ != LKM_OK) return s__;                                                                    \
  } while (0)

static inline lkm_status fail(lkm_ctx* ctx, lkm_status s, const char* msg) {
  ctx->err = msg;
  return s;
}

}  // namespace lkm

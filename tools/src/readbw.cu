// Read-only HBM bandwidth by access method (B200): LDG.128, 1-D bulk copies, TMA tiles.  Build:
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tools/bin/readbw tools/src/readbw.cu
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstring>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t c) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(c) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory"); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t ph) {
  uint32_t ok = 0;
  while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(ph) : "memory");
}

__global__ void __launch_bounds__(512) ldg_kernel(const float4* __restrict__ x, size_t n4, float* out) {
  float s = 0.f;
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  for (; i + 7 * stride < n4; i += 8 * stride) {
    float4 v[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) v[q] = __ldg(x + i + q * stride);
#pragma unroll
    for (int q = 0; q < 8; ++q) s += v[q].x + v[q].w;
  }
  for (; i < n4; i += stride) s += __ldg(x + i).x;
  if (s == 12345.678f) out[0] = s;
}

// mode 0: 1-D bulk copies of `chunk` bytes; mode 1: TMA 2-D box (one op of `chunk` bytes per stage)
template <int MODE>
__global__ void __launch_bounds__(160) ring_kernel(const unsigned char* __restrict__ x, size_t bytes, int chunk, int stages,
                                                    const __grid_constant__ CUtensorMap tmap, int box_rows, int n_col_boxes, float* out) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[16], empty[16];
  const int t = threadIdx.x, warp = t >> 5, lane = t & 31;
  if (t == 0) { for (int s = 0; s < 16; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 4); } asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
  __syncthreads();
  const size_t n_chunks = bytes / chunk;
  const size_t per = (n_chunks + gridDim.x - 1) / gridDim.x;
  const size_t j0 = per * blockIdx.x < n_chunks ? per * blockIdx.x : n_chunks, j1 = j0 + per < n_chunks ? j0 + per : n_chunks;
  float acc = 0.f;
  if (warp == 4) {
    if (lane == 0) {
      int s = 0; uint32_t use = 0;
      for (size_t j = j0; j < j1; ++j) {
        if (use > 0) mbar_wait(&empty[s], (use - 1) & 1);
        mbar_expect(&full[s], (uint32_t)chunk);
        if (MODE == 0) {
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                       ::"r"(smem_u32(smem + (size_t)s * chunk)), "l"(x + j * (size_t)chunk), "r"((uint32_t)chunk), "r"(smem_u32(&full[s])) : "memory");
        } else {
          // chunk j = (row block j / n_col_boxes, column box j % n_col_boxes)
          const int c0 = (int)(j % n_col_boxes) * 32, c1 = (int)(j / n_col_boxes) * box_rows;
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                       ::"r"(smem_u32(smem + (size_t)s * chunk)), "l"(&tmap), "r"(smem_u32(&full[s])), "r"(c0), "r"(c1) : "memory");
        }
        if (++s == stages) { s = 0; ++use; }
      }
    }
  } else {
    int s = 0; uint32_t ph = 0;
    for (size_t j = j0; j < j1; ++j) {
      mbar_wait(&full[s], ph);
      acc += reinterpret_cast<const float*>(smem + (size_t)s * chunk)[t];
      __syncwarp();
      if (lane == 0) mbar_arrive(&empty[s]);
      if (++s == stages) { s = 0; ph ^= 1; }
    }
  }
  if (acc == 12345.678f) out[0] = acc;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  const size_t bytes = (size_t)4 << 30;
  unsigned char* x; float* out;
  CK(cudaMalloc(&x, bytes)); CK(cudaMalloc(&out, 4));
  CK(cudaMemset(x, 1, bytes));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  void* fn = nullptr; cudaDriverEntryPointQueryResult q;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q));
  EncodeTiledFn encode = (EncodeTiledFn)fn;
  auto report = [&](const char* name, float ms) { printf("%-64s %8.3f ms  %7.1f GB/s\n", name, ms, bytes / ms / 1e6); };
  float ms;
  for (int occ : {2, 4}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaEventRecord(e0);
      ldg_kernel<<<148 * occ, 512>>>((const float4*)x, bytes / 16, out);
      cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
    }
    char nm[96]; snprintf(nm, 96, "LDG.128 x8, %d CTAs of 512 per SM", occ); report(nm, ms);
  }
  CUtensorMap dummy; memset(&dummy, 0, sizeof(dummy));
  CK(cudaFuncSetAttribute(ring_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(ring_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  for (int chunk : {4096, 16384, 32768}) {
    for (int stages : {4, 8, 12}) {
      if ((size_t)chunk * stages > 196608) continue;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        ring_kernel<0><<<148, 160, chunk * stages>>>(x, bytes, chunk, stages, dummy, 0, 1, out);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
      }
      char nm[96]; snprintf(nm, 96, "bulk 1-D %5d B x %2d stages, 1 CTA/SM", chunk, stages); report(nm, ms);
    }
  }
  struct Cfg { int d, box_cols, box_rows; CUtensorMapSwizzle sw; const char* name; };
  const Cfg cfgs[] = {
      {32, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, "TMA d=32  box 32x128 SW128 (16 KB, contiguous)"},
      {32, 32, 128, CU_TENSOR_MAP_SWIZZLE_NONE, "TMA d=32  box 32x128 no swizzle"},
      {32, 32, 256, CU_TENSOR_MAP_SWIZZLE_128B, "TMA d=32  box 32x256 SW128 (32 KB)"},
      {128, 32, 128, CU_TENSOR_MAP_SWIZZLE_128B, "TMA d=128 box 32x128 SW128 (16 KB, 512 B stride)"},
      {128, 32, 256, CU_TENSOR_MAP_SWIZZLE_128B, "TMA d=128 box 32x256 SW128 (32 KB, 512 B stride)"},
      {128, 128, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "TMA d=128 box 128x32 no swizzle (16 KB, full rows)"},
      {128, 32, 32, CU_TENSOR_MAP_SWIZZLE_NONE, "TMA d=128 box 32x32 no swizzle (4 KB)"},
  };
  for (const Cfg& c : cfgs) {
    const size_t n = bytes / (c.d * 4);
    CUtensorMap map;
    const cuuint64_t dims[2] = {(cuuint64_t)c.d, (cuuint64_t)n};
    const cuuint64_t strides[1] = {(cuuint64_t)c.d * 4};
    const cuuint32_t box[2] = {(cuuint32_t)c.box_cols, (cuuint32_t)c.box_rows};
    const cuuint32_t estr[2] = {1, 1};
    if (encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, c.sw,
               CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS) { printf("encode failed: %s\n", c.name); continue; }
    const int chunk = c.box_cols * c.box_rows * 4;
    for (int stages : {4, 8}) {
      if ((size_t)chunk * stages > 196608) continue;
      for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0);
        ring_kernel<1><<<148, 160, chunk * stages>>>(x, bytes, chunk, stages, map, c.box_rows, c.d / c.box_cols, out);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
      }
      char nm[128]; snprintf(nm, 128, "%s x %d stages", c.name, stages); report(nm, ms);
    }
  }
  return 0;
}

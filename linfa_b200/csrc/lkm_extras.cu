// Second translation unit of liblkm.so: the callers and data formats either side of the K-Means path
// (SURVEY 8(f) "next" rows): silhouette score (src/metrics_clustering.rs:65-135) and .npy files
// (algorithms/linfa-clustering/examples/kmeans.rs:40-42, ndarray-npy `write_npy`).
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "lkm_ctx.hpp"
#include "lkm_metrics.cuh"

#define LKM_EXPORT extern "C" __attribute__((visibility("default")))

using namespace lkm;

namespace lkm {

template <typename F>
static lkm_status silhouette_impl(lkm_ctx* ctx, const F* x, uint64_t n, uint64_t d, int64_t row_stride, const uint64_t* labels,
                                  F* score_out) {
  if (!x || !labels || !score_out) return fail(ctx, LKM_ERR_INVALID_ARG, "null argument");
  if (n == 0 || d == 0) return fail(ctx, LKM_ERR_SHAPE, "empty dataset");
  if (row_stride < (int64_t)d) return fail(ctx, LKM_ERR_INVALID_ARG, "row stride smaller than the row");
  CK(cudaSetDevice(ctx->device));
  // label_count (src/dataset/impl_dataset.rs): distinct labels -> dense ids in order of first appearance is irrelevant for
  // the score (min over the other clusters); ids here follow the sorted label values
  std::vector<uint64_t> order(n);
  std::iota(order.begin(), order.end(), 0ull);
  std::stable_sort(order.begin(), order.end(), [&](uint64_t p, uint64_t q) { return labels[p] < labels[q]; });
  std::vector<unsigned long long> off;
  std::vector<uint32_t> cluster_of(n);
  for (uint64_t i = 0; i < n; ++i) {
    if (i == 0 || labels[order[i]] != labels[order[i - 1]]) off.push_back(i);
    cluster_of[i] = (uint32_t)(off.size() - 1);
  }
  const int n_clusters = (int)off.size();
  off.push_back(n);
  // "Single label dataset, all points are in the same cluster" (:77-80)
  if (n_clusters == 1) { *score_out = (F)1; return LKM_OK; }
  std::vector<F> xs((size_t)n * d);
  for (uint64_t i = 0; i < n; ++i) std::memcpy(&xs[i * d], x + order[i] * (uint64_t)row_stride, d * sizeof(F));
  const size_t xi_bytes = (size_t)kSilThreads * (d | 1) * sizeof(F);
  const size_t limit = ctx->smem_optin - 1024;
  if (xi_bytes + d * sizeof(F) > limit) return fail(ctx, LKM_ERR_UNSUPPORTED, "row too wide for the silhouette tile");
  const int tj = (int)std::min<size_t>(256, (limit - xi_bytes) / (d * sizeof(F)));
  const size_t smem = xi_bytes + (size_t)tj * d * sizeof(F);
  DevBuf dx, doff, dcl, dsc;
  lkm_status st = LKM_OK;
  auto run = [&]() -> lkm_status {
    CK(dx.ensure(xs.size() * sizeof(F)));
    CK(doff.ensure(off.size() * 8));
    CK(dcl.ensure(n * 4));
    CK(dsc.ensure(n * sizeof(F)));
    CK(cudaMemcpyAsync(dx.p, xs.data(), xs.size() * sizeof(F), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(doff.p, off.data(), off.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dcl.p, cluster_of.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream));
    SilArgs<F> a{};
    a.X = dx.as<F>(); a.n = n; a.d = (int)d; a.n_clusters = n_clusters; a.off = doff.as<unsigned long long>();
    a.cluster_of = dcl.as<uint32_t>(); a.tj = tj; a.scores = dsc.as<F>();
    CK(cudaFuncSetAttribute(silhouette_kernel<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    silhouette_kernel<F><<<(unsigned)((n + kSilThreads - 1) / kSilThreads), kSilThreads, smem, ctx->stream>>>(a);
    CK(cudaGetLastError());
    ctx->launches++;
    std::vector<F> sc(n), unsorted(n);
    CK(cudaMemcpyAsync(sc.data(), dsc.p, n * sizeof(F), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    // `.sum::<F>()` over sample_iter (:128): sample order, in F
    for (uint64_t i = 0; i < n; ++i) unsorted[order[i]] = sc[i];
    F total = (F)0;
    for (uint64_t i = 0; i < n; ++i) total = total + unsorted[i];
    *score_out = total / (F)n;
    return LKM_OK;
  };
  st = run();
  dx.release(); doff.release(); dcl.release(); dsc.release();
  return st;
}

// ---- .npy (format 1.0, little endian, C order) ------------------------------------------------
static const char* npy_descr(int dtype) { return dtype == 0 ? "<f4" : dtype == 1 ? "<f8" : dtype == 2 ? "<u8" : nullptr; }
static size_t npy_itemsize(int dtype) { return dtype == 0 ? 4 : 8; }

}  // namespace lkm

LKM_EXPORT lkm_status lkm_silhouette_score_f32(lkm_ctx* ctx, const float* x, uint64_t n, uint64_t d, int64_t row_stride,
                                               const uint64_t* labels, float* score_out) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  return silhouette_impl<float>(ctx, x, n, d, row_stride, labels, score_out);
}
LKM_EXPORT lkm_status lkm_silhouette_score_f64(lkm_ctx* ctx, const double* x, uint64_t n, uint64_t d, int64_t row_stride,
                                               const uint64_t* labels, double* score_out) {
  if (!ctx) return LKM_ERR_INVALID_ARG;
  return silhouette_impl<double>(ctx, x, n, d, row_stride, labels, score_out);
}

LKM_EXPORT lkm_status lkm_write_npy(const char* path, int dtype, const void* data, int ndim, const uint64_t* shape) {
  if (!path || !data || ndim < 0 || ndim > 8 || (ndim > 0 && !shape) || !npy_descr(dtype)) return LKM_ERR_INVALID_ARG;
  std::string dict = std::string("{'descr': '") + npy_descr(dtype) + "', 'fortran_order': False, 'shape': (";
  size_t count = 1;
  for (int i = 0; i < ndim; ++i) {
    dict += std::to_string((unsigned long long)shape[i]);
    if (ndim == 1 || i + 1 < ndim) dict += ",";
    if (i + 1 < ndim) dict += " ";
    count *= (size_t)shape[i];
  }
  dict += "), }";
  // magic (6) + version (2) + header length (2) + dict + padding + '\n' is a multiple of 64
  size_t total = 10 + dict.size() + 1;
  const size_t pad = (64 - total % 64) % 64;
  dict.append(pad, ' ');
  dict += "\n";
  if (dict.size() > 65535) return LKM_ERR_INVALID_ARG;
  FILE* f = std::fopen(path, "wb");
  if (!f) return LKM_ERR_INVALID_ARG;
  const unsigned char head[10] = {0x93, 'N', 'U', 'M', 'P', 'Y', 1, 0, (unsigned char)(dict.size() & 255), (unsigned char)(dict.size() >> 8)};
  bool ok = std::fwrite(head, 1, 10, f) == 10 && std::fwrite(dict.data(), 1, dict.size(), f) == dict.size();
  const size_t bytes = count * npy_itemsize(dtype);
  ok = ok && (bytes == 0 || std::fwrite(data, 1, bytes, f) == bytes);
  ok = (std::fclose(f) == 0) && ok;
  return ok ? LKM_OK : LKM_ERR_INVALID_ARG;
}

// reads the header: dtype code, ndim, shape (up to 8 dims); data_offset = first byte of the array
LKM_EXPORT lkm_status lkm_read_npy_header(const char* path, int* dtype_out, int* ndim_out, uint64_t* shape_out, uint64_t* data_offset_out) {
  if (!path || !dtype_out || !ndim_out || !shape_out || !data_offset_out) return LKM_ERR_INVALID_ARG;
  FILE* f = std::fopen(path, "rb");
  if (!f) return LKM_ERR_INVALID_ARG;
  unsigned char head[12];
  lkm_status st = LKM_ERR_INVALID_ARG;
  do {
    if (std::fread(head, 1, 10, f) != 10 || std::memcmp(head, "\x93NUMPY", 6) != 0) break;
    size_t hlen = head[8] | ((size_t)head[9] << 8), hoff = 10;
    if (head[6] >= 2) {   // versions 2.0 / 3.0: four-byte header length
      if (std::fread(head + 10, 1, 2, f) != 2) break;
      hlen |= ((size_t)head[10] << 16) | ((size_t)head[11] << 24);
      hoff = 12;
    }
    std::string dict(hlen, '\0');
    if (std::fread(&dict[0], 1, hlen, f) != hlen) break;
    const size_t dp = dict.find("'descr'");
    if (dp == std::string::npos) break;
    const size_t q0 = dict.find('\'', dp + 7), q1 = q0 == std::string::npos ? q0 : dict.find('\'', q0 + 1);
    if (q1 == std::string::npos) break;
    const std::string descr = dict.substr(q0 + 1, q1 - q0 - 1);
    int dt = -1;
    if (descr == "<f4") dt = 0; else if (descr == "<f8") dt = 1; else if (descr == "<u8") dt = 2;
    if (dt < 0) { st = LKM_ERR_UNSUPPORTED; break; }
    if (dict.find("'fortran_order': False") == std::string::npos) { st = LKM_ERR_UNSUPPORTED; break; }
    const size_t sp = dict.find("'shape'");
    const size_t p0 = sp == std::string::npos ? sp : dict.find('(', sp), p1 = p0 == std::string::npos ? p0 : dict.find(')', p0);
    if (p1 == std::string::npos) break;
    int nd = 0;
    bool bad = false;
    size_t i = p0 + 1;
    while (i < p1) {
      while (i < p1 && (dict[i] == ' ' || dict[i] == ',')) ++i;
      if (i >= p1) break;
      if (dict[i] < '0' || dict[i] > '9' || nd == 8) { bad = true; break; }
      unsigned long long v = 0;
      while (i < p1 && dict[i] >= '0' && dict[i] <= '9') v = v * 10 + (unsigned)(dict[i++] - '0');
      shape_out[nd++] = v;
    }
    if (bad) break;
    *dtype_out = dt; *ndim_out = nd; *data_offset_out = hoff + hlen;
    st = LKM_OK;
  } while (0);
  std::fclose(f);
  return st;
}

LKM_EXPORT lkm_status lkm_read_npy_data(const char* path, uint64_t data_offset, void* out, uint64_t bytes) {
  if (!path || (!out && bytes)) return LKM_ERR_INVALID_ARG;
  FILE* f = std::fopen(path, "rb");
  if (!f) return LKM_ERR_INVALID_ARG;
  bool ok = std::fseek(f, (long)data_offset, SEEK_SET) == 0 && (bytes == 0 || std::fread(out, 1, bytes, f) == bytes);
  std::fclose(f);
  return ok ? LKM_OK : LKM_ERR_INVALID_ARG;
}

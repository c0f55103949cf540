/* lkm.h — C ABI of the B200-native K-Means hot path ("linfa k-means", lkm).
 *
 * This is the drop-in boundary for linfa-clustering's K-Means.  The reference
 * (rust-ml/linfa v0.8.1) has no FFI of its own: the path sits behind the linfa
 * trait system.  Each entry point below names the reference item it replaces;
 * the Rust crate in rust/linfa-clustering-b200 (and the ctypes binding in
 * linfa_b200/_ffi.py) bind exactly these symbols.  Paths are relative to the
 * reference root; KM = algorithms/linfa-clustering/src/k_means.
 *
 * Conventions: every function returns an lkm_status (0 = OK) and never throws
 * or unwinds.  All pointers are plain host pointers unless the name says
 * `_device`.  The caller owns every output buffer; the library never keeps a
 * host pointer after the call returns.  One context drives one GPU (one
 * process per GPU); a context may be used from one thread at a time.
 * F is float (`_f32`) or double (`_f64`) — linfa::Float is implemented for
 * exactly these two (src/dataset/mod.rs:68-74).  Matrices are row-major.
 */
#ifndef LKM_H
#define LKM_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct lkm_ctx lkm_ctx;

/* Status codes.  10..14 mirror KMeansParamsError (KM/errors.rs:5-19),
 * 1 mirrors KMeansError::InertiaError (KM/errors.rs:27-28), 20 mirrors
 * IncrKMeansError::NotConverged (KM/errors.rs:40-42).  3..6 are the places
 * where the reference panics (assert!/expect/unwrap). */
typedef enum lkm_status {
  LKM_OK = 0,
  LKM_ERR_INERTIA = 1,        /* no run produced a finite inertia (KM/algorithm.rs:336,365) */
  LKM_ERR_INVALID_ARG = 2,
  LKM_ERR_SAMPLE_AMOUNT = 3,  /* Random init with k > n: rand::seq::index::sample panics */
  LKM_ERR_ZERO_WEIGHTS = 4,   /* assert_ne!(weights.sum(), 0) (KM/init.rs:127) */
  LKM_ERR_BAD_WEIGHTS = 5,    /* WeightedIndex::new(..).expect("invalid weights") (KM/init.rs:131-133) */
  LKM_ERR_SHAPE = 6,          /* shape asserts: Precomputed (KM/init.rs:96-97), predict (KM/algorithm.rs:744-748) */
  LKM_ERR_N_CLUSTERS = 10,    /* KMeansParamsError::NClusters  (KM/hyperparams.rs:125) */
  LKM_ERR_N_RUNS = 11,        /* KMeansParamsError::NRuns      (KM/hyperparams.rs:127) */
  LKM_ERR_TOLERANCE = 12,     /* KMeansParamsError::Tolerance  (KM/hyperparams.rs:129) */
  LKM_ERR_MAX_ITERATIONS = 13,/* KMeansParamsError::MaxIterations (KM/hyperparams.rs:131) */
  LKM_ERR_INCREMENTAL_HAMERLY = 14, /* KMeansParamsError::IncrementalHamerly (KM/algorithm.rs:640-644) */
  LKM_NOT_CONVERGED = 20,     /* fit_with: Err(NotConverged(model)); outputs ARE written */
  LKM_ERR_CUDA = 30,
  LKM_ERR_NCCL = 31,
  LKM_ERR_NO_DATA = 32,       /* fit/init called before set_data / gen_blobs */
  LKM_ERR_UNSUPPORTED = 33
} lkm_status;

/* KMeansInit (KM/init.rs:22-35) */
typedef enum lkm_init_method {
  LKM_INIT_RANDOM = 0,
  LKM_INIT_PRECOMPUTED = 1,
  LKM_INIT_KMEANS_PLUSPLUS = 2,
  LKM_INIT_KMEANS_PARA = 3
} lkm_init_method;

/* KMeansAlgorithm (KM/init.rs:54-79).  LKM_ALGO_HAMERLY runs fit_hamerly
 * (KM/algorithm.rs:248-291): per-row upper / lower bounds decide which rows are
 * looked at after the first iteration, inertia is compute_inertia against the
 * final centroids (:275,570-577) and cluster_count comes from the best run's
 * memberships (:281-291).  Row-sharded (multi-GPU) runs execute Lloyd's
 * iterations and return Hamerly's result fields. */
typedef enum lkm_algorithm { LKM_ALGO_LLOYD = 0, LKM_ALGO_HAMERLY = 1 } lkm_algorithm;

/* How the D^2 / weighted pick of k-means++ walks the prefix sum.
 * EXACT reproduces rand 0.8.5 WeightedIndex bit for bit (sequential prefix sum
 * in F on one GPU thread) — same row indices as the reference for the same
 * RNG.  FAST uses a fixed-order parallel f64 scan with the same uniform draw. */
typedef enum lkm_pick_mode { LKM_PICK_EXACT = 0, LKM_PICK_FAST = 1, LKM_PICK_AUTO = 2 } lkm_pick_mode;
/* Cost of EXACT: every pick walks the n weights on ONE GPU thread (that is what makes the prefix sum the reference's,
 * bit for bit): about 1 ms per million weights and pick, i.e. k * n / 1e9 seconds per seeding — fine for the reference's
 * own problem sizes, hours for 100M rows x 1024 clusters.  AUTO uses EXACT up to 2^22 weights and FAST above; the
 * bindings default to EXACT so that small fits reproduce linfa's seeds, and say so in their docs. */

/* Trampolines into the caller's `R: Rng` (KM/hyperparams.rs:37 `rng: R`).
 * The reference clones the rng at the top of every fit (KM/algorithm.rs:298);
 * the binding clones before the call and passes the clone here. */
typedef struct lkm_rng {
  void* user;
  uint32_t (*next_u32)(void* user);
  uint64_t (*next_u64)(void* user);
} lkm_rng;

/* KMeansValidParams (KM/hyperparams.rs:19-42); defaults at :71-82 are
 * n_runs 10, tolerance 1e-4, max_n_iterations 300, init KMeansPlusPlus. */
typedef struct lkm_params {
  uint64_t n_clusters;
  uint64_t n_runs;
  uint64_t max_n_iterations;
  double tolerance;          /* cast to F inside, like F::cast(1e-4) */
  int32_t init_method;       /* lkm_init_method */
  int32_t algorithm;         /* lkm_algorithm */
  int32_t pick_mode;         /* lkm_pick_mode (not in the reference; default EXACT = 0) */
  int32_t reserved;
  const void* precomputed;   /* KMeansInit::Precomputed: k x d of F, host, row-major */
  uint64_t precomputed_rows; /* shape check against n_clusters / d (KM/init.rs:96-97) */
  uint64_t precomputed_cols;
} lkm_params;

/* Timings and counters of the last lkm_lloyd_iters_* / lkm_fit_* call. */
typedef struct lkm_stats {
  uint64_t n_iter;           /* Lloyd iterations executed (last run) */
  double last_shift;         /* ||C_t - C_{t+1}||_F of the last iteration */
  double inertia_total;      /* sum of min squared distances of the last assignment */
  float device_ms;           /* CUDA-event time of the iteration loop on the context stream */
  uint32_t kernel_launches;  /* kernels this library launched in the call */
  uint32_t assign_path;      /* 0 exact CUDA-core, 1 certified FMA, 2 tcgen05 3xTF32, 3 f32-screened f64 (small d), 4 f64 DMMA */
  uint64_t fallback_rows;    /* rows re-evaluated in exact form by the certified paths */
  float main_kernel_ms;      /* with lkm_set_l2_flush: summed duration of the assign/fused kernel alone */
  uint32_t update_path;      /* how the last iteration summed the rows: 0 inside the assignment kernel, 1 streaming update pass, 2 sorted update (unordered rows) */
  uint64_t rows_scanned;     /* Hamerly: rows that went through two_closest_centroids after iteration 1 (summed over iterations) */
  uint64_t rows_tightened;   /* Hamerly: rows whose upper bound was recomputed (one exact distance) */
} lkm_stats;

/* -- context ------------------------------------------------------------- */
lkm_status lkm_create(lkm_ctx** out, int device);
lkm_status lkm_destroy(lkm_ctx* ctx);
const char* lkm_last_error(const lkm_ctx* ctx);
/* the CUDA stream (cudaStream_t) every kernel of this context is launched on */
void* lkm_stream(lkm_ctx* ctx);
lkm_status lkm_synchronize(lkm_ctx* ctx);
const char* lkm_version(void);

/* -- multi-GPU (new; the reference is single-process).  Rows are sharded,
 * every rank holds full centroids, and each Lloyd iteration all-gathers the
 * per-rank (k*d sums, k counts, inertia) partials and reduces them in rank
 * order.  id128 is an ncclUniqueId (128 bytes) made on rank 0 and shipped by
 * the caller (torch.distributed / MPI / a file). */
lkm_status lkm_comm_unique_id(void* id128);
lkm_status lkm_comm_init(lkm_ctx* ctx, const void* id128, int rank, int world_size);
/* global row offset / global row count of this rank's shard, used by seeding
 * (row indices drawn by the RNG are global) and by inertia = total / n_global */
lkm_status lkm_set_shard(lkm_ctx* ctx, uint64_t row_offset, uint64_t n_global);

/* -- several devices from ONE process: a device group.  SURVEY 8(b) asked for lkm_create(device_ids, n_devices) so
 * that one `fit(&dataset)` call of the Rust binding can use the whole box.  Rows are split into contiguous shards (rank
 * r gets rows [r n / G, (r + 1) n / G)), one context and one host thread per device; every Lloyd iteration exchanges the
 * per-device partials over NVLink peer access exactly as the one-process-per-GPU mode does, the seeding rendezvous
 * happens on the host, and every rank sees the same random draws (the caller's rng is read once per draw).  The result
 * equals the single-device fit up to the summation order of the partials (it is identical on every rank). */
typedef struct lkm_multi lkm_multi;
lkm_status lkm_create_multi(lkm_multi** out, const int* device_ids, int n_devices);
lkm_status lkm_destroy_multi(lkm_multi* m);
const char* lkm_multi_last_error(const lkm_multi* m);
int lkm_multi_size(const lkm_multi* m);
lkm_ctx* lkm_multi_ctx(lkm_multi* m, int i);   /* rank i's context, for the per-device knobs */
lkm_status lkm_multi_set_data_f32(lkm_multi* m, const float* x, uint64_t n, uint64_t d, int64_t row_stride);
lkm_status lkm_multi_set_data_f64(lkm_multi* m, const double* x, uint64_t n, uint64_t d, int64_t row_stride);
/* Fit::fit over the sharded matrix; outputs as lkm_fit_* */
lkm_status lkm_multi_fit_f32(lkm_multi* m, const lkm_params* p, lkm_rng rng, float* centroids_out, float* cluster_count_out,
                             float* inertia_out, lkm_stats* stats_out);
lkm_status lkm_multi_fit_f64(lkm_multi* m, const lkm_params* p, lkm_rng rng, double* centroids_out, double* cluster_count_out,
                             double* inertia_out, lkm_stats* stats_out);
/* update_memberships_and_dists with the rows split over the devices (no exchange); either output may be NULL */
lkm_status lkm_multi_assign_f32(lkm_multi* m, const float* centroids, uint64_t k, const float* x, uint64_t n, uint64_t d,
                                int64_t row_stride, uint64_t* labels_out, float* min_sq_dist_out);
lkm_status lkm_multi_assign_f64(lkm_multi* m, const double* centroids, uint64_t k, const double* x, uint64_t n, uint64_t d,
                                int64_t row_stride, uint64_t* labels_out, double* min_sq_dist_out);

/* -- observations: `dataset.records()` (KM/algorithm.rs:299) -------------- */
/* copies a host matrix (any row stride, like an ndarray view) to the device */
lkm_status lkm_set_data_f32(lkm_ctx* ctx, const float* x, uint64_t n, uint64_t d, int64_t row_stride);
lkm_status lkm_set_data_f64(lkm_ctx* ctx, const double* x, uint64_t n, uint64_t d, int64_t row_stride);
/* borrows a device-resident contiguous n x d matrix (no copy) */
lkm_status lkm_set_data_device_f32(lkm_ctx* ctx, const void* x_device, uint64_t n, uint64_t d);
lkm_status lkm_set_data_device_f64(lkm_ctx* ctx, const void* x_device, uint64_t n, uint64_t d);
/* device-side linfa_datasets::generate::blobs (datasets/src/generate.rs:28-58):
 * blob b owns rows [b*m, (b+1)*m) of the GLOBAL matrix (m = n_global / n_blobs,
 * the last blob takes the remainder), row = centre_b + sigma * N(0, I) from a
 * counter-based generator keyed by (seed, global row, feature).  This rank
 * materialises global rows [row_offset, row_offset + n). */
lkm_status lkm_gen_blobs_f32(lkm_ctx* ctx, uint64_t n, uint64_t d, uint64_t n_blobs, const float* centres,
                             float sigma, uint64_t seed, uint64_t row_offset, uint64_t n_global);
lkm_status lkm_gen_blobs_f64(lkm_ctx* ctx, uint64_t n, uint64_t d, uint64_t n_blobs, const double* centres,
                             double sigma, uint64_t seed, uint64_t row_offset, uint64_t n_global);
/* copy rows [row0, row0+rows) of the resident matrix back to the host */
lkm_status lkm_get_data_f32(lkm_ctx* ctx, uint64_t row0, uint64_t rows, float* out);
lkm_status lkm_get_data_f64(lkm_ctx* ctx, uint64_t row0, uint64_t rows, double* out);

/* -- Fit::fit -> fit_lloyd + get_kmeans_result (KM/algorithm.rs:294-367,370-388)
 * centroids_out k*d, cluster_count_out k (as F, from the LAST run's final
 * assignment — KM/algorithm.rs:304,342,355-357), inertia_out = best total / n. */
lkm_status lkm_fit_f32(lkm_ctx* ctx, const lkm_params* p, lkm_rng rng, float* centroids_out,
                       float* cluster_count_out, float* inertia_out, lkm_stats* stats_out);
lkm_status lkm_fit_f64(lkm_ctx* ctx, const lkm_params* p, lkm_rng rng, double* centroids_out,
                       double* cluster_count_out, double* inertia_out, lkm_stats* stats_out);

/* -- KMeansInit::run (KM/init.rs:83-101): seeding only -------------------- */
lkm_status lkm_init_f32(lkm_ctx* ctx, const lkm_params* p, lkm_rng rng, float* centroids_out);
lkm_status lkm_init_f64(lkm_ctx* ctx, const lkm_params* p, lkm_rng rng, double* centroids_out);

/* -- the loop body of fit_lloyd (KM/algorithm.rs:314-331) as a benchmark /
 * resume entry: runs Lloyd iterations from centroids_inout until the shift
 * drops below `tolerance` or `max_iters` iterations ran.  tolerance = 0 runs
 * exactly max_iters iterations (a shift is never < 0).  cluster_count_out
 * and inertia_total_out (sum, not mean) may be NULL. */
lkm_status lkm_lloyd_iters_f32(lkm_ctx* ctx, float* centroids_inout, uint64_t k, uint64_t max_iters,
                               double tolerance, float* cluster_count_out, float* inertia_total_out,
                               lkm_stats* stats_out);
lkm_status lkm_lloyd_iters_f64(lkm_ctx* ctx, double* centroids_inout, uint64_t k, uint64_t max_iters,
                               double tolerance, double* cluster_count_out, double* inertia_total_out,
                               lkm_stats* stats_out);

/* The loop body of fit_hamerly (KM/algorithm.rs:262-279) from centroids_inout: same arguments as lkm_lloyd_iters_*;
 * inertia_total_out is compute_inertia against the RETURNED centroids (sum, not mean).  stats.rows_scanned /
 * rows_tightened say how much of the matrix the bounds let the iterations skip. */
lkm_status lkm_hamerly_iters_f32(lkm_ctx* ctx, float* centroids_inout, uint64_t k, uint64_t max_iters,
                                 double tolerance, float* cluster_count_out, float* inertia_total_out,
                                 lkm_stats* stats_out);
lkm_status lkm_hamerly_iters_f64(lkm_ctx* ctx, double* centroids_inout, uint64_t k, uint64_t max_iters,
                                 double tolerance, double* cluster_count_out, double* inertia_total_out,
                                 lkm_stats* stats_out);

/* -- PredictInplace (KM/algorithm.rs:735-777) -> update_cluster_memberships
 * (:838-849) and Transformer (:718-733) -> update_min_dists (:852-863).
 * x == NULL means "the resident matrix" (n and d must match it). */
lkm_status lkm_predict_f32(lkm_ctx* ctx, const float* centroids, uint64_t k, const float* x, uint64_t n,
                           uint64_t d, int64_t row_stride, uint64_t* labels_out);
lkm_status lkm_predict_f64(lkm_ctx* ctx, const double* centroids, uint64_t k, const double* x, uint64_t n,
                           uint64_t d, int64_t row_stride, uint64_t* labels_out);
lkm_status lkm_transform_f32(lkm_ctx* ctx, const float* centroids, uint64_t k, const float* x, uint64_t n,
                             uint64_t d, int64_t row_stride, float* min_sq_dist_out);
lkm_status lkm_transform_f64(lkm_ctx* ctx, const double* centroids, uint64_t k, const double* x, uint64_t n,
                             uint64_t d, int64_t row_stride, double* min_sq_dist_out);
/* update_memberships_and_dists (KM/algorithm.rs:866-881): both at once */
lkm_status lkm_assign_f32(lkm_ctx* ctx, const float* centroids, uint64_t k, const float* x, uint64_t n,
                          uint64_t d, int64_t row_stride, uint64_t* labels_out, float* min_sq_dist_out);
lkm_status lkm_assign_f64(lkm_ctx* ctx, const double* centroids, uint64_t k, const double* x, uint64_t n,
                          uint64_t d, int64_t row_stride, uint64_t* labels_out, double* min_sq_dist_out);

/* -- FitWith::fit_with (KM/algorithm.rs:635-715): one mini-batch step on the
 * resident matrix.  has_model = 0 initialises from params (best of n_runs
 * seedings by dists.sum(), :659-678); the buffers are then n_clusters x d and
 * n_clusters.  has_model = 1: centroids_inout is model_rows x model_cols and
 * cluster_count_inout has model_rows entries — like the reference, the model's
 * own shape is used and p->n_clusters is not looked at; model_cols must equal
 * the batch's n_features (LKM_ERR_SHAPE otherwise).  Parameters are checked
 * first (the ParamGuard blanket impl, src/param_guard.rs:75-86), then Hamerly
 * is rejected (:640-644).  Returns LKM_OK when the shift is below tolerance,
 * LKM_NOT_CONVERGED otherwise; outputs are written in both cases. */
lkm_status lkm_fit_with_f32(lkm_ctx* ctx, const lkm_params* p, lkm_rng rng, int has_model, uint64_t model_rows,
                            uint64_t model_cols, float* centroids_inout, float* cluster_count_inout,
                            float* inertia_out);
lkm_status lkm_fit_with_f64(lkm_ctx* ctx, const lkm_params* p, lkm_rng rng, int has_model, uint64_t model_rows,
                            uint64_t model_cols, double* centroids_inout, double* cluster_count_inout,
                            double* inertia_out);

/* -- SilhouetteScore::silhouette_score (src/metrics_clustering.rs:65-135) ---
 * mean over the samples of (b - a) / max(a, b), a = mean distance to the other
 * samples of the own cluster, b = smallest mean distance to another cluster;
 * 1 when there is a single label (:77-80).  labels: any u64 values (the
 * `usize` targets KMeans::predict produces).  O(n^2 d) on the device; the
 * per-sample scores follow the reference's arithmetic operation by operation. */
lkm_status lkm_silhouette_score_f32(lkm_ctx* ctx, const float* x, uint64_t n, uint64_t d, int64_t row_stride,
                                    const uint64_t* labels, float* score_out);
lkm_status lkm_silhouette_score_f64(lkm_ctx* ctx, const double* x, uint64_t n, uint64_t d, int64_t row_stride,
                                    const uint64_t* labels, double* score_out);

/* -- GaussianMixtureModel, full covariances (algorithms/linfa-clustering/src/gaussian_mixture/algorithm.rs): the
 * first caller of the K-Means path inside the reference (GmmInitMethod::KMeans, :137-146).  Fit::fit (:411-467) on the
 * resident matrix: initial responsibilities from a K-Means fit (KMeans::params_with_rng(k, rng) with its defaults,
 * then predict) or random, then EM — E-step (:291-298, 339-352) and M-step (:300-322) on the device, the k Cholesky
 * factorizations (:261-276) on the host.  Defaults (gaussian_mixture/hyperparams.rs:88-101): tolerance 1e-3,
 * reg_covariance 1e-6, n_runs 1, max_n_iterations 100, init KMeans.  Status codes 40..48 mirror GmmError. */
typedef struct lkm_gmm_params {
  uint64_t n_clusters;
  double tolerance;
  double reg_covariance;
  uint64_t n_runs;
  uint64_t max_n_iterations;
  int32_t init_method;        /* 0 = GmmInitMethod::KMeans, 1 = GmmInitMethod::Random */
  int32_t kmeans_pick_mode;   /* lkm_pick_mode of the K-Means initialisation (default EXACT = 0) */
} lkm_gmm_params;
enum {
  LKM_GMM_ERR_N_CLUSTERS = 40,      /* GmmError::InvalidValue, in the order of check_ref (hyperparams.rs:167-193) */
  LKM_GMM_ERR_TOLERANCE = 41,
  LKM_GMM_ERR_REG_COVARIANCE = 42,
  LKM_GMM_ERR_N_RUNS = 43,
  LKM_GMM_ERR_MAX_ITERATIONS = 44,
  LKM_GMM_ERR_EMPTY_CLUSTER = 45,   /* GmmError::EmptyCluster (algorithm.rs:224-229) */
  LKM_GMM_ERR_NOT_POSITIVE_DEFINITE = 46,   /* GmmError::LinalgError(NotPositiveDefinite) */
  LKM_GMM_ERR_NOT_CONVERGED = 47,   /* GmmError::NotConverged (algorithm.rs:459-465) */
  LKM_GMM_ERR_LOWER_BOUND = 48      /* GmmError::LowerBoundError */
};
/* outputs: weights k, means k x d, covariances / precisions / precisions_chol k x d x d; n_iter_out = the iteration at
 * which the best run converged, lower_bound_out = its mean log-likelihood (both may be NULL) */
lkm_status lkm_gmm_fit_f32(lkm_ctx* ctx, const lkm_gmm_params* p, lkm_rng rng, float* weights_out, float* means_out,
                           float* covariances_out, float* precisions_out, float* precisions_chol_out, uint64_t* n_iter_out,
                           float* lower_bound_out);
lkm_status lkm_gmm_fit_f64(lkm_ctx* ctx, const lkm_gmm_params* p, lkm_rng rng, double* weights_out, double* means_out,
                           double* covariances_out, double* precisions_out, double* precisions_chol_out, uint64_t* n_iter_out,
                           double* lower_bound_out);
/* estimate_log_prob_resp (:339-352) for a given model: labels_out = argmax of the responsibilities (PredictInplace,
 * :470-484), proba_out = n x k responsibilities (predict_proba, :205-208), log_prob_norm_mean_out = the E-step's lower
 * bound; any of the three may be NULL.  x == NULL: the resident matrix. */
lkm_status lkm_gmm_predict_f32(lkm_ctx* ctx, uint64_t k, const float* weights, const float* means, const float* precisions_chol,
                               const float* x, uint64_t n, uint64_t d, int64_t row_stride, uint64_t* labels_out,
                               float* proba_out, float* log_prob_norm_mean_out);
lkm_status lkm_gmm_predict_f64(lkm_ctx* ctx, uint64_t k, const double* weights, const double* means, const double* precisions_chol,
                               const double* x, uint64_t n, uint64_t d, int64_t row_stride, uint64_t* labels_out,
                               double* proba_out, double* log_prob_norm_mean_out);

/* -- .npy files: ndarray_npy::write_npy as used by the reference's example
 * (algorithms/linfa-clustering/examples/kmeans.rs:40-42: the records and the
 * memberships as u64).  Format 1.0, little endian, C order.  dtype: 0 = f32
 * ("<f4"), 1 = f64 ("<f8"), 2 = u64 ("<u8").  Host only, no context needed. */
lkm_status lkm_write_npy(const char* path, int dtype, const void* data, int ndim, const uint64_t* shape);
lkm_status lkm_read_npy_header(const char* path, int* dtype_out, int* ndim_out, uint64_t* shape_out /* [8] */,
                               uint64_t* data_offset_out);
lkm_status lkm_read_npy_data(const char* path, uint64_t data_offset, void* out, uint64_t bytes);

/* -- knobs that are not in the reference (all have defaults) --------------- */
/* force an assignment path: -1 auto (default), 0 exact, 1 certified FMA, 2 tcgen05, 3 f32-screened f64 (lkm_f64s.cuh),
 * 4 f64 scores on the FP64 tensor pipe (DMMA, lkm_dmma.cuh) */
lkm_status lkm_set_assign_path(lkm_ctx* ctx, int path);
/* parity-test aid: when on, every Lloyd kernel also writes the label each row was accumulated under in the LAST
 * executed iteration (the `memberships` array of KM/algorithm.rs:304,316-322, which the fused kernels otherwise never
 * materialise); lkm_get_labels copies the first n of them to the host (u32 per row, this rank's shard). */
lkm_status lkm_set_keep_labels(lkm_ctx* ctx, int on);
lkm_status lkm_get_labels(lkm_ctx* ctx, uint32_t* labels_out, uint64_t n);
/* lkm_gen_blobs_* layout: 0 (default) = the contiguous blobs of datasets/src/generate.rs:28-58, 1 = every row draws
 * its blob from a hash of its global index (unordered data; same noise) */
lkm_status lkm_set_blob_layout(lkm_ctx* ctx, int layout);
/* measurement hygiene for benchmarks: when bytes > 0 every Lloyd iteration is
 * timed with its own CUDA events and a scratch buffer of `bytes` is overwritten
 * between iterations (outside the timed spans) so that the observations are
 * read from HBM, not from L2.  0 (default) turns it off. */
lkm_status lkm_set_l2_flush(lkm_ctx* ctx, uint64_t bytes);
/* with lkm_set_l2_flush active: also time the assign/fused kernel on its own (lkm_stats.main_kernel_ms).
 * The extra event between the fused kernel and the combine kernel defeats their programmatic
 * dependent launch, so it is off by default and used only for the roofline figure. */
lkm_status lkm_set_kernel_timing(lkm_ctx* ctx, int on);
/* tensor-core scoring passes of the certified assignment kernels: 0 (default) adaptive — start with the cheapest
 * variant and switch to 3xTF32 when more than 2 % of the rows of an iteration needed the exact re-scan; 1 pins the
 * one-product screening variant (pair kernel for d in {64, 128}, pipeline kernel for d = 32), 2 the two-product
 * variant (d = 32 pipeline kernel only; the pair kernel then runs 3xTF32), 3 pins 3xTF32.  Results do not depend on it (uncertified rows are re-scanned exactly); only speed does. */
lkm_status lkm_set_scoring_passes(lkm_ctx* ctx, int passes);

#ifdef __cplusplus
}
#endif
#endif /* LKM_H */

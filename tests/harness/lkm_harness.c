/* Drives include/lkm.h the way rust/linfa-clustering-b200/src/lib.rs does — plain C (the header must stay C-clean), the
 * same call sequence and ownership rules: create -> set_data -> fit (rng through the two trampolines) -> predict ->
 * transform -> fit_with -> Hamerly -> device group -> destroy.  Prints one line per result; tests/test_harness.py compares
 * them with the Python binding's.  Usage: lkm_harness <n> <d> <k> [device]
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "lkm.h"

/* rand_xoshiro 0.6 Xoshiro256Plus seeded by SplitMix64 (seed_from_u64), as KMeans::params uses it (KM/algorithm.rs:211) */
typedef struct { uint64_t s[4]; } xoshiro;
static uint64_t splitmix(uint64_t* x) {
  uint64_t z = (*x += 0x9E3779B97F4A7C15ULL);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
static void seed_from_u64(xoshiro* g, uint64_t seed) { for (int i = 0; i < 4; ++i) g->s[i] = splitmix(&seed); }
static uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
static uint64_t next_u64(void* u) {
  xoshiro* g = (xoshiro*)u;
  const uint64_t r = g->s[0] + g->s[3], t = g->s[1] << 17;
  g->s[2] ^= g->s[0]; g->s[3] ^= g->s[1]; g->s[1] ^= g->s[2]; g->s[0] ^= g->s[3];
  g->s[2] ^= t; g->s[3] = rotl(g->s[3], 45);
  return r;
}
static uint32_t next_u32(void* u) { return (uint32_t)(next_u64(u) >> 32); }

#define CHECK(call) do { lkm_status st__ = (call); if (st__ != LKM_OK) { fprintf(stderr, "%s -> %d: %s\n", #call, (int)st__, ctx ? lkm_last_error(ctx) : ""); return 1; } } while (0)

int main(int argc, char** argv) {
  const uint64_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 3000, d = argc > 2 ? strtoull(argv[2], 0, 10) : 4, k = argc > 3 ? strtoull(argv[3], 0, 10) : 5;
  const int device = argc > 4 ? atoi(argv[4]) : 0;
  lkm_ctx* ctx = NULL;
  CHECK(lkm_create(&ctx, device));
  printf("version %s\n", lkm_version());
  /* blobs the caller owns: k centres on a line, unit-ish noise from the same xoshiro */
  double* x = (double*)malloc(n * d * sizeof(double));
  xoshiro data_rng;
  seed_from_u64(&data_rng, 7);
  for (uint64_t i = 0; i < n; ++i)
    for (uint64_t j = 0; j < d; ++j) {
      double u = 0;
      for (int q = 0; q < 12; ++q) u += (double)(next_u64(&data_rng) >> 11) * (1.0 / 9007199254740992.0);
      x[i * d + j] = 10.0 * (double)(i % k) + (u - 6.0);
    }
  CHECK(lkm_set_data_f64(ctx, x, n, d, (int64_t)d));
  lkm_params p;
  memset(&p, 0, sizeof(p));
  p.n_clusters = k; p.n_runs = 3; p.max_n_iterations = 300; p.tolerance = 1e-4; p.init_method = LKM_INIT_KMEANS_PLUSPLUS;
  p.algorithm = LKM_ALGO_LLOYD; p.pick_mode = LKM_PICK_EXACT;
  double* cen = (double*)calloc(k * d, sizeof(double));
  double* cnt = (double*)calloc(k, sizeof(double));
  double inertia = 0;
  lkm_stats st;
  xoshiro g;
  seed_from_u64(&g, 42);
  lkm_rng rng = {&g, next_u32, next_u64};   /* the binding clones its rng and passes the clone (KM/algorithm.rs:298) */
  CHECK(lkm_fit_f64(ctx, &p, rng, cen, cnt, &inertia, &st));
  printf("fit inertia %.17g n_iter %llu\n", inertia, (unsigned long long)st.n_iter);
  for (uint64_t c = 0; c < k; ++c) { printf("centroid %llu count %.17g:", (unsigned long long)c, cnt[c]); for (uint64_t j = 0; j < d; ++j) printf(" %.17g", cen[c * d + j]); printf("\n"); }
  /* Predict / Transformer on a host matrix the caller owns */
  uint64_t* labels = (uint64_t*)malloc(n * sizeof(uint64_t));
  double* dist = (double*)malloc(n * sizeof(double));
  CHECK(lkm_predict_f64(ctx, cen, k, x, n, d, (int64_t)d, labels));
  CHECK(lkm_transform_f64(ctx, cen, k, x, n, d, (int64_t)d, dist));
  uint64_t lsum = 0; double dsum = 0;
  for (uint64_t i = 0; i < n; ++i) { lsum += labels[i] * (i % 97 + 1); dsum += dist[i]; }
  printf("predict checksum %llu transform sum %.17g\n", (unsigned long long)lsum, dsum);
  /* Hamerly: same optimum, post-update inertia */
  p.algorithm = LKM_ALGO_HAMERLY;
  seed_from_u64(&g, 42);
  double hin = 0;
  double* hcen = (double*)calloc(k * d, sizeof(double));
  CHECK(lkm_fit_f64(ctx, &p, rng, hcen, cnt, &hin, &st));
  printf("hamerly inertia %.17g\n", hin);
  /* FitWith: one mini-batch step from the fitted model; Hamerly is rejected there */
  double inc = 0;
  lkm_status fw = lkm_fit_with_f64(ctx, &p, rng, 1, k, d, hcen, cnt, &inc);
  printf("fit_with hamerly status %d\n", (int)fw);
  p.algorithm = LKM_ALGO_LLOYD;
  fw = lkm_fit_with_f64(ctx, &p, rng, 1, k, d, hcen, cnt, &inc);
  printf("fit_with status %d inertia %.17g\n", (int)fw, inc);
  /* parameter errors come back as the reference's variants */
  p.n_clusters = 0;
  printf("n_clusters 0 -> %d\n", (int)lkm_fit_f64(ctx, &p, rng, cen, cnt, &inertia, &st));
  p.n_clusters = k;
  /* a device group on the same device twice: sharded fit, same optimum */
  {
    lkm_multi* m = NULL;
    int ids[2] = {device, device};
    lkm_status ms = lkm_create_multi(&m, ids, 2);
    if (ms != LKM_OK) { fprintf(stderr, "lkm_create_multi -> %d\n", (int)ms); return 1; }
    ms = lkm_multi_set_data_f64(m, x, n, d, (int64_t)d);
    seed_from_u64(&g, 42);
    double min = 0;
    if (ms == LKM_OK) ms = lkm_multi_fit_f64(m, &p, rng, hcen, cnt, &min, &st);
    if (ms != LKM_OK) { fprintf(stderr, "multi -> %d: %s\n", (int)ms, lkm_multi_last_error(m)); return 1; }
    printf("multi inertia %.17g\n", min);
    lkm_destroy_multi(m);
  }
  CHECK(lkm_destroy(ctx));
  free(x); free(cen); free(cnt); free(labels); free(dist); free(hcen);
  return 0;
}

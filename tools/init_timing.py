"""Seeding wall time on a benchmark shape (GPU box): k-means++ (fast pick) and k-means|| with / without the tensor-core candidate histogram."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_b200 as lb
import bench

wl = os.environ.get("WL", "cfg3")
rows, d, k, dt, _ = bench.WORKLOADS[wl]
n = int(os.environ.get("ROWS", rows))
ctx = lb.Context.default()
centres = bench.centres_for(k, d, dt)
ctx.gen_blobs(n, d, centres, 1.0, seed=2024)
X = None
for name, method in (("k-means++ fast", lb.KMeansInit.KMeansPlusPlus), ("k-means||", lb.KMeansInit.KMeansPara)):
    for rep in range(2):
        p = lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(40 + rep)).n_runs(1).init_method(method).pick_mode(lb.PickMode.Fast).check()
        t0 = time.perf_counter()
        c = p.init_centroids_resident(ctx)
        ms = 1e3 * (time.perf_counter() - t0)
        cost = float(ctx.assign(c, None, want_labels=False)[1].astype(np.float64).sum()) / n
        print(f"{wl} n={n} k={k} {name:15s} rep {rep}: {ms:9.1f} ms, mean cost {cost:10.3f} (LKM_PARA_TC={os.environ.get('LKM_PARA_TC','0')})", flush=True)

"""k-means|| seeding + a few Lloyd iterations on the cfg4 per-GPU shard (12.5M x 64 f32, k = 1024) and k-means++ (fast pick)
on cfg2: wall time and quality against the blob structure (GPU box)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_b200 as lb

def run(name, n, d, k, dtype, init, pick=None, iters=5):
    ctx = lb.Context(0)
    g = np.random.default_rng(1234)
    centres = g.uniform(-30, 30, (k, d)).astype(dtype)
    ctx.gen_blobs(n, d, centres, 1.0, seed=2024)
    p = lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(40)).n_runs(1).max_n_iterations(iters).tolerance(1e-9).init_method(init)
    if pick is not None:
        p = p.pick_mode(pick)
    v = p.check()
    t1 = time.perf_counter()
    model = v.fit_resident(ctx)
    t2 = time.perf_counter()
    st = ctx.last_stats
    print(f"{name}: fit (init + {st.n_iter} Lloyd iterations) {1e3 * (t2 - t1):9.1f} ms wall, device {st.device_ms:9.1f} ms, "
          f"inertia / (n d) = {float(model.inertia()) / d:7.4f}, empty clusters {int((model.cluster_count() == 0).sum())}")
    ctx.close()

run("cfg2 k-means++ fast", 1_000_000, 32, 64, np.float32, lb.KMeansInit.KMeansPlusPlus, lb.PickMode.Fast)
run("cfg2 k-means||     ", 1_000_000, 32, 64, np.float32, lb.KMeansInit.KMeansPara)
run("cfg4 shard k-means||", 12_500_000, 64, 1024, np.float32, lb.KMeansInit.KMeansPara, iters=3)

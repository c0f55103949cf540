"""cfg5 row shape (f64, d = 8, k = 16) with blob-ordered vs row-shuffled observations, new and old kernel (GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_b200 as lb
import bench

n, d, k = 16_000_000, 8, 16
ctx = lb.Context.default()
centres = bench.centres_for(k, d, np.float64)
c0 = bench.start_centroids(centres)
ctx.gen_blobs(n, d, centres, 1.0, seed=2024)
X = ctx.get_data(0, n)
g = np.random.default_rng(0)
for name, data in (("blob order", X), ("shuffled", X[g.permutation(n)].copy())):
    ctx.set_data(data)
    for path in (3, 1):
        ctx.set_assign_path(path)
        ctx.set_l2_flush(512 * 1024 * 1024)
        ctx.set_kernel_timing(True)
        ctx.lloyd_iters(c0, 2, 0.0)
        _, _, _, st = ctx.lloyd_iters(c0, 5, 0.0)
        ctx.set_kernel_timing(False)
        ctx.set_l2_flush(0)
        ms = st.main_kernel_ms / 5
        print(f"{name:10s} path {path}: fused kernel {1e3 * ms:8.1f} us / iteration = {n * d * 8 / ms / 1e6:7.1f} GB/s, re-scanned rows {st.fallback_rows}")

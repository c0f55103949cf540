"""cfg2 shape with blob-ordered vs row-shuffled observations: fused-kernel time per Lloyd iteration (GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_b200 as lb
import bench

n, d, k = 1_000_000, 32, 64
ctx = lb.Context.default()
centres = bench.centres_for(k, d, np.float32)
c0 = bench.start_centroids(centres)
ctx.gen_blobs(n, d, centres, 1.0, seed=2024)
X = ctx.get_data(0, n)
g = np.random.default_rng(0)
for name, data in (("blob order", X), ("shuffled", X[g.permutation(n)].copy())):
    ctx.set_data(data)
    ctx.set_l2_flush(512 * 1024 * 1024)
    ctx.set_kernel_timing(True)
    ctx.lloyd_iters(c0, 3, 0.0)
    _, _, _, st = ctx.lloyd_iters(c0, 20, 0.0)
    ctx.set_kernel_timing(False)
    ctx.set_l2_flush(0)
    print(f"{name:10s}: fused kernel {1e3 * st.main_kernel_ms / 20:7.2f} us / iteration, step {1e3 * st.device_ms / 20:7.2f} us, "
          f"re-scanned rows {st.fallback_rows}")

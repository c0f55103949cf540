"""cfg2 shape, a few Lloyd iterations (for ncu captures of the pipeline kernel)."""
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
ctx.lloyd_iters(c0, 2, 0.0)
_, _, _, st = ctx.lloyd_iters(c0, 4, 0.0)
print(f"step {1e3 * st.device_ms / 4:9.1f} us, re-scanned rows {st.fallback_rows}, path {st.assign_path}")

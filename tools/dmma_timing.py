"""f64 Lloyd iteration with a large contraction: FP64 tensor pipe (DMMA, path 4) against the CUDA-core kernels (GPU box)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_b200 as lb
import bench

for n, d, k in ((2_000_000, 64, 256), (1_000_000, 128, 256), (4_000_000, 16, 128)):
    ctx = lb.Context(0)
    centres = bench.centres_for(k, d, np.float64)
    c0 = bench.start_centroids(centres)
    ctx.gen_blobs(n, d, centres, 1.0, seed=2024)
    for name, env in (("DMMA", "1"), ("CUDA cores", "0")):
        os.environ["LKM_DMMA"] = env
        c2 = lb.Context(0)
        c2.set_data_device_from(ctx) if hasattr(c2, "set_data_device_from") else None
        c2.gen_blobs(n, d, centres, 1.0, seed=2024)
        c2.lloyd_iters(c0, 2, 0.0)
        _, _, _, st = c2.lloyd_iters(c0, 5, 0.0)
        flops = 2.0 * n * k * d * 5 / (st.device_ms * 1e-3) / 1e12
        print(f"n={n} d={d} k={k} {name:10s}: {st.device_ms / 5:8.3f} ms / iteration, path {st.assign_path}, {flops:6.2f} TFLOP/s (2nkd), "
              f"re-scanned {st.fallback_rows & ((1 << 40) - 1)}", flush=True)
        c2.close()
    ctx.close()

#!/bin/bash
python -m pytest tests/test_gpu_parity.py tests/test_golden.py tests/test_gpu_rowwise.py -m gpu -q -x 2>&1 | tail -3
python - <<'PY'
import time, os, sys
import numpy as np
sys.path.insert(0, '.')
import linfa_b200 as lb, bench
for env in ("1", "0"):
    os.environ["LKM_ROWS_ASSIGN"] = env
    ctx = lb.Context(0)
    n, d, k = 4_000_000, 128, 256
    centres = bench.centres_for(k, d, np.float32)
    ctx.gen_blobs(n, d, centres, 1.0, seed=1)
    c = bench.start_centroids(centres)
    ctx.assign(c, None)
    t0 = time.perf_counter(); ctx.assign(c, None); dt = time.perf_counter() - t0
    print(f"LKM_ROWS_ASSIGN={env}: assign of {n} x {d}, k={k}: {1e3*dt:.1f} ms wall", flush=True)
    ctx.close()
PY

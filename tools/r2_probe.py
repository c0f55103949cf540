"""One workload, a few Lloyd iterations, for per-kernel launch lists under ncu (GPU box).
env: WL=cfg3 ROWS=2000000 HASHED=0/1 KPP=0/1 HAMERLY=0/1 ITERS=4 PASSES=0/1/3 (LKM_TC5_FUSE etc. are read by the library)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_b200 as lb
import bench

wl = os.environ.get("WL", "cfg3")
rows, d, k, dt, _ = bench.WORKLOADS[wl]
n = int(os.environ.get("ROWS", rows))
k = int(os.environ.get("K", k))
iters = int(os.environ.get("ITERS", 4))
ctx = lb.Context.default()
centres = bench.centres_for(k, d, dt)
ctx.set_blob_layout(bool(int(os.environ.get("HASHED", "0"))))
ctx.gen_blobs(n, d, centres, 1.0, seed=2024)
c0 = bench.start_centroids(centres)
if int(os.environ.get("KPP", "0")):
    p = lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(40)).n_runs(1).pick_mode(lb.PickMode.Fast).check()
    c0 = p.init_centroids_resident(ctx)
if int(os.environ.get("PASSES", "0")):
    ctx.set_scoring_passes(int(os.environ["PASSES"]))
if int(os.environ.get("PATH_FORCE", "-1")) >= 0:
    ctx.set_assign_path(int(os.environ["PATH_FORCE"]))
run = ctx.hamerly_iters if int(os.environ.get("HAMERLY", "0")) else ctx.lloyd_iters
run(c0, 2, 0.0)
_, _, _, st = run(c0, iters, 0.0)
print(f"{wl} n={n} k={k} hashed={os.environ.get('HASHED','0')} kpp={os.environ.get('KPP','0')}: step {1e3 * st.device_ms / iters:9.1f} us, "
      f"re-scanned rows {st.fallback_rows & ((1 << 40) - 1)}, path {st.assign_path}, launches {st.kernel_launches}, "
      f"hamerly rows scanned {st.rows_scanned} tightened {st.rows_tightened}")

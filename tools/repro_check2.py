"""Does a fit depend on what the context did before?    python tools/repro_check2.py"""
import os, sys, subprocess
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import linfa_b200 as lb
    n, d, k = 3000, 4, 5
    g = lb.Xoshiro256Plus.seed_from_u64(7)
    x = np.empty((n, d))
    for i in range(n):
        for j in range(d):
            u = 0.0
            for _ in range(12):   # not sum(): Python >= 3.12 compensates float sums, the C harness adds in order
                u += (g.next_u64() >> 11) * (1.0 / 9007199254740992.0)
            x[i, j] = 10.0 * (i % k) + (u - 6.0)
    for runs in [int(v) for v in sys.argv[1].split(",")]:
        m = lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(42)).n_runs(runs).fit(lb.DatasetBase(x))
        st = lb.Context.default().last_stats
        print(runs, m.inertia().hex(), [v.hex() for v in m.centroids().ravel()[:4]], st.n_iter, st.assign_path)
else:
    for arg in ("3", "1,3", "3,3", "1,1,3"):
        print(arg, "->")
        print(subprocess.run([sys.executable, __file__, arg], capture_output=True, text=True, env=dict(os.environ, LKM_STEP_TRACE="0")).stdout)

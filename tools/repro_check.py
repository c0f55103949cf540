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
    for runs in (1, 3):
        m = lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(42)).n_runs(runs).fit(lb.DatasetBase(x))
        print(runs, m.inertia().hex(), [v.hex() for v in m.centroids().ravel()[:6]], lb.Context.default().last_stats.n_iter)
else:
    from tests.test_harness import build
    exe = build()
    a = subprocess.run([exe, "3000", "4", "5"], capture_output=True, text=True).stdout
    b = subprocess.run([exe, "3000", "4", "5"], capture_output=True, text=True).stdout
    print("harness twice equal:", a == b)
    print(a[:600])
    p1 = subprocess.run([sys.executable, __file__, "x"], capture_output=True, text=True).stdout
    p2 = subprocess.run([sys.executable, __file__, "x"], capture_output=True, text=True).stdout
    print("python twice equal:", p1 == p2)
    print(p1)
    for v in a.splitlines()[2:4]:
        print([float(t).hex() for t in v.split(":")[1].split()])

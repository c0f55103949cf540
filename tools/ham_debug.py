import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import linfa_b200 as lb
from tests import oracle_lib
from tests.test_hamerly import _blobs
from tests.oracle_lib import INIT_PRECOMPUTED
orc = oracle_lib.load()
n, d, k, dt = 40000, 32, 64, np.float32
X, _ = _blobs(n, d, k, dt, seed=n + d)
g = np.random.default_rng(1)
init = X[g.choice(n, k, replace=False)].copy()
ctx = lb.Context(0)
ctx.set_data(X)
for mi in (7, 8, 10, 12, 16, 20, 30, 34, 36, 60):
    ref = orc.fit_hamerly(X, k, orc.rng(1), n_runs=1, max_iter=mi, tol=1e-4, method=INIT_PRECOMPUTED, precomputed=init, acc_f64=True)
    cen, counts, inertia, st = ctx.hamerly_iters(init, mi, 1e-4)
    lab = ctx.get_labels(n)
    cl, _, _, stl = ctx.lloyd_iters(init, mi, 1e-4)
    print(mi, "gpu n_iter", st.n_iter, "shift", st.last_shift, "ref n_iter", ref["n_iter"], "lloyd n_iter", stl.n_iter, "lloyd shift", stl.last_shift,
          "label diffs", int((lab != ref["labels"]).sum()), "max cen diff", float(np.abs(cen - ref["centroids"]).max()),
          "vs lloyd", float(np.abs(cen - cl).max()), "nan", int(np.isnan(cen).sum()), "scanned", st.rows_scanned, "tight", st.rows_tightened, flush=True)

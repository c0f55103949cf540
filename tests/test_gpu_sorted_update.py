"""The sorted update (linfa_b200/csrc/lkm_sortupd.cuh): compute_centroids (KM/algorithm.rs:785-810) for rows in no particular
order — a stable counting sort of the labels, then a segmented sum over a gather of whole rows.  Same bar as the streaming update:
labels of every row identical to `closest_centroid`, counts exact, centroids / inertia within 1e-5 of the f64-accumulated oracle,
bit-identical repeats; the choice between the two updates is made from the data (label changes between neighbouring rows)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lb():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import linfa_b200

    return linfa_b200


@pytest.fixture(scope="module")
def oracle():
    from tests import oracle_lib

    return oracle_lib.load()


def shuffled_blobs(n, d, kt, seed, spread=30.0):
    g = np.random.default_rng(seed)
    centres = g.uniform(-spread, spread, (kt, d))
    X = (centres[g.integers(0, kt, n)] + g.standard_normal((n, d))).astype(np.float32)
    return X, g


def context(lb, mode):
    """LKM_SORT_UPD is read when the context is created."""
    old = os.environ.get("LKM_SORT_UPD")
    os.environ["LKM_SORT_UPD"] = str(mode)
    try:
        return lb.Context(0)
    finally:
        if old is None:
            del os.environ["LKM_SORT_UPD"]
        else:
            os.environ["LKM_SORT_UPD"] = old


def compare_iterations(ctx, oracle, X, init, iters, want_update_path):
    k = init.shape[0]
    cur = init
    for _ in range(iters):
        nxt, counts, inertia, st = ctx.lloyd_iters(cur, 1, 0.0)
        assert st.update_path == want_update_path, (st.update_path, want_update_path)
        lab, dist = oracle.assign(X, cur)
        got = ctx.get_labels(X.shape[0])
        assert (got == lab).all(), f"{(got != lab).sum()} rows labelled differently from closest_centroid"
        ref = oracle.compute_centroids(cur, X, lab, acc_f64=True)
        np.testing.assert_allclose(nxt, ref, rtol=1e-5, atol=1e-5 * np.abs(ref).max())
        assert counts.tolist() == np.bincount(lab.astype(np.int64), minlength=k).tolist()
        np.testing.assert_allclose(inertia, dist.astype(np.float64).sum(), rtol=1e-5)
        cur = nxt
    return cur


@pytest.mark.parametrize("n,d,k,kt", [
    (20000, 64, 40, 40),        # one centroid chunk
    (30001, 64, 300, 150),      # ragged tail, two centroid chunks
    (25000, 128, 256, 256),     # cfg3's shape
    (6000, 128, 1024, 64),      # more clusters than rows per warp: many entries per warp, empty clusters
    (4097, 64, 2048, 16),       # the largest k the sort takes; nearly every cluster empty or tiny
    (70000, 128, 16, 3),        # few large clusters: many warps per cluster
])
def test_sorted_update_matches_oracle(lb, oracle, n, d, k, kt):
    X, g = shuffled_blobs(n, d, kt, seed=n + d + k)
    init = X[g.choice(n, k, replace=False)].copy()
    ctx = context(lb, 1)
    try:
        ctx.set_data(X)
        ctx.set_keep_labels(True)
        compare_iterations(ctx, oracle, X, init, 3, want_update_path=2)
    finally:
        ctx.close()


def test_sorted_update_handles_ordered_rows_too(lb, oracle):
    """Forced on blob-ordered rows (one label change per blob): every warp sees one or two clusters."""
    n, d, k = 40000, 64, 32
    X, g = shuffled_blobs(n, d, k, seed=5)
    init = X[g.choice(n, k, replace=False)].copy()
    lab, _ = oracle.assign(X, init)
    X = np.ascontiguousarray(X[np.argsort(lab, kind="stable")])
    ctx = context(lb, 1)
    try:
        ctx.set_data(X)
        ctx.set_keep_labels(True)
        compare_iterations(ctx, oracle, X, init, 2, want_update_path=2)
    finally:
        ctx.close()


def test_update_is_chosen_from_the_data(lb, oracle):
    """Automatic mode, a matrix past the 32 MB threshold: hashed rows take the sorted update from the first iteration of the call
    on, blob-ordered rows stay on the streaming update; both agree with the oracle and with each other's cluster sizes."""
    n, d, k = 150000, 64, 64
    X, g = shuffled_blobs(n, d, k, seed=11)
    init = X[g.choice(n, k, replace=False)].copy()
    ctx = context(lb, -1)
    try:
        ctx.set_data(X)
        ctx.set_keep_labels(True)
        c3, counts, inertia, st = ctx.lloyd_iters(init, 3, 0.0)
        assert st.update_path == 2
        # the same three iterations one at a time (each call probes afresh, between its first assignment and update)
        cur = init
        for _ in range(3):
            cur = compare_iterations(ctx, oracle, X, cur, 1, want_update_path=2)
        assert (c3 == cur).all()
        # repeat: same bits
        again = ctx.lloyd_iters(init, 3, 0.0)
        assert (again[0] == c3).all() and again[2] == inertia
        order = np.argsort(oracle.assign(X, init)[0], kind="stable")
        Xo = np.ascontiguousarray(X[order])
        ctx.set_data(Xo)
        c3o, counts_o, _, st_o = ctx.lloyd_iters(init, 3, 0.0)
        assert st_o.update_path in (0, 1)
        assert counts_o.tolist() == counts.tolist()
        np.testing.assert_allclose(c3o, c3, rtol=1e-5, atol=1e-5 * np.abs(c3).max())
    finally:
        ctx.close()


def test_sorted_update_through_fit(lb, oracle):
    """Fit::fit end to end on hashed rows with the sorted update forced: same optimum as the oracle's fit."""
    n, d, k = 50000, 64, 16
    X, g = shuffled_blobs(n, d, k, seed=3)
    ctx = context(lb, 1)
    try:
        model = lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(9)).n_runs(2).max_n_iterations(50).fit(lb.DatasetBase(X), ctx)
        from tests.oracle_lib import INIT_KMEANSPP

        ref = oracle.fit(X, k, oracle.rng(9), n_runs=2, max_iter=50, tol=1e-4, method=INIT_KMEANSPP, acc_f64=True)
        np.testing.assert_allclose(model.inertia(), ref["inertia"], rtol=1e-5)
        np.testing.assert_allclose(model.centroids(), ref["centroids"], rtol=1e-4, atol=1e-4 * np.abs(ref["centroids"]).max())
    finally:
        ctx.close()

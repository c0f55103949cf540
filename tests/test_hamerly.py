"""Hamerly's algorithm (KM/algorithm.rs:248-291, 396-603, 889-919): the oracle's restatement against the reference's own
test literals (CPU), and the CUDA path against the oracle (GPU, through the C ABI)."""
import numpy as np
import pytest

from tests.oracle_lib import INIT_KMEANSPARA, INIT_KMEANSPP, INIT_PRECOMPUTED, INIT_RANDOM
from tests.test_oracle_kats import make_1d_dataset

RTOL = {np.float32: 1e-5, np.float64: 1e-10}


# ---- reference literals, oracle side (no GPU) ----
def test_two_closest_centroids_l2(oracle):   # algorithm.rs:1346-1354
    idx, a, b = oracle.two_closest(np.array([[0.0, 0.0], [10.0, 0.0], [0.0, 10.0]]), np.array([1.0, 1.0]))
    assert idx == 0 and abs(a - np.sqrt(2.0)) < 1e-10 and abs(b - np.sqrt(82.0)) < 1e-10


def test_two_closest_centroids_single(oracle):   # algorithm.rs:1366-1374
    assert oracle.two_closest(np.array([[5.0, 5.0]]), np.array([1.0, 1.0])) == (0, 0.0, 0.0)


def test_two_closest_centroids_obs_is_centroid(oracle):   # algorithm.rs:1376-1384
    idx, a, b = oracle.two_closest(np.array([[0.0, 0.0], [3.0, 4.0], [10.0, 0.0]]), np.array([3.0, 4.0]))
    assert idx == 1 and abs(a) < 1e-10 and abs(b - 5.0) < 1e-10


def test_two_closest_centroids_equidistant(oracle):   # algorithm.rs:1386-1395: index 1 on a tie of the first two
    idx, a, b = oracle.two_closest(np.array([[2.0, 0.0], [0.0, 2.0]]), np.array([1.0, 1.0]))
    assert idx == 1 and abs(a - np.sqrt(2.0)) < 1e-10 and abs(b - np.sqrt(2.0)) < 1e-10


def test_two_farthest_indices(oracle):   # algorithm.rs:1397-1419
    f = oracle.two_farthest
    assert f(np.array([1.0, 5.0, 3.0, 2.0])) == (1, 2)
    assert f(np.array([3.0, 3.0, 3.0])) == (2, 1)
    assert f(np.array([2.0, 7.0])) == (1, 0)
    assert f(np.array([7.0, 2.0])) == (0, 1)
    assert f(np.array([8.0, 1.0, 2.0, 9.0])) == (3, 0)
    assert f(np.array([9.0, 1.0, 2.0, 8.0])) == (0, 3)
    assert f(np.array([1.0])) == (0, 0)


def test_compute_inertia(oracle):   # algorithm.rs:1512-1520
    assert oracle.compute_inertia(np.array([[0.0, 0.0], [3.0, 4.0]]), [0, 0], np.array([[1.0, 1.0]])) == 15.0


def test_hamerly_strategy_new_and_recompute(oracle):
    # algorithm.rs:1485-1497 (memberships [0, 0, 1], sums [[1, 0], [10, 10]]) followed by one recompute_centroids:
    # m_k-means (sum + old) / (count + 1) = [[1/3, 0], [10, 10]]
    X = np.array([[0.0, 0.0], [1.0, 0.0], [10.0, 10.0]])
    c0 = np.array([[0.0, 0.0], [10.0, 10.0]])
    r = oracle.fit_hamerly(X, 2, oracle.rng(1), n_runs=1, max_iter=1, tol=1e-4, method=INIT_PRECOMPUTED, precomputed=c0)
    assert r["status"] == 0 and r["labels"].tolist() == [0, 0, 1] and r["cluster_count"].tolist() == [2.0, 1.0]
    np.testing.assert_allclose(r["centroids"], [[1.0 / 3.0, 0.0], [10.0, 10.0]], atol=1e-12)


@pytest.mark.parametrize("method", [INIT_RANDOM, INIT_KMEANSPP, INIT_KMEANSPARA])
def test_hamerly_lloyd_equivalence(oracle, method):   # algorithm.rs:1266-1345
    data, rng = make_1d_dataset(oracle)
    pre = None
    if method == INIT_KMEANSPARA:   # :1309-1324: k-means|| is not deterministic, both start from the same precomputed seeds
        pre = oracle.init(INIT_KMEANSPARA, 6, data, oracle.rng(7))
        method = INIT_PRECOMPUTED
    # both fits clone the rng (`self.rng().clone()`, KM/algorithm.rs:252,298): the oracle entry points leave it untouched
    lloyd = oracle.fit(data, 6, rng, n_runs=3, method=method, precomputed=pre)
    ham = oracle.fit_hamerly(data, 6, rng, n_runs=3, method=method, precomputed=pre)
    assert lloyd["status"] == 0 and ham["status"] == 0
    assert abs(lloyd["inertia"] - ham["inertia"]) < 1e-4
    sort = lambda c: c[np.lexsort(c.T[::-1])]
    np.testing.assert_allclose(sort(lloyd["centroids"]), sort(ham["centroids"]), atol=1e-4)


def test_hamerly_skips_rows(oracle):
    """the bounds do their job: on separated blobs nearly no row is scanned after the first iteration"""
    g = np.random.default_rng(5)
    centres = g.uniform(-30, 30, (8, 4))
    X = np.concatenate([c + g.standard_normal((500, 4)) for c in centres])
    r = oracle.fit_hamerly(X, 8, oracle.rng(3), n_runs=1, max_iter=50, method=INIT_KMEANSPP)
    assert r["status"] == 0 and r["scans"] < 0.5 * X.shape[0] * r["n_iter"]


# ---- CUDA path against the oracle ----
def _blobs(n, d, k, dt, seed, sigma=1.0, spread=30.0):
    g = np.random.default_rng(seed)
    centres = g.uniform(-spread, spread, (k, d))
    lab = g.integers(0, k, n)
    return (centres[lab] + sigma * g.standard_normal((n, d))).astype(dt), centres.astype(dt)


@pytest.mark.gpu
@pytest.mark.parametrize("n,d,k,dt,sigma", [
    (6000, 2, 3, np.float64, 1.0), (20000, 8, 16, np.float64, 3.0), (40000, 32, 64, np.float32, 1.0),
    (30000, 128, 40, np.float32, 6.0), (25000, 64, 300, np.float32, 4.0), (9000, 5, 7, np.float32, 4.0),
    (9000, 20, 9, np.float64, 6.0), (5000, 3, 1, np.float64, 1.0)])
def test_hamerly_iters_match_oracle(oracle, n, d, k, dt, sigma):
    """lkm_hamerly_iters against the oracle's fit_hamerly from the same (poor) start: labels of the final assignment
    row by row, centroids, post-update inertia, counts, iteration count; and the bounds let rows be skipped."""
    import linfa_b200 as lb

    X, centres = _blobs(n, d, k, dt, seed=n + d)
    g = np.random.default_rng(1)
    init = X[g.choice(n, k, replace=False)].copy()   # random rows: several seeds per blob, many rows move at first
    ctx = lb.Context(0)
    try:
        ctx.set_data(X)
        for max_iter in (1, 2, 7, 60):
            ref = oracle.fit_hamerly(X, k, oracle.rng(1), n_runs=1, max_iter=max_iter, tol=1e-4, method=INIT_PRECOMPUTED,
                                     precomputed=init, acc_f64=True)
            assert ref["status"] == 0
            cen, counts, inertia, st = ctx.hamerly_iters(init, max_iter, 1e-4)
            assert abs(int(st.n_iter) - ref["n_iter"]) <= (1 if max_iter == 60 else 0)
            if int(st.n_iter) != ref["n_iter"]:
                continue   # the shift crossed the tolerance one iteration apart (summation order); nothing to compare row by row
            labels = ctx.get_labels(n)
            assert (labels == ref["labels"]).all(), (max_iter, int((labels != ref["labels"]).sum()))
            tol = RTOL[dt]
            np.testing.assert_allclose(cen, ref["centroids"], rtol=tol, atol=tol * np.abs(ref["centroids"]).max())
            assert counts.tolist() == ref["cluster_count"].tolist()
            np.testing.assert_allclose(float(inertia) / n, float(ref["inertia"]), rtol=max(tol, 2e-6 if dt == np.float32 else 0))
            # same labels as Lloyd's iterations from the same start
            cl, _, _, stl = ctx.lloyd_iters(init, max_iter, 1e-4)
            if int(stl.n_iter) == int(st.n_iter):
                np.testing.assert_allclose(cen, cl, rtol=tol, atol=tol * np.abs(cl).max())
            if max_iter == 60 and k > 1 and int(st.n_iter) > 4:
                print(f"n={n} d={d} k={k}: {st.n_iter} iterations, rows scanned {st.rows_scanned}, tightened {st.rows_tightened}, "
                      f"oracle scans {ref['scans']}")
                assert int(st.rows_scanned) < 0.6 * n * (int(st.n_iter) - 1)
    finally:
        ctx.close()


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64])
def test_fit_hamerly_best_of_runs(oracle, dt):
    """Fit::fit with KMeansAlgorithm::Hamerly: best of n_runs by the post-update inertia, cluster_count from the best
    run's memberships (KM/algorithm.rs:281-291) — against the oracle with the same rng (k-means++, exact pick)."""
    import linfa_b200 as lb

    X, _ = _blobs(4000, 6, 9, dt, seed=3, sigma=5.0)
    ref = oracle.fit_hamerly(X, 9, oracle.rng(21), n_runs=4, max_iter=300, tol=1e-4, method=INIT_KMEANSPP, acc_f64=True)
    model = (lb.KMeans.params_with_rng(9, lb.Xoshiro256Plus.seed_from_u64(21)).n_runs(4)
             .algorithm(lb.KMeansAlgorithm.Hamerly).fit(lb.DatasetBase(X)))
    tol = RTOL[dt]
    np.testing.assert_allclose(model.centroids(), ref["centroids"], rtol=10 * tol, atol=10 * tol * np.abs(ref["centroids"]).max())
    assert model.cluster_count().tolist() == ref["cluster_count"].tolist()
    np.testing.assert_allclose(model.inertia(), ref["inertia"], rtol=10 * tol)
    # Lloyd from the same rng reaches the same optimum; its inertia is the pre-update one, slightly larger
    lloyd = lb.KMeans.params_with_rng(9, lb.Xoshiro256Plus.seed_from_u64(21)).n_runs(4).fit(lb.DatasetBase(X))
    assert abs(lloyd.inertia() - model.inertia()) < 1e-3 * model.inertia()


@pytest.mark.gpu
def test_hamerly_reproducible(oracle):
    """fixed-point running sums: two runs give the same bits"""
    import linfa_b200 as lb

    X, _ = _blobs(50000, 16, 32, np.float32, seed=9, sigma=5.0)
    init = X[:32].copy()
    ctx = lb.Context(0)
    try:
        ctx.set_data(X)
        a = ctx.hamerly_iters(init, 25, 1e-6)
        b = ctx.hamerly_iters(init, 25, 1e-6)
        assert (a[0] == b[0]).all() and a[2] == b[2] and (a[1] == b[1]).all()
    finally:
        ctx.close()

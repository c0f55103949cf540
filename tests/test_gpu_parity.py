"""GPU parity tests: the CUDA path, called through the C ABI, against the CPU oracle.

Bar (BASELINE.json north_star): labels identical to the reference on the same
inputs; min squared distances bit-identical (the kernels evaluate the winning
distance in the reference's direct form); centroids and inertia within 1e-5
(f32) / 1e-10 (f64) relative of the f64-accumulated oracle.
"""
import numpy as np
import pytest

from tests.oracle_lib import INIT_KMEANSPARA, INIT_KMEANSPP, INIT_PRECOMPUTED, INIT_RANDOM

pytestmark = pytest.mark.gpu

RTOL = {np.float32: 1e-5, np.float64: 1e-10}
DTYPES = [np.float32, np.float64]


@pytest.fixture(scope="module")
def lb():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import linfa_b200

    return linfa_b200


@pytest.fixture(scope="module")
def ctx(lb):
    return lb.Context.default()


def blobs(n, d, k_true, dt, seed=0, spread=30.0):
    g = np.random.default_rng(seed)
    centres = g.uniform(-spread, spread, (k_true, d))
    m = n // k_true
    parts = []
    for b in range(k_true):
        rows = m if b < k_true - 1 else n - m * (k_true - 1)
        parts.append(g.standard_normal((rows, d)) + centres[b])
    return np.concatenate(parts).astype(dt), centres.astype(dt)


def function_test_1d(x):
    return np.where(x < 0.4, x * x, np.where((x >= 0.4) & (x < 0.8), 3.0 * x + 1.0, np.sin(10.0 * x)))


def make_1d_dataset(oracle, seed=42):
    rng = oracle.rng(seed)
    xt = rng.uniform_f64(0.0, 1.0, (100, 1))
    return np.concatenate([xt, function_test_1d(xt)], axis=1)


# ---- literal KATs through the C ABI ----
@pytest.mark.parametrize("dt", DTYPES)
def test_kat_min_dists_and_argmin(lb, dt):
    m = lb.KMeans(np.array([[0.0, 1.0], [40.0, 10.0]], dtype=dt), np.zeros(2), 0.0)
    assert m.transform(np.array([[3.0, 4.0], [1.0, 3.0], [25.0, 15.0]], dtype=dt)).tolist() == [18.0, 5.0, 250.0]
    m = lb.KMeans(np.array([[0, 0], [1, 2], [20, 0], [0, 20]], dtype=dt), np.zeros(4), 0.0)
    assert m.predict(np.array([[1, 0.6], [20, 2], [20, 0], [7, 20]], dtype=dt)).tolist() == [0, 2, 2, 3]
    assert m.predict(np.array([20.0, 2.0], dtype=dt)) == 2  # Ix1 form


def test_kat_self_nearest_and_ties(lb, oracle):
    Cn = oracle.rng(42).uniform_f64(-100.0, 100.0, (20, 5))
    m = lb.KMeans(Cn, np.zeros(20), 0.0)
    assert m.predict(Cn).tolist() == list(range(20))
    assert (m.transform(Cn) == 0).all()
    m = lb.KMeans(np.array([[2.0, 0.0], [0.0, 2.0], [2.0, 0.0]]), np.zeros(3), 0.0)
    assert m.predict(np.array([[1.0, 1.0], [2.0, 0.0]])).tolist() == [0, 0]  # strict '<': lowest index


def test_kat_membership_counts(lb, ctx):
    Cn = np.array([[0.0, 1.0], [40.0, 10.0], [3.0, 9.0]])
    X = np.array([[3.0, 4.0], [1.0, 3.0], [25.0, 15.0]])
    labels, _ = ctx.assign(Cn, X)
    assert np.bincount(labels.astype(int), minlength=3).tolist() == [2, 1, 0]


@pytest.mark.parametrize("dt", DTYPES)
def test_kat_tolerance_step(lb, dt):
    X = np.array([[1.0, 1.0], [11.0, 11.0]], dtype=dt)
    p = (lb.KMeans.params_with_rng(1, lb.Xoshiro256Plus.seed_from_u64(45)).tolerance(8.5)
         .init_method(lb.KMeansInit.Precomputed(np.array([[0.0, 0.0]], dtype=dt))))
    m = p.fit(lb.DatasetBase(X))
    assert m.centroids().tolist() == [[4.0, 4.0]]
    assert lb.Context.default().last_stats.n_iter == 1
    assert p.fit_with(None, lb.DatasetBase(X)) is not None  # converges on the first batch


def test_kat_six_points(lb):
    X = np.array([[0, 0], [1, 0], [0, 1], [10, 10], [11, 10], [10, 11]], dtype=np.float64)
    m = (lb.KMeans.params(2).n_runs(1).tolerance(1e-4)
         .init_method(lb.KMeansInit.Precomputed([[0.0, 0.0], [10.0, 10.0]])).fit(lb.DatasetBase(X)))
    assert lb.Context.default().last_stats.n_iter == 8
    np.testing.assert_allclose(m.centroids(), [[0.33332825, 0.33332825], [10.33332825, 10.33332825]], atol=1e-8)
    assert m.predict(X).tolist() == [0, 0, 0, 1, 1, 1]
    assert m.cluster_count().tolist() == [3.0, 3.0]
    assert abs(m.inertia() - 0.4444444453) < 1e-9


def test_kat_empty_cluster_keeps_centroid(lb, ctx):
    # algorithm.rs:1117-1125 through one Lloyd iteration
    ctx.set_data(np.array([[1.0, 2.0]]))
    c, counts, inertia, st = ctx.lloyd_iters(np.ones((2, 2)), 1, 0.0)
    assert c.tolist() == [[1.0, 1.5], [1.0, 1.0]] and counts.tolist() == [1.0, 0.0] and inertia == 1.0


def test_kat_incremental(lb):
    # algorithm.rs:1180-1202
    model = lb.KMeans(np.array([[-1.0, -1.0], [3.0, 4.0], [7.0, 8.0]]), np.zeros(3), 0.0)
    p = lb.KMeans.params_with_rng(3, lb.Xoshiro256Plus.seed_from_u64(45)).tolerance(100.0)
    model = p.fit_with(model, lb.DatasetBase(np.array([[-1.0, -3.0], [0.0, 0.0], [3.0, 5.0], [5.0, 5.0]])))
    np.testing.assert_allclose(model.centroids(), [[-0.5, -1.5], [4, 5], [7, 8]], atol=1e-12)
    model = p.fit_with(model, lb.DatasetBase(np.array([[-5.0, -5.0], [0.0, 0.0], [10.0, 10.0]])))
    np.testing.assert_allclose(model.centroids(), [[-6 / 4, -8 / 4], [4, 5], [10, 10]], atol=1e-12)
    assert model.cluster_count().tolist() == [4.0, 2.0, 1.0]


# ---- assignment parity on random blobs ----
SHAPES = [(1000, 2, 3), (777, 3, 1), (5000, 8, 16), (3001, 32, 64), (2049, 64, 256), (1500, 128, 256),
          (900, 5, 7), (4100, 50, 5), (1024, 16, 1024), (300, 1, 4), (2600, 12, 33)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n,d,k", SHAPES)
def test_assign_matches_oracle_bitwise(lb, ctx, oracle, n, d, k, dt):
    X, _ = blobs(n, d, max(1, min(k, 20)), dt, seed=n + d + k)
    g = np.random.default_rng(k)
    Cn = X[g.choice(n, size=k, replace=k > n)] + (g.standard_normal((k, d)) * 0.5).astype(dt)
    labels, dists = ctx.assign(Cn.astype(dt), X)
    ref_l, ref_d = oracle.assign(X, Cn.astype(dt))
    assert (labels == ref_l).all()
    assert (dists == ref_d).all()  # bit-identical


@pytest.mark.parametrize("dt", DTYPES)
def test_assign_strided_and_resident(lb, ctx, oracle, dt):
    X, _ = blobs(1200, 6, 4, dt, seed=5)
    wide = np.zeros((1200, 10), dtype=dt)
    wide[:, :6] = X
    view = wide[:, :6]  # row stride 10
    Cn = X[:9].copy()
    l1, d1 = ctx.assign(Cn, view)
    ref_l, ref_d = oracle.assign(X, Cn)
    assert (l1 == ref_l).all() and (d1 == ref_d).all()
    ctx.set_data(view)
    l2, d2 = ctx.assign(Cn, None)
    assert (l2 == ref_l).all() and (d2 == ref_d).all()
    np.testing.assert_array_equal(ctx.get_data(), X)


def test_predict_length_mismatch_rejected(lb):
    m = lb.KMeans(np.zeros((2, 2)), np.zeros(2), 0.0)
    with pytest.raises(AssertionError):
        m.predict_inplace(np.zeros((5, 2)), np.zeros(4, dtype=np.uint64))


# ---- Lloyd parity from Precomputed centroids ----
LLOYD_CASES = [(10000, 2, 3, 3), (4000, 3, 10, 10), (30000, 3, 10, 10), (20000, 32, 64, 64), (6000, 8, 16, 16),
               (5000, 16, 12, 30), (3000, 64, 8, 8), (2000, 128, 256, 16), (9000, 4, 40, 13)]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n,d,k,k_true", LLOYD_CASES)
def test_lloyd_matches_oracle(lb, ctx, oracle, n, d, k, k_true, dt):
    X, _ = blobs(n, d, k_true, dt, seed=n + k)
    g = np.random.default_rng(d)
    init = X[g.choice(n, size=k, replace=False)].copy()
    tol = 1e-4
    ref = oracle.fit(X, k, oracle.rng(1), n_runs=1, max_iter=25, tol=tol, method=INIT_PRECOMPUTED, precomputed=init,
                     acc_f64=True)
    faithful = oracle.fit(X, k, oracle.rng(1), n_runs=1, max_iter=25, tol=tol, method=INIT_PRECOMPUTED,
                          precomputed=init)
    m = (lb.KMeans.params(k).n_runs(1).max_n_iterations(25).tolerance(tol)
         .init_method(lb.KMeansInit.Precomputed(init)).fit(lb.DatasetBase(X)))
    st = ctx.last_stats
    assert abs(int(st.n_iter) - ref["n_iter"]) <= 1, (st.n_iter, ref["n_iter"], st.last_shift)
    rt = RTOL[dt]
    if int(st.n_iter) == ref["n_iter"]:
        scale = np.abs(ref["centroids"]).max()
        np.testing.assert_allclose(m.centroids(), ref["centroids"], rtol=rt, atol=rt * scale)
        np.testing.assert_allclose(m.inertia(), ref["inertia"], rtol=rt)
        assert m.cluster_count().tolist() == ref["cluster_count"].tolist()
        # the F-faithful oracle's own deviation from the f64-accumulated one, for the record
        dev = np.abs(faithful["centroids"] - ref["centroids"]).max() / scale
        assert dev < 1e-3
    # labels: identical to the reference's assignment for the same centroids
    labels = m.predict(X)
    ref_l, ref_d = oracle.assign(X, m.centroids())
    assert (labels == ref_l).all()
    assert (m.transform(X) == ref_d).all()


@pytest.mark.parametrize("dt", DTYPES)
def test_lloyd_fixed_iterations_no_convergence(lb, ctx, oracle, dt):
    X, _ = blobs(8000, 8, 5, dt, seed=11)
    init = X[[1, 50, 900, 4000, 7000, 7999]].copy()
    ctx.set_data(X)
    c, counts, inertia, st = ctx.lloyd_iters(init, 4, 0.0)  # tolerance 0: exactly 4 iterations
    assert st.n_iter == 4
    ref = oracle.fit(X, 6, oracle.rng(1), n_runs=1, max_iter=4, tol=float(np.finfo(dt).tiny), method=INIT_PRECOMPUTED,
                     precomputed=init, acc_f64=True)
    assert ref["n_iter"] == 4
    rt = RTOL[dt]
    np.testing.assert_allclose(c, ref["centroids"], rtol=rt, atol=rt * np.abs(ref["centroids"]).max())
    np.testing.assert_allclose(inertia / 8000, ref["inertia"], rtol=rt)
    assert counts.tolist() == ref["cluster_count"].tolist()


def test_lloyd_is_run_to_run_reproducible(lb, ctx):
    X, _ = blobs(50000, 16, 20, np.float32, seed=3)
    init = X[::2500].copy()
    ctx.set_data(X)
    a = ctx.lloyd_iters(init, 10, 0.0)
    b = ctx.lloyd_iters(init, 10, 0.0)
    assert (a[0] == b[0]).all() and a[2] == b[2] and (a[1] == b[1]).all()


# ---- edge cases ----
def test_edge_cases(lb, oracle):
    X = np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0], [7.0, 8.0]])
    m = lb.KMeans.params(1).fit(lb.DatasetBase(X))
    np.testing.assert_allclose(m.centroids(), [[4.0, 5.0]], atol=1e-4)
    Y = np.array([[1.0, 2.0], [10.0, 20.0], [-5.0, -5.0], [100.0, 0.0]])
    m = lb.KMeans.params(4).init_method(lb.KMeansInit.Precomputed(Y)).fit(lb.DatasetBase(Y))
    assert m.inertia() == 0.0
    m = lb.KMeans.params(1).fit(lb.DatasetBase(np.full((4, 2), 5.0)))
    np.testing.assert_allclose(m.centroids(), [[5.0, 5.0]], atol=1e-10)
    assert abs(m.inertia()) < 1e-10
    m = lb.KMeans.params(1).fit(lb.DatasetBase(np.array([[3.0, 7.0]])))
    assert m.centroids().tolist() == [[3.0, 7.0]] and m.inertia() == 0.0


def test_nan_data_is_inertia_error(lb):
    with pytest.raises(lb.KMeansError) as e:
        (lb.KMeans.params(2).n_runs(2).init_method(lb.KMeansInit.Precomputed(np.zeros((2, 2))))
         .fit(lb.DatasetBase(np.full((8, 2), np.nan))))
    assert e.value.kind == "InertiaError"


def test_precomputed_wrong_shape_panics(lb):
    with pytest.raises(AssertionError):
        lb.KMeans.params(3).init_method(lb.KMeansInit.Precomputed(np.zeros((2, 2)))).fit(np.zeros((10, 2)))


def test_native_param_errors(lb, ctx):
    # the C ABI validates too (a Rust binding may skip its own ParamGuard)
    from linfa_b200 import _ffi
    import ctypes as C

    ctx.set_data(np.zeros((4, 2)))
    for field, val, code in [("n_clusters", 0, 10), ("n_runs", 0, 11), ("tolerance", 0.0, 12),
                             ("max_n_iterations", 0, 13)]:
        p = _ffi.lkm_params(1, 1, 5, 1e-4, _ffi.INIT_RANDOM, 0, 0, 0, None, 0, 0)
        setattr(p, field, val)
        out = np.zeros(8)
        st = _ffi.lib.lkm_fit_f64(ctx.h, C.byref(p), _ffi.make_rng(lb.Xoshiro256Plus.seed_from_u64(1)),
                                  _ffi.fptr(out, C.c_double), _ffi.fptr(out, C.c_double), _ffi.fptr(out, C.c_double), None)
        assert st == code


# ---- seeding ----
@pytest.mark.parametrize("dt", DTYPES)
def test_random_init_matches_oracle(lb, oracle, dt):
    X, _ = blobs(5000, 4, 5, dt, seed=8)
    for k, seed in [(3, 42), (11, 7), (40, 9), (200, 3)]:
        got = lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(seed)).init_method(
            lb.KMeansInit.Random).check().init_centroids(X)
        want = oracle.init(INIT_RANDOM, k, X, oracle.rng(seed))
        assert (got == want).all()


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("n,d,k", [(3000, 3, 10), (20000, 8, 24), (100000, 32, 16)])
def test_kmeanspp_exact_matches_oracle(lb, oracle, n, d, k, dt):
    X, _ = blobs(n, d, k, dt, seed=n)
    got = lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(40)).init_method(
        lb.KMeansInit.KMeansPlusPlus).check().init_centroids(X)
    want = oracle.init(INIT_KMEANSPP, k, X, oracle.rng(40))
    assert (got == want).all()  # same row picks, in the same order


def test_kmeanspp_default_fit_matches_oracle(lb, oracle):
    # the reference's bench configuration (benches/k_means.rs:36-77): k-means++, n_runs 10, tol 1e-3
    X, _ = blobs(4000, 3, 10, np.float64, seed=40)
    m = (lb.KMeans.params_with_rng(10, lb.Xoshiro256Plus.seed_from_u64(40)).max_n_iterations(1000).tolerance(1e-3)
         .fit(lb.DatasetBase(X)))
    ref = oracle.fit(X, 10, oracle.rng(40), n_runs=10, max_iter=1000, tol=1e-3, method=INIT_KMEANSPP, acc_f64=True)
    np.testing.assert_allclose(m.centroids(), ref["centroids"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(m.inertia(), ref["inertia"], rtol=1e-10)
    assert m.cluster_count().tolist() == ref["cluster_count"].tolist()


def test_weighted_zero_weight_never_picked_via_para(lb):
    # statistical tests of init.rs:372-449 for ++ (fast mode) and ||
    g = np.random.default_rng(0)
    X = np.concatenate([g.standard_normal((50, 2)) + c for c in (20.0, -1000.0, 1000.0)])
    for init in (lb.KMeansInit.KMeansPlusPlus, lb.KMeansInit.KMeansPara):
        for mode in (lb.PickMode.Exact, lb.PickMode.Fast):
            for seed in range(5):
                out = (lb.KMeans.params_with_rng(3, lb.Xoshiro256Plus.seed_from_u64(seed)).init_method(init)
                       .pick_mode(mode).check().init_centroids(X))
                found = set()
                for row in out:
                    idx = np.where((X == row).all(axis=1))[0]
                    assert idx.size >= 1
                    found.add(int(idx[0]) // 50)
                assert found == {0, 1, 2}
    # degenerate n < k does not fail and duplicates the point (init.rs:374-377)
    for init in (lb.KMeansInit.KMeansPlusPlus, lb.KMeansInit.KMeansPara):
        out = lb.KMeans.params(2).init_method(init).check().init_centroids(np.array([[1.0, 2.0]]))
        assert out.tolist() == [[1.0, 2.0], [1.0, 2.0]]


def test_init_losses_beat_random(lb, ctx, oracle):
    g = np.random.default_rng(1)
    X = np.concatenate([g.standard_normal((50, 2)) + c for c in (20.0, -1000.0, 1000.0)])
    rng = lb.Xoshiro256Plus.seed_from_u64(42)

    def loss(init, mode=lb.PickMode.Exact):
        c = lb.KMeans.params_with_rng(3, rng.clone()).init_method(init).pick_mode(mode).check().init_centroids(X)
        return ctx.assign(c, X)[1].sum()

    l_rand = loss(lb.KMeansInit.Random)
    assert loss(lb.KMeansInit.KMeansPlusPlus) < l_rand
    assert loss(lb.KMeansInit.KMeansPlusPlus, lb.PickMode.Fast) < l_rand
    assert loss(lb.KMeansInit.KMeansPara) < l_rand


def test_fast_pick_cost_close_to_exact(lb, ctx):
    X, _ = blobs(200000, 16, 32, np.float32, seed=2)
    costs = {}
    for mode in (lb.PickMode.Exact, lb.PickMode.Fast):
        c = (lb.KMeans.params_with_rng(32, lb.Xoshiro256Plus.seed_from_u64(5)).pick_mode(mode).check()
             .init_centroids(X))
        costs[mode] = float(ctx.assign(c, X)[1].astype(np.float64).sum())
    assert costs[lb.PickMode.Fast] <= 1.5 * costs[lb.PickMode.Exact]


# ---- self-consistency families (algorithm.rs:1015-1077) ----
@pytest.mark.parametrize("init", ["Random", "KMeansPlusPlus", "KMeansPara"])
def test_n_runs_consistency(lb, oracle, init):
    data = make_1d_dataset(oracle)
    rng = lb.Xoshiro256Plus.seed_from_u64(42)
    for _ in range(100):
        rng.next_u64()
    im = getattr(lb.KMeansInit, init)
    m = lb.KMeans.params_with(3, rng.clone(), lb.L2Dist()).n_runs(1).init_method(im).fit(lb.DatasetBase(data))
    clusters = m.predict(lb.DatasetBase(data))
    inertia = sum(oracle.sq_l2_dist(data[i], m.centroids()[int(clusters.targets[i])]) for i in range(100))
    assert abs(inertia - m.transform(data).sum()) < 1e-5
    assert m.predict(data[0]) == clusters.targets[0]
    m2 = lb.KMeans.params_with(3, rng.clone(), lb.L2Dist()).init_method(im).fit(lb.DatasetBase(data))
    if init == "Random":
        assert m2.transform(data).sum() <= m.transform(data).sum()
    if init != "KMeansPara":  # k-means|| has no deterministic reference
        ref = oracle.fit(data, 3, _skip100(oracle), n_runs=1,
                         method=INIT_RANDOM if init == "Random" else INIT_KMEANSPP, acc_f64=True)
        np.testing.assert_allclose(m.centroids(), ref["centroids"], rtol=1e-9, atol=1e-9)


def _skip100(oracle):
    r = oracle.rng(42)
    for _ in range(100):
        r.next_u64()
    return r


def test_max_n_iterations(lb, oracle):
    data = make_1d_dataset(oracle)
    m = (lb.KMeans.params_with(6, _py_skip100(lb), lb.L2Dist()).n_runs(1).max_n_iterations(5)
         .init_method(lb.KMeansInit.Random).fit(lb.DatasetBase(data)))
    assert lb.Context.default().last_stats.n_iter <= 5 and m.centroids().shape == (6, 2)


def _py_skip100(lb):
    r = lb.Xoshiro256Plus.seed_from_u64(42)
    for _ in range(100):
        r.next_u64()
    return r


# ---- device-side blobs ----
@pytest.mark.parametrize("dt", DTYPES)
def test_gen_blobs_layout_and_determinism(lb, ctx, dt):
    centres = np.array([[0.0, 1.0], [-10.0, 20.0], [-1.0, 10.0]], dtype=dt)
    ctx.gen_blobs(10000, 2, centres, 1.0, seed=7)
    a = ctx.get_data()
    ctx.gen_blobs(10000, 2, centres, 1.0, seed=7)
    assert (ctx.get_data() == a).all()
    m = 10000 // 3
    for b in range(3):
        rows = a[b * m:(b + 1) * m] if b < 2 else a[2 * m:]
        assert np.abs(rows.mean(0) - centres[b]).max() < 0.1
        assert np.abs(rows.std(0) - 1.0).max() < 0.05
    # a shard generated with a row offset equals the same rows of the whole
    ctx.gen_blobs(3000, 2, centres, 1.0, seed=7, row_offset=2500, n_global=10000)
    assert (ctx.get_data() == a[2500:5500]).all()


# ---- tcgen05 3xTF32 path (f32, d = 32): certified argmin + exact re-evaluation ----
TC_CASES = [(20000, 64, 64, 30.0), (5000, 1, 4, 30.0), (7777, 3, 3, 30.0), (33333, 100, 40, 30.0),
            (12800, 128, 128, 30.0), (15000, 64, 8, 2.0), (9000, 37, 5, 0.5), (129, 2, 2, 30.0)]


@pytest.mark.parametrize("n,k,k_true,spread", TC_CASES)
def test_tc_path_matches_oracle(lb, ctx, oracle, n, k, k_true, spread):
    """spread 2.0 / 0.5 make the blobs overlap heavily, so many rows fall inside the certification
    margin and take the exact fallback; labels must still be identical."""
    dt = np.float32
    X, _ = blobs(n, 32, k_true, dt, seed=n + k, spread=spread)
    g = np.random.default_rng(k)
    init = X[g.choice(n, size=k, replace=False)].copy()
    ctx.set_assign_path(2)
    try:
        ctx.set_data(X)
        c, counts, inertia, st = ctx.lloyd_iters(init, 3, 0.0)
        assert st.assign_path == 2 and st.n_iter == 3
    finally:
        ctx.set_assign_path(-1)
    # replay the three iterations with the oracle, checking the assignment of every iteration
    cur = init
    for it in range(3):
        lab, d = oracle.assign(X, cur)
        nxt = oracle.compute_centroids(cur, X, lab, acc_f64=True)
        if it == 2:
            ref_counts = np.bincount(lab.astype(np.int64), minlength=k)
            ref_inertia = d.astype(np.float64).sum()
        cur = nxt
    scale = np.abs(cur).max()
    np.testing.assert_allclose(c, cur, rtol=1e-5, atol=1e-5 * scale)
    assert counts.tolist() == ref_counts.tolist()  # identical labels in the last assignment
    np.testing.assert_allclose(inertia, ref_inertia, rtol=1e-5)
    print(f"fallback rows over 3 iterations: {st.fallback_rows} of {3 * n}")


def test_tc_path_ties_and_duplicates(lb, ctx, oracle):
    # duplicate centroids => exact ties => the lowest index must win, as in the reference
    X, _ = blobs(4096, 32, 4, np.float32, seed=1)
    init = np.concatenate([X[:3], X[:3], X[5:6]]).copy()
    ctx.set_assign_path(2)
    try:
        ctx.set_data(X)
        c, counts, inertia, st = ctx.lloyd_iters(init, 1, 0.0)
    finally:
        ctx.set_assign_path(-1)
    lab, d = oracle.assign(X, init)
    assert counts.tolist() == np.bincount(lab.astype(np.int64), minlength=7).tolist()
    assert counts[3:6].sum() == 0
    np.testing.assert_allclose(inertia, d.astype(np.float64).sum(), rtol=1e-6)


def test_tc_path_nan_rows(lb, ctx):
    X, _ = blobs(1000, 32, 4, np.float32, seed=2)
    X[17, 5] = np.nan
    ctx.set_assign_path(2)
    try:
        ctx.set_data(X)
        c, counts, inertia, st = ctx.lloyd_iters(X[:4].copy(), 1, 0.0)
    finally:
        ctx.set_assign_path(-1)
    assert np.isnan(inertia)  # a NaN row poisons the inertia, as dists.sum() does in the reference
    assert counts.sum() == 1000


def test_tc_path_is_reproducible(lb, ctx):
    X, _ = blobs(100000, 32, 64, np.float32, seed=3)
    init = X[::1563][:64].copy()
    ctx.set_data(X)
    a = ctx.lloyd_iters(init, 5, 0.0)
    b = ctx.lloyd_iters(init, 5, 0.0)
    assert a[3].assign_path == 2
    assert (a[0] == b[0]).all() and a[2] == b[2] and (a[1] == b[1]).all()


# ---- tensor-core assignment for wide rows / many clusters (two-pass: tc_assign_kernel + update pass) ----
TCA_CASES = [(6000, 128, 256, 16, 30.0, False), (5000, 64, 1024, 40, 30.0, False), (3000, 128, 300, 10, 1.0, False),
             (2500, 64, 700, 5, 0.5, False), (1000, 128, 1024, 7, 30.0, False),
             # update_rows_kernel: several sub-tiles per CTA with a ragged tail, blob order (one segment per sub-tile) and
             # shuffled rows (many one-row segments), one / two / four cluster chunks
             (148 * 32 * 3 + 19, 128, 256, 256, 30.0, False), (148 * 32 * 3 + 19, 128, 256, 256, 30.0, True),
             (40000, 64, 1024, 64, 30.0, True), (30001, 128, 1000, 50, 30.0, True), (20000, 64, 512, 300, 30.0, True),
             (2047, 64, 4000, 3, 30.0, False), (3000, 64, 4500, 5, 30.0, False), (2000, 128, 8200, 4, 30.0, True),   # 18 / 33 chunk launches
             # CTA-pair assignment kernel (assign_tc5_kernel: d in {64, 128}, k <= 256): heavy overlap (many exact re-scans),
             # k below one CTA's half, a single partial tile (the second CTA of the pair idles), odd tile counts
             (128 * 7 + 5, 128, 200, 20, 1.0, True), (100, 64, 256, 4, 30.0, False), (128 * 148 * 2 + 1, 64, 130, 130, 30.0, True),
             (5000, 128, 100, 100, 0.3, True), (128 * 75, 128, 256, 64, 30.0, False), (4000, 64, 40, 3, 30.0, True),
             # several launches per iteration (one per 256 centroids): a last chunk of one centroid, of 255; one row; one tile + 1
             (129, 128, 257, 5, 30.0, False), (3000, 64, 511, 20, 5.0, True), (1, 128, 40, 1, 30.0, False), (20000, 128, 600, 600, 30.0, False),
             # fused update of the screening kernel: long runs of uniform tiles, cluster changes inside a CTA's tile sequence,
             # a ragged last tile, more tiles than one pass over the CTA pairs
             (40000, 128, 32, 16, 30.0, False), (128 * 74 * 2 * 3 + 70, 128, 20, 7, 30.0, False), (60000, 128, 256, 3, 30.0, False)]


@pytest.mark.parametrize("passes", [0, 1, 3])
@pytest.mark.parametrize("n,d,k,k_true,spread,shuffle", TCA_CASES)
def test_tc_assign_path_matches_oracle(lb, ctx, oracle, n, d, k, k_true, spread, shuffle, passes):
    """passes: 0 = adaptive, 1 = one-product screening variant of the pair kernel, 3 = 3xTF32 (same results)."""
    if passes and not (d in (64, 128) and k <= 16384):
        pytest.skip("scoring passes only select between the variants of the pair kernel")
    dt = np.float32
    X, _ = blobs(n, d, k_true, dt, seed=n + k, spread=spread)
    if shuffle:
        X = X[np.random.default_rng(5).permutation(n)].copy()
    g = np.random.default_rng(k)
    init = X[g.choice(n, size=k, replace=k > n)].copy()
    if k > n:
        init += (g.standard_normal(init.shape) * 0.01).astype(dt)
    ctx.set_data(X)
    ctx.set_scoring_passes(passes)
    try:
        c, counts, inertia, st = ctx.lloyd_iters(init, 2, 0.0)
    finally:
        ctx.set_scoring_passes(0)
    assert st.assign_path == 2 and st.n_iter == 2
    cur = init
    for it in range(2):
        lab, dd = oracle.assign(X, cur)
        nxt = oracle.compute_centroids(cur, X, lab, acc_f64=True)
        if it == 1:
            ref_counts = np.bincount(lab.astype(np.int64), minlength=k)
            ref_inertia = dd.astype(np.float64).sum()
        cur = nxt
    scale = np.abs(cur).max()
    np.testing.assert_allclose(c, cur, rtol=1e-5, atol=1e-5 * scale)
    assert counts.tolist() == ref_counts.tolist()
    np.testing.assert_allclose(inertia, ref_inertia, rtol=1e-5)
    print(f"fallback rows over 2 iterations: {st.fallback_rows} of {2 * n}")


@pytest.mark.parametrize("passes", [1, 3])
@pytest.mark.parametrize("scale,offset", [(1e-3, 0.0), (1e3, 0.0), (1.0, 5e3), (1e-20, 0.0), (1e15, 0.0)])
def test_pair_kernel_magnitudes(lb, ctx, oracle, scale, offset, passes):
    """assign_tc5_kernel: score window and certificate over 35 orders of magnitude and far from the origin."""
    dt = np.float32
    X, _ = blobs(9000, 128, 40, dt, seed=13, spread=30.0)
    X = (X.astype(np.float64) * scale + offset).astype(dt)
    g = np.random.default_rng(6)
    X = X[g.permutation(len(X))].copy()
    init = X[g.choice(len(X), size=40, replace=False)].copy()
    ctx.set_data(X)
    ctx.set_scoring_passes(passes)
    try:
        c, counts, inertia, st = ctx.lloyd_iters(init, 2, 0.0)
    finally:
        ctx.set_scoring_passes(0)
    assert st.assign_path == 2
    cur = init
    for _ in range(2):
        lab, d = oracle.assign(X, cur)
        cur = oracle.compute_centroids(cur, X, lab, acc_f64=True)
    assert counts.tolist() == np.bincount(lab.astype(np.int64), minlength=40).tolist()
    np.testing.assert_allclose(c, cur, rtol=1e-5, atol=1e-5 * np.abs(cur).max())
    ref_inertia = d.astype(np.float64).sum()
    if np.isfinite(ref_inertia):
        np.testing.assert_allclose(inertia, ref_inertia, rtol=2e-5)


@pytest.mark.parametrize("passes", [1, 3])
def test_pair_kernel_nan_inf_and_duplicates(lb, ctx, oracle, passes):
    dt = np.float32
    ctx.set_scoring_passes(passes)
    try:
        _pair_kernel_nan_inf_and_duplicates(ctx, oracle)
    finally:
        ctx.set_scoring_passes(0)


def _pair_kernel_nan_inf_and_duplicates(ctx, oracle):
    dt = np.float32
    X, _ = blobs(3000, 64, 6, dt, seed=4)
    idx = list(range(0, 3000, 100)) + [0, 600]                  # 32 centroids, the last two duplicate earlier ones:
    init = X[idx].copy()                                        # exact ties, the lowest index wins
    ctx.set_data(X)
    c, counts, inertia, st = ctx.lloyd_iters(init, 1, 0.0)
    assert st.assign_path == 2
    lab, d = oracle.assign(X, init)
    assert counts.tolist() == np.bincount(lab.astype(np.int64), minlength=32).tolist()
    assert counts[30] == 0 and counts[31] == 0
    X[100, 3] = np.inf
    X[2000, 30] = np.nan
    ctx.set_data(X)
    c, counts, inertia, st = ctx.lloyd_iters(init, 1, 0.0)
    lab, d = oracle.assign(X, init)
    assert counts.tolist() == np.bincount(lab.astype(np.int64), minlength=32).tolist()
    assert not np.isfinite(inertia)


def test_pair_kernel_fused_update_is_reproducible_and_matches_two_pass(lb, ctx, oracle):
    """The fused update (uniform tiles summed inside the screening kernel) against the two-pass path (3xTF32 + update pass):
    same labels, so the same counts; sums within the f32 tolerance; run-to-run bit-identical."""
    X, _ = blobs(50000, 128, 20, np.float32, seed=17, spread=30.0)
    init = X[::2500][:20].copy()
    ctx.set_data(X)
    a = ctx.lloyd_iters(init, 3, 0.0)
    b = ctx.lloyd_iters(init, 3, 0.0)
    assert a[3].assign_path == 2
    assert (a[0] == b[0]).all() and a[2] == b[2] and (a[1] == b[1]).all()
    ctx.set_scoring_passes(3)
    try:
        c = ctx.lloyd_iters(init, 3, 0.0)
    finally:
        ctx.set_scoring_passes(0)
    assert (a[1] == c[1]).all()
    np.testing.assert_allclose(a[0], c[0], rtol=1e-5, atol=1e-5 * np.abs(c[0]).max())
    np.testing.assert_allclose(a[2], c[2], rtol=1e-5)


# ---- warp-specialised pipeline kernel (tc4: f32, d = 32, k <= 64) ----
def _oracle_iters(oracle, X, init, iters):
    cur = init
    for _ in range(iters):
        lab, d = oracle.assign(X, cur)
        nxt = oracle.compute_centroids(cur, X, lab, acc_f64=True)
        last = (lab, d)
        cur = nxt
    return cur, last


PIPE_CASES = [
    # n, k, k_true, spread, shuffle
    (40000, 64, 64, 30.0, True),     # every tile has mixed labels: general accumulation path
    (40000, 64, 64, 30.0, False),    # blob order: uniform tiles, register fast path
    (128 * 148 * 3 + 77, 33, 33, 30.0, True),   # three tiles per CTA plus a tail tile, k padded to 64
    (1000, 17, 17, 5.0, True),       # fewer tiles than CTAs, k <= 32 (32-column variant)
    (127, 2, 2, 30.0, False),        # a single partial tile
    (128, 64, 8, 1.0, True),         # exactly one tile, many near ties
    (50000, 48, 6, 0.7, True),       # heavy overlap: a large share of rows takes the exact re-scan
]


@pytest.mark.parametrize("passes", [0, 1, 2, 3])
@pytest.mark.parametrize("n,k,k_true,spread,shuffle", PIPE_CASES)
def test_pipeline_kernel_matches_oracle(lb, ctx, oracle, n, k, k_true, spread, shuffle, passes):
    dt = np.float32
    X, _ = blobs(n, 32, k_true, dt, seed=7 * n + k, spread=spread)
    g = np.random.default_rng(n + k)
    if shuffle:
        X = X[g.permutation(n)].copy()
    init = X[g.choice(n, size=k, replace=False)].copy()
    ctx.set_assign_path(2)
    ctx.set_scoring_passes(passes)
    try:
        ctx.set_data(X)
        c, counts, inertia, st = ctx.lloyd_iters(init, 4, 0.0)
        assert st.assign_path == 2 and st.n_iter == 4
    finally:
        ctx.set_assign_path(-1)
        ctx.set_scoring_passes(0)
    cur, (lab, d) = _oracle_iters(oracle, X, init, 4)
    scale = np.abs(cur).max()
    np.testing.assert_allclose(c, cur, rtol=1e-5, atol=1e-5 * scale)
    assert counts.tolist() == np.bincount(lab.astype(np.int64), minlength=k).tolist()
    np.testing.assert_allclose(inertia, d.astype(np.float64).sum(), rtol=1e-5)
    print(f"re-scanned rows over 4 iterations: {st.fallback_rows} of {4 * n}")


@pytest.mark.parametrize("scale,offset", [(1e-3, 0.0), (1e3, 0.0), (1.0, 5e3), (1e-20, 0.0), (1e15, 0.0)])
def test_pipeline_kernel_magnitudes(lb, ctx, oracle, scale, offset):
    """The binade window of the packed scores follows max|x|, max|c|: tiny, huge and far-from-origin data."""
    dt = np.float32
    X, _ = blobs(20000, 32, 10, dt, seed=11, spread=30.0)
    X = (X.astype(np.float64) * scale + offset).astype(dt)
    g = np.random.default_rng(5)
    X = X[g.permutation(len(X))].copy()
    init = X[g.choice(len(X), size=10, replace=False)].copy()
    ctx.set_assign_path(2)
    try:
        ctx.set_data(X)
        c, counts, inertia, st = ctx.lloyd_iters(init, 2, 0.0)
    finally:
        ctx.set_assign_path(-1)
    cur, (lab, d) = _oracle_iters(oracle, X, init, 2)
    assert counts.tolist() == np.bincount(lab.astype(np.int64), minlength=10).tolist()
    np.testing.assert_allclose(c, cur, rtol=1e-5, atol=1e-5 * np.abs(cur).max())
    ref_inertia = d.astype(np.float64).sum()
    if np.isfinite(ref_inertia):
        np.testing.assert_allclose(inertia, ref_inertia, rtol=2e-5)


def test_pipeline_kernel_inf_rows(lb, ctx, oracle):
    dt = np.float32
    X, _ = blobs(3000, 32, 5, dt, seed=3)
    X[100, 3] = np.inf
    X[2000, 30] = -np.inf
    init = X[[0, 700, 1300, 1900, 2500]].copy()
    ctx.set_assign_path(2)
    try:
        ctx.set_data(X)
        c, counts, inertia, st = ctx.lloyd_iters(init, 1, 0.0)
    finally:
        ctx.set_assign_path(-1)
    lab, d = oracle.assign(X, init)
    assert counts.tolist() == np.bincount(lab.astype(np.int64), minlength=5).tolist()
    assert not np.isfinite(inertia)


def test_pipeline_kernel_full_fit_matches_oracle(lb, oracle):
    """Through the public API (fit -> predict), convergence by tolerance, k-means++ seeding shared with the oracle."""
    X, _ = blobs(30000, 32, 12, np.float32, seed=21, spread=10.0)
    rng = lb.Xoshiro256Plus.seed_from_u64(9)
    model = (lb.KMeans.params_with_rng(12, rng).n_runs(2).max_n_iterations(50).tolerance(1e-4)
             .init_method(lb.KMeansInit.KMeansPlusPlus).fit(lb.DatasetBase(X)))
    ref = oracle.fit(X, 12, oracle.rng(9), n_runs=2, max_iter=50, tol=1e-4, method=INIT_KMEANSPP, acc_f64=True)
    assert ref["status"] == 0
    np.testing.assert_allclose(model.centroids(), ref["centroids"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(model.inertia(), ref["inertia"], rtol=1e-5)
    assert model.cluster_count().tolist() == ref["cluster_count"].tolist()
    lab, _ = oracle.assign(X, model.centroids())
    assert (model.predict(X) == lab).all()


# ---- f32-screened small-d f64 kernel (lkm_f64s.cuh: d in {2, 4, 8, 16}, k <= 64; BASELINE cfg5 shape) ----
F64S_CASES = [
    # n, d, k, k_true, spread, shuffle
    (100000, 8, 16, 16, 30.0, False),    # cfg5 shape, blob order: one flush per blob
    (100000, 8, 16, 16, 30.0, True),     # every step of the update walk changes cluster
    (64 * 8 * 296 * 2 + 37, 8, 16, 16, 30.0, True),   # two tiles per warp of a full grid plus a ragged tail
    (70001, 8, 13, 13, 3.0, True),       # odd k (padded pair), overlapping blobs
    (50000, 8, 64, 8, 1.0, True),        # k = 64, many near ties -> f64 certified scan / exact re-scan
    (33000, 2, 3, 3, 10.0, False),       # cfg1-like rows of 16 bytes
    (40000, 4, 7, 7, 30.0, True),
    (40000, 16, 32, 32, 30.0, True),
    (63, 8, 16, 4, 30.0, False),         # a single partial tile (forced path)
    (1, 8, 1, 1, 30.0, False),           # one row, one centroid
]


@pytest.mark.parametrize("packed", [1, 0])
@pytest.mark.parametrize("n,d,k,k_true,spread,shuffle", F64S_CASES)
def test_f64_screened_kernel_matches_oracle(lb, oracle, n, d, k, k_true, spread, shuffle, packed, monkeypatch):
    dt = np.float64
    monkeypatch.setenv("LKM_F64S_PACKED", str(packed))
    ctx = lb.Context(0)
    X, _ = blobs(n, d, k_true, dt, seed=3 * n + k, spread=spread)
    g = np.random.default_rng(n + k)
    if shuffle:
        X = X[g.permutation(n)].copy()
    init = X[g.choice(n, size=k, replace=(n < k))].copy() if n >= k else np.repeat(X, k, axis=0)[:k].copy()
    ctx.set_assign_path(3)
    ctx.set_data(X)
    c, counts, inertia, st = ctx.lloyd_iters(init, 4, 0.0)
    assert st.assign_path == 3 and st.n_iter == 4
    cur, (lab, dist) = _oracle_iters(oracle, X, init, 4)
    assert counts.tolist() == np.bincount(lab.astype(np.int64), minlength=k).tolist()
    np.testing.assert_allclose(c, cur, rtol=1e-10, atol=1e-10 * np.abs(cur).max())
    np.testing.assert_allclose(inertia, dist.sum(), rtol=1e-10)
    # the kernel it replaces gives the same labels and (up to summation order) the same sums
    ctx.set_assign_path(1)
    c1, counts1, inertia1, st1 = ctx.lloyd_iters(init, 4, 0.0)
    assert st1.assign_path == 1
    assert counts1.tolist() == counts.tolist()
    np.testing.assert_allclose(c, c1, rtol=1e-12, atol=1e-12 * np.abs(c1).max())
    print(f"re-scanned rows over 4 iterations: {st.fallback_rows} of {4 * n}")


def test_f64_screened_kernel_is_default_and_reproducible(lb, oracle):
    X, _ = blobs(200000, 8, 16, np.float64, seed=5)
    g = np.random.default_rng(1)
    init = X[g.choice(len(X), size=16, replace=False)].copy()
    ctx = lb.Context(0)
    ctx.set_data(X)
    a = ctx.lloyd_iters(init, 3, 0.0)
    b = ctx.lloyd_iters(init, 3, 0.0)
    assert a[3].assign_path == 3
    assert (a[0] == b[0]).all() and (a[1] == b[1]).all() and a[2] == b[2]


@pytest.mark.parametrize("scale,offset", [(1e-3, 0.0), (1e3, 0.0), (1.0, 5e3), (1e-20, 0.0), (1e-160, 0.0), (1e15, 0.0), (1e100, 0.0), (1.0, 1e9)])
def test_f64_screened_kernel_magnitudes(lb, oracle, scale, offset):
    """f32 overflow / underflow / cancellation far from the origin void the certificate; the f64 scan takes over."""
    X, _ = blobs(40000, 8, 10, np.float64, seed=11, spread=30.0)
    X = X * scale + offset
    g = np.random.default_rng(5)
    X = X[g.permutation(len(X))].copy()
    init = X[g.choice(len(X), size=10, replace=False)].copy()
    ctx = lb.Context(0)
    ctx.set_assign_path(3)
    ctx.set_data(X)
    c, counts, inertia, st = ctx.lloyd_iters(init, 2, 0.0)
    cur, (lab, dist) = _oracle_iters(oracle, X, init, 2)
    assert counts.tolist() == np.bincount(lab.astype(np.int64), minlength=10).tolist()
    np.testing.assert_allclose(c, cur, rtol=1e-10, atol=1e-10 * np.abs(cur).max())
    if np.isfinite(dist.sum()):
        np.testing.assert_allclose(inertia, dist.sum(), rtol=1e-10)


def test_f64_screened_kernel_nan_inf_ties(lb, oracle):
    X, _ = blobs(40000, 8, 5, np.float64, seed=3)
    init = X[[0, 9000, 17000, 25000, 33000]].copy()
    init = np.concatenate([init, init[:2]])          # duplicate centroids: exact ties -> lowest index
    X[100, 3] = np.inf
    X[20000, 7] = -np.inf
    X[30000, 0] = np.nan
    ctx = lb.Context(0)
    ctx.set_assign_path(3)
    ctx.set_data(X)
    c, counts, inertia, st = ctx.lloyd_iters(init, 1, 0.0)
    lab, dist = oracle.assign(X, init)
    assert counts.tolist() == np.bincount(lab.astype(np.int64), minlength=7).tolist()
    assert counts[5] == 0 and counts[6] == 0
    assert not np.isfinite(inertia)
    ctx.set_assign_path(1)
    c1, counts1, _, _ = ctx.lloyd_iters(init, 1, 0.0)
    assert counts1.tolist() == counts.tolist()
    np.testing.assert_array_equal(np.isnan(c), np.isnan(c1))


def test_f64_screened_kernel_full_fit_matches_oracle(lb, oracle):
    """Through the public API (fit -> predict -> transform): f64, d = 8, convergence by tolerance, k-means++ seeding and
    best-of-two runs shared with the oracle; n is above the automatic threshold of the screened kernel."""
    X, _ = blobs(60000, 8, 12, np.float64, seed=21, spread=10.0)
    g = np.random.default_rng(2)
    X = X[g.permutation(len(X))].copy()
    rng = lb.Xoshiro256Plus.seed_from_u64(9)
    model = (lb.KMeans.params_with_rng(12, rng).n_runs(2).max_n_iterations(60).tolerance(1e-6)
             .init_method(lb.KMeansInit.KMeansPlusPlus).fit(lb.DatasetBase(X)))
    assert lb.Context.default().last_stats.assign_path == 3
    ref = oracle.fit(X, 12, oracle.rng(9), n_runs=2, max_iter=60, tol=1e-6, method=INIT_KMEANSPP, acc_f64=True)
    assert ref["status"] == 0
    np.testing.assert_allclose(model.centroids(), ref["centroids"], rtol=1e-10, atol=1e-9)
    np.testing.assert_allclose(model.inertia(), ref["inertia"], rtol=1e-10)
    assert model.cluster_count().tolist() == ref["cluster_count"].tolist()
    lab, dist = oracle.assign(X, model.centroids())
    assert (model.predict(X) == lab).all()
    assert (model.transform(X) == dist).all()


def test_kmeans_para_candidate_histogram_through_the_pair_kernel(lb, monkeypatch):
    """k-means|| (KM/init.rs:181-234): for f32 rows of 64 / 128 features the candidate weights (cluster_membership_counts,
    :273-285) come from one tensor-core Lloyd iteration over the candidates; the certified labels make them identical to the
    exact CUDA-core histogram, hence the same seeding for the same rng."""
    X, _ = blobs(120000, 64, 40, np.float32, seed=17, spread=20.0)
    X = X[np.random.default_rng(3).permutation(len(X))].copy()
    out = []
    for flag in ("1", "0"):
        monkeypatch.setenv("LKM_PARA_TC", flag)
        ctx = lb.Context(0)
        params = (lb.KMeans.params_with_rng(40, lb.Xoshiro256Plus.seed_from_u64(5)).init_method(lb.KMeansInit.KMeansPara)
                  .context(ctx).check())
        out.append(params.init_centroids(lb.DatasetBase(X), ctx))
        ctx.close()
    assert (out[0] == out[1]).all()
    # every blob is represented (statistical test of the reference, KM/init.rs:372-449, at this separation)
    d2 = ((X[::1000, None, :] - out[0][None, :, :]) ** 2).sum(-1).min(1)
    assert d2.max() < 64 * 9.0

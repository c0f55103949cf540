"""Pins the CPU oracle against the reference's own test literals.

Each test names the reference test it replays
(algorithms/linfa-clustering/src/k_means/{algorithm,init,hyperparams}.rs).
"""
import numpy as np
import pytest

from tests.oracle_lib import INIT_KMEANSPARA, INIT_KMEANSPP, INIT_PRECOMPUTED, INIT_RANDOM

DTYPES = [np.float32, np.float64]


def function_test_1d(x):
    # algorithm.rs:971-983
    return np.where(x < 0.4, x * x, np.where((x >= 0.4) & (x < 0.8), 3.0 * x + 1.0, np.sin(10.0 * x)))


def make_1d_dataset(oracle, seed=42):
    # algorithm.rs:1016-1019: xt = Array::random_using(100, Uniform::new(0., 1.0), &mut rng)
    rng = oracle.rng(seed)
    xt = rng.uniform_f64(0.0, 1.0, (100, 1))
    return np.concatenate([xt, function_test_1d(xt)], axis=1), rng


# ---- RNG known-answer tests (public algorithm vectors) ----
def test_xoshiro256plus_reference_vector(oracle):
    from tests.oracle_lib import Xoshiro

    g = Xoshiro(oracle.lib, state=[1, 2, 3, 4])
    want = [5, 211106232532999, 211106635186183, 9223759065350669058, 9250833439874351877,
            13862484359527728515, 2346507365006083650, 1168864526675804870, 34095955243042024,
            3466914240207415127]
    assert [g.next_u64() for _ in want] == want


def test_seed_from_u64_splitmix(oracle):
    g = oracle.rng(42)
    assert [int(v) for v in g.state] == [0xBDD732262FEB6E95, 0x28EFE333B266F103, 0x47526757130F9F52,
                                          0x581CE1FF0E4AE394]
    assert [g.next_u64() for _ in range(4)] == [1581911519303979561, 5726079574540882823,
                                                1154208747244521758, 5653213587482834094]
    g = oracle.rng(42)
    assert g.u01_f64() == 0.08575559529546095
    g = oracle.rng(42)
    assert g.next_u32() == 368317477
    g = oracle.rng(40)
    assert int(g.state[0]) == 0x369EAE0B0CA19112
    assert g.next_u64() == 4591798564541617206


# ---- arithmetic KATs ----
@pytest.mark.parametrize("dt", DTYPES)
def test_min_dists(oracle, dt):
    # algorithm.rs:1003-1010
    Cn = np.array([[0.0, 1.0], [40.0, 10.0]], dtype=dt)
    X = np.array([[3.0, 4.0], [1.0, 3.0], [25.0, 15.0]], dtype=dt)
    _, d = oracle.assign(X, Cn)
    assert d.tolist() == [18.0, 5.0, 250.0]


@pytest.mark.parametrize("dt", DTYPES)
def test_oracle_test_for_closest_centroid(oracle, dt):
    # algorithm.rs:1150-1160
    Cn = np.array([[0, 0], [1, 2], [20, 0], [0, 20]], dtype=dt)
    X = np.array([[1, 0.6], [20, 2], [20, 0], [7, 20]], dtype=dt)
    lab, _ = oracle.assign(X, Cn)
    assert lab.tolist() == [0, 2, 2, 3]


def test_nothing_is_closer_than_self(oracle):
    # algorithm.rs:1127-1143
    rng = oracle.rng(42)
    Cn = rng.uniform_f64(-100.0, 100.0, (20, 5))
    lab, d = oracle.assign(Cn, Cn)
    assert lab.tolist() == list(range(20))
    assert np.all(d == 0)


def test_ties_pick_lowest_index(oracle):
    # strict '<' at algorithm.rs:940
    Cn = np.array([[2.0, 0.0], [0.0, 2.0], [2.0, 0.0]])
    lab, _ = oracle.assign(np.array([[1.0, 1.0], [2.0, 0.0]]), Cn)
    assert lab.tolist() == [0, 0]


@pytest.mark.parametrize("dt", DTYPES)
def test_compute_centroids_works(oracle, dt):
    # algorithm.rs:1079-1115 (data from numpy; the reference draws with thread_rng)
    g = np.random.default_rng(0)
    c1 = g.uniform(-100, 100, (100, 4)).astype(dt)
    c2 = g.uniform(-100, 100, (100, 4)).astype(dt)
    X = np.concatenate([c1, c2])
    lab = np.concatenate([np.zeros(100), np.ones(100)]).astype(np.uint64)
    out = oracle.compute_centroids(np.zeros((2, 4), dtype=dt), X, lab)
    tol = 1e-5 if dt == np.float64 else 1e-3
    np.testing.assert_allclose(out[0], c1.astype(np.float64).sum(0) / 101, atol=tol)
    np.testing.assert_allclose(out[1], c2.astype(np.float64).sum(0) / 101, atol=tol)


@pytest.mark.parametrize("dt", DTYPES)
def test_compute_extra_centroids(oracle, dt):
    # algorithm.rs:1117-1125: empty cluster keeps its centroid
    out = oracle.compute_centroids(np.ones((2, 2), dtype=dt), np.array([[1.0, 2.0]], dtype=dt), [0])
    assert out.tolist() == [[1.0, 1.5], [1.0, 1.0]]


def test_compute_inertia(oracle):
    # algorithm.rs:1512-1520
    X = np.array([[0.0, 0.0], [3.0, 4.0]])
    _, d = oracle.assign(X, np.array([[1.0, 1.0]]))
    assert oracle.sum(d) == 15.0


def test_ndarray_sum_order(oracle):
    # ndarray 0.16 unrolled_fold: 8 lanes, (p0+p4)+(p1+p5)+(p2+p6)+(p3+p7), then the tail
    x = (np.arange(1, 20, dtype=np.float32) * np.float32(0.1)).astype(np.float32)
    p = [np.float32(0)] * 8
    for blk in range(2):
        for l in range(8):
            p[l] = np.float32(p[l] + x[8 * blk + l])
    acc = np.float32(0)
    for a, b in ((0, 4), (1, 5), (2, 6), (3, 7)):
        acc = np.float32(acc + np.float32(p[a] + p[b]))
    for v in x[16:]:
        acc = np.float32(acc + v)
    assert oracle.sum(x) == acc


def test_compute_centroids_incremental(oracle):
    # algorithm.rs:1167-1178
    X = np.array([[-1.0, -3.0], [0.0, 0.0], [3.0, 5.0], [5.0, 5.0]])
    Cn, counts = oracle.compute_centroids_incremental(X, [0, 0, 1, 1], [[-1.0, -1.0], [3.0, 4.0], [7.0, 8.0]],
                                                      [3.0, 0.0, 1.0])
    np.testing.assert_allclose(Cn, [[-4 / 5, -6 / 5], [4, 5], [7, 8]], atol=1e-12)
    assert counts.tolist() == [5.0, 2.0, 1.0]


def test_incremental_kmeans(oracle):
    # algorithm.rs:1180-1202
    rng = oracle.rng(45)
    m = ([[-1.0, -1.0], [3.0, 4.0], [7.0, 8.0]], [0.0, 0.0, 0.0])
    r = oracle.fit_with(np.array([[-1.0, -3.0], [0.0, 0.0], [3.0, 5.0], [5.0, 5.0]]), 3, rng, model=m, tol=100.0)
    assert r["status"] == 0
    np.testing.assert_allclose(r["centroids"], [[-0.5, -1.5], [4, 5], [7, 8]], atol=1e-12)
    r = oracle.fit_with(np.array([[-5.0, -5.0], [0.0, 0.0], [10.0, 10.0]]), 3, rng,
                        model=(r["centroids"], r["cluster_count"]), tol=100.0)
    assert r["status"] == 0
    np.testing.assert_allclose(r["centroids"], [[-6 / 4, -8 / 4], [4, 5], [10, 10]], atol=1e-12)


def test_tolerance(oracle):
    # algorithm.rs:1220-1231 and its Hamerly twin :1722-1733 (C = [[4,4]] +- 0.1)
    X = np.array([[1.0, 1.0], [11.0, 11.0]])
    r = oracle.fit_with(X, 1, oracle.rng(45), tol=8.5, method=INIT_PRECOMPUTED, precomputed=[[0.0, 0.0]])
    assert r["status"] == 0
    r = oracle.fit(X, 1, oracle.rng(45), tol=8.5, method=INIT_PRECOMPUTED, precomputed=[[0.0, 0.0]])
    assert r["status"] == 0 and r["n_iter"] == 1
    assert r["centroids"].tolist() == [[4.0, 4.0]]


def test_derived_six_point_kat(oracle):
    # SURVEY Appendix B "derived" row; data from algorithm.rs:1586-1596
    X = np.array([[0, 0], [1, 0], [0, 1], [10, 10], [11, 10], [10, 11]], dtype=np.float64)
    r = oracle.fit(X, 2, oracle.rng(42), n_runs=1, tol=1e-4, method=INIT_PRECOMPUTED,
                   precomputed=[[0.0, 0.0], [10.0, 10.0]])
    assert r["status"] == 0 and r["n_iter"] == 8
    np.testing.assert_allclose(r["centroids"], [[0.33332825, 0.33332825], [10.33332825, 10.33332825]], atol=1e-8)
    assert r["labels"].tolist() == [0, 0, 0, 1, 1, 1]
    assert r["cluster_count"].tolist() == [3.0, 3.0]
    assert abs(r["inertia"] - 0.4444444453) < 1e-9


def test_n_clusters_eq_n_samples(oracle):
    # algorithm.rs:1636-1647 pattern: Precomputed(data) => inertia 0
    X = np.array([[1.0, 2.0], [10.0, 20.0], [-5.0, -5.0], [100.0, 0.0]])
    r = oracle.fit(X, 4, oracle.rng(42), method=INIT_PRECOMPUTED, precomputed=X)
    assert r["status"] == 0 and r["inertia"] == 0.0


def test_single_cluster_and_identical_rows(oracle):
    # algorithm.rs:1624-1673 (Hamerly twins; Lloyd shares the m_k-means update)
    X = np.array([[1.0, 2.0], [3.0, 4.0], [5.0, 6.0], [7.0, 8.0]])
    r = oracle.fit(X, 1, oracle.rng(42))
    np.testing.assert_allclose(r["centroids"], [[4.0, 5.0]], atol=1e-4)
    r = oracle.fit(np.full((4, 2), 5.0), 1, oracle.rng(42))
    np.testing.assert_allclose(r["centroids"], [[5.0, 5.0]], atol=1e-10)
    assert abs(r["inertia"]) < 1e-10
    r = oracle.fit(np.array([[3.0, 7.0]]), 1, oracle.rng(42))
    assert r["centroids"].tolist() == [[3.0, 7.0]] and r["inertia"] == 0.0


def test_param_errors(oracle):
    # hyperparams.rs:204-232
    X = np.zeros((4, 2))
    assert oracle.fit(X, 0, oracle.rng(1))["status"] == 10
    assert oracle.fit(X, 1, oracle.rng(1), n_runs=0)["status"] == 11
    assert oracle.fit(X, 1, oracle.rng(1), tol=0.0)["status"] == 12
    assert oracle.fit(X, 1, oracle.rng(1), tol=-1.0)["status"] == 12
    assert oracle.fit(X, 1, oracle.rng(1), max_iter=0)["status"] == 13


def test_nan_data_is_inertia_error(oracle):
    # algorithm.rs:336,365: NaN inertia never beats +inf => Err(InertiaError)
    X = np.full((8, 2), np.nan)
    r = oracle.fit(X, 2, oracle.rng(1), n_runs=2, method=INIT_PRECOMPUTED, precomputed=np.zeros((2, 2)))
    assert r["status"] == 1


# ---- seeding ----
def test_cluster_membership_counts(oracle):
    # init.rs:326-332
    Cn = np.array([[0.0, 1.0], [40.0, 10.0], [3.0, 9.0]])
    X = np.array([[3.0, 4.0], [1.0, 3.0], [25.0, 15.0]])
    assert oracle.membership_counts(Cn, X).tolist() == [2.0, 1.0, 0.0]


def test_precomputed(oracle):
    # init.rs:305-317
    Cn = np.array([[0.0, 1.0], [40.0, 10.0]])
    X = np.array([[3.0, 4.0], [1.0, 3.0], [25.0, 15.0]])
    out = oracle.init(INIT_PRECOMPUTED, 2, X, oracle.rng(40), precomputed=Cn)
    assert out.tolist() == Cn.tolist()


@pytest.mark.parametrize("rt", [0, 1, 4])
def test_sample_subsequent_candidates(oracle, rt):
    # init.rs:319-324: probabilities 0, 3.55, 4.44 => rows 1 and 2 whatever the RNG says
    assert oracle.sample_subsequent(np.array([0.0, 0.4, 0.5]), 8.0, 0, rayon_threads=rt).tolist() == [1, 2]


@pytest.mark.parametrize("dt", DTYPES)
def test_weighted_kmeans_plusplus(oracle, dt):
    # init.rs:337-357: zero-weight points are never chosen
    g = np.random.default_rng(42)
    X = (g.standard_normal((1000, 2)) * 100).astype(dt)
    w = np.zeros(1000, dtype=dt)
    w[0], w[1] = 2.0, 3.0
    for seed in range(20):
        _, picked = oracle.weighted_kmeanspp(2, X, w, oracle.rng(seed))
        assert sorted(picked.tolist()) == [0, 1]


def test_weighted_index_partition_rule(oracle):
    # rand 0.8.5 WeightedIndex: #{j : cum[j] <= u}; zero-weight items are never returned
    w = np.array([0.0, 1.0, 0.0, 0.0, 3.0, 0.0])
    seen = set()
    for seed in range(200):
        seen.add(oracle.weighted_index(w, oracle.rng(seed)))
    assert seen == {1, 4}
    assert oracle.weighted_index(np.zeros(4), oracle.rng(0)) == -1  # AllWeightsZero
    assert oracle.weighted_index(np.array([1.0, -1.0]), oracle.rng(0)) == -1  # InvalidWeight
    assert oracle.weighted_index(np.array([1.0, np.nan]), oracle.rng(0)) == -1


def _three_blobs(seed=0):
    g = np.random.default_rng(seed)
    return np.concatenate([g.standard_normal((50, 2)) + c for c in (20.0, -1000.0, 1000.0)])


@pytest.mark.parametrize("method", [INIT_KMEANSPP, INIT_KMEANSPARA])
def test_verify_init(oracle, method):
    # init.rs:372-411
    rng = oracle.rng(42)
    deg = np.array([[1.0, 2.0]])
    out = oracle.init(method, 2, deg, rng)
    assert out.tolist() == [[1.0, 2.0], [1.0, 2.0]]
    X = _three_blobs()
    out = oracle.init(method, 3, X, rng)
    blobs = set()
    for row in out:
        idx = np.where((X == row).all(axis=1))[0]
        assert idx.size >= 1
        blobs.add(int(idx[0]) // 50)
    assert blobs == {0, 1, 2}


def test_compare_init_losses(oracle):
    # init.rs:421-439: loss(++) < loss(random), loss(||) < loss(random)
    X = _three_blobs(1)
    rng = oracle.rng(42)

    def loss(Cn):
        return oracle.sum(oracle.assign(X, Cn)[1])

    l_rand = loss(oracle.init(INIT_RANDOM, 3, X, rng.clone()))
    l_pp = loss(oracle.init(INIT_KMEANSPP, 3, X, rng.clone()))
    l_para = loss(oracle.init(INIT_KMEANSPARA, 3, X, rng))
    assert l_pp < l_rand and l_para < l_rand


def test_random_init_picks_distinct_rows(oracle):
    # init.rs:105-113 / rand 0.8.5 seq::index::sample (Floyd for amount < 12; the
    # other branches for larger amounts)
    for length, amount in [(100, 3), (100, 6), (100, 40), (1000, 60), (600000, 20), (5000, 200), (100000, 170)]:
        idx = oracle.index_sample(oracle.rng(7), length, amount)
        assert len(set(idx.tolist())) == amount and idx.max() < length
    with pytest.raises(ValueError):
        oracle.index_sample(oracle.rng(7), 3, 4)


# ---- self-consistency families (algorithm.rs:1015-1077, 1233-1248) ----
@pytest.mark.parametrize("method", [INIT_RANDOM, INIT_KMEANSPP, INIT_KMEANSPARA])
def test_n_runs(oracle, method):
    data, rng = make_1d_dataset(oracle)
    r1 = oracle.fit(data, 3, rng.clone(), n_runs=1, method=method)
    assert r1["status"] == 0
    lab, dists = oracle.assign(data, r1["centroids"])
    inertia = sum(oracle.sq_l2_dist(data[i], r1["centroids"][int(lab[i])]) for i in range(100))
    assert abs(inertia - dists.sum()) < 1e-5
    one, _ = oracle.assign(data[:1], r1["centroids"])
    assert one[0] == lab[0]
    r2 = oracle.fit(data, 3, rng.clone(), method=method)
    _, dists2 = oracle.assign(data, r2["centroids"])
    if method == INIT_RANDOM:
        assert dists2.sum() <= dists.sum()


def test_max_n_iterations(oracle):
    data, rng = make_1d_dataset(oracle)
    r = oracle.fit(data, 6, rng.clone(), n_runs=1, max_iter=5, method=INIT_RANDOM)
    assert r["status"] == 0 and r["n_iter"] <= 5


def test_tutorial_blobs(oracle):
    # algorithm.rs:122-154 (blobs drawn with numpy; the assertion is on geometry)
    g = np.random.default_rng(42)
    centres = np.array([[0.0, 1.0], [-10.0, 20.0], [-1.0, 10.0]])
    X = np.concatenate([g.standard_normal((100, 2)) + c for c in centres])
    r = oracle.fit(X, 3, oracle.rng(42), max_iter=200, tol=1e-2)
    lab, _ = oracle.assign(np.array([[-9.0, 20.5]]), r["centroids"])
    np.testing.assert_allclose(r["centroids"][int(lab[0])], [-10.0, 20.0], atol=0.3)
    assert r["cluster_count"].sum() == 300


@pytest.mark.parametrize("dt", DTYPES)
def test_f64_accumulated_variant_close_to_faithful(oracle, dt):
    g = np.random.default_rng(3)
    centres = g.uniform(-30, 30, (5, 3))
    X = np.concatenate([g.standard_normal((400, 3)) + c for c in centres]).astype(dt)
    init = X[[0, 400, 800, 1200, 1600]]
    a = oracle.fit(X, 5, oracle.rng(1), n_runs=1, method=INIT_PRECOMPUTED, precomputed=init)
    b = oracle.fit(X, 5, oracle.rng(1), n_runs=1, method=INIT_PRECOMPUTED, precomputed=init, acc_f64=True)
    assert (a["labels"] == b["labels"]).all()
    rtol = 1e-5 if dt == np.float32 else 1e-12
    np.testing.assert_allclose(a["centroids"], b["centroids"], rtol=rtol, atol=rtol)
    np.testing.assert_allclose(a["inertia"], b["inertia"], rtol=rtol)

"""The rows either side of the K-Means path: silhouette score (src/metrics_clustering.rs), .npy files
(examples/kmeans.rs:40-42) and generate::blobs (datasets/src/generate.rs)."""
import numpy as np
import pytest


def _silhouette_cases():
    # test_silhouette_score (src/metrics_clustering.rs:145-184)
    a = np.concatenate([np.linspace(0, 1, 10), np.linspace(10000, 10001, 10)])[:, None]
    yield np.concatenate([a, a], axis=1), np.array([0] * 10 + [1] * 10), lambda s: abs(s - 1.0) < 1e-3
    b = np.concatenate([np.linspace(0, 1, 5), np.linspace(1, 2, 5), np.linspace(10000, 10001, 5), np.linspace(10001, 10002, 5)])[:, None]
    yield np.concatenate([b, b], axis=1), np.array([0] * 5 + [1] * 5 + [0] * 5 + [1] * 5), lambda s: s < 0
    c = np.linspace(0, 10, 100)[:, None]
    yield np.concatenate([c, c], axis=1), (np.arange(100) + 3) % 48, lambda s: s < -0.5


def test_oracle_silhouette_reference_literals(oracle):
    for X, labels, ok in _silhouette_cases():
        assert ok(oracle.silhouette(X, labels))
    # test_empty_labels_as_single_label (:186-192)
    assert oracle.silhouette(np.linspace(0, 1, 10)[:, None], np.zeros(10)) == 1.0


def test_npy_round_trip_against_numpy(tmp_path):
    import linfa_b200 as lb

    g = np.random.default_rng(0)
    for arr in (g.standard_normal((7, 3)), g.standard_normal((5, 4)).astype(np.float32), g.integers(0, 9, 11).astype(np.uint64),
                np.zeros((0, 3)), g.standard_normal((2, 3, 4))):
        p = tmp_path / "a.npy"
        lb.write_npy(p, arr)
        back = np.load(p)                       # numpy reads what the library wrote
        assert back.dtype == arr.dtype and back.shape == arr.shape and (back == arr).all()
        assert p.stat().st_size == 128 + arr.nbytes or (p.stat().st_size - arr.nbytes) % 64 == 0
        np.save(p, arr)
        mine = lb.read_npy(p)                   # and the library reads what numpy wrote
        assert mine.dtype == arr.dtype and mine.shape == arr.shape and (mine == arr).all()
    with pytest.raises(TypeError):
        lb.write_npy(tmp_path / "b.npy", np.zeros(3, dtype=np.int32))
    np.save(tmp_path / "c.npy", np.asfortranarray(g.standard_normal((3, 2))))
    with pytest.raises(lb.LkmError):
        lb.read_npy(tmp_path / "c.npy")


@pytest.mark.gpu
def test_silhouette_reference_literals_on_gpu(oracle):
    import linfa_b200 as lb

    for X, labels, ok in _silhouette_cases():
        s = lb.DatasetBase(X, labels).silhouette_score()
        assert ok(s) and s == oracle.silhouette(X, labels)
    assert lb.DatasetBase(np.linspace(0, 1, 10)[:, None]).silhouette_score() == 1.0


@pytest.mark.gpu
@pytest.mark.parametrize("dt", [np.float32, np.float64])
@pytest.mark.parametrize("n,d,k", [(700, 3, 4), (1500, 16, 7), (900, 37, 3), (300, 130, 5)])
def test_silhouette_matches_oracle_bitwise(oracle, dt, n, d, k):
    """same operations in the same order as the reference (per-pair sum in ndarray's unrolled order, sqrt, running totals
    per cluster in sample order, scores added in sample order): bit-identical to the oracle"""
    import linfa_b200 as lb

    g = np.random.default_rng(n + d)
    centres = g.uniform(-10, 10, (k, d))
    lab = g.integers(0, k, n)
    X = (centres[lab] + g.standard_normal((n, d))).astype(dt)
    labels = (lab * 7 + 3).astype(np.uint64)     # arbitrary label values
    labels[5] = 1000                              # a cluster of one sample: a = 0
    got = lb.DatasetBase(X, labels).silhouette_score()
    want = oracle.silhouette(X, labels)
    assert got == want, (got, want)


@pytest.mark.gpu
def test_example_pipeline_blobs_fit_predict_silhouette_npy(tmp_path):
    """examples/kmeans.rs: blobs -> fit -> predict -> write_npy, plus the silhouette of the result"""
    import linfa_b200 as lb

    rng = lb.Xoshiro256Plus.seed_from_u64(42)
    expected = np.array([[10.0, 10.0], [1.0, 12.0], [20.0, 30.0], [-20.0, 30.0]])
    X = lb.generate.blobs(1000, expected, rng)
    assert X.shape == (4000, 2) and X.dtype == np.float64
    for b in range(4):   # blob b owns rows [b * size, (b + 1) * size) and is centred on its centroid with unit variance
        blob = X[b * 1000:(b + 1) * 1000]
        assert np.abs(blob.mean(axis=0) - expected[b]).max() < 0.15 and abs((blob - expected[b]).std() - 1.0) < 0.1
    model = lb.KMeans.params_with_rng(4, rng).max_n_iterations(200).tolerance(1e-5).fit(lb.DatasetBase(X))
    ds = model.predict(lb.DatasetBase(X))
    got = np.sort(model.centroids(), axis=0)
    np.testing.assert_allclose(got, np.sort(expected, axis=0), atol=0.2)
    lb.write_npy(tmp_path / "clustered_dataset.npy", ds.records)
    lb.write_npy(tmp_path / "clustered_memberships.npy", ds.targets.astype(np.uint64))
    assert (np.load(tmp_path / "clustered_dataset.npy") == X).all()
    assert (np.load(tmp_path / "clustered_memberships.npy") == ds.targets).all()
    assert ds.silhouette_score() > 0.6

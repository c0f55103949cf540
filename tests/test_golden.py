"""Committed golden fixtures (tests/golden/): the reference's own test literals pin the oracle, the frozen oracle
outputs pin both the oracle (CPU, here) and the CUDA path through the C ABI (GPU)."""
import json
import os

import numpy as np
import pytest

from tests.oracle_lib import INIT_PRECOMPUTED, Xoshiro

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
KATS = json.load(open(os.path.join(GOLD, "reference_kats.json")))
FITS = np.load(os.path.join(GOLD, "oracle_fits.npz"))
FIT_NAMES = sorted({k.split("/")[0] for k in FITS.files})
DTYPES = [np.float32, np.float64]


# ---------------------------------------------------------------- oracle vs the reference's literals (CPU)
@pytest.mark.parametrize("dt", DTYPES)
def test_oracle_reference_kats(oracle, dt):
    k = KATS["min_sq_dists"]
    _, d = oracle.assign(np.array(k["observations"], dtype=dt), np.array(k["centroids"], dtype=dt))
    assert d.tolist() == k["expected"]
    k = KATS["closest_centroid"]
    lab, _ = oracle.assign(np.array(k["observations"], dtype=dt), np.array(k["centroids"], dtype=dt))
    assert lab.tolist() == k["expected"]
    k = KATS["compute_extra_centroids"]
    c = oracle.compute_centroids(np.array(k["old_centroids"], dtype=dt), np.array(k["observations"], dtype=dt),
                                 np.array(k["memberships"], dtype=np.uint64))
    assert c.tolist() == k["expected"]
    k = KATS["compute_inertia"]
    X, C0 = np.array(k["observations"], dtype=dt), np.array(k["centroids"], dtype=dt)
    assert sum(oracle.sq_l2_dist(X[i], C0[m]) for i, m in enumerate(k["memberships"])) == k["expected"]
    k = KATS["membership_counts"]
    assert oracle.membership_counts(np.array(k["centroids"], dtype=dt), np.array(k["observations"], dtype=dt)).tolist() == k["expected"]
    k = KATS["tolerance"]
    r = oracle.fit(np.array(k["observations"], dtype=dt), 1, oracle.rng(42), n_runs=1, max_iter=300, tol=k["tolerance"],
                   method=INIT_PRECOMPUTED, precomputed=np.array(k["init"], dtype=dt))
    assert r["status"] == 0 and r["n_iter"] == k["expected_n_iter"]
    np.testing.assert_allclose(r["centroids"], k["expected_centroids"], atol=1e-6)
    k = KATS["derived_six_points"]
    r = oracle.fit(np.array(k["observations"], dtype=np.float64), 2, oracle.rng(42), n_runs=1, max_iter=300, tol=k["tolerance"],
                   method=INIT_PRECOMPUTED, precomputed=np.array(k["init"], dtype=np.float64))
    assert r["n_iter"] == k["expected_n_iter"]
    np.testing.assert_allclose(r["centroids"], k["expected_centroids"], atol=1e-7)
    np.testing.assert_allclose(r["inertia"], k["expected_inertia"], rtol=1e-8)


def test_oracle_rng_kats(oracle):
    k = KATS["xoshiro256plus_state_1234"]
    g = Xoshiro(oracle.lib, state=k["state"])
    assert [g.next_u64() for _ in k["expected"]] == k["expected"]
    k = KATS["seed_from_u64_42"]
    g = oracle.rng(42)
    assert [hex(int(v)) for v in g.state] == k["expected_state"]
    assert [g.next_u64() for _ in k["expected_first_u64"]] == k["expected_first_u64"]


@pytest.mark.parametrize("name", FIT_NAMES)
def test_oracle_reproduces_frozen_fits(oracle, name):
    X = FITS[name + "/X"]
    k, seed, n_runs, max_iter, method = (int(v) for v in FITS[name + "/params"])
    r = oracle.fit(X, k, oracle.rng(seed), n_runs=n_runs, max_iter=max_iter, tol=float(FITS[name + "/tol"][0]), method=method,
                   acc_f64=True)
    assert r["status"] == 0
    assert (r["centroids"] == FITS[name + "/centroids"]).all()
    assert r["inertia"] == FITS[name + "/inertia"][0]
    lab, d = oracle.assign(X, r["centroids"])
    assert (lab == FITS[name + "/labels"]).all() and (d == FITS[name + "/min_sq_dists"]).all()


# ---------------------------------------------------------------- CUDA path vs the frozen fits (GPU)
@pytest.mark.gpu
@pytest.mark.parametrize("name", FIT_NAMES)
def test_gpu_matches_frozen_fits(name):
    import linfa_b200 as lb

    X = FITS[name + "/X"]
    k, seed, n_runs, max_iter, method = (int(v) for v in FITS[name + "/params"])
    init = [lb.KMeansInit.Random, None, lb.KMeansInit.KMeansPlusPlus][method]
    tol = {np.dtype(np.float32): 1e-5, np.dtype(np.float64): 1e-10}[X.dtype]
    model = (lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(seed)).n_runs(n_runs).max_n_iterations(max_iter)
             .tolerance(float(FITS[name + "/tol"][0])).init_method(init).fit(lb.DatasetBase(X)))
    want = FITS[name + "/centroids"]
    # Two runs that reach the same optimum from different seedings differ only in the order of the centroids and
    # in the last bits of the inertia, so which of them "wins" the strict < of KM/algorithm.rs:336 depends on the
    # summation order (unpinned by the reference): compare the solutions in canonical (sorted) order.
    got = model.centroids()
    og, ow = np.lexsort(got.T[::-1]), np.lexsort(want.T[::-1])
    np.testing.assert_allclose(got[og], want[ow], rtol=tol, atol=tol * np.abs(want).max())
    np.testing.assert_allclose(model.inertia(), FITS[name + "/inertia"][0], rtol=tol)
    assert model.cluster_count()[og].tolist() == FITS[name + "/cluster_count"][ow].tolist()
    # labels / min squared distances against the frozen centroids: identical / bit-identical
    frozen = lb.KMeans(want, FITS[name + "/cluster_count"], float(FITS[name + "/inertia"][0]))
    assert (frozen.predict(X) == FITS[name + "/labels"]).all()
    assert (frozen.transform(X) == FITS[name + "/min_sq_dists"]).all()

"""Gaussian mixture (algorithms/linfa-clustering/src/gaussian_mixture/): the numpy oracle against the reference's own test
expectations (CPU), and the CUDA path against the oracle (GPU, through the C ABI)."""
import numpy as np
import pytest

from oracle import gmm_oracle


def _two_gaussians(n=500, seed=42):
    # test_gmm_fit (gaussian_mixture/algorithm.rs:541-580): weights [0.5, 0.5], means [[0, 0], [20, 20]], the two covariances
    g = np.random.default_rng(seed)
    c0, c1 = np.array([[1.0, 0.7], [0.7, 1.0]]), np.array([[1.0, -0.6], [-0.6, 1.0]])
    a = g.multivariate_normal([0.0, 0.0], c0, n)
    b = g.multivariate_normal([20.0, 20.0], c1, n)
    return np.concatenate([a, b]), (c0, c1)


def _kmeans_resp(X, k, seed=0):
    from sklearn.cluster import KMeans as SK

    lab = SK(k, n_init=3, random_state=seed).fit_predict(X)
    resp = np.zeros((X.shape[0], k))
    resp[np.arange(X.shape[0]), lab] = 1.0
    return resp


def test_oracle_gmm_fit_reference_expectations():
    X, covs = _two_gaussians()
    m = gmm_oracle.fit(X, 2, _kmeans_resp(X, 2))
    order = np.argsort(m["means"][:, 0])
    np.testing.assert_allclose(m["weights"][order], [0.5, 0.5], atol=1e-1)
    np.testing.assert_allclose(m["means"][order], [[0.0, 0.0], [20.0, 20.0]], atol=1e-1 * 3)
    for c, want in zip(order, covs):
        np.testing.assert_allclose(m["covariances"][c], want, atol=0.2)
    # test_predict_proba (:760-781): rows of the responsibilities sum to one
    _, log_resp = gmm_oracle.estimate_log_prob_resp(X, m["weights"], m["means"], m["precisions_chol"])
    np.testing.assert_allclose(np.exp(log_resp).sum(axis=1), 1.0, atol=1e-6)


def test_oracle_zeroed_reg_covar_failure():
    # test_zeroed_reg_covar_failure (:627-651): identical points, reg_covar = 0 -> NotPositiveDefinite; the default 1e-6 fits
    X = np.tile(np.array([[1.0, 1.0]]), (30, 1))
    with pytest.raises(gmm_oracle.GmmError):
        gmm_oracle.fit(X, 1, np.ones((30, 1)), reg_covar=0.0)
    assert gmm_oracle.fit(X, 1, np.ones((30, 1)))["n_iter"] is not None


def test_gmm_param_checks_without_gpu():
    import linfa_b200 as lb

    P = lb.GaussianMixtureModel.params
    for bad in (P(0), P(1).tolerance(0.0), P(1).reg_covariance(-1e-6), P(1).n_runs(0), P(1).max_n_iterations(0)):
        with pytest.raises(lb.GmmError) as e:   # test_invalid_* (:697-757)
            bad.check()
        assert e.value.kind == "InvalidValue"


@pytest.mark.gpu
@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-8), (np.float32, 2e-4)])
def test_gmm_e_step_and_predict_match_oracle(dt, tol):
    import linfa_b200 as lb

    g = np.random.default_rng(5)
    n, d, k = 3000, 5, 4
    X = (g.uniform(-5, 5, (k, d))[g.integers(0, k, n)] + g.standard_normal((n, d))).astype(dt)
    model = gmm_oracle.fit(X.astype(np.float64), k, _kmeans_resp(X, k))
    cast = {kk: (v.astype(dt) if isinstance(v, np.ndarray) else v) for kk, v in model.items()}
    gm = lb.GaussianMixtureModel(cast["weights"], cast["means"], cast["covariances"], cast["precisions"], cast["precisions_chol"])
    lpn, log_resp = gmm_oracle.estimate_log_prob_resp(X, cast["weights"], cast["means"], cast["precisions_chol"])
    proba = gm.predict_proba(X)
    np.testing.assert_allclose(proba, np.exp(log_resp), atol=tol, rtol=tol)
    np.testing.assert_allclose(proba.sum(axis=1), 1.0, atol=1e-5)
    want = np.exp(log_resp).argmax(axis=1)
    got = gm.predict(X)
    clear = np.sort(np.exp(log_resp), axis=1)[:, -1] - np.sort(np.exp(log_resp), axis=1)[:, -2] > 1e-4
    assert (got[clear] == want[clear]).all()
    np.testing.assert_allclose(gm.score(X), lpn.mean(), rtol=max(tol, 1e-6))


@pytest.mark.gpu
@pytest.mark.parametrize("dt,tol", [(np.float64, 1e-8), (np.float32, 5e-4)])
def test_gmm_fit_matches_oracle_from_random_init(dt, tol):
    """GmmInitMethod::Random: the responsibilities come from the rng stream (Uniform(0, 1), row-normalised, :148-157), which
    the test reproduces with the oracle's xoshiro — same EM trajectory on both sides."""
    import linfa_b200 as lb
    from tests import oracle_lib

    X, _ = _two_gaussians(400, seed=3)
    X = X.astype(dt)
    n, k = X.shape[0], 2
    rng = oracle_lib.load().rng(7)
    r = rng.uniform_f64(0.0, 1.0, (n, k))
    resp0 = (r / r.sum(axis=1, keepdims=True)).astype(dt)
    ref = gmm_oracle.fit(X, k, resp0, tolerance=1e-4, max_n_iterations=200)
    m = (lb.GaussianMixtureModel.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(7)).init_method(lb.GmmInitMethod.Random)
         .tolerance(1e-4).max_n_iterations(200).fit(lb.DatasetBase(X)))
    assert abs(m.n_iter - ref["n_iter"]) <= 1
    np.testing.assert_allclose(m.weights(), ref["weights"], atol=tol)
    np.testing.assert_allclose(m.means(), ref["means"], atol=tol * 20)
    np.testing.assert_allclose(m.covariances(), ref["covariances"], atol=tol * 20)
    np.testing.assert_allclose(m.precisions(), ref["precisions"], rtol=tol * 50, atol=tol * 20)
    np.testing.assert_allclose(m.lower_bound, ref["lower_bound"], rtol=max(tol, 1e-6))


@pytest.mark.gpu
def test_gmm_fit_kmeans_init_reference_tests():
    import linfa_b200 as lb

    # test_gmm_fit (:541-580)
    X, covs = _two_gaussians()
    m = lb.GaussianMixtureModel.params(2).fit(lb.DatasetBase(X))
    order = np.argsort(m.means()[:, 0])
    np.testing.assert_allclose(m.weights()[order], [0.5, 0.5], atol=1e-1)
    np.testing.assert_allclose(m.means()[order], [[0.0, 0.0], [20.0, 20.0]], atol=0.3)
    for c, want in zip(order, covs):
        np.testing.assert_allclose(m.covariances()[c], want, atol=0.2)
    np.testing.assert_allclose(m.precisions()[order[0]], np.linalg.inv(m.covariances()[order[0]]), rtol=1e-6, atol=1e-8)
    # test_centroids_prediction (:672-695): blobs around four centres, centroids within 1 of them
    rng = lb.Xoshiro256Plus.seed_from_u64(42)
    centres = np.array([[0.0, 1.0], [-10.0, 20.0], [-1.0, 10.0], [30.0, 0.0]])
    Xb = lb.generate.blobs(1000, centres, rng)
    gm = lb.GaussianMixtureModel.params_with_rng(4, rng).fit(lb.DatasetBase(Xb))
    got = gm.centroids()[np.lexsort(gm.centroids().T[::-1])]
    np.testing.assert_allclose(got, centres[np.lexsort(centres.T[::-1])], atol=1.0)
    # predict labels the blobs consistently
    lab = gm.predict(Xb)
    assert all(len(set(lab[b * 1000:(b + 1) * 1000].tolist())) == 1 for b in range(4))
    # test_zeroed_reg_covar_failure (:627-651)
    same = np.tile(np.array([[1.0, 1.0]]), (30, 1))
    with pytest.raises(lb.GmmError) as e:
        lb.GaussianMixtureModel.params(1).reg_covariance(0.0).fit(lb.DatasetBase(same))
    assert e.value.kind == "LinalgError"
    assert lb.GaussianMixtureModel.params(1).fit(lb.DatasetBase(same)).weights().tolist() == [1.0]
    # test_invalid_* through the C ABI
    with pytest.raises(lb.GmmError):
        lb.GaussianMixtureModel.params(0).fit(lb.DatasetBase(X))

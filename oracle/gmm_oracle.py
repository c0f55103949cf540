"""ORACLE — TEST INFRASTRUCTURE ONLY.  numpy restatement of linfa-clustering's Gaussian mixture model
(algorithms/linfa-clustering/src/gaussian_mixture/algorithm.rs), full covariances.  Only tests/ may import it.

Every function cites the reference lines it follows.  The arithmetic is numpy's in the array dtype (f32 / f64); the
summation order inside `dot` / `sum` is numpy's, not ndarray's ("parity unpinned" like the K-Means oracle's sums: the
reference's own GMM tests compare at 1e-1 ... 1e-6, tests/test_gmm.py replays them)."""
import numpy as np


class GmmError(Exception):
    pass


def estimate_gaussian_parameters(X, resp, reg_covar):
    """:217-259"""
    F = X.dtype.type
    nk = resp.sum(axis=0)
    if nk.min() < F(10.0) * np.finfo(F).eps:
        raise GmmError(f"EmptyCluster #{int(nk.argmin()) + 1}")
    means = resp.T.dot(X) / nk[:, None]
    k, d = means.shape
    cov = np.zeros((k, d, d), dtype=X.dtype)
    for c in range(k):
        diff = X - means[c]
        m = diff.T * resp[:, c]
        cov_c = m.dot(diff) / nk[c]
        cov_c[np.diag_indices(d)] += F(reg_covar)
        cov[c] = cov_c
    return nk, means, cov


def compute_precisions_cholesky_full(cov):
    """:261-276: decomp = cholesky(cov_k); sol = L^-1; precisions_chol_k = sol^T"""
    k, d, _ = cov.shape
    out = np.zeros_like(cov)
    for c in range(k):
        try:
            L = np.linalg.cholesky(cov[c].astype(np.float64)).astype(cov.dtype)
        except np.linalg.LinAlgError as e:
            raise GmmError("NotPositiveDefinite") from e
        sol = np.linalg.solve(L.astype(np.float64), np.eye(d)).astype(cov.dtype)
        out[c] = sol.T
    return out


def compute_precisions_full(pchol):
    """:278-289"""
    return np.stack([p.dot(p.T) for p in pchol])


def estimate_log_prob_resp(X, weights, means, pchol):
    """:339-409"""
    F = X.dtype.type
    n, d = X.shape
    k = means.shape[0]
    log_det = np.log(np.stack([np.diag(p) for p in pchol])).sum(axis=1)
    log_prob = np.zeros((n, k), dtype=X.dtype)
    for c in range(k):
        diff = (X - means[c]).dot(pchol[c])
        log_prob[:, c] = (diff * diff).sum(axis=1)
    log_prob = F(-0.5) * (log_prob + F(d * np.log(2.0 * np.pi))) + log_det
    wlp = log_prob + np.log(weights)
    log_max = wlp.max(axis=1)
    shifted = wlp - log_max[:, None]
    log_prob_norm = np.log(np.exp(shifted).sum(axis=1)) + log_max
    log_resp = wlp - log_prob_norm[:, None]
    return log_prob_norm, log_resp


def fit(X, k, resp0, tolerance=1e-3, reg_covar=1e-6, n_runs=1, max_n_iterations=100):
    """GaussianMixtureModel::new (:124-178) from given initial responsibilities, then Fit::fit (:411-467).
    Returns dict(weights, means, covariances, precisions, precisions_chol, lower_bound, n_iter)."""
    F = X.dtype.type
    n = X.shape[0]
    nk, means, cov = estimate_gaussian_parameters(X, resp0.astype(X.dtype), reg_covar)
    weights = nk / F(n)
    pchol = compute_precisions_cholesky_full(cov)
    max_lb, best, best_iter = -np.inf, None, None
    for _ in range(n_runs):
        lb, conv = -np.inf, None
        for it in range(max_n_iterations):
            prev = lb
            lpn, log_resp = estimate_log_prob_resp(X, weights, means, pchol)
            nk, means, cov = estimate_gaussian_parameters(X, np.exp(log_resp), reg_covar)
            weights = nk / F(n)
            pchol = compute_precisions_cholesky_full(cov)
            lb = lpn.mean()
            if abs(lb - prev) < tolerance:
                conv = it
                break
        if lb > max_lb:
            max_lb = lb
            best = dict(weights=weights.copy(), means=means.copy(), covariances=cov.copy(), precisions_chol=pchol.copy(),
                        precisions=compute_precisions_full(pchol), lower_bound=lb)
            best_iter = conv
    if best_iter is None:
        raise GmmError("NotConverged")
    best["n_iter"] = best_iter
    return best


def predict(X, model):
    """:470-484"""
    _, log_resp = estimate_log_prob_resp(X, model["weights"], model["means"], model["precisions_chol"])
    return np.exp(log_resp).argmax(axis=1)

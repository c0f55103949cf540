"""Full-size checks (BASELINE.json configs[1..4] per-GPU shapes) through size-independent properties.

The oracle cannot run these sizes in seconds, so the CUDA path is checked against what the blob layout of
datasets/src/generate.rs:28-58 (restated by gen_blobs_kernel) implies: blob b owns rows [b m, (b + 1) m), m = n / k,
the last blob takes the remainder; centres ~ U(-30, 30)^d, noise N(0, 1).  Starting from centres + N(0, 0.5) every row's
nearest centroid is its own blob's, so after one Lloyd iteration (KM/algorithm.rs:314-331, 785-810)

  * the cluster counts are exactly the blob sizes (every row assigned once, labels right),
  * centroid b = (sum of blob b + old centroid) / (m + 1) lies within a few standard errors of the blob centre,
  * inertia / n is the chi-square mean d (+ the offset of the perturbed start) within a percent,
  * the result is bit-identical run to run, and a second iteration moves the centroids by a vanishing shift,
  * (cfg2) the fused kernel's inertia equals the sum of the bit-exact per-row distances of the assignment pass, and the
    label histogram of that pass equals the counts.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CONFIGS = [
    # name, n, d, k, dtype, assign path expected
    ("cfg2", 1_000_000, 32, 64, np.float32, 2),
    ("cfg3", 10_000_000, 128, 256, np.float32, 2),
    ("cfg4-shard", 12_500_000, 64, 1024, np.float32, 2),
    ("cfg5-shard", 125_000_000, 8, 16, np.float64, 3),
]


@pytest.fixture(scope="module")
def lb():
    import torch

    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    import linfa_b200

    return linfa_b200


def _separated_centres(k, d, dtype, seed):
    """U(-30, 30)^d centres (benches/k_means.rs:48-49); re-drawn until every pair is >= 20 sigma apart (the first draw, practically)."""
    g = np.random.default_rng(seed)
    for _ in range(200):
        c = g.uniform(-30.0, 30.0, (k, d))
        d2 = ((c[:, None, :] - c[None, :, :]) ** 2).sum(-1) + np.eye(k) * 1e9
        if d2.min() >= 400.0:
            return c.astype(dtype)
    raise AssertionError("no separated centres found")


@pytest.mark.parametrize("name,n,d,k,dtype,path", CONFIGS)
def test_full_size_blob_properties(lb, name, n, d, k, dtype, path):
    ctx = lb.Context(0)
    try:
        centres = _separated_centres(k, d, dtype, seed=1234 + d)
        g = np.random.default_rng(99)
        start = (centres.astype(np.float64) + 0.5 * g.standard_normal(centres.shape)).astype(dtype)
        ctx.gen_blobs(n, d, centres, 1.0, seed=2024)
        c1, counts, inertia, st = ctx.lloyd_iters(start, 1, 0.0)
        assert st.assign_path == path and st.n_iter == 1
        m = n // k
        sizes = [m] * (k - 1) + [n - m * (k - 1)]
        assert counts.tolist() == sizes, "every row must land in its own blob's cluster"
        # centroid b = (sum + old) / (m + 1): within 6 standard errors of the centre (+ the old centroid's 1 / (m + 1) share)
        tol_c = 6.0 / np.sqrt(m) + 4.0 / (m + 1)
        assert np.abs(c1.astype(np.float64) - centres.astype(np.float64)).max() < tol_c
        # inertia of the FIRST assignment is against the perturbed start: E = n d sigma^2 + sum_b m_b |start_b - centre_b|^2
        off = ((start.astype(np.float64) - centres.astype(np.float64)) ** 2).sum(1)
        expect = n * d + float(np.dot(sizes, off))
        assert abs(float(inertia) - expect) / expect < 1e-2
        # bit-reproducible
        c1b, countsb, inertiab, _ = ctx.lloyd_iters(start, 1, 0.0)
        assert (c1 == c1b).all() and (counts == countsb).all() and inertia == inertiab
        # second iteration: same labels, centroids at the blob means, inertia = chi-square mean, tiny shift afterwards
        c2, counts2, inertia2, st2 = ctx.lloyd_iters(c1, 1, 0.0)
        assert counts2.tolist() == sizes
        assert abs(float(inertia2) / (n * d) - 1.0) < 1e-2
        assert np.abs(c2.astype(np.float64) - c1.astype(np.float64)).max() < 2.0 * tol_c / np.sqrt(m) + 1e-3 * tol_c
        if name == "cfg2":
            labels, dists = ctx.assign(start)
            assert np.bincount(labels.astype(np.int64), minlength=k).tolist() == sizes
            assert (labels == np.repeat(np.arange(k), sizes)).all()
            np.testing.assert_allclose(float(inertia), dists.astype(np.float64).sum(), rtol=1e-5)
    finally:
        ctx.close()

"""Multi-GPU parity inside `pytest -m gpu`: spawns tests/multi_gpu_check.py under torchrun on every
GPU count in {2, 4, 8} the box offers.  One process per GPU, rows sharded, per-rank partials exchanged over NVLink
and combined in rank order: cluster counts and per-row labels identical to one GPU holding every row, centroids within
tolerance of it and BIT-IDENTICAL on every rank.  Skipped (not passed) on a single-GPU box; the host-side protocol is
covered on CPU by tests/test_distributed_cpu.py (gloo, world_size 2)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _gpu_count():
    import torch

    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world,exchange", [(2, "tagged"), (2, "flags"), (4, "tagged"), (8, "tagged")])
def test_row_sharded_fit_matches_single_gpu(world, exchange):
    """exchange: "tagged" = the self-validating 16-byte slots of finalize_cl_kernel (default), "flags" = LKM_P2P_LL=0, plain
    stores + system fence + flags (the protocol finalize_x_kernel uses for dense partials either way)."""
    have = _gpu_count()
    if have < world:
        pytest.skip(f"needs {world} GPUs, this box has {have}")
    port = 29500 + world + (os.getpid() % 200) + (300 if exchange == "flags" else 0)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    env = dict(os.environ, LKM_P2P_LL="0" if exchange == "flags" else "1")
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900, env=env)
    sys.stdout.write(res.stdout[-6000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    assert "MISMATCH" not in res.stdout
    assert res.stdout.count("[multi_gpu_check]") >= 8


def test_device_group_in_one_process(oracle):
    """lkm_create_multi: the sharded fit from ONE process (what the Rust binding's `fit(&dataset)` uses).  Two ranks on the
    devices the box has (the same device twice on a single-GPU box: the rendezvous, the replayed rng and the NVLink-style
    exchange all run, only the link is local)."""
    import numpy as np

    import linfa_b200 as lb

    have = _gpu_count()
    ids = [0, 1] if have >= 2 else [0, 0]
    g = np.random.default_rng(0)
    for dt, tol in ((np.float64, 1e-10), (np.float32, 1e-5)):
        centres = g.uniform(-30, 30, (12, 16))
        X = (centres[g.integers(0, 12, 40000)] + g.standard_normal((40000, 16))).astype(dt)
        grp = lb.DeviceGroup(ids)
        try:
            for init in (lb.KMeansInit.KMeansPlusPlus, lb.KMeansInit.Random, lb.KMeansInit.KMeansPara):
                params = lb.KMeans.params_with_rng(12, lb.Xoshiro256Plus.seed_from_u64(5)).n_runs(2).init_method(init)
                multi = grp.fit(params, lb.DatasetBase(X))
                single = params.fit(lb.DatasetBase(X))
                if init != lb.KMeansInit.KMeansPara:   # k-means|| draws per-shard streams: same quality, not the same seeds
                    np.testing.assert_allclose(multi.centroids(), single.centroids(), rtol=tol, atol=tol * 30)
                    assert multi.cluster_count().tolist() == single.cluster_count().tolist()
                    np.testing.assert_allclose(multi.inertia(), single.inertia(), rtol=tol)
                else:
                    assert multi.inertia() < 1.5 * single.inertia()
            lab, dist = grp.assign(single.centroids(), X)
            ref_lab, ref_dist = oracle.assign(X, single.centroids())
            assert (lab == ref_lab).all() and (dist == ref_dist).all()
        finally:
            grp.close()
    # wide rows: the pair kernels, rows_top2_kernel with every centroid in shared memory (132 KB), the seeding pass (203 KB) —
    # kernels whose shared-memory opt-in has to be made on EVERY device of the group, not once per process
    centres = g.uniform(-30, 30, (64, 128))
    X = (centres[g.integers(0, 64, 60000)] + g.standard_normal((60000, 128))).astype(np.float32)
    grp = lb.DeviceGroup(ids)
    try:
        for algo in (lb.KMeansAlgorithm.Lloyd, lb.KMeansAlgorithm.Hamerly):
            params = lb.KMeans.params_with_rng(256, lb.Xoshiro256Plus.seed_from_u64(6)).n_runs(1).max_n_iterations(8).algorithm(algo)
            multi = grp.fit(params, lb.DatasetBase(X))
            single = params.fit(lb.DatasetBase(X))
            np.testing.assert_allclose(multi.inertia(), single.inertia(), rtol=1e-4)
            assert multi.cluster_count().sum() == 60000
        lab, dist = grp.assign(single.centroids(), X)
        ref_lab, ref_dist = oracle.assign(X, single.centroids())
        assert (lab == ref_lab).all() and (dist == ref_dist).all()
    finally:
        grp.close()

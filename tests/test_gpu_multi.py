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


@pytest.mark.parametrize("world", [2, 4, 8])
def test_row_sharded_fit_matches_single_gpu(world):
    have = _gpu_count()
    if have < world:
        pytest.skip(f"needs {world} GPUs, this box has {have}")
    port = 29500 + world + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    res = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    sys.stdout.write(res.stdout[-6000:])
    assert res.returncode == 0, res.stdout[-4000:] + res.stderr[-4000:]
    assert "MISMATCH" not in res.stdout
    assert res.stdout.count("[multi_gpu_check]") >= 8

"""world_size-2 gloo tests of the sharded (multi-GPU) protocol, on CPU.

Each rank computes the partial vector of its row shard with the oracle, the partials are
all-gathered and combined in rank order (linfa_b200.distributed.combine_partials, the numpy
restatement of finalize_kernel) and the result must equal a single-process iteration."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from linfa_b200 import distributed as D
        from tests import oracle_lib

        orc = oracle_lib.load()
        # the unique-id exchange path (payload stands in for an ncclUniqueId)
        payload = bytes(range(128)) if rank == 0 else None
        got = D.broadcast_bytes(payload, 128, 0)
        assert got == bytes(range(128))

        g = np.random.default_rng(0)
        n, d, k = 10007, 5, 7
        centres = g.uniform(-30, 30, (k, d))
        X = np.concatenate([g.standard_normal((n // k + (1 if b < n % k else 0), d)) + centres[b] for b in range(k)])
        assert X.shape[0] == n
        cur = X[g.choice(n, k, replace=False)].copy()
        lo, hi = D.shard_range(n, rank, world)
        ranges = [D.shard_range(n, r, world) for r in range(world)]
        assert ranges[0][0] == 0 and ranges[-1][1] == n and all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
        shard = X[lo:hi]
        ref = cur.copy()
        for it in range(6):
            # rank-local partial: what lloyd_kernel + finalize_kernel<REDUCE_ONLY> publish
            lab, dists = orc.assign(shard, cur)
            sums = np.zeros((k, d))
            np.add.at(sums, lab.astype(np.int64), shard)
            counts = np.bincount(lab.astype(np.int64), minlength=k).astype(np.float64)
            part = torch.from_numpy(D.pack_partial(sums, counts, dists.sum()))
            gathered = [torch.zeros_like(part) for _ in range(world)]
            dist.all_gather(gathered, part)
            cur, cnt, inertia, shift, done = D.combine_partials([t.numpy() for t in gathered], cur, 1e-4, it, 100)
            # single-process reference iteration on the whole matrix
            lab_all, d_all = orc.assign(X, ref)
            ref = orc.compute_centroids(ref, X, lab_all, acc_f64=True)
            np.testing.assert_allclose(cur, ref, rtol=1e-12, atol=1e-12)
            assert cnt.tolist() == np.bincount(lab_all.astype(np.int64), minlength=k).tolist()
            np.testing.assert_allclose(inertia, d_all.sum(), rtol=1e-12)
        # every rank ends with bit-identical centroids (no broadcast needed)
        mine = torch.from_numpy(cur.copy())
        allc = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allc, mine)
        assert all(torch.equal(allc[0], c) for c in allc)
        ret[rank] = "ok"
    finally:
        dist.destroy_process_group()


def test_sharded_protocol_world2():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        ret = m.dict()
        mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
        assert dict(ret) == {0: "ok", 1: "ok"}


@pytest.mark.parametrize("n,world", [(10, 3), (1000003, 8), (7, 8), (0, 2)])
def test_shard_range_covers_rows(n, world):
    from linfa_b200.distributed import shard_range

    spans = [shard_range(n, r, world) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    sizes = [hi - lo for lo, hi in spans]
    assert max(sizes) - min(sizes) <= 1

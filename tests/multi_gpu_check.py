"""Multi-GPU parity check, run under torchrun on a box with >= 2 GPUs:

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
      tests/multi_gpu_check.py

Row-sharded Lloyd iterations (NCCL all-gather of the per-rank partials inside the library) must
give the same labels (counts), and centroids / inertia within tolerance, as one GPU holding the
whole matrix.  Every rank must end with bit-identical centroids."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import linfa_b200 as lb
    from linfa_b200 import distributed as D

    ok = True
    for dtype, n, d, k, rtol in [(np.float32, 400_003, 32, 64, 1e-5), (np.float64, 300_001, 8, 16, 1e-10),
                                 (np.float32, 100_000, 20, 12, 1e-5), (np.float32, 300_007, 128, 256, 1e-5),
                                 (np.float32, 200_000, 64, 600, 1e-5), (np.float32, 1, 32, 1, 1e-5)]:
        ctx = lb.Context(local)
        D.init_comm(ctx, device=torch.device("cuda", local))
        g = np.random.default_rng(5)
        centres = g.uniform(-30, 30, (k, d)).astype(dtype)
        init = (centres + 0.5 * g.standard_normal((k, d))).astype(dtype)
        lo, hi = D.shard_range(n, rank, world)
        ctx.gen_blobs(hi - lo, d, centres, 1.0, seed=11, row_offset=lo, n_global=n)
        ctx.set_keep_labels(True)
        c, counts, inertia, st = ctx.lloyd_iters(init, 6, 0.0)
        # per-row labels of the last iteration, gathered in rank order
        lab = torch.from_numpy(ctx.get_labels(hi - lo).astype(np.int64)).cuda()
        sizes = [D.shard_range(n, r, world)[1] - D.shard_range(n, r, world)[0] for r in range(world)]
        parts = [torch.zeros(s_, dtype=torch.int64, device="cuda") for s_ in sizes]
        pad = max(sizes)
        buf = [torch.zeros(pad, dtype=torch.int64, device="cuda") for _ in range(world)]
        mine = torch.zeros(pad, dtype=torch.int64, device="cuda")
        mine[: hi - lo] = lab
        dist.all_gather(buf, mine)
        all_labels = torch.cat([buf[r][: sizes[r]] for r in range(world)]).cpu().numpy()
        # bit-identical across ranks
        t = torch.from_numpy(c.copy()).cuda()
        allc = [torch.zeros_like(t) for _ in range(world)]
        dist.all_gather(allc, t)
        same = all(torch.equal(allc[0], x) for x in allc)
        if rank == 0:
            solo = lb.Context(local)
            solo.gen_blobs(n, d, centres, 1.0, seed=11)
            solo.set_keep_labels(True)
            c1, counts1, inertia1, st1 = solo.lloyd_iters(init, 6, 0.0)
            labels_equal = bool((solo.get_labels(n).astype(np.int64) == all_labels).all())
            scale = np.abs(c1).max()
            good = (same and labels_equal and st.n_iter == 6 and counts.tolist() == counts1.tolist()
                    and np.allclose(c, c1, rtol=rtol, atol=rtol * scale) and np.isclose(inertia, inertia1, rtol=rtol))
            print(f"[multi_gpu_check] {np.dtype(dtype).name} n={n} d={d} k={k} world={world}: "
                  f"{'ok' if good else 'MISMATCH'} path={st.assign_path} "
                  f"max|dC|/scale={np.abs(c - c1).max() / scale:.3e} counts_equal={counts.tolist() == counts1.tolist()} "
                  f"labels_equal_row_by_row={labels_equal} bit_identical_across_ranks={same}",
                  flush=True)
            ok = ok and good
            solo.close()
        ctx.close()
    # ---- sharded seeding: same rng => same picks as one GPU holding every row ----
    for dtype, n, d, k in [(np.float64, 50_001, 6, 9), (np.float32, 120_000, 32, 16)]:
        ctx = lb.Context(local)
        D.init_comm(ctx, device=torch.device("cuda", local))
        g = np.random.default_rng(9)
        centres = g.uniform(-30, 30, (k, d)).astype(dtype)
        lo, hi = D.shard_range(n, rank, world)
        ctx.gen_blobs(hi - lo, d, centres, 1.0, seed=21, row_offset=lo, n_global=n)
        solo = lb.Context(local)
        solo.gen_blobs(n, d, centres, 1.0, seed=21)
        res = {}
        for name, init, mode in [("random", lb.KMeansInit.Random, lb.PickMode.Exact),
                                 ("kmeans++ exact", lb.KMeansInit.KMeansPlusPlus, lb.PickMode.Exact),
                                 ("kmeans++ fast", lb.KMeansInit.KMeansPlusPlus, lb.PickMode.Fast),
                                 ("kmeans||", lb.KMeansInit.KMeansPara, lb.PickMode.Fast)]:
            p = (lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(40)).n_runs(1).max_n_iterations(10)
                 .init_method(init).pick_mode(mode).check())
            m_sh = p.fit_resident(ctx)
            m_1 = p.fit_resident(solo)
            t = torch.from_numpy(m_sh.centroids().copy()).cuda()
            allc = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(allc, t)
            same = all(torch.equal(allc[0], x) for x in allc)
            scale = np.abs(m_1.centroids()).max()
            if name in ("random", "kmeans++ exact"):
                good = same and np.allclose(m_sh.centroids(), m_1.centroids(), rtol=1e-5, atol=1e-5 * scale) \
                    and m_sh.cluster_count().tolist() == m_1.cluster_count().tolist()
            else:  # different (but legal) picks: compare quality
                good = same and float(m_sh.inertia()) <= 1.5 * float(m_1.inertia()) + 1e-9
            res[name] = good
            if rank == 0:
                print(f"[multi_gpu_check] seeding {np.dtype(dtype).name} n={n} k={k} {name}: {'ok' if good else 'MISMATCH'} "
                      f"inertia sharded={float(m_sh.inertia()):.6g} solo={float(m_1.inertia()):.6g}", flush=True)
            ok = ok and good
        ctx.close()
        solo.close()
    flag = torch.tensor([1 if ok else 0], device="cuda")
    dist.broadcast(flag, 0)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) == 1 else 1)


if __name__ == "__main__":
    main()

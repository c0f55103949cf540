"""Host-side plumbing of the row-sharded (multi-GPU) path: one process per GPU.

The data path has ONE exchange step per Lloyd iteration: every rank publishes its partial
vector [k*d sums | k counts | inertia] (f64) into an all-gather, and every rank then reduces the
gathered partials in RANK ORDER and applies the m_k-means update (finalize_kernel) — identical
inputs on every rank, so no broadcast of the centroids is needed.  The library runs that
all-gather itself over NCCL (lkm_comm_init); this module only (1) splits the rows, (2) ships the
NCCL unique id with whatever torch.distributed backend is up (nccl or gloo) and (3) restates the
combine step in numpy so the protocol can be checked on CPU (tests/test_distributed_cpu.py).
"""
import ctypes as C

import numpy as np


def shard_range(n, rank, world):
    """Contiguous row block of `rank`: [lo, hi).  The first n % world ranks get one extra row."""
    base, rem = divmod(int(n), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def partial_length(k, d):
    return k * d + k + 1


def pack_partial(sums, counts, inertia):
    """[k*d sums | k counts | inertia] as f64 — the layout finalize_kernel<REDUCE_ONLY> publishes."""
    k, d = sums.shape
    out = np.empty(partial_length(k, d), dtype=np.float64)
    out[: k * d] = np.asarray(sums, dtype=np.float64).ravel()
    out[k * d: k * d + k] = counts
    out[k * d + k] = inertia
    return out


def combine_partials(gathered, centroids, tol, n_iter, max_iter):
    """numpy restatement of finalize_kernel<UPDATE> over rank partials: rank-order f64 sums,
    new = (sum + old) / (count + 1) (KM/algorithm.rs:785-810), shift = ||old - new||_F evaluated
    on the F-rounded centroids, stop rule of KM/algorithm.rs:328."""
    centroids = np.asarray(centroids)
    k, d = centroids.shape
    tot = np.zeros(partial_length(k, d), dtype=np.float64)
    for part in gathered:  # ascending rank
        tot += part
    sums = tot[: k * d].reshape(k, d)
    counts = tot[k * d: k * d + k]
    new = ((sums + centroids.astype(np.float64)) / (counts[:, None] + 1.0)).astype(centroids.dtype)
    diff = (centroids - new).astype(np.float64)
    shift = centroids.dtype.type(np.sqrt((diff * diff).sum()))
    done = bool(shift < centroids.dtype.type(tol)) or (n_iter + 1 >= max_iter)
    return new, counts, tot[k * d + k], float(shift), done


def broadcast_bytes(payload, nbytes, src=0, device=None):
    """Ships `nbytes` bytes from rank `src` to every rank over torch.distributed (any backend)."""
    import torch
    import torch.distributed as dist

    buf = torch.zeros(nbytes, dtype=torch.uint8)
    if dist.get_rank() == src:
        buf = torch.frombuffer(bytearray(payload), dtype=torch.uint8).clone()
    if device is not None:
        buf = buf.to(device)
    dist.broadcast(buf, src)
    return bytes(buf.cpu().numpy().tobytes())


def init_comm(ctx, device=None):
    """Creates the NCCL communicator of `ctx` (lkm_comm_init) for the current process group."""
    import torch.distributed as dist

    from . import _ffi

    rank, world = dist.get_rank(), dist.get_world_size()
    payload = None
    if rank == 0:
        buf = C.create_string_buffer(128)
        st = _ffi.lib.lkm_comm_unique_id(buf)
        if st != 0:
            raise RuntimeError("lkm_comm_unique_id failed (libnccl not loadable?)")
        payload = buf.raw
    uid = broadcast_bytes(payload, 128, 0, device)
    ctx.comm_init(uid, rank, world)
    return rank, world

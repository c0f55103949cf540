#!/usr/bin/env python
"""bench.py — K-Means Lloyd samples*iterations/s on B200 (BASELINE.json metric).

Workload (N = 1): BASELINE.json configs[1] — KMeans k=64 on 1M x 32 f32 Gaussian
blobs (linfa_datasets::generate::blobs layout, generated on the device).  A "step"
is ONE Lloyd iteration (assign every sample to its nearest centroid, recompute
centroids and inertia) over the whole matrix.  N > 1: one process per GPU, each
rank holds its own 1M x 32 shard (weak scaling), per-iteration partials are
all-gathered over NCCL and reduced in rank order.

  value      device-timed (CUDA events on the library's stream), observations resident
             in HBM, L2 overwritten between steps (outside the timed spans)
  e2e        the same metric through the public API with HOST buffers: every step copies
             the observations from pinned host memory, runs one iteration and reads the
             centroids/inertia back
  roofline   fused kernel: n*d*sizeof(F) algorithmic bytes per launch / its mean duration,
             against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline / --impl reference: the C++ restatement of linfa's rayon/ndarray K-Means
             (oracle/), all host threads, on a bounded row sample of the same workload
"""
import argparse
import ctypes as C
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (rows per GPU, d, k, dtype)
    "cfg2": (1_000_000, 32, 64, np.float32),
    "cfg3": (10_000_000, 128, 256, np.float32),   # 10M x 128 total at N=1
    "cfg4": (12_500_000, 64, 1024, np.float32),   # 100M x 64 over 8 GPUs
    "cfg5": (125_000_000, 8, 16, np.float64),     # 1B x 8 over 8 GPUs
    "cfg1": (10_000, 2, 3, np.float64),
}
WORKLOAD_TEXT = {
    "cfg2": "KMeans k=64 on 1M x 32 f32 blobs per GPU (BASELINE configs[1]); step = 1 Lloyd iteration",
    "cfg3": "KMeans k=256 on 10M x 128 f32 blobs per GPU (BASELINE configs[2]); step = 1 Lloyd iteration",
    "cfg4": "KMeans k=1024 on 12.5M x 64 f32 blobs per GPU (BASELINE configs[3] shard); step = 1 Lloyd iteration",
    "cfg5": "KMeans k=16 on 125M x 8 f64 blobs per GPU (BASELINE configs[4] shard); step = 1 Lloyd iteration",
    "cfg1": "KMeans k=3 on 10k x 2 f64 blobs (BASELINE configs[0]); step = 1 Lloyd iteration",
}


DEFAULT_KERNELS = {
    "cfg2": "lloyd_tc4_kernel (fused assign+update, tcgen05 pipeline)",
    "cfg3": "assign_tc5s_kernel + update_rows_kernel (CTA-pair tcgen05 assignment, then the streaming update: two passes over X)",
    "cfg4": "4 x assign_tc5s_kernel (one per 256 centroids) + update_rows_kernel x 2 cluster chunks (six passes over X)",
    "cfg5": "lloyd_f64s_kernel<8, 1, 2> (f32-screened fused assign+update, f64 d=8 k<=64)",
    "cfg1": "lloyd_kernel<double> (exact, CUDA cores)",
}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return float(p["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def load_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (None if not captured)."""
    try:
        with open(os.path.join(ROOT, "profiles", "r1_traffic.json")) as f:
            t = json.load(f).get(workload)
        return (int(t["dram_bytes_read"]) + int(t["dram_bytes_write"]), t["kernel"]) if t else (None, None)
    except Exception:
        return None, None


def centres_for(k, d, dtype, seed=1234):
    g = np.random.default_rng(seed)
    return g.uniform(-30.0, 30.0, (k, d)).astype(dtype)  # benches/k_means.rs:48-49: Uniform(-30, 30)


def start_centroids(centres, seed=99):
    # fixed starting point for the timed Lloyd steps: blob centres + N(0, 0.5) noise
    g = np.random.default_rng(seed)
    return (centres.astype(np.float64) + 0.5 * g.standard_normal(centres.shape)).astype(centres.dtype)


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons through NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.stop_flag, self.max_mhz = index, [], set(), False, None
        self.ok = False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            uuid = None
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = index
            if vis:
                ent = vis.split(",")[index].strip()
                if ent.isdigit():
                    idx = int(ent)
                else:
                    uuid = ent
            self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid) if uuid else pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception as e:  # pragma: no cover
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            nv.nvmlClocksThrottleReasonHwSlowdown: "hw_slowdown",
            nv.nvmlClocksThrottleReasonHwThermalSlowdown: "hw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwThermalSlowdown: "sw_thermal_slowdown",
            nv.nvmlClocksThrottleReasonSwPowerCap: "sw_power_cap",
            nv.nvmlClocksThrottleReasonHwPowerBrakeSlowdown: "hw_power_brake",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, nm in names.items():
                    if r & bit:
                        self.reasons.add(nm)
            except Exception:
                pass
            time.sleep(0.005)

    def result(self):
        self.stop_flag = True
        if self.ok:
            self.join(timeout=1.0)
        med = float(np.median(self.samples)) if self.samples else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def time_oracle(orc, oracle_lib, n, d, k, dtype, steps, warmup, budget_s=15.0):
    """Times Lloyd iterations of the CPU oracle (all host threads) on a bounded row sample."""
    centres = centres_for(k, d, dtype)
    g = np.random.default_rng(7)
    probe_rows = min(n, 32768)

    def make(rows):
        m = max(1, rows // k)
        lab = np.minimum(np.arange(rows) // m, k - 1)
        return (centres[lab].astype(np.float64) + g.standard_normal((rows, d))).astype(dtype)

    c0 = start_centroids(centres)

    def iterate(x, c, iters):
        # one Lloyd iteration of the reference (KM/algorithm.rs:315-329): assignment on all host
        # threads, single-threaded centroid update, Frobenius shift, dists.sum()
        for _ in range(iters):
            labels, dists = orc.assign(x, c)
            cn = orc.compute_centroids(c, x, labels)
            orc.l2_distance(c, cn)
            orc.sum(dists)
            c = cn
        return c

    xp = make(probe_rows)
    t0 = time.perf_counter()
    iterate(xp, c0, 1)
    per_row = (time.perf_counter() - t0) / probe_rows
    rows = int(min(n, max(probe_rows, budget_s / max(per_row, 1e-12) / max(steps + warmup, 1))))
    x = xp if rows == probe_rows else make(rows)
    c = iterate(x, c0, warmup)
    t0 = time.perf_counter()
    iterate(x, c, steps)
    dt = time.perf_counter() - t0
    return rows * steps / dt, rows, dt, orc.threads


def run_reference(args, rank, world):
    if rank != 0:
        return
    from tests import oracle_lib

    orc = oracle_lib.load()
    n, d, k, dtype = WORKLOADS[args.workload]
    value, rows, dt, threads = time_oracle(orc, oracle_lib, n, d, k, dtype, args.steps, args.warmup, budget_s=60.0)
    sample = f"{rows} of {n} rows of the same blobs (k={k}, d={d}), {args.steps} Lloyd iterations, {threads} threads"
    line = {
        "impl": "reference", "metric": "kmeans_lloyd_samples_iter_per_s", "value": value, "unit": "samples*iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD_TEXT[args.workload], "k": k, "d": d, "rows_timed": rows,
                   "what": "C++ restatement of linfa's rayon/ndarray K-Means (oracle/); linfa itself needs cargo, absent here"},
        "cpu_baseline": {"value": value, "unit": "samples*iter/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples*iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--assign-path", type=int, default=-1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-init", action="store_true", help="skip the separately timed k-means++ seeding")
    ap.add_argument("--no-flush", action="store_true", help="keep L2 warm between steps (not the headline)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch

    import linfa_b200 as lb
    from linfa_b200 import _ffi

    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    ctx = lb.Context(local_rank)
    if world > 1:
        from linfa_b200 import distributed as D

        D.init_comm(ctx, device=torch.device("cuda", local_rank))

    n, d, k, dtype = WORKLOADS[args.workload]
    itemsize = np.dtype(dtype).itemsize
    centres = centres_for(k, d, dtype)
    c0 = start_centroids(centres)
    tiny = 0.0  # lkm_lloyd_iters: tolerance 0 = run exactly `steps` iterations
    if args.assign_path >= 0:
        ctx.set_assign_path(args.assign_path)
    ctx.gen_blobs(n, d, centres, 1.0, seed=2024, row_offset=rank * n, n_global=world * n)

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def allmax(v):
        if dist is None:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- seeding, timed separately (init is not part of the Lloyd metric) ----
    init_ms = None
    if world == 1 and n <= 20_000_000 and not args.no_init:
        p = lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(40)).n_runs(1).pick_mode(lb.PickMode.Fast).check()
        pp, _keep = p._native_params(dtype)
        out = np.zeros((k, d), dtype=dtype)
        sfx, ct = _ffi.sfx_of(dtype)
        t0 = time.perf_counter()
        st = getattr(_ffi.lib, f"lkm_init_{sfx}")(ctx.h, C.byref(pp), _ffi.make_rng(lb.Xoshiro256Plus.seed_from_u64(40)),
                                                  _ffi.fptr(out, ct))
        ctx.check(st)
        init_ms = 1e3 * (time.perf_counter() - t0)

    # ---- device-resident leg ----
    flush_bytes = 0 if args.no_flush else 512 * 1024 * 1024
    ctx.set_l2_flush(flush_bytes)
    ctx.lloyd_iters(c0, args.warmup, tiny)  # W untimed warm-up steps
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    t0 = time.perf_counter()
    cen, counts, inertia, st = ctx.lloyd_iters(c0, args.steps, tiny)
    barrier()
    wall_ms = 1e3 * (time.perf_counter() - t0)
    assert st.n_iter == args.steps, (st.n_iter, args.steps)
    dev_ms = allmax(float(st.device_ms))
    launches = int(st.kernel_launches)
    # the dominant kernel on its own (extra event per step; not the run `value` comes from)
    main_ms = None
    if flush_bytes:
        ctx.set_kernel_timing(True)
        ctx.lloyd_iters(c0, args.warmup, tiny)
        barrier()
        _, _, _, st_k = ctx.lloyd_iters(c0, args.steps, tiny)
        main_ms = allmax(float(st_k.main_kernel_ms))
        ctx.set_kernel_timing(False)
    # keep the same load running ~1.5 s so that NVML sees the clocks of this kernel mix
    t_end = time.perf_counter() + 1.5
    while time.perf_counter() < t_end:
        ctx.lloyd_iters(c0, args.steps, tiny)
    clocks = sampler.result()
    value = world * n * args.steps / (dev_ms * 1e-3)

    # L2-warm figure (what a real fit sees after the first iteration), for the record
    ctx.set_l2_flush(0)
    ctx.lloyd_iters(c0, args.warmup, tiny)
    barrier()
    _, _, _, st_warm = ctx.lloyd_iters(c0, args.steps, tiny)
    warm_ms = allmax(float(st_warm.device_ms))

    # ---- roofline of the dominant (fused assign+update) kernel ----
    peak, peak_src = load_peaks()
    roofline = None
    if main_ms:
        per_launch_s = main_ms * 1e-3 / args.steps
        achieved = n * d * itemsize / per_launch_s / 1e9
        traffic, kname = load_traffic(args.workload)
        roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": traffic, "kernel": kname or DEFAULT_KERNELS.get(args.workload, "fused assign+update kernel"),
                    "algorithmic_bytes_per_launch": n * d * itemsize, "mean_launch_us": per_launch_s * 1e6,
                    "peak_source": peak_src}

    # ---- end-to-end leg: host buffers through the public API, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        e2e_rows = n
        host = torch.empty((e2e_rows, d), dtype=torch.float32 if dtype == np.float32 else torch.float64,
                           pin_memory=True)
        hx = host.numpy()
        hx[...] = ctx.get_data(0, e2e_rows)
        e2e_steps = max(3, min(args.steps, 10))
        cur = c0.copy()
        for _ in range(2):
            ctx.set_data(hx)
            ctx.lloyd_iters(cur, 1, tiny)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.set_data(hx)                                   # H2D of the step's observations (pinned)
            cur, cnts, inert, _s = ctx.lloyd_iters(cur, 1, tiny)  # H2D centroids, 1 iteration, D2H result
        barrier()
        e2e_s = allmax(time.perf_counter() - t0)
        e2e = {"value": world * e2e_rows * e2e_steps / e2e_s, "unit": "samples*iter/s",
               "h2d_bytes_per_step": int(e2e_rows * d * itemsize + k * d * itemsize),
               "d2h_bytes_per_step": int(k * d * itemsize + k * 8 + 64), "steps": e2e_steps,
               "ms_per_step": 1e3 * e2e_s / e2e_steps}

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from tests import oracle_lib

        orc = oracle_lib.load()
        v, rows, dt, threads = time_oracle(orc, oracle_lib, n, d, k, dtype, 3, 1, budget_s=15.0)
        cpu = {"value": v, "unit": "samples*iter/s", "cores": threads, "kind": "port",
               "sample": f"{rows} of {n} rows of the same blob distribution, 3 Lloyd iterations (+1 warm-up), "
                         f"C++ restatement of linfa (oracle/), {threads} threads"}

    if rank == 0:
        line = {
            "metric": "kmeans_lloyd_samples_iter_per_s", "value": value, "unit": "samples*iter/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD_TEXT[args.workload], "rows_per_gpu": n, "d": d, "k": k,
                       "init": "blob centres + N(0,0.5) (Precomputed); k-means++ timed separately",
                       "kmeanspp_fast_init_ms": init_ms,
                       "l2": ("512 MiB scratch overwritten between steps, then a 256 MiB read-only sweep so that no dirty lines remain; both outside the timed spans" if flush_bytes
                              else "no flush (L2 warm)"),
                       "l2_warm_ms_per_step": warm_ms / args.steps,
                       "wall_ms_per_step_incl_flush": wall_ms / args.steps,
                       "parallelism": f"row-sharded x{world}, all-gather of per-rank partials" if world > 1 else "1 GPU"},
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()

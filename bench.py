#!/usr/bin/env python
"""bench.py — K-Means Lloyd samples*iterations/s on B200 (BASELINE.json metric).

Headline workload: BASELINE.json configs[2], the config the metric is quoted on ("1/2/4/8 B200") —
KMeans k=256 on 10M x 128 f32 Gaussian blobs (linfa_datasets::generate::blobs layout, generated on the
device).  A "step" is ONE Lloyd iteration (assign every sample to its nearest centroid, recompute
centroids and inertia) over the whole matrix.  `--gpus N`: one process per GPU, the SAME 10M rows split
into N contiguous row shards (strong scaling); every iteration the per-rank (k*d sums, k counts, inertia)
partials are exchanged over NVLink and combined in rank order on every rank.

  value      device-timed (CUDA events on the library's stream), observations resident in HBM (5.12 GB,
             40x the L2; the smaller side workloads overwrite the L2 between steps, outside the timed spans)
  e2e        the same metric through the C ABI with HOST buffers: every step copies the observations
             from pinned host memory, runs one iteration and reads centroids / counts / inertia back
  roofline   dominant kernel: algorithmic bytes (n*d*sizeof(F)) and flops (2*n*k*d) per launch / its mean
             duration (CUDA events around the launch) against MEASURED_PEAKS.json; the bound is the
             roofline with the larger minimum time
  cpu_baseline / --impl reference: the C++ restatement of linfa's rayon/ndarray K-Means (oracle/),
             all host threads, on a bounded row sample of the same workload
  roofline.other_workloads: the other BASELINE configs in the same run and JSON line — cfg2 (1M x 32, k=64),
             cfg4 shard (12.5M x 64, k=1024 per GPU = the 100M config at N=8), cfg5 shard (125M x 8 f64, k=16
             per GPU = the 1B config at N=8), each weak-scaled (fixed rows per GPU), plus the headline shape
             with hashed (unordered) rows, from a k-means++ start, under KMeansAlgorithm::Hamerly, and an f64
             shape on the FP64 tensor pipe (DMMA)
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
    # name: (rows, d, k, dtype, scaling) — "strong": rows are the TOTAL, split over the GPUs; "weak": rows per GPU
    "cfg3": (10_000_000, 128, 256, np.float32, "strong"),
    "cfg2": (1_000_000, 32, 64, np.float32, "weak"),
    "cfg4": (12_500_000, 64, 1024, np.float32, "weak"),   # 100M x 64 over 8 GPUs
    "cfg5": (125_000_000, 8, 16, np.float64, "weak"),     # 1B x 8 over 8 GPUs
    "cfg1": (10_000, 2, 3, np.float64, "weak"),
}
WORKLOAD_TEXT = {
    "cfg3": "KMeans k=256 on 10M x 128 f32 blobs, rows split over the GPUs (BASELINE configs[2]); step = 1 Lloyd iteration",
    "cfg2": "KMeans k=64 on 1M x 32 f32 blobs per GPU (BASELINE configs[1]); step = 1 Lloyd iteration",
    "cfg4": "KMeans k=1024 on 12.5M x 64 f32 blobs per GPU (BASELINE configs[3] = 100M x 64 at 8 GPUs); step = 1 Lloyd iteration",
    "cfg5": "KMeans k=16 on 125M x 8 f64 blobs per GPU (BASELINE configs[4] = 1B x 8 at 8 GPUs); step = 1 Lloyd iteration",
    "cfg1": "KMeans k=3 on 10k x 2 f64 blobs (BASELINE configs[0]); step = 1 Lloyd iteration",
}
KERNELS = {
    "cfg3": "assign_tc5s_kernel<128,true,256> (CTA-pair tcgen05 screening assignment with the fused update; undecided rows through rows_top2_kernel, listed tiles through update_rows_kernel<128>, finalize_cl_kernel)",
    "cfg2": "lloyd_tc4_kernel<64,1> (fused assign+update, warp-specialised tcgen05 pipeline)",
    "cfg4": "4 x assign_tc5s_kernel<64,*,256> (one launch per 256 centroids, the last one with the fused update)",
    "cfg5": "lloyd_f64s_kernel<8,true,2> (f32-screened fused assign+update for f64 rows)",
    "cfg1": "lloyd_kernel<double,2,FUSED> (exact, CUDA cores)",
}
L2_TEXT = {
    True: "512 MiB scratch overwritten between steps, then a 256 MiB read-only sweep so that no dirty lines remain; both outside the timed spans",
    False: "inputs larger than L2 (no flush): X is 40x the 126 MB L2",
}


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return {"hbm": float(p["hbm_gbs"]), "bf16_burst": float(p["bf16_tflops"]),
                "bf16_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])),
                "source": "measured (MEASURED_PEAKS.json)"}
    except Exception:
        return {"hbm": 6650.0, "bf16_burst": 1650.0, "bf16_sustained": 1650.0, "source": "fallback (B200_PROFILING.md)"}


def load_traffic(workload):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture (None if not captured)."""
    for name in ("r2c_traffic.json", "r2_traffic.json", "r1_traffic.json"):
        try:
            with open(os.path.join(ROOT, "profiles", name)) as f:
                t = json.load(f).get(workload)
            if t:
                return int(t["dram_bytes_read"]) + int(t["dram_bytes_write"]), f"profiles/{name}: {t.get('kernel', '')} ({t.get('capture', 'ncu --set full')})"
        except Exception:
            pass
    return None, None


def centres_for(k, d, dtype, seed=1234):
    g = np.random.default_rng(seed)
    return g.uniform(-30.0, 30.0, (k, d)).astype(dtype)  # benches/k_means.rs:48-49: Uniform(-30, 30)


def start_centroids(centres, seed=99):
    # fixed starting point for the timed Lloyd steps: blob centres + N(0, 0.5) noise
    g = np.random.default_rng(seed)
    return (centres.astype(np.float64) + 0.5 * g.standard_normal(centres.shape)).astype(centres.dtype)


def shard_of(workload, rank, world):
    rows, d, k, dtype, scaling = WORKLOADS[workload]
    if scaling == "strong":
        lo = rows * rank // world
        hi = rows * (rank + 1) // world
        return hi - lo, lo, rows
    return rows, rank * rows, world * rows


def config_of(workload, world, flush):
    """The workload description both arms print (identical keys and values)."""
    rows, d, k, dtype, scaling = WORKLOADS[workload]
    n_total = rows if scaling == "strong" else rows * world
    return {"workload": WORKLOAD_TEXT[workload], "rows_total": n_total, "d": d, "k": k, "scaling": scaling,
            "init": "blob centres + N(0,0.5) (Precomputed); seeding timed separately",
            "l2": L2_TEXT[bool(flush)]}


def needs_flush(workload, world):
    n, _, _ = shard_of(workload, 0, world)
    _, d, _, dtype, _ = WORKLOADS[workload]
    return n * d * np.dtype(dtype).itemsize < 4 * 126 * 1024 * 1024


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


def time_oracle(orc, n, d, k, dtype, steps, warmup, budget_s=15.0):
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
    rows, d, k, dtype, scaling = WORKLOADS[args.workload]
    n_total = rows if scaling == "strong" else rows * args.gpus
    value, used, dt, threads = time_oracle(orc, n_total, d, k, dtype, args.steps, args.warmup, budget_s=60.0)
    sample = f"{used} of {n_total} rows of the same blob distribution (k={k}, d={d}), {args.steps} Lloyd iterations, {threads} threads"
    line = {
        "impl": "reference", "metric": "kmeans_lloyd_samples_iter_per_s", "value": value, "unit": "samples*iter/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None,
        "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
        "config": config_of(args.workload, args.gpus, needs_flush(args.workload, args.gpus)),
        "detail": {"rows_timed": used,
                   "what": "C++ restatement of linfa's rayon/ndarray K-Means (oracle/), assignment on all host threads, serial update; "
                           "linfa itself needs cargo, absent here"},
        "cpu_baseline": {"value": value, "unit": "samples*iter/s", "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "samples*iter/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


class Runner:
    def __init__(self, args, rank, world, local_rank):
        import torch

        import linfa_b200 as lb

        self.torch, self.lb = torch, lb
        self.args, self.rank, self.world, self.local_rank = args, rank, world, local_rank
        torch.cuda.set_device(local_rank)
        self.dist = None
        if world > 1:
            import torch.distributed as dist

            dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
            self.dist = dist
        self.ctx = lb.Context(local_rank)
        if world > 1:
            from linfa_b200 import distributed as D

            D.init_comm(self.ctx, device=torch.device("cuda", local_rank))
        if args.assign_path >= 0:
            self.ctx.set_assign_path(args.assign_path)
        self.peaks = load_peaks()

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist is not None:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def allmax(self, v):
        if self.dist is None:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def generate(self, workload, shuffled=False):
        rows, d, k, dtype, scaling = WORKLOADS[workload]
        n, off, n_total = shard_of(workload, self.rank, self.world)
        centres = centres_for(k, d, dtype)
        self.ctx.set_blob_layout(shuffled)
        self.ctx.gen_blobs(n, d, centres, 1.0, seed=2024, row_offset=off, n_global=n_total)
        self.ctx.set_blob_layout(False)
        return n, n_total, d, k, dtype, centres

    def timed_steps(self, c0, steps, warmup, flush, kernel_timing=True, hold_s=1.0):
        """W warm-up + K timed Lloyd steps; returns device ms (max over ranks), kernel ms, stats, clocks."""
        ctx = self.ctx
        ctx.set_l2_flush(512 * 1024 * 1024 if flush else 0)
        ctx.set_kernel_timing(False)
        ctx.lloyd_iters(c0, warmup, 0.0)
        sampler = ClockSampler(self.local_rank)
        sampler.start()
        self.barrier()
        t0 = time.perf_counter()
        _, _, _, st = ctx.lloyd_iters(c0, steps, 0.0)
        self.barrier()
        wall_ms = 1e3 * (time.perf_counter() - t0)
        assert st.n_iter == steps, (st.n_iter, steps)
        dev_ms = self.allmax(float(st.device_ms))
        launches, fallback = int(st.kernel_launches), int(st.fallback_rows & ((1 << 40) - 1))
        main_ms = None
        if kernel_timing:
            # the dominant kernel on its own (extra event per step; not the run `value` comes from).  The per-step event
            # mode needs the flush path, so shapes larger than L2 use a 1-byte "flush".
            ctx.set_l2_flush(512 * 1024 * 1024 if flush else 64)
            ctx.set_kernel_timing(True)
            ctx.lloyd_iters(c0, warmup, 0.0)
            self.barrier()
            _, _, _, st_k = ctx.lloyd_iters(c0, steps, 0.0)
            main_ms = self.allmax(float(st_k.main_kernel_ms))
            ctx.set_kernel_timing(False)
            ctx.set_l2_flush(512 * 1024 * 1024 if flush else 0)
        # keep the same load running so that NVML sees the clocks of this kernel mix
        t_end = time.perf_counter() + hold_s
        while time.perf_counter() < t_end:
            ctx.lloyd_iters(c0, steps, 0.0)
        clocks = sampler.result()
        ctx.set_l2_flush(0)
        return dev_ms, main_ms, launches, fallback, wall_ms, clocks

    def roofline(self, workload, n, d, k, dtype, main_ms, steps):
        if not main_ms:
            return None
        itemsize = np.dtype(dtype).itemsize
        per_launch_s = main_ms * 1e-3 / steps
        gbs = n * d * itemsize / per_launch_s / 1e9
        tfl = 2.0 * n * k * d / per_launch_s / 1e12
        pk = self.peaks
        hbm = {"achieved": gbs, "peak": pk["hbm"], "unit": "GB/s", "frac": gbs / pk["hbm"]}
        out = {"kernel": KERNELS.get(workload), "algorithmic_bytes_per_launch": n * d * itemsize,
               "algorithmic_flops_per_launch": 2 * n * k * d, "mean_launch_us": per_launch_s * 1e6,
               "peak_source": pk["source"], "hbm": hbm}
        if dtype == np.float32 and d >= 32:
            # TF32 dense peak = half of the measured bf16 rate; the sustained figure, because the kernel is timed inside a
            # long loop (the clocks line shows the power cap it runs under)
            tpk = 0.5 * pk["bf16_sustained"]
            out["tensor"] = {"achieved": tfl, "peak": tpk, "unit": "TFLOP/s", "frac": tfl / tpk,
                             "peak_burst": 0.5 * pk["bf16_burst"], "frac_of_burst": tfl / (0.5 * pk["bf16_burst"]),
                             "what": "2*n*k*d flops of ONE TF32 product (the screening pass the kernel executes) against half the measured bf16 rate"}
        # the binding roofline is the one with the larger minimum time
        t_hbm = n * d * itemsize / (pk["hbm"] * 1e9)
        t_tc = (2.0 * n * k * d / (out["tensor"]["peak"] * 1e12)) if "tensor" in out else 0.0
        b = out["tensor"] if t_tc > t_hbm else hbm
        out.update({"bound": "tensor" if t_tc > t_hbm else "hbm", "achieved": b["achieved"], "peak": b["peak"],
                    "unit": b["unit"], "frac": b["frac"]})
        traffic, src = load_traffic(workload)
        out["traffic"] = traffic
        out["traffic_source"] = src
        return out

    def side(self, workload, steps, warmup, shuffled=False, kmeanspp_start=False):
        """One side workload: value, kernel time and roofline fractions, clocks."""
        n, n_total, d, k, dtype, centres = self.generate(workload, shuffled)
        c0 = start_centroids(centres)
        extra = {}
        if kmeanspp_start:
            lb = self.lb
            p = lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(40)).n_runs(1).pick_mode(lb.PickMode.Fast).check()
            t0 = time.perf_counter()
            c0 = p.init_centroids_resident(self.ctx)
            extra["kmeanspp_fast_init_ms"] = 1e3 * (time.perf_counter() - t0)
        flush = needs_flush(workload, self.world)
        dev_ms, main_ms, launches, fallback, _, clocks = self.timed_steps(c0, steps, warmup, flush, hold_s=0.5)
        rf = self.roofline(workload, n, d, k, dtype, main_ms, steps)
        out = {"workload": WORKLOAD_TEXT[workload], "rows_per_gpu": n, "rows_total": n_total, "d": d, "k": k,
               "dtype": "f32" if dtype == np.float32 else "f64",
               "rows": "hashed blob per row (unordered)" if shuffled else "blob-ordered (generate::blobs)",
               "start": "k-means++ (fast pick) on the resident rows" if kmeanspp_start else "blob centres + N(0,0.5)",
               "value": n_total * steps / (dev_ms * 1e-3), "unit": "samples*iter/s", "ms_per_step": dev_ms / steps,
               "steps": steps, "warmup": warmup, "l2": L2_TEXT[flush], "gpu_launches": launches,
               "rows_rescanned_exactly": fallback,
               "kernel_us": (main_ms * 1e3 / steps) if main_ms else None,
               "hbm_frac": rf["hbm"]["frac"] if rf else None,
               "tensor_frac": rf["tensor"]["frac"] if rf and "tensor" in rf else None,
               "bound": rf["bound"] if rf else None, "clocks": clocks}
        out.update(extra)
        return out


def side_hamerly(R, steps):
    """Hamerly (KM/algorithm.rs:248-291) on the headline shape from a k-means++ start, next to Lloyd from the same start:
    device time per iteration of both, and how many rows the bounds let the iterations skip."""
    lb = R.lb
    n, n_total, d, k, dtype, centres = R.generate("cfg3")
    p = lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(40)).n_runs(1).pick_mode(lb.PickMode.Fast).check()
    c0 = p.init_centroids_resident(R.ctx)
    iters = max(steps, 20)
    R.ctx.hamerly_iters(c0, 3, 0.0)
    _, _, _, st = R.ctx.hamerly_iters(c0, iters, 0.0)
    _, _, _, sl = R.ctx.lloyd_iters(c0, iters, 0.0)
    return {"workload": WORKLOAD_TEXT["cfg3"] + " — KMeansAlgorithm::Hamerly, k-means++ (fast pick) start", "rows_per_gpu": n, "d": d, "k": k,
            "iterations": int(st.n_iter), "value": n * int(st.n_iter) / (float(st.device_ms) * 1e-3), "unit": "samples*iter/s",
            "ms_per_step": float(st.device_ms) / int(st.n_iter), "lloyd_ms_per_step_same_start": float(sl.device_ms) / int(sl.n_iter),
            "rows_scanned_per_iteration": int(st.rows_scanned) / max(int(st.n_iter) - 1, 1),
            "rows_tightened_per_iteration": int(st.rows_tightened) / max(int(st.n_iter) - 1, 1), "gpu_launches": int(st.kernel_launches)}


def side_predict(R, steps):
    """PredictInplace + Transformer (KM/algorithm.rs:718-777) on the resident headline matrix: labels and min squared distances of
    every row, read back to the host (wall clock: the call a user makes)."""
    n, n_total, d, k, dtype, centres = R.generate("cfg3")
    c = start_centroids(centres)
    R.ctx.assign(c, None)
    t0 = time.perf_counter()
    reps = 3
    for _ in range(reps):
        labels, dists = R.ctx.assign(c, None)
    dt = (time.perf_counter() - t0) / reps
    return {"workload": WORKLOAD_TEXT["cfg3"] + " — update_memberships_and_dists on the resident rows, labels (u64) and distances read back",
            "rows_per_gpu": n, "d": d, "k": k, "value": n / dt, "unit": "samples/s", "ms_per_call": 1e3 * dt,
            "d2h_bytes_per_call": int(n * (8 + np.dtype(dtype).itemsize))}


def side_dmma(R, steps):
    """f64 with a large contraction (north_star: DMMA for f64): 2M x 64 f64, k = 256, Lloyd iterations on the FP64 tensor pipe."""
    n, d, k = 2_000_000, 64, 256
    centres = centres_for(k, d, np.float64)
    R.ctx.gen_blobs(n, d, centres, 1.0, seed=2024)
    c0 = start_centroids(centres)
    R.ctx.lloyd_iters(c0, 3, 0.0)
    _, _, _, st = R.ctx.lloyd_iters(c0, max(steps // 2, 5), 0.0)
    it = int(st.n_iter)
    ms = float(st.device_ms) / it
    return {"workload": "KMeans k=256 on 2M x 64 f64 blobs (f64, large k*d): assign_dmma_kernel + update", "rows_per_gpu": n, "d": d, "k": k,
            "dtype": "f64", "value": n * it / (float(st.device_ms) * 1e-3), "unit": "samples*iter/s", "ms_per_step": ms,
            "assign_path": int(st.assign_path), "fp64_tflops_2nkd": 2.0 * n * k * d / (ms * 1e-3) / 1e12, "gpu_launches": int(st.kernel_launches)}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg3", choices=sorted(WORKLOADS))
    ap.add_argument("--assign-path", type=int, default=-1)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-init", action="store_true", help="skip the separately timed seeding")
    ap.add_argument("--no-others", action="store_true", help="skip roofline.other_workloads")
    ap.add_argument("--shuffled", action="store_true", help="headline workload with hashed (unordered) rows")
    ap.add_argument("--rows", type=int, default=0, help="override the workload's row count (tuning aid, not the headline)")
    args = ap.parse_args()
    if args.rows:
        w = WORKLOADS[args.workload]
        WORKLOADS[args.workload] = (args.rows,) + w[1:]
        WORKLOAD_TEXT[args.workload] += f" [rows overridden: {args.rows}]"
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    R = Runner(args, rank, world, local_rank)
    ctx, lb, torch = R.ctx, R.lb, R.torch
    from linfa_b200 import _ffi

    n, n_total, d, k, dtype, centres = R.generate(args.workload, args.shuffled)
    itemsize = np.dtype(dtype).itemsize
    c0 = start_centroids(centres)
    flush = needs_flush(args.workload, world)

    # ---- seeding, timed separately (init is not part of the Lloyd metric) ----
    init = {}
    if world == 1 and not args.no_init:
        for name, method in (("kmeanspp_fast_init_ms", lb.KMeansInit.KMeansPlusPlus), ("kmeans_para_init_ms", lb.KMeansInit.KMeansPara)):
            p = (lb.KMeans.params_with_rng(k, lb.Xoshiro256Plus.seed_from_u64(40)).n_runs(1).init_method(method)
                 .pick_mode(lb.PickMode.Fast).check())
            t0 = time.perf_counter()
            p.init_centroids_resident(ctx)
            init[name] = 1e3 * (time.perf_counter() - t0)

    # ---- device-resident leg ----
    dev_ms, main_ms, launches, fallback, wall_ms, clocks = R.timed_steps(c0, args.steps, args.warmup, flush)
    value = n_total * args.steps / (dev_ms * 1e-3)
    roofline = R.roofline(args.workload, n, d, k, dtype, main_ms, args.steps)

    # ---- end-to-end leg: host buffers through the C ABI, copies inside the timed region ----
    e2e = None
    if not args.no_e2e:
        host = torch.empty((n, d), dtype=torch.float32 if dtype == np.float32 else torch.float64, pin_memory=True)
        hx = host.numpy()
        chunk = max(1, (256 << 20) // (d * itemsize))
        for r0 in range(0, n, chunk):
            hx[r0:r0 + chunk] = ctx.get_data(r0, min(chunk, n - r0))
        e2e_steps = max(3, min(args.steps, 10))
        cur = c0.copy()
        for _ in range(2):
            ctx.set_data_sharded(hx, shard_of(args.workload, rank, world)[1], n_total)
            ctx.lloyd_iters(cur, 1, 0.0)
        R.barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            ctx.set_data_sharded(hx, shard_of(args.workload, rank, world)[1], n_total)   # H2D of the step's observations (pinned)
            cur, cnts, inert, _s = ctx.lloyd_iters(cur, 1, 0.0)                          # H2D centroids, 1 iteration, D2H result
        R.barrier()
        e2e_s = R.allmax(time.perf_counter() - t0)
        e2e = {"value": n_total * e2e_steps / e2e_s, "unit": "samples*iter/s",
               "h2d_bytes_per_step": int(n * d * itemsize + k * d * itemsize),
               "d2h_bytes_per_step": int(k * d * itemsize + k * itemsize + itemsize + 64), "steps": e2e_steps,
               "ms_per_step": 1e3 * e2e_s / e2e_steps, "bytes_are": "per rank"}
        del host, hx

    # ---- the other BASELINE configs and the unordered / k-means++-started variants, same run ----
    others = None
    if not args.no_others:
        others = {}
        s2, w2 = args.steps, args.warmup
        plan = [("cfg3_hashed_rows", "cfg3", True, False), ("cfg2", "cfg2", False, False), ("cfg2_hashed_rows", "cfg2", True, False),
                ("cfg4_shard", "cfg4", False, False), ("cfg5_shard", "cfg5", False, False)]
        if world == 1:
            plan.insert(1, ("cfg3_kmeanspp_start", "cfg3", False, True))
        for name, wl, sh, kpp in plan:
            try:
                others[name] = R.side(wl, s2, w2, shuffled=sh, kmeanspp_start=kpp)
            except Exception as e:  # a side workload must not take the headline down
                others[name] = {"error": repr(e)}
        if world == 1:
            for name, fn in (("cfg3_hamerly", side_hamerly), ("cfg3_predict", side_predict), ("f64_dmma", side_dmma)):
                try:
                    others[name] = fn(R, s2)
                except Exception as e:
                    others[name] = {"error": repr(e)}

    # ---- CPU baseline (rank 0, N = 1 only) ----
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        from tests import oracle_lib

        orc = oracle_lib.load()
        v, rows, dt, threads = time_oracle(orc, n_total, d, k, dtype, 3, 1, budget_s=15.0)
        cpu = {"value": v, "unit": "samples*iter/s", "cores": threads, "kind": "port",
               "sample": f"{rows} of {n_total} rows of the same blob distribution, 3 Lloyd iterations (+1 warm-up), "
                         f"C++ restatement of linfa (oracle/), {threads} threads"}

    if rank == 0:
        cfg = config_of(args.workload, world, flush)
        if args.shuffled:
            cfg["rows"] = "hashed blob per row (unordered)"
        # the other workloads ride inside `roofline` (a key the driver keeps): `config` stays identical in both arms
        if roofline is not None:
            roofline["other_workloads"] = others
            roofline["seeding_ms"] = init or None   # k-means++ (fast pick) and k-means|| on the headline matrix, wall clock
        line = {
            "metric": "kmeans_lloyd_samples_iter_per_s", "value": value, "unit": "samples*iter/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": dev_ms / args.steps,
            "higher_is_better": True, "scaling": WORKLOADS[args.workload][4], "vs_baseline": None,
            "dtype": "f32" if dtype == np.float32 else "f64", "data": "synthetic",
            "config": cfg,
            "detail": {"rows_per_gpu": n, "wall_ms_per_step": wall_ms / args.steps, "rows_rescanned_exactly": fallback,
                       "parallelism": (f"row-sharded x{world}, per-rank partials exchanged over NVLink inside the finalize kernel (tagged 16-byte "
                                       "slots for sparse partials, stores + fence + flags for dense ones) and combined in rank order"
                                       if world > 1 else "1 GPU"),
                       # what a step spends outside the dominant kernel: the small dependent launches and, row-sharded, the exchange
                       "step_tail_us": (dev_ms / args.steps * 1e3 - roofline["mean_launch_us"]) if roofline is not None else None,
                       **init},
            "clocks": clocks, "gpu_launches": launches, "e2e": e2e, "roofline": roofline, "cpu_baseline": cpu,
        }
        print(json.dumps(line), flush=True)
    if R.dist is not None:
        R.dist.barrier()
        R.dist.destroy_process_group()
    ctx.close()


if __name__ == "__main__":
    main()

"""Builds the native libraries with nvcc for sm_100a, in-tree:

  linfa_b200/liblkm.so        the product: the C-ABI library of include/lkm.h
  linfa_b200/liblkm_probe.so  the same sources + the tcgen05 probes / micro-benchmarks of csrc/lkm_probe.cuh
                              (-DLKM_WITH_PROBES); loaded only by tests/test_gpu_tc_probe.py and tools/*.py

The .so files are git-ignored but travel to the GPU box with the snapshot.
"""
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "lkm_api.cu"), os.path.join(HERE, "csrc", "lkm_extras.cu")]
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [
    os.path.join(ROOT, "include", "lkm.h")]
OUT = os.environ.get("LKM_BUILD_OUT") or os.path.join(HERE, "liblkm.so")
OUT_PROBE = os.path.join(HERE, "liblkm_probe.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-fmad=true",  # contraction is controlled per operation with the *_rn intrinsics where the reference's order matters
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2,-Wall",
    "-shared", "-cudart", "static", "--threads", "2",
    "-Xptxas", "-v" if os.environ.get("LKM_PTXAS_V") else "-O3",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build(out=None):
    out = out or OUT
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(p) > t for p in DEPS + [os.path.abspath(__file__)])


def _compile(out, extra, verbose):
    cmd = [nvcc()] + NVCC_FLAGS + extra + os.environ.get("LKM_EXTRA_NVCC", "").split() + ["-o", out] + SRC + ["-ldl"]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed building {os.path.basename(out)}")
    if verbose or os.environ.get("LKM_PTXAS_V"):
        sys.stderr.write(res.stdout + res.stderr)
    return out


def build(force=False, verbose=False, probes=True):
    jobs = []
    if force or needs_build(OUT):
        jobs.append((OUT, []))
    if probes and (force or needs_build(OUT_PROBE)):
        jobs.append((OUT_PROBE, ["-DLKM_WITH_PROBES"]))
    if jobs:
        with ThreadPoolExecutor(max_workers=len(jobs)) as ex:
            for f in [ex.submit(_compile, o, e, verbose) for o, e in jobs]:
                f.result()
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

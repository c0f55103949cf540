"""Builds linfa_b200/liblkm.so (the C-ABI library of include/lkm.h) with nvcc for sm_100a.

In-tree build: the .so is git-ignored but travels to the GPU box with the snapshot.
"""
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = [os.path.join(HERE, "csrc", "lkm_api.cu")]
DEPS = [os.path.join(HERE, "csrc", f) for f in os.listdir(os.path.join(HERE, "csrc"))] + [
    os.path.join(ROOT, "include", "lkm.h")]
OUT = os.environ.get("LKM_BUILD_OUT") or os.path.join(HERE, "liblkm.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--use_fast_math=false" if False else "-fmad=true",  # contraction is controlled per-op with *_rn intrinsics
    "-Xcompiler", "-fPIC,-fvisibility=hidden,-O2,-Wall",
    "-shared", "-cudart", "static",
    "-Xptxas", "-v" if os.environ.get("LKM_PTXAS_V") else "-O3",
]


def nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def needs_build():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(p) > t for p in DEPS + [os.path.abspath(__file__)])


def build(force=False, verbose=False):
    if not force and not needs_build():
        return OUT
    cmd = [nvcc()] + NVCC_FLAGS + os.environ.get("LKM_EXTRA_NVCC", "").split() + ["-o", OUT] + SRC + ["-ldl"]
    if verbose:
        print(" ".join(cmd))
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("nvcc failed building liblkm.so")
    if verbose or os.environ.get("LKM_PTXAS_V"):
        sys.stderr.write(res.stdout + res.stderr)
    return OUT


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))

"""tcgen05 TF32 MMA issue-rate micro-benchmark (GPU box): cycles per MMA for several N and batch sizes."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linfa_b200 import _ffi, _probe
ctx = _probe.ProbeContext()
fn = _probe.lib.lkm_debug_mma_bench
fn.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
def run(n, reps, per, mode, grid=148):
    out = np.zeros(grid, dtype=np.int64)
    st = fn(ctx.h, n, reps, per, mode, grid, out.ctypes.data_as(C.POINTER(C.c_longlong)))
    assert st == 0, ctx.error()
    return out.mean() / (reps * per)
for n in (64, 128, 256):
    for per in (1, 4, 13, 52):
        for mode in (0, 1, 2):
            run(n, 20, per, mode)
            print(f"N={n:3d} per_batch={per:2d} mode={mode}: {run(n, 200, per, mode):7.1f} cycles/MMA  (1 CTA: {run(n, 200, per, mode, grid=1):7.1f})")

import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linfa_b200 import _ffi, _probe
ctx = _probe.ProbeContext()
bf = _probe.lib.lkm_debug_mma_bench
bf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
def run(n, reps, per, mode, grid=148):
    out = np.zeros(grid, dtype=np.int64)
    assert bf(ctx.h, n, reps, per, mode, grid, out.ctypes.data_as(C.POINTER(C.c_longlong))) == 0, ctx.error()
    return out.mean() / (reps * per)
for n in (64, 128, 256):
    for per in (7, 13, 104):
        for mode, nm in ((0, "tf32 SS"), (8, "f16 SS")):
            run(n, 20, per, mode)
            print(f"N={n:3d} per_batch={per:3d} {nm}: {run(n, 100, per, mode):7.1f} cycles/MMA")

import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linfa_b200 import _ffi, _probe
ctx = _probe.ProbeContext()
bf = _probe.lib.lkm_debug_mma_bench
bf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
def run(n, reps, mode=16, grid=148):
    out = np.zeros(grid, dtype=np.int64)
    assert bf(ctx.h, n, reps, 13, mode, grid, out.ctypes.data_as(C.POINTER(C.c_longlong))) == 0, ctx.error()
    return out.mean() / reps
for mode, nm in ((16, "tf32 SS"), (17, "tf32 TS"), (18, "f16 SS"), (19, "f16 TS")):
    for n in (32, 64, 128, 256):
        run(n, 20, mode)
        print(f"{nm} N={n:3d}: {run(n, 200, mode):7.1f} cycles per 13-MMA batch incl. commit+wait round trip")

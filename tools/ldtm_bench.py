import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linfa_b200 import _ffi, _probe
ctx = _probe.ProbeContext()
bf = _probe.lib.lkm_debug_mma_bench
bf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
for nw in (1, 2, 4, 8):
    out = np.zeros(8, dtype=np.int64)
    assert bf(ctx.h, nw * 16, 100, 0, 32, 1, out.ctypes.data_as(C.POINTER(C.c_longlong))) == 0, ctx.error()
    print(f"{nw} warps: cycles per 4 KB tcgen05.ld (per warp):", (out[:nw] / 800).round(1))

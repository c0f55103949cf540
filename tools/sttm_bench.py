"""TMEM store / load throughput on one SM and on all SMs (cycles per 4 KB tcgen05.st / ld.32x32b.x32 per warp)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linfa_b200 import _ffi, _probe
ctx = _probe.ProbeContext()
bf = _probe.lib.lkm_debug_mma_bench
bf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
for grid in (1, 148):
    for mode, name in ((32, "ld"), (33, "st"), (34, "st(w0-3)+ld(w4-7)")):
        for nw in ((1, 4, 8) if mode != 34 else (8,)):
            out = np.zeros(8, dtype=np.int64)
            assert bf(ctx.h, nw * 16, 100, 0, mode, grid, out.ctypes.data_as(C.POINTER(C.c_longlong))) == 0, ctx.error()
            print(f"grid {grid:3d} {name:18s} {nw} warps: cycles per 4 KB op (per warp):", (out[:nw] / 800).round(1))

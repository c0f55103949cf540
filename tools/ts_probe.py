"""A-from-TMEM (TS mode) probe + issue-rate benchmark (GPU box)."""
import ctypes as C, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linfa_b200 import _ffi, _probe
ctx = _probe.ProbeContext()
fn = _probe.lib.lkm_debug_tc_ts
fn.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.POINTER(C.c_float)]
g = np.random.default_rng(1)
x = (g.standard_normal((128, 32)) * 30).astype(np.float32)
c = (g.standard_normal((64, 32)) * 30).astype(np.float32)
s = np.zeros((128, 64), dtype=np.float32)
st = fn(ctx.h, _ffi.fptr(x, C.c_float), _ffi.fptr(c, C.c_float), _ffi.fptr(s, C.c_float))
print("status", st, ctx.error() if st else "")
tr = lambda a: (a.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
ref = tr(x).astype(np.float64) @ tr(c).astype(np.float64).T
print("TS result vs trunc(x).trunc(c): max abs err", np.abs(s - ref).max(), "scale", np.abs(ref).max())
bf = _probe.lib.lkm_debug_mma_bench
bf.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_longlong)]
def run(n, reps, per, mode, grid=148):
    out = np.zeros(grid, dtype=np.int64)
    assert bf(ctx.h, n, reps, per, mode, grid, out.ctypes.data_as(C.POINTER(C.c_longlong))) == 0, ctx.error()
    return out.mean() / (reps * per)
for n in (64, 128):
    for per in (13, 52):
        for mode in (0, 4):
            run(n, 20, per, mode)
            print(f"N={n:3d} per_batch={per:2d} {'TS' if mode else 'SS'}: {run(n, 200, per, mode):7.1f} cycles/MMA")

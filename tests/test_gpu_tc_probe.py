"""tcgen05 building-block probe: S = X * C^T by 3xTF32 against float64 dot products."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n_c", [64, 16, 128, 256])
def test_tc_scores_probe(n_c):
    from linfa_b200 import _ffi, _probe

    ctx = _probe.ProbeContext()
    g = np.random.default_rng(n_c)
    x = (g.uniform(-30, 30, (128, 32)) + g.standard_normal((128, 32))).astype(np.float32)
    c = g.uniform(-30, 30, (n_c, 32)).astype(np.float32)
    s = np.zeros((128, n_c), dtype=np.float32)
    fn = _probe.lib.lkm_debug_tc_scores
    fn.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_float)]
    st = fn(ctx.h, _ffi.fptr(x, C.c_float), _ffi.fptr(c, C.c_float), n_c, _ffi.fptr(s, C.c_float))
    assert st == 0, ctx.error()
    ref = x.astype(np.float64) @ c.astype(np.float64).T
    scale = np.linalg.norm(x.astype(np.float64), axis=1)[:, None] * np.linalg.norm(c.astype(np.float64), axis=1)[None, :]
    err = np.abs(s - ref) / scale
    print("max rel err (vs |x||c|):", err.max(), "plain tf32 would be ~1e-3")
    ctx.close()
    assert err.max() < 2e-6

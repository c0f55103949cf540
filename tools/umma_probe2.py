"""Round-1 tcgen05 probes (run on the GPU box): (1) does kind::tf32 truncate or round raw fp32 operands,
(2) is SBO = 0 legal for a K-major SW128 A operand (all 8-row groups alias one 1 KB atom),
(3) how the fp32 accumulator rounds when a small product is added to a large value."""
import ctypes as C
import sys, os
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from linfa_b200 import _ffi, _probe

ctx = _probe.ProbeContext()
fn = _probe.lib.lkm_debug_umma_raw
fn.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_ulonglong, C.c_ulonglong, C.c_int, C.c_int,
               C.c_uint, C.c_int, C.c_int, C.POINTER(C.c_float)]


def sw128_off(r, kk):
    c = kk >> 2
    return (r >> 3) * 1024 + (r & 7) * 128 + ((c ^ (r & 7)) << 4) + (kk & 3) * 4


def image(mat):  # mat [rows][32] f32 -> K-major SW128 bytes
    rows = mat.shape[0]
    img = np.zeros(((rows + 7) // 8) * 1024, dtype=np.uint8)
    v = img.view(np.float32)
    for r in range(rows):
        for kk in range(32):
            v[sw128_off(r, kk) // 4] = mat[r, kk]
    return img


def desc_hi(sbo, layout=2):
    return ((sbo >> 4) << 32) | (1 << 46) | (layout << 61)


def idesc(m, n):
    return (1 << 4) | (2 << 7) | (2 << 10) | ((n >> 3) << 17) | ((m >> 4) << 24)


def run(a_img, b_img, a_hi, b_hi, ksteps, ncols, n):
    dump = np.zeros((128, ncols), dtype=np.float32)
    st = fn(ctx.h, a_img.ctypes.data, a_img.size, b_img.ctypes.data, b_img.size, a_hi, b_hi, 32, 32, idesc(128, n), ksteps,
            ncols, dump.ctypes.data_as(C.POINTER(C.c_float)))
    if st != 0:
        print("status", st, ctx.error())
        return None
    return dump


g = np.random.default_rng(0)
# (1) truncation vs rounding of raw fp32 A
A = (g.standard_normal((128, 32)) * 30).astype(np.float32)
B = np.eye(32, dtype=np.float32)
d = run(image(A), image(B), desc_hi(1024), desc_hi(1024), 4, 32, 32)
bits = A.view(np.uint32)
trunc = (bits & np.uint32(0xFFFFE000)).view(np.float32)
rne = ((bits + np.uint32(0xFFF) + ((bits >> 13) & 1)) & np.uint32(0xFFFFE000)).view(np.float32)
rna = ((bits + np.uint32(0x1000)) & np.uint32(0xFFFFE000)).view(np.float32)
print("(1) A raw: equals trunc:", np.array_equal(d, trunc), " equals rne:", np.array_equal(d, rne), " equals rna:", np.array_equal(d, rna))
# same for B raw
Bm = (g.standard_normal((32, 32)) * 30).astype(np.float32)
Ai = np.zeros((128, 32), dtype=np.float32); Ai[:32] = np.eye(32)
d = run(image(Ai), image(Bm), desc_hi(1024), desc_hi(1024), 4, 32, 32)
bb = Bm.view(np.uint32)
print("(1) B raw: equals trunc:", np.array_equal(d[:32], (bb & np.uint32(0xFFFFE000)).view(np.float32).T))

# (2) SBO = 0 on A: 1 KB image, every 8-row group aliases it
A8 = (g.standard_normal((8, 32)) * 30).astype(np.float32)
A8 = (A8.view(np.uint32) & np.uint32(0xFFFFE000)).view(np.float32)
d = run(image(A8), image(B), desc_hi(0), desc_hi(1024), 4, 32, 32)
if d is not None:
    exp = np.tile(A8, (16, 1))
    print("(2) SBO=0 aliasing works:", np.array_equal(d, exp))
    if not np.array_equal(d, exp):
        print(d[:12, :4]); print(exp[:12, :4])

# (3) accumulator rounding. two k-steps (two MMAs): K0 then small; and same-MMA variant
for K0 in (16384.0, -16384.0):
    for same in (False, True):
        A3 = np.zeros((128, 32), dtype=np.float32)
        B3 = np.zeros((32, 32), dtype=np.float32)
        k_small = 1 if same else 8
        A3[:, 0] = 1.0; A3[:, k_small] = 1.0
        B3[:, 0] = K0
        small = (np.arange(32) - 16) * 2.0 ** -11   # quarter ulps of 2^14 (ulp = 2^-9)
        B3[:, k_small] = small
        d = run(image(A3), image(B3), desc_hi(1024), desc_hi(1024), 2, 32, 32)
        got = (d[0].astype(np.float64) - K0) / 2.0 ** -9
        exact = small / 2.0 ** -9
        print(f"(3) K0={K0} same_mma={same}: exact(ulps) -> got(ulps)")
        print("   ", " ".join(f"{e:+.2f}>{q:+.0f}" for e, q in zip(exact, got)))
# (3b) many products inside one MMA: 8 products each 0.25 ulp -> 2 ulps if summed before rounding
A4 = np.zeros((128, 32), dtype=np.float32); B4 = np.zeros((32, 32), dtype=np.float32)
A4[:, :8] = 1.0
B4[:, 0] = 16384.0
B4[:, 1:8] = 2.0 ** -11
d = run(image(A4), image(B4), desc_hi(1024), desc_hi(1024), 1, 32, 32)
print("(3b) K0 + 7 * 0.25ulp in one MMA -> ulps:", (float(d[0, 0]) - 16384.0) / 2.0 ** -9)

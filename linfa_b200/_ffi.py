"""ctypes binding of liblkm.so (include/lkm.h).

There is no CPU fallback: importing this module fails loudly when the CUDA
library has not been built (`python -c "import __graft_entry__ as g; g.build()"`),
and creating a context fails when no CUDA device is present.
"""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("LKM_LIB_PATH") or os.path.join(HERE, "liblkm.so")   # LKM_LIB_PATH: A/B builds while tuning

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: the B200 K-Means path has no CPU fallback. "
        "Build it with `python linfa_b200/build.py` (needs nvcc)."
    )

lib = C.CDLL(LIB_PATH)

NEXT_U32 = C.CFUNCTYPE(C.c_uint32, C.c_void_p)
NEXT_U64 = C.CFUNCTYPE(C.c_uint64, C.c_void_p)


class lkm_rng(C.Structure):
    _fields_ = [("user", C.c_void_p), ("next_u32", NEXT_U32), ("next_u64", NEXT_U64)]


class lkm_params(C.Structure):
    _fields_ = [
        ("n_clusters", C.c_uint64),
        ("n_runs", C.c_uint64),
        ("max_n_iterations", C.c_uint64),
        ("tolerance", C.c_double),
        ("init_method", C.c_int32),
        ("algorithm", C.c_int32),
        ("pick_mode", C.c_int32),
        ("reserved", C.c_int32),
        ("precomputed", C.c_void_p),
        ("precomputed_rows", C.c_uint64),
        ("precomputed_cols", C.c_uint64),
    ]


class lkm_gmm_params(C.Structure):
    _fields_ = [
        ("n_clusters", C.c_uint64),
        ("tolerance", C.c_double),
        ("reg_covariance", C.c_double),
        ("n_runs", C.c_uint64),
        ("max_n_iterations", C.c_uint64),
        ("init_method", C.c_int32),
        ("kmeans_pick_mode", C.c_int32),
    ]


class lkm_stats(C.Structure):
    _fields_ = [
        ("n_iter", C.c_uint64),
        ("last_shift", C.c_double),
        ("inertia_total", C.c_double),
        ("device_ms", C.c_float),
        ("kernel_launches", C.c_uint32),
        ("assign_path", C.c_uint32),
        ("fallback_rows", C.c_uint64),
        ("main_kernel_ms", C.c_float),
        ("update_path", C.c_uint32),
        ("rows_scanned", C.c_uint64),
        ("rows_tightened", C.c_uint64),
    ]


# status codes (include/lkm.h)
OK, ERR_INERTIA, ERR_INVALID_ARG, ERR_SAMPLE_AMOUNT, ERR_ZERO_WEIGHTS, ERR_BAD_WEIGHTS, ERR_SHAPE = 0, 1, 2, 3, 4, 5, 6
ERR_N_CLUSTERS, ERR_N_RUNS, ERR_TOLERANCE, ERR_MAX_ITERATIONS, ERR_INCREMENTAL_HAMERLY = 10, 11, 12, 13, 14
NOT_CONVERGED, ERR_CUDA, ERR_NCCL, ERR_NO_DATA, ERR_UNSUPPORTED = 20, 30, 31, 32, 33

INIT_RANDOM, INIT_PRECOMPUTED, INIT_KMEANS_PLUSPLUS, INIT_KMEANS_PARA = 0, 1, 2, 3
ALGO_LLOYD, ALGO_HAMERLY = 0, 1
PICK_EXACT, PICK_FAST, PICK_AUTO = 0, 1, 2

# every symbol include/lkm.h declares (tests check that the library exports them all)
UNTYPED_SYMBOLS = [
    "lkm_create", "lkm_destroy", "lkm_last_error", "lkm_stream", "lkm_synchronize", "lkm_version",
    "lkm_comm_unique_id", "lkm_comm_init", "lkm_set_shard", "lkm_set_assign_path", "lkm_set_l2_flush", "lkm_set_kernel_timing", "lkm_set_scoring_passes",
    "lkm_set_keep_labels", "lkm_get_labels", "lkm_set_blob_layout",
    "lkm_write_npy", "lkm_read_npy_header", "lkm_read_npy_data",
    "lkm_create_multi", "lkm_destroy_multi", "lkm_multi_last_error", "lkm_multi_size", "lkm_multi_ctx",
]
TYPED_SYMBOLS = [
    "lkm_set_data", "lkm_set_data_device", "lkm_gen_blobs", "lkm_get_data", "lkm_fit", "lkm_init",
    "lkm_lloyd_iters", "lkm_hamerly_iters", "lkm_predict", "lkm_transform", "lkm_assign", "lkm_fit_with",
    "lkm_silhouette_score", "lkm_gmm_fit", "lkm_gmm_predict", "lkm_multi_set_data", "lkm_multi_fit", "lkm_multi_assign",
]

_vp, _u64, _i64, _i32 = C.c_void_p, C.c_uint64, C.c_int64, C.c_int
_u64p = C.POINTER(C.c_uint64)
_pp = C.POINTER(lkm_params)
_sp = C.POINTER(lkm_stats)

lib.lkm_create.argtypes = [C.POINTER(_vp), _i32]
lib.lkm_destroy.argtypes = [_vp]
lib.lkm_last_error.argtypes = [_vp]
lib.lkm_last_error.restype = C.c_char_p
lib.lkm_stream.argtypes = [_vp]
lib.lkm_stream.restype = _vp
lib.lkm_synchronize.argtypes = [_vp]
lib.lkm_version.restype = C.c_char_p
lib.lkm_comm_unique_id.argtypes = [_vp]
lib.lkm_comm_init.argtypes = [_vp, _vp, _i32, _i32]
lib.lkm_set_shard.argtypes = [_vp, _u64, _u64]
lib.lkm_set_assign_path.argtypes = [_vp, _i32]
lib.lkm_set_l2_flush.argtypes = [_vp, _u64]
lib.lkm_set_kernel_timing.argtypes = [_vp, _i32]
lib.lkm_set_scoring_passes.argtypes = [_vp, _i32]
lib.lkm_set_keep_labels.argtypes = [_vp, _i32]
lib.lkm_get_labels.argtypes = [_vp, C.POINTER(C.c_uint32), _u64]
lib.lkm_set_blob_layout.argtypes = [_vp, _i32]
lib.lkm_create_multi.argtypes = [C.POINTER(_vp), C.POINTER(C.c_int), _i32]
lib.lkm_destroy_multi.argtypes = [_vp]
lib.lkm_multi_last_error.argtypes = [_vp]
lib.lkm_multi_last_error.restype = C.c_char_p
lib.lkm_multi_size.argtypes = [_vp]
lib.lkm_multi_ctx.argtypes = [_vp, _i32]
lib.lkm_multi_ctx.restype = _vp
lib.lkm_write_npy.argtypes = [C.c_char_p, _i32, _vp, _i32, _u64p]
lib.lkm_read_npy_header.argtypes = [C.c_char_p, C.POINTER(C.c_int), C.POINTER(C.c_int), _u64p, _u64p]
lib.lkm_read_npy_data.argtypes = [C.c_char_p, _u64, _vp, _u64]

for _sfx, _ct in (("f32", C.c_float), ("f64", C.c_double)):
    _fp = C.POINTER(_ct)
    getattr(lib, f"lkm_set_data_{_sfx}").argtypes = [_vp, _fp, _u64, _u64, _i64]
    getattr(lib, f"lkm_set_data_device_{_sfx}").argtypes = [_vp, _vp, _u64, _u64]
    getattr(lib, f"lkm_gen_blobs_{_sfx}").argtypes = [_vp, _u64, _u64, _u64, _fp, _ct, _u64, _u64, _u64]
    getattr(lib, f"lkm_get_data_{_sfx}").argtypes = [_vp, _u64, _u64, _fp]
    getattr(lib, f"lkm_fit_{_sfx}").argtypes = [_vp, _pp, lkm_rng, _fp, _fp, _fp, _sp]
    getattr(lib, f"lkm_init_{_sfx}").argtypes = [_vp, _pp, lkm_rng, _fp]
    getattr(lib, f"lkm_lloyd_iters_{_sfx}").argtypes = [_vp, _fp, _u64, _u64, C.c_double, _fp, _fp, _sp]
    getattr(lib, f"lkm_hamerly_iters_{_sfx}").argtypes = [_vp, _fp, _u64, _u64, C.c_double, _fp, _fp, _sp]
    getattr(lib, f"lkm_predict_{_sfx}").argtypes = [_vp, _fp, _u64, _fp, _u64, _u64, _i64, _u64p]
    getattr(lib, f"lkm_transform_{_sfx}").argtypes = [_vp, _fp, _u64, _fp, _u64, _u64, _i64, _fp]
    getattr(lib, f"lkm_assign_{_sfx}").argtypes = [_vp, _fp, _u64, _fp, _u64, _u64, _i64, _u64p, _fp]
    getattr(lib, f"lkm_fit_with_{_sfx}").argtypes = [_vp, _pp, lkm_rng, _i32, _u64, _u64, _fp, _fp, _fp]
    getattr(lib, f"lkm_silhouette_score_{_sfx}").argtypes = [_vp, _fp, _u64, _u64, _i64, _u64p, _fp]
    getattr(lib, f"lkm_multi_set_data_{_sfx}").argtypes = [_vp, _fp, _u64, _u64, _i64]
    getattr(lib, f"lkm_multi_fit_{_sfx}").argtypes = [_vp, _pp, lkm_rng, _fp, _fp, _fp, _sp]
    getattr(lib, f"lkm_multi_assign_{_sfx}").argtypes = [_vp, _fp, _u64, _fp, _u64, _u64, _i64, _u64p, _fp]
    getattr(lib, f"lkm_gmm_fit_{_sfx}").argtypes = [_vp, C.POINTER(lkm_gmm_params), lkm_rng, _fp, _fp, _fp, _fp, _fp, _u64p, _fp]
    getattr(lib, f"lkm_gmm_predict_{_sfx}").argtypes = [_vp, _u64, _fp, _fp, _fp, _fp, _u64, _u64, _i64, _u64p, _fp, _fp]


def sfx_of(dtype):
    dtype = np.dtype(dtype)
    if dtype == np.float32:
        return "f32", C.c_float
    if dtype == np.float64:
        return "f64", C.c_double
    raise TypeError(f"linfa::Float is f32 or f64, got {dtype}")


def fptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


def make_rng(rng):
    """Trampolines into a Python `Rng` (anything with next_u32()/next_u64())."""
    n32 = NEXT_U32(lambda _u: rng.next_u32() & 0xFFFFFFFF)
    n64 = NEXT_U64(lambda _u: rng.next_u64() & 0xFFFFFFFFFFFFFFFF)
    s = lkm_rng(None, n32, n64)
    s._keep = (n32, n64, rng)  # keep the callbacks alive for the duration of the call
    return s

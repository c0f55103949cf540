"""ctypes view of oracle/_build/liboracle.so — the CPU restatement of linfa's K-Means.

TEST INFRASTRUCTURE: imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs (see oracle/kmeans_oracle.cpp).
"""
import ctypes as C
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "_build", "liboracle.so")

INIT_RANDOM, INIT_PRECOMPUTED, INIT_KMEANSPP, INIT_KMEANSPARA = 0, 1, 2, 3

_u64p = C.POINTER(C.c_uint64)


def build(force=False):
    src = os.path.join(ORACLE_DIR, "kmeans_oracle.cpp")
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-B"], stdout=subprocess.DEVNULL)
    return LIB_PATH


def _ptr(a, ct):
    return a.ctypes.data_as(C.POINTER(ct)) if a is not None else None


class Xoshiro:
    """Xoshiro256Plus state held as 4 u64, advanced by the oracle library."""

    def __init__(self, lib, seed=None, state=None):
        self._lib = lib
        self.state = np.zeros(4, dtype=np.uint64)
        if state is not None:
            self.state[:] = state
        else:
            lib.orc_xoshiro_seed(C.c_uint64(seed), _ptr(self.state, C.c_uint64))

    def clone(self):
        return Xoshiro(self._lib, state=self.state.copy())

    def next_u64(self):
        return int(self._lib.orc_xoshiro_next_u64(_ptr(self.state, C.c_uint64)))

    def next_u32(self):
        return int(self._lib.orc_xoshiro_next_u32(_ptr(self.state, C.c_uint64)))

    def u01_f64(self):
        return float(self._lib.orc_u01_f64(_ptr(self.state, C.c_uint64)))

    def uniform_f64(self, low, high, size):
        """ndarray-rand `Array::random_using(size, Uniform::new(low, high), rng)`:
        row-major draws of u01*scale+low (rand 0.8.5 UniformFloat)."""
        n = int(np.prod(size))
        scale = high - low
        out = np.empty(n, dtype=np.float64)
        for i in range(n):
            out[i] = self.u01_f64() * scale + low
        return out.reshape(size)


class Oracle:
    def __init__(self, lib):
        self.lib = lib
        L = lib
        L.orc_xoshiro_next_u64.restype = C.c_uint64
        L.orc_xoshiro_next_u32.restype = C.c_uint32
        L.orc_u01_f64.restype = C.c_double
        L.orc_u01_f32.restype = C.c_float
        L.orc_gen_range_u64.restype = C.c_uint64
        L.orc_gen_range_u64.argtypes = [_u64p, C.c_uint64, C.c_uint64]
        L.orc_index_sample.argtypes = [_u64p, C.c_uint64, C.c_uint64, _u64p]
        L.orc_hardware_threads.restype = C.c_int
        for sfx, ct in (("f32", C.c_float), ("f64", C.c_double)):
            fp = C.POINTER(ct)
            u64, i32 = C.c_uint64, C.c_int

            def sig(name, restype, argtypes):
                fn = getattr(L, f"orc_{name}_{sfx}")
                fn.restype = restype
                fn.argtypes = argtypes

            sig("sq_l2_dist", ct, [fp, fp, u64])
            sig("l2_distance", ct, [fp, fp, u64])
            sig("sum", ct, [fp, u64])
            sig("assign", None, [fp, u64, u64, fp, u64, _u64p, fp, i32])
            sig("compute_centroids", None, [fp, fp, _u64p, u64, u64, u64, fp, i32])
            sig("compute_centroids_incremental", None, [fp, _u64p, u64, u64, u64, fp, fp, fp])
            sig("weighted_index", C.c_int64, [fp, u64, _u64p])
            sig("init", i32, [i32, u64, fp, u64, u64, fp, _u64p, fp, i32, i32])
            sig("weighted_kmeanspp", i32, [u64, fp, u64, u64, fp, _u64p, fp, _u64p, i32])
            sig("sample_subsequent", u64, [fp, u64, ct, u64, i32, _u64p])
            sig("membership_counts", None, [fp, u64, fp, u64, u64, fp, i32])
            sig("fit", i32, [fp, u64, u64, u64, u64, u64, ct, i32, fp, _u64p, fp, fp, fp, _u64p, _u64p, i32, i32, i32])
            sig("fit_with", i32, [fp, u64, u64, u64, u64, ct, i32, fp, _u64p, i32, fp, fp, fp, i32, i32])
            sig("fit_hamerly", i32, [fp, u64, u64, u64, u64, u64, ct, i32, fp, _u64p, fp, fp, fp, _u64p, _u64p, _u64p, i32, i32, i32])
            sig("two_closest", None, [fp, u64, u64, fp, _u64p, fp, fp])
            sig("two_farthest", None, [fp, u64, _u64p, _u64p])
            sig("compute_inertia", ct, [fp, u64, u64, _u64p, fp])
            sig("silhouette", ct, [fp, u64, u64, _u64p, fp])
        self.threads = max(1, int(L.orc_hardware_threads()))

    # -- helpers -----------------------------------------------------------
    @staticmethod
    def _sfx(dtype):
        dtype = np.dtype(dtype)
        if dtype == np.float32:
            return "f32", C.c_float
        if dtype == np.float64:
            return "f64", C.c_double
        raise TypeError(dtype)

    def _fn(self, name, sfx):
        return getattr(self.lib, f"orc_{name}_{sfx}")

    def rng(self, seed):
        return Xoshiro(self.lib, seed=seed)

    def sq_l2_dist(self, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b, dtype=a.dtype)
        sfx, ct = self._sfx(a.dtype)
        return self._fn("sq_l2_dist", sfx)(_ptr(a, ct), _ptr(b, ct), a.size)

    def l2_distance(self, a, b):
        a = np.ascontiguousarray(a)
        b = np.ascontiguousarray(b, dtype=a.dtype)
        sfx, ct = self._sfx(a.dtype)
        return self._fn("l2_distance", sfx)(_ptr(a, ct), _ptr(b, ct), a.size)

    def sum(self, x):
        x = np.ascontiguousarray(x)
        sfx, ct = self._sfx(x.dtype)
        return self._fn("sum", sfx)(_ptr(x, ct), x.size)

    def assign(self, X, Cn, threads=None):
        """update_memberships_and_dists: (labels u64, min squared dists F)."""
        X = np.ascontiguousarray(X)
        Cn = np.ascontiguousarray(Cn, dtype=X.dtype)
        sfx, ct = self._sfx(X.dtype)
        n, d = X.shape
        labels = np.empty(n, dtype=np.uint64)
        dists = np.empty(n, dtype=X.dtype)
        self._fn("assign", sfx)(_ptr(X, ct), n, d, _ptr(Cn, ct), Cn.shape[0], _ptr(labels, C.c_uint64),
                                _ptr(dists, ct), threads or self.threads)
        return labels, dists

    def compute_centroids(self, C_old, X, labels, acc_f64=False):
        X = np.ascontiguousarray(X)
        C_old = np.ascontiguousarray(C_old, dtype=X.dtype)
        labels = np.ascontiguousarray(labels, dtype=np.uint64)
        sfx, ct = self._sfx(X.dtype)
        n, d = X.shape
        out = np.empty_like(C_old)
        self._fn("compute_centroids", sfx)(_ptr(C_old, ct), _ptr(X, ct), _ptr(labels, C.c_uint64), n, d,
                                           C_old.shape[0], _ptr(out, ct), int(acc_f64))
        return out

    def compute_centroids_incremental(self, X, labels, C_old, counts):
        X = np.ascontiguousarray(X)
        C_old = np.ascontiguousarray(C_old, dtype=X.dtype)
        counts = np.ascontiguousarray(counts, dtype=X.dtype).copy()
        labels = np.ascontiguousarray(labels, dtype=np.uint64)
        sfx, ct = self._sfx(X.dtype)
        n, d = X.shape
        out = np.empty_like(C_old)
        self._fn("compute_centroids_incremental", sfx)(_ptr(X, ct), _ptr(labels, C.c_uint64), n, d, C_old.shape[0],
                                                       _ptr(C_old, ct), _ptr(counts, ct), _ptr(out, ct))
        return out, counts

    def weighted_index(self, w, rng):
        w = np.ascontiguousarray(w)
        sfx, ct = self._sfx(w.dtype)
        return int(self._fn("weighted_index", sfx)(_ptr(w, ct), w.size, _ptr(rng.state, C.c_uint64)))

    def index_sample(self, rng, length, amount):
        out = np.empty(amount, dtype=np.uint64)
        st = self.lib.orc_index_sample(_ptr(rng.state, C.c_uint64), length, amount, _ptr(out, C.c_uint64))
        if st:
            raise ValueError("amount > length")
        return out

    def init(self, method, k, X, rng, precomputed=None, threads=None, rayon_threads=1):
        X = np.ascontiguousarray(X)
        sfx, ct = self._sfx(X.dtype)
        n, d = X.shape
        out = np.empty((k, d), dtype=X.dtype)
        pre = np.ascontiguousarray(precomputed, dtype=X.dtype) if precomputed is not None else None
        st = self._fn("init", sfx)(method, k, _ptr(X, ct), n, d, _ptr(pre, ct), _ptr(rng.state, C.c_uint64),
                                   _ptr(out, ct), threads or self.threads, rayon_threads)
        if st:
            raise RuntimeError(f"oracle init status {st}")
        return out

    def weighted_kmeanspp(self, k, X, w, rng, threads=None):
        X = np.ascontiguousarray(X)
        w = np.ascontiguousarray(w, dtype=X.dtype)
        sfx, ct = self._sfx(X.dtype)
        n, d = X.shape
        out = np.empty((k, d), dtype=X.dtype)
        picked = np.empty(k, dtype=np.uint64)
        st = self._fn("weighted_kmeanspp", sfx)(k, _ptr(X, ct), n, d, _ptr(w, ct), _ptr(rng.state, C.c_uint64),
                                                _ptr(out, ct), _ptr(picked, C.c_uint64), threads or self.threads)
        if st:
            raise RuntimeError(f"oracle kmeans++ status {st}")
        return out, picked

    def sample_subsequent(self, dists, multiplier, seed, rayon_threads=1):
        dists = np.ascontiguousarray(dists)
        sfx, ct = self._sfx(dists.dtype)
        out = np.empty(dists.size, dtype=np.uint64)
        m = self._fn("sample_subsequent", sfx)(_ptr(dists, ct), dists.size, ct(multiplier), seed, rayon_threads,
                                               _ptr(out, C.c_uint64))
        return out[: int(m)].copy()

    def membership_counts(self, Cn, X, threads=None):
        X = np.ascontiguousarray(X)
        Cn = np.ascontiguousarray(Cn, dtype=X.dtype)
        sfx, ct = self._sfx(X.dtype)
        n, d = X.shape
        out = np.empty(Cn.shape[0], dtype=X.dtype)
        self._fn("membership_counts", sfx)(_ptr(Cn, ct), Cn.shape[0], _ptr(X, ct), n, d, _ptr(out, ct),
                                           threads or self.threads)
        return out

    def fit(self, X, k, rng, n_runs=10, max_iter=300, tol=1e-4, method=INIT_KMEANSPP, precomputed=None,
            threads=None, rayon_threads=1, acc_f64=False):
        """fit_lloyd.  Returns dict(status, centroids, cluster_count, inertia, n_iter, labels)."""
        X = np.ascontiguousarray(X)
        sfx, ct = self._sfx(X.dtype)
        n, d = X.shape
        Cn = np.zeros((max(k, 1), d), dtype=X.dtype)
        counts = np.zeros(max(k, 1), dtype=X.dtype)
        inertia = np.zeros(1, dtype=X.dtype)
        n_iter = np.zeros(1, dtype=np.uint64)
        labels = np.zeros(n, dtype=np.uint64)
        pre = np.ascontiguousarray(precomputed, dtype=X.dtype) if precomputed is not None else None
        st = self._fn("fit", sfx)(
            _ptr(X, ct), n, d, k, n_runs, max_iter, ct(tol), method, _ptr(pre, ct), _ptr(rng.state, C.c_uint64),
            _ptr(Cn, ct), _ptr(counts, ct), _ptr(inertia, ct), _ptr(n_iter, C.c_uint64), _ptr(labels, C.c_uint64),
            threads or self.threads, rayon_threads, int(acc_f64))
        return dict(status=int(st), centroids=Cn[:k], cluster_count=counts[:k], inertia=inertia[0],
                    n_iter=int(n_iter[0]), labels=labels)

    def fit_hamerly(self, X, k, rng, n_runs=10, max_iter=300, tol=1e-4, method=INIT_KMEANSPP, precomputed=None,
                    threads=None, rayon_threads=1, acc_f64=False):
        """fit_hamerly (KM/algorithm.rs:248-291).  Returns dict(status, centroids, cluster_count, inertia, n_iter, labels
        (the best run's), scans (two_closest_centroids calls after the first iteration))."""
        X = np.ascontiguousarray(X)
        sfx, ct = self._sfx(X.dtype)
        n, d = X.shape
        Cn = np.zeros((max(k, 1), d), dtype=X.dtype)
        counts = np.zeros(max(k, 1), dtype=X.dtype)
        inertia = np.zeros(1, dtype=X.dtype)
        n_iter = np.zeros(1, dtype=np.uint64)
        scans = np.zeros(1, dtype=np.uint64)
        labels = np.zeros(n, dtype=np.uint64)
        pre = np.ascontiguousarray(precomputed, dtype=X.dtype) if precomputed is not None else None
        st = self._fn("fit_hamerly", sfx)(
            _ptr(X, ct), n, d, k, n_runs, max_iter, ct(tol), method, _ptr(pre, ct), _ptr(rng.state, C.c_uint64),
            _ptr(Cn, ct), _ptr(counts, ct), _ptr(inertia, ct), _ptr(n_iter, C.c_uint64), _ptr(labels, C.c_uint64),
            _ptr(scans, C.c_uint64), threads or self.threads, rayon_threads, int(acc_f64))
        return dict(status=int(st), centroids=Cn[:k], cluster_count=counts[:k], inertia=inertia[0],
                    n_iter=int(n_iter[0]), labels=labels, scans=int(scans[0]))

    def two_closest(self, Cn, x):
        Cn = np.ascontiguousarray(Cn)
        x = np.ascontiguousarray(x, dtype=Cn.dtype)
        sfx, ct = self._sfx(Cn.dtype)
        idx = np.zeros(1, dtype=np.uint64)
        a, b = np.zeros(1, dtype=Cn.dtype), np.zeros(1, dtype=Cn.dtype)
        self._fn("two_closest", sfx)(_ptr(Cn, ct), Cn.shape[0], Cn.shape[1], _ptr(x, ct), _ptr(idx, C.c_uint64), _ptr(a, ct), _ptr(b, ct))
        return int(idx[0]), a[0], b[0]

    def two_farthest(self, moved):
        moved = np.ascontiguousarray(moved)
        sfx, ct = self._sfx(moved.dtype)
        f1, f2 = np.zeros(1, dtype=np.uint64), np.zeros(1, dtype=np.uint64)
        self._fn("two_farthest", sfx)(_ptr(moved, ct), moved.size, _ptr(f1, C.c_uint64), _ptr(f2, C.c_uint64))
        return int(f1[0]), int(f2[0])

    def compute_inertia(self, X, labels, Cn):
        X = np.ascontiguousarray(X)
        Cn = np.ascontiguousarray(Cn, dtype=X.dtype)
        lab = np.ascontiguousarray(labels, dtype=np.uint64)
        sfx, ct = self._sfx(X.dtype)
        return self._fn("compute_inertia", sfx)(_ptr(X, ct), X.shape[0], X.shape[1], _ptr(lab, C.c_uint64), _ptr(Cn, ct))

    def silhouette(self, X, labels):
        X = np.ascontiguousarray(X)
        lab = np.ascontiguousarray(labels).astype(np.uint64)
        sfx, ct = self._sfx(X.dtype)
        return self._fn("silhouette", sfx)(_ptr(X, ct), X.shape[0], X.shape[1], _ptr(lab, C.c_uint64), None)

    def fit_with(self, X, k, rng, model=None, n_runs=10, tol=1e-4, method=INIT_KMEANSPP, precomputed=None,
                 threads=None, rayon_threads=1):
        X = np.ascontiguousarray(X)
        sfx, ct = self._sfx(X.dtype)
        n, d = X.shape
        if model is not None:
            Cn = np.ascontiguousarray(model[0], dtype=X.dtype).copy()
            counts = np.ascontiguousarray(model[1], dtype=X.dtype).copy()
        else:
            Cn = np.zeros((k, d), dtype=X.dtype)
            counts = np.zeros(k, dtype=X.dtype)
        inertia = np.zeros(1, dtype=X.dtype)
        pre = np.ascontiguousarray(precomputed, dtype=X.dtype) if precomputed is not None else None
        st = self._fn("fit_with", sfx)(
            _ptr(X, ct), n, d, k, n_runs, ct(tol), method, _ptr(pre, ct), _ptr(rng.state, C.c_uint64),
            int(model is not None), _ptr(Cn, ct), _ptr(counts, ct), _ptr(inertia, ct), threads or self.threads,
            rayon_threads)
        return dict(status=int(st), centroids=Cn, cluster_count=counts, inertia=inertia[0])


_cached = None


def load():
    global _cached
    if _cached is None:
        _cached = Oracle(C.CDLL(build()))
    return _cached

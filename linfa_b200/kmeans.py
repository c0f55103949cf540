"""Python mirror of linfa-clustering's K-Means surface over the lkm C ABI.

Names, argument meaning, defaults and error behaviour follow the reference
(KM = algorithms/linfa-clustering/src/k_means):
  KMeans::params / params_with_rng / params_with      KM/algorithm.rs:209-222
  KMeansParams builder + ParamGuard                   KM/hyperparams.rs:52-142
  KMeansInit / KMeansAlgorithm                        KM/init.rs:22-79
  Fit / FitWith / PredictInplace / Transformer impls  KM/algorithm.rs:370-388, 614-777
  KMeansParamsError / KMeansError / IncrKMeansError   KM/errors.rs
All arithmetic on the data happens in the CUDA library; this file only holds
parameters, converts arrays and maps status codes to the reference's errors.
"""
import copy
import ctypes as C

import numpy as np

from . import _ffi
from .rng import Xoshiro256Plus


# ---- errors (KM/errors.rs) -------------------------------------------------
class KMeansParamsError(Exception):
    NClusters = "n_clusters cannot be 0"
    NRuns = "n_runs cannot be 0"
    Tolerance = "tolerance must be greater than 0"
    MaxIterations = "max_n_iterations cannot be 0"
    IncrementalHamerly = ("only KMeansAlgorithm::Lloyd is supported by fit_with (Mini-Batch K-means); "
                          "Hamerly requires a persistent dataset across iterations and cannot be used incrementally")

    def __init__(self, kind):
        self.kind = kind
        super().__init__(getattr(KMeansParamsError, kind))


class KMeansError(Exception):
    """InvalidParams(KMeansParamsError) | InertiaError | LinfaError"""

    def __init__(self, kind, inner=None, msg=None):
        self.kind, self.inner = kind, inner
        text = {"InvalidParams": f"Invalid hyperparameter: {inner}",
                "InertiaError": "Fitting failed: No inertia improvement (-inf)"}.get(kind, msg or kind)
        super().__init__(text)


class IncrKMeansError(Exception):
    """InvalidParams(KMeansParamsError) | NotConverged(model) | LinfaError"""

    def __init__(self, kind, inner=None, model=None):
        self.kind, self.inner, self.model = kind, inner, model
        super().__init__("Algorithm has not yet converged, Keep on running the algorithm."
                         if kind == "NotConverged" else f"Invalid hyperparameter: {inner}")


_PARAM_ERRORS = {_ffi.ERR_N_CLUSTERS: "NClusters", _ffi.ERR_N_RUNS: "NRuns", _ffi.ERR_TOLERANCE: "Tolerance",
                 _ffi.ERR_MAX_ITERATIONS: "MaxIterations", _ffi.ERR_INCREMENTAL_HAMERLY: "IncrementalHamerly"}


class LkmError(RuntimeError):
    """CUDA / NCCL / unsupported-shape failures of the native library (no reference counterpart)."""

    def __init__(self, status, msg):
        self.status = status
        super().__init__(f"lkm status {status}: {msg}")


# ---- small stand-ins for linfa core types ----------------------------------
class L2Dist:
    """linfa_nn::distance::L2Dist (linfa-nn/src/distance.rs:60-81); the only metric of the GPU path."""

    def __eq__(self, other):
        return isinstance(other, L2Dist)


class DatasetBase:
    """linfa::DatasetBase<records, targets> (src/dataset/mod.rs:176-186), records/targets only."""

    def __init__(self, records, targets=None):
        self.records = records
        self.targets = targets

    @classmethod
    def from_(cls, records):
        return cls(records)

    def nsamples(self):
        return self.records.shape[0]

    def silhouette_score(self, ctx=None):
        """SilhouetteScore::silhouette_score (src/metrics_clustering.rs:65-135) of (records, targets) on the GPU; a
        dataset without targets counts as one cluster (test_empty_labels_as_single_label, :186-192)."""
        x = np.ascontiguousarray(self.records)
        sfx, ct = _ffi.sfx_of(x.dtype)
        n, d = x.shape
        labels = np.zeros(n, dtype=np.uint64) if self.targets is None else np.ascontiguousarray(self.targets).astype(np.uint64)
        out = np.zeros(1, dtype=x.dtype)
        ctx = ctx or Context.default()
        ctx.check(getattr(_ffi.lib, f"lkm_silhouette_score_{sfx}")(ctx.h, _ffi.fptr(x, ct), n, d, d, labels.ctypes.data_as(_ffi._u64p),
                                                                   _ffi.fptr(out, ct)))
        return out[0]


Dataset = DatasetBase

NPY_DTYPES = {np.dtype(np.float32): 0, np.dtype(np.float64): 1, np.dtype(np.uint64): 2}


def write_npy(path, array):
    """ndarray_npy::write_npy as the reference's example uses it (examples/kmeans.rs:40-42), through the library's writer."""
    a = np.ascontiguousarray(array)
    if a.dtype not in NPY_DTYPES:
        raise TypeError(f"write_npy: f32, f64 or u64, got {a.dtype}")
    shape = np.asarray(a.shape, dtype=np.uint64)
    st = _ffi.lib.lkm_write_npy(str(path).encode(), NPY_DTYPES[a.dtype], a.ctypes.data_as(C.c_void_p), a.ndim,
                                shape.ctypes.data_as(_ffi._u64p))
    if st != _ffi.OK:
        raise LkmError(st, f"write_npy({path}) failed")


def read_npy(path):
    dt, nd, off = C.c_int(), C.c_int(), C.c_uint64()
    shape = np.zeros(8, dtype=np.uint64)
    st = _ffi.lib.lkm_read_npy_header(str(path).encode(), C.byref(dt), C.byref(nd), shape.ctypes.data_as(_ffi._u64p), C.byref(off))
    if st != _ffi.OK:
        raise LkmError(st, f"read_npy({path}): not a little-endian C-order f32 / f64 / u64 .npy file")
    dtype = [np.float32, np.float64, np.uint64][dt.value]
    out = np.empty(tuple(int(v) for v in shape[: nd.value]), dtype=dtype)
    st = _ffi.lib.lkm_read_npy_data(str(path).encode(), off.value, out.ctypes.data_as(C.c_void_p), out.nbytes)
    if st != _ffi.OK:
        raise LkmError(st, f"read_npy({path}): truncated file")
    return out


class generate:
    """linfa_datasets::generate (datasets/src/generate.rs)."""

    @staticmethod
    def blobs(blob_size, blob_centroids, rng, ctx=None, sigma=1.0):
        """generate::blobs (datasets/src/generate.rs:12-46): `blob_size` points around each row of `blob_centroids`
        (standard normal noise), blob after blob, as an (n_blobs * blob_size) x n_features matrix.  Generated on the GPU by
        the library's counter-based generator, keyed by one u64 drawn from `rng`; the layout and the distribution are the
        reference's, the random stream is not (rand_distr's ziggurat tables are not restated: data is an INPUT of the path)."""
        c = np.ascontiguousarray(blob_centroids)
        if c.dtype not in (np.float32, np.float64):
            c = c.astype(np.float64)
        ctx = ctx or Context.default()
        n = int(blob_size) * c.shape[0]
        ctx.gen_blobs(n, c.shape[1], c, sigma, seed=int(rng.next_u64()))
        return ctx.get_data(0, n)


class KMeansInit:
    """KMeansInit<F> (KM/init.rs:22-35)."""

    def __init__(self, kind, centroids=None):
        self.kind = kind
        self.centroids = centroids

    @staticmethod
    def Precomputed(centroids):
        return KMeansInit("Precomputed", np.asarray(centroids))

    def __eq__(self, other):
        return (isinstance(other, KMeansInit) and self.kind == other.kind
                and (self.centroids is None) == (other.centroids is None)
                and (self.centroids is None or np.array_equal(self.centroids, other.centroids)))

    def __repr__(self):
        return f"KMeansInit::{self.kind}"


KMeansInit.Random = KMeansInit("Random")
KMeansInit.KMeansPlusPlus = KMeansInit("KMeansPlusPlus")
KMeansInit.KMeansPara = KMeansInit("KMeansPara")
_INIT_CODE = {"Random": _ffi.INIT_RANDOM, "Precomputed": _ffi.INIT_PRECOMPUTED,
              "KMeansPlusPlus": _ffi.INIT_KMEANS_PLUSPLUS, "KMeansPara": _ffi.INIT_KMEANS_PARA}


class KMeansAlgorithm:
    Lloyd = "Lloyd"
    Hamerly = "Hamerly"


class PickMode:
    """Not in the reference: how the k-means++ weighted pick walks its prefix sum (include/lkm.h)."""
    Exact = _ffi.PICK_EXACT   # rand 0.8.5 WeightedIndex bit for bit (one GPU thread per pick: ~1 ms per million rows and pick)
    Fast = _ffi.PICK_FAST     # fixed-order parallel scan, same uniform draw
    Auto = _ffi.PICK_AUTO     # Exact up to 2^22 rows, Fast above


# ---- device context ---------------------------------------------------------
class Context:
    """One lkm_ctx = one GPU.  `Context.default()` is what fit/predict use."""

    _default = None

    def __init__(self, device=0):
        h = C.c_void_p()
        st = _ffi.lib.lkm_create(C.byref(h), int(device))
        if st != _ffi.OK:
            raise LkmError(st, "lkm_create failed: no usable CUDA device (this path has no CPU fallback)")
        self.h = h
        self.device = device
        self.rank, self.world = 0, 1
        self.last_stats = None

    @classmethod
    def default(cls):
        if cls._default is None:
            cls._default = cls(0)
        return cls._default

    @classmethod
    def set_default(cls, ctx):
        cls._default = ctx

    def close(self):
        if self.h:
            _ffi.lib.lkm_destroy(self.h)
            self.h = None

    def error(self):
        return _ffi.lib.lkm_last_error(self.h).decode()

    def check(self, st):
        if st != _ffi.OK:
            raise LkmError(st, self.error())

    def comm_init(self, uid_bytes, rank, world):
        buf = C.create_string_buffer(bytes(uid_bytes), 128)
        self.check(_ffi.lib.lkm_comm_init(self.h, buf, rank, world))
        self.rank, self.world = rank, world

    def set_shard(self, row_offset, n_global):
        self.check(_ffi.lib.lkm_set_shard(self.h, row_offset, n_global))

    def set_assign_path(self, path):
        self.check(_ffi.lib.lkm_set_assign_path(self.h, path))

    def set_l2_flush(self, nbytes):
        self.check(_ffi.lib.lkm_set_l2_flush(self.h, int(nbytes)))

    def set_kernel_timing(self, on):
        self.check(_ffi.lib.lkm_set_kernel_timing(self.h, int(bool(on))))

    def set_keep_labels(self, on):
        """Parity-test aid: the Lloyd kernels also write the labels of the last iteration (see get_labels)."""
        self.check(_ffi.lib.lkm_set_keep_labels(self.h, int(bool(on))))

    def get_labels(self, n=None):
        """Labels every row was accumulated under in the last executed Lloyd iteration (needs set_keep_labels)."""
        n = self.shape[0] if n is None else n
        out = np.empty(n, dtype=np.uint32)
        self.check(_ffi.lib.lkm_get_labels(self.h, out.ctypes.data_as(C.POINTER(C.c_uint32)), n))
        return out

    def set_blob_layout(self, shuffled):
        """gen_blobs layout: False = contiguous blobs (linfa's generate::blobs), True = hashed blob per row."""
        self.check(_ffi.lib.lkm_set_blob_layout(self.h, int(bool(shuffled))))

    def set_scoring_passes(self, passes):
        """0 adaptive (default), 1 / 2 / 3 pin the tensor-core scoring variant (speed only, never results)."""
        self.check(_ffi.lib.lkm_set_scoring_passes(self.h, int(passes)))

    def set_data(self, x):
        x = np.asarray(x)
        if x.ndim != 2:
            raise ValueError("records must be 2-dimensional (n_observations, n_features)")
        sfx, ct = _ffi.sfx_of(x.dtype)
        x = _row_major_view(x)
        self._keep = x
        stride = x.strides[0] // x.itemsize if x.shape[0] > 1 else x.shape[1]
        self.check(getattr(_ffi.lib, f"lkm_set_data_{sfx}")(self.h, _ffi.fptr(x, ct), x.shape[0], x.shape[1], stride))
        self.dtype, self.shape = x.dtype, x.shape

    def set_data_sharded(self, x_local, row_offset, n_global):
        """This rank's row block [row_offset, row_offset + len(x_local)) of an n_global-row matrix."""
        self.set_data(x_local)
        self.set_shard(row_offset, n_global)

    def set_data_device(self, ptr, n, d, dtype):
        sfx, _ = _ffi.sfx_of(dtype)
        self.check(getattr(_ffi.lib, f"lkm_set_data_device_{sfx}")(self.h, C.c_void_p(ptr), n, d))
        self.dtype, self.shape = np.dtype(dtype), (n, d)

    def gen_blobs(self, n, d, centres, sigma=1.0, seed=0, row_offset=0, n_global=0):
        centres = np.ascontiguousarray(centres)
        sfx, ct = _ffi.sfx_of(centres.dtype)
        self.check(getattr(_ffi.lib, f"lkm_gen_blobs_{sfx}")(self.h, n, d, centres.shape[0], _ffi.fptr(centres, ct),
                                                             ct(sigma), seed, row_offset, n_global))
        self.dtype, self.shape = centres.dtype, (n, d)

    def get_data(self, row0=0, rows=None):
        n, d = self.shape
        rows = n - row0 if rows is None else rows
        sfx, ct = _ffi.sfx_of(self.dtype)
        out = np.empty((rows, d), dtype=self.dtype)
        self.check(getattr(_ffi.lib, f"lkm_get_data_{sfx}")(self.h, row0, rows, _ffi.fptr(out, ct)))
        return out

    def lloyd_iters(self, centroids, max_iters, tolerance, entry="lkm_lloyd_iters"):
        """fit_lloyd's loop from given centroids; returns (centroids, counts, inertia_total, stats)."""
        c = np.ascontiguousarray(centroids, dtype=self.dtype).copy()
        sfx, ct = _ffi.sfx_of(self.dtype)
        k = c.shape[0]
        counts = np.zeros(k, dtype=self.dtype)
        inertia = np.zeros(1, dtype=self.dtype)
        stats = _ffi.lkm_stats()
        self.check(getattr(_ffi.lib, f"{entry}_{sfx}")(self.h, _ffi.fptr(c, ct), k, max_iters, float(tolerance),
                                                       _ffi.fptr(counts, ct), _ffi.fptr(inertia, ct), C.byref(stats)))
        self.last_stats = stats
        return c, counts, inertia[0], stats

    def hamerly_iters(self, centroids, max_iters, tolerance):
        """fit_hamerly's loop (KM/algorithm.rs:262-279) from given centroids: same returns as lloyd_iters, the inertia
        being compute_inertia against the returned centroids; stats.rows_scanned / rows_tightened count the bound misses."""
        return self.lloyd_iters(centroids, max_iters, tolerance, entry="lkm_hamerly_iters")

    def assign(self, centroids, x=None, want_labels=True, want_dists=True):
        """update_memberships_and_dists on `x` (host array) or on the resident matrix."""
        c = np.ascontiguousarray(centroids)
        sfx, ct = _ffi.sfx_of(c.dtype)
        if x is None:
            n, d = self.shape
            xp, stride = None, d
        else:
            x = _row_major_view(np.asarray(x, dtype=c.dtype))
            n, d = x.shape
            xp, stride = _ffi.fptr(x, ct), (x.strides[0] // x.itemsize if n > 1 else d)
        if c.shape[1] != d:
            raise ValueError("centroids and observations disagree on n_features")
        labels = np.empty(n, dtype=np.uint64) if want_labels else None
        dists = np.empty(n, dtype=c.dtype) if want_dists else None
        self.check(getattr(_ffi.lib, f"lkm_assign_{sfx}")(self.h, _ffi.fptr(c, ct), c.shape[0], xp, n, d, stride,
                                                          _ffi.fptr(labels, C.c_uint64), _ffi.fptr(dists, ct)))
        return labels, dists


def _row_major_view(x):
    """An ndarray view whose rows are contiguous (any row stride >= n_features) is passed as is,
    like an ndarray ArrayView2; anything else is copied to standard layout."""
    if (x.shape[0] <= 1 or x.shape[1] == 0 or x.strides[1] != x.itemsize
            or x.strides[0] < x.shape[1] * x.itemsize or x.strides[0] % x.itemsize):
        return np.ascontiguousarray(x)
    return x


def _records(dataset):
    return dataset.records if isinstance(dataset, DatasetBase) else dataset


class DeviceGroup:
    """lkm_multi: several GPUs driven from this one process (include/lkm.h, lkm_create_multi).  Rows are split into
    contiguous shards; `fit` is Fit::fit over the sharded matrix, `assign` is update_memberships_and_dists."""

    def __init__(self, device_ids):
        ids = (C.c_int * len(device_ids))(*device_ids)
        h = C.c_void_p()
        st = _ffi.lib.lkm_create_multi(C.byref(h), ids, len(device_ids))
        if st != _ffi.OK:
            raise LkmError(st, "lkm_create_multi failed: no usable CUDA device (this path has no CPU fallback)")
        self.h, self.device_ids, self.dtype, self.shape = h, list(device_ids), None, None

    def close(self):
        if self.h:
            _ffi.lib.lkm_destroy_multi(self.h)
            self.h = None

    def check(self, st):
        if st != _ffi.OK:
            raise LkmError(st, _ffi.lib.lkm_multi_last_error(self.h).decode())

    def set_data(self, x):
        x = _row_major_view(np.asarray(x))
        sfx, ct = _ffi.sfx_of(x.dtype)
        self.check(getattr(_ffi.lib, f"lkm_multi_set_data_{sfx}")(self.h, _ffi.fptr(x, ct), x.shape[0], x.shape[1], x.strides[0] // x.itemsize))
        self.dtype, self.shape = x.dtype, x.shape

    def fit(self, params, dataset):
        """params: KMeansParams / KMeansValidParams (checked here like the ParamGuard blanket impl does)."""
        v = params.check() if hasattr(params, "check") else params
        self.set_data(_records(dataset))
        sfx, ct = _ffi.sfx_of(self.dtype)
        k, d = v.n_clusters(), self.shape[1]
        p, keep = v._native_params(self.dtype)
        rng = copy.deepcopy(v._rng)
        cen, cnt, inertia = np.zeros((k, d), self.dtype), np.zeros(k, self.dtype), np.zeros(1, self.dtype)
        stats = _ffi.lkm_stats()
        self.check(getattr(_ffi.lib, f"lkm_multi_fit_{sfx}")(self.h, C.byref(p), _ffi.make_rng(rng), _ffi.fptr(cen, ct), _ffi.fptr(cnt, ct),
                                                              _ffi.fptr(inertia, ct), C.byref(stats)))
        self.last_stats = stats
        return KMeans(cen, cnt, inertia[0], v._dist_fn, None)

    def assign(self, centroids, x):
        c = np.ascontiguousarray(centroids)
        x = np.ascontiguousarray(x, dtype=c.dtype)
        sfx, ct = _ffi.sfx_of(c.dtype)
        labels, dists = np.zeros(x.shape[0], np.uint64), np.zeros(x.shape[0], c.dtype)
        self.check(getattr(_ffi.lib, f"lkm_multi_assign_{sfx}")(self.h, _ffi.fptr(c, ct), c.shape[0], _ffi.fptr(x, ct), x.shape[0], x.shape[1],
                                                                 x.shape[1], labels.ctypes.data_as(_ffi._u64p), _ffi.fptr(dists, ct)))
        return labels, dists


# ---- hyper-parameters --------------------------------------------------------
class KMeansValidParams:
    """KMeansValidParams (KM/hyperparams.rs:19-42) + the Fit / FitWith impls."""

    def __init__(self, n_clusters, rng, dist_fn):
        # defaults: KM/hyperparams.rs:71-82
        self._n_runs = 10
        self._tolerance = 1e-4
        self._max_n_iterations = 300
        self._n_clusters = n_clusters
        self._init = KMeansInit.KMeansPlusPlus
        self._rng = rng
        self._dist_fn = dist_fn
        self._algorithm = KMeansAlgorithm.Lloyd
        self._pick_mode = PickMode.Exact
        self._ctx = None

    def n_runs(self):
        return self._n_runs

    def tolerance(self):
        return self._tolerance

    def max_n_iterations(self):
        return self._max_n_iterations

    def n_clusters(self):
        return self._n_clusters

    def init_method(self):
        return self._init

    def rng(self):
        return self._rng

    def dist_fn(self):
        return self._dist_fn

    def algorithm(self):
        return self._algorithm

    # -- native call plumbing
    def _native_params(self, dtype, check_shape_d=None):
        p = _ffi.lkm_params()
        p.n_clusters, p.n_runs, p.max_n_iterations = self._n_clusters, self._n_runs, self._max_n_iterations
        p.tolerance = float(np.dtype(dtype).type(self._tolerance))
        p.init_method = _INIT_CODE[self._init.kind]
        p.algorithm = _ffi.ALGO_HAMERLY if self._algorithm == KMeansAlgorithm.Hamerly else _ffi.ALGO_LLOYD
        p.pick_mode = self._pick_mode
        keep = None
        if self._init.kind == "Precomputed":
            keep = np.ascontiguousarray(self._init.centroids, dtype=dtype)
            p.precomputed = keep.ctypes.data
            p.precomputed_rows, p.precomputed_cols = keep.shape
        return p, keep

    def fit(self, dataset, ctx=None):
        """Fit::fit (KM/algorithm.rs:370-388) -> fit_lloyd (:294-343)."""
        x = np.asarray(_records(dataset))
        ctx = ctx or self._ctx or Context.default()
        ctx.set_data(x)
        return self.fit_resident(ctx)

    def fit_resident(self, ctx):
        """fit on the matrix already resident in `ctx` (set_data / gen_blobs / set_data_device)."""
        dtype = ctx.dtype
        sfx, ct = _ffi.sfx_of(dtype)
        k, d = self._n_clusters, ctx.shape[1]
        p, keep = self._native_params(dtype)
        rng = copy.deepcopy(self._rng)  # `let mut rng = self.rng().clone()` (KM/algorithm.rs:298)
        cen = np.zeros((max(k, 1), d), dtype=dtype)
        cnt = np.zeros(max(k, 1), dtype=dtype)
        inertia = np.zeros(1, dtype=dtype)
        stats = _ffi.lkm_stats()
        st = getattr(_ffi.lib, f"lkm_fit_{sfx}")(ctx.h, C.byref(p), _ffi.make_rng(rng), _ffi.fptr(cen, ct),
                                                 _ffi.fptr(cnt, ct), _ffi.fptr(inertia, ct), C.byref(stats))
        ctx.last_stats = stats
        if st in _PARAM_ERRORS:
            raise KMeansError("InvalidParams", KMeansParamsError(_PARAM_ERRORS[st]))
        if st == _ffi.ERR_INERTIA:
            raise KMeansError("InertiaError")
        if st in (_ffi.ERR_SHAPE, _ffi.ERR_SAMPLE_AMOUNT, _ffi.ERR_ZERO_WEIGHTS, _ffi.ERR_BAD_WEIGHTS):
            raise AssertionError(ctx.error())  # the reference panics here (assert!/expect)
        ctx.check(st)
        return KMeans(cen[:k], cnt[:k], inertia[0], self._dist_fn, ctx)

    def fit_with(self, model, dataset, ctx=None):
        """FitWith::fit_with (KM/algorithm.rs:635-715): Ok(model) or raises IncrKMeansError."""
        if self._algorithm == KMeansAlgorithm.Hamerly:  # first thing the reference checks (:640-644)
            raise IncrKMeansError("InvalidParams", KMeansParamsError("IncrementalHamerly"))
        x = np.asarray(_records(dataset))
        ctx = ctx or self._ctx or Context.default()
        dtype = x.dtype
        sfx, ct = _ffi.sfx_of(dtype)
        k, d = self._n_clusters, x.shape[1]
        p, keep = self._native_params(dtype)
        ctx.set_data(x)
        if model is not None:
            # the reference never looks at n_clusters here: the model's own shape rules (KM/algorithm.rs:680-690)
            cen = np.ascontiguousarray(model.centroids(), dtype=dtype).copy()
            cnt = np.ascontiguousarray(model.cluster_count(), dtype=dtype).copy()
            if cen.ndim != 2 or cnt.shape != (cen.shape[0],):
                raise AssertionError("model.cluster_count must have one entry per centroid")
            k = cen.shape[0]
        else:
            cen = np.zeros((max(k, 1), d), dtype=dtype)
            cnt = np.zeros(max(k, 1), dtype=dtype)
        inertia = np.zeros(1, dtype=dtype)
        rng = copy.deepcopy(self._rng)
        st = getattr(_ffi.lib, f"lkm_fit_with_{sfx}")(ctx.h, C.byref(p), _ffi.make_rng(rng), int(model is not None),
                                                      cen.shape[0], cen.shape[1],
                                                      _ffi.fptr(cen, ct), _ffi.fptr(cnt, ct), _ffi.fptr(inertia, ct))
        if st in _PARAM_ERRORS:
            raise IncrKMeansError("InvalidParams", KMeansParamsError(_PARAM_ERRORS[st]))
        if st in (_ffi.ERR_SHAPE, _ffi.ERR_SAMPLE_AMOUNT, _ffi.ERR_ZERO_WEIGHTS, _ffi.ERR_BAD_WEIGHTS):
            raise AssertionError(ctx.error())
        out = KMeans(cen[:k], cnt[:k], inertia[0], self._dist_fn, ctx)
        if st == _ffi.NOT_CONVERGED:
            raise IncrKMeansError("NotConverged", model=out)
        ctx.check(st)
        return out

    def init_centroids(self, dataset, ctx=None):
        """KMeansInit::run (KM/init.rs:83-101) alone."""
        x = np.asarray(_records(dataset))
        ctx = ctx or self._ctx or Context.default()
        ctx.set_data(x)
        return self.init_centroids_resident(ctx)

    def init_centroids_resident(self, ctx):
        """KMeansInit::run on the matrix already resident in `ctx`."""
        dtype, shape = ctx.dtype, ctx.shape
        sfx, ct = _ffi.sfx_of(dtype)
        p, keep = self._native_params(dtype)
        out = np.zeros((self._n_clusters, shape[1]), dtype=dtype)
        rng = copy.deepcopy(self._rng)
        st = getattr(_ffi.lib, f"lkm_init_{sfx}")(ctx.h, C.byref(p), _ffi.make_rng(rng), _ffi.fptr(out, ct))
        if st in (_ffi.ERR_SHAPE, _ffi.ERR_SAMPLE_AMOUNT, _ffi.ERR_ZERO_WEIGHTS, _ffi.ERR_BAD_WEIGHTS):
            raise AssertionError(ctx.error())
        ctx.check(st)
        return out


class KMeansParams:
    """KMeansParams newtype + consuming builder + ParamGuard (KM/hyperparams.rs:50-142)."""

    def __init__(self, n_clusters, rng, dist_fn):
        self._v = KMeansValidParams(n_clusters, rng, dist_fn)

    def _set(self, name, value):
        setattr(self._v, name, value)
        return self

    def n_runs(self, n_runs):
        return self._set("_n_runs", n_runs)

    def tolerance(self, tolerance):
        return self._set("_tolerance", tolerance)

    def max_n_iterations(self, max_n_iterations):
        return self._set("_max_n_iterations", max_n_iterations)

    def init_method(self, init):
        return self._set("_init", init)

    def algorithm(self, algorithm):
        return self._set("_algorithm", algorithm)

    def pick_mode(self, mode):
        return self._set("_pick_mode", mode)

    def context(self, ctx):
        return self._set("_ctx", ctx)

    # ParamGuard (src/param_guard.rs:15-35); checks in the reference's order (KM/hyperparams.rs:124-136)
    def check_ref(self):
        v = self._v
        if v._n_clusters == 0:
            raise KMeansParamsError("NClusters")
        if v._n_runs == 0:
            raise KMeansParamsError("NRuns")
        if v._tolerance <= 0:
            raise KMeansParamsError("Tolerance")
        if v._max_n_iterations == 0:
            raise KMeansParamsError("MaxIterations")
        return v

    def check(self):
        return self.check_ref()

    def check_unwrap(self):
        return self.check_ref()

    # blanket impls: Fit / FitWith for P: ParamGuard (src/param_guard.rs:55-86)
    def fit(self, dataset, ctx=None):
        try:
            v = self.check_ref()
        except KMeansParamsError as e:
            raise KMeansError("InvalidParams", e) from None
        return v.fit(dataset, ctx)

    def fit_with(self, model, dataset, ctx=None):
        try:
            v = self.check_ref()
        except KMeansParamsError as e:
            raise IncrKMeansError("InvalidParams", e) from None
        return v.fit_with(model, dataset, ctx)


# ---- model ---------------------------------------------------------------------
class KMeans:
    """KMeans<F, D> (KM/algorithm.rs:202-241)."""

    def __init__(self, centroids, cluster_count, inertia, dist_fn=None, ctx=None):
        self._centroids = np.ascontiguousarray(centroids)
        self._cluster_count = np.asarray(cluster_count)
        self._inertia = inertia
        self._dist_fn = dist_fn or L2Dist()
        self._ctx = ctx

    @staticmethod
    def params(n_clusters):
        return KMeansParams(n_clusters, Xoshiro256Plus.seed_from_u64(42), L2Dist())

    @staticmethod
    def params_with_rng(n_clusters, rng):
        return KMeansParams(n_clusters, rng, L2Dist())

    @staticmethod
    def params_with(n_clusters, rng, dist_fn):
        if not isinstance(dist_fn, L2Dist):
            raise NotImplementedError("the B200 path implements L2Dist only; other metrics stay on linfa-clustering")
        return KMeansParams(n_clusters, rng, dist_fn)

    def centroids(self):
        return self._centroids

    def cluster_count(self):
        return self._cluster_count

    def inertia(self):
        return self._inertia

    def _context(self):
        return self._ctx or Context.default()

    def predict_inplace(self, observations, memberships):
        """PredictInplace<Array2, Array1<usize>> (KM/algorithm.rs:743-756)."""
        x = np.asarray(observations, dtype=self._centroids.dtype)
        assert x.shape[0] == memberships.shape[0], "The number of data points must match the number of memberships."
        labels, _ = self._context().assign(self._centroids, x, want_dists=False)
        memberships[...] = labels

    def predict(self, x):
        """The Predict blanket impls (src/dataset/impl_dataset.rs:1153-1202)."""
        if isinstance(x, DatasetBase):
            rec = np.asarray(x.records)
            t = np.zeros(rec.shape[0], dtype=np.uint64)
            self.predict_inplace(rec, t)
            return DatasetBase(x.records, t)
        a = np.asarray(x, dtype=self._centroids.dtype)
        if a.ndim == 1:  # PredictInplace<Array1, usize> (KM/algorithm.rs:763-777)
            labels, _ = self._context().assign(self._centroids, a[None, :], want_dists=False)
            return int(labels[0])
        t = np.zeros(a.shape[0], dtype=np.uint64)
        self.predict_inplace(a, t)
        return t

    def transform(self, observations):
        """Transformer::transform (KM/algorithm.rs:718-733): min squared distance per observation."""
        x = np.asarray(observations, dtype=self._centroids.dtype)
        _, d = self._context().assign(self._centroids, x, want_labels=False)
        return d

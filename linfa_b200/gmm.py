"""Python mirror of linfa-clustering's GaussianMixtureModel (algorithms/linfa-clustering/src/gaussian_mixture/):
`GaussianMixtureModel.params(k).tolerance(..).reg_covariance(..).n_runs(..).max_n_iterations(..).init_method(..)
.with_rng(..).fit(dataset)`, accessors, `predict`, `predict_proba` — over the C ABI of include/lkm.h (lkm_gmm_*)."""
import copy
import ctypes as C

import numpy as np

from . import _ffi
from .kmeans import Context, DatasetBase, LkmError, PickMode, _records
from .rng import Xoshiro256Plus


class GmmError(Exception):
    """GmmError (gaussian_mixture/errors.rs)."""

    def __init__(self, kind, message=""):
        super().__init__(f"{kind}: {message}" if message else kind)
        self.kind = kind


class GmmInitMethod:
    KMeans = 0
    Random = 1


class GmmCovarType:
    Full = "Full"


_GMM_ERRORS = {40: "InvalidValue", 41: "InvalidValue", 42: "InvalidValue", 43: "InvalidValue", 44: "InvalidValue",
               45: "EmptyCluster", 46: "LinalgError", 47: "NotConverged", 48: "LowerBoundError"}


class GmmParams:
    """GmmParams / GmmValidParams (gaussian_mixture/hyperparams.rs); defaults :88-101."""

    def __init__(self, n_clusters, rng=None):
        self._n_clusters = n_clusters
        self._covar_type = GmmCovarType.Full
        self._tolerance = 1e-3
        self._reg_covar = 1e-6
        self._n_runs = 1
        self._max_n_iter = 100
        self._init_method = GmmInitMethod.KMeans
        self._rng = rng if rng is not None else Xoshiro256Plus.seed_from_u64(42)
        self._pick_mode = PickMode.Exact

    def _set(self, name, v):
        q = copy.copy(self)
        setattr(q, name, v)
        return q

    def tolerance(self, v):
        return self._set("_tolerance", v)

    def reg_covariance(self, v):
        return self._set("_reg_covar", v)

    def n_runs(self, v):
        return self._set("_n_runs", v)

    def max_n_iterations(self, v):
        return self._set("_max_n_iter", v)

    def init_method(self, v):
        return self._set("_init_method", v)

    def with_rng(self, rng):
        return self._set("_rng", rng)

    def kmeans_pick_mode(self, v):   # not in the reference: how the K-Means initialisation walks its prefix sums
        return self._set("_pick_mode", v)

    def check(self):
        """ParamGuard::check_ref (hyperparams.rs:167-193)."""
        if self._n_clusters == 0:
            raise GmmError("InvalidValue", "`n_clusters` cannot be 0!")
        if self._tolerance <= 0:
            raise GmmError("InvalidValue", "`tolerance` must be greater than 0!")
        if self._reg_covar < 0:
            raise GmmError("InvalidValue", "`reg_covar` must be positive!")
        if self._n_runs == 0:
            raise GmmError("InvalidValue", "`n_runs` cannot be 0!")
        if self._max_n_iter == 0:
            raise GmmError("InvalidValue", "`max_n_iterations` cannot be 0!")
        return self

    def fit(self, dataset, ctx=None):
        """Fit::fit (gaussian_mixture/algorithm.rs:411-467)."""
        x = np.asarray(_records(dataset))
        ctx = ctx or Context.default()
        ctx.set_data(x)
        dtype = ctx.dtype
        sfx, ct = _ffi.sfx_of(dtype)
        k, d = self._n_clusters, x.shape[1]
        p = _ffi.lkm_gmm_params(k, float(self._tolerance), float(self._reg_covar), self._n_runs, self._max_n_iter,
                                self._init_method, self._pick_mode)
        kk = max(k, 1)
        w, mu = np.zeros(kk, dtype), np.zeros((kk, d), dtype)
        cov, prec, pchol = (np.zeros((kk, d, d), dtype) for _ in range(3))
        n_iter, lb = C.c_uint64(0), np.zeros(1, dtype)
        rng = copy.deepcopy(self._rng)   # `self.rng()` hands out a clone (hyperparams.rs:83-85)
        st = getattr(_ffi.lib, f"lkm_gmm_fit_{sfx}")(ctx.h, C.byref(p), _ffi.make_rng(rng), _ffi.fptr(w, ct), _ffi.fptr(mu, ct),
                                                     _ffi.fptr(cov, ct), _ffi.fptr(prec, ct), _ffi.fptr(pchol, ct), C.byref(n_iter),
                                                     _ffi.fptr(lb, ct))
        if st in _GMM_ERRORS:
            raise GmmError(_GMM_ERRORS[st], ctx.error())
        ctx.check(st)
        return GaussianMixtureModel(w[:k], mu[:k], cov[:k], prec[:k], pchol[:k], ctx, int(n_iter.value), lb[0])


class GaussianMixtureModel:
    """GaussianMixtureModel<F> (gaussian_mixture/algorithm.rs:102-110)."""

    def __init__(self, weights, means, covariances, precisions, precisions_chol, ctx=None, n_iter=0, lower_bound=None):
        self._w, self._mu, self._cov, self._prec, self._pchol = weights, means, covariances, precisions, precisions_chol
        self._ctx, self.n_iter, self.lower_bound = ctx, n_iter, lower_bound

    @staticmethod
    def params(n_clusters):
        return GmmParams(n_clusters)

    @staticmethod
    def params_with_rng(n_clusters, rng):
        return GmmParams(n_clusters, rng)

    def weights(self):
        return self._w

    def means(self):
        return self._mu

    def covariances(self):
        return self._cov

    def precisions(self):
        return self._prec

    def centroids(self):
        return self._mu

    def _run(self, x, want_labels, want_proba):
        x = np.ascontiguousarray(x, dtype=self._mu.dtype)
        sfx, ct = _ffi.sfx_of(x.dtype)
        n, d = x.shape
        k = self._mu.shape[0]
        labels = np.zeros(n, dtype=np.uint64) if want_labels else None
        proba = np.zeros((n, k), dtype=x.dtype) if want_proba else None
        lpn = np.zeros(1, dtype=x.dtype)
        ctx = self._ctx or Context.default()
        st = getattr(_ffi.lib, f"lkm_gmm_predict_{sfx}")(
            ctx.h, k, _ffi.fptr(np.ascontiguousarray(self._w), ct), _ffi.fptr(np.ascontiguousarray(self._mu), ct),
            _ffi.fptr(np.ascontiguousarray(self._pchol), ct), _ffi.fptr(x, ct), n, d, d,
            labels.ctypes.data_as(_ffi._u64p) if want_labels else None, _ffi.fptr(proba, ct), _ffi.fptr(lpn, ct))
        ctx.check(st)
        return labels, proba, lpn[0]

    def predict(self, x):
        """PredictInplace (:470-484): argmax of the responsibilities."""
        if isinstance(x, DatasetBase):
            return DatasetBase(x.records, self._run(np.asarray(x.records), True, False)[0])
        return self._run(x, True, False)[0]

    def predict_proba(self, observations):
        """:205-208"""
        return self._run(observations, False, True)[1]

    def score(self, observations):
        """mean log-likelihood of the observations (the E-step's lower bound, :291-298)"""
        return self._run(observations, False, False)[2]

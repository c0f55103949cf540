"""linfa_b200 — B200-native K-Means hot path behind linfa-clustering's API surface.

`KMeans.params(k)...fit(dataset)`, `predict`, `transform` as in linfa-clustering; the
arithmetic runs in hand-written sm_100a CUDA kernels behind the C ABI of include/lkm.h.
"""
from .kmeans import (Context, Dataset, DatasetBase, DeviceGroup, IncrKMeansError, KMeans, KMeansAlgorithm, KMeansError,  # noqa: F401
                     KMeansInit, KMeansParams, KMeansParamsError, KMeansValidParams, L2Dist, LkmError, PickMode, generate,
                     read_npy, write_npy)
from .rng import Xoshiro256Plus  # noqa: F401
from .gmm import GaussianMixtureModel, GmmCovarType, GmmError, GmmInitMethod, GmmParams  # noqa: F401,E402

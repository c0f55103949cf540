"""CPU-only checks of the boundary: the C-ABI library loads and exports every symbol
include/lkm.h declares, the host-side mirror validates parameters like the reference,
and the binding's RNG matches the oracle's."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lb():
    import __graft_entry__ as g

    g.build()
    import linfa_b200

    return linfa_b200


def test_library_exports_every_declared_symbol(lb):
    hdr = open(os.path.join(ROOT, "include", "lkm.h")).read()
    declared = set(re.findall(r"^(?:lkm_status|const char\*|void\*|int|lkm_ctx\*)\s+(lkm_[a-z0-9_]+)\(", hdr, flags=re.M))
    assert len(declared) >= 30
    lib = C.CDLL(os.path.join(ROOT, "linfa_b200", "liblkm.so"))
    missing = [s for s in sorted(declared) if not hasattr(lib, s)]
    assert not missing, missing
    from linfa_b200 import _ffi

    listed = set(_ffi.UNTYPED_SYMBOLS) | {f"{s}_{t}" for s in _ffi.TYPED_SYMBOLS for t in ("f32", "f64")}
    assert listed == declared


def test_version_string(lb):
    assert b"sm_100a" in lb._ffi.lib.lkm_version()


def test_struct_layouts_match_header(lb):
    from linfa_b200 import _ffi

    assert C.sizeof(_ffi.lkm_params) == 72
    assert C.sizeof(_ffi.lkm_rng) == 24
    assert C.sizeof(_ffi.lkm_stats) == 72


def test_no_device_fails_loudly(lb):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(lb.LkmError):
        lb.Context(0)


def test_param_validation_order_and_errors(lb):
    # hyperparams.rs:204-232
    P = lb.KMeans.params
    for params, kind in [(P(0), "NClusters"), (P(1).n_runs(0), "NRuns"), (P(1).tolerance(0.0), "Tolerance"),
                         (P(1).tolerance(-1.0), "Tolerance"), (P(1).max_n_iterations(0), "MaxIterations"),
                         (P(0).n_runs(0).tolerance(0.0), "NClusters"), (P(1).n_runs(0).tolerance(0.0), "NRuns")]:
        with pytest.raises(lb.KMeansParamsError) as e:
            params.check()
        assert e.value.kind == kind
        with pytest.raises(lb.KMeansError) as e2:
            params.fit(np.zeros((4, 2)))
        assert e2.value.kind == "InvalidParams" and e2.value.inner.kind == kind


def test_defaults_match_reference(lb):
    # hyperparams.rs:71-82, algorithm.rs:211
    v = lb.KMeans.params(7).check()
    assert (v.n_clusters(), v.n_runs(), v.tolerance(), v.max_n_iterations()) == (7, 10, 1e-4, 300)
    assert v.init_method() == lb.KMeansInit.KMeansPlusPlus and v.algorithm() == lb.KMeansAlgorithm.Lloyd
    assert v.rng().s == lb.Xoshiro256Plus.seed_from_u64(42).s
    assert isinstance(v.dist_fn(), lb.L2Dist)


def test_fit_with_rejects_hamerly(lb):
    # algorithm.rs:1204-1218
    p = (lb.KMeans.params_with_rng(2, lb.Xoshiro256Plus.seed_from_u64(45)).algorithm(lb.KMeansAlgorithm.Hamerly)
         .init_method(lb.KMeansInit.Precomputed([[0.0, 0.0], [10.0, 10.0]])))
    with pytest.raises(lb.IncrKMeansError) as e:
        p.fit_with(None, lb.DatasetBase(np.array([[1.0, 1.0], [11.0, 11.0]])))
    assert e.value.kind == "InvalidParams" and e.value.inner.kind == "IncrementalHamerly"


def test_python_xoshiro_matches_oracle(lb, oracle):
    for seed in (0, 40, 42, 2**63 + 5):
        a, b = lb.Xoshiro256Plus.seed_from_u64(seed), oracle.rng(seed)
        assert a.s == [int(v) for v in b.state]
        assert [a.next_u64() for _ in range(50)] == [b.next_u64() for _ in range(50)]
        assert a.next_u32() == b.next_u32()

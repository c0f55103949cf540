"""Writes the committed golden fixtures of tests/golden/.

  reference_kats.json   literal inputs / expected outputs copied from the reference's own unit tests
                        (algorithms/linfa-clustering/src/k_means/*.rs; each entry cites file:line).  These pin the
                        oracle; the reference itself (Rust) cannot be built in this image.
  oracle_fits.npz       seeded inputs and the outputs of the CPU oracle (oracle/kmeans_oracle.cpp) for a few
                        small fits.  They freeze the oracle's behaviour (summation order, RNG stream) so that a
                        change to the oracle shows up as a fixture diff, and give the GPU tests committed targets.

Run from the repository root:  python tests/golden/make_golden.py
"""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests import oracle_lib  # noqa: E402

KM = "algorithms/linfa-clustering/src/k_means"

KATS = {
    "min_sq_dists": {"src": f"{KM}/algorithm.rs:1003-1013 (test_min_dists)",
                     "centroids": [[0, 1], [40, 10]], "observations": [[3, 4], [1, 3], [25, 15]], "expected": [18, 5, 250]},
    "closest_centroid": {"src": f"{KM}/algorithm.rs:1150-1160 (oracle_test_for_closest_centroid)",
                         "centroids": [[0, 0], [1, 2], [20, 0], [0, 20]],
                         "observations": [[1, 0.6], [20, 2], [20, 0], [7, 20]], "expected": [0, 2, 2, 3]},
    "compute_extra_centroids": {"src": f"{KM}/algorithm.rs:1117-1125 (test_compute_extra_centroids)",
                                "observations": [[1, 2]], "memberships": [0], "old_centroids": [[1, 1], [1, 1]],
                                "expected": [[1, 1.5], [1, 1]]},
    "compute_inertia": {"src": f"{KM}/algorithm.rs:1512-1520 (test_compute_inertia)",
                        "observations": [[0, 0], [3, 4]], "memberships": [0, 0], "centroids": [[1, 1]], "expected": 15},
    "tolerance": {"src": f"{KM}/algorithm.rs:1220-1231 (test_tolerance)", "observations": [[1, 1], [11, 11]],
                  "init": [[0, 0]], "tolerance": 8.5, "expected_centroids": [[4, 4]], "expected_n_iter": 1},
    "membership_counts": {"src": f"{KM}/init.rs:326-332 (test_cluster_membership_counts)",
                          "centroids": [[0, 1], [40, 10], [3, 9]], "observations": [[3, 4], [1, 3], [25, 15]],
                          "expected": [2, 1, 0]},
    "xoshiro256plus_state_1234": {"src": "rand_xoshiro 0.6 xoshiro256plus reference vector (SURVEY A.8)",
                                  "state": [1, 2, 3, 4],
                                  "expected": [5, 211106232532999, 211106635186183, 9223759065350669058,
                                               9250833439874351877, 13862484359527728515, 2346507365006083650,
                                               1168864526675804870, 34095955243042024, 3466914240207415127]},
    "seed_from_u64_42": {"src": "rand_core 0.6 SeedableRng::seed_from_u64 (SplitMix64), default rng of KM/algorithm.rs:211",
                         "expected_state": ["0xbdd732262feb6e95", "0x28efe333b266f103", "0x47526757130f9f52",
                                            "0x581ce1ff0e4ae394"],
                         "expected_first_u64": [1581911519303979561, 5726079574540882823, 1154208747244521758,
                                                5653213587482834094]},
    "derived_six_points": {"src": f"data of {KM}/algorithm.rs:1586-1596; expected values derived from SURVEY A.1/A.5 in f64",
                           "observations": [[0, 0], [1, 0], [0, 1], [10, 10], [11, 10], [10, 11]],
                           "init": [[0, 0], [10, 10]], "tolerance": 1e-4, "expected_n_iter": 8,
                           "expected_centroids": [[0.33332825, 0.33332825], [10.33332825, 10.33332825]],
                           "expected_labels": [0, 0, 0, 1, 1, 1], "expected_inertia": 0.4444444453},
}


def blobs(n, d, k_true, dt, seed, spread):
    g = np.random.default_rng(seed)
    centres = g.uniform(-spread, spread, (k_true, d))
    m = n // k_true
    parts = [g.standard_normal((m if b < k_true - 1 else n - m * (k_true - 1), d)) + centres[b] for b in range(k_true)]
    return np.concatenate(parts).astype(dt)


FIT_CASES = [
    # name, n, d, k, k_true, dtype, spread, init method, rng seed, n_runs, max_iter, tol
    ("f32_d32_k8_pp", 640, 32, 8, 8, np.float32, 30.0, oracle_lib.INIT_KMEANSPP, 3, 2, 30, 1e-4),
    ("f32_d32_k20_random", 900, 32, 20, 10, np.float32, 4.0, oracle_lib.INIT_RANDOM, 11, 1, 25, 1e-3),
    ("f64_d8_k16_pp", 1200, 8, 16, 16, np.float64, 30.0, oracle_lib.INIT_KMEANSPP, 5, 2, 40, 1e-6),
    ("f64_d2_k3_random", 3000, 2, 3, 3, np.float64, 10.0, oracle_lib.INIT_RANDOM, 40, 3, 100, 1e-4),
    ("f32_d5_k4_pp", 999, 5, 4, 4, np.float32, 8.0, oracle_lib.INIT_KMEANSPP, 1, 1, 50, 1e-4),
]


def main():
    with open(os.path.join(HERE, "reference_kats.json"), "w") as f:
        json.dump(KATS, f, indent=1)
    oracle_lib.build()
    orc = oracle_lib.load()
    out = {}
    for name, n, d, k, k_true, dt, spread, method, seed, n_runs, max_iter, tol in FIT_CASES:
        X = blobs(n, d, k_true, dt, seed=1000 + seed, spread=spread)
        r = orc.fit(X, k, orc.rng(seed), n_runs=n_runs, max_iter=max_iter, tol=tol, method=method, acc_f64=True)
        assert r["status"] == 0, name
        labels, dists = orc.assign(X, r["centroids"])
        out[name + "/X"] = X
        out[name + "/params"] = np.array([k, seed, n_runs, max_iter, method], dtype=np.int64)
        out[name + "/tol"] = np.array([tol])
        out[name + "/centroids"] = r["centroids"]
        out[name + "/inertia"] = np.array([r["inertia"]], dtype=np.float64)
        out[name + "/cluster_count"] = np.asarray(r["cluster_count"], dtype=np.float64)
        out[name + "/labels"] = labels.astype(np.uint32)
        out[name + "/min_sq_dists"] = dists
    np.savez_compressed(os.path.join(HERE, "oracle_fits.npz"), **out)
    print("wrote", len(KATS), "reference KATs and", len(FIT_CASES), "oracle fit fixtures")


if __name__ == "__main__":
    main()

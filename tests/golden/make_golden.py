"""Generates tests/golden/*.npz (run once, here, where scipy is installed; outputs are committed).

Why scipy: the reference's own implementation of this path is absent from /root/reference
(raft::distance and pylibraft.distance were removed upstream in 26.02, CHANGELOG.md:59), so the
reference itself cannot be imported to generate vectors.  The historical pylibraft tests compared
against scipy.spatial.distance.cdist ([RECALLED]) and BASELINE.json configs[0]
names the same check, so cdist output on seeded inputs is what is pinned here, together with the
known-answer vectors that DO survive in the reference's tests:
  cpp/tests/matrix/argmin.cu:71-75            row-argmin {0,3,3}
  cpp/tests/core/operators_host.cpp:134-152   argmin_op tie-break cases
"""
import os
import numpy as np
from scipy.spatial.distance import cdist

HERE = os.path.dirname(os.path.abspath(__file__))
SCIPY = {"L2Expanded": "sqeuclidean", "L2SqrtExpanded": "euclidean", "CosineExpanded": "cosine",
         "L1": "cityblock", "L2Unexpanded": "sqeuclidean", "L2SqrtUnexpanded": "euclidean",
         "Linf": "chebyshev", "Canberra": "canberra", "CorrelationExpanded": "correlation"}


def blobs(rows, cols, seed, centers=None):
    rng = np.random.default_rng(seed)
    if centers is None:
        centers = rng.uniform(-10, 10, size=(5, cols))
    lab = rng.integers(0, 5, size=rows)
    return (centers[lab] + rng.standard_normal((rows, cols))).astype(np.float32), centers


def main():
    out = {}
    for name, (m, n, k) in {"small": (37, 29, 19), "cfg1": (128, 96, 32)}.items():
        x, c = blobs(m, k, 1234)
        y, _ = blobs(n, k, 4321, c)
        out[f"{name}_x"], out[f"{name}_y"] = x, y
        x64, y64 = x.astype(np.float64), y.astype(np.float64)
        for metric, sp in SCIPY.items():
            out[f"{name}_{metric}"] = cdist(x64, y64, sp)
        out[f"{name}_LpUnexpanded_p3"] = cdist(x64, y64, "minkowski", p=3.0)
        out[f"{name}_InnerProduct"] = x64 @ y64.T
        d = cdist(x64, y64, "sqeuclidean")
        out[f"{name}_nn_idx"] = np.argmin(d, axis=1).astype(np.int32)
        out[f"{name}_nn_val"] = d.min(axis=1)
    # reference known answers
    out["ref_argmin_in"] = np.array([0.1, 0.2, 0.3, 0.4, 0.4, 0.3, 0.2, 0.1, 0.2, 0.3, 0.5, 0.0],
                                    dtype=np.float32).reshape(3, 4)   # argmin.cu:71-72
    out["ref_argmin_out"] = np.array([0, 3, 3], dtype=np.int32)
    np.savez_compressed(os.path.join(HERE, "distance_golden.npz"), **out)
    print("wrote", os.path.join(HERE, "distance_golden.npz"), len(out), "arrays")


if __name__ == "__main__":
    main()

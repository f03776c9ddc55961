"""Generates the committed golden vectors (run here, in the container that has /root/reference):

    python tests/golden/make_golden.py

Inputs are regenerated from seeds by tests/golden/cases.py; only the expected OUTPUTS are stored.
Expected assignments come from the C oracle (oracle/kmcuda_oracle.c), which the GPU suite separately
pins bit-for-bit against the unmodified reference rebuilt for sm_100 (tests/test_parity_gpu.py::
test_oracle_matches_reference); expected neighbours come from scikit-learn, the reference's own pin
(reference src/test.py:598-606).
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)

from oracle import oracle as O  # noqa: E402
import cases  # noqa: E402


def main():
    out = {}
    for name, (n, d, k, seed, kind) in cases.ASSIGN_CASES.items():
        X, C = cases.make_assign_case(n, d, k, seed, kind)
        a, _, changed = O.assign_lloyd(X, C)
        _, best, second = O.assign_truth(X, C)
        out["assign/" + name] = a
        out["assign_ties/" + name] = O.tie_exempt(best, second)
        print(name, "changed", changed, "ties", int(out["assign_ties/" + name].sum()))
    X = cases.blobs()
    from sklearn.neighbors import NearestNeighbors
    nb = NearestNeighbors(n_neighbors=11).fit(X)
    out["knn/blobs_k10"] = nb.kneighbors(X)[1][:, 1:].astype(np.uint32)
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote", os.path.join(HERE, "golden.npz"), os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes")


if __name__ == "__main__":
    main()

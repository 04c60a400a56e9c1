#!/usr/bin/env python
"""A/B check of the two 10-NN + covariance kernels of the GICP path: the cell-centric kernel (GFS_GICP_KNN_CELLS=1)
must give bit-identical covariances and alignment results to the per-query shell search (default)."""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoflowslam_b200 import RegistrationGICP, synth  # noqa: E402


def run(per_query, pairs):
    if per_query:
        os.environ.pop("GFS_GICP_KNN_CELLS", None)
    else:
        os.environ["GFS_GICP_KNN_CELLS"] = "1"
    out = []
    for tgt, src, _ in pairs:
        reg = RegistrationGICP(max_points=max(len(tgt), len(src)), max_pairs=1)
        t0 = time.perf_counter()
        r = reg.RegisterPointClouds(tgt, src)
        dt = time.perf_counter() - t0
        out.append((r, reg.cloud(0), reg.cloud(1), reg.knn_stats(0), reg.knn_stats(1), dt))
        reg.close()
    return out


def main():
    pairs = [synth.gicp_pair(2000 + i, n_target=50000) for i in range(3)] + [synth.gicp_pair(2100, n_target=3000)]
    a, b = run(False, pairs), run(True, pairs)
    for i, (x, y) in enumerate(zip(a, b)):
        same = all(np.array_equal(x[k][1], y[k][1]) and np.array_equal(x[k][0], y[k][0]) for k in (1, 2))
        same = same and np.array_equal(x[0]["T"], y[0]["T"]) and x[0]["iterations"] == y[0]["iterations"]
        print("pair %d: points %d/%d cells/per-query %s %s identical=%s  %.1f ms vs %.1f ms" %
              (i, len(x[1][0]), len(x[2][0]), x[3], x[4], same, 1e3 * x[5], 1e3 * y[5]))
        assert same
    print("A/B ok")


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""CPU study for the next GICP step: how often would a float32 prefilter with CELL-LOCAL coordinates
leave the exact fp64 selection ambiguous, and how many candidates do the 1-NN / 10-NN searches look at?  Pure numpy /
scipy on the synthetic configs[2] clouds; nothing here runs on the GPU or feeds the product.
  python scripts/gicp_prefilter_study.py [seed]"""
import os
import sys

import numpy as np
from scipy.spatial import cKDTree

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoflowslam_b200 import synth  # noqa: E402
from oracle import oracle as O  # noqa: E402

CELL, VOXEL, MAXD, K = 0.1, 0.02, 0.1, 10


def cell_local_f32_d2(P, Q):
    """squared distances |P - Q|^2 the way the prefilter would compute them: both points relative to the origin of P's
    grid cell (fp64 subtraction, rounded to float32), differences / products / sums in float32"""
    origin = np.floor(P / CELL) * CELL
    p = (P - origin).astype(np.float32); q = (Q - origin).astype(np.float32)
    d = p - q
    return (d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]) + d[:, 2] * d[:, 2]


def main():
    seed = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
    tgt, src, T_true = synth.gicp_pair(seed, n_target=50000)
    T = O.voxelgrid(tgt, VOXEL); S = O.voxelgrid(src, VOXEL)
    tree = cKDTree(T)
    print("clouds: %d / %d points after the %.2f m voxel filter" % (len(T), len(S), VOXEL))
    # ---- 1-NN (k_nn_corr): best and second best
    d, idx = tree.query(S, k=2)
    ok = d[:, 0] <= MAXD
    d2_64 = (d[ok] ** 2)
    f1 = cell_local_f32_d2(T[idx[ok, 0]], S[ok]); f2 = cell_local_f32_d2(T[idx[ok, 1]], S[ok])
    err = np.abs(f1.astype(np.float64) - d2_64[:, 0])
    A = 1.2 * 2 * np.sqrt(3.0) * 2.0 ** -24 * (2 * CELL + MAXD)  # guaranteed bound: A sqrt(v) + 5e-7 v + 1e-14
    bound = lambda v: A * np.sqrt(np.maximum(v, 0)) + 5e-7 * v + 1e-14
    tol = bound(f1.astype(np.float64))
    print("1-NN: float32 cell-local d^2 error: max %.2e, max error / bound %.3f (must be < 1)" % (err.max(), (err / tol).max()))
    ambiguous = (f2.astype(np.float64) - bound(f2.astype(np.float64))) <= (f1.astype(np.float64) + tol)
    print("1-NN: queries whose runner-up falls inside the bound (need the exact fp64 comparison): %d of %d = %.4f %%; "
          "warps of 32 consecutive queries with at least one: %.2f %%" %
          (ambiguous.sum(), ok.sum(), 100.0 * ambiguous.mean(),
           100.0 * np.mean([ambiguous[i:i + 32].any() for i in range(0, len(ambiguous), 32)])))
    # emulation of the decision the kernel makes (`k_nn_corr_f32` on the git branch r2-gicp-f32-prefilter: compiled and emulated here, not yet run on a GPU): best / runner-up by
    # float32 value among the 8 exact nearest (a superset of everything that matters), clear winner <=> runner-up's lower
    # bound above the best's upper bound; every clear winner must be the exact nearest neighbour
    d8, i8 = tree.query(S[ok], k=8)
    v8 = np.stack([cell_local_f32_d2(T[i8[:, j]], S[ok]) for j in range(8)], 1).astype(np.float64)
    order = np.argsort(v8, 1, kind="stable")
    b1 = np.take_along_axis(v8, order[:, :1], 1)[:, 0]; b2 = np.take_along_axis(v8, order[:, 1:2], 1)[:, 0]
    id1 = np.take_along_axis(i8, order[:, :1], 1)[:, 0]
    clear = (b2 - bound(b2)) > (b1 + bound(b1))
    wrong = clear & (id1 != i8[:, 0])
    print("1-NN decision emulation: clear winners %.4f %%, of which not the exact nearest neighbour: %d (must be 0)" %
          (100.0 * clear.mean(), int(wrong.sum())))
    # candidates a 27-cell scan looks at vs. what a search pruned by the final distance needs
    cells = np.floor(T / CELL).astype(np.int64)
    key = (cells[:, 0] + 512) | ((cells[:, 1] + 512) << 20) | ((cells[:, 2] + 512) << 40)
    uniq, cnt = np.unique(key, return_counts=True)
    lut = dict(zip(uniq.tolist(), cnt.tolist()))
    qs = S[ok][::16]; ds = d[ok][::16, 0]
    qc = np.floor(qs / CELL).astype(np.int64)
    full, pruned = [], []
    for (cx, cy, cz), q, r in zip(qc, qs, ds):
        n27 = npr = 0
        for dz in (-1, 0, 1):
            for dy in (-1, 0, 1):
                for dx in (-1, 0, 1):
                    c = lut.get(int((cx + dx + 512) | ((cy + dy + 512) << 20) | ((cz + dz + 512) << 40)), 0)
                    if not c:
                        continue
                    n27 += c
                    lo = np.array([cx + dx, cy + dy, cz + dz]) * CELL
                    gap = np.maximum(np.maximum(lo - q, q - (lo + CELL)), 0.0)
                    if (gap ** 2).sum() <= r * r:
                        npr += c
        full.append(n27); pruned.append(npr)
    print("1-NN: records in the 3x3x3 block: mean %.1f; in the cells within the final NN distance: mean %.1f (median %d)" %
          (np.mean(full), np.mean(pruned), int(np.median(pruned))))
    # the same count for finer 1-NN grids: occupied cells probed / records read when only the cells within the final
    # NN distance of the query are visited (what a search seeded with the previous correspondence converges to)
    Sa = S @ np.asarray(T_true)[:3, :3].T + np.asarray(T_true)[:3, 3]       # the source cloud at the true alignment
    da, _ = tree.query(Sa, k=1)
    oka = da <= MAXD
    print("1-NN distance: first iteration median %.4f m, at the true alignment median %.4f m" % (np.median(d[ok][:, 0]), np.median(da[oka])))
    for label, qs_, ds_ in (("first iteration", qs[::4], ds[::4]), ("aligned", Sa[oka][::64], da[oka][::64])):
        for g in (0.1, 0.05, 0.03):
            cg = np.floor(T / g).astype(np.int64)
            kg = (cg[:, 0] + 4096) | ((cg[:, 1] + 4096) << 20) | ((cg[:, 2] + 4096) << 40)
            u, c = np.unique(kg, return_counts=True)
            lg = dict(zip(u.tolist(), c.tolist()))
            probes, recs = [], []
            for q, r in zip(qs_, ds_):
                lo_c = np.floor((q - r) / g).astype(np.int64); hi_c = np.floor((q + r) / g).astype(np.int64)
                np_, nr = 0, 0
                for z in range(lo_c[2], hi_c[2] + 1):
                    for y in range(lo_c[1], hi_c[1] + 1):
                        for x in range(lo_c[0], hi_c[0] + 1):
                            lo = np.array([x, y, z]) * g
                            gap = np.maximum(np.maximum(lo - q, q - (lo + g)), 0.0)
                            if (gap ** 2).sum() > r * r:
                                continue
                            np_ += 1
                            nr += lg.get(int((x + 4096) | ((y + 4096) << 20) | ((z + 4096) << 40)), 0)
                probes.append(np_); recs.append(nr)
            print("1-NN grid %.3f m, %s: hash probes per query mean %.1f, records read mean %.1f" % (g, label, np.mean(probes), np.mean(recs)))
    # ---- 10-NN (k_knn_cov): list length if the radius comes from float32 upper bounds
    dk, ik = tree.query(T[::8], k=K + 7)
    q = T[::8]
    up = np.stack([cell_local_f32_d2(T[ik[:, j]], q) for j in range(K + 7)], 1).astype(np.float64)
    up = up + bound(up)
    rho = np.sort(up, 1)[:, K - 1]                               # k-th smallest upper bound >= true k-th distance
    n_in = ((dk ** 2) <= rho[:, None]).sum(1)
    print("10-NN: exact candidates inside the float32-derived radius: mean %.3f, max %d (list capacity 16); queries with more "
          "than %d: %.4f %%" % (n_in.mean(), n_in.max(), K, 100.0 * np.mean(n_in > K)))
    print("10-NN: k-th neighbour distance: median %.4f m, 99th percentile %.4f m (cell %.2f m)" %
          (np.median(dk[:, K - 1]), np.percentile(dk[:, K - 1], 99), CELL))


if __name__ == "__main__":
    main()

"""N > 1 host logic on CPU: two gloo ranks.  (a) frame sharding covers the batch exactly once and the
whole-job aggregate is the sum over ranks; (b) the landmark-partitioned BA normal equations: the sum
over ranks of the per-shard Schur-reduced systems equals the single-shard system."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from geoflowslam_b200 import synth
    from geoflowslam_b200.parallel import landmark_owner, shard_range
    from oracle import oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (a) sharding
        n = 1027
        s, e = shard_range(n, rank, world)
        cover = torch.zeros(n, dtype=torch.int32)
        cover[s:e] = 1
        dist.all_reduce(cover)
        frames_done = torch.tensor([float(e - s)])
        dist.all_reduce(frames_done)
        ok_a = bool((cover == 1).all()) and frames_done.item() == n
        # (b) partitioned BA system
        p = synth.ba_problem(seed=3020, n_kf=6, n_points=150)
        Hpp, bp, Hll, bl, Hpl = O.ba_system(p)
        nP = len(bp)
        lam = 1e-2

        def reduced(owned):
            Hs = np.zeros((nP, nP)); bs = np.zeros(nP)
            for j in range(p["n_points"]):
                if not owned(j):
                    continue
                Dinv = np.linalg.inv(Hll[j] + lam * np.eye(3))
                es = [e for e in range(p["n_obs"]) if p["obs_pt"][e] == j and p["obs_kf"][e] < p["n_opt_kf"]]
                for e1 in es:
                    o1 = 15 * int(p["obs_kf"][e1])
                    bs[o1:o1 + 6] -= Hpl[e1] @ Dinv @ bl[j]
                    for e2 in es:
                        o2 = 15 * int(p["obs_kf"][e2])
                        Hs[o1:o1 + 6, o2:o2 + 6] -= Hpl[e1] @ Dinv @ Hpl[e2].T
            return Hs, bs

        # the pose-side terms of an edge live with its landmark's owner; inertial terms with rank 0.
        # Hpp is linear in the edges, so split it as (inertial part on rank 0) + (visual part by owner):
        q_vis = dict(p); q_vis["n_inertial"] = 0
        Hpp_vis, bp_vis = O.ba_system(q_vis)[:2]
        mine = dict(p)
        keep = np.array([landmark_owner(int(j), world) == rank for j in p["obs_pt"]])
        for k in ("obs_kf", "obs_pt", "obs_uvr", "obs_inv_sigma2"):
            mine[k] = p[k][keep]
        mine["n_obs"] = int(keep.sum()); mine["n_inertial"] = 0
        Hpp_mine, bp_mine = O.ba_system(mine)[:2]
        Hs, bs = reduced(lambda j: landmark_owner(j, world) == rank)
        Hs += Hpp_mine; bs += bp_mine
        if rank == 0:
            Hs += (Hpp - Hpp_vis) + lam * np.eye(nP); bs += bp - bp_vis
        t = torch.from_numpy(np.concatenate([Hs.ravel(), bs]))
        dist.all_reduce(t)
        Hs_sum = t[:nP * nP].numpy().reshape(nP, nP); bs_sum = t[nP * nP:].numpy()
        Hs_full, bs_full = reduced(lambda j: True)
        Hs_full += Hpp + lam * np.eye(nP); bs_full += bp
        ok_b = np.allclose(Hs_sum, Hs_full, rtol=1e-10, atol=1e-8) and np.allclose(bs_sum, bs_full, rtol=1e-10, atol=1e-8)
        # and the reduced solve reproduces the oracle's pose step
        okx, x = O.ba_step(p, lam)
        xp = np.linalg.solve(Hs_sum, bs_sum)
        ok_c = okx and np.allclose(xp, x[:nP], rtol=1e-6, atol=1e-9)
        q.put((rank, ok_a, ok_b, ok_c))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_rank_gloo_sharding_and_partitioned_ba():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=240) for _ in range(2)]
    for pr in procs:
        pr.join(timeout=60)
    assert sorted(r[0] for r in res) == [0, 1]
    for rank, a, b, c in res:
        assert a, "sharding"
        assert b, "partitioned system"
        assert c, "reduced solve"


def test_shard_range_properties():
    from geoflowslam_b200.parallel import shard_range
    for n in (0, 1, 7, 8, 1024, 1027):
        for w in (1, 2, 4, 8):
            r = [shard_range(n, k, w) for k in range(w)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(w - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1

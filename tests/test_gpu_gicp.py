"""GPU parity: batched GICP (through the C ABI) vs the CPU oracle.  Integer results equal;
floating point within 1e-4 relative (north_star), in practice ~1e-9."""
import numpy as np
import pytest

from geoflowslam_b200 import synth

pytestmark = pytest.mark.gpu
RTOL = 1e-4


def _check(g, o, tight=True):
    assert g["converged"] == o["converged"] and g["iterations"] == o["iterations"]
    assert g["num_inliers"] == o["num_inliers"] and g["inner_evals"] == o["inner_evals"]
    assert g["n_target"] == o["n_target"] and g["n_source"] == o["n_source"]
    tol = 1e-7 if tight else RTOL
    assert np.allclose(g["T"], o["T"], rtol=tol, atol=tol)
    assert np.allclose(g["H"], o["H"], rtol=tol, atol=tol * np.abs(o["H"]).max())
    assert np.allclose(g["b"], o["b"], rtol=tol, atol=tol * max(np.abs(o["b"]).max(), 1.0))
    assert np.isclose(g["error"], o["error"], rtol=tol, atol=1e-9)


def test_preprocess_stage_parity():
    from geoflowslam_b200 import RegistrationGICP
    from oracle import oracle as O
    tgt, src, _ = synth.gicp_pair(2003, n_target=12000)
    reg = RegistrationGICP(max_points=16384, max_pairs=1)
    reg.RegisterPointClouds(tgt, src)
    for ci, cloud in enumerate((tgt, src)):
        xyz, cov = reg.cloud(ci)
        ref = O.voxelgrid(cloud, 0.02)
        assert len(xyz) == len(ref)
        og, orr = np.lexsort(xyz.T), np.lexsort(ref.T)
        assert np.array_equal(xyz[og], ref[orr])            # per-voxel means are bit-exact (fixed sum order)
        cref = O.covariances(ref, 10)
        assert np.allclose(cov[og], cref[orr], rtol=0, atol=1e-9)


@pytest.mark.parametrize("seed,n", [(2000, 20000), (2004, 50000), (2005, 5000)])
def test_align_matches_oracle(seed, n):
    from geoflowslam_b200 import RegistrationGICP
    from oracle import oracle as O
    tgt, src, T = synth.gicp_pair(seed, n_target=n)
    reg = RegistrationGICP(max_points=65536, max_pairs=1)
    g = reg.RegisterPointClouds(tgt, src)
    o = O.gicp_align(tgt, src)
    _check(g, o)
    if seed == 2000:  # (other seeds may slide along a wall: GICP's own ambiguity, same in the oracle)
        assert np.abs(g["T"] - T).max() < 2e-3


def test_batch_with_ragged_sizes_and_inits():
    from geoflowslam_b200 import RegistrationGICP
    from oracle import oracle as O
    pairs = [synth.gicp_pair(2010 + i, n_target=[8000, 3000, 12000, 500][i]) for i in range(4)]
    stride = max(max(len(t), len(s)) for t, s, _ in pairs)
    P = len(pairs)
    tg = np.zeros((P, stride, 4), np.float32); sr = np.zeros((P, stride, 4), np.float32)
    nt = np.zeros(P, np.int32); ns = np.zeros(P, np.int32)
    T0 = np.tile(np.eye(4), (P, 1, 1))
    for i, (t, s, T) in enumerate(pairs):
        tg[i, :len(t)] = t; sr[i, :len(s)] = s; nt[i] = len(t); ns[i] = len(s)
    T0[1] = pairs[1][2]                                    # start one pair at the true pose
    T0[2][:3, 3] += 0.3                                    # and one far enough to need several iterations
    reg = RegistrationGICP(max_points=stride, max_pairs=P)
    res = reg.align_batch(tg, nt, sr, ns, T0)
    from geoflowslam_b200.gicp import _res_to_dict
    for i, (t, s, _) in enumerate(pairs):
        _check(_res_to_dict(res[i]), O.gicp_align(t, s, T0=T0[i]))


def test_degenerate_inputs():
    from geoflowslam_b200 import RegistrationGICP
    from oracle import oracle as O
    tgt, src, _ = synth.gicp_pair(2002, n_target=3000)
    reg = RegistrationGICP(max_points=8192, max_pairs=1)
    far = src.copy(); far[:, :3] += 50.0
    _check(reg.RegisterPointClouds(tgt, far), O.gicp_align(tgt, far))       # zero inliers, LM stalls
    _check(reg.RegisterPointClouds(tgt, tgt), O.gicp_align(tgt, tgt))       # identical clouds
    few = tgt[:4]
    _check(reg.RegisterPointClouds(few, few), O.gicp_align(few, few))       # < 5 points: identity covariances
    sparse = tgt[::200]                                                     # neighbours beyond the shell limit
    _check(reg.RegisterPointClouds(sparse, sparse + np.float32([0.01, 0, 0, 0])),
           O.gicp_align(sparse, sparse + np.float32([0.01, 0, 0, 0])))
    from geoflowslam_b200 import GfsError
    with pytest.raises(GfsError):
        reg.RegisterPointClouds(np.zeros((10000, 4), np.float32), src)      # beyond max_points


def _pack(clouds):
    stride = max(len(c) for c in clouds)
    a = np.zeros((len(clouds), stride, 4), np.float32); n = np.zeros(len(clouds), np.int32)
    for i, c in enumerate(clouds):
        a[i, :len(c)] = c; n[i] = len(c)
    return a, n


def test_search_variants_are_bit_identical(monkeypatch):
    """The query order (point / grid-cell order), the correspondence kernels (shell walk, ball walk, ball walk over octants, each
    in fp64 or with the float32 prefilter) and the 10-NN kernels (thread per query, warp per cell with octant skipping) are
    different routes to the same exact neighbours: every result field and every covariance is equal bit for bit."""
    from geoflowslam_b200 import RegistrationGICP
    pairs = [synth.gicp_pair(2030 + i, n_target=[20000, 6000, 900][i]) for i in range(3)]
    tg, nt = _pack([p[0] for p in pairs]); sr, ns = _pack([p[1] for p in pairs])
    stride = max(tg.shape[1], sr.shape[1])
    tg = np.pad(tg, ((0, 0), (0, stride - tg.shape[1]), (0, 0))); sr = np.pad(sr, ((0, 0), (0, stride - sr.shape[1]), (0, 0)))
    T0 = np.tile(np.eye(4), (3, 1, 1))
    out = {}
    # (query order, correspondence kernel, 10-NN kernel, grid): grid 1 = the dense sorted grid (only the default kernels use it)
    variants = ((0, 0, 0, 0), (1, 0, 0, 0), (1, 1, 0, 0), (1, 2, 0, 0), (0, 2, 0, 0), (1, 3, 1, 0), (1, 4, 1, 0), (0, 3, 1, 0), (1, 1, 1, 0),
                (1, 5, 0, 0), (1, 6, 0, 0), (0, 5, 1, 0), (1, 7, 0, 0), (0, 7, 0, 0), (1, 7, 1, 0),
                (1, 7, 2, 0), (1, 7, 3, 0), (1, 5, 4, 0), (1, 7, 5, 0), (1, 7, 2, 1), (1, 8, 3, 1), (1, 9, 4, 1))
    for order, nn, knn, grid in variants:
        monkeypatch.setenv("GFS_GICP_ORDER", str(order)); monkeypatch.setenv("GFS_GICP_NN", str(nn)); monkeypatch.setenv("GFS_GICP_KNN", str(knn))
        monkeypatch.setenv("GFS_GICP_GRID", str(grid))
        reg = RegistrationGICP(max_points=stride, max_pairs=3)
        out[(order, nn, knn, grid)] = reg.align_batch(tg, nt, sr, ns, T0).tobytes()
        out[(order, nn, knn, grid, "cov")] = [reg.cloud(c)[1].tobytes() for c in range(6)]
        if knn == 1:   # the warp-per-cell kernel must do the bulk of the work itself, not hand everything over
            cells, handed = reg.knn_stats(0)
            assert handed < 0.25 * nt[0], (cells, handed)
        reg.close()
    base = out[(0, 0, 0, 0)]
    for k, v in out.items():
        if len(k) == 4:
            assert v == base, "variant order=%d nn=%d knn=%d grid=%d differs from the first-generation search" % k
        else:
            assert v == out[(0, 0, 0, 0, "cov")], "covariances of variant order=%d nn=%d knn=%d grid=%d differ" % k[:4]


def test_dense_grid_with_a_clamped_region(monkeypatch):
    """Clouds whose bounding box does not fit the dense grid (outliers tens of metres away; a deliberately tiny grid capacity, so
    that most of the room lies OUTSIDE the 16-cell region and is clamped into its border cells) give the hash grid's bytes:
    results, covariances, for the pairwise call and for tracking mode."""
    from geoflowslam_b200 import RegistrationGICP
    rng = np.random.default_rng(7)
    clouds = []
    for k in range(2):
        t, s_, _ = synth.gicp_pair(2090 + k, n_target=[9000, 4000][k])
        far_t = np.c_[rng.uniform(-60, 60, (25, 3)), np.ones(25)].astype(np.float32)
        far_s = np.c_[rng.uniform(-60, 60, (25, 3)), np.ones(25)].astype(np.float32)
        clouds.append((np.vstack([t, far_t, far_t[:12] + np.float32(0.01)]), np.vstack([s_, far_s, far_t[:5] + np.float32(0.02)])))
    tg, nt = _pack([c[0] for c in clouds]); sr, ns = _pack([c[1] for c in clouds])
    stride = max(tg.shape[1], sr.shape[1])
    tg = np.pad(tg, ((0, 0), (0, stride - tg.shape[1]), (0, 0))); sr = np.pad(sr, ((0, 0), (0, stride - sr.shape[1]), (0, 0)))
    T0 = np.tile(np.eye(4), (2, 1, 1))
    res = {}
    for grid, cap in ((0, None), (1, None), (1, 4096), (1, 64)):
        monkeypatch.setenv("GFS_GICP_GRID", str(grid))
        if cap: monkeypatch.setenv("GFS_GICP_DENSE_CAP", str(cap))
        else: monkeypatch.delenv("GFS_GICP_DENSE_CAP", raising=False)
        reg = RegistrationGICP(max_points=stride, max_pairs=2)
        r = reg.align_batch(tg, nt, sr, ns, T0)
        covs = [reg.cloud(c)[1].tobytes() for c in range(4)]
        reg.track_reset(); reg.track_batch(tg, nt); q = reg.track_batch(sr, ns, T0)
        res[(grid, cap)] = (r.tobytes(), covs, q.tobytes())
        reg.close()
    for k, v in res.items():
        assert v == res[(0, None)], "dense grid (capacity %s) differs from the hash grid" % (k[1],)
    assert res[(0, None)][0] == res[(0, None)][2]


def test_dense_and_sparse_clouds_take_the_hand_over_paths(monkeypatch):
    """Clouds the warp-per-cell 10-NN kernel cannot finish on its own -- a coarse voxel size puts more points into a 3x3x3 block
    than it stages; a very sparse cloud has its neighbours outside the block -- still give the thread-per-query results."""
    from geoflowslam_b200 import RegistrationGICP
    t, s_, _ = synth.gicp_pair(2033, n_target=30000)
    sparse_t, sparse_s = t[::40].copy(), s_[::40].copy()
    for cloud_t, cloud_s, kw in ((t, s_, dict(downsampling_resolution=0.005)), (sparse_t, sparse_s, {})):
        res = {}
        for knn in (0, 1, 2, 3):
            monkeypatch.setenv("GFS_GICP_KNN", str(knn))
            reg = RegistrationGICP(max_points=32768, max_pairs=1, **kw)
            r = reg.RegisterPointClouds(cloud_t, cloud_s)
            res[knn] = (r["T"].tobytes(), r["iterations"], r["num_inliers"], reg.cloud(0)[1].tobytes(), reg.cloud(1)[1].tobytes())
            reg.close()
        assert res[0] == res[1] == res[2] == res[3]


def test_track_mode_equals_pairwise_align():
    """gfs_gicp_track_batch: each cloud preprocessed once, registered against the previous call's -> the same bytes as
    gfs_gicp_align_batch on the pairs (cloud[k-1], cloud[k])."""
    from geoflowslam_b200 import RegistrationGICP
    # two sequences of four clouds each: views of one scene from a slowly moving camera
    seqs = []
    for s in range(2):
        t0, s0, _ = synth.gicp_pair(2040 + s, n_target=[15000, 7000][s])
        t1, s1, _ = synth.gicp_pair(2050 + s, n_target=[15000, 7000][s])
        seqs.append([t0, s0, t0[::2].copy(), s0])   # cloud 2 is a subsample of cloud 0, cloud 3 repeats cloud 1
    K = 4
    packed = [_pack([seqs[s][k] for s in range(2)]) for k in range(K)]
    stride = max(a.shape[1] for a, _ in packed)
    packed = [(np.pad(a, ((0, 0), (0, stride - a.shape[1]), (0, 0))), n) for a, n in packed]
    T0 = np.tile(np.eye(4), (2, 1, 1)); T0[1, :3, 3] += 0.01
    trk = RegistrationGICP(max_points=stride, max_pairs=2)
    ref = RegistrationGICP(max_points=stride, max_pairs=2)
    assert trk.track_batch(*packed[0]) is None
    for k in range(1, K):
        r = trk.track_batch(packed[k][0], packed[k][1], T0)
        e = ref.align_batch(packed[k - 1][0], packed[k - 1][1], packed[k][0], packed[k][1], T0)
        assert r is not None and r.tobytes() == e.tobytes(), "track call %d differs from the pairwise align" % k
    # a reset starts a new chain; align_batch on the same handle also invalidates the stored clouds
    trk.track_reset()
    assert trk.track_batch(*packed[2]) is None
    r = trk.track_batch(packed[3][0], packed[3][1], T0)
    assert r.tobytes() == ref.align_batch(packed[2][0], packed[2][1], packed[3][0], packed[3][1], T0).tobytes()


def test_counts_beyond_stride_are_rejected():
    from geoflowslam_b200 import RegistrationGICP
    from geoflowslam_b200._lib import GfsError
    reg = RegistrationGICP(max_points=4096, max_pairs=1)
    a = np.zeros((1, 1024, 4), np.float32)
    with pytest.raises(GfsError):
        reg.align_batch(a, np.array([2000], np.int32), a, np.array([10], np.int32), np.eye(4)[None])
    with pytest.raises(GfsError):
        reg.align_batch(a, np.array([10], np.int32), a, np.array([-1], np.int32), np.eye(4)[None])


def test_track_mode_edge_cases():
    """Tracking mode with an empty cloud, a tiny cloud (fewer than 10 points: PredictStateICP refuses such frames, the library
    must still not misbehave) and a cloud of identical points in the chain: same bytes as the pairwise align, no crash."""
    from geoflowslam_b200 import RegistrationGICP
    t, s_, _ = synth.gicp_pair(2060, n_target=4000)
    tiny = t[:7].copy()
    same = np.tile(t[:1], (300, 1))
    chain = [t, np.zeros((0, 4), np.float32), s_, tiny, t, same, s_]
    stride = max(len(c) for c in chain)
    T0 = np.eye(4)[None]
    trk = RegistrationGICP(max_points=stride, max_pairs=1)
    ref = RegistrationGICP(max_points=stride, max_pairs=1)

    def pack(c):
        a = np.zeros((1, stride, 4), np.float32); a[0, :len(c)] = c
        return a, np.array([len(c)], np.int32)
    assert trk.track_batch(*pack(chain[0])) is None
    for k in range(1, len(chain)):
        a, n = pack(chain[k]); pa, pn = pack(chain[k - 1])
        r = trk.track_batch(a, n, T0)
        e = ref.align_batch(pa, pn, a, n, T0)
        assert r.tobytes() == e.tobytes(), "chain step %d" % k
        assert np.isfinite(r["T"]).all()


def test_results_are_run_to_run_reproducible():
    """Fixed summation orders everywhere (per-warp partials, last-block reductions): two solves of the same batch give the same
    bytes, also when the batch is processed next to other pairs."""
    from geoflowslam_b200 import RegistrationGICP
    pairs = [synth.gicp_pair(2070 + i, n_target=9000) for i in range(3)]
    tg, nt = _pack([p[0] for p in pairs]); sr, ns = _pack([p[1] for p in pairs])
    stride = max(tg.shape[1], sr.shape[1])
    tg = np.pad(tg, ((0, 0), (0, stride - tg.shape[1]), (0, 0))); sr = np.pad(sr, ((0, 0), (0, stride - sr.shape[1]), (0, 0)))
    T0 = np.tile(np.eye(4), (3, 1, 1))
    reg = RegistrationGICP(max_points=stride, max_pairs=3)
    a = reg.align_batch(tg, nt, sr, ns, T0).tobytes()
    b = reg.align_batch(tg, nt, sr, ns, T0).tobytes()
    assert a == b
    one = RegistrationGICP(max_points=stride, max_pairs=1)
    c = one.align_batch(tg[1:2], nt[1:2], sr[1:2], ns[1:2], T0[1:2]).tobytes()
    assert c == reg.align_batch(tg, nt, sr, ns, T0)[1:2].tobytes()

"""Pins the matching oracle: Hamming + BF against cv2, GMS against a literal numpy restatement."""
import numpy as np
import pytest

from oracle import oracle as O

cv2 = pytest.importorskip("cv2")


def test_descriptor_distance_is_popcount():
    rng = np.random.default_rng(0)
    for _ in range(200):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert O.descriptor_distance(a, b) == int(np.unpackbits(a ^ b).sum())
    z = np.zeros(32, np.uint8)
    assert O.descriptor_distance(z, z) == 0 and O.descriptor_distance(z, ~z) == 256


@pytest.mark.parametrize("nq,nt,seed", [(1000, 1000, 0), (37, 1024, 1), (1, 1, 2), (513, 77, 3)])
def test_bf_match_equals_cv2(nq, nt, seed):
    rng = np.random.default_rng(seed)
    dq = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    dt = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    dt[rng.integers(0, nt, nt // 4)] = dt[rng.integers(0, nt, nt // 4)]  # duplicate rows -> ties
    m = cv2.BFMatcher(cv2.NORM_HAMMING).match(dq, dt)
    idx, dist = O.bf_match(dq, dt)
    assert [x.queryIdx for x in m] == list(range(nq))
    assert [x.trainIdx for x in m] == idx.tolist()
    assert [int(x.distance) for x in m] == dist.tolist()


def _gms_numpy(p1, s1, p2, s2, matches):
    """gms_matcher::run(1) restated literally with a dense 400x400 table (gms_matcher.h:385-419)."""
    G = 20
    n1 = p1.astype(np.float32) / np.array(s1, np.float32)
    n2 = p2.astype(np.float32) / np.array(s2, np.float32)
    nm = len(matches)
    mask = np.zeros(nm, bool)
    pairs = np.zeros((nm, 2), np.int64)

    def nb9(i):
        x, y = i % G, i // G
        out = []
        for yi in (-1, 0, 1):
            for xi in (-1, 0, 1):
                xx, yy = x + xi, y + yi
                out.append(-1 if (xx < 0 or xx >= G or yy < 0 or yy >= G) else xx + yy * G)
        return out

    for t in (1, 2, 3, 4):
        stat = np.zeros((400, 400), np.int64)
        nleft = np.zeros(400, np.int64)
        for i, (q, tr) in enumerate(matches):
            fx = np.float32(n1[q, 0] * np.float32(G)); fy = np.float32(n1[q, 1] * np.float32(G))
            x = int(np.floor(float(fx) + (0.5 if t in (2, 4) else 0.0)))
            y = int(np.floor(float(fy) + (0.5 if t in (3, 4) else 0.0)))
            l = -1 if (x >= G or y >= G) else x + y * G
            pairs[i, 0] = l
            if t == 1:
                pairs[i, 1] = int(np.floor(np.float32(n2[tr, 0] * np.float32(G)))) + \
                              int(np.floor(np.float32(n2[tr, 1] * np.float32(G)))) * G
            r = pairs[i, 1]
            if l < 0 or r < 0 or l >= 400 or r >= 400:
                continue
            stat[l, r] += 1
            nleft[l] += 1
        cp = np.full(400, -1, np.int64)
        for i in range(400):
            if stat[i].sum() == 0:
                continue
            cp[i] = int(np.argmax(stat[i]))
            score = 0; th = 0.0; npair = 0
            for ll, rr in zip(nb9(i), nb9(cp[i])):
                if ll == -1 or rr == -1:
                    continue
                score += stat[ll, rr]; th += nleft[ll]; npair += 1
            if score < 6 * np.sqrt(th / npair):
                cp[i] = -2
        for i in range(nm):
            if pairs[i, 0] >= 0 and cp[pairs[i, 0]] == pairs[i, 1]:
                mask[i] = True
    return mask


def _match_set(seed, n=1000, inlier_frac=0.6):
    rng = np.random.default_rng(seed)
    p1 = np.stack([rng.uniform(19, 620, n), rng.uniform(19, 460, n)], 1).astype(np.float32)
    p2 = p1 + rng.normal(0, 1.0, p1.shape).astype(np.float32) + np.float32([6, -4])
    p2 = np.clip(p2, 0, [639, 479]).astype(np.float32)
    tr = np.arange(n)
    bad = rng.random(n) > inlier_frac
    tr[bad] = rng.integers(0, n, bad.sum())
    return p1, p2, np.stack([np.arange(n), tr], 1).astype(np.int32)


@pytest.mark.parametrize("seed,n,frac", [(0, 1000, 0.6), (1, 1000, 0.1), (2, 300, 0.9), (3, 3000, 0.5)])
def test_gms_against_dense_numpy(seed, n, frac):
    p1, p2, m = _match_set(seed, n, frac)
    mask, cnt = O.gms_filter(p1, (640, 480), p2, (640, 480), m)
    ref = _gms_numpy(p1, (640, 480), p2, (640, 480), m)
    assert cnt == int(mask.sum())
    assert np.array_equal(mask, ref)
    if frac >= 0.5:
        assert cnt > 0.3 * n * frac  # sanity: coherent matches survive


def test_gms_empty_and_degenerate():
    p = np.zeros((0, 2), np.float32)
    mask, cnt = O.gms_filter(p, (640, 480), p, (640, 480), np.zeros((0, 2), np.int32))
    assert cnt == 0 and len(mask) == 0
    # every keypoint in one cell: single dense cell pair survives
    p1 = np.full((500, 2), 100.0, np.float32)
    m = np.stack([np.arange(500), np.arange(500)], 1).astype(np.int32)
    mask, cnt = O.gms_filter(p1, (640, 480), p1, (640, 480), m)
    assert cnt == 500 and mask.all()

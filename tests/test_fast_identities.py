"""Bit-level identities the FAST kernel (`k_fast_cells2`, geoflowslam_b200/csrc/orb.cu) relies on,
checked in plain Python against the scalar definition of cv::FAST:

* the antipodal-pair reject on eight pixels per thread: even / odd bytes of two adjacent words as
  16-bit lanes, `v > c + th` / `v < c - th` as bit 15 of one 32-bit add per two pixels;
* the packed two-polarity arc network: `q = v * 0xFFFF + ((c + 256) | (256 - c) << 16)` holds
  `c - v + 256` and `v - c + 256`, per-half min / max give max over the 16 nine-arcs of
  max(min d, min -d).
The GPU parity tests (tests/test_gpu_orb.py) check the kernel itself; this file documents and pins
the arithmetic on the CPU.
"""
import numpy as np

M = 0xFFFFFFFF


def prmt(a, b, sel):
    by = [(a >> (8 * i)) & 255 for i in range(4)] + [(b >> (8 * i)) & 255 for i in range(4)]
    r = 0
    for i in range(4):
        r |= by[(sel >> (4 * i)) & 7] << (8 * i)
    return r


def reject(K, c, u, d, lf, rt):
    HB, LB = (c + K) & M, (K - c) & M
    nb = (((HB - u) & M) & ((HB - d) & M)) | (((HB - lf) & M) & ((HB - rt) & M))
    nd = (((LB + u) & M) & ((LB + d) & M)) | (((LB + lf) & M) & ((LB + rt) & M))
    return (~(nb & nd)) & 0x80008000


def test_antipodal_reject_eight_pixels():
    rng = np.random.default_rng(1)
    E = lambda w: prmt(w, 0, 0x4240)
    O = lambda w: prmt(w, 0, 0x4341)
    for trial in range(3000):
        th = int(rng.choice([7, 25, 0, 255, 100]))
        mode = trial % 3
        if mode == 0:
            row = rng.integers(0, 256, size=(7, 16))
        elif mode == 1:
            row = np.clip(128 + rng.integers(-40, 40, size=(7, 16)), 0, 255)
        else:
            row = rng.choice([0, 255, th, 255 - th, 128], size=(7, 16))
        word = lambda y, w: int(sum(int(row[y, 4 * w + i]) << (8 * i) for i in range(4)))
        Wm, C0, C1, Wp = word(3, 0), word(3, 1), word(3, 2), word(3, 3)
        U0, U1, D0, D1 = word(0, 1), word(0, 2), word(6, 1), word(6, 2)
        K = ((th + 0x8000) * 0x00010001) & M
        eM, oM, e0, o0, e1, o1, eP, oP = E(Wm), O(Wm), E(C0), O(C0), E(C1), O(C1), E(Wp), O(Wp)
        t0e = reject(K, e0, E(U0), E(D0), oM, prmt(o0, o1, 0x5432))
        t0o = reject(K, o0, O(U0), O(D0), prmt(eM, e0, 0x5432), e1)
        t1e = reject(K, e1, E(U1), E(D1), o0, prmt(o1, oP, 0x5432))
        t1o = reject(K, o1, O(U1), O(D1), prmt(e0, e1, 0x5432), eP)
        ta, tb = (t0e >> 1) | t0o, (t1e >> 1) | t1o
        bits = [bool(t & m) for t in (ta, tb) for m in (0x4000, 0x8000, 0x40000000, 0x80000000)]
        for j in range(8):
            x = 4 + j
            c = int(row[3, x]); hi, lo = c + th, c - th
            v0, v8, v4, v12 = int(row[6, x]), int(row[0, x]), int(row[3, x + 3]), int(row[3, x - 3])
            br = ((v0 > hi) | (v8 > hi)) & ((v4 > hi) | (v12 > hi))
            dk = ((v0 < lo) | (v8 < lo)) & ((v4 < lo) | (v12 < lo))
            assert bool(br | dk) == bits[j], (trial, j)


def _s16(x):
    x &= 0xFFFF
    return x - 65536 if x >= 32768 else x


def _vmin(a, b):
    return (min(_s16(a), _s16(b)) & 0xFFFF) | ((min(_s16(a >> 16), _s16(b >> 16)) & 0xFFFF) << 16)


def _vmax(a, b):
    return (max(_s16(a), _s16(b)) & 0xFFFF) | ((max(_s16(a >> 16), _s16(b >> 16)) & 0xFFFF) << 16)


def test_packed_arc_network_equals_fast_score():
    rng = np.random.default_rng(2)
    for trial in range(1500):
        c = int(rng.integers(0, 256))
        v = [int(x) for x in rng.choice([0, 255, c, min(255, c + 30), max(0, c - 30), int(rng.integers(0, 256))], size=16)]
        CC = ((c + 256) + ((256 - c) << 16)) & M
        q = [(vv * 0xFFFF + CC) & M for vv in v]
        assert all((x & 0xFFFF) == c - vv + 256 and (x >> 16) == vv - c + 256 for x, vv in zip(q, v))
        a2 = [_vmin(q[k], q[(k + 1) & 15]) for k in range(16)]
        a4 = [_vmin(a2[k], a2[(k + 2) & 15]) for k in range(16)]
        best = 0
        for k in range(16):
            best = _vmax(best, _vmin(_vmin(a4[k], a4[(k + 4) & 15]), q[(k + 8) & 15]))
        got = max(best & 0xFFFF, best >> 16) - 256
        d = [c - x for x in v]
        ref = max(max(min(d[(k + i) & 15] for i in range(9)), min(-d[(k + i) & 15] for i in range(9))) for k in range(16))
        assert got == ref

"""Host-side models of two index schemes the CUDA kernels rely on (no GPU needed):
* warp_reduce_vec (geoflowslam_b200/csrc/gicp.cu): the recursive-halving reduction of 29 per-lane sums with 31 shuffles;
* k_klt_pad_level / win_tma_ok (geoflowslam_b200/csrc/klt.cu): the reflect-101 padded level copies the optical-flow windows are
  fetched from, and the 16-byte-column box arithmetic of the tensor-map loads.
"""
import numpy as np


def warp_reduce_vec_model(vals):
    """vals[lane][k], k < NV <= 32 -> per-lane result; follows the kernel: in the step with offset o a lane keeps one half of its
    remaining values and hands the other half to lane ^ o."""
    nv = vals.shape[1]
    v = np.zeros((32, 32))
    v[:, :nv] = vals
    o = 16
    while o >= 1:
        new = v.copy()
        for lane in range(32):
            upper = (lane & o) != 0
            for k in range(o):
                send_from_partner = v[lane ^ o, k] if ((lane ^ o) & o) != 0 else v[lane ^ o, k + o]   # what the partner sends
                keep = v[lane, k + o] if upper else v[lane, k]
                new[lane, k] = keep + send_from_partner
        v = new
        o >>= 1
    return v[:, 0]


def test_recursive_halving_puts_sum_l_in_lane_l():
    rng = np.random.default_rng(3)
    vals = rng.integers(-1000, 1000, (32, 29)).astype(np.float64)     # integers: the sums are exact in any order
    out = warp_reduce_vec_model(vals)
    assert np.array_equal(out[:29], vals.sum(0))
    assert np.all(out[29:] == 0)


def reflect101(p, n):
    if n == 1:
        return 0
    while p < 0 or p >= n:
        p = -p if p < 0 else 2 * (n - 1) - p
    return p


def test_padded_level_equals_numpy_reflect_and_windows_are_boxes():
    PAD_X, PAD_Y = 48, 40
    rng = np.random.default_rng(5)
    for w, h in ((80, 60), (40, 30), (37, 21)):
        img = rng.integers(0, 256, (h, w)).astype(np.uint8)
        pw = (w + 2 * PAD_X + 15) // 16 * 16
        pad = np.zeros((h + 2 * PAD_Y, pw), np.uint8)
        for y in range(pad.shape[0]):
            for x in range(pw):
                pad[y, x] = img[reflect101(y - PAD_Y, h), reflect101(x - PAD_X, w)]
        if w > PAD_X and h > PAD_Y:   # single reflection: numpy's 'reflect' is BORDER_REFLECT_101
            ref = np.pad(img, ((PAD_Y, PAD_Y), (PAD_X, PAD_X)), mode="reflect")
            assert np.array_equal(pad[:, : w + 2 * PAD_X], ref)
        # every window the tracker may ask for (origin in [-win, w) x [-win, h), 36 x 36) is a box of the padded copy whose
        # 64-byte-wide fetch starts on a 16-byte column and contains it
        win, ww, pitch = 35, 36, 64
        for x0, y0 in ((-win, -win), (w - 1, h - 1), (-7, h - 3), (w - 20, -win), (3, 5)):
            assert x0 >= -PAD_X and y0 >= -PAD_Y and x0 + ww <= w + PAD_X and y0 + ww <= h + PAD_Y       # win_tma_ok
            bx, off = (x0 + PAD_X) & ~15, (x0 + PAD_X) & 15
            assert bx % 16 == 0 and off + ww <= pitch
            box = np.zeros((ww, pitch), np.uint8)
            for r in range(ww):
                for c in range(pitch):
                    box[r, c] = pad[y0 + PAD_Y + r, bx + c] if bx + c < pw else 0       # the tensor path fills out-of-bounds with zeros
            want = np.array([[img[reflect101(y0 + r, h), reflect101(x0 + c, w)] for c in range(ww)] for r in range(ww)], np.uint8)
            assert np.array_equal(box[:, off:off + ww], want)


def test_derivative_window_offsets():
    # derivative window: 40 entries wide from the multiple of 4 at or below ipx (also for negative ipx: two's complement floor)
    for ipx in (-35, -4, -1, 0, 3, 17, 602):
        xa, off = ipx & ~3, ipx & 3
        assert xa % 4 == 0 and xa <= ipx and xa + off == ipx and off + 36 <= 40

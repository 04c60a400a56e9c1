"""Seeded synthetic inputs of the shapes BASELINE.json names.  numpy only."""
import numpy as np


def _value_noise(rng, h, w):
    img = np.zeros((h, w), np.float32)
    amp = 1.0
    for cell in (64, 32, 16, 8, 4):
        gh, gw = h // cell + 2, w // cell + 2
        g = rng.random((gh, gw), dtype=np.float32)
        ys = np.arange(h, dtype=np.float32) / cell
        xs = np.arange(w, dtype=np.float32) / cell
        y0 = ys.astype(np.int32); x0 = xs.astype(np.int32)
        fy = (ys - y0)[:, None]; fx = (xs - x0)[None, :]
        a = g[y0][:, x0]; b = g[y0][:, x0 + 1]; c = g[y0 + 1][:, x0]; d = g[y0 + 1][:, x0 + 1]
        img += amp * ((a * (1 - fx) + b * fx) * (1 - fy) + (c * (1 - fx) + d * fx) * fy)
        amp *= 0.6
    img -= img.min()
    img /= max(img.max(), 1e-6)
    return img


def scene(seed, h=560, w=720, nrect=200):
    """Textured scene: multi-octave value noise + random rectangles (corner sources)."""
    rng = np.random.default_rng(seed)
    img = 40.0 + 120.0 * _value_noise(rng, h, w)
    for _ in range(nrect):
        rw, rh = rng.integers(6, 60, 2)
        x0 = rng.integers(0, w - rw); y0 = rng.integers(0, h - rh)
        img[y0:y0 + rh, x0:x0 + rw] = rng.integers(0, 256)
    return img.astype(np.float32)


def _homography(rng, w, h, max_px):
    """Homography moving the 4 image corners by <= max_px (DLT on 4 points)."""
    src = np.array([[0, 0], [w, 0], [w, h], [0, h]], np.float64)
    dst = src + rng.uniform(-max_px, max_px, (4, 2))
    A = []
    for (x, y), (u, v) in zip(src, dst):
        A.append([x, y, 1, 0, 0, 0, -u * x, -u * y, -u])
        A.append([0, 0, 0, x, y, 1, -v * x, -v * y, -v])
    _, _, vt = np.linalg.svd(np.array(A))
    H = vt[-1].reshape(3, 3)
    return H / H[2, 2]


def frame_from_scene(sc, rng, w=640, h=480, max_px=5.0, noise=1.5):
    """Bilinear sample of the scene under a small random homography, plus sensor noise."""
    H = _homography(rng, w, h, max_px)
    ys, xs = np.mgrid[0:h, 0:w].astype(np.float64)
    den = H[2, 0] * xs + H[2, 1] * ys + H[2, 2]
    u = (H[0, 0] * xs + H[0, 1] * ys + H[0, 2]) / den + (sc.shape[1] - w) / 2
    v = (H[1, 0] * xs + H[1, 1] * ys + H[1, 2]) / den + (sc.shape[0] - h) / 2
    u0 = np.clip(np.floor(u).astype(np.int32), 0, sc.shape[1] - 2)
    v0 = np.clip(np.floor(v).astype(np.int32), 0, sc.shape[0] - 2)
    fu = (u - u0).astype(np.float32); fv = (v - v0).astype(np.float32)
    img = (sc[v0, u0] * (1 - fu) + sc[v0, u0 + 1] * fu) * (1 - fv) + \
          (sc[v0 + 1, u0] * (1 - fu) + sc[v0 + 1, u0 + 1] * fu) * fv
    img = img + rng.normal(0, noise, img.shape).astype(np.float32)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def orb_frames(n, w=640, h=480, group=8, seed0=1000):
    """configs[1]: n gray frames; frames of one group view one scene through different small
    homographies (so consecutive frames match); frame i uses default_rng(seed0 + i)."""
    out = np.empty((n, h, w), np.uint8)
    sc = None
    for i in range(n):
        if i % group == 0:
            sc = scene(seed0 + i, h + 80, w + 80)
        out[i] = frame_from_scene(sc, np.random.default_rng(seed0 + i), w, h)
    return out


# ------------------------------------------------------------------------------------------------
# RGB-D clouds for GICP (BASELINE configs[2]): a room (3 planes) + 5 boxes seen from two poses.
# ------------------------------------------------------------------------------------------------
def _rot(rv):
    th = np.linalg.norm(rv)
    if th < 1e-12:
        return np.eye(3)
    k = rv / th
    K = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * K + (1 - np.cos(th)) * K @ K


def _raycast(o, d, boxes):
    """o: (3,), d: (N,3) unit-z-normalised rays.  Returns depth along each ray to the nearest surface of
    the room x in [-3,3], y in [-1.5,1.5], z in [..,6] or of the boxes."""
    t = np.full(len(d), np.inf)
    with np.errstate(divide="ignore", invalid="ignore"):
        for axis, val in ((2, 6.0), (1, 1.5), (0, -3.0), (0, 3.0), (1, -1.5)):
            tt = (val - o[axis]) / d[:, axis]
            tt[~(tt > 1e-6)] = np.inf
            t = np.minimum(t, tt)
        for (lo, hi) in boxes:
            t0 = (lo - o) / d
            t1 = (hi - o) / d
            tn = np.minimum(t0, t1).max(1)
            tf = np.maximum(t0, t1).min(1)
            hit = (tf >= tn) & (tn > 1e-6)
            t = np.where(hit, np.minimum(t, tn), t)
    return t


def gicp_pair(seed, n_target=50000, max_rot_deg=3.0, max_trans=0.05, noise_z=0.002):
    """-> (target float32 (N,4), source float32 (M,4), T_target_source_true (4,4) float64).
    Both clouds are expressed in their own camera frame (x right, y down, z forward)."""
    rng = np.random.default_rng(seed)
    boxes = []
    for _ in range(5):
        c = np.array([rng.uniform(-2, 2), rng.uniform(0.3, 1.2), rng.uniform(1.5, 5.0)])
        s = rng.uniform(0.2, 0.6, 3)
        boxes.append((c - s, c + s))
    w = int(round(np.sqrt(n_target * 4 / 3)))
    h = int(round(w * 3 / 4))
    fx = fy = 0.95 * w
    cx, cy = w / 2 - 0.5, h / 2 - 0.5
    us, vs = np.meshgrid(np.arange(w), np.arange(h))
    dirs = np.stack([(us.ravel() - cx) / fx, (vs.ravel() - cy) / fy, np.ones(w * h)], 1)

    def view(R, tvec):
        # camera-to-world: Xw = R Xc + t
        dw = dirs @ R.T
        depth = _raycast(tvec, dw, boxes)  # parametrised so that Xc = depth * dirs (dirs has z = 1)
        ok = np.isfinite(depth) & (depth > 0.3) & (depth < 10.0)
        z = depth[ok] + rng.normal(0, noise_z, ok.sum())
        pts = dirs[ok] * z[:, None]
        return np.concatenate([pts, np.ones((len(pts), 1))], 1).astype(np.float32)

    R1, t1 = np.eye(3), np.zeros(3)
    rv = np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg, 3))
    R2 = _rot(rv)
    t2 = rng.uniform(-max_trans, max_trans, 3)
    tgt = view(R1, t1)
    src = view(R2, t2)
    T = np.eye(4)
    T[:3, :3] = R2  # X_target = R2 X_source + t2
    T[:3, 3] = t2
    return tgt, src, T

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


def depth_frames(seed, n_frames=4, w=640, h=480, stride=2, n_points=50052, max_rot_deg=3.0, max_trans=0.05, noise_z=0.002,
                 depth_factor=1000.0, cam=None):
    """n_frames 16-bit depth images (h, w) of ONE scene (the room and boxes of gicp_pair) seen from n_frames nearby camera
    poses -- what an RGB-D driver hands to System::TrackRGBD (depth in 1 / depth_factor metres, 0 = invalid).  Even frames
    sit at the base pose, every odd frame at its own U(+-max_rot_deg, +-max_trans) perturbation of it, so CONSECUTIVE
    frames (also last -> first of an even-length ring) differ by exactly the configs[2] perturbation (+-3 deg, +-5 cm);
    each frame has its own depth noise.  Only the pixels Frame::ConvertDepthToPointCloud samples
    (every `stride`-th row and column, reference src/Frame.cc:606-607) inside a centred window are rendered, sized so that a
    frame yields about n_points cloud points (configs[2]: 50k); everything else is 0 like the invalid border of a real
    sensor.  -> (depth uint16 (n_frames, h, w), poses [(R, t) camera-to-world])."""
    cam = cam or G1_CAM
    rng = np.random.default_rng(seed)
    boxes = []
    for _ in range(5):
        c = np.array([rng.uniform(-2, 2), rng.uniform(0.3, 1.2), rng.uniform(1.5, 5.0)])
        sz = rng.uniform(0.2, 0.6, 3)
        boxes.append((c - sz, c + sz))
    gw, gh = (w + stride - 1) // stride, (h + stride - 1) // stride
    ww = min(gw, int(round(np.sqrt(n_points * 4 / 3))))
    wh = min(gh, int(round(ww * 3 / 4)))
    u0, v0 = (gw - ww) // 2, (gh - wh) // 2
    us, vs = np.meshgrid((u0 + np.arange(ww)) * stride, (v0 + np.arange(wh)) * stride)
    dirs = np.stack([(us.ravel() - cam["cx"]) / cam["fx"], (vs.ravel() - cam["cy"]) / cam["fy"], np.ones(ww * wh)], 1)
    out = np.zeros((n_frames, h, w), np.uint16)
    poses = []
    for k in range(n_frames):
        R = _rot(np.deg2rad(rng.uniform(-max_rot_deg, max_rot_deg, 3))) if k & 1 else np.eye(3)
        t = rng.uniform(-max_trans, max_trans, 3) if k & 1 else np.zeros(3)
        depth = _raycast(t, dirs @ R.T, boxes)
        ok = np.isfinite(depth) & (depth > 0.3) & (depth < 10.0)
        z = np.where(ok, depth + rng.normal(0, noise_z, len(depth)), 0.0)
        out[k, vs.ravel(), us.ravel()] = np.clip(np.rint(z * depth_factor), 0, 65535).astype(np.uint16)
        poses.append((R, t))
    return out, poses


# ------------------------------------------------------------------------------------------------
# LocalInertialBA problem (BASELINE configs[3]): 20 KFs on an arc, 3000 points, ~15k stereo
# observations, 200 Hz IMU preintegrated in float32 as IMU::Preintegrated does
# (reference src/ImuTypes.cc:184-246), 1 fixed predecessor keyframe.
# ------------------------------------------------------------------------------------------------
PRE_STRIDE = 292  # dR9 dV3 dP3 JRg9 JVg9 JVa9 JPg9 JPa9 C225 dT1 b6(bax bay baz bwx bwy bwz)
G1_CAM = dict(fx=606.986, fy=607.011, cx=311.519, cy=247.260, bf=606.986 * 0.0745)  # g1 yaml :25-28,54
IMU_NOISE = dict(ng=2.443e-3, na=1.176e-2, ngw=1e-4, naw=1e-3, freq=200.0)         # g1 yaml :106-110


def _f32(x):
    return np.asarray(x, np.float32)


def _hat32(v):
    return _f32([[0, -v[2], v[1]], [v[2], 0, -v[0]], [-v[1], v[0], 0]])


def _polar32(R):
    U, _, Vt = np.linalg.svd(R.astype(np.float32))
    return (U @ Vt).astype(np.float32)


def preintegrate(acc, gyr, dt, bias, noise=IMU_NOISE):
    """IMU::Preintegrated::IntegrateNewMeasurement over the samples, float32 throughout.
    bias = (bax, bay, baz, bwx, bwy, bwz).  Returns the packed PRE_STRIDE record."""
    f = np.float32
    sf = np.sqrt(f(noise["freq"]))
    ng, na = f(noise["ng"]) * sf, f(noise["na"]) * sf
    ngw, naw = f(noise["ngw"]) / sf, f(noise["naw"]) / sf
    Nga = np.diag(_f32([ng * ng] * 3 + [na * na] * 3))
    NgaWalk = np.diag(_f32([ngw * ngw] * 3 + [naw * naw] * 3))
    dR = np.eye(3, dtype=f); dV = np.zeros(3, f); dP = np.zeros(3, f)
    JRg = np.zeros((3, 3), f); JVg = np.zeros((3, 3), f); JVa = np.zeros((3, 3), f)
    JPg = np.zeros((3, 3), f); JPa = np.zeros((3, 3), f)
    C = np.zeros((15, 15), f)
    dT = f(0)
    b = _f32(bias)
    dt = f(dt)
    for a_m, w_m in zip(_f32(acc), _f32(gyr)):
        A = np.eye(9, dtype=f); B = np.zeros((9, 6), f)
        a = a_m - b[:3]
        Wacc = _hat32(a)
        dP = dP + dV * dt + f(0.5) * (dR @ a) * dt * dt
        dV = dV + (dR @ a) * dt
        A[3:6, 0:3] = -dR * dt @ Wacc
        A[6:9, 0:3] = f(-0.5) * dR * dt * dt @ Wacc
        A[6:9, 3:6] = np.eye(3, dtype=f) * dt
        B[3:6, 3:6] = dR * dt
        B[6:9, 3:6] = f(0.5) * dR * dt * dt
        JPa = JPa + JVa * dt - f(0.5) * dR * dt * dt
        JPg = JPg + JVg * dt - f(0.5) * dR * dt * dt @ Wacc @ JRg
        JVa = JVa - dR * dt
        JVg = JVg - dR * dt @ Wacc @ JRg
        v = (w_m - b[3:]) * dt
        d2 = f(v @ v); d = np.sqrt(d2)
        W = _hat32(v)
        if d < 1e-4:
            dRi = np.eye(3, dtype=f) + W; rJ = np.eye(3, dtype=f)
        else:
            dRi = np.eye(3, dtype=f) + W * (np.sin(d) / d) + W @ W * ((f(1) - np.cos(d)) / d2)
            rJ = np.eye(3, dtype=f) - W * ((f(1) - np.cos(d)) / d2) + W @ W * ((d - np.sin(d)) / (d2 * d))
        dR = _polar32(dR @ dRi)
        A[0:3, 0:3] = dRi.T
        B[0:3, 0:3] = rJ * dt
        C[0:9, 0:9] = A @ C[0:9, 0:9] @ A.T + B @ Nga @ B.T
        C[9:15, 9:15] += NgaWalk
        JRg = dRi.T @ JRg - rJ * dt
        dT = dT + dt
    rec = np.concatenate([dR.ravel(), dV, dP, JRg.ravel(), JVg.ravel(), JVa.ravel(), JPg.ravel(), JPa.ravel(),
                          C.ravel(), [dT], b]).astype(np.float32)
    assert rec.size == PRE_STRIDE
    return rec


def ba_problem(seed=3000, n_kf=20, n_points=3000, obs_per_point=5, kf_dt=0.5, imu_rate=200, b_large=True,
               rot_noise_deg=1.0, trans_noise=0.02, point_noise=0.02, outlier_frac=0.01, n_icp=0):
    """Synthetic LocalInertialBA problem as plain arrays (the flattened export of the reference's
    KeyFrame / MapPoint graph, see include/gfs_b200.h GfsBaProblem).  Keyframe 0 is the newest
    (vpOptimizableKFs order, Optimizer.cc:3078-3086); the fixed predecessor is index n_kf."""
    rng = np.random.default_rng(seed)
    g = np.array([0, 0, -9.81])
    Rbc = np.array([[0, 0, 1.0], [-1, 0, 0], [0, -1, 0]])  # camera z forward / x right / y down in a FLU body
    tbc = np.array([0.05, 0.02, 0.01])
    Rcb = Rbc.T; tcb = -Rcb @ tbc
    nk = n_kf + 1
    per = int(round(kf_dt * imu_rate))
    dt = 1.0 / imu_rate
    T_total = kf_dt * (nk - 1)
    w_yaw = (2.0 / 5.0) / T_total  # 2 m arc on a 5 m radius circle

    def pose(t):
        yaw = w_yaw * t + 0.02 * np.sin(1.3 * t)
        pitch = 0.03 * np.sin(0.9 * t); roll = 0.02 * np.cos(1.1 * t)
        R = _rot(np.array([0, 0, yaw])) @ _rot(np.array([0, pitch, 0])) @ _rot(np.array([roll, 0, 0]))
        p = np.array([5 * np.sin(w_yaw * t), 5 * (1 - np.cos(w_yaw * t)), 0.05 * np.sin(0.7 * t)])
        return R, p

    h = 1e-4
    ts = np.arange(0, (nk - 1) * per) * dt + dt / 2  # mid-point samples
    acc = np.zeros((len(ts), 3)); gyr = np.zeros((len(ts), 3))
    bg_true = np.array([0.002, -0.001, 0.0015]); ba_true = np.array([0.02, -0.03, 0.01])
    for i, t in enumerate(ts):
        R0, p0 = pose(t); Rp, pp = pose(t + h); Rm, pm = pose(t - h)
        a_w = (pp - 2 * p0 + pm) / (h * h)
        dRm = R0.T @ (Rp - Rm) / (2 * h)
        gyr[i] = [dRm[2, 1], dRm[0, 2], dRm[1, 0]]
        acc[i] = R0.T @ (a_w - g)
    ns = IMU_NOISE
    gyr += bg_true + rng.normal(0, ns["ng"] * np.sqrt(imu_rate), gyr.shape)
    acc += ba_true + rng.normal(0, ns["na"] * np.sqrt(imu_rate), acc.shape)

    # chronological keyframes c = 0 (fixed predecessor) .. n_kf (newest); window index i = n_kf - c
    Rwb_t, twb_t, vel_t = [], [], []
    for c in range(nk):
        t = c * kf_dt
        R, p = pose(t); _, pp = pose(t + h); _, pm = pose(t - h)
        Rwb_t.append(R); twb_t.append(p); vel_t.append((pp - pm) / (2 * h))
    order = list(range(n_kf, 0, -1)) + [0]  # problem index -> chronological index
    bias_est = np.concatenate([ba_true, bg_true]) + rng.normal(0, [2e-3] * 3 + [2e-4] * 3)
    pre = np.zeros((n_kf, PRE_STRIDE), np.float32)
    in_kf1 = np.zeros(n_kf, np.int32); in_kf2 = np.zeros(n_kf, np.int32); down = np.zeros(n_kf, np.uint8)
    for i in range(n_kf):  # edge i links problem kf i (cur) with its predecessor
        c = order[i]
        seg = slice((c - 1) * per, c * per)
        pre[i] = preintegrate(acc[seg], gyr[seg], dt, bias_est)
        in_kf2[i] = i
        in_kf1[i] = i + 1  # predecessor: next window index, or the fixed KF (index n_kf)
        down[i] = 1 if i == n_kf - 1 else 0

    # points in front of the trajectory
    cam = G1_CAM
    pts = np.zeros((n_points, 3))
    for j in range(n_points):  # unproject a random pixel of a random keyframe at 2..8 m
        c = int(rng.integers(0, nk))
        u, v, z = rng.uniform(20, 620), rng.uniform(20, 460), rng.uniform(2.0, 8.0)
        Xc = np.array([(u - cam["cx"]) / cam["fx"] * z, (v - cam["cy"]) / cam["fy"] * z, z])
        pts[j] = Rwb_t[c] @ (Rbc @ Xc + tbc) + twb_t[c]
    obs_kf, obs_pt, obs_uvr, obs_is2 = [], [], [], []
    for j in range(n_points):
        vis = []
        for i in range(nk):
            c = order[i]
            Rcw = Rcb @ Rwb_t[c].T
            Xc = Rcw @ (pts[j] - twb_t[c]) + tcb
            if Xc[2] < 0.3:
                continue
            u = cam["fx"] * Xc[0] / Xc[2] + cam["cx"]; v = cam["fy"] * Xc[1] / Xc[2] + cam["cy"]
            if 0 <= u < 640 and 0 <= v < 480:
                vis.append((i, u, v, u - cam["bf"] / Xc[2]))
        if len(vis) < 2:
            continue
        k = min(len(vis), obs_per_point + int(rng.integers(-1, 2)))
        s = int(rng.integers(0, len(vis) - k + 1))
        for (i, u, v, ur) in vis[s:s + k]:
            octv = int(rng.integers(0, 4))
            sig = 1.2 ** octv
            nz = rng.normal(0, sig, 3)
            if rng.random() < outlier_frac:
                nz += rng.normal(0, 30.0, 3)
            mono = rng.random() < 0.05  # a few depth-less observations -> EdgeMono
            obs_kf.append(i); obs_pt.append(j)
            obs_uvr.append([np.float32(u + nz[0]), np.float32(v + nz[1]), -1.0 if mono else np.float32(ur + nz[2])])
            obs_is2.append(np.float32(1.0) / np.float32(np.float32(sig) * np.float32(sig)))

    # initial estimates: float32 keyframe states, as the reference stores them
    def f64(x):
        return np.asarray(x, np.float32).astype(np.float64)
    kf_Rwb = np.zeros((nk, 9)); kf_twb = np.zeros((nk, 3)); kf_Rcw = np.zeros((nk, 9)); kf_tcw = np.zeros((nk, 3))
    kf_vel = np.zeros((nk, 3)); kf_bg = np.zeros((nk, 3)); kf_ba = np.zeros((nk, 3))
    for i in range(nk):
        c = order[i]
        fixed = i == n_kf
        R = Rwb_t[c] if fixed else Rwb_t[c] @ _rot(np.deg2rad(rng.uniform(-rot_noise_deg, rot_noise_deg, 3)))
        p = twb_t[c] if fixed else twb_t[c] + rng.uniform(-trans_noise, trans_noise, 3)
        R32 = _polar32(R.astype(np.float32)); p32 = p.astype(np.float32)
        Rcw32 = (Rcb.astype(np.float32) @ R32.T).astype(np.float32)
        tcw32 = (Rcb.astype(np.float32) @ (-(R32.T @ p32)) + tcb.astype(np.float32)).astype(np.float32)
        kf_Rwb[i] = f64(R32).ravel(); kf_twb[i] = f64(p32)
        kf_Rcw[i] = f64(Rcw32).ravel(); kf_tcw[i] = f64(tcw32)
        kf_vel[i] = f64(vel_t[c] + (0 if fixed else rng.uniform(-0.02, 0.02, 3)))
        kf_bg[i] = f64(bias_est[3:]); kf_ba[i] = f64(bias_est[:3])
    pt_xyz = f64(pts + rng.normal(0, point_noise, pts.shape))
    # optional EdgeICP factors between consecutive keyframes (vertex 0 = previous KF, vertex 1 = KF):
    # T_c1_c2 = T_c1w * T_c2w^-1 from the true poses plus GICP-sized noise
    icp_kf1, icp_kf2, icp_Rt = [], [], []
    for i in range(min(n_icp, n_kf)):
        c2, c1 = order[i], order[i + 1]
        def Tcw(c):
            R = Rcb @ Rwb_t[c].T
            return R, R @ (-twb_t[c]) + tcb
        R1, t1 = Tcw(c1); R2, t2 = Tcw(c2)
        R12 = R1 @ R2.T @ _rot(rng.normal(0, 5e-4, 3)); t12 = t1 - R1 @ R2.T @ t2 + rng.normal(0, 2e-3, 3)
        icp_kf1.append(i + 1); icp_kf2.append(i); icp_Rt.append(np.concatenate([R12.ravel(), t12]))
    icp = dict(n_icp=len(icp_kf1), icp_kf1=np.array(icp_kf1, np.int32), icp_kf2=np.array(icp_kf2, np.int32),
               icp_Rt=np.array(icp_Rt, np.float64).reshape(-1, 12))
    return dict(icp, 
        n_opt_kf=n_kf, n_fixed_kf=1, n_points=n_points, n_obs=len(obs_kf), n_inertial=n_kf,
        iterations=4 if b_large else 8, b_large=int(b_large), lambda_init=1e-2 if b_large else 1.0,
        Rcb=f64(Rcb).ravel(), tcb=f64(tcb), Rbc=f64(Rbc).ravel(), tbc=f64(tbc),
        fx=np.float32(cam["fx"]), fy=np.float32(cam["fy"]), cx=np.float32(cam["cx"]), cy=np.float32(cam["cy"]),
        bf=float(np.float32(cam["bf"])),
        kf_Rwb=kf_Rwb, kf_twb=kf_twb, kf_Rcw=kf_Rcw, kf_tcw=kf_tcw, kf_vel=kf_vel, kf_bg=kf_bg, kf_ba=kf_ba,
        kf_has_imu=np.ones(nk, np.uint8),
        pt_xyz=pt_xyz, pt_close=np.ones(n_points, np.uint8),
        obs_kf=np.array(obs_kf, np.int32), obs_pt=np.array(obs_pt, np.int32),
        obs_uvr=np.array(obs_uvr, np.float64).reshape(-1, 3), obs_inv_sigma2=np.array(obs_is2, np.float32),
        in_kf1=in_kf1, in_kf2=in_kf2, in_pre=pre, in_downweight=down,
        truth=dict(Rwb=np.array([Rwb_t[c] for c in order]), twb=np.array([twb_t[c] for c in order]), pts=pts,
                   vel=np.array([vel_t[c] for c in order]), bg=bg_true, ba=ba_true))


# ------------------------------------------------------------------------------------------------
# PoseOptimization problem: one frame, ~400 map-point observations (mostly
# RGB-D "stereo" observations, some monocular), a few gross outliers, pose prior off by ~1 deg / 3 cm.
# ------------------------------------------------------------------------------------------------
def _quat_from_R(R):
    """rotation matrix -> unit quaternion (w, x, y, z), w >= 0"""
    t = np.trace(R)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        q = np.array([0.25 * s, (R[2, 1] - R[1, 2]) / s, (R[0, 2] - R[2, 0]) / s, (R[1, 0] - R[0, 1]) / s])
    else:
        i = int(np.argmax(np.diag(R)))
        j, k = (i + 1) % 3, (i + 2) % 3
        s = np.sqrt(R[i, i] - R[j, j] - R[k, k] + 1.0) * 2
        v = np.zeros(3)
        v[i] = 0.25 * s
        v[j] = (R[j, i] + R[i, j]) / s
        v[k] = (R[k, i] + R[i, k]) / s
        q = np.array([(R[k, j] - R[j, k]) / s, v[0], v[1], v[2]])
    q = q / np.linalg.norm(q)
    return q if q[0] >= 0 else -q


def pose_problem(seed=5000, n_obs=400, outlier_frac=0.1, mono_frac=0.2, rot_deg=1.0, trans=0.03, w=640, h=480):
    """-> dict(n_obs, q_wxyz f32, t f32, fx.., bf, Xw f64 (n,3), uvr f32 (n,3), inv_sigma2 f32 (n), truth...)"""
    rng = np.random.default_rng(seed)
    cam = G1_CAM
    R_true = _rot(np.deg2rad(rng.uniform(-20, 20, 3)))
    t_true = rng.uniform(-1, 1, 3)
    u = rng.uniform(20, w - 20, n_obs)
    v = rng.uniform(20, h - 20, n_obs)
    z = rng.uniform(0.6, 8.0, n_obs)
    Xc = np.stack([(u - cam["cx"]) / cam["fx"] * z, (v - cam["cy"]) / cam["fy"] * z, z], 1)
    Xw = (Xc - t_true) @ R_true  # Xc = R Xw + t
    octave = rng.integers(0, 8, n_obs)
    sigma = 1.2 ** octave
    uu = u + rng.normal(0, 1, n_obs) * sigma * 0.7
    vv = v + rng.normal(0, 1, n_obs) * sigma * 0.7
    ur = uu - cam["bf"] / z + rng.normal(0, 0.3, n_obs)
    bad = rng.random(n_obs) < outlier_frac
    uu[bad] += rng.uniform(8, 60, bad.sum()) * rng.choice([-1, 1], bad.sum())
    vv[bad] += rng.uniform(8, 60, bad.sum()) * rng.choice([-1, 1], bad.sum())
    mono = rng.random(n_obs) < mono_frac
    ur[mono] = -1.0
    R0 = _rot(np.deg2rad(rng.uniform(-rot_deg, rot_deg, 3))) @ R_true
    t0 = t_true + rng.uniform(-trans, trans, 3)
    return dict(n_obs=n_obs, q_wxyz=_quat_from_R(R0).astype(np.float32), t=t0.astype(np.float32),
                fx=np.float32(cam["fx"]), fy=np.float32(cam["fy"]), cx=np.float32(cam["cx"]), cy=np.float32(cam["cy"]),
                bf=np.float32(cam["bf"]), Xw=np.ascontiguousarray(Xw, np.float64),
                uvr=np.ascontiguousarray(np.stack([uu, vv, ur], 1), np.float32),
                inv_sigma2=(1.0 / sigma ** 2).astype(np.float32), truth_R=R_true, truth_t=t_true, truth_bad=bad)


# ------------------------------------------------------------------------------------------------
# PoseInertialOptimizationLastKeyFrame / LastFrame problem: the frame's body
# state and ~n_obs map-point observations, the previous keyframe (fixed) or previous frame (free, with
# the ConstraintPoseImu prior), and the IMU preintegration between them.
# ------------------------------------------------------------------------------------------------
def pose_inertial_problem(seed=6000, mode=0, n_obs=400, outlier_frac=0.1, mono_frac=0.2, rot_deg=0.5, trans=0.02,
                          n_rounds=4, rec_init=0, imu_rate=200, prior_H=None, w=640, h=480):
    """-> dict in the GfsPoseInertialProblem layout (include/gfs_b200.h) plus `truth`.
    mode 0 = LastKeyFrame (0.4 s of IMU since the keyframe), mode 1 = LastFrame (one 30 Hz frame interval)."""
    rng = np.random.default_rng(seed)
    g = np.array([0, 0, -9.81])
    cam = G1_CAM
    Rbc = np.array([[0, 0, 1.0], [-1, 0, 0], [0, -1, 0]])
    tbc = np.array([0.05, 0.02, 0.01])
    Rcb = Rbc.T; tcb = -Rcb @ tbc
    dt = 1.0 / imu_rate
    ph = rng.uniform(0, 6.28, 4)

    def pose(t):
        yaw = 0.15 * t + 0.05 * np.sin(1.3 * t + ph[0])
        pitch = 0.03 * np.sin(0.9 * t + ph[1]); roll = 0.02 * np.cos(1.1 * t + ph[2])
        R = _rot(np.array([0, 0, yaw])) @ _rot(np.array([0, pitch, 0])) @ _rot(np.array([roll, 0, 0]))
        p = np.array([0.6 * t + 0.1 * np.sin(0.8 * t + ph[3]), 0.2 * np.sin(0.5 * t), 0.05 * np.sin(0.7 * t)])
        return R, p

    hh = 1e-4

    def vel(t):
        return (pose(t + hh)[1] - pose(t - hh)[1]) / (2 * hh)

    t_kf = 1.0
    n_kf = 80                      # IMU samples since the last keyframe (0.4 s)
    n_fr = max(1, imu_rate // 30)  # IMU samples since the previous frame
    t1 = t_kf + n_kf * dt
    t0 = t_kf if mode == 0 else t1 - n_fr * dt
    ts = t_kf + np.arange(n_kf) * dt + dt / 2
    acc = np.zeros((n_kf, 3)); gyr = np.zeros((n_kf, 3))
    bg_true = np.array([0.002, -0.001, 0.0015]); ba_true = np.array([0.02, -0.03, 0.01])
    for i, t in enumerate(ts):
        R0, p0 = pose(t); Rp, pp = pose(t + hh); Rm, pm = pose(t - hh)
        a_w = (pp - 2 * p0 + pm) / (hh * hh)
        dRm = R0.T @ (Rp - Rm) / (2 * hh)
        gyr[i] = [dRm[2, 1], dRm[0, 2], dRm[1, 0]]
        acc[i] = R0.T @ (a_w - g)
    gyr += bg_true + rng.normal(0, IMU_NOISE["ng"] * np.sqrt(imu_rate), gyr.shape)
    acc += ba_true + rng.normal(0, IMU_NOISE["na"] * np.sqrt(imu_rate), acc.shape)
    bias_est = np.concatenate([ba_true, bg_true]) + rng.normal(0, [2e-3] * 3 + [2e-4] * 3)
    pre_kf = preintegrate(acc, gyr, dt, bias_est)
    pre = pre_kf if mode == 0 else preintegrate(acc[-n_fr:], gyr[-n_fr:], dt, bias_est)
    Ckf = pre_kf[60:285].reshape(15, 15)

    def f64(x):
        return np.asarray(x, np.float32).astype(np.float64)

    R1, p1 = pose(t1); R0_, p0_ = pose(t0)
    # current frame: perturbed float32 state, Tcw composed in float32 as the Frame stores it
    Ri = R1 @ _rot(np.deg2rad(rng.uniform(-rot_deg, rot_deg, 3)))
    pi_ = p1 + rng.uniform(-trans, trans, 3)
    R32 = _polar32(Ri.astype(np.float32)); p32 = pi_.astype(np.float32)
    Rcw32 = (Rcb.astype(np.float32) @ R32.T).astype(np.float32)
    tcw32 = (Rcb.astype(np.float32) @ (-(R32.T @ p32)) + tcb.astype(np.float32)).astype(np.float32)
    pert = 0.0 if mode == 0 else 1.0
    Rp32 = _polar32((R0_ @ _rot(np.deg2rad(pert * rng.uniform(-0.1, 0.1, 3)))).astype(np.float32))
    pp32 = (p0_ + pert * rng.uniform(-0.005, 0.005, 3)).astype(np.float32)
    # observations of the current frame
    u = rng.uniform(20, w - 20, n_obs); v = rng.uniform(20, h - 20, n_obs); z = rng.uniform(0.6, 12.0, n_obs)
    Xc = np.stack([(u - cam["cx"]) / cam["fx"] * z, (v - cam["cy"]) / cam["fy"] * z, z], 1)
    Xw = (Xc @ Rbc.T + tbc) @ R1.T + p1
    octave = rng.integers(0, 8, n_obs)
    sigma = 1.2 ** octave
    uu = u + rng.normal(0, 1, n_obs) * sigma * 0.7
    vv = v + rng.normal(0, 1, n_obs) * sigma * 0.7
    ur = uu - cam["bf"] / z + rng.normal(0, 0.3, n_obs)
    bad = rng.random(n_obs) < outlier_frac
    uu[bad] += rng.uniform(8, 60, bad.sum()) * rng.choice([-1, 1], bad.sum())
    vv[bad] += rng.uniform(8, 60, bad.sum()) * rng.choice([-1, 1], bad.sum())
    mono = rng.random(n_obs) < mono_frac
    ur[mono] = -1.0
    if prior_H is None:
        A = rng.normal(0, 1, (15, 15)) * np.sqrt([3e3] * 6 + [3e2] * 3 + [1e5] * 3 + [3e3] * 3)
        prior_H = A.T @ A / 15 + np.diag([1e4] * 6 + [1e3] * 3 + [1e6] * 3 + [1e4] * 3)
    prior_H = np.asarray(prior_H, np.float64).reshape(15, 15)
    return dict(
        mode=int(mode), n_obs=int(n_obs), n_rounds=int(n_rounds), rec_init=int(rec_init),
        fx=np.float32(cam["fx"]), fy=np.float32(cam["fy"]), cx=np.float32(cam["cx"]), cy=np.float32(cam["cy"]),
        bf=np.float32(cam["bf"]), Rcb=f64(Rcb).ravel(), tcb=f64(tcb), tbc=f64(tbc),
        Rwb=f64(R32).ravel(), twb=f64(p32), Rcw=f64(Rcw32).ravel(), tcw=f64(tcw32),
        vel=f64(vel(t1) + rng.uniform(-0.05, 0.05, 3)), bg=f64(bias_est[3:]), ba=f64(bias_est[:3]),
        p_Rwb=f64(Rp32).ravel(), p_twb=f64(pp32), p_vel=f64(vel(t0) + pert * rng.uniform(-0.01, 0.01, 3)),
        p_bg=f64(bias_est[3:]), p_ba=f64(bias_est[:3]),
        pre=np.ascontiguousarray(pre, np.float32),
        rw_Cg=np.ascontiguousarray(Ckf[9:12, 9:12], np.float32).ravel(), rw_Ca=np.ascontiguousarray(Ckf[12:15, 12:15], np.float32).ravel(),
        c_Rwb=f64(Rp32).ravel(), c_twb=f64(pp32), c_vwb=f64(vel(t0)), c_bg=f64(bias_est[3:]), c_ba=f64(bias_est[:3]),
        c_H=np.ascontiguousarray(prior_H, np.float64).ravel(),
        Xw=np.ascontiguousarray(np.asarray(Xw, np.float32), np.float64), uvr=np.ascontiguousarray(np.stack([uu, vv, ur], 1), np.float32),
        inv_sigma2=(1.0 / sigma ** 2).astype(np.float32), close=(z < 10.0).astype(np.uint8),
        truth=dict(Rwb=R1, twb=p1, vel=vel(t1), bg=bg_true, ba=ba_true, bad=bad))


def imu_calib_noise(noise=IMU_NOISE):
    """(ng, na, ngw, naw) as IMU::Calib receives them: densities scaled by sqrt(frequency) (Settings / Tracking)"""
    sf = np.sqrt(np.float32(noise["freq"]))
    return np.float32(noise["ng"]) * sf, np.float32(noise["na"]) * sf, np.float32(noise["ngw"]) / sf, np.float32(noise["naw"]) / sf


def imu_samples(rng, n, dt=1 / 200, w_scale=0.3):
    """n random (acc, gyro) samples -> (acc (n,3), gyr (n,3), rows (n,7) float32: ax ay az wx wy wz dt)"""
    acc = rng.normal(0, 1, (n, 3)) + [0, 0, 9.81]
    gyr = rng.normal(0, w_scale, (n, 3))
    return acc, gyr, np.concatenate([acc, gyr, np.full((n, 1), dt)], 1).astype(np.float32)


def lba_problem(seed=7000, n_kf=12, n_fixed=3, n_points=1500, **kw):
    """Synthetic Optimizer::LocalBundleAdjustment problem (reference src/Optimizer.cc:1588-2040): the non-inertial local
    BA of the RGB-D configuration.  Same flattened layout as ba_problem with vertex_se3 = 1: g2o::VertexSE3Expmap
    keyframes (kf_Rcw / kf_tcw), no inertial edges, 10 Levenberg iterations from g2o's default lambda, several fixed
    covisible keyframes (lFixedCameras)."""
    p = ba_problem(seed=seed, n_kf=n_kf + n_fixed - 1, n_points=n_points, b_large=False, **kw)
    nk = n_kf + n_fixed
    assert len(p["kf_Rcw"]) == nk
    # fixed keyframes were optimised before: they hold their ground-truth pose (float32, as the KeyFrame stores it)
    Rcb32, tcb32 = p["Rcb"].reshape(3, 3).astype(np.float32), p["tcb"].astype(np.float32)
    for i in range(n_kf, nk):
        R32 = _polar32(p["truth"]["Rwb"][i].astype(np.float32)); t32 = p["truth"]["twb"][i].astype(np.float32)
        p["kf_Rwb"][i] = R32.astype(np.float64).ravel(); p["kf_twb"][i] = t32.astype(np.float64)
        p["kf_Rcw"][i] = (Rcb32 @ R32.T).astype(np.float32).astype(np.float64).ravel()
        p["kf_tcw"][i] = (Rcb32 @ (-(R32.T @ t32)) + tcb32).astype(np.float32).astype(np.float64)
    p.update(vertex_se3=1, n_opt_kf=n_kf, n_fixed_kf=n_fixed, n_inertial=0, iterations=10, lambda_init=0.0, b_large=0,
             in_kf1=np.zeros(0, np.int32), in_kf2=np.zeros(0, np.int32), in_pre=np.zeros((0, PRE_STRIDE), np.float32),
             in_downweight=np.zeros(0, np.uint8), kf_has_imu=np.zeros(nk, np.uint8), n_icp=0)
    return p


# ------------------------------------------------------------------------------------------------
# A closed-loop test sequence (BASELINE configs[4] in miniature): a textured wall seen by a moving RGB-D-inertial rig.
# Exact geometry, so landmarks, depth and IMU samples are mutually consistent.
# ------------------------------------------------------------------------------------------------
def vio_sequence(seed=8000, n_frames=20, w=640, h=480, fps=30, imu_rate=200, wall_x=3.0, ppm=260.0):
    """-> dict(frames (n,h,w) u8, depth (n,h,w) f32, stamps, Rwb/twb/vel truth per frame, imu rows per frame interval
    [(m,7) ax ay az wx wy wz dt], calibration).  Body frame: x forward, y left, z up; camera z forward."""
    rng = np.random.default_rng(seed)
    cam = G1_CAM
    Rbc = np.array([[0, 0, 1.0], [-1, 0, 0], [0, -1, 0]])
    tbc = np.array([0.05, 0.02, 0.01])
    tex = scene(seed, 1400, 1800, nrect=900)
    y0, z0 = -tex.shape[1] / (2 * ppm), -tex.shape[0] / (2 * ppm)
    g = np.array([0, 0, -9.81])
    ph = rng.uniform(0, 6.28, 4)

    def pose(t):
        yaw = 0.10 * np.sin(1.1 * t + ph[0]); pitch = 0.04 * np.sin(0.9 * t + ph[1]); roll = 0.03 * np.cos(1.3 * t + ph[2])
        R = _rot(np.array([0, 0, yaw])) @ _rot(np.array([0, pitch, 0])) @ _rot(np.array([roll, 0, 0]))
        p = np.array([0.15 * np.sin(0.8 * t + ph[3]), 0.35 * np.sin(0.7 * t), 0.10 * np.sin(0.9 * t)])
        return R, p

    hh = 1e-4
    vs, us = np.mgrid[0:h, 0:w].astype(np.float64)
    dc = np.stack([(us - cam["cx"]) / cam["fx"], (vs - cam["cy"]) / cam["fy"], np.ones_like(us)], -1)
    frames = np.zeros((n_frames, h, w), np.uint8); depth = np.zeros((n_frames, h, w), np.float32)
    Rs, ps, vels, stamps = [], [], [], []
    for k in range(n_frames):
        t = k / fps
        R, p = pose(t)
        Rwc = R @ Rbc; twc = R @ tbc + p
        dw = dc @ Rwc.T
        s = (wall_x - twc[0]) / dw[..., 0]
        P = twc + s[..., None] * dw
        tu = (P[..., 1] - y0) * ppm; tv = (P[..., 2] - z0) * ppm
        u0 = np.clip(np.floor(tu).astype(np.int32), 0, tex.shape[1] - 2); v0 = np.clip(np.floor(tv).astype(np.int32), 0, tex.shape[0] - 2)
        fu = (tu - u0).astype(np.float32); fv = (tv - v0).astype(np.float32)
        img = (tex[v0, u0] * (1 - fu) + tex[v0, u0 + 1] * fu) * (1 - fv) + (tex[v0 + 1, u0] * (1 - fu) + tex[v0 + 1, u0 + 1] * fu) * fv
        img = img + rng.normal(0, 1.0, img.shape).astype(np.float32)
        frames[k] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        depth[k] = s.astype(np.float32)
        Rs.append(R); ps.append(p); vels.append((pose(t + hh)[1] - pose(t - hh)[1]) / (2 * hh)); stamps.append(t)
    # IMU: mid-point samples between consecutive frames (the last interval of each frame is shortened to end on the frame)
    bg_true = np.array([0.002, -0.001, 0.0015]); ba_true = np.array([0.02, -0.03, 0.01])
    imu = []
    for k in range(1, n_frames):
        ta, tb = stamps[k - 1], stamps[k]
        edges = list(np.arange(ta, tb - 1e-9, 1.0 / imu_rate)) + [tb]
        rows = []
        for a, b in zip(edges[:-1], edges[1:]):
            tm = 0.5 * (a + b)
            R0, p0 = pose(tm); Rp, pp = pose(tm + hh); Rm, pm = pose(tm - hh)
            a_w = (pp - 2 * p0 + pm) / (hh * hh)
            dRm = R0.T @ (Rp - Rm) / (2 * hh)
            gyr = np.array([dRm[2, 1], dRm[0, 2], dRm[1, 0]]) + bg_true + rng.normal(0, 2e-4, 3)
            acc = R0.T @ (a_w - g) + ba_true + rng.normal(0, 2e-3, 3)
            rows.append(np.concatenate([acc, gyr, [b - a]]))
        imu.append(np.array(rows, np.float32))
    return dict(frames=frames, depth=depth, stamps=np.array(stamps), Rwb=np.array(Rs), twb=np.array(ps), vel=np.array(vels), imu=imu,
                bg=bg_true, ba=ba_true, Rbc=Rbc, tbc=tbc, cam=cam, gravity=g)


def room_sequence(seed=8200, n_frames=30, w=640, h=480, fps=30, imu_rate=200, ppm=260.0, depth_noise=0.0):
    """Like vio_sequence, but the camera looks into a room CORNER -- front wall x = 3, side wall y = 1.3, floor z = -0.6, each
    textured -- so that the depth clouds constrain all six degrees of freedom (a single wall lets GICP slide along it).
    -> the vio_sequence dict plus odom rows per frame interval (vx vy vz in the body frame, 30 Hz)."""
    rng = np.random.default_rng(seed)
    cam = G1_CAM
    Rbc = np.array([[0, 0, 1.0], [-1, 0, 0], [0, -1, 0]])
    tbc = np.array([0.05, 0.02, 0.01])
    tex = [scene(seed + i, 1400, 1800, nrect=900) for i in range(3)]
    g = np.array([0, 0, -9.81])
    ph = rng.uniform(0, 6.28, 4)
    # plane = (axis, value, in-plane axes, texture origin in metres)
    planes = [(0, 3.0, (1, 2), (-3.5, -2.7)), (1, 1.3, (0, 2), (-1.0, -2.7)), (2, -0.6, (0, 1), (-1.0, -3.5))]

    def pose(t):
        yaw = 0.08 * np.sin(1.1 * t + ph[0]); pitch = 0.04 * np.sin(0.9 * t + ph[1]); roll = 0.03 * np.cos(1.3 * t + ph[2])
        R = _rot(np.array([0, 0, yaw])) @ _rot(np.array([0, pitch, 0])) @ _rot(np.array([roll, 0, 0]))
        p = np.array([0.15 * np.sin(0.8 * t + ph[3]), 0.25 * np.sin(0.7 * t), 0.08 * np.sin(0.9 * t)])
        return R, p

    hh = 1e-4
    vs, us = np.mgrid[0:h, 0:w].astype(np.float64)
    dc = np.stack([(us - cam["cx"]) / cam["fx"], (vs - cam["cy"]) / cam["fy"], np.ones_like(us)], -1)
    frames = np.zeros((n_frames, h, w), np.uint8); depth = np.zeros((n_frames, h, w), np.float32)
    Rs, ps, vels, stamps = [], [], [], []
    for k in range(n_frames):
        t = k / fps
        R, p = pose(t)
        Rwc = R @ Rbc; twc = R @ tbc + p
        dw = dc @ Rwc.T
        best = np.full((h, w), np.inf); img = np.zeros((h, w), np.float32)
        for (ax, val, (a, b), (oa, ob)), tx in zip(planes, tex):
            with np.errstate(divide="ignore", invalid="ignore"):
                sp = (val - twc[ax]) / dw[..., ax]
            hit = (sp > 0.05) & (sp < best)
            P = twc + sp[..., None] * dw
            tu = np.nan_to_num((P[..., a] - oa) * ppm, nan=-1.0, posinf=-1.0, neginf=-1.0)
            tv = np.nan_to_num((P[..., b] - ob) * ppm, nan=-1.0, posinf=-1.0, neginf=-1.0)
            hit &= (tu >= 0) & (tu < tx.shape[1] - 1) & (tv >= 0) & (tv < tx.shape[0] - 1)
            tu = np.clip(tu, -1.0, tx.shape[1]); tv = np.clip(tv, -1.0, tx.shape[0])
            u0 = np.clip(np.floor(tu), 0, tx.shape[1] - 2).astype(np.int32); v0 = np.clip(np.floor(tv), 0, tx.shape[0] - 2).astype(np.int32)
            fu = (tu - u0).astype(np.float32); fv = (tv - v0).astype(np.float32)
            col = (tx[v0, u0] * (1 - fu) + tx[v0, u0 + 1] * fu) * (1 - fv) + (tx[v0 + 1, u0] * (1 - fu) + tx[v0 + 1, u0 + 1] * fu) * fv
            img = np.where(hit, col, img); best = np.where(hit, sp, best)
        img = img + rng.normal(0, 1.0, img.shape).astype(np.float32)
        frames[k] = np.clip(np.rint(img), 0, 255).astype(np.uint8)
        d = np.where(np.isfinite(best), best, 0.0)
        if depth_noise > 0:
            d = np.where(d > 0, d + rng.normal(0, depth_noise, d.shape), 0.0)
        depth[k] = d.astype(np.float32)
        Rs.append(R); ps.append(p); vels.append((pose(t + hh)[1] - pose(t - hh)[1]) / (2 * hh)); stamps.append(t)
    bg_true = np.array([0.002, -0.001, 0.0015]); ba_true = np.array([0.02, -0.03, 0.01])
    imu, odom = [], []
    for k in range(1, n_frames):
        ta, tb = stamps[k - 1], stamps[k]
        edges = list(np.arange(ta, tb - 1e-9, 1.0 / imu_rate)) + [tb]
        rows = []
        for a, b in zip(edges[:-1], edges[1:]):
            tm = 0.5 * (a + b)
            R0, p0 = pose(tm); Rp, pp = pose(tm + hh); Rm, pm = pose(tm - hh)
            a_w = (pp - 2 * p0 + pm) / (hh * hh)
            dRm = R0.T @ (Rp - Rm) / (2 * hh)
            gyr = np.array([dRm[2, 1], dRm[0, 2], dRm[1, 0]]) + bg_true + rng.normal(0, 2e-4, 3)
            acc = R0.T @ (a_w - g) + ba_true + rng.normal(0, 2e-3, 3)
            rows.append(np.concatenate([acc, gyr, [b - a]]))
        imu.append(np.array(rows, np.float32))
        Rm_, _ = pose(0.5 * (ta + tb))
        odom.append((Rm_.T @ ((ps[k] - ps[k - 1]) / (tb - ta))).astype(np.float32).reshape(1, 3))
    return dict(frames=frames, depth=depth, stamps=np.array(stamps), Rwb=np.array(Rs), twb=np.array(ps), vel=np.array(vels), imu=imu,
                odom=odom, bg=bg_true, ba=ba_true, Rbc=Rbc, tbc=tbc, cam=cam, gravity=g)

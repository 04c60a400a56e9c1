"""A minimal closed-loop visual-inertial odometry chain over the library's entry points, for end-to-end checks
(BASELINE.json: "ATE vs ref").  It is NOT the reference's Tracking state machine (out of scope): frame 0
seeds the landmarks from its depth map; every later frame runs, exactly as TrackLocalMap's inertial branch would,
    fbKltTracking(previous image -> current image)                      ORBmatcher.cc:2186
    IMU preintegration of the samples since the previous frame           ImuTypes.cc:184
    IMU prediction of the body state (Tracking::PredictStateIMU)         Tracking.cc, mpImuPreintegratedFrame
    PoseInertialOptimizationLastKeyFrame (frame 1) / LastFrame (later)   Optimizer.cc:5899 / 6762
and hands the marginalised 15x15 prior to the next frame.  The numerical work goes through a `backend` with three
callables, so the same chain runs on the CUDA library and, in the tests, on the CPU oracle:
    backend.fb_klt(prev_img, cur_img, kps, priors) -> (tracked positions, status)
    backend.preintegrate(rows, bias6) -> 292-float record
    backend.pose_inertial(problem dict) -> result dict"""
import numpy as np


def select_landmarks(img, depth, n_max=600, cell=24, border=40):
    """deterministic corner-ish pixels: the strongest min(|Ix|,|Iy|)-type response of every grid cell"""
    f = img.astype(np.float32)
    ix = np.abs(f[1:-1, 2:] - f[1:-1, :-2]); iy = np.abs(f[2:, 1:-1] - f[:-2, 1:-1])
    resp = np.zeros_like(f); resp[1:-1, 1:-1] = np.minimum(ix, iy)
    h, w = img.shape
    pts = []
    for y0 in range(border, h - border - cell, cell):
        for x0 in range(border, w - border - cell, cell):
            blk = resp[y0:y0 + cell, x0:x0 + cell]
            j = int(np.argmax(blk))
            if blk.flat[j] > 12:
                pts.append((x0 + j % cell, y0 + j // cell, float(blk.flat[j])))
    pts.sort(key=lambda p: -p[2])
    pts = np.array([(p[0], p[1]) for p in pts[:n_max]], np.float32)
    z = depth[pts[:, 1].astype(int), pts[:, 0].astype(int)]
    return pts, z


def run_chain(seq, backend, n_frames=None):
    """-> dict(twb (n,3), Rwb (n,3,3), n_tracked, n_inliers): the estimated body trajectory (frame 0 = ground truth)"""
    from . import synth
    cam, Rbc, tbc, g = seq["cam"], seq["Rbc"], seq["tbc"], seq["gravity"]
    Rcb = Rbc.T; tcb = -Rcb @ tbc
    n = n_frames or len(seq["frames"])
    f32 = lambda x: np.asarray(x, np.float32).astype(np.float64)
    ng, na, ngw, naw = synth.imu_calib_noise()
    # frame 0: ground-truth state, landmarks from its depth map
    R, p, v = seq["Rwb"][0].copy(), seq["twb"][0].copy(), seq["vel"][0].copy()
    bg, ba = np.zeros(3), np.zeros(3)
    kps, z = select_landmarks(seq["frames"][0], seq["depth"][0])
    Xc = np.stack([(kps[:, 0] - cam["cx"]) / cam["fx"] * z, (kps[:, 1] - cam["cy"]) / cam["fy"] * z, z], 1)
    Xw = f32((Xc @ Rbc.T + tbc) @ R.T + p)
    alive = np.ones(len(kps), bool)
    out = dict(twb=[p.copy()], Rwb=[R.copy()], n_tracked=[int(alive.sum())], n_inliers=[int(alive.sum())])
    prior = None
    prevR, prevp, prevv, prevbg, prevba = R, p, v, bg, ba
    for k in range(1, n):
        a, b = seq["frames"][k - 1], seq["frames"][k]
        idx = np.nonzero(alive)[0]
        pr, st = backend.fb_klt(a, b, kps[idx], kps[idx])
        alive[idx[~st]] = False
        kps[idx[st]] = pr[st]
        idx = idx[st]
        rows = seq["imu"][k - 1]
        rec = backend.preintegrate(rows, np.concatenate([prevba, prevbg]))
        dR, dV, dP, dT = rec[0:9].reshape(3, 3).astype(np.float64), rec[9:12].astype(np.float64), rec[12:15].astype(np.float64), float(rec[285])
        # Tracking::PredictStateIMU with the frame-to-frame preintegration
        R0 = prevR @ dR
        p0 = prevp + prevv * dT + 0.5 * g * dT * dT + prevR @ dP
        v0 = prevv + g * dT + prevR @ dV
        R32 = synth._polar32(R0.astype(np.float32)); p32 = p0.astype(np.float32)
        Rcw32 = (Rcb.astype(np.float32) @ R32.T).astype(np.float32)
        tcw32 = (Rcb.astype(np.float32) @ (-(R32.T @ p32)) + tcb.astype(np.float32)).astype(np.float32)
        u, vv = kps[idx, 0], kps[idx, 1]
        zz = seq["depth"][k][np.clip(np.rint(vv).astype(int), 0, 479), np.clip(np.rint(u).astype(int), 0, 639)]
        ur = (u - np.float32(cam["bf"]) / zz).astype(np.float32)
        C = rec[60:285].reshape(15, 15)
        prob = dict(mode=0 if prior is None else 1, n_obs=len(idx), n_rounds=4, rec_init=0,
                    fx=np.float32(cam["fx"]), fy=np.float32(cam["fy"]), cx=np.float32(cam["cx"]), cy=np.float32(cam["cy"]), bf=np.float32(cam["bf"]),
                    Rcb=f32(Rcb).ravel(), tcb=f32(tcb), tbc=f32(tbc),
                    Rwb=f32(R32).ravel(), twb=f32(p32), Rcw=f32(Rcw32).ravel(), tcw=f32(tcw32), vel=f32(v0), bg=f32(prevbg), ba=f32(prevba),
                    p_Rwb=f32(prevR).ravel(), p_twb=f32(prevp), p_vel=f32(prevv), p_bg=f32(prevbg), p_ba=f32(prevba),
                    pre=rec, rw_Cg=np.ascontiguousarray(C[9:12, 9:12], np.float32).ravel(), rw_Ca=np.ascontiguousarray(C[12:15, 12:15], np.float32).ravel(),
                    c_Rwb=f32(prevR).ravel(), c_twb=f32(prevp), c_vwb=f32(prevv), c_bg=f32(prevbg), c_ba=f32(prevba),
                    c_H=(np.zeros(225) if prior is None else np.asarray(prior, np.float64).ravel()),
                    Xw=np.ascontiguousarray(Xw[idx]), uvr=np.ascontiguousarray(np.stack([u, vv, ur], 1), np.float32),
                    inv_sigma2=np.ones(len(idx), np.float32), close=np.ones(len(idx), np.uint8))
        r = backend.pose_inertial(prob)
        alive[idx[r["outlier"]]] = False
        # the reference narrows the optimised state to float when it writes it back (SetImuPoseVelocity, IMU::Bias)
        prevR, prevp, prevv = f32(r["Rwb"]), f32(r["twb"]), f32(r["vel"])
        prevbg, prevba = f32(r["bg"]), f32(r["ba"])
        prior = r["H"]
        out["twb"].append(prevp.copy()); out["Rwb"].append(prevR.copy())
        out["n_tracked"].append(int(len(idx))); out["n_inliers"].append(int(r["n_inliers"]))
    out["twb"] = np.array(out["twb"]); out["Rwb"] = np.array(out["Rwb"])
    return out


class CudaBackend:
    """the library's own entry points"""

    def __init__(self, max_points=1024):
        from .klt import KltTracker
        from .pose_inertial import PoseInertialOptimizer
        self._klt = KltTracker(max_points=max_points, max_batch=1)
        self._pin = PoseInertialOptimizer(max_obs=max_points, max_batch=1)

    def fb_klt(self, a, b, kps, priors):
        return self._klt.fbKltTracking(a, b, kps, priors)

    def preintegrate(self, rows, bias6):
        from . import imu, synth
        return imu.preintegrate_batch([rows], [bias6], *synth.imu_calib_noise())[0]

    def pose_inertial(self, prob):
        return self._pin.optimize_batch([prob])[0]

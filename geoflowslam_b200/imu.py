"""Host-side mirror of IMU::Preintegrated (reference include/ImuTypes.h:172-274, src/ImuTypes.cc:163-246) over the C
ABI, batched over independent intervals, and of the TUM trajectory format System::SaveTrajectoryTUM writes
(src/System.cc:1083-1144)."""
import numpy as np

from . import _lib
from ._lib import check, ptr

PRE_STRIDE = 292
FIELDS = dict(dR=(0, 9), dV=(9, 12), dP=(12, 15), JRg=(15, 24), JVg=(24, 33), JVa=(33, 42), JPg=(42, 51), JPa=(51, 60), C=(60, 285),
              dT=(285, 286), b=(286, 292))


def preintegrate_batch(measurements, biases, ng, na, ngw, naw, stream=None):
    """measurements: list of (m_i, 7) arrays (ax ay az wx wy wz dt); biases: (n, 6) bax..bwz -> (n, 292) float32 records"""
    _lib.require_device()
    n = len(measurements)
    rows = [np.ascontiguousarray(m, np.float32).reshape(-1, 7) for m in measurements]
    off = np.zeros(n + 1, np.int32)
    off[1:] = np.cumsum([len(r) for r in rows])
    meas = np.ascontiguousarray(np.concatenate(rows, 0) if off[-1] else np.zeros((1, 7), np.float32))
    b = np.ascontiguousarray(biases, np.float32).reshape(n, 6)
    out = np.zeros((n, PRE_STRIDE), np.float32)
    check(_lib.lib().gfs_imu_preintegrate_batch(stream, ptr(meas), ptr(off), ptr(b), n, float(ng), float(na), float(ngw), float(naw), ptr(out)))
    return out


def unpack(record):
    r = np.asarray(record, np.float32)
    out = {k: r[a:b].copy() for k, (a, b) in FIELDS.items()}
    for k in ("dR", "JRg", "JVg", "JVa", "JPg", "JPa"):
        out[k] = out[k].reshape(3, 3)
    out["C"] = out["C"].reshape(15, 15)
    out["dT"] = float(out["dT"][0])
    return out


# ---- trajectory files
def tum_line(t_seconds, twc, q_xyzw):
    """One line of System::SaveTrajectoryTUM (src/System.cc:1136-1138): `fixed`, the timestamp in ms with 4 decimals,
    then the float32 translation and quaternion (x y z w) with 9 decimals."""
    v = [float(np.float32(x)) for x in list(twc) + list(q_xyzw)]
    return "%.4f %s" % (float(t_seconds) * 1e3, " ".join("%.9f" % x for x in v))


def save_trajectory_tum(path, stamps, Twc_list):
    """Twc_list: iterable of 4x4 camera-to-world matrices (the Tcw^-1 the reference composes from the reference keyframe
    and the relative frame pose, :1121-1134); quaternion as Eigen::Quaternionf(R) normalised."""
    from .synth import _quat_from_R
    with open(path, "w") as f:
        for t, T in zip(stamps, Twc_list):
            T = np.asarray(T, np.float64)
            q = _quat_from_R(T[:3, :3])  # w x y z, w >= 0
            f.write(tum_line(t, T[:3, 3], [q[1], q[2], q[3], q[0]]) + "\n")


def ate_rmse(est_xyz, gt_xyz):
    """Absolute trajectory error after the closed-form rigid alignment (Horn / Umeyama without scale) the reference's
    evaluation script applies (evaluation/PoseEvaluatorTUM.py:519-535): RMSE of the translational residuals."""
    P = np.asarray(est_xyz, np.float64); Q = np.asarray(gt_xyz, np.float64)
    mp, mq = P.mean(0), Q.mean(0)
    U, _, Vt = np.linalg.svd((Q - mq).T @ (P - mp))
    S = np.eye(3)
    if np.linalg.det(U) * np.linalg.det(Vt) < 0:
        S[2, 2] = -1
    R = U @ S @ Vt
    res = (P - mp) @ R.T + mq - Q
    return float(np.sqrt((res ** 2).sum(1).mean()))

"""CPU oracle of IMU::Preintegrated (oracle/imu_oracle.cpp) and the trajectory-format helpers.  The reference has no
tests for this path (parity unpinned): the C++ restatement is checked against the independent numpy restatement
(geoflowslam_b200.synth.preintegrate, SVD-based NormalizeRotation) and against closed-form integrals."""
import numpy as np

from geoflowslam_b200 import imu, synth
from oracle import oracle as O


def test_matches_numpy_restatement():
    rng = np.random.default_rng(0)
    bias = np.array([0.02, -0.03, 0.01, 0.002, -0.001, 0.0015])
    for n, ws in ((80, 0.3), (7, 0.3), (200, 1.5), (30, 1e-3)):  # the last one runs the d < eps branch of IntegratedRotation
        acc, gyr, meas = synth.imu_samples(rng, n, w_scale=ws)
        ref = synth.preintegrate(acc, gyr, 1 / 200, bias)
        got = O.imu_preintegrate(meas, bias, *synth.imu_calib_noise())
        for k, (a, b) in imu.FIELDS.items():
            scale = max(np.abs(ref[a:b]).max(), 1e-12)
            assert np.abs(got[a:b] - ref[a:b]).max() <= 5e-6 * scale, k


def test_constant_motion_closed_form():
    # constant acceleration a (bias-free), no rotation: dV = a T, dP = a T^2 / 2, dR = I, JVa = -T I, JPa = -T^2/2 I
    n, dt = 100, 0.005
    a = np.array([0.3, -0.2, 9.81], np.float32)
    meas = np.tile(np.concatenate([a, [0, 0, 0], [dt]]).astype(np.float32), (n, 1))
    r = imu.unpack(O.imu_preintegrate(meas, np.zeros(6), *synth.imu_calib_noise()))
    T = n * dt
    assert np.allclose(r["dR"], np.eye(3), atol=1e-7) and abs(r["dT"] - T) < 1e-6
    assert np.allclose(r["dV"], a * T, rtol=1e-5) and np.allclose(r["dP"], a * T * T / 2, rtol=1e-5)
    assert np.allclose(r["JVa"], -T * np.eye(3), atol=1e-6) and np.allclose(r["JPa"], -T * T / 2 * np.eye(3), atol=1e-6)
    assert np.allclose(r["JRg"], -T * np.eye(3), atol=1e-6)
    C = r["C"]
    assert np.allclose(C, C.T, atol=1e-12) and np.linalg.eigvalsh(C[:9, :9].astype(np.float64)).min() > 0
    ng, na, ngw, naw = synth.imu_calib_noise()
    assert np.allclose(np.diag(C)[9:12], n * ngw * ngw, rtol=1e-5) and np.allclose(np.diag(C)[12:], n * naw * naw, rtol=1e-5)
    # empty interval = Initialize()
    z = imu.unpack(O.imu_preintegrate(np.zeros((0, 7), np.float32), [1, 2, 3, 4, 5, 6], *synth.imu_calib_noise()))
    assert np.array_equal(z["dR"], np.eye(3, dtype=np.float32)) and z["dT"] == 0 and not z["C"].any() and list(z["b"]) == [1, 2, 3, 4, 5, 6]


def test_tum_format_and_ate(tmp_path):
    # C++: f << fixed << setprecision(4) << t * 1e3 << " " << setprecision(9) << twc(0) ... q.w()
    line = imu.tum_line(1403636579.763555527, [0.1, -2.5, 3.0], [0.0, 0.0, np.sin(0.25), np.cos(0.25)])
    assert line == "1403636579763.5554 0.100000001 -2.500000000 3.000000000 0.000000000 0.000000000 0.247403964 0.968912423"
    rng = np.random.default_rng(1)
    gt = np.cumsum(rng.normal(0, 0.1, (50, 3)), 0)
    R = synth._rot(np.array([0.3, -0.2, 0.5])); t = np.array([1.0, 2.0, -0.5])
    est = gt @ R.T + t
    assert imu.ate_rmse(est, gt) < 1e-12
    assert abs(imu.ate_rmse(est + np.array([0.01, 0, 0]) * (np.arange(50) % 2)[:, None], gt) - 0.005) < 1e-3
    Ts = [np.block([[R, p[:, None]], [np.zeros((1, 3)), np.ones((1, 1))]]) for p in gt[:3]]
    imu.save_trajectory_tum(tmp_path / "traj.txt", [0.1, 0.2, 0.3], Ts)
    rows = [l.split() for l in open(tmp_path / "traj.txt")]
    assert len(rows) == 3 and rows[0][0] == "100.0000" and all(len(r) == 8 for r in rows)
    q = np.array(rows[0][4:], float)
    assert abs(np.linalg.norm(q) - 1) < 1e-6

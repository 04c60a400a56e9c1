"""Sequence plumbing (geoflowslam_b200/dataset.py): the TUM-style layout of BASELINE configs[0] / [4] written and read back
the way the reference's example drivers read it (Examples/RGB-D/rgbd_tum.cc, Examples/RGB-D-Inertial/rgbd_inertial.cc)."""
import numpy as np


def _sequence(n=6):
    from geoflowslam_b200 import synth
    seq = synth.vio_sequence(seed=8100, n_frames=n, w=160, h=120)
    # absolute-time IMU rows [t, acc, gyro]: every sample stamped at the END of its interval
    rows, t = [], float(seq["stamps"][0])
    for k, block in enumerate(seq["imu"]):
        t = float(seq["stamps"][k])
        for r in block:
            t += float(r[6])
            rows.append([t, *r[0:3], *r[3:6]])
    imu = np.array(rows)
    odom = np.array([[s, 0.1, 0.0, 0.0, 0.0, 0.0, 0.01] for s in np.arange(0.0, float(seq["stamps"][-1]) + 1e-9, 1 / 30.0)])
    return seq, imu, odom


def test_inertial_sequence_round_trip(tmp_path):
    from geoflowslam_b200 import dataset as D
    seq, imu, odom = _sequence()
    root = str(tmp_path / "seq")
    assoc = D.write_sequence(root, seq["stamps"], seq["frames"], seq["depth"], imu=imu, odom=odom, inertial=True)
    rgb, dep, ts = D.load_images(assoc, inertial=True)
    assert len(rgb) == len(seq["frames"]) and np.allclose(ts, seq["stamps"], atol=1e-7)
    for i in (0, len(rgb) - 1):
        gray, d = D.read_frame(root, rgb[i], dep[i])
        assert np.array_equal(gray, seq["frames"][i])                       # gray -> BGR -> gray is lossless
        assert d.dtype == np.float32 and np.abs(d - seq["depth"][i]).max() <= 0.5e-3 + 1e-6   # millimetre quantisation
    t_imu, acc, gyr = D.load_imu(root + "/imu/imu.txt")
    assert np.allclose(t_imu, imu[:, 0], atol=1e-7) and np.allclose(acc, imu[:, 1:4], rtol=1e-6) and np.allclose(gyr, imu[:, 4:7], rtol=1e-6)
    t_od, vpos = D.load_odom(root + "/imu/odom.txt")
    assert vpos.shape == (len(odom), 3) and np.allclose(vpos[:, 0], 0.1)
    # frame loop: the first frame has no samples and is skipped; every later frame gets the samples up to its stamp
    groups = D.frame_measurements(ts, t_imu, t_od)
    assert [g[0] for g in groups] == list(range(1, len(ts)))
    for (ni, (a, b), (oa, ob)) in groups:
        assert b - a == len(seq["imu"][ni - 1]) and np.all(t_imu[a:b] <= ts[ni] + 1e-12) and (a == 0 or t_imu[a - 1] <= ts[ni - 1] + 1e-12)
        assert np.all(t_od[oa:ob] <= ts[ni] + 1e-12)
    assert groups[-1][1][1] == len(t_imu)


def test_tum_flavour_and_sparse_imu(tmp_path):
    from geoflowslam_b200 import dataset as D
    seq, imu, _ = _sequence(4)
    root = str(tmp_path / "tum")
    depth = seq["depth"].copy(); depth[0, :5, :5] = np.nan; depth[0, 5:8, :5] = -1.0
    assoc = D.write_sequence(root, seq["stamps"] + 1305031102.0, seq["frames"], depth, inertial=False)
    rgb, dep, ts = D.load_images(assoc, inertial=False)
    assert np.allclose(ts, seq["stamps"] + 1305031102.0, atol=1e-4) and rgb[0].startswith("rgb/1305031102.0000")
    _, d = D.read_frame(root, rgb[0], dep[0])
    assert np.all(d[:8, :5] == 0)                                            # invalid depth stays 0 (Frame.cc:601 skips d <= 0)
    # frames with fewer than three IMU samples are not tracked (rgbd_inertial.cc:171)
    g = D.frame_measurements(np.array([0.0, 0.1, 0.2]), np.array([0.01, 0.02, 0.03, 0.15, 0.16]))
    assert [x[0] for x in g] == [1]
    # rgbd_inertial.cc:80-85: the samples at or before frame 0 are skipped, then `first_imu--` / `first_odom--`: the LAST
    # sample at or before frame 0 opens frame 1's group (PreintegrateIMU interpolates from it)
    t_imu = np.array([-0.02, -0.01, 0.0, 0.03, 0.06, 0.09, 0.12, 0.15, 0.18])
    t_od = np.array([-0.05, 0.0, 0.05, 0.1, 0.15, 0.2])
    g = D.frame_measurements(np.array([0.0, 0.1, 0.2]), t_imu, t_od)
    assert g == [(1, (2, 6), (1, 4)), (2, (6, 9), (4, 6))]
    # two samples after frame 0 + the one at its stamp make three: the frame is tracked (it was dropped before the fix)
    g = D.frame_measurements(np.array([0.0, 0.1]), np.array([0.0, 0.04, 0.08]))
    assert g == [(1, (0, 3), (0, 0))]


def test_config0_plumbing_chain_from_disk(tmp_path):
    """configs[0] / [4] plumbing end to end on the CPU: sequence -> disk -> loaders / frame loop -> closed-loop chain on the
    oracle -> SaveTrajectoryTUM file + ATE (scripts/config0_plumbing.py)."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location(
        "config0_plumbing", os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts", "config0_plumbing.py"))
    mod = importlib.util.module_from_spec(spec); spec.loader.exec_module(mod)
    r = mod.run(8, str(tmp_path / "seq"))
    assert r["ate"] < 2e-3 and r["est"]["n_tracked"][-1] > 0.9 * r["est"]["n_tracked"][0]
    # the inertial samples came back from imu.txt grouped exactly as they were generated
    assert [len(b) for b in r["seq"]["imu"]] == [len(b) for b in r["truth"]["imu"]]
    assert np.allclose(np.concatenate(r["seq"]["imu"])[:, :6], np.concatenate(r["truth"]["imu"])[:, :6], rtol=1e-6, atol=1e-7)
    assert np.allclose(np.concatenate(r["seq"]["imu"])[:, 6], np.concatenate(r["truth"]["imu"])[:, 6], atol=2e-7)
    lines = open(r["trajectory"]).read().strip().split("\n")
    assert len(lines) == 8 and all(len(ln.split()) == 8 for ln in lines)
    assert abs(float(lines[1].split()[0]) - 1e3 / 30) < 1e-3                 # stamps in ms, 4 decimals (System.cc:1136)

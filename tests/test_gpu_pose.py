"""GPU parity: batched PoseOptimization (through the C ABI) vs the CPU oracle.  Rounds and outlier flags
equal, Levenberg iteration counts within one (rounding at convergence), pose within 1e-8, chi2 within
1e-4 relative (north_star asks 1e-4)."""
import numpy as np
import pytest

from geoflowslam_b200 import synth

pytestmark = pytest.mark.gpu


def _check(g, o):
    assert g["rounds_done"] == o["rounds_done"]
    # g2o stops a round on `rho == 0` / three relative decreases below 1e-3: at convergence the chi2
    # differences are pure rounding, so the parallel summation order may stop one iteration earlier or
    # later than the sequential oracle -- with the same pose to 1e-8 (checked below)
    assert all(abs(a - b) <= 1 for a, b in zip(g["lm_iterations"], o["lm_iterations"]))
    assert np.allclose(g["q_wxyz"], o["q_wxyz"], atol=1e-8) and np.allclose(g["t"], o["t"], atol=1e-8)
    assert np.allclose(g["chi2"], o["chi2"], rtol=1e-4, atol=1e-5)
    d = g["outlier"] != o["outlier"]
    if d.any():  # only edges sitting on the threshold may differ
        c = o["chi2"][d]
        assert np.all(np.minimum(np.abs(c - 5.991), np.abs(c - 7.815)) < 1e-3)
    else:
        assert g["n_inliers"] == o["n_inliers"] and g["n_bad"] == o["n_bad"] and g["n_good"] == o["n_good"]
        assert np.isclose(g["avg_reproj_error"], o["avg_reproj_error"], rtol=1e-4)


def test_single_frame_matches_oracle():
    from geoflowslam_b200 import PoseOptimizer
    from oracle import oracle as O
    p = synth.pose_problem(5000)
    opt = PoseOptimizer(max_obs=1024, max_batch=1)
    g = opt.PoseOptimization(p)
    _check(g, O.pose_optimize(p))
    assert opt.last_launches() == 1
    g2 = opt.PoseOptimization(p)     # fixed summation order: same bits again
    assert np.array_equal(g["q_wxyz"], g2["q_wxyz"]) and np.array_equal(g["chi2"], g2["chi2"])


def test_batch_of_ragged_frames_matches_oracle():
    from geoflowslam_b200 import PoseOptimizer
    from oracle import oracle as O
    probs = [synth.pose_problem(5100 + i, n_obs=n, outlier_frac=f, mono_frac=m)
             for i, (n, f, m) in enumerate([(400, 0.1, 0.2), (37, 0.0, 0.0), (1000, 0.3, 0.5), (8, 0.0, 1.0), (2, 0.0, 0.0),
                                            (250, 0.6, 0.1), (129, 0.05, 1.0), (0, 0.0, 0.0)])]
    opt = PoseOptimizer(max_obs=1000, max_batch=8)
    gs = opt.optimize_batch(probs)
    for p, g in zip(probs, gs):
        _check(g, O.pose_optimize(p))
    assert gs[4]["n_inliers"] == 0 and gs[7]["n_inliers"] == 0 and gs[3]["rounds_done"] == 1


def test_capacity_errors_are_loud():
    from geoflowslam_b200 import GfsError, PoseOptimizer
    opt = PoseOptimizer(max_obs=100, max_batch=1)
    with pytest.raises(GfsError):
        opt.PoseOptimization(synth.pose_problem(5200, n_obs=101))
    with pytest.raises(GfsError):
        opt.optimize_batch([synth.pose_problem(5200, n_obs=10)] * 2)

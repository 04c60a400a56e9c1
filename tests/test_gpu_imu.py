"""GPU parity: batched IMU preintegration (through the C ABI) vs the CPU oracle, within 2e-6 of each field's scale
(float32 recurrences of up to 200 steps; the double sin / cos of the device and of glibc may differ in the last place)."""
import numpy as np
import pytest

from geoflowslam_b200 import imu, synth

pytestmark = pytest.mark.gpu


def test_batch_of_ragged_intervals_matches_oracle():
    from oracle import oracle as O
    rng = np.random.default_rng(3)
    lens = [80, 7, 200, 0, 1, 33, 100, 6]
    ws = [0.3, 0.3, 1.5, 0.3, 0.3, 1e-3, 0.8, 0.1]
    meas = [synth.imu_samples(rng, n, w_scale=w)[2] for n, w in zip(lens, ws)]
    bias = rng.normal(0, 0.01, (len(lens), 6)).astype(np.float32)
    got = imu.preintegrate_batch(meas, bias, *synth.imu_calib_noise())
    for i in range(len(lens)):
        ref = O.imu_preintegrate(meas[i], bias[i], *synth.imu_calib_noise())
        for k, (a, b) in imu.FIELDS.items():
            scale = max(np.abs(ref[a:b]).max(), 1e-12)
            assert np.abs(got[i, a:b] - ref[a:b]).max() <= 2e-6 * scale, (i, k)
    # the record plugs into the inertial optimisers: same layout as synth.preintegrate
    acc, gyr, m = synth.imu_samples(rng, 60)
    rec = imu.preintegrate_batch([m], [[0.02, -0.03, 0.01, 0.002, -0.001, 0.0015]], *synth.imu_calib_noise())[0]
    ref = synth.preintegrate(acc, gyr, 1 / 200, [0.02, -0.03, 0.01, 0.002, -0.001, 0.0015])
    assert np.allclose(rec, ref, rtol=0, atol=5e-6 * np.abs(ref).max())


def test_large_batch_is_position_independent():
    rng = np.random.default_rng(4)
    uniq = [synth.imu_samples(rng, n)[2] for n in (40, 90, 7, 150)]
    bias = rng.normal(0, 0.01, (4, 6)).astype(np.float32)
    order = rng.integers(0, 4, 2048)
    got = imu.preintegrate_batch([uniq[j] for j in order], bias[order], *synth.imu_calib_noise())
    one = imu.preintegrate_batch(uniq, bias, *synth.imu_calib_noise())
    assert np.array_equal(got, one[order])

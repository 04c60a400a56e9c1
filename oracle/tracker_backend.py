"""The CPU oracle behind the `backend` interface of geoflowslam_b200/tracker.py and chain.py: the same closed-loop host logic
runs once on the CUDA library and once on this, and every decision / the trajectory are compared.  TEST INFRASTRUCTURE
(tests/, bench.py's ATE sub-line, scripts/config4_closed_loop.py)."""
import numpy as np

from geoflowslam_b200 import synth

from . import oracle as O


class OracleBackend:
    def __init__(self, gicp_threads=4):
        self._orb = O.OrbOracle(1000, 1.2, 8, 25, 7)
        self.gicp_threads = gicp_threads

    def orb(self, img):
        k, d, _ = self._orb.extract(img)
        return k, d

    def depth_to_cloud(self, depth, stride):
        c = synth.G1_CAM
        return O.depth_to_cloud(depth, stride, c["fx"], c["fy"], c["cx"], c["cy"])

    def gicp(self, target, source, T0):
        return O.gicp_align(target, source, T0=T0, threads=self.gicp_threads)

    def search_by_projection(self, mode, q, kps, u_right, desc, occupied, grid, nnratio, check_orientation):
        return O.search_by_projection(mode, q, kps, u_right, desc, occupied, grid, nnratio=nnratio, check_orientation=check_orientation)

    def search_with_gms(self, k1, d1, k2, d2, size):
        idx, _ = O.bf_match(d1, d2)
        m = np.stack([np.arange(len(idx), dtype=np.int32), idx], 1)
        mask, cnt = O.gms_filter(np.stack([k1["x"], k1["y"]], 1), size, np.stack([k2["x"], k2["y"]], 1), size, m)
        return m, mask, cnt

    def pose_optimization(self, prob):
        return O.pose_optimize(prob)

    def preintegrate(self, rows, bias6):
        return O.imu_preintegrate(rows, bias6, *synth.imu_calib_noise())

    def pose_inertial(self, prob):
        return O.pose_inertial_optimize(prob)

    def local_inertial_ba(self, prob):
        return O.ba_solve(prob)

    # chain.py's three callables
    def fb_klt(self, a, b, kps, priors):
        pa, pb = O.klt_build_pyramid(a, 3), O.klt_build_pyramid(b, 3)
        return O.fb_klt_tracking(pa, pb, a.shape[1], a.shape[0], 3, kps, priors)

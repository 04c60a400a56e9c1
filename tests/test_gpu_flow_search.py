"""ORBmatcher::SearchByProjectionWithOF / FilterOutliers (reference src/ORBmatcher.cc:2303-2497, 208-246) composed over the CUDA
tracker: the product's array version driven by KltTracker.fbKltTracking (the GPU kernels) against the oracle's statement-by-
statement restatement driven by the oracle's CPU tracker.  OpenCV's findFundamentalMat stays on the host in both, as in the
reference; because the CUDA tracker is bit-exact with the oracle's, the two runs must agree in every accepted id, position and
mask pixel."""
import importlib.util
import os

import numpy as np
import pytest

_spec = importlib.util.spec_from_file_location("_flow_search_cases", os.path.join(os.path.dirname(os.path.abspath(__file__)), "test_flow_search.py"))
_cases = importlib.util.module_from_spec(_spec)
_spec.loader.exec_module(_cases)
BOUNDS, K, _scenario = _cases.BOUNDS, _cases.K, _cases._scenario

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [0, 1, 5])
def test_with_of_on_the_cuda_tracker_equals_oracle(seed):
    from geoflowslam_b200 import KltTracker
    from geoflowslam_b200.matcher import search_by_projection_with_of
    from oracle import oracle as O
    fr, cur, last, state, X, R, t = _scenario(seed)
    trk = KltTracker(max_size=(640, 480), levels=3, max_points=1024, max_batch=1)   # Frame::Frame builds 3 levels (src/Frame.cc:373)

    def cuda_tracker(prev_img, cur_img, kps, priors, win, nlvl, ferr, max_dist):
        return trk.fbKltTracking(prev_img, cur_img, kps, priors, nwinsize=win, nbpyrlvl=nlvl, ferr=ferr, fmax_fbklt_dist=max_dist)

    m0 = np.zeros((480, 640), np.uint8)
    mask_g = m0.copy()
    n_g, ids_g, pts_g = search_by_projection_with_of(cuda_tracker, cur, last, state, X, R, t, K, BOUNDS, fr[0], fr[1], mask_g)
    n_o, tracked_o, mask_o = O.search_by_projection_with_of(cur, last, state, X, R, t, K, BOUNDS, fr[0], fr[1], m0.copy(),
                                                            tracker=O.fb_klt_tracking_images)
    assert n_g == n_o and n_g > 30
    assert [int(i) for i in ids_g] == [i for i, _ in tracked_o]
    assert np.array_equal(pts_g, np.array([p for _, p in tracked_o], np.float32).reshape(-1, 2))
    assert np.array_equal(mask_g, mask_o)


def test_filter_outliers_after_cuda_tracking():
    """FilterOutliers on keypoints the CUDA tracker moved: same flags as the oracle on the oracle-tracked keypoints."""
    from geoflowslam_b200 import KltTracker
    from geoflowslam_b200.matcher import filter_outliers
    from oracle import oracle as O
    fr, cur, last, state, X, R, t = _scenario(3)
    trk = KltTracker(max_size=(640, 480), levels=3, max_points=1024, max_batch=1)
    pr_g, st_g = trk.fbKltTracking(fr[0], fr[1], last, last)
    pr_o, st_o = O.fb_klt_tracking_images(fr[0], fr[1], last, last)
    assert np.array_equal(st_g, st_o) and np.array_equal(pr_g, pr_o)
    has = (state == 1) & st_g
    n_g, out_g = filter_outliers(pr_g, has, X, R, t, K, 1.0)
    n_o, out_o = O.filter_outliers(pr_o, has, X, R, t, K, 1.0)
    assert n_g == n_o and np.array_equal(out_g, out_o) and n_g > 20

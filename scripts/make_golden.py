#!/usr/bin/env python
"""Generates tests/golden/*.npz: outputs of the OpenCV calls the reference makes on this path, taken from the cv2 wheel
of this container (the reference leaves the OpenCV version unpinned, CMakeLists.txt:66; cv2.__version__ is stored in
every file).  The reference itself is C++ and cannot be built here, so these are the only outputs of
reference-side code that can be committed.  Inputs are small seeded images so the fixtures stay a few hundred KB.

  python scripts/make_golden.py          (re-creates the files; tests/test_golden.py consumes them)
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")


def image(seed, h, w):
    """smooth blobs + rectangles + noise: corners, plateaus and texture at a small size"""
    rng = np.random.default_rng(seed)
    img = cv2.GaussianBlur(rng.integers(0, 256, (h, w)).astype(np.float32), (0, 0), 3.0)
    img = (img - img.min()) / (img.max() - img.min()) * 200 + 20
    for _ in range(25):
        x, y = int(rng.integers(0, w - 8)), int(rng.integers(0, h - 8))
        img[y:y + int(rng.integers(4, 24)), x:x + int(rng.integers(4, 24))] = rng.integers(0, 256)
    img += rng.normal(0, 2.0, img.shape)
    return np.clip(img, 0, 255).astype(np.uint8)


def main():
    os.makedirs(OUT, exist_ok=True)
    ver = np.array(cv2.__version__)
    a = image(1, 120, 160)
    # second frame: the first one shifted by a sub-pixel affine warp (so optical flow has a known, small motion)
    M = np.array([[1.0, 0.01, 1.7], [-0.01, 1.0, -1.2]], np.float32)
    b = cv2.warpAffine(a, M, (160, 120), flags=cv2.INTER_LINEAR, borderMode=cv2.BORDER_REFLECT_101)
    # --- ORB extractor primitives (ORBextractor.cc:809,826,1189,1240)
    fast = {}
    for thr in (25, 7):
        kps = cv2.FastFeatureDetector_create(thr, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(a)
        fast["fast%d" % thr] = np.array([[k.pt[0], k.pt[1], k.response] for k in kps], np.float32).reshape(-1, 3)
    np.savez_compressed(os.path.join(OUT, "orb_primitives.npz"), cv2_version=ver, img=a,
                        resize_area=cv2.resize(a, (133, 100), interpolation=cv2.INTER_AREA),
                        blur7=cv2.GaussianBlur(a, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101), **fast)
    # --- BFMatcher (ORBmatcher.cc:755)
    rng = np.random.default_rng(2)
    dq = rng.integers(0, 256, (200, 32), dtype=np.uint8)
    dt = rng.integers(0, 256, (180, 32), dtype=np.uint8)
    dt[:60] = dq[:60] ^ (rng.integers(0, 256, (60, 32), dtype=np.uint8) & rng.integers(0, 256, (60, 32), dtype=np.uint8) & 3)
    dt[60:70] = dt[50:60]  # exact ties: the first index must win
    m = cv2.BFMatcher(cv2.NORM_HAMMING).match(dq, dt)
    np.savez_compressed(os.path.join(OUT, "bf_hamming.npz"), cv2_version=ver, dq=dq, dt=dt,
                        train_idx=np.array([x.trainIdx for x in m], np.int32), dist=np.array([x.distance for x in m], np.int32))
    # --- optical-flow front end (Frame.cc:366-373, ORBmatcher.cc:2224,2271)
    _, pyr = cv2.buildOpticalFlowPyramid(a, (21, 21), 2)
    pts = cv2.goodFeaturesToTrack(a, 120, 0.01, 5).reshape(-1, 2).astype(np.float32)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    fl = cv2.OPTFLOW_USE_INITIAL_FLOW + cv2.OPTFLOW_LK_GET_MIN_EIGENVALS
    nxt, st, err = cv2.calcOpticalFlowPyrLK(a, b, pts, pts.copy(), winSize=(21, 21), maxLevel=2, criteria=crit, flags=fl)
    np.savez_compressed(os.path.join(OUT, "optical_flow.npz"), cv2_version=ver, prev=a, cur=b, clahe=cv2.createCLAHE(3.0, (8, 8)).apply(a),
                        pyr_img1=pyr[2], pyr_img2=pyr[4], der0=pyr[1], der1=pyr[3], der2=pyr[5], pts=pts, next=nxt,
                        status=st.ravel().astype(np.uint8), min_eig=err.ravel().astype(np.float32))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)), "bytes")


if __name__ == "__main__":
    main()

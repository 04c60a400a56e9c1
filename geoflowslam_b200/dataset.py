"""Sequence plumbing either side of the hot path (BASELINE configs[0] / [4]): the on-disk layout the
reference's example drivers read, restated from their loaders --

    rgb/<t>.png            8UC3 (or 8UC1) colour image        Examples/RGB-D/rgbd_tum.cc:313-338,
    depth/<t>.png          16U depth, metres * DepthMapFactor  Examples/RGB-D-Inertial/rgbd_inertial.cc:326-349
    associate.txt          "<t> <rgb path> <depth path>" per frame; t in seconds (rgbd_tum) or milliseconds (rgbd_inertial)
    imu/imu.txt            "t_ms,ax,ay,az,gx,gy,gz", '#' comments  rgbd_inertial.cc:352-384
    imu/odom.txt           "t_ms,vx,vy,vz,wx,wy,wz"                rgbd_inertial.cc:386-412 (only columns 1-3 are kept)

-- and the frame loop's grouping of the inertial / odometry samples (rgbd_inertial.cc:80, 151-173): samples up to and
including the first frame's stamp are dropped, frame ni > 0 gets every sample with stamp <= its own, frames with fewer
than three IMU samples are skipped.  Host-side only (numpy + cv2 for the PNGs); nothing here touches the GPU."""
import os

import numpy as np


def write_sequence(root, stamps_s, images, depth_m, imu=None, odom=None, inertial=True, depth_factor=1000.0):
    """images: (n,h,w) or (n,h,w,3) u8; depth_m: (n,h,w) float metres (<= 0 / nan -> 0 = invalid); imu: (m,7) rows
    [t_s, ax, ay, az, gx, gy, gz]; odom: (k,7) rows [t_s, vx, vy, vz, wx, wy, wz].  inertial=True writes the
    rgbd_inertial flavour (stamps in milliseconds), False the rgbd_tum flavour (seconds).  Returns the association path."""
    import cv2
    os.makedirs(os.path.join(root, "rgb"), exist_ok=True)
    os.makedirs(os.path.join(root, "depth"), exist_ok=True)
    lines = []
    for t, img, d in zip(stamps_s, images, depth_m):
        ts = float(t) * 1e3 if inertial else float(t)
        name = "%.4f.png" % ts
        img = np.asarray(img, np.uint8)
        cv2.imwrite(os.path.join(root, "rgb", name), img if img.ndim == 3 else cv2.cvtColor(img, cv2.COLOR_GRAY2BGR))
        d = np.nan_to_num(np.asarray(d, np.float64), nan=0.0, posinf=0.0, neginf=0.0)
        d16 = np.clip(np.rint(np.maximum(d, 0.0) * depth_factor), 0, 65535).astype(np.uint16)
        cv2.imwrite(os.path.join(root, "depth", name), d16)
        lines.append("%.4f rgb/%s depth/%s" % (ts, name, name))
    assoc = os.path.join(root, "associate.txt")
    with open(assoc, "w") as f:
        f.write("\n".join(lines) + "\n")
    if imu is not None or odom is not None:
        os.makedirs(os.path.join(root, "imu"), exist_ok=True)
    for rows, fname, hdr in ((imu, "imu.txt", "#t_ms,ax,ay,az,gx,gy,gz"), (odom, "odom.txt", "#t_ms,vx,vy,vz,wx,wy,wz")):
        if rows is None:
            continue
        with open(os.path.join(root, "imu", fname), "w") as f:
            f.write(hdr + "\n")
            for r in np.asarray(rows, np.float64).reshape(-1, 7):
                f.write("%.4f,%s\n" % (r[0] * 1e3, ",".join("%.9g" % v for v in r[1:])))
    return assoc


def load_images(association_path, inertial=True):
    """LoadImages -> (rgb paths, depth paths, stamps in seconds)."""
    rgb, dep, ts = [], [], []
    with open(association_path) as f:
        for s in f.read().split("\n"):
            if not s:
                continue
            tok = s.split()
            ts.append(float(tok[0]) / 1e3 if inertial else float(tok[0]))
            rgb.append(tok[1]); dep.append(tok[2])
    return rgb, dep, np.array(ts, np.float64)


def _load_csv7(path):
    rows = []
    with open(path) as f:
        for s in f.read().split("\n"):
            if not s or s[0] == "#":
                continue
            rows.append([float(x) for x in s.split(",")[:7]])
    return np.array(rows, np.float64).reshape(-1, 7)


def load_imu(path):
    """LoadIMU -> (stamps s, acc (m,3) float32, gyro (m,3) float32); the file's columns are t_ms, acc, gyro."""
    r = _load_csv7(path)
    return r[:, 0] / 1e3, r[:, 1:4].astype(np.float32), r[:, 4:7].astype(np.float32)


def load_odom(path):
    """LoadOdom -> (stamps s, vPos (k,3) float32): only columns 1-3 are kept, as in the reference."""
    r = _load_csv7(path)
    return r[:, 0] / 1e3, r[:, 1:4].astype(np.float32)


def read_frame(root, rgb_path, depth_path, depth_factor=1000.0):
    """cv::imread(IMREAD_UNCHANGED) of both images + the conversions Tracking::GrabImageRGBD applies: colour -> gray
    (BGR order, Tracking.cc GrabImageRGBD) and depth * (1 / DepthMapFactor) as float32."""
    import cv2
    im = cv2.imread(os.path.join(root, rgb_path), cv2.IMREAD_UNCHANGED)
    d = cv2.imread(os.path.join(root, depth_path), cv2.IMREAD_UNCHANGED)
    if im is None or d is None:
        raise FileNotFoundError("failed to load image at: %s / %s" % (rgb_path, depth_path))
    gray = im if im.ndim == 2 else cv2.cvtColor(im, cv2.COLOR_BGR2GRAY if im.shape[2] == 3 else cv2.COLOR_BGRA2GRAY)
    return gray, d.astype(np.float32) * np.float32(1.0 / depth_factor)


def frame_measurements(frame_stamps, imu_stamps, odom_stamps=None, min_imu=3):
    """The frame loop of rgbd_inertial.cc: -> list of (frame index, imu index range, odom index range) for the frames that
    are handed to TrackRGBD; index ranges are half-open [a, b) into the loaded sample arrays.

    Before the loop the reference skips the samples stamped at or before frame 0 and then steps back by one
    (`first_imu--`, rgbd_inertial.cc:80-81; the same for the odometry, :82-85): the LAST sample at or before frame 0 is the
    first sample of frame 1's group -- Tracking::PreintegrateIMU interpolates from it at the start of the interval.
    (With no sample at or before frame 0 the reference's index would go to -1; clamped to 0 here.)  Frame 0 itself gets no
    samples (`if (ni > 0)`, :151,162) and is therefore dropped by the `vImuMeas.size() < 3` test (:171)."""
    first_imu = 0
    while first_imu < len(imu_stamps) and imu_stamps[first_imu] <= frame_stamps[0]:         # :80
        first_imu += 1
    first_imu = max(first_imu - 1, 0)                                                        # :81
    first_odom = 0
    if odom_stamps is not None:
        while first_odom < len(odom_stamps) and odom_stamps[first_odom] <= frame_stamps[0]:  # :83
            first_odom += 1
        first_odom = max(first_odom - 1, 0)                                                  # :84
    out = []
    for ni in range(len(frame_stamps)):
        a, oa = first_imu, first_odom
        if ni > 0:
            while first_imu < len(imu_stamps) and imu_stamps[first_imu] <= frame_stamps[ni]:  # :151-159
                first_imu += 1
            if odom_stamps is not None:
                while first_odom < len(odom_stamps) and odom_stamps[first_odom] <= frame_stamps[ni]:  # :162-168
                    first_odom += 1
        if first_imu - a < min_imu:                                                           # :171
            continue
        out.append((ni, (a, first_imu), (oa, first_odom)))
    return out

"""A minimal closed-loop RGB-D-inertial tracker over the library's entry points (BASELINE.json configs[4], "ATE vs ref").

It is the host logic of SURVEY.md 3.5 -- the part of Tracking / LocalMapping that decides WHICH hot-path call runs WHEN and
on WHAT -- restated on plain arrays as far as needed to drive every stage of the path in the reference's order with the
reference's gates, NOT the reference's state machine (relocalisation, map reset, culling, loop closing, IMU initialisation
and the threads are out of scope).  Per frame (System::TrackRGBD -> Tracking::Track, src/Tracking.cc:2042):

  Frame::Frame                 ORB extraction, depth per keypoint (ComputeStereoFromRGBD, src/Frame.cc:1314-1332),
                               depth -> cloud (ConvertDepthToPointCloud with the yaml's Downsample 3, :590-623)
  frame 0                      StereoInitialization: one map point per keypoint with depth, first keyframe (:2697-2823)
  Tracking::PreintegrateIMU    frame-to-frame and keyframe-to-frame preintegration (:1724-1830)
  TrackWithMotionModelICP      PredictStateIMU (:1876-1950) -> PredictStateICP: GICP of the two clouds, accepted iff
                               `converged && num_inliers > 200` (:3364-3413) -> SearchByProjection(cur, last, th = 15), retried
                               with 2 th below 25 matches (:3622-3654), SearchWithGMS against the last frame's map points if
                               that still leaves fewer than 15 (TrackReferenceKeyFrame's matcher, :3126) ->
                               Optimizer::PoseOptimization, which only classifies outliers (src/Optimizer.cc:1090-1097)
  TrackLocalMap                SearchLocalPoints: isInFrustum + SearchByProjection(F, local points, th) (:4294-4359,
                               src/Frame.cc:876-931) -> PoseInertialOptimizationLastKeyFrame after a map update, else
                               LastFrame with the marginalised prior (:3729-3799)
  every kf_every-th frame      CreateNewKeyFrame: new map points for the closest keypoints with depth (:4168-4292), then
                               LocalMapping's Optimizer::LocalInertialBA over the last <= 10 keyframes (src/LocalMapping.cc:223,
                               src/Optimizer.cc:3056-3702): window, fixed predecessor, fixed covisible observers, EdgeStereo /
                               EdgeMono observations, inertial edges from the keyframe-to-keyframe preintegrations; outlier
                               observations erased, states written back (as float), last frame re-predicted from its keyframe
                               (Tracking::UpdateFrameIMU, :4900-4960).

Everything numerical goes through a `backend` (CudaBackend below = this library; the tests' OracleBackend = the CPU oracle),
so the SAME host logic runs on both and every integer decision (match lists, outlier flags, accept / retry gates, keyframe
contents) can be compared one to one.  Simplifications against the reference, all on the host side: the IMU is taken as
initialised from frame 0 (ground-truth first pose and velocity, zero biases); keyframes are inserted every kf_every-th
frame instead of by NeedNewKeyFrame's covisibility heuristics; local BA runs synchronously inside the frame that inserts
the keyframe; a map point's descriptor / viewing direction / scale range come from the keyframe that created it."""
import numpy as np

from ._lib import KP_DTYPE

TH_HIGH = 100
NLEVELS, SCALE = 8, 1.2
SF = (np.float32(SCALE) ** np.arange(NLEVELS)).astype(np.float32)
for _i in range(1, NLEVELS):
    SF[_i] = np.float32(SF[_i - 1] * np.float32(SCALE))
INV_SIGMA2 = (np.float32(1.0) / (SF * SF)).astype(np.float32)
PROJ_QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("ur", "<f4"), ("angle", "<f4"),
                             ("min_level", "<i4"), ("max_level", "<i4"), ("blocks", "<i4"), ("desc", "u1", (32,))])


def f32_64(x):
    """a float32 value widened to double: how the reference stores poses / points and loads them into g2o"""
    return np.asarray(x, np.float32).astype(np.float64)


def _polar32(R):
    U, _, Vt = np.linalg.svd(np.asarray(R, np.float32))
    return (U @ Vt).astype(np.float32)


class Frame:
    """the arrays of ORB_SLAM3::Frame the path reads (include/Frame.h)"""

    def __init__(self, idx, img, depth, backend, cam, cloud_stride):
        self.id = idx
        self.img = img
        self.keys, self.desc = backend.orb(img)                                   # Frame::ExtractORB
        n = len(self.keys)
        self.N = n
        u = self.keys["x"].astype(np.int32); v = self.keys["y"].astype(np.int32)  # imDepth.at<float>(v, u): truncation
        d = depth[np.clip(v, 0, depth.shape[0] - 1), np.clip(u, 0, depth.shape[1] - 1)].astype(np.float32)
        self.depth = np.where(d > 0, d, np.float32(-1)).astype(np.float32)        # mvDepth
        with np.errstate(divide="ignore"):
            self.u_right = np.where(d > 0, self.keys["x"] - np.float32(cam["bf"]) / d, np.float32(-1)).astype(np.float32)
        self.cloud = backend.depth_to_cloud(depth, cloud_stride)                  # source_points
        self.mp = np.full(n, -1, np.int64)        # mvpMapPoints (map point id or -1)
        self.outlier = np.zeros(n, bool)          # mvbOutlier
        # body state (mImuBias, pose, velocity); float32 like Frame::SetImuPoseVelocity stores them
        self.Rwb = self.twb = self.vel = None
        self.bg = np.zeros(3); self.ba = np.zeros(3)
        self.prior_H = None                       # mpcpi of the previous PoseInertialOptimization (15 x 15)
        self.pre_frame = None                     # mpImuPreintegratedFrame
        self.pre_kf = None                        # mpImuPreintegrated (from the last keyframe)


class Tracker:
    def __init__(self, seq, backend, kf_every=5, cloud_stride=3, window=10, use_icp=True, log=None):
        self.seq, self.be = seq, backend
        self.cam = seq["cam"]
        self.Rbc = np.asarray(seq["Rbc"], np.float64); self.tbc = np.asarray(seq["tbc"], np.float64)
        self.Rcb = self.Rbc.T; self.tcb = -self.Rcb @ self.tbc
        self.g = np.asarray(seq["gravity"], np.float64)
        self.kf_every, self.cloud_stride, self.window, self.use_icp = kf_every, cloud_stride, window, use_icp
        self.w, self.h = seq["frames"].shape[2], seq["frames"].shape[1]
        self.grid = (0.0, 0.0, 64.0 / self.w, 48.0 / self.h)                      # mnMinX, mnMinY, mfGridElement{Width,Height}Inv
        from . import synth
        self.cal = synth.imu_calib_noise()
        # the map: map points as parallel lists, keyframes as dicts
        self.mp_X = []; self.mp_desc = []; self.mp_normal = []; self.mp_maxd = []; self.mp_mind = []; self.mp_obs = []; self.mp_bad = []
        self.kfs = []
        self.map_updated = False
        self.decisions = []                       # integer decisions, for one-to-one comparison between backends
        self.traj = []
        self.log = log

    # ---- geometry helpers (float32 like Frame::SetPose / GetPose)
    def Tcw_of(self, Rwb, twb):
        R32 = _polar32(Rwb); p32 = np.asarray(twb, np.float32)
        Rcw = (self.Rcb.astype(np.float32) @ R32.T).astype(np.float32)
        tcw = (self.Rcb.astype(np.float32) @ (-(R32.T @ p32)) + self.tcb.astype(np.float32)).astype(np.float32)
        return Rcw, tcw

    def body_of(self, Rcw, tcw):
        """ImuPose from the camera pose (Frame::UpdatePoseMatrices + GetImuPose)"""
        Rwc = np.asarray(Rcw, np.float64).T
        Rwb = Rwc @ self.Rcb
        twb = Rwc @ (self.tcb - np.asarray(tcw, np.float64))
        return Rwb, twb

    def project(self, Rcw, tcw, Xw):
        Xc = (np.asarray(Xw, np.float32) @ np.asarray(Rcw, np.float32).T + np.asarray(tcw, np.float32)).astype(np.float32)
        with np.errstate(divide="ignore", invalid="ignore"):
            invz = (np.float32(1.0) / Xc[:, 2]).astype(np.float32)
            u = (np.float32(self.cam["fx"]) * Xc[:, 0] * invz + np.float32(self.cam["cx"])).astype(np.float32)
            v = (np.float32(self.cam["fy"]) * Xc[:, 1] * invz + np.float32(self.cam["cy"])).astype(np.float32)
        return Xc, invz, u, v

    # ---- map points
    def new_map_point(self, Xw, desc, octave, Ow, kf_idx, key_idx):
        """MapPoint::MapPoint(Pos, pRefKF) + UpdateNormalAndDepth (src/MapPoint.cc): normal = viewing direction of the creating
        keyframe, mfMaxDistance = dist * scale[octave], mfMinDistance = mfMaxDistance / scale[nlevels - 1]"""
        Xw = np.asarray(Xw, np.float32)
        PO = Xw - np.asarray(Ow, np.float32)
        dist = np.float32(np.linalg.norm(PO))
        self.mp_X.append(Xw); self.mp_desc.append(np.array(desc, np.uint8)); self.mp_normal.append((PO / dist).astype(np.float32))
        maxd = np.float32(dist * SF[octave])
        self.mp_maxd.append(maxd); self.mp_mind.append(np.float32(maxd / SF[NLEVELS - 1]))
        self.mp_obs.append({kf_idx: key_idx}); self.mp_bad.append(False)
        return len(self.mp_X) - 1

    # ---- frame 0
    def initialize(self, F):
        """StereoInitialization (:2697-2823): needs more than 500 keypoints; a map point for every keypoint with depth"""
        assert F.N > 500, "StereoInitialization needs more than 500 keypoints"
        s = self.seq
        F.Rwb = f32_64(_polar32(s["Rwb"][0])); F.twb = f32_64(s["twb"][0]); F.vel = f32_64(s["vel"][0])
        Rcw, tcw = self.Tcw_of(F.Rwb, F.twb)
        Rwc = Rcw.T.astype(np.float32); Ow = (-(Rwc @ tcw)).astype(np.float32)
        kf = self.make_keyframe(F, None)
        for i in range(F.N):
            z = F.depth[i]
            if z > 0:
                x3c = np.array([(F.keys["x"][i] - np.float32(self.cam["cx"])) * z / np.float32(self.cam["fx"]),
                                (F.keys["y"][i] - np.float32(self.cam["cy"])) * z / np.float32(self.cam["fy"]), z], np.float32)
                Xw = (Rwc @ x3c + Ow).astype(np.float32)                           # Frame::UnprojectStereo
                m = self.new_map_point(Xw, F.desc[i], int(F.keys["octave"][i]), Ow, kf["idx"], i)
                F.mp[i] = m; kf["mp"][i] = m
        self.decisions.append(("init", F.N, int((F.mp >= 0).sum())))

    def make_keyframe(self, F, pre_from_prev):
        kf = dict(idx=len(self.kfs), frame_id=F.id, keys=F.keys, desc=F.desc, depth=F.depth, u_right=F.u_right, mp=F.mp.copy(),
                  Rwb=F.Rwb.copy(), twb=F.twb.copy(), vel=F.vel.copy(), bg=F.bg.copy(), ba=F.ba.copy(), pre=pre_from_prev,
                  cloud=F.cloud, matches_inliers=0)
        self.kfs.append(kf)
        return kf

    # ---- Tracking::PredictStateIMU (:1876-1950)
    def predict_state_imu(self, F, last):
        if self.map_updated:   # from the last keyframe with the keyframe-to-frame preintegration
            kf = self.kfs[-1]
            R1, p1, v1, rec = kf["Rwb"], kf["twb"], kf["vel"], F.pre_kf
            F.bg, F.ba = kf["bg"].copy(), kf["ba"].copy()
        else:                  # from the last frame with the frame-to-frame preintegration
            R1, p1, v1, rec = last.Rwb, last.twb, last.vel, F.pre_frame
            F.bg, F.ba = last.bg.copy(), last.ba.copy()
        dR = rec[0:9].reshape(3, 3).astype(np.float64); dV = rec[9:12].astype(np.float64); dP = rec[12:15].astype(np.float64)
        t = float(rec[285])
        R2 = R1 @ dR
        p2 = p1 + v1 * t + 0.5 * t * t * self.g + R1 @ dP
        v2 = v1 + t * self.g + R1 @ dV
        F.Rwb = f32_64(_polar32(R2)); F.twb = f32_64(p2); F.vel = f32_64(v2)

    # ---- TrackWithMotionModelICP's matcher calls
    def search_last_frame(self, F, last, Rcw, tcw, th):
        """ORBmatcher::SearchByProjection(CurrentFrame, LastFrame, th, false) (src/ORBmatcher.cc:1853-2063) flattened"""
        idx = np.nonzero((last.mp >= 0) & ~last.outlier)[0]
        q = np.zeros(len(idx), PROJ_QUERY_DTYPE)
        if len(idx) == 0:
            return np.full(F.N, -1, np.int32), 0, idx
        Rl, tl = self.Tcw_of(last.Rwb, last.twb)
        twc = (-(Rcw.T @ tcw)).astype(np.float32)
        tlc = (Rl @ twc + tl).astype(np.float32)
        b = np.float32(self.cam["bf"]) / np.float32(self.cam["fx"])
        fwd, bwd = tlc[2] > b, -tlc[2] > b
        Xw = np.stack([self.mp_X[m] for m in last.mp[idx]])
        Xc, invz, u, v = self.project(Rcw, tcw, Xw)
        ok = (invz >= 0) & (u >= 0) & (u <= self.w) & (v >= 0) & (v <= self.h)
        octv = last.keys["octave"][idx].astype(np.int32)
        q["u"], q["v"] = u, v
        q["radius"] = np.where(ok, np.float32(th) * SF[octv], np.float32(-1))
        q["ur"] = (u - np.float32(self.cam["bf"]) * invz).astype(np.float32)
        q["angle"] = last.keys["angle"][idx]
        if fwd:
            q["min_level"], q["max_level"] = octv, -1
        elif bwd:
            q["min_level"], q["max_level"] = 0, octv
        else:
            q["min_level"], q["max_level"] = octv - 1, octv + 1
        q["blocks"] = [1 if len(self.mp_obs[m]) > 0 else 0 for m in last.mp[idx]]
        q["desc"] = np.stack([self.mp_desc[m] for m in last.mp[idx]])
        assign, nm = self.be.search_by_projection(0, q, F.keys, F.u_right, F.desc, None, self.grid, nnratio=0.9, check_orientation=True)
        return assign, nm, idx

    def search_local_points(self, F, Rcw, tcw, th):
        """Tracking::SearchLocalPoints (:4294-4359): Frame::isInFrustum(pMP, 0.5) for the local map points that are not matched
        yet, then ORBmatcher(0.8).SearchByProjection(F, vpMapPoints, th)"""
        have = set(int(m) for m in F.mp[F.mp >= 0])
        local = sorted(set(int(m) for kf in self.kfs[-self.window:] for m in kf["mp"][kf["mp"] >= 0]) - have)   # UpdateLocalPoints
        local = [m for m in local if not self.mp_bad[m]]
        if not local:
            return 0
        Xw = np.stack([self.mp_X[m] for m in local])
        Xc, invz, u, v = self.project(Rcw, tcw, Xw)
        Ow = (-(Rcw.T @ tcw)).astype(np.float32)
        PO = (Xw - Ow).astype(np.float32)
        dist = np.linalg.norm(PO, axis=1).astype(np.float32)
        maxd = np.array([self.mp_maxd[m] for m in local], np.float32); mind = np.array([self.mp_mind[m] for m in local], np.float32)
        nrm = np.stack([self.mp_normal[m] for m in local])
        with np.errstate(divide="ignore", invalid="ignore"):
            view_cos = (np.einsum("ij,ij->i", PO, nrm) / dist).astype(np.float32)
            ratio = maxd / dist
            lvl = np.ceil(np.log(ratio) / np.float32(np.log(np.float32(SCALE)))).astype(np.int32)                # MapPoint::PredictScale
        lvl = np.clip(lvl, 0, NLEVELS - 1)
        ok = (Xc[:, 2] >= 0) & (u >= 0) & (u <= self.w) & (v >= 0) & (v <= self.h) & (dist >= np.float32(0.8) * mind) & \
            (dist <= np.float32(1.2) * maxd) & (view_cos >= 0.5)
        r = np.where(view_cos > 0.998, np.float32(2.5), np.float32(4.0)) * np.float32(th)                        # RadiusByViewingCos
        q = np.zeros(len(local), PROJ_QUERY_DTYPE)
        q["u"], q["v"] = u, v
        q["radius"] = np.where(ok, r * SF[lvl], np.float32(-1))
        q["ur"] = (u - np.float32(self.cam["bf"]) * invz).astype(np.float32)
        q["min_level"], q["max_level"] = lvl - 1, lvl
        q["blocks"] = 1
        q["desc"] = np.stack([self.mp_desc[m] for m in local])
        occupied = (F.mp >= 0).astype(np.uint8)
        assign, nm = self.be.search_by_projection(1, q, F.keys, F.u_right, F.desc, occupied, self.grid, nnratio=0.8, check_orientation=True)
        for i in np.nonzero(assign >= 0)[0]:
            F.mp[i] = local[assign[i]]
        return nm

    def observations(self, F, sel):
        """(Xw, uvr, inv_sigma2) of the matched keypoints `sel`"""
        if len(sel) == 0:
            return np.zeros((0, 3)), np.zeros((0, 3), np.float32), np.zeros(0, np.float32)
        Xw = np.ascontiguousarray(np.stack([self.mp_X[m] for m in F.mp[sel]]).astype(np.float64))
        uvr = np.ascontiguousarray(np.stack([F.keys["x"][sel], F.keys["y"][sel], F.u_right[sel]], 1), np.float32)
        return Xw, uvr, INV_SIGMA2[F.keys["octave"][sel].astype(np.int32)]

    def pose_optimization(self, F, Rcw, tcw):
        """Optimizer::PoseOptimization(&mCurrentFrame): classifies outliers, the pose is NOT written back (Optimizer.cc:1090-1097)"""
        sel = np.nonzero(F.mp >= 0)[0]
        if len(sel) < 3:
            return 0
        Xw, uvr, is2 = self.observations(F, sel)
        from . import synth
        prob = dict(n_obs=len(sel), q_wxyz=synth._quat_from_R(np.asarray(Rcw, np.float64)).astype(np.float32), t=np.asarray(tcw, np.float32),
                    fx=np.float32(self.cam["fx"]), fy=np.float32(self.cam["fy"]), cx=np.float32(self.cam["cx"]), cy=np.float32(self.cam["cy"]),
                    bf=np.float32(self.cam["bf"]), Xw=Xw, uvr=uvr, inv_sigma2=is2)
        r = self.be.pose_optimization(prob)
        F.outlier[:] = False
        F.outlier[sel] = np.asarray(r["outlier"], bool)
        return int(r["n_inliers"])

    def pose_inertial(self, F, last):
        """PoseInertialOptimizationLastKeyFrame (after a map update) / LastFrame (src/Optimizer.cc:5899, 6762)"""
        sel = np.nonzero(F.mp >= 0)[0]
        Xw, uvr, is2 = self.observations(F, sel)
        Rcw, tcw = self.Tcw_of(F.Rwb, F.twb)
        if self.map_updated or last.prior_H is None:
            mode, rec = 0, F.pre_kf
            kf = self.kfs[-1]
            pR, pp, pv, pbg, pba = kf["Rwb"], kf["twb"], kf["vel"], kf["bg"], kf["ba"]
            H = np.zeros(225)
        else:
            mode, rec = 1, F.pre_frame
            pR, pp, pv, pbg, pba = last.Rwb, last.twb, last.vel, last.bg, last.ba
            H = np.asarray(last.prior_H, np.float64).ravel()
        C = rec[60:285].reshape(15, 15)
        c32 = lambda x: np.float32(self.cam[x])
        prob = dict(mode=mode, n_obs=len(sel), n_rounds=4, rec_init=0, fx=c32("fx"), fy=c32("fy"), cx=c32("cx"), cy=c32("cy"), bf=c32("bf"),
                    Rcb=f32_64(self.Rcb).ravel(), tcb=f32_64(self.tcb), tbc=f32_64(self.tbc),
                    Rwb=f32_64(F.Rwb).ravel(), twb=f32_64(F.twb), Rcw=f32_64(Rcw).ravel(), tcw=f32_64(tcw), vel=f32_64(F.vel),
                    bg=f32_64(F.bg), ba=f32_64(F.ba),
                    p_Rwb=f32_64(pR).ravel(), p_twb=f32_64(pp), p_vel=f32_64(pv), p_bg=f32_64(pbg), p_ba=f32_64(pba),
                    pre=rec, rw_Cg=np.ascontiguousarray(C[9:12, 9:12], np.float32).ravel(), rw_Ca=np.ascontiguousarray(C[12:15, 12:15], np.float32).ravel(),
                    c_Rwb=f32_64(pR).ravel(), c_twb=f32_64(pp), c_vwb=f32_64(pv), c_bg=f32_64(pbg), c_ba=f32_64(pba), c_H=H,
                    Xw=Xw, uvr=uvr, inv_sigma2=is2, close=np.ones(len(sel), np.uint8))
        r = self.be.pose_inertial(prob)
        F.outlier[:] = False
        F.outlier[sel] = np.asarray(r["outlier"], bool)
        F.Rwb, F.twb, F.vel = f32_64(_polar32(r["Rwb"])), f32_64(r["twb"]), f32_64(r["vel"])   # SetImuPoseVelocity: float
        F.bg, F.ba = f32_64(r["bg"]), f32_64(r["ba"])
        F.prior_H = r["H"]
        return mode, int(r["n_inliers"])

    # ---- Tracking::CreateNewKeyFrame (:4168-4292) + LocalMapping's LocalInertialBA
    def create_keyframe(self, F, n_inliers):
        kf = self.make_keyframe(F, F.pre_kf)
        kf["matches_inliers"] = n_inliers
        for i in np.nonzero(F.mp >= 0)[0]:
            if not F.outlier[i]:
                self.mp_obs[F.mp[i]][kf["idx"]] = int(i)                          # KeyFrame::AddMapPoint / MapPoint::AddObservation
            else:
                kf["mp"][i] = -1
        Rcw, tcw = self.Tcw_of(F.Rwb, F.twb)
        Rwc = Rcw.T.astype(np.float32); Ow = (-(Rwc @ tcw)).astype(np.float32)
        order = sorted((float(F.depth[i]), int(i)) for i in range(F.N) if F.depth[i] > 0)
        th_depth = np.float32(self.cam["bf"]) * np.float32(40.0) / np.float32(self.cam["fx"])   # mThDepth = mbf * ThDepth / fx
        n_points = created = 0
        for z, i in order:
            if kf["mp"][i] < 0 or F.outlier[i]:
                zz = np.float32(z)
                x3c = np.array([(F.keys["x"][i] - np.float32(self.cam["cx"])) * zz / np.float32(self.cam["fx"]),
                                (F.keys["y"][i] - np.float32(self.cam["cy"])) * zz / np.float32(self.cam["fy"]), zz], np.float32)
                m = self.new_map_point((Rwc @ x3c + Ow).astype(np.float32), F.desc[i], int(F.keys["octave"][i]), Ow, kf["idx"], i)
                F.mp[i] = m; kf["mp"][i] = m; F.outlier[i] = False
                created += 1
            n_points += 1
            if z > th_depth and n_points > 100:
                break
        self.decisions.append(("kf", kf["idx"], int((kf["mp"] >= 0).sum()), created))
        return kf

    def local_inertial_ba(self):
        """Optimizer::LocalInertialBA(pKF, ..., bLarge) flattened into the GfsBaProblem arrays (src/Optimizer.cc:3062-3577)"""
        kfs = self.kfs
        if len(kfs) < 3:                                                            # LocalMapping.cc:184: KeyFramesInMap() > 2
            return None
        cur = kfs[-1]
        b_large = cur["matches_inliers"] > 100                                      # LocalMapping.cc:217-221
        Nd = min(len(kfs) - 2, 20 if b_large else 10)                               # :3062-3068, maxOpt / opt_it
        opt = [kfs[-1 - j] for j in range(min(Nd, len(kfs)))]                       # pKF, then its mPrevKF chain (:3078-3086)
        oldest = opt[-1]
        if oldest["idx"] > 0:
            fixed = [kfs[oldest["idx"] - 1]]                                        # the window's predecessor is fixed (:3106-3113)
        else:
            fixed = [opt.pop()]                                                     # no predecessor: the oldest one is fixed (:3114-3120)
        opt_ids = {k["idx"]: i for i, k in enumerate(opt)}
        pts = sorted(set(int(m) for k in opt for m in k["mp"][k["mp"] >= 0] if not self.mp_bad[m]))   # lLocalMapPoints (:3090-3103)
        fixed_ids = {fixed[0]["idx"]: len(opt)}
        for m in pts:                                                               # fixed covisible observers (:3123-3146)
            for kidx in self.mp_obs[m]:
                if kidx not in opt_ids and kidx not in fixed_ids and len(fixed_ids) < 200:
                    fixed_ids[kidx] = len(opt) + len(fixed); fixed.append(kfs[kidx])
        allk = opt + fixed
        slot = dict(opt_ids); slot.update(fixed_ids)
        nk = len(allk)
        A = lambda f, w_: np.ascontiguousarray(np.stack([f(k) for k in allk]).reshape(nk, w_), np.float64)
        camT = [self.Tcw_of(k["Rwb"], k["twb"]) for k in allk]
        pt_slot = {m: j for j, m in enumerate(pts)}
        obs_kf, obs_pt, obs_uvr, obs_is2, obs_ref = [], [], [], [], []
        for m in pts:                                                               # edges in map-point order, observations in keyframe order
            for kidx in sorted(self.mp_obs[m]):
                if kidx not in slot:
                    continue
                k = kfs[kidx]; i = self.mp_obs[m][kidx]
                obs_kf.append(slot[kidx]); obs_pt.append(pt_slot[m])
                obs_uvr.append([k["keys"]["x"][i], k["keys"]["y"][i], k["u_right"][i]])
                obs_is2.append(INV_SIGMA2[int(k["keys"]["octave"][i])]); obs_ref.append((m, kidx))
        n_in = len(opt) if fixed[0]["idx"] == opt[-1]["idx"] - 1 else len(opt) - 1
        in_kf1, in_kf2, pre, down = [], [], [], []
        for i, k in enumerate(opt):                                                 # EdgeInertial between k and its predecessor (:3328-3401)
            if k["idx"] == 0 or (k["idx"] - 1) not in slot or k["pre"] is None:
                continue
            in_kf1.append(slot[k["idx"] - 1]); in_kf2.append(i); pre.append(k["pre"]); down.append(1 if i == len(opt) - 1 else 0)
        c32 = lambda x: np.float32(self.cam[x])
        prob = dict(n_opt_kf=len(opt), n_fixed_kf=len(fixed), n_points=len(pts), n_obs=len(obs_kf), n_inertial=len(in_kf1),
                    iterations=4 if b_large else 8, b_large=int(b_large), lambda_init=1e-2 if b_large else 1.0,
                    Rcb=f32_64(self.Rcb).ravel(), tcb=f32_64(self.tcb), Rbc=f32_64(self.Rbc).ravel(), tbc=f32_64(self.tbc),
                    fx=c32("fx"), fy=c32("fy"), cx=c32("cx"), cy=c32("cy"), bf=float(c32("bf")),
                    kf_Rwb=A(lambda k: f32_64(k["Rwb"]).ravel(), 9), kf_twb=A(lambda k: f32_64(k["twb"]), 3),
                    kf_Rcw=np.ascontiguousarray(np.stack([f32_64(c[0]).ravel() for c in camT])), kf_tcw=np.ascontiguousarray(np.stack([f32_64(c[1]) for c in camT])),
                    kf_vel=A(lambda k: f32_64(k["vel"]), 3), kf_bg=A(lambda k: f32_64(k["bg"]), 3), kf_ba=A(lambda k: f32_64(k["ba"]), 3),
                    kf_has_imu=np.ones(nk, np.uint8),
                    pt_xyz=np.ascontiguousarray(np.stack([self.mp_X[m] for m in pts]).astype(np.float64)), pt_close=np.ones(len(pts), np.uint8),
                    obs_kf=np.array(obs_kf, np.int32), obs_pt=np.array(obs_pt, np.int32),
                    obs_uvr=np.array(obs_uvr, np.float64).reshape(-1, 3), obs_inv_sigma2=np.array(obs_is2, np.float32),
                    in_kf1=np.array(in_kf1, np.int32), in_kf2=np.array(in_kf2, np.int32),
                    in_pre=np.ascontiguousarray(np.stack(pre).astype(np.float32)) if pre else np.zeros((0, 292), np.float32),
                    in_downweight=np.array(down, np.uint8),
                    n_icp=0, icp_kf1=np.zeros(0, np.int32), icp_kf2=np.zeros(0, np.int32), icp_Rt=np.zeros((0, 12)))
        r = self.be.local_inertial_ba(prob)
        n_out = int(np.asarray(r["obs_outlier"]).sum())
        self.decisions.append(("ba", len(opt), len(fixed), len(pts), len(obs_kf), len(in_kf1), int(r["lm_trials"]), n_out, int(r["failed"])))
        if r["failed"]:
            return r
        for j, (m, kidx) in enumerate(obs_ref):                                     # erase outlier observations (:3638-3650)
            if r["obs_outlier"][j]:
                i = self.mp_obs[m].pop(kidx)
                kfs[kidx]["mp"][i] = -1
                if len(self.mp_obs[m]) == 0:
                    self.mp_bad[m] = True
        for i, k in enumerate(opt):                                                 # write back as float (:3654-3690)
            k["Rwb"] = f32_64(_polar32(np.asarray(r["kf_Rwb"][i]).reshape(3, 3))); k["twb"] = f32_64(r["kf_twb"][i])
            k["vel"] = f32_64(r["kf_vel"][i]); k["bg"] = f32_64(r["kf_bg"][i]); k["ba"] = f32_64(r["kf_ba"][i])
        for j, m in enumerate(pts):
            self.mp_X[m] = np.asarray(r["pt_xyz"][j], np.float32)
        self.map_updated = True
        return r

    # ---- the frame loop
    def run(self, n_frames=None):
        s = self.seq
        n = n_frames or len(s["frames"])
        last = None
        rows_since_kf = []
        for k in range(n):
            F = Frame(k, s["frames"][k], s["depth"][k], self.be, self.cam, self.cloud_stride)
            if k == 0:
                self.initialize(F)
                self.traj.append((F.Rwb.copy(), F.twb.copy()))
                last = F
                continue
            kf = self.kfs[-1]
            rows = s["imu"][k - 1]
            rows_since_kf.append(rows)
            bias_last = np.concatenate([last.ba, last.bg])
            F.pre_frame = self.be.preintegrate(rows, bias_last)                                       # mpImuPreintegratedFrame
            F.pre_kf = self.be.preintegrate(np.concatenate(rows_since_kf), np.concatenate([kf["ba"], kf["bg"]]))   # mpImuPreintegratedFromLastKF
            if self.map_updated:
                # Tracking::UpdateFrameIMU after local BA: the last frame's state is re-predicted from its (updated) keyframe
                self.update_last_frame(last, rows_since_kf[:-1])
            self.predict_state_imu(F, last)
            Rcw, tcw = self.Tcw_of(F.Rwb, F.twb)
            icp_ok = False
            if self.use_icp:
                from .gicp import predict_state_icp
                Tl = np.eye(4, dtype=np.float32); Tl[:3, :3], Tl[:3, 3] = self.Tcw_of(last.Rwb, last.twb)
                Tc = np.eye(4, dtype=np.float32); Tc[:3, :3], Tc[:3, 3] = Rcw, tcw
                icp = predict_state_icp(self.be.gicp, Tl, Tc, last.cloud, F.cloud)
                icp_ok = bool(icp["ok"])
                res = icp["result"]
                self.decisions.append(("icp", k, int(icp_ok), -1 if res is None else int(res["iterations"]), -1 if res is None else int(res["num_inliers"])))
                if icp_ok:
                    Rcw, tcw = icp["Tcw"][:3, :3].copy(), icp["Tcw"][:3, 3].copy()
                    Rwb, twb = self.body_of(Rcw, tcw)
                    F.Rwb, F.twb = f32_64(_polar32(Rwb)), f32_64(twb)
            th = 15
            assign, nm, idx = self.search_last_frame(F, last, Rcw, tcw, th)
            if nm < 25:                                                                               # :3640-3654
                assign, nm, idx = self.search_last_frame(F, last, Rcw, tcw, 2 * th)
            how = "proj"
            if nm >= 15:
                sel = np.nonzero(assign >= 0)[0]
                F.mp[sel] = last.mp[idx[assign[sel]]]
            else:
                # fallback on the matcher of TrackReferenceKeyFrame: SearchWithGMS over the last frame's map-point keypoints
                how = "gms"
                have = np.nonzero((last.mp >= 0) & ~last.outlier)[0]
                m, mask, cnt = self.be.search_with_gms(last.keys[have], last.desc[have], F.keys, F.desc, (self.w, self.h))
                for (qi, ti), ok in zip(m, mask):
                    if ok and F.mp[ti] < 0:
                        F.mp[ti] = last.mp[have[qi]]
                nm = int((F.mp >= 0).sum())
            n_pose = self.pose_optimization(F, Rcw, tcw)
            drop = np.nonzero((F.mp >= 0) & F.outlier)[0]                                             # discard outliers (:3688-3711)
            F.mp[drop] = -1; F.outlier[drop] = False
            n_local = self.search_local_points(F, Rcw, tcw, th=2)                                     # IMU initialised: th = 2
            mode, n_inl = self.pose_inertial(F, last)
            self.decisions.append(("track", k, how, int(nm), int(n_pose), int(len(drop)), int(n_local), int(mode), int(n_inl),
                                   int(F.outlier.sum())))
            self.map_updated = False
            self.traj.append((F.Rwb.copy(), F.twb.copy()))
            if k % self.kf_every == 0:
                self.create_keyframe(F, n_inl)
                rows_since_kf = []
                self.local_inertial_ba()
            last = F
        return dict(Rwb=np.array([t[0] for t in self.traj]), twb=np.array([t[1] for t in self.traj]), decisions=self.decisions,
                    n_keyframes=len(self.kfs), n_map_points=len(self.mp_X))

    def update_last_frame(self, last, rows_to_last):
        """Tracking::UpdateFrameIMU (:4900-4960) for the last frame: IMU propagation from its keyframe's new state, keyframe bias"""
        kf = self.kfs[-1]
        if last.id == kf["frame_id"]:
            last.Rwb, last.twb, last.vel = kf["Rwb"].copy(), kf["twb"].copy(), kf["vel"].copy()
        elif rows_to_last:
            rec = self.be.preintegrate(np.concatenate(rows_to_last), np.concatenate([kf["ba"], kf["bg"]]))
            dR = rec[0:9].reshape(3, 3).astype(np.float64); dV = rec[9:12].astype(np.float64); dP = rec[12:15].astype(np.float64)
            t = float(rec[285])
            last.Rwb = f32_64(_polar32(kf["Rwb"] @ dR))
            last.twb = f32_64(kf["twb"] + kf["vel"] * t + 0.5 * t * t * self.g + kf["Rwb"] @ dP)
            last.vel = f32_64(kf["vel"] + t * self.g + kf["Rwb"] @ dV)
        last.bg, last.ba = kf["bg"].copy(), kf["ba"].copy()


class CudaBackend:
    """every numerical step through this library's C ABI"""

    def __init__(self, max_points=2048, max_cloud=65536, ba_kf=32, ba_points=8192, ba_obs=65536):
        from .gicp import RegistrationGICP
        from .matcher import ORBmatcher
        from .optimizer import Optimizer
        from .orb import ORBextractor
        from .pose import PoseOptimizer
        from .pose_inertial import PoseInertialOptimizer
        self._orb = ORBextractor(1000, 1.2, 8, 25, 7, max_size=(640, 480), max_batch=1)
        self._gicp = RegistrationGICP(max_points=max_cloud, max_pairs=1)
        self._pose = PoseOptimizer(max_obs=max_points, max_batch=1)
        self._pin = PoseInertialOptimizer(max_obs=max_points, max_batch=1)
        self._ba = Optimizer(max_kf=ba_kf, max_points=ba_points, max_obs=ba_obs, max_inertial=ba_kf, max_batch=1)
        self._matcher = ORBmatcher()
        self.cam = None

    def orb(self, img):
        _, k, d = self._orb(img)
        return k, d

    def depth_to_cloud(self, depth, stride):
        from . import matcher, synth
        c = synth.G1_CAM
        return matcher.depth_to_cloud(depth, stride, c["fx"], c["fy"], c["cx"], c["cy"])

    def gicp(self, target, source, T0):
        return self._gicp.RegisterPointClouds(target, source, T0)

    def search_by_projection(self, mode, q, kps, u_right, desc, occupied, grid, nnratio, check_orientation):
        from . import matcher
        return matcher.search_by_projection(mode, q, kps, u_right, desc, occupied, grid, nnratio=nnratio, check_orientation=check_orientation)

    def search_with_gms(self, k1, d1, k2, d2, size):
        return self._matcher.SearchWithGMS(k1, d1, k2, d2, size)

    def pose_optimization(self, prob):
        return self._pose.PoseOptimization(prob)

    def preintegrate(self, rows, bias6):
        from . import imu, synth
        return imu.preintegrate_batch([rows], [bias6], *synth.imu_calib_noise())[0]

    def pose_inertial(self, prob):
        return self._pin.optimize_batch([prob])[0]

    def local_inertial_ba(self, prob):
        return self._ba.LocalInertialBA(prob)


def run_tracker(seq, backend, n_frames=None, **kw):
    return Tracker(seq, backend, **kw).run(n_frames)

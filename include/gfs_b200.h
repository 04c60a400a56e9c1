/* gfs_b200.h -- C ABI of the B200-native GeoFlow-SLAM hot path (libgfs_b200.so).
 *
 * The reference (HorizonRobotics/GeoFlowSlam) has no plugin/FFI layer: the drop-in boundary is
 * four C++ call sites.  Each entry point below names the reference interface it
 * replaces (file:line into the reference tree).  Plain pointers and sizes only; no C++/torch
 * types.  All functions return 0 on success or a negative GFS_ERR_* code; none throws.
 *
 * Pointer-space convention: names ending in _device take DEVICE pointers and enqueue work on
 * `stream` without synchronising (the caller owns synchronisation); the un-suffixed variants
 * take HOST pointers, stage through pinned memory, and return after the results are on the host.
 * `stream` is a cudaStream_t passed as void* (NULL = default stream).
 *
 * Threading: a handle owns its workspace; use one handle per calling thread (the reference calls
 * the extractor from the tracking thread or two pool threads, System.cc:570-596).
 */
#ifndef GFS_B200_H_
#define GFS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GFS_OK 0
#define GFS_ERR_INVALID (-1)  /* bad argument */
#define GFS_ERR_CUDA (-2)     /* CUDA runtime error, see gfs_last_error() */
#define GFS_ERR_CAPACITY (-3) /* input exceeds the capacity the handle was created with */
#define GFS_ERR_EMPTY (-4)    /* empty image: the reference returns -1 (ORBextractor.cc:1150) */
#define GFS_ERR_NODEVICE (-5) /* no CUDA device: there is NO CPU fallback */

const char* gfs_last_error(void);
/* 0 when a CUDA device is usable, GFS_ERR_NODEVICE otherwise. */
int gfs_device_check(void);
/* library / build identification, e.g. "gfs_b200 0.1 sm_100a" */
const char* gfs_version(void);

/* ------------------------------------------------------------------------------------------
 * ORB extractor -- replaces ORB_SLAM3::ORBextractor (include/ORBextractor.h:46-120,
 * src/ORBextractor.cc:421-479 ctor, :1145-1225 operator()).
 * ---------------------------------------------------------------------------------------- */

/* cv::KeyPoint fields the reference fills (pt, size, angle, response, octave). 24 bytes. */
typedef struct GfsKeyPoint {
  float x, y;     /* level-0 pixel coordinates (pt *= scale, ORBextractor.cc:1204) */
  float size;     /* PATCH_SIZE * mvScaleFactor[octave] truncated to int (:860) */
  float angle;    /* degrees, cv::fastAtan2 convention (:94) */
  float response; /* FAST score */
  int32_t octave;
} GfsKeyPoint;

typedef struct GfsOrb GfsOrb;

/* ORBextractor::ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)
 * (src/ORBextractor.cc:421).  max_w/max_h/max_batch size the device workspace. */
int gfs_orb_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int max_w,
                   int max_h, int max_batch, GfsOrb** out);
int gfs_orb_destroy(GfsOrb* h);
/* Row stride (in keypoints) of the per-frame output arrays: nfeatures + 3*nlevels rounded up to
 * a multiple of 32 (the quadtree may return up to N+2 points per level, ORBextractor.cc:703-707). */
int gfs_orb_max_keypoints(const GfsOrb* h);
/* GetScaleFactors() / mnFeaturesPerLevel (include/ORBextractor.h:69-83, :108). arrays of nlevels. */
int gfs_orb_tables(const GfsOrb* h, float* scale_factors, int* features_per_level);
/* level size as ComputePyramid computes it (src/ORBextractor.cc:1229-1231) */
int gfs_orb_level_size(const GfsOrb* h, int w, int h_img, int level, int* lw, int* lh);

/* int ORBextractor::operator()(image, mask, keypoints, descriptors, vLappingArea)
 * (include/ORBextractor.h:61-64).  Batched: `batch` gray images of w x h, row pitch `pitch`
 * bytes, consecutive images `img_stride` bytes apart.
 *   out_kp   [batch][gfs_orb_max_keypoints]       keypoints, mono from the front, lapping from the back
 *   out_desc [batch][gfs_orb_max_keypoints][32]   rBRIEF descriptors, same order
 *   out_n    [batch]  number of keypoints;  out_mono [batch]  monoIndex (the reference's return value)
 */
int gfs_orb_extract_batch_device(GfsOrb* h, void* stream, const uint8_t* d_imgs, int batch, int w, int h_img,
                                 int pitch, size_t img_stride, int lap0, int lap1, GfsKeyPoint* d_out_kp,
                                 uint8_t* d_out_desc, int* d_out_n, int* d_out_mono);
int gfs_orb_extract_batch(GfsOrb* h, void* stream, const uint8_t* imgs, int batch, int w, int h_img, int pitch,
                          size_t img_stride, int lap0, int lap1, GfsKeyPoint* out_kp, uint8_t* out_desc, int* out_n,
                          int* out_mono);
/* single frame, host pointers; returns GFS_ERR_EMPTY for a null/0-sized image. */
int gfs_orb_extract(GfsOrb* h, void* stream, const uint8_t* img, int w, int h_img, int pitch, int lap0, int lap1,
                    GfsKeyPoint* out_kp, uint8_t* out_desc, int* out_n, int* out_mono);

/* mvImagePyramid (public member, include/ORBextractor.h:82) and the blurred working copy
 * (ORBextractor.cc:1188-1189) of frame `frame` of the last batch, un-bordered, tightly packed
 * lw x lh bytes, copied to the host.  Debug / parity hook. */
int gfs_orb_get_level(GfsOrb* h, void* stream, int frame, int level, int blurred, uint8_t* out);
/* FAST candidates (vToDistributeKeys, ORBextractor.cc:853-870) of (frame, level) of the last batch
 * as (x, y, response) float triples relative to minBorder; returns the count through *n. */
int gfs_orb_get_candidates(GfsOrb* h, void* stream, int frame, int level, float* out_xyr, int cap, int* n);
/* Per-stage device timing of the last gfs_orb_extract_batch_device call (CUDA events on the call's
 * stream): ms6 = {pyramid, fast_cells, octree, blur, orient_desc, pack_lapping}.  Measurement aid
 * (bench.py roofline); the reference's analogue is REGISTER_TIMES (Frame.cc:352-365). */
int gfs_orb_set_profiling(GfsOrb* h, int enable);
int gfs_orb_get_profile(GfsOrb* h, float* ms6);
/* number of kernels one gfs_orb_extract_batch_device call launches (for bench gpu_launches). */
int gfs_orb_launches_per_call(const GfsOrb* h, int lap0, int lap1);

/* ------------------------------------------------------------------------------------------
 * Matcher -- replaces ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2536-2550), the
 * cv::BFMatcher(NORM_HAMMING).match calls and gms_matcher::GetInlierMask(false,false) inside
 * ORBmatcher::SearchWithGMS / SearchForInitializationWithGMS / SearchForTriangulationWithGMS
 * (src/ORBmatcher.cc:744-778, 786-841, 852-913; Thirdparty/GMS/include/gms_matcher.h).
 * ---------------------------------------------------------------------------------------- */

/* Batched brute-force Hamming match: for pair p, query descriptors dq[p] (nq[p] rows of 32 B) vs
 * train descriptors dt[p] (nt[p] rows).  out_idx[p][q] = argmin train index (ties -> lowest
 * index, as cv::BFMatcher), out_dist[p][q] = Hamming distance; -1/-1 when nt[p] == 0.
 * Descriptor arrays are [pairs][stride][32]; out arrays are [pairs][stride]. */
int gfs_match_bf_hamming_batch_device(void* stream, const uint8_t* d_dq, const int* d_nq, const uint8_t* d_dt,
                                      const int* d_nt, int pairs, int stride, int* d_out_idx, int* d_out_dist);
int gfs_match_bf_hamming(void* stream, const uint8_t* dq, int nq, const uint8_t* dt, int nt, int* out_idx,
                         int* out_dist);

/* Batched GMS: matches of pair p are (q, out_idx[p][q]) for q < nq[p] (what BFMatcher::match
 * returns).  kp1/kp2 are [pairs][stride] keypoints (only x, y are read), image sizes w x h.
 * out_inlier[p][q] in {0,1}; out_count[p] = number of inliers (the reference's return value). */
int gfs_gms_filter_batch_device(void* stream, const GfsKeyPoint* d_kp1, const int* d_n1, const GfsKeyPoint* d_kp2,
                                const int* d_n2, const int* d_train_idx, int pairs, int stride, int w1, int h1,
                                int w2, int h2, uint8_t* d_out_inlier, int* d_out_count);
/* single pair, host pointers, explicit match list (query, train) pairs as cv::DMatch order */
int gfs_gms_filter(void* stream, const GfsKeyPoint* kp1, int n1, int w1, int h1, const GfsKeyPoint* kp2, int n2,
                   int w2, int h2, const int* matches_qt, int nm, uint8_t* out_inlier, int* out_count);

/* Projection-window search -- replaces ORBmatcher::SearchByProjection(Frame&, const Frame&, th,
 * bMono) (mode 0, src/ORBmatcher.cc:1853-2063) and SearchByProjection(Frame&, vector<MapPoint*>&,
 * th, bFarPoints, thFarPoints) (mode 1, :43-207) for the RGB-D / monocular layout (Nleft == -1),
 * including Frame::GetFeaturesInArea / AssignFeaturesToGrid (src/Frame.cc:1007-1071, 734-761).
 * The shim flattens each candidate MapPoint into a query; the projection itself (Sophus / camera
 * model) stays with the caller. */
typedef struct GfsProjQuery {
  float u, v;     /* projected position (uv / mTrackProjX,Y) */
  float radius;   /* search radius (th * scale[octave], or r * scale[level]); < 0 = skip this map point */
  float ur;       /* projected right coordinate (uv.x - bf * invz / mTrackProjXR) */
  float angle;    /* keypoint angle in the last frame (rotation histogram, mode 0) */
  int32_t min_level, max_level; /* GetFeaturesInArea level window (-1 = open) */
  int32_t blocks; /* MapPoint::Observations() > 0: a keypoint it takes is skipped by later map points */
  uint8_t desc[32];
} GfsProjQuery;     /* 64 bytes */
/* Frame side: undistorted keypoints (x, y, octave, angle), mvuRight, descriptors, occupied[i] =
 * (mvpMapPoints[i] && Observations() > 0) before the call (may be NULL), grid parameters mnMinX,
 * mnMinY, mfGridElementWidthInv, mfGridElementHeightInv.  out_assign[i] = index of the query now
 * stored in mvpMapPoints[i] or -1 (untouched / cleared); *out_nmatches = the reference's return value. */
int gfs_search_by_projection(void* stream, int mode, float nnratio, int check_orientation, const GfsProjQuery* queries, int nq,
                             const GfsKeyPoint* kps_un, const float* u_right, const uint8_t* desc, const uint8_t* occupied,
                             int n, float min_x, float min_y, float inv_w, float inv_h, int* out_assign, int* out_nmatches);
int gfs_search_by_projection_batch_device(void* stream, int mode, float nnratio, int check_orientation,
                                          const GfsProjQuery* d_queries, const int* d_nq, int qstride,
                                          const GfsKeyPoint* d_kps_un, const float* d_u_right, const uint8_t* d_desc,
                                          const uint8_t* d_occupied, const int* d_n, int kstride, int frames, float min_x,
                                          float min_y, float inv_w, float inv_h, int* d_scratch_claim, int* d_out_assign,
                                          int* d_out_nmatches);

/* Frame::ConvertDepthToPointCloud (src/Frame.cc:590-623): every `stride`-th pixel with 0 < d < 10 m,
 * scan order, float4 (x, y, z, 1) -- the clouds RegistrationGICP consumes. */
int gfs_depth_to_cloud(void* stream, const float* depth, int w, int h, int stride, float fx, float fy, float cx, float cy,
                       float* out_xyz1, int cap, int* out_n);
int gfs_depth_to_cloud_batch_device(void* stream, const float* d_depth, int frames, int w, int h, int pitch_floats,
                                    size_t frame_stride_floats, int stride, float fx, float fy, float cx, float cy,
                                    float* d_out_xyz1, int cap, int* d_out_n);

/* ------------------------------------------------------------------------------------------
 * Batched tracking front-end (BASELINE.json configs[1]): ORB extraction of `batch` independent
 * frames, then BF-Hamming + GMS from frame i to frame i+1 -- the Frame::ExtractORB
 * (src/Frame.cc:768-777) + ORBmatcher::SearchWithGMS (src/ORBmatcher.cc:744-778) sequence of
 * System::TrackRGBD, for a whole batch per call.  Match outputs have batch-1 rows.
 * ---------------------------------------------------------------------------------------- */
typedef struct GfsFrontend GfsFrontend;
int gfs_frontend_create(int nfeatures, float scale_factor, int nlevels, int ini_th_fast, int min_th_fast, int max_w,
                        int max_h, int max_batch, GfsFrontend** out);
int gfs_frontend_destroy(GfsFrontend* f);
int gfs_frontend_max_keypoints(const GfsFrontend* f);
GfsOrb* gfs_frontend_orb(GfsFrontend* f);
int gfs_frontend_run_device(GfsFrontend* f, void* stream, const uint8_t* d_imgs, int batch, int w, int h_img,
                            int pitch, size_t img_stride, GfsKeyPoint* d_kp, uint8_t* d_desc, int* d_n, int* d_mono,
                            int* d_train_idx, int* d_dist, uint8_t* d_inlier, int* d_inlier_count);
int gfs_frontend_run(GfsFrontend* f, void* stream, const uint8_t* imgs, int batch, int w, int h_img, int pitch,
                     size_t img_stride, GfsKeyPoint* out_kp, uint8_t* out_desc, int* out_n, int* out_mono,
                     int* out_train_idx, int* out_dist, uint8_t* out_inlier, int* out_inlier_count);
/* ms8 = {pyramid, fast_cells, octree, blur, orient_desc, pack_lapping, bf_hamming, gms} of the last call */
int gfs_frontend_set_profiling(GfsFrontend* f, int enable);
int gfs_frontend_get_profile(GfsFrontend* f, float* ms8);
int gfs_frontend_launches_per_call(const GfsFrontend* f, int batch);

/* ------------------------------------------------------------------------------------------
 * GICP -- replaces RegistrationGICP::RegisterPointClouds (include/RegistrationGICP.h:25-28,
 * src/RegistrationGICP.cc:5-20) = small_gicp::align(target, source, init_T, setting) with
 * type GICP (Thirdparty/small_gicp/src/small_gicp/registration/registration_helper.cpp:56-68,
 * :111-120): voxel downsampling, 10-NN covariances, exact nearest-neighbour correspondences,
 * Levenberg-Marquardt.  Points are float4 (x, y, z, 1) like std::vector<Eigen::Vector4f>.
 * ---------------------------------------------------------------------------------------- */
typedef struct GfsGicpSetting {        /* small_gicp::RegistrationSetting (registration_helper.hpp:38-49) */
  double downsampling_resolution;      /* 0.02  (RegistrationGICP.cc:11) */
  double max_correspondence_distance;  /* 0.1   (RegistrationGICP.cc:12) */
  double rotation_eps;                 /* 0.1 deg in rad */
  double translation_eps;              /* 1e-3 */
  int num_neighbors;                   /* 10 (registration_helper.cpp:59) */
  int max_iterations;                  /* 20 */
  int num_threads;                     /* 4 in the reference; ignored */
} GfsGicpSetting;

typedef struct GfsGicpResult {  /* small_gicp::RegistrationResult (registration_result.hpp:11-30) */
  double T[16];                 /* T_target_source, row-major 4x4 */
  double H[36];                 /* information matrix of the last linearisation */
  double b[6];
  double error;
  int iterations;               /* zero-based index of the last outer iteration */
  int num_inliers;
  int converged;
  int n_target, n_source;       /* points after downsampling (measurement aid) */
  int inner_evals;              /* total lambda trials = error evaluations (measurement aid) */
} GfsGicpResult;

typedef struct GfsGicp GfsGicp;
void gfs_gicp_default_setting(GfsGicpSetting* s);
/* setting may be NULL (defaults above). max_points = per-cloud capacity, max_pairs = batch capacity. */
int gfs_gicp_create(const GfsGicpSetting* setting, int max_points, int max_pairs, GfsGicp** out);
int gfs_gicp_destroy(GfsGicp* h);
/* RegisterPointClouds(target, source, init_T_target_source) -> result; T0 row-major 4x4. */
int gfs_gicp_align(GfsGicp* h, void* stream, const float* target, int nt, const float* source, int ns, const double* T0,
                   GfsGicpResult* out);
/* Batched: clouds are [pairs][stride][4] floats with nt[p] / ns[p] valid points, T0 is [pairs][16].
 * The device variant enqueues on `stream` but synchronises it once per LM trial to read two
 * control counters (the optimizer's data-dependent loop, optimizer.hpp:97-141). */
int gfs_gicp_align_batch(GfsGicp* h, void* stream, const float* target, const int* nt, const float* source, const int* ns,
                         int pairs, int stride, const double* T0, GfsGicpResult* out);
int gfs_gicp_align_batch_device(GfsGicp* h, void* stream, const float* d_target, const int* d_nt, const float* d_source,
                                const int* d_ns, int pairs, int stride, const double* d_T0, GfsGicpResult* d_out);
/* Tracking mode -- the call pattern of Tracking::PredictStateICP (src/Tracking.cc:3364-3413): every frame the current
 * cloud (source) is registered against the last frame's cloud (target).  d_cloud is [seqs][stride][4] floats with d_n[s]
 * valid points: the NEW cloud of each of `seqs` independent sequences.  Each cloud is preprocessed (voxel grid, k-NN
 * grid, covariances) once and kept in HBM, instead of twice as small_gicp::align does (registration_helper.cpp:56-68);
 * results equal gfs_gicp_align_batch on (previous cloud, new cloud) bit for bit.  The first call after create / reset /
 * gfs_gicp_align* only stores the clouds (d_T0 / d_out may be NULL, nothing is written).  Same stream semantics as
 * gfs_gicp_align_batch_device.  gfs_gicp_track_batch is the host-pointer variant (copies in, results out, synchronises). */
int gfs_gicp_track_reset(GfsGicp* h);
int gfs_gicp_track_calls(const GfsGicp* h);   /* clouds stored per sequence since the last reset (0: the next call only stores) */
int gfs_gicp_track_batch_device(GfsGicp* h, void* stream, const float* d_cloud, const int* d_n, int seqs, int stride,
                                const double* d_T0, GfsGicpResult* d_out);
int gfs_gicp_track_batch(GfsGicp* h, void* stream, const float* cloud, const int* n, int seqs, int stride, const double* T0,
                         GfsGicpResult* out);
/* parity hook: downsampled points + covariances of cloud (2*pair = target, 2*pair+1 = source) */
int gfs_gicp_get_cloud(GfsGicp* h, void* stream, int cloud, double* out_xyz, double* out_cov6, int cap, int* n);
int gfs_gicp_last_launches(const GfsGicp* h);
/* measurement aid (bench.py's roofline): with profiling on, every align / track call records a CUDA event after each
 * stage on the caller's stream and synchronises once more at its end; ms8 / launches8[0..4] = grouping + voxel means + cell
 * pack, 10-NN covariances (k_knn_cov), correspondence search (k_nn_corr*), linearisation, LM bookkeeping of the last call */
int gfs_gicp_set_profiling(GfsGicp* h, int on);
int gfs_gicp_get_profile(const GfsGicp* h, float* ms8, int* launches8);
/* diagnostics: grid cells of `cloud` and the queries its cell-centric 10-NN pass handed to the per-query kernel */
int gfs_gicp_get_knn_stats(GfsGicp* h, void* stream, int cloud, int* n_cells, int* n_per_query);

/* ------------------------------------------------------------------------------------------
 * Local inertial bundle adjustment -- replaces the numerical core of
 * Optimizer::LocalInertialBA (include/Optimizer.h:180-183, src/Optimizer.cc:3056-3702): the g2o
 * problem it builds (VertexPose/Velocity/GyroBias/AccBias + VertexSBAPointXYZ, EdgeMono /
 * EdgeStereo with Huber, EdgeInertial with Huber, EdgeGyroRW / EdgeAccRW; src/G2oTypes.cc,
 * include/G2oTypes.h) and the Levenberg-Marquardt / Schur solve g2o runs on it
 * (Thirdparty/g2o/g2o/core/optimization_algorithm_levenberg.cpp:59-164, block_solver.hpp:354-560).
 * The pointer graph (KeyFrame*, MapPoint*, observations) is flattened by the shim into the plain
 * arrays below; window selection, outlier erasure and write-back stay with the caller.
 * ---------------------------------------------------------------------------------------- */
#define GFS_BA_PRE_STRIDE 292 /* floats per inertial edge: dR9 dV3 dP3 JRg9 JVg9 JVa9 JPg9 JPa9 C225 dT1
                                 b6 (bax bay baz bwx bwy bwz) -- IMU::Preintegrated members, ImuTypes.h */
typedef struct GfsBaProblem {
  int n_opt_kf;   /* optimizable keyframes (vpOptimizableKFs order, index 0 = pKF; Optimizer.cc:3078-3086) */
  int n_fixed_kf; /* fixed keyframes (lFixedKeyFrames, :3106-3165); stored after the optimizable ones */
  int n_points;   /* lLocalMapPoints */
  int n_obs;      /* visual edges, in creation order (:3446-3577) */
  int n_inertial; /* EdgeInertial (+ EdgeGyroRW + EdgeAccRW each), :3328-3401 */
  int iterations; /* optimizer.optimize(opt_it): 4 if bLarge else 8 (:3063-3068) */
  int b_large;
  double lambda_init; /* setUserLambdaInit: 1e-2 if bLarge else 1e0 (:3178-3187) */
  /* ImuCamPose calibration (G2oTypes.cc:48-56); Pinhole parameters are floats (Pinhole.cpp:36-42) */
  double Rcb[9], tcb[3], Rbc[9], tbc[3];
  float fx, fy, cx, cy;
  double bf;
  /* keyframe states [n_opt_kf + n_fixed_kf]; matrices row-major; float values widened like the
   * reference does when it loads them (G2oTypes.cc:30-52) */
  const double *kf_Rwb, *kf_twb, *kf_Rcw, *kf_tcw, *kf_vel, *kf_bg, *kf_ba;
  const uint8_t* kf_has_imu;
  const double* pt_xyz;    /* [n_points][3] */
  const uint8_t* pt_close; /* pMP->mTrackDepth < 10 (:3607) */
  const int *obs_kf, *obs_pt;
  const double* obs_uvr;       /* [n_obs][3] u, v, u_right; u_right < 0 -> EdgeMono (:3477, :3509) */
  const float* obs_inv_sigma2; /* mvInvLevelSigma2[octave] / unc2 (:3494, :3528) */
  const int *in_kf1, *in_kf2;  /* keyframe indices of (prev, cur) */
  const float* in_pre;         /* [n_inertial][GFS_BA_PRE_STRIDE] */
  const uint8_t* in_downweight; /* 1 -> information * 1e-2 (i == N-1, :3371) */
  /* optional EdgeICP factors (pbICPFlag, :3260-3321): vertex 0 = previous keyframe, vertex 1 = keyframe,
   * measurement T_c1_c2 = GICP result (row-major R then t, 12 doubles); information 1e2 * I, Huber
   * sqrt(0.4); Jacobians by g2o's central differences (delta 1e-9, base_binary_edge.hpp:124-190). */
  int n_icp;
  const int *icp_kf1, *icp_kf2;
  const double* icp_Rt; /* [n_icp][12] */
  /* 0: VertexPose keyframes (LocalInertialBA).  1: g2o::VertexSE3Expmap keyframes = Optimizer::LocalBundleAdjustment
   * (src/Optimizer.cc:1588-2040): EdgeSE3ProjectXYZ / g2o::EdgeStereoSE3ProjectXYZ, poses kf_Rcw /
   * kf_tcw updated as exp(u) * Tcw, no inertial / ICP edges (n_inertial = n_icp = 0), iterations = 10, lambda_init <= 0
   * (g2o's default tau * max diagonal), every observation with chi2 > 5.991 / 7.815 or non-positive depth flagged. */
  int vertex_se3;
} GfsBaProblem;

typedef struct GfsBaResult {
  /* optimised states, same layout as the inputs ([n_opt_kf + n_fixed_kf] keyframes) */
  double *kf_Rwb, *kf_twb, *kf_Rcw, *kf_tcw, *kf_vel, *kf_bg, *kf_ba, *pt_xyz;
  double* obs_chi2;            /* [n_obs] e->chi2() as the reference reads it after optimize (:3611, :3624) */
  uint8_t* obs_depth_positive; /* [n_obs] e->isDepthPositive() (mono edges; 1 for stereo) */
  uint8_t* obs_outlier;        /* [n_obs] the vToErase decision (:3603-3626) */
  float err, err_end;          /* activeRobustChi2 before / after (:3590-3592) */
  int failed;                  /* "FAIL LOCAL-INERTIAL BA" guard (:3632-3636): no write-back */
  int iterations_done;         /* g2o optimize() return value */
  int lm_trials;               /* total LM trials (measurement aid) */
  double lambda_final;
} GfsBaResult;

typedef struct GfsBa GfsBa;
/* capacity: keyframes (optimizable + fixed), points, observations, inertial edges per problem, problems per batch */
int gfs_ba_create(int max_kf, int max_points, int max_obs, int max_inertial, int max_batch, GfsBa** out);
int gfs_ba_destroy(GfsBa* h);
/* Solve `batch` independent problems (host pointers inside GfsBaProblem / GfsBaResult). */
int gfs_ba_solve_batch(GfsBa* h, void* stream, const GfsBaProblem* problems, GfsBaResult* results, int batch);
int gfs_ba_solve(GfsBa* h, void* stream, const GfsBaProblem* problem, GfsBaResult* result);
/* Two-phase variant for resident problems: upload once, then solve (re-solve) on the device;
 * gfs_ba_download copies the results of the last gfs_ba_solve_uploaded back. */
int gfs_ba_upload(GfsBa* h, void* stream, const GfsBaProblem* problems, int batch);
int gfs_ba_solve_uploaded(GfsBa* h, void* stream);
int gfs_ba_download(GfsBa* h, void* stream, GfsBaResult* results, int batch);
int gfs_ba_last_launches(const GfsBa* h);
/* Edge-partitioned mode: this rank owns the landmarks p with p % world == rank; the
 * reduced pose system is summed over ranks by `allreduce(buf, count_doubles, user)` (e.g. a
 * thin wrapper over ncclAllReduce / torch.distributed.all_reduce) once per LM trial. */
typedef int (*GfsAllReduceFn)(double* device_buf, int count, void* user);
int gfs_ba_set_partition(GfsBa* h, int rank, int world, GfsAllReduceFn allreduce, void* user);
/* The same partition with NCCL called directly: the per-trial sums -- [Hs | bs] before the reduced solve and
 * [chi2 partials | robust weights | gain-ratio scale] after the update -- go out as ONE grouped ncclAllReduce launch each on the
 * solve stream (NVLink / NVSwitch), with no host synchronisation and no callback.  libnccl.so.2 is bound with dlopen when this
 * is first called.  Rank 0 obtains the 128-byte id from gfs_nccl_unique_id and distributes it (e.g. a torch.distributed
 * broadcast); every rank then calls gfs_ba_set_partition_nccl with its CUDA device current.  world == 1 leaves partitioned mode. */
int gfs_nccl_unique_id(void* out128);
int gfs_ba_set_partition_nccl(GfsBa* h, int rank, int world, const void* unique_id128);
int gfs_ba_last_nccl_calls(const GfsBa* h);

/* ------------------------------------------------------------------------------------------------
 * Motion-only bundle adjustment of the tracking thread.
 * Replaces the g2o block of Optimizer::PoseOptimization (reference include/Optimizer.h:62-65,
 * src/Optimizer.cc:763-1099): one VertexSE3Expmap, EdgeSE3ProjectXYZOnlyPose (monocular, pinhole)
 * and g2o::EdgeStereoSE3ProjectXYZOnlyPose (stereo / RGB-D) edges with Huber kernels, 4 rounds of
 * 10 Levenberg iterations (dense 6x6 LDLT) each restarted from the frame's pose, chi2 classification
 * after every round (5.991 / 7.815), outliers moved to level 1, kernels removed after the third round.
 * The fork does not write the optimised pose back (the SetPose calls are commented out, :1086-1097):
 * the results the caller consumes are mvbOutlier, the average reprojection error and the inlier count.
 * Observations are the frame's features with a MapPoint, in feature order.
 * ------------------------------------------------------------------------------------------------ */
typedef struct GfsPoseProblem {
  int n_obs;               /* nInitialCorrespondences */
  float q_wxyz[4], t[3];   /* pFrame->GetPose(): unit quaternion (w, x, y, z) and translation of Tcw */
  float fx, fy, cx, cy, bf;/* Frame::fx, fy, cx, cy, mbf = pinhole parameters */
  const double* Xw;        /* [n_obs][3] pMP->GetWorldPos().cast<double>() */
  const float* uvr;        /* [n_obs][3] kpUn.pt.x, kpUn.pt.y, mvuRight (< 0: monocular observation) */
  const float* inv_sigma2; /* [n_obs] mvInvLevelSigma2[kpUn.octave] */
} GfsPoseProblem;
typedef struct GfsPoseResult {
  int n_inliers;           /* return value: nInitialCorrespondences - nBad (0 when n_obs < 3) */
  int n_bad, n_good;       /* nBad of the last round; nGood accumulates over the rounds as the reference's does */
  float avg_reproj_error;  /* SetFrame2FrameReprojError / SetFrame2MapReprojError value of the last round */
  int rounds_done;
  int lm_iterations[4];    /* Levenberg iterations g2o ran in each round */
  double q_wxyz[4], t[3];  /* vSE3_recov->estimate() */
  uint8_t* outlier;        /* [n_obs] mvbOutlier */
  float* chi2;             /* [n_obs] chi2 of the last round (may be NULL) */
} GfsPoseResult;
typedef struct GfsPose GfsPose;
int gfs_pose_create(int max_obs, int max_batch, GfsPose** out);
int gfs_pose_destroy(GfsPose* h);
/* Host pointers inside the structs; one CUDA block per frame, the whole optimisation in one launch. */
int gfs_pose_optimize_batch(GfsPose* h, void* stream, const GfsPoseProblem* problems, int batch, GfsPoseResult* results);
int gfs_pose_optimize(GfsPose* h, void* stream, const GfsPoseProblem* problem, GfsPoseResult* result);
int gfs_pose_last_launches(const GfsPose* h);

/* ------------------------------------------------------------------------------------------------
 * Inertial pose-only optimisers of the tracking thread.
 * Replace the g2o blocks of Optimizer::PoseInertialOptimizationLastKeyFrame (reference
 * include/Optimizer.h:80-83, src/Optimizer.cc:5899-6284) and Optimizer::PoseInertialOptimizationLastFrame
 * (include/Optimizer.h:92-95, src/Optimizer.cc:6762-7172; Tracking.cc:3777,3792 call them from
 * TrackLocalMap): VertexPose / VertexVelocity / VertexGyroBias / VertexAccBias of the frame (and, LastFrame,
 * of the previous frame), EdgeMonoOnlyPose / EdgeStereoOnlyPose with Huber (G2oTypes.cc:362-383,418-443),
 * EdgeInertial with Huber 6.0 (G2oTypes.cc:473-719), EdgeGyroRW / EdgeAccRW, EdgePriorPoseImu with Huber 5
 * (G2oTypes.cc:927-995), Gauss-Newton with the dense Eigen LDLT over all 15 / 30 unknowns
 * (optimization_algorithm_gauss_newton.cpp:49-93, linear_solver_dense.h:66-114), nIterations rounds of 10
 * iterations with chi2 classification, and the Hessian hand-over: GetHessian2 blocks (LastKeyFrame) or the
 * 30x30 Hessian marginalised over the previous frame (Optimizer::Marginalize, :4408-4487), clamped like
 * the ConstraintPoseImu constructor (G2oTypes.h:857-868).
 * ------------------------------------------------------------------------------------------------ */
#define GFS_PIN_LAST_KEYFRAME 0
#define GFS_PIN_LAST_FRAME 1
typedef struct GfsPoseInertialProblem {
  int mode;     /* GFS_PIN_LAST_KEYFRAME: previous state fixed; GFS_PIN_LAST_FRAME: previous frame free + prior */
  int n_obs;    /* features of the frame with a MapPoint, in feature order */
  int n_rounds; /* nIterations (<= 4) */
  int rec_init; /* bRecInit */
  float fx, fy, cx, cy, bf;       /* Pinhole parameters, Frame::mbf */
  double Rcb[9], tcb[3], tbc[3];  /* mImuCalib.mTcb / mTbc (G2oTypes.cc:96-101); row-major */
  /* the frame: ImuCamPose(Frame*) + VertexVelocity / GyroBias / AccBias (floats widened, G2oTypes.cc:76-101) */
  double Rwb[9], twb[3], Rcw[9], tcw[3], vel[3], bg[3], ba[3];
  /* mpLastKeyFrame (mode 0) or mpPrevFrame (mode 1) */
  double p_Rwb[9], p_twb[3], p_vel[3], p_bg[3], p_ba[3];
  const float* pre;         /* [GFS_BA_PRE_STRIDE] mpImuPreintegrated (mode 0) / mpImuPreintegratedFrame (mode 1) */
  float rw_Cg[9], rw_Ca[9]; /* pFrame->mpImuPreintegrated->C.block<3,3>(9,9) and (12,12): random-walk covariances */
  /* pFp->mpcpi (mode 1 only): ConstraintPoseImu members; H row-major 15x15 */
  double c_Rwb[9], c_twb[3], c_vwb[3], c_bg[3], c_ba[3], c_H[225];
  const double* Xw;         /* [n_obs][3] pMP->GetWorldPos().cast<double>() */
  const float* uvr;         /* [n_obs][3] kpUn.pt.x, kpUn.pt.y, mvuRight (< 0: EdgeMonoOnlyPose) */
  const float* inv_sigma2;  /* [n_obs] mvInvLevelSigma2[octave] / unc2 */
  const uint8_t* close;     /* [n_obs] mTrackDepth < 10 (monocular edges only) */
} GfsPoseInertialProblem;
typedef struct GfsPoseInertialResult {
  int n_inliers;            /* return value: nInitialCorrespondences - nBad */
  int n_bad, n_inliers_last;/* nBad as returned; nInliers of the last round */
  float avg_reproj_error;   /* SetFrame2FrameReprojError / SetFrame2MapReprojError value of the last round */
  int rounds_done;
  int gn_iterations[4];     /* iterations g2o ran in each round */
  double Rwb[9], twb[3], vel[3], bg[3], ba[3]; /* VP / VV / VG / VA estimates (the caller narrows to float) */
  double H[225];            /* pFrame->mpcpi->H: the prior handed to the next frame, after the eigenvalue clamp */
  uint8_t* outlier;         /* [n_obs] mvbOutlier */
  float* chi2;              /* [n_obs] chi2 of the last round (may be NULL) */
} GfsPoseInertialResult;
typedef struct GfsPoseInertial GfsPoseInertial;
int gfs_pose_inertial_create(int max_obs, int max_batch, GfsPoseInertial** out);
int gfs_pose_inertial_destroy(GfsPoseInertial* h);
/* Host pointers inside the structs; one CUDA block per frame, the whole optimisation in one launch. */
int gfs_pose_inertial_optimize_batch(GfsPoseInertial* h, void* stream, const GfsPoseInertialProblem* problems, int batch,
                                     GfsPoseInertialResult* results);
int gfs_pose_inertial_optimize(GfsPoseInertial* h, void* stream, const GfsPoseInertialProblem* problem,
                               GfsPoseInertialResult* result);
int gfs_pose_inertial_last_launches(const GfsPoseInertial* h);

/* ------------------------------------------------------------------------------------------------
 * Optical-flow front end -- replaces cv::buildOpticalFlowPyramid(image, mImGray, winSize, 3)
 * (reference src/Frame.cc:370-373) and ORBmatcher::fbKltTracking / Tracking::fbKltTracking
 * (include/ORBmatcher.h:56-60, src/ORBmatcher.cc:2186-2293; src/Tracking.cc:3262-3360): forward
 * cv::calcOpticalFlowPyrLK over nbpyrlvl levels (OPTFLOW_USE_INITIAL_FLOW | OPTFLOW_LK_GET_MIN_EIGENVALS, 30
 * iterations, eps 0.01), status / min-eigenvalue / inBorder filter, backward pass at level 0, forward-backward
 * distance check.  The F-matrix RANSAC that follows in SearchByProjectionWithOF (cv::findFundamentalMat,
 * ORBmatcher.cc:2399,2463) is NOT part of this entry point.
 * A frame's pyramid is one caller-owned device block of gfs_klt_pyramid_bytes() bytes: the images of levels
 * 0..levels tightly packed (level l is ((w+1)/2, (h+1)/2) of level l-1), then, 16-byte aligned, the int16
 * (dI/dx, dI/dy) Scharr derivatives in the same order, then every level once more with the reflect-101 border
 * cv::buildOpticalFlowPyramid's copyMakeBorder gives mImGray (48 x 40 pixels, row pitch a multiple of 16 bytes; the
 * tracker fetches its windows from these copies with tensor-map TMA loads).  Reads outside the image follow the
 * border rules of the padded OpenCV pyramid: reflect-101 image, zero derivative.  gfs_klt_pyramid_layout describes
 * the first two parts (the ones with an OpenCV counterpart a caller may want to read).
 * ------------------------------------------------------------------------------------------------ */
typedef struct GfsKlt GfsKlt;
int gfs_klt_create(int max_w, int max_h, int levels, int max_points, int max_batch, GfsKlt** out);
int gfs_klt_destroy(GfsKlt* h);
size_t gfs_klt_pyramid_bytes(const GfsKlt* h, int w, int h_img);
/* level sizes, pixel offsets of the levels and the byte offset of the derivative block; arrays of levels + 1 */
int gfs_klt_pyramid_layout(const GfsKlt* h, int w, int h_img, int* level_w, int* level_h, int* level_off, size_t* deriv_offset);
/* cv::buildOpticalFlowPyramid for `batch` gray frames (device pointers); d_pyr is [batch][gfs_klt_pyramid_bytes] */
int gfs_klt_build_pyramid_batch_device(GfsKlt* h, void* stream, const uint8_t* d_imgs, int batch, int w, int h_img, int pitch,
                                       size_t img_stride, uint8_t* d_pyr);
/* ORBmatcher::fbKltTracking for `batch` frame pairs: kps [batch][stride][2] (vkps), priors in/out (vpriorkps),
 * n [batch] points per pair, status [batch][stride] (vkpstatus) */
int gfs_klt_fb_track_batch_device(GfsKlt* h, void* stream, const uint8_t* d_prev_pyr, const uint8_t* d_cur_pyr, int batch, int w, int h_img,
                                  const float* d_kps, float* d_priors, const int* d_n, int stride, int win, int nbpyrlvl, float ferr,
                                  float fmax_fbklt_dist, uint8_t* d_status);
/* cv::calcOpticalFlowPyrLK on prebuilt pyramids with OPTFLOW_LK_GET_MIN_EIGENVALS (d_err = minimum eigenvalues, may
 * be NULL) and optionally OPTFLOW_USE_INITIAL_FLOW; parity hook against OpenCV itself */
int gfs_klt_calc_batch_device(GfsKlt* h, void* stream, const uint8_t* d_prev_pyr, const uint8_t* d_cur_pyr, int batch, int w, int h_img,
                              const float* d_pts, float* d_next, const int* d_n, int stride, int win, int max_level, int max_count,
                              float eps, int use_initial_flow, uint8_t* d_status, float* d_err);
/* one frame pair, HOST pointers: images in, tracks out (both pyramids are built on the way) */
int gfs_klt_fb_track(GfsKlt* h, void* stream, const uint8_t* prev_img, const uint8_t* cur_img, int w, int h_img, int pitch, const float* kps,
                     float* priors, int n, int win, int nbpyrlvl, float ferr, float fmax_fbklt_dist, uint8_t* status);
int gfs_klt_last_launches(const GfsKlt* h);
/* cv::createCLAHE(clip_limit, Size(tiles_x, tiles_y))->apply on 8-bit gray frames -- replaces the optional in-place
 * CLAHE pass of Frame::Frame (reference src/Frame.cc:366-368, 498-500: clip 3.0, 8x8 tiles).  d_dst may alias d_src. */
int gfs_clahe_apply_batch_device(void* stream, const uint8_t* d_src, int batch, int w, int h_img, int pitch, size_t img_stride,
                                 double clip_limit, int tiles_x, int tiles_y, uint8_t* d_dst, int dst_pitch, size_t dst_stride);
int gfs_clahe_apply(void* stream, const uint8_t* src, int w, int h_img, int pitch, double clip_limit, int tiles_x, int tiles_y, uint8_t* dst);

/* ------------------------------------------------------------------------------------------------
 * IMU preintegration -- replaces IMU::Preintegrated::Initialize + IntegrateNewMeasurement over
 * an interval's measurements (reference include/ImuTypes.h:172-274, src/ImuTypes.cc:163-246; IntegratedRotation
 * :87-112; Calib::Set :399-412) for a batch of independent intervals: what Tracking::PreintegrateIMU,
 * Preintegrated::Reintegrate (:176-182) and MergePrevious (:248-269) compute one interval at a time.
 *   meas    [offsets[n]][7]  acceleration xyz, angular velocity xyz, dt -- the (acc, angVel, tstep) triples the
 *                            caller forms from consecutive IMU samples (Tracking.cc mid-point rule)
 *   offsets [n + 1]          interval i owns rows offsets[i] .. offsets[i+1]-1 (an empty interval gives the
 *                            Initialize() state)
 *   bias    [n][6]           bax bay baz bwx bwy bwz (IMU::Bias)
 *   ng, na, ngw, naw         the IMU::Calib noise densities (already scaled by sqrt(frequency) as Settings does)
 *   out     [n][GFS_BA_PRE_STRIDE]  dR9 dV3 dP3 JRg9 JVg9 JVa9 JPg9 JPa9 C225 dT1 b6: the record GfsBaProblem.in_pre
 *                            and GfsPoseInertialProblem.pre take
 * ------------------------------------------------------------------------------------------------ */
int gfs_imu_preintegrate_batch(void* stream, const float* meas, const int* offsets, const float* bias, int n, float ng, float na, float ngw,
                               float naw, float* out);
int gfs_imu_preintegrate_batch_device(void* stream, const float* d_meas, const int* d_offsets, const float* d_bias, int n, float ng,
                                      float na, float ngw, float naw, float* d_out);

#ifdef __cplusplus
}
#endif
#endif /* GFS_B200_H_ */

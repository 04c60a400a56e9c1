// Local inertial bundle adjustment on sm_100a, batched over independent problems.
//
// Replaces the numerical core of Optimizer::LocalInertialBA (reference src/Optimizer.cc:3056-3702):
// the g2o graph of VertexPose/Velocity/GyroBias/AccBias + VertexSBAPointXYZ with EdgeMono /
// EdgeStereo / EdgeInertial / EdgeGyroRW / EdgeAccRW (src/G2oTypes.cc, include/G2oTypes.h) and the
// Levenberg-Marquardt + Schur solve g2o runs on it (optimization_algorithm_levenberg.cpp:59-164,
// block_solver.hpp:354-560).  fp64 throughout, float32 where the reference is (IMU deltas).
//
//   k_vis_error      thread per visual edge: residual, chi2, Huber rho -> block partial sums
//   k_in_error       thread per inertial edge (+ its two random-walk edges)
//   k_lin_points     thread per landmark: Jacobians of its edges, H_ll, b_l, H_pl blocks, per-edge pose terms
//   k_lin_kf         warp per keyframe: fixed-order sum of its edges' 6x6 / 6x1 pose terms
//   k_lin_inertial   CTA per problem: 9x24 inertial Jacobians, J^T W J into the dense pose block
//   k_schur_prep     thread per landmark: (H_ll + lambda I)^-1, D^-1 b_l
//   k_schur_pairs    warp per keyframe pair: sum over shared landmarks of (B_i D^-1) B_j^T
//   k_schur_rhs      warp per keyframe
//   k_ldlt_solve     CTA per problem: dense LDL^T of the reduced pose system (<= 15 * n_kf unknowns)
//   k_backsub_update landmark back-substitution and the vertices' oplus into the trial state
//   k_lm_*           the Levenberg control flow, one thread per problem
// Every sum has a fixed order (no floating-point atomics): results are run-to-run reproducible.
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

#include <dlfcn.h>

#include <mutex>

#include "common.cuh"
#include "inertial.cuh"

namespace gfs {

static const int ERR_THREADS = 128;
// per-problem scalar state
enum { D_LAMBDA = 0, D_NI, D_CUR, D_TEMP, D_INI, D_RHO, D_LAST, D_ERR0, D_SCALE, D_NSTATE = 12 };
enum { J_ACTIVE = 0, J_NEED, J_NBAD, J_QMAX, J_IT, J_DONE, J_TRIALS, J_OK, J_PHASE, J_NSTATE = 12 };

struct BaCalib {
  double Rcb[9], tcb[3], Rbc[9], tbc[3];
  double fx, fy, cx, cy, bf, lambda_init;
  double deltaMono, deltaStereo;  // (float)sqrt(5.991), (float)sqrt(7.815) widened (Optimizer.cc:3427-3429)
  int nOpt, nKf, nPt, nObs, nIn, nIcp, iterations, bLarge, dimP;
  int se3;  // 1: g2o::VertexSE3Expmap keyframes (Optimizer::LocalBundleAdjustment), 0: VertexPose (LocalInertialBA)
};

struct BaDev {
  int maxKf, maxOpt, maxPt, maxObs, maxIn, maxDim, nblk;  // capacities (per problem); maxOpt = optimizable keyframes
  int rank, world;                                 // landmark partition (p % world == rank), default 0 / 1
  BaCalib* calib;       // [B]
  double *kf, *kfBak;   // [B][maxKf][KF_STRIDE]
  double *pt, *ptBak;   // [B][maxPt][3]
  const uint8_t *kfImu, *ptClose;          // [B][maxKf], [B][maxPt]
  const int *obsKf, *obsPt;                // [B][maxObs]
  const double* obsUvr;                    // [B][maxObs][3]
  const float* obsW;                       // [B][maxObs]
  const int *inKf1, *inKf2;                // [B][maxIn]
  const float* inPre;                      // [B][maxIn][292]
  const double *infoIn, *infoG, *infoA;    // [B][maxIn][81], [9], [9]
  const int *icpKf1, *icpKf2;              // [B][maxIn] EdgeICP vertices (previous keyframe, keyframe)
  const double* icpRt;                     // [B][maxIn][12] measured T_c1_c2
  double* icpRho;                          // [B][maxIn]
  const int *ptStart, *ptEdges;            // CSR landmarks -> edges: [B][maxPt+1], [B][maxObs]
  const int *kfStart, *kfEdges;            // CSR optimizable keyframes -> edges: [B][maxKf+1], [B][maxObs]
  const int* ptKfEdge;                     // [B][maxPt][maxOpt] edge id of (landmark, optimizable keyframe) or -1
  double* chi2;         // [B][maxObs] chi2 of every visual edge at the last evaluation
  double* inRho;        // [B][maxIn] robustified chi2 of inertial edge + its RW edges
  double* partChi;      // [B][nblk]
  double* E;            // [B][maxObs][18]  H_pl block of the edge (6x3)
  double* Ae;           // [B][maxObs][27]  pose terms of the edge: 21 (upper 6x6) + 6
  double *Hll, *bl, *Dinv, *db;  // [B][maxPt][9|3|9|3]
  double *Hpp, *bp, *Hs, *bs;    // [B][maxDim^2], [B][maxDim]
  double* Hin;          // [B][maxIn][24*24 + 24]
  double* x;            // [B][maxDim + 3 maxPt]
  double* partScale;    // [B][nblk]
  double* dstate;       // [B][D_NSTATE]
  int* istate;          // [B][J_NSTATE]
  int* counters;        // [4]
};

// ------------------------------------------------------------------------------------------------
// small fp64 helpers (same formulas as oracle/ba_oracle.cpp)
// ------------------------------------------------------------------------------------------------
#define BA_DELTA_INERTIAL 4.0 /* sqrt(16.0), Optimizer.cc:3372 */

// ---- g2o::SE3Quat restated (Thirdparty/g2o/g2o/types/se3quat.h:40-215) for EdgeICP (include/G2oTypes.h:508-572)
struct Quat { double w, x, y, z; };
struct SE3Q { Quat r; double t[3]; };
__device__ Quat quat_from_R(const double* m) {  // Eigen::Quaterniond(Matrix3d)
  Quat q;
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m[7] - m[5]) * t; q.y = (m[2] - m[6]) * t; q.z = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
    double v[3];
    v[i] = 0.5 * t;
    t = 0.5 / t;
    q.w = (m[3 * k + j] - m[3 * j + k]) * t;
    v[j] = (m[3 * j + i] + m[3 * i + j]) * t;
    v[k] = (m[3 * k + i] + m[3 * i + k]) * t;
    q.x = v[0]; q.y = v[1]; q.z = v[2];
  }
  return q;
}
__device__ void quat_normalize_rot(Quat& q) {
  if (q.w < 0) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  const double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  q.w /= n; q.x /= n; q.y /= n; q.z /= n;
}
__device__ Quat quat_mul(const Quat& a, const Quat& b) {
  return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
__device__ void quat_rot(const Quat& q, const double* v, double* o) {
  const double ux = 2 * (q.y * v[2] - q.z * v[1]), uy = 2 * (q.z * v[0] - q.x * v[2]), uz = 2 * (q.x * v[1] - q.y * v[0]);
  o[0] = v[0] + q.w * ux + (q.y * uz - q.z * uy);
  o[1] = v[1] + q.w * uy + (q.z * ux - q.x * uz);
  o[2] = v[2] + q.w * uz + (q.x * uy - q.y * ux);
}
__device__ void quat_to_R(const Quat& q, double* R) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y,
               tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ SE3Q se3q_make(const double* R, const double* t) {
  SE3Q s; s.r = quat_from_R(R); quat_normalize_rot(s.r);
  s.t[0] = t[0]; s.t[1] = t[1]; s.t[2] = t[2];
  return s;
}
__device__ SE3Q se3q_mul(const SE3Q& a, const SE3Q& b) {
  SE3Q r = a;
  double rt[3];
  quat_rot(a.r, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] += rt[i];
  r.r = quat_mul(a.r, b.r);
  quat_normalize_rot(r.r);
  return r;
}
__device__ SE3Q se3q_inv(const SE3Q& a) {
  SE3Q r;
  r.r = Quat{a.r.w, -a.r.x, -a.r.y, -a.r.z};
  const double nt[3] = {-a.t[0], -a.t[1], -a.t[2]};
  quat_rot(r.r, nt, r.t);
  return r;
}
__device__ void se3q_log(const SE3Q& a, double* res) {
  double R[9];
  quat_to_R(a.r, R);
  const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
  const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  double omega[3], Om[9], Om2[9], Vi[9];
  if (d > 0.99999) {
    for (int i = 0; i < 3; i++) omega[i] = 0.5 * dR[i];
    skew3(omega, Om);
    mm3(Om, Om, Om2);
    for (int i = 0; i < 9; i++) Vi[i] = (i % 4 == 0 ? 1.0 : 0.0) - 0.5 * Om[i] + (1. / 12.) * Om2[i];
  } else {
    const double theta = acos(d);
    const double k = theta / (2 * sqrt(1 - d * d));
    for (int i = 0; i < 3; i++) omega[i] = k * dR[i];
    skew3(omega, Om);
    mm3(Om, Om, Om2);
    const double c = (1 - theta / (2 * tan(theta / 2))) / (theta * theta);
    for (int i = 0; i < 9; i++) Vi[i] = (i % 4 == 0 ? 1.0 : 0.0) - 0.5 * Om[i] + c * Om2[i];
  }
  double ups[3];
  mv3(Vi, a.t, ups);
  for (int i = 0; i < 3; i++) { res[i] = omega[i]; res[3 + i] = ups[i]; }
}
// EdgeICP::computeError: log(T_c1c2^-1 * T_c1w * T_c2w^-1); s1, s2 = keyframe state records
__device__ void icp_error(const double* Rt, const double* s1, const double* s2, double* err6) {
  const SE3Q Tm = se3q_make(Rt, Rt + 9), T1 = se3q_make(s1 + K_RCW, s1 + K_TCW), T2 = se3q_make(s2 + K_RCW, s2 + K_TCW);
  se3q_log(se3q_mul(se3q_mul(se3q_inv(Tm), T1), se3q_inv(T2)), err6);
}
#define BA_DELTA_ICP ((double)(float)0.632455532033675866) /* (float)sqrt(0.4), Optimizer.cc:3258 */

// visual residual: obs - Project[Stereo](Xw) (G2oTypes.cc:172-188, Pinhole.cpp:36-42)
__device__ __forceinline__ int vis_error(const BaCalib& C, const double* kf, const double* Xw, const double* obs, double* err,
                                         double* Xc) {
  mv3(kf + K_RCW, Xw, Xc);
  Xc[0] += kf[K_TCW]; Xc[1] += kf[K_TCW + 1]; Xc[2] += kf[K_TCW + 2];
  const double u = C.fx * Xc[0] / Xc[2] + C.cx, v = C.fy * Xc[1] / Xc[2] + C.cy;
  err[0] = obs[0] - u;
  err[1] = obs[1] - v;
  if (obs[2] < 0) return 2;
  if (C.se3) {  // g2o::EdgeStereoSE3ProjectXYZ::cam_project narrows 1/z to float (types_six_dof_expmap.cpp:213-221)
    const float invz = (float)(1.0 / Xc[2]);
    const double us = Xc[0] * (double)invz * C.fx + C.cx;
    err[0] = obs[0] - us;
    err[1] = obs[1] - (Xc[1] * (double)invz * C.fy + C.cy);
    err[2] = obs[2] - (us - C.bf * (double)invz);
    return 3;
  }
  const double invZ = 1 / Xc[2];
  err[2] = obs[2] - (u - C.bf * invZ);
  return 3;
}

// SE3Quat::exp (Thirdparty/g2o/g2o/types/se3quat.h:223-257)
__device__ SE3Q se3q_exp(const double* u) {
  const double theta = sqrt(u[0] * u[0] + u[1] * u[1] + u[2] * u[2]);
  double O[9], O2[9], R[9], V[9];
  skew3(u, O);
  mm3(O, O, O2);
  if (theta < 0.00001) {
    for (int i = 0; i < 9; i++) { R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / pow(theta, 3.0);
    for (int i = 0; i < 9; i++) {
      const double I = (i % 4 == 0) ? 1.0 : 0.0;
      R[i] = I + a * O[i] + b * O2[i];
      V[i] = I + b * O[i] + c * O2[i];
    }
  }
  SE3Q s;
  s.r = quat_from_R(R);
  quat_normalize_rot(s.r);
  mv3(V, u + 3, s.t);
  return s;
}
// ImuCamPose::Update (G2oTypes.cc:191-217) on a keyframe state record (pose part); in SE3 mode
// g2o::VertexSE3Expmap::oplusImpl (types_six_dof_expmap.h:55-70): Tcw <- exp(u) * Tcw
__device__ void kf_oplus(const BaCalib& C, double* st, const double* u) {
  if (C.se3) {
    const SE3Q T = se3q_mul(se3q_exp(u), se3q_make(st + K_RCW, st + K_TCW));
    quat_to_R(T.r, st + K_RCW);
    st[K_TCW] = T.t[0]; st[K_TCW + 1] = T.t[1]; st[K_TCW + 2] = T.t[2];
    return;
  }
  double t[3], E[9];
  mv3(st + K_RWB, u + 3, t);
  st[K_TWB] += t[0]; st[K_TWB + 1] += t[1]; st[K_TWB + 2] += t[2];
  exp_so3(u, E);
  mm3(st + K_RWB, E, st + K_RWB);
  // (NormalizeRotation(Rwb) every third update: its result is discarded in the reference, :203-207)
  double Rbw[9], tbw[3];
  mt3(st + K_RWB, Rbw);
  mv3(Rbw, st + K_TWB, tbw);
  tbw[0] = -tbw[0]; tbw[1] = -tbw[1]; tbw[2] = -tbw[2];
  mm3(C.Rcb, Rbw, st + K_RCW);
  mv3(C.Rcb, tbw, st + K_TCW);
  st[K_TCW] += C.tcb[0]; st[K_TCW + 1] += C.tcb[1]; st[K_TCW + 2] += C.tcb[2];
}

// fixed-order block sum of one double per thread -> out (thread 0 writes)
template <int THREADS>
__device__ __forceinline__ void block_sum_store(double v, double* out) {
  __shared__ double s_w[THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  if (lane == 0) s_w[warp] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0;
#pragma unroll
    for (int w = 0; w < THREADS / 32; w++) s += s_w[w];
    *out = s;
  }
}

__device__ __forceinline__ bool problem_on(const BaDev& D, int b, int phaseFlag) { return D.istate[b * J_NSTATE + phaseFlag] != 0; }
__device__ __forceinline__ bool owned(const BaDev& D, int j) { return D.world <= 1 || (j % D.world) == D.rank; }

// ------------------------------------------------------------------------------------------------
// errors
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(ERR_THREADS) k_vis_error(BaDev D, int flag) {
  const int b = blockIdx.y;
  if (!problem_on(D, b, flag)) return;
  const BaCalib& C = D.calib[b];
  const int e = blockIdx.x * ERR_THREADS + threadIdx.x;
  double r0 = 0;
  if (e < C.nObs) {
    const int k = D.obsKf[(size_t)b * D.maxObs + e], j = D.obsPt[(size_t)b * D.maxObs + e];
    if (owned(D, j)) {
      double r[3], Xc[3];
      const int d = vis_error(C, D.kf + ((size_t)b * D.maxKf + k) * KF_STRIDE, D.pt + ((size_t)b * D.maxPt + j) * 3,
                              D.obsUvr + ((size_t)b * D.maxObs + e) * 3, r, Xc);
      const double w = (double)D.obsW[(size_t)b * D.maxObs + e];
      double c2 = 0;
      for (int a = 0; a < d; a++) c2 += r[a] * w * r[a];
      D.chi2[(size_t)b * D.maxObs + e] = c2;
      double rho[2];
      huber(c2, d == 2 ? C.deltaMono : C.deltaStereo, rho);
      r0 = rho[0];
    }
  }
  block_sum_store<ERR_THREADS>(r0, D.partChi + (size_t)b * D.nblk + blockIdx.x);
}

// EdgeInertial::computeError (G2oTypes.cc:495-522); err9 out
__device__ void inertial_error(const BaDev& D, int b, int e, double* err9) {
  const int k1 = D.inKf1[(size_t)b * D.maxIn + e], k2 = D.inKf2[(size_t)b * D.maxIn + e];
  inertial_error_core(D.kf + ((size_t)b * D.maxKf + k1) * KF_STRIDE, D.kf + ((size_t)b * D.maxKf + k2) * KF_STRIDE,
                      D.inPre + ((size_t)b * D.maxIn + e) * GFS_BA_PRE_STRIDE, err9);
}

__global__ void k_in_error(BaDev D, int flag) {
  const int b = blockIdx.y;
  if (!problem_on(D, b, flag)) return;
  const BaCalib& C = D.calib[b];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < C.nIcp) {  // EdgeICP: information 1e2 * I, Huber sqrt(0.4) (Optimizer.cc:3258, 3306-3314)
    double v = 0;
    if (D.rank == 0) {
      double r[6], rho[2];
      icp_error(D.icpRt + ((size_t)b * D.maxIn + e) * 12, D.kf + ((size_t)b * D.maxKf + D.icpKf1[(size_t)b * D.maxIn + e]) * KF_STRIDE,
                D.kf + ((size_t)b * D.maxKf + D.icpKf2[(size_t)b * D.maxIn + e]) * KF_STRIDE, r);
      double c2 = 0;
      for (int a = 0; a < 6; a++) c2 += r[a] * 1e2 * r[a];
      huber(c2, BA_DELTA_ICP, rho);
      v = rho[0];
    }
    D.icpRho[(size_t)b * D.maxIn + e] = v;
  }
  if (e >= C.nIn) return;
  double out = 0;
  if (D.rank == 0) {  // inertial edges belong to rank 0 in partitioned mode
    double r[9], rho[2];
    inertial_error(D, b, e, r);
    const double* I = D.infoIn + ((size_t)b * D.maxIn + e) * 81;
    double c2 = 0;
    for (int a = 0; a < 9; a++)
      for (int c = 0; c < 9; c++) c2 += r[a] * I[9 * a + c] * r[c];
    huber(c2, BA_DELTA_INERTIAL, rho);
    out = rho[0];
    const int k1 = D.inKf1[(size_t)b * D.maxIn + e], k2 = D.inKf2[(size_t)b * D.maxIn + e];
    const double* s1 = D.kf + ((size_t)b * D.maxKf + k1) * KF_STRIDE;
    const double* s2 = D.kf + ((size_t)b * D.maxKf + k2) * KF_STRIDE;
    const double *G = D.infoG + ((size_t)b * D.maxIn + e) * 9, *A = D.infoA + ((size_t)b * D.maxIn + e) * 9;
    double rg[3], ra[3], cg = 0, ca = 0;
    for (int i = 0; i < 3; i++) { rg[i] = s2[K_BG + i] - s1[K_BG + i]; ra[i] = s2[K_BA + i] - s1[K_BA + i]; }
    for (int a = 0; a < 3; a++)
      for (int c = 0; c < 3; c++) { cg += rg[a] * G[3 * a + c] * rg[c]; ca += ra[a] * A[3 * a + c] * ra[c]; }
    out += cg + ca;
  }
  D.inRho[(size_t)b * D.maxIn + e] = out;
}

// ------------------------------------------------------------------------------------------------
// linearisation
// ------------------------------------------------------------------------------------------------
// EdgeMono / EdgeStereo linearizeOplus (G2oTypes.cc:335-360, 385-416) + constructQuadraticForm
// (base_binary_edge.hpp:56-118) for every edge of one landmark.
__global__ void __launch_bounds__(128) k_lin_points(BaDev D) {
  const int b = blockIdx.y;
  if (!problem_on(D, b, J_ACTIVE)) return;
  const BaCalib& C = D.calib[b];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= C.nPt) return;
  double Hll[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0}, bl[3] = {0, 0, 0};
  if (owned(D, j)) {
    const double* Xw = D.pt + ((size_t)b * D.maxPt + j) * 3;
    const int s = D.ptStart[(size_t)b * (D.maxPt + 1) + j], t = D.ptStart[(size_t)b * (D.maxPt + 1) + j + 1];
    for (int q = s; q < t; q++) {
      const int e = D.ptEdges[(size_t)b * D.maxObs + q];
      const int k = D.obsKf[(size_t)b * D.maxObs + e];
      const double* kf = D.kf + ((size_t)b * D.maxKf + k) * KF_STRIDE;
      double r[3], Xc[3], rho[2];
      const int d = vis_error(C, kf, Xw, D.obsUvr + ((size_t)b * D.maxObs + e) * 3, r, Xc);
      const double w0 = (double)D.obsW[(size_t)b * D.maxObs + e];
      double c2 = 0;
      for (int a = 0; a < d; a++) c2 += r[a] * w0 * r[a];
      huber(c2, d == 2 ? C.deltaMono : C.deltaStereo, rho);
      const double w = rho[1] * w0;
      double pj[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
      pj[0] = C.fx / Xc[2]; pj[2] = -C.fx * Xc[0] / (Xc[2] * Xc[2]);
      pj[4] = C.fy / Xc[2]; pj[5] = -C.fy * Xc[1] / (Xc[2] * Xc[2]);
      if (d == 3) { pj[6] = pj[0]; pj[7] = pj[1]; pj[8] = pj[2] + C.bf * (1.0 / (Xc[2] * Xc[2])); }
      double Jl[9];
      for (int a = 0; a < 3; a++)
        for (int c = 0; c < 3; c++)
          Jl[3 * a + c] = (a < d) ? -(pj[3 * a] * kf[K_RCW + c] + pj[3 * a + 1] * kf[K_RCW + 3 + c] + pj[3 * a + 2] * kf[K_RCW + 6 + c]) : 0.0;
      for (int a = 0; a < 3; a++) {
        double tt = 0;
        for (int q2 = 0; q2 < d; q2++) tt += Jl[3 * q2 + a] * (-w * r[q2]);
        bl[a] += tt;
        for (int c = 0; c < 3; c++) {
          double h = 0;
          for (int q2 = 0; q2 < d; q2++) h += Jl[3 * q2 + a] * w * Jl[3 * q2 + c];
          Hll[3 * a + c] += h;
        }
      }
      if (k < C.nOpt) {
        // VertexPose: proj_jac * Rcb * SE3deriv(Xb); VertexSE3Expmap (EdgeSE3ProjectXYZ, g2o::EdgeStereoSE3ProjectXYZ):
        // -proj_jac * SE3deriv(Xc)
        double Xb[3];
        if (C.se3) {
          Xb[0] = Xc[0]; Xb[1] = Xc[1]; Xb[2] = Xc[2];
        } else {
          mv3(C.Rbc, Xc, Xb);
          Xb[0] += C.tbc[0]; Xb[1] += C.tbc[1]; Xb[2] += C.tbc[2];
        }
        const double Dm[18] = {0.0, Xb[2], -Xb[1], 1.0, 0.0, 0.0, -Xb[2], 0.0, Xb[0], 0.0, 1.0, 0.0, Xb[1], -Xb[0], 0.0, 0.0, 0.0, 1.0};
        double PR[9], Jp[18];
        for (int a = 0; a < 3; a++)
          for (int c = 0; c < 3; c++)
            PR[3 * a + c] = (a < d) ? (C.se3 ? -pj[3 * a + c] : pj[3 * a] * C.Rcb[c] + pj[3 * a + 1] * C.Rcb[3 + c] + pj[3 * a + 2] * C.Rcb[6 + c]) : 0.0;
        for (int a = 0; a < 3; a++)
          for (int c = 0; c < 6; c++) Jp[6 * a + c] = PR[3 * a] * Dm[c] + PR[3 * a + 1] * Dm[6 + c] + PR[3 * a + 2] * Dm[12 + c];
        double* Eo = D.E + ((size_t)b * D.maxObs + e) * 18;
        double* Ao = D.Ae + ((size_t)b * D.maxObs + e) * 27;
        int n = 0;
        for (int a = 0; a < 6; a++) {
          for (int c = a; c < 6; c++) {
            double h = 0;
            for (int q2 = 0; q2 < d; q2++) h += Jp[6 * q2 + a] * w * Jp[6 * q2 + c];
            Ao[n++] = h;
          }
          for (int c = 0; c < 3; c++) {
            double h = 0;
            for (int q2 = 0; q2 < d; q2++) h += Jp[6 * q2 + a] * w * Jl[3 * q2 + c];
            Eo[3 * a + c] = h;
          }
        }
        for (int a = 0; a < 6; a++) {
          double tt = 0;
          for (int q2 = 0; q2 < d; q2++) tt += Jp[6 * q2 + a] * (-w * r[q2]);
          Ao[21 + a] = tt;
        }
      }
    }
  }
  double* Ho = D.Hll + ((size_t)b * D.maxPt + j) * 9;
  for (int i = 0; i < 9; i++) Ho[i] = Hll[i];
  double* bo = D.bl + ((size_t)b * D.maxPt + j) * 3;
  bo[0] = bl[0]; bo[1] = bl[1]; bo[2] = bl[2];
}

// warp per optimizable keyframe: H_pp diagonal 6x6 block and b_p[0:6] = fixed-order sum over its edges
__global__ void __launch_bounds__(128) k_lin_kf(BaDev D) {
  const int b = blockIdx.y;
  if (!problem_on(D, b, J_ACTIVE)) return;
  const BaCalib& C = D.calib[b];
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= C.nOpt) return;
  const int s = D.kfStart[(size_t)b * (D.maxKf + 1) + k], t = D.kfStart[(size_t)b * (D.maxKf + 1) + k + 1];
  double acc[27];
#pragma unroll
  for (int i = 0; i < 27; i++) acc[i] = 0;
  for (int q = s + lane; q < t; q += 32) {
    const int e = D.kfEdges[(size_t)b * D.maxObs + q];
    if (!owned(D, D.obsPt[(size_t)b * D.maxObs + e])) continue;
    const double* A = D.Ae + ((size_t)b * D.maxObs + e) * 27;
#pragma unroll
    for (int i = 0; i < 27; i++) acc[i] += A[i];
  }
#pragma unroll
  for (int i = 0; i < 27; i++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[i] += __shfl_down_sync(0xffffffffu, acc[i], o);
  if (lane == 0) {
    double* H = D.Hpp + (size_t)b * D.maxDim * D.maxDim;
    double* bp = D.bp + (size_t)b * D.maxDim;
    const int o = 15 * k, n = C.dimP;
    int idx = 0;
    for (int a = 0; a < 6; a++)
      for (int c = a; c < 6; c++) { H[(size_t)(o + a) * n + o + c] = acc[idx]; H[(size_t)(o + c) * n + o + a] = acc[idx]; idx++; }
    for (int a = 0; a < 6; a++) bp[o + a] = acc[21 + a];
  }
}

// EdgeInertial::linearizeOplus (G2oTypes.cc:524-719): J 9x24, columns [pose1 6, vel1 3, bg1 3, ba1 3, pose2 6, vel2 3]
__device__ void inertial_jacobian(const BaDev& D, int b, int e, double* J, double* r) {
  const int k1 = D.inKf1[(size_t)b * D.maxIn + e], k2 = D.inKf2[(size_t)b * D.maxIn + e];
  inertial_jacobian_core(D.kf + ((size_t)b * D.maxKf + k1) * KF_STRIDE, D.kf + ((size_t)b * D.maxKf + k2) * KF_STRIDE,
                         D.inPre + ((size_t)b * D.maxIn + e) * GFS_BA_PRE_STRIDE, J, r);
}

// CTA per problem.  Phase 1: one warp per inertial edge builds J^T (rho' Omega) J (24x24) and -J^T rho' Omega r.
// Phase 2: the edges are added to the dense pose block one after the other (fixed order), then the
// random-walk edges (BaseMultiEdge / BaseBinaryEdge::constructQuadraticForm).
static const int INERTIAL_THREADS = 256;
__global__ void __launch_bounds__(INERTIAL_THREADS) k_lin_inertial(BaDev D) {
  __shared__ double s_J[INERTIAL_THREADS / 32][216];
  __shared__ double s_WJ[INERTIAL_THREADS / 32][216];
  __shared__ double s_Wr[INERTIAL_THREADS / 32][9];
  const int b = blockIdx.x;
  if (!problem_on(D, b, J_ACTIVE) || D.rank != 0) return;
  const BaCalib& C = D.calib[b];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = INERTIAL_THREADS / 32;
  for (int e0 = 0; e0 < C.nIn; e0 += nw) {
    const int e = e0 + warp;
    if (e < C.nIn) {
      if (lane == 0) {
        double r[9], rho[2];
        inertial_jacobian(D, b, e, s_J[warp], r);
        const double* I = D.infoIn + ((size_t)b * D.maxIn + e) * 81;
        double c2 = 0;
        for (int a = 0; a < 9; a++)
          for (int c = 0; c < 9; c++) c2 += r[a] * I[9 * a + c] * r[c];
        huber(c2, BA_DELTA_INERTIAL, rho);
        for (int a = 0; a < 9; a++) {
          double t = 0;
          for (int c = 0; c < 9; c++) t += I[9 * a + c] * r[c];
          s_Wr[warp][a] = -rho[1] * t;
        }
        s_WJ[warp][0] = rho[1];  // stash; overwritten below after the warp reads it
      }
      __syncwarp();
      const double rho1 = s_WJ[warp][0];
      __syncwarp();
      const double* I = D.infoIn + ((size_t)b * D.maxIn + e) * 81;
      for (int i = lane; i < 216; i += 32) {
        const int a = i / 24, c = i - a * 24;
        double t = 0;
        for (int q = 0; q < 9; q++) t += I[9 * a + q] * s_J[warp][q * 24 + c];
        s_WJ[warp][i] = rho1 * t;
      }
      __syncwarp();
      double* Ho = D.Hin + ((size_t)b * D.maxIn + e) * 600;
      for (int i = lane; i < 576; i += 32) {
        const int c1 = i / 24, c2 = i - c1 * 24;
        double h = 0;
        for (int a = 0; a < 9; a++) h += s_J[warp][a * 24 + c1] * s_WJ[warp][a * 24 + c2];
        Ho[i] = h;
      }
      if (lane < 24) {
        double t = 0;
        for (int a = 0; a < 9; a++) t += s_J[warp][a * 24 + lane] * s_Wr[warp][a];
        Ho[576 + lane] = t;
      }
    }
    __syncwarp();
  }
  __threadfence_block();
  __syncthreads();
  double* H = D.Hpp + (size_t)b * D.maxDim * D.maxDim;
  double* bp = D.bp + (size_t)b * D.maxDim;
  const int n = C.dimP;
  for (int e = 0; e < C.nIn; e++) {
    const int k1 = D.inKf1[(size_t)b * D.maxIn + e], k2 = D.inKf2[(size_t)b * D.maxIn + e];
    const double* Ho = D.Hin + ((size_t)b * D.maxIn + e) * 600;
    for (int i = threadIdx.x; i < 600; i += INERTIAL_THREADS) {
      const int c1 = i < 576 ? i / 24 : i - 576, c2 = i < 576 ? i - c1 * 24 : -1;
      const int g1 = c1 < 15 ? (k1 < C.nOpt ? 15 * k1 + c1 : -1) : (k2 < C.nOpt ? 15 * k2 + (c1 - 15) : -1);
      if (g1 < 0) continue;
      if (c2 < 0) { bp[g1] += Ho[i]; continue; }
      const int g2 = c2 < 15 ? (k1 < C.nOpt ? 15 * k1 + c2 : -1) : (k2 < C.nOpt ? 15 * k2 + (c2 - 15) : -1);
      if (g2 < 0) continue;
      H[(size_t)g1 * n + g2] += Ho[i];
    }
    __syncthreads();
    // EdgeGyroRW / EdgeAccRW (G2oTypes.h:782-852): r = b2 - b1, J = [-I, I], information from C (Optimizer.cc:3384-3396)
    if (threadIdx.x < 18) {
      const int which = threadIdx.x / 9, a = (threadIdx.x % 9) / 3, c = threadIdx.x % 3;
      const double* Om = (which == 0 ? D.infoG : D.infoA) + ((size_t)b * D.maxIn + e) * 9;
      const double* s1 = D.kf + ((size_t)b * D.maxKf + k1) * KF_STRIDE;
      const double* s2 = D.kf + ((size_t)b * D.maxKf + k2) * KF_STRIDE;
      const int fo = which == 0 ? K_BG : K_BA;
      const int o1 = (k1 < C.nOpt) ? 15 * k1 + 9 + 3 * which : -1, o2 = (k2 < C.nOpt) ? 15 * k2 + 9 + 3 * which : -1;
      if (o1 >= 0) H[(size_t)(o1 + a) * n + o1 + c] += Om[3 * a + c];
      if (o2 >= 0) H[(size_t)(o2 + a) * n + o2 + c] += Om[3 * a + c];
      if (o1 >= 0 && o2 >= 0) {
        H[(size_t)(o1 + a) * n + o2 + c] += -Om[3 * a + c];
        H[(size_t)(o2 + a) * n + o1 + c] += -Om[3 * c + a];
      }
      if (c == 0) {
        double rr[3];
        for (int i = 0; i < 3; i++) rr[i] = s2[fo + i] - s1[fo + i];
        const double Or = Om[3 * a] * rr[0] + Om[3 * a + 1] * rr[1] + Om[3 * a + 2] * rr[2];
        if (o1 >= 0) bp[o1 + a] += Or;
        if (o2 >= 0) bp[o2 + a] += -Or;
      }
    }
    __syncthreads();
  }
  // EdgeICP: numeric Jacobians by central differences through the vertices' oplus, delta = 1e-9
  // (BaseBinaryEdge::linearizeOplus, base_binary_edge.hpp:124-190), then the robust quadratic form
  double* sJ = &s_J[0][0];   // [6][12]
  double* sR = &s_Wr[0][0];  // r[6], then w
  for (int e = 0; e < C.nIcp; e++) {
    const int k1 = D.icpKf1[(size_t)b * D.maxIn + e], k2 = D.icpKf2[(size_t)b * D.maxIn + e];
    const double* Rt = D.icpRt + ((size_t)b * D.maxIn + e) * 12;
    const double* s1 = D.kf + ((size_t)b * D.maxKf + k1) * KF_STRIDE;
    const double* s2 = D.kf + ((size_t)b * D.maxKf + k2) * KF_STRIDE;
    if (threadIdx.x < 12) {
      const int v = threadIdx.x / 6, d = threadIdx.x % 6;
      const int k = v == 0 ? k1 : k2;
      double col[6] = {0, 0, 0, 0, 0, 0};
      if (k < C.nOpt) {
        const double delta = 1e-9, scalar = 1.0 / (2 * delta);
        double sp[KF_STRIDE], sm[KF_STRIDE], add[6] = {0, 0, 0, 0, 0, 0}, ep[6], em[6];
        const double* sv = v == 0 ? s1 : s2;
        for (int q = 0; q < KF_STRIDE; q++) { sp[q] = sv[q]; sm[q] = sv[q]; }
        add[d] = delta; kf_oplus(C, sp, add);
        add[d] = -delta; kf_oplus(C, sm, add);
        icp_error(Rt, v == 0 ? sp : s1, v == 0 ? s2 : sp, ep);
        icp_error(Rt, v == 0 ? sm : s1, v == 0 ? s2 : sm, em);
        for (int a = 0; a < 6; a++) col[a] = scalar * (ep[a] - em[a]);
      }
      for (int a = 0; a < 6; a++) sJ[a * 12 + threadIdx.x] = col[a];
    } else if (threadIdx.x == 32) {
      double r[6], rho[2];
      icp_error(Rt, s1, s2, r);
      double c2 = 0;
      for (int a = 0; a < 6; a++) c2 += r[a] * 1e2 * r[a];
      huber(c2, BA_DELTA_ICP, rho);
      for (int a = 0; a < 6; a++) sR[a] = r[a];
      sR[6] = rho[1] * 1e2;
    }
    __syncthreads();
    if (threadIdx.x < 156) {
      const double w = sR[6];
      const int c1 = threadIdx.x < 144 ? threadIdx.x / 12 : threadIdx.x - 144, c2 = threadIdx.x < 144 ? threadIdx.x % 12 : -1;
      const int g1 = c1 < 6 ? (k1 < C.nOpt ? 15 * k1 + c1 : -1) : (k2 < C.nOpt ? 15 * k2 + c1 - 6 : -1);
      if (g1 >= 0) {
        if (c2 < 0) {
          double t = 0;
          for (int a = 0; a < 6; a++) t += sJ[a * 12 + c1] * (-w * sR[a]);
          bp[g1] += t;
        } else {
          const int g2 = c2 < 6 ? (k1 < C.nOpt ? 15 * k1 + c2 : -1) : (k2 < C.nOpt ? 15 * k2 + c2 - 6 : -1);
          if (g2 >= 0) {
            double hh = 0;
            for (int a = 0; a < 6; a++) hh += sJ[a * 12 + c1] * w * sJ[a * 12 + c2];
            H[(size_t)g1 * n + g2] += hh;
          }
        }
      }
    }
    __syncthreads();
  }
}

// ------------------------------------------------------------------------------------------------
// Schur complement + solve (block_solver.hpp:354-486)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) k_schur_prep(BaDev D) {
  const int b = blockIdx.y;
  if (!problem_on(D, b, J_NEED)) return;
  const BaCalib& C = D.calib[b];
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= C.nPt) return;
  const double lambda = D.dstate[b * D_NSTATE + D_LAMBDA];
  double Dm[9];
  const double* H = D.Hll + ((size_t)b * D.maxPt + j) * 9;
  for (int i = 0; i < 9; i++) Dm[i] = H[i];
  Dm[0] += lambda; Dm[4] += lambda; Dm[8] += lambda;
  double* Di = D.Dinv + ((size_t)b * D.maxPt + j) * 9;
  inv3(Dm, Di);
  mv3(Di, D.bl + ((size_t)b * D.maxPt + j) * 3, D.db + ((size_t)b * D.maxPt + j) * 3);
}

// Hs = Hpp (+ lambda on the diagonal of real vertices, identity on unused dofs); bs = bp
__global__ void __launch_bounds__(256) k_hs_init(BaDev D) {
  const int b = blockIdx.y;
  if (!problem_on(D, b, J_NEED)) return;
  const BaCalib& C = D.calib[b];
  const int n = C.dimP;
  const double lambda = D.dstate[b * D_NSTATE + D_LAMBDA];
  const double* H = D.Hpp + (size_t)b * D.maxDim * D.maxDim;
  double* Hs = D.Hs + (size_t)b * D.maxDim * D.maxDim;
  const bool root = D.rank == 0;  // partitioned mode: damping and vertex bookkeeping are added once
  for (int i = blockIdx.x * 256 + threadIdx.x; i < n * n; i += gridDim.x * 256) {
    const int r = i / n, c = i - r * n;
    double v = H[i];
    if (r == c && root) {
      const int k = r / 15, d = r - 15 * k;
      const bool used = d < 6 || D.kfImu[(size_t)b * D.maxKf + k];
      v = used ? v + lambda : 1.0;
    }
    Hs[i] = v;
  }
  if (blockIdx.x == 0)
    for (int i = threadIdx.x; i < n; i += 256) {
      const int k = i / 15, d = i - 15 * k;
      const bool used = d < 6 || D.kfImu[(size_t)b * D.maxKf + k];
      D.bs[(size_t)b * D.maxDim + i] = used ? D.bp[(size_t)b * D.maxDim + i] : 0.0;
    }
}

// warp per ordered keyframe pair (i1 <= i2): S = sum over landmarks seen by both of (B1 Dinv) B2^T
__global__ void __launch_bounds__(128) k_schur_pairs(BaDev D) {
  const int b = blockIdx.y;
  if (!problem_on(D, b, J_NEED)) return;
  const BaCalib& C = D.calib[b];
  const int lane = threadIdx.x & 31;
  const int pairIdx = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int nO = C.nOpt;
  if (pairIdx >= nO * (nO + 1) / 2) return;
  int i1 = 0, rem = pairIdx;
  while (rem >= nO - i1) { rem -= nO - i1; i1++; }
  const int i2 = i1 + rem;
  const int s = D.kfStart[(size_t)b * (D.maxKf + 1) + i1], t = D.kfStart[(size_t)b * (D.maxKf + 1) + i1 + 1];
  double S[36];
#pragma unroll
  for (int i = 0; i < 36; i++) S[i] = 0;
  for (int q = s + lane; q < t; q += 32) {
    const int e1 = D.kfEdges[(size_t)b * D.maxObs + q];
    const int j = D.obsPt[(size_t)b * D.maxObs + e1];
    if (!owned(D, j)) continue;
    const int e2 = D.ptKfEdge[((size_t)b * D.maxPt + j) * D.maxOpt + i2];
    if (e2 < 0) continue;
    const double* B1 = D.E + ((size_t)b * D.maxObs + e1) * 18;
    const double* B2 = D.E + ((size_t)b * D.maxObs + e2) * 18;
    const double* Di = D.Dinv + ((size_t)b * D.maxPt + j) * 9;
    double BD[18];
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
      for (int c = 0; c < 3; c++) BD[3 * a + c] = B1[3 * a] * Di[c] + B1[3 * a + 1] * Di[3 + c] + B1[3 * a + 2] * Di[6 + c];
#pragma unroll
    for (int a = 0; a < 6; a++)
#pragma unroll
      for (int c = 0; c < 6; c++) S[6 * a + c] += BD[3 * a] * B2[3 * c] + BD[3 * a + 1] * B2[3 * c + 1] + BD[3 * a + 2] * B2[3 * c + 2];
  }
#pragma unroll
  for (int i = 0; i < 36; i++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) S[i] += __shfl_down_sync(0xffffffffu, S[i], o);
  if (lane == 0) {
    double* Hs = D.Hs + (size_t)b * D.maxDim * D.maxDim;
    const int n = C.dimP, o1 = 15 * i1, o2 = 15 * i2;
    for (int a = 0; a < 6; a++)
      for (int c = 0; c < 6; c++) {
        if (i1 == i2) {
          Hs[(size_t)(o1 + a) * n + o1 + c] -= S[6 * a + c];
        } else {
          Hs[(size_t)(o1 + a) * n + o2 + c] -= S[6 * a + c];
          Hs[(size_t)(o2 + c) * n + o1 + a] -= S[6 * a + c];
        }
      }
  }
}

// warp per keyframe: bs[pose] -= sum over its edges of B (Dinv bl)
__global__ void __launch_bounds__(128) k_schur_rhs(BaDev D) {
  const int b = blockIdx.y;
  if (!problem_on(D, b, J_NEED)) return;
  const BaCalib& C = D.calib[b];
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (k >= C.nOpt) return;
  const int s = D.kfStart[(size_t)b * (D.maxKf + 1) + k], t = D.kfStart[(size_t)b * (D.maxKf + 1) + k + 1];
  double acc[6] = {0, 0, 0, 0, 0, 0};
  for (int q = s + lane; q < t; q += 32) {
    const int e = D.kfEdges[(size_t)b * D.maxObs + q];
    const int j = D.obsPt[(size_t)b * D.maxObs + e];
    if (!owned(D, j)) continue;
    const double* B = D.E + ((size_t)b * D.maxObs + e) * 18;
    const double* d = D.db + ((size_t)b * D.maxPt + j) * 3;
#pragma unroll
    for (int a = 0; a < 6; a++) acc[a] += B[3 * a] * d[0] + B[3 * a + 1] * d[1] + B[3 * a + 2] * d[2];
  }
#pragma unroll
  for (int a = 0; a < 6; a++)
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) acc[a] += __shfl_down_sync(0xffffffffu, acc[a], o);
  if (lane == 0)
    for (int a = 0; a < 6; a++) D.bs[(size_t)b * D.maxDim + 15 * k + a] -= acc[a];
}

// CTA per problem: blocked right-looking LDL^T (no pivoting) of the reduced pose system, then the
// solve.  Block column of 32: (1) one warp factors the 32x32 diagonal block in shared memory, (2) every
// thread solves one row of the panel below it, (3) the trailing lower triangle gets its rank-32 update
// from the panel held in shared memory (8x8 register tiles), so the matrix in L2 is read-modify-written
// once per 32 columns.  g2o solves this system with Eigen::SimplicialLDLT (linear_solver_eigen.h:60-76).
static const int LDLT_THREADS = 512, LDLT_NB = 32, LDLT_PITCH = 33, LDLT_MAXN = 320;
__global__ void __launch_bounds__(LDLT_THREADS) k_ldlt_solve(BaDev D) {
  extern __shared__ double s_ld[];
  double* Lp = s_ld;                              // [LDLT_MAXN][33] panel: unit-lower L columns of this block
  double* Wp = Lp + LDLT_MAXN * LDLT_PITCH;       // [LDLT_MAXN][33] panel times D
  double* dv = Wp + LDLT_MAXN * LDLT_PITCH;       // [32] pivots of this block
  double* y = dv + 32;                            // [LDLT_MAXN] right-hand side / solution
  __shared__ int s_ok;
  const int b = blockIdx.x;
  if (!problem_on(D, b, J_NEED)) return;
  const BaCalib& C = D.calib[b];
  const int n = C.dimP, tid = threadIdx.x, lane = tid & 31;
  double* A = D.Hs + (size_t)b * D.maxDim * D.maxDim;
  double* x = D.x + (size_t)b * (D.maxDim + 3 * D.maxPt);
  if (tid == 0) s_ok = 1;
  for (int i = tid; i < n; i += LDLT_THREADS) y[i] = D.bs[(size_t)b * D.maxDim + i];
  __syncthreads();
  for (int kb = 0; kb < n; kb += LDLT_NB) {
    const int w = min(LDLT_NB, n - kb);
    const int rows = n - kb;  // panel rows kb .. n-1, local index r = i - kb
    for (int idx = tid; idx < rows * LDLT_NB; idx += LDLT_THREADS) {
      const int r = idx >> 5, c = idx & 31;
      Lp[r * LDLT_PITCH + c] = (c < w) ? A[(size_t)(kb + r) * n + kb + c] : 0.0;
    }
    __syncthreads();
    // (1) diagonal block: lane r owns row r (r < w)
    if (tid < 32) {
      for (int j = 0; j < w; j++) {
        const double d = Lp[j * LDLT_PITCH + j];
        if (d == 0.0 || !isfinite(d)) { if (lane == 0) s_ok = 0; break; }
        if (lane == 0) dv[j] = d;
        if (lane > j && lane < w) {
          const double l = Lp[lane * LDLT_PITCH + j] / d;
          for (int c = j + 1; c <= lane; c++) Lp[lane * LDLT_PITCH + c] -= l * Lp[c * LDLT_PITCH + j];  // column j still unscaled
        }
        __syncwarp();
        if (lane > j && lane < w) Lp[lane * LDLT_PITCH + j] /= d;
        __syncwarp();
      }
    }
    __syncthreads();
    if (!s_ok) break;
    // (2) panel rows below the diagonal block: L[r][c] = (A[r][c] - sum_{j<c} L[r][j] d_j L[c][j]) / d_c
    for (int r = w + tid; r < rows; r += LDLT_THREADS) {
      double* row = Lp + r * LDLT_PITCH;
      for (int c = 0; c < w; c++) {
        double v = row[c];
        for (int j = 0; j < c; j++) v -= row[j] * dv[j] * Lp[c * LDLT_PITCH + j];
        row[c] = v / dv[c];
      }
    }
    __syncthreads();
    // forward substitution with this block column: unit-lower solve on the diagonal block (warp 0),
    // then y[rows below] -= L_panel * y_block
    if (tid < 32) {
      for (int j = 0; j < w; j++) {
        const double yj = y[kb + j];
        if (lane > j && lane < w) y[kb + lane] -= Lp[lane * LDLT_PITCH + j] * yj;
        __syncwarp();
      }
    }
    __syncthreads();
    for (int r = w + tid; r < rows; r += LDLT_THREADS) {
      double v = 0;
      for (int j = 0; j < w; j++) v += Lp[r * LDLT_PITCH + j] * y[kb + j];
      y[kb + r] -= v;
    }
    // write the factor columns back (unit diagonal implied, pivots on the diagonal) and build W = L D
    for (int idx = tid; idx < rows * LDLT_NB; idx += LDLT_THREADS) {
      const int r = idx >> 5, c = idx & 31;
      if (c < w) {
        const double l = Lp[r * LDLT_PITCH + c];
        if (r > c) A[(size_t)(kb + r) * n + kb + c] = l;
        else if (r == c) A[(size_t)(kb + r) * n + kb + c] = dv[c];
        Wp[r * LDLT_PITCH + c] = (r >= w) ? l * dv[c] : 0.0;
      } else {
        Wp[r * LDLT_PITCH + c] = 0.0;
      }
    }
    __syncthreads();
    // (3) trailing update on the lower triangle, rows/cols local index >= w, 8x8 tiles
    const int m = rows - w;
    if (m > 0) {
      const int nt = (m + 7) >> 3;
      const int ntiles = nt * (nt + 1) / 2;
      for (int tile = tid; tile < ntiles; tile += LDLT_THREADS) {
        int ti = (int)((sqrt(8.0 * tile + 1.0) - 1.0) * 0.5);
        while ((ti + 1) * (ti + 2) / 2 <= tile) ti++;
        while (ti * (ti + 1) / 2 > tile) ti--;
        const int tk = tile - ti * (ti + 1) / 2;
        const int r0 = w + ti * 8, k0 = w + tk * 8;
        double acc[8][8];
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
          for (int v = 0; v < 8; v++) acc[u][v] = 0.0;
        for (int j = 0; j < w; j++) {
          double li[8], wk[8];
#pragma unroll
          for (int u = 0; u < 8; u++) li[u] = (r0 + u < rows) ? Lp[(r0 + u) * LDLT_PITCH + j] : 0.0;
#pragma unroll
          for (int v = 0; v < 8; v++) wk[v] = (k0 + v < rows) ? Wp[(k0 + v) * LDLT_PITCH + j] : 0.0;
#pragma unroll
          for (int u = 0; u < 8; u++)
#pragma unroll
            for (int v = 0; v < 8; v++) acc[u][v] += li[u] * wk[v];
        }
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
          for (int v = 0; v < 8; v++) {
            const int r = r0 + u, k = k0 + v;
            if (r < rows && k <= r) A[(size_t)(kb + r) * n + kb + k] -= acc[u][v];
          }
      }
    }
    __syncthreads();
  }
  __syncthreads();
  const int ok = s_ok;
  if (tid == 0) D.istate[b * J_NSTATE + J_OK] = ok;
  if (!ok) {
    for (int i = tid; i < n; i += LDLT_THREADS) x[i] = 0.0;
    return;
  }
  // diagonal scaling, then blocked backward substitution with L^T: block columns from last to first,
  // each panel re-read (coalesced) from the factor in L2
  for (int i = tid; i < n; i += LDLT_THREADS) y[i] /= A[(size_t)i * n + i];
  __syncthreads();
  const int nblocks = (n + LDLT_NB - 1) / LDLT_NB;
  for (int bi = nblocks - 1; bi >= 0; bi--) {
    const int kb = bi * LDLT_NB;
    const int w = min(LDLT_NB, n - kb), rows = n - kb;
    for (int idx = tid; idx < rows * LDLT_NB; idx += LDLT_THREADS) {
      const int r = idx >> 5, c = idx & 31;
      Lp[r * LDLT_PITCH + c] = (c < w && r > c) ? A[(size_t)(kb + r) * n + kb + c] : 0.0;
    }
    __syncthreads();
    // x[kb + c] -= sum_{r >= w} L[r][c] * x[kb + r]: warp per column group, lanes over rows
    {
      const int warp = tid >> 5;
      for (int c = warp; c < w; c += LDLT_THREADS / 32) {
        double v = 0;
        for (int r = w + lane; r < rows; r += 32) v += Lp[r * LDLT_PITCH + c] * y[kb + r];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
        if (lane == 0) y[kb + c] -= v;
      }
    }
    __syncthreads();
    // unit-upper (L^T) solve inside the diagonal block, last column first
    if (tid < 32) {
      for (int j = w - 1; j >= 0; j--) {
        const double xj = y[kb + j];
        if (lane < j) y[kb + lane] -= Lp[j * LDLT_PITCH + lane] * xj;
        __syncwarp();
      }
    }
    __syncthreads();
  }
  __syncthreads();
  for (int i = tid; i < n; i += LDLT_THREADS) {
    const int k = i / 15, d = i - 15 * k;
    const bool used = d < 6 || D.kfImu[(size_t)b * D.maxKf + k];
    x[i] = used ? y[i] : 0.0;
  }
}

// ImuCamPose::Update (G2oTypes.cc:191-217) and the additive vertices; state <- backup (+) x
__global__ void __launch_bounds__(128) k_backsub_update(BaDev D) {
  const int b = blockIdx.y;
  if (!problem_on(D, b, J_NEED)) return;
  const BaCalib& C = D.calib[b];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int ok = D.istate[b * J_NSTATE + J_OK];
  double* x = D.x + (size_t)b * (D.maxDim + 3 * D.maxPt);
  const double lambda = D.dstate[b * D_NSTATE + D_LAMBDA];
  double sc = 0;
  if (i < C.nPt) {
    const int j = i;
    double* Xo = D.pt + ((size_t)b * D.maxPt + j) * 3;
    const double* Xb = D.ptBak + ((size_t)b * D.maxPt + j) * 3;
    double xl[3] = {0, 0, 0};
    if (ok && owned(D, j)) {
      const double* bl = D.bl + ((size_t)b * D.maxPt + j) * 3;
      double c[3] = {bl[0], bl[1], bl[2]};
      const int s = D.ptStart[(size_t)b * (D.maxPt + 1) + j], t = D.ptStart[(size_t)b * (D.maxPt + 1) + j + 1];
      for (int q = s; q < t; q++) {
        const int e = D.ptEdges[(size_t)b * D.maxObs + q];
        const int k = D.obsKf[(size_t)b * D.maxObs + e];
        if (k >= C.nOpt) continue;
        const double* B = D.E + ((size_t)b * D.maxObs + e) * 18;
        const double* xp = x + 15 * k;
        for (int cc = 0; cc < 3; cc++)
          for (int a = 0; a < 6; a++) c[cc] -= B[3 * a + cc] * xp[a];
      }
      mv3(D.Dinv + ((size_t)b * D.maxPt + j) * 9, c, xl);
      for (int a = 0; a < 3; a++) sc += xl[a] * (lambda * xl[a] + bl[a]);
    }
    for (int a = 0; a < 3; a++) { x[C.dimP + 3 * j + a] = xl[a]; Xo[a] = Xb[a] + xl[a]; }
  } else if (i - C.nPt < C.nOpt) {
    const int k = i - C.nPt;
    const double* u = x + 15 * k;
    const double* sb = D.kfBak + ((size_t)b * D.maxKf + k) * KF_STRIDE;
    double* so = D.kf + ((size_t)b * D.maxKf + k) * KF_STRIDE;
    double st[KF_STRIDE];
    for (int q = 0; q < KF_STRIDE; q++) st[q] = sb[q];
    if (ok) {
      kf_oplus(C, st, u);
      if (D.kfImu[(size_t)b * D.maxKf + k])
        for (int a = 0; a < 3; a++) { st[K_VEL + a] += u[6 + a]; st[K_BG + a] += u[9 + a]; st[K_BA + a] += u[12 + a]; }
      if (D.rank == 0) {  // computeScale over the pose part, counted once
        const double* bp = D.bp + (size_t)b * D.maxDim + 15 * k;
        for (int a = 0; a < 15; a++) sc += u[a] * (lambda * u[a] + bp[a]);
      }
    }
    for (int q = 0; q < KF_STRIDE; q++) so[q] = st[q];
  }
  block_sum_store<128>(sc, D.partScale + (size_t)b * D.nblk + blockIdx.x);
}

__global__ void k_backup(BaDev D, int flag) {  // push(): backup <- state
  const int b = blockIdx.y;
  if (!problem_on(D, b, flag)) return;
  const BaCalib& C = D.calib[b];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C.nKf * KF_STRIDE) D.kfBak[(size_t)b * D.maxKf * KF_STRIDE + i] = D.kf[(size_t)b * D.maxKf * KF_STRIDE + i];
  if (i < C.nPt * 3) D.ptBak[(size_t)b * D.maxPt * 3 + i] = D.pt[(size_t)b * D.maxPt * 3 + i];
}
__global__ void k_restore(BaDev D) {  // pop() for the problems whose last trial was rejected (phase flag 2)
  const int b = blockIdx.y;
  if (D.istate[b * J_NSTATE + J_PHASE] != 2) return;
  const BaCalib& C = D.calib[b];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < C.nKf * KF_STRIDE) D.kf[(size_t)b * D.maxKf * KF_STRIDE + i] = D.kfBak[(size_t)b * D.maxKf * KF_STRIDE + i];
  if (i < C.nPt * 3) D.pt[(size_t)b * D.maxPt * 3 + i] = D.ptBak[(size_t)b * D.maxPt * 3 + i];
}

// ------------------------------------------------------------------------------------------------
// Levenberg control (optimization_algorithm_levenberg.cpp:59-164, sparse_optimizer.cpp:354-420)
// ------------------------------------------------------------------------------------------------
__device__ double total_chi(const BaDev& D, int b, const BaCalib& C) {
  double s = 0;
  for (int e = 0; e < C.nIn; e++) s += D.inRho[(size_t)b * D.maxIn + e];
  for (int e = 0; e < C.nIcp; e++) s += D.icpRho[(size_t)b * D.maxIn + e];
  const int nb = (C.nObs + ERR_THREADS - 1) / ERR_THREADS;
  for (int i = 0; i < nb; i++) s += D.partChi[(size_t)b * D.nblk + i];
  return s;
}
// mode 0: initial error (err); mode 1: start of an outer iteration; mode 2: after a trial
__global__ void k_lm_control(BaDev D, int batch, int mode, int it) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  const BaCalib& C = D.calib[b];
  double* ds = D.dstate + b * D_NSTATE;
  int* is = D.istate + b * J_NSTATE;
  if (mode == 0) {
    const double chi = total_chi(D, b, C);
    ds[D_ERR0] = chi; ds[D_LAST] = chi;
    return;
  }
  if (mode == 1) {
    if (!is[J_ACTIVE]) return;
    const double chi = total_chi(D, b, C);
    ds[D_CUR] = chi; ds[D_INI] = chi; ds[D_TEMP] = chi; ds[D_LAST] = chi;
    if (it == 0) { ds[D_LAMBDA] = C.lambda_init; ds[D_NI] = 2; is[J_NBAD] = 0; }
    ds[D_RHO] = 0;
    is[J_QMAX] = 0;
    is[J_NEED] = 1;
    is[J_IT] = it;
    is[J_PHASE] = 0;
    atomicAdd(&D.counters[0], 1);
    return;
  }
  if (!is[J_NEED]) { is[J_PHASE] = 0; return; }
  const bool ok2 = is[J_OK] != 0;
  double tempChi = total_chi(D, b, C);
  ds[D_LAST] = tempChi;
  if (!ok2) tempChi = DBL_MAX;
  double rho = ds[D_CUR] - tempChi;
  double scale = 0;
  const int nb = (C.nPt + C.nOpt + 127) / 128;
  for (int i = 0; i < nb; i++) scale += D.partScale[(size_t)b * D.nblk + i];
  scale += 1e-3;
  rho /= scale;
  if (rho > 0 && isfinite(tempChi)) {
    double alpha = 1. - pow((2 * rho - 1), 3);
    alpha = fmin(alpha, 2. / 3.);
    const double sf = fmax(1. / 3., alpha);
    ds[D_LAMBDA] *= sf;
    ds[D_NI] = 2;
    ds[D_CUR] = tempChi;
    is[J_PHASE] = 1;  // accepted: keep the trial state
  } else {
    ds[D_LAMBDA] *= ds[D_NI];
    ds[D_NI] *= 2;
    is[J_PHASE] = 2;  // rejected: k_restore pops the backup
  }
  ds[D_RHO] = rho;
  is[J_QMAX]++;
  is[J_TRIALS]++;
  if (rho < 0 && is[J_QMAX] < 10) {
    atomicAdd(&D.counters[0], 1);  // another trial with the larger lambda
    return;
  }
  // the outer iteration is over
  is[J_NEED] = 0;
  is[J_DONE]++;
  bool stop = false;
  if (is[J_QMAX] == 10 || rho == 0) stop = true;
  else {
    if ((ds[D_INI] - ds[D_CUR]) * 1e3 < ds[D_INI]) is[J_NBAD]++;
    else is[J_NBAD] = 0;
    if (is[J_NBAD] >= 3) stop = true;
  }
  if (stop || it + 1 >= C.iterations) is[J_ACTIVE] = 0;
  else atomicAdd(&D.counters[1], 1);
}

// computeLambdaInit (optimization_algorithm_levenberg.cpp:166-178) for the problems without a user lambda
// (lambda_init <= 0): tau * max |H_jj| over the pose blocks and the landmark blocks, tau = 1e-5.  CTA per problem.
__global__ void __launch_bounds__(256) k_lambda_auto(BaDev D) {
  __shared__ double s_max[256];
  const int b = blockIdx.x, tid = threadIdx.x;
  const BaCalib& C = D.calib[b];
  if (C.lambda_init > 0 || !problem_on(D, b, J_ACTIVE)) return;
  double m = 0;
  const double* H = D.Hpp + (size_t)b * D.maxDim * D.maxDim;
  for (int j = tid; j < C.dimP; j += 256) m = fmax(m, fabs(H[(size_t)j * C.dimP + j]));
  for (int j = tid; j < C.nPt * 3; j += 256) m = fmax(m, fabs(D.Hll[((size_t)b * D.maxPt + j / 3) * 9 + 4 * (j % 3)]));
  s_max[tid] = m;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (tid < o) s_max[tid] = fmax(s_max[tid], s_max[tid + o]);
    __syncthreads();
  }
  if (tid == 0) D.dstate[b * D_NSTATE + D_LAMBDA] = 1e-5 * s_max[0];
}

__global__ void k_ba_init(BaDev D, int batch) {
  const int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= batch) return;
  for (int i = 0; i < D_NSTATE; i++) D.dstate[b * D_NSTATE + i] = 0;
  for (int i = 0; i < J_NSTATE; i++) D.istate[b * J_NSTATE + i] = 0;
  D.istate[b * J_NSTATE + J_ACTIVE] = D.calib[b].iterations > 0 ? 1 : 0;
  D.istate[b * J_NSTATE + J_OK] = 1;
}

struct BaOutDev {
  uint8_t *depthPos, *outlier;  // [B][maxObs]
  float* errs;                  // [B][2]
  int* info;                    // [B][4] failed, iterations_done, lm_trials
  double* lambda;               // [B]
};
__global__ void __launch_bounds__(128) k_ba_finish(BaDev D, BaOutDev O) {
  const int b = blockIdx.y;
  const BaCalib& C = D.calib[b];
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) {
    const float err = (float)D.dstate[b * D_NSTATE + D_ERR0], err_end = (float)D.dstate[b * D_NSTATE + D_LAST];
    O.errs[2 * b] = err; O.errs[2 * b + 1] = err_end;
    O.info[4 * b] = (!C.se3 && (2 * err < err_end || isnan(err) || isnan(err_end)) && !C.bLarge) ? 1 : 0;
    O.info[4 * b + 1] = D.istate[b * J_NSTATE + J_DONE];
    O.info[4 * b + 2] = D.istate[b * J_NSTATE + J_TRIALS];
    O.lambda[b] = D.dstate[b * D_NSTATE + D_LAMBDA];
  }
  if (e >= C.nObs) return;
  const int k = D.obsKf[(size_t)b * D.maxObs + e], j = D.obsPt[(size_t)b * D.maxObs + e];
  const bool mono = D.obsUvr[((size_t)b * D.maxObs + e) * 3 + 2] < 0;
  const double c2 = D.chi2[(size_t)b * D.maxObs + e];
  const double* kf = D.kf + ((size_t)b * D.maxKf + k) * KF_STRIDE;
  const double* X = D.pt + ((size_t)b * D.maxPt + j) * 3;
  const double z = kf[K_RCW + 6] * X[0] + kf[K_RCW + 7] * X[1] + kf[K_RCW + 8] * X[2] + kf[K_TCW + 2];
  const bool dpos = (mono || C.se3) ? (z > 0.0) : true;
  const float chi2Mono2 = 5.991f, chi2Stereo2 = 7.815f;  // Optimizer.cc:3428, 3430
  bool out;
  if (C.se3) {  // LocalBundleAdjustment (Optimizer.cc:1972, 1996)
    out = c2 > (mono ? 5.991 : 7.815) || !dpos;
  } else if (mono) {
    const bool close = D.ptClose[(size_t)b * D.maxPt + j] != 0;
    out = (c2 > chi2Mono2 && !close) || (c2 > 1.5f * chi2Mono2 && close) || !dpos;
  } else {
    out = c2 > chi2Stereo2;
  }
  O.depthPos[(size_t)b * D.maxObs + e] = dpos;
  O.outlier[(size_t)b * D.maxObs + e] = out;
}

}  // namespace gfs

// ================================================================================================
// Host side: flattening, inertial information matrices (EdgeInertial ctor), launch sequence
// ================================================================================================
using namespace gfs;

struct GfsBa {
  BaDev dev;
  BaOutDev out;
  int maxBatch = 0, batch = 0;
  std::vector<BaCalib> calib;
  std::vector<DevBuf> bufs;
  // host staging of the flattened batch
  std::vector<double> h_kf, h_pt, h_uvr, h_infoIn, h_infoG, h_infoA;
  std::vector<uint8_t> h_kfImu, h_ptClose;
  std::vector<int> h_obsKf, h_obsPt, h_inKf1, h_inKf2, h_ptStart, h_ptEdges, h_kfStart, h_kfEdges, h_ptKfEdge, h_icpKf1, h_icpKf2;
  std::vector<double> h_icpRt;
  std::vector<float> h_obsW, h_inPre;
  PinnedBuf h_counters;
  int launches = 0;
  GfsAllReduceFn allreduce = nullptr;
  void* allreduceUser = nullptr;
  void* ncclComm = nullptr;  // ncclComm_t of the partition (gfs_ba_set_partition_nccl): collectives go straight onto the solve stream
  int ncclCalls = 0;         // grouped NCCL launches of the last solve
  DevBuf redBuf;
};

// ---- NCCL, bound at run time (dlopen): the library must load on machines without NCCL (single-GPU use, the CPU-only build
// check), so nothing is linked.  Prototypes as in <nccl.h> (2.x ABI: ncclUniqueId is 128 bytes, passed by value).
namespace ncclrt {
struct UniqueId { char internal[128]; };
typedef int (*GetUniqueIdFn)(UniqueId*);
typedef int (*CommInitRankFn)(void**, int, UniqueId, int);
typedef int (*CommDestroyFn)(void*);
typedef int (*AllReduceFn)(const void*, void*, size_t, int, int, void*, cudaStream_t);
typedef int (*GroupFn)();
typedef const char* (*ErrStrFn)(int);
static const int kDouble = 8, kSum = 0;  // ncclFloat64, ncclSum
struct Api {
  void* lib = nullptr;
  GetUniqueIdFn getUniqueId = nullptr;
  CommInitRankFn commInitRank = nullptr;
  CommDestroyFn commDestroy = nullptr;
  AllReduceFn allReduce = nullptr;
  GroupFn groupStart = nullptr, groupEnd = nullptr;
  ErrStrFn errStr = nullptr;
};
static Api* api() {
  static Api a;
  static std::once_flag once;
  std::call_once(once, [] {
    // the copy the process already has (torch's bundled NCCL), else the system one
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!lib) lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!lib) lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (lib) {
      a.lib = lib;
      a.getUniqueId = (GetUniqueIdFn)dlsym(lib, "ncclGetUniqueId");
      a.commInitRank = (CommInitRankFn)dlsym(lib, "ncclCommInitRank");
      a.commDestroy = (CommDestroyFn)dlsym(lib, "ncclCommDestroy");
      a.allReduce = (AllReduceFn)dlsym(lib, "ncclAllReduce");
      a.groupStart = (GroupFn)dlsym(lib, "ncclGroupStart");
      a.groupEnd = (GroupFn)dlsym(lib, "ncclGroupEnd");
      a.errStr = (ErrStrFn)dlsym(lib, "ncclGetErrorString");
    }
  });
  return (a.lib && a.getUniqueId && a.commInitRank && a.commDestroy && a.allReduce && a.groupStart && a.groupEnd) ? &a : nullptr;
}
}  // namespace ncclrt

template <class T>
static int dev_alloc(GfsBa* h, T** p, size_t count) {
  h->bufs.emplace_back();
  int rc = h->bufs.back().reserve(std::max<size_t>(count, 1) * sizeof(T));
  if (rc) return rc;
  *p = (T*)h->bufs.back().p;
  return GFS_OK;
}

extern "C" {

int gfs_ba_create(int max_kf, int max_points, int max_obs, int max_inertial, int max_batch, GfsBa** out) {
  GFS_REQUIRE(out, GFS_ERR_INVALID, "out is null");
  *out = nullptr;
  GFS_REQUIRE(max_kf > 0 && max_points > 0 && max_obs > 0 && max_inertial >= 0 && max_batch > 0, GFS_ERR_INVALID, "bad capacity");
  // the reference optimises at most 20 keyframes (maxOpt, Optimizer.cc:3062-3068) next to up to 200 fixed ones
  int rc = gfs_device_check();
  if (rc) return rc;
  GfsBa* h = new GfsBa();
  h->bufs.reserve(64);
  BaDev& D = h->dev;
  memset(&D, 0, sizeof(D));
  D.maxKf = max_kf; D.maxPt = max_points; D.maxObs = max_obs; D.maxIn = std::max(max_inertial, 1);
  D.maxOpt = std::min(max_kf, 21);
  D.maxDim = 15 * D.maxOpt;
  D.nblk = std::max(div_up(max_obs, ERR_THREADS), div_up(max_points + max_kf, 128)) + 1;
  D.rank = 0; D.world = 1;
  h->maxBatch = max_batch;
  const size_t B = max_batch;
#define AL(field, type, count)                                  \
  if ((rc = dev_alloc<type>(h, (type**)&D.field, (count)))) {   \
    gfs_ba_destroy(h);                                          \
    return rc;                                                  \
  }
  AL(calib, BaCalib, B)
  AL(kf, double, B * D.maxKf * KF_STRIDE) AL(kfBak, double, B * D.maxKf * KF_STRIDE)
  AL(pt, double, B * D.maxPt * 3) AL(ptBak, double, B * D.maxPt * 3)
  AL(kfImu, uint8_t, B * D.maxKf) AL(ptClose, uint8_t, B * D.maxPt)
  AL(obsKf, int, B * D.maxObs) AL(obsPt, int, B * D.maxObs) AL(obsUvr, double, B * D.maxObs * 3) AL(obsW, float, B * D.maxObs)
  AL(inKf1, int, B * D.maxIn) AL(inKf2, int, B * D.maxIn) AL(inPre, float, B * D.maxIn * GFS_BA_PRE_STRIDE)
  AL(infoIn, double, B * D.maxIn * 81) AL(infoG, double, B * D.maxIn * 9) AL(infoA, double, B * D.maxIn * 9)
  AL(icpKf1, int, B * D.maxIn) AL(icpKf2, int, B * D.maxIn) AL(icpRt, double, B * D.maxIn * 12) AL(icpRho, double, B * D.maxIn)
  AL(ptStart, int, B * (D.maxPt + 1)) AL(ptEdges, int, B * D.maxObs)
  AL(kfStart, int, B * (D.maxKf + 1)) AL(kfEdges, int, B * D.maxObs)
  AL(ptKfEdge, int, B * D.maxPt * D.maxOpt)
  AL(chi2, double, B * D.maxObs) AL(inRho, double, B * D.maxIn) AL(partChi, double, B * D.nblk)
  AL(E, double, B * D.maxObs * 18) AL(Ae, double, B * D.maxObs * 27)
  AL(Hll, double, B * D.maxPt * 9) AL(bl, double, B * D.maxPt * 3) AL(Dinv, double, B * D.maxPt * 9) AL(db, double, B * D.maxPt * 3)
  AL(Hpp, double, B * D.maxDim * D.maxDim) AL(bp, double, B * D.maxDim)
  AL(Hs, double, B * D.maxDim * D.maxDim) AL(bs, double, B * D.maxDim)
  AL(Hin, double, B * D.maxIn * 600)
  AL(x, double, B * (D.maxDim + 3 * D.maxPt)) AL(partScale, double, B * D.nblk)
  AL(dstate, double, B * D_NSTATE) AL(istate, int, B * J_NSTATE) AL(counters, int, 4)
  BaOutDev& O = h->out;
  if ((rc = dev_alloc<uint8_t>(h, &O.depthPos, B * D.maxObs)) || (rc = dev_alloc<uint8_t>(h, &O.outlier, B * D.maxObs)) ||
      (rc = dev_alloc<float>(h, &O.errs, B * 2)) || (rc = dev_alloc<int>(h, &O.info, B * 4)) ||
      (rc = dev_alloc<double>(h, &O.lambda, B))) {
    gfs_ba_destroy(h);
    return rc;
  }
#undef AL
  if ((rc = h->h_counters.reserve(16))) { gfs_ba_destroy(h); return rc; }
  *out = h;
  return GFS_OK;
}

int gfs_ba_destroy(GfsBa* h) {
  if (!h) return GFS_OK;
  for (DevBuf& b : h->bufs) b.release();
  h->redBuf.release();
  h->h_counters.release();
  if (h->ncclComm) {
    if (ncclrt::Api* n = ncclrt::api()) n->commDestroy(h->ncclComm);
  }
  delete h;
  return GFS_OK;
}

int gfs_ba_last_launches(const GfsBa* h) { return h ? h->launches : GFS_ERR_INVALID; }

int gfs_ba_set_partition(GfsBa* h, int rank, int world, GfsAllReduceFn allreduce, void* user) {
  GFS_REQUIRE(h, GFS_ERR_INVALID, "null handle");
  GFS_REQUIRE(world >= 1 && rank >= 0 && rank < world, GFS_ERR_INVALID, "bad rank/world");
  GFS_REQUIRE(world == 1 || allreduce, GFS_ERR_INVALID, "partitioned mode needs an all-reduce callback");
  if (h->ncclComm) {
    if (ncclrt::Api* n = ncclrt::api()) n->commDestroy(h->ncclComm);
    h->ncclComm = nullptr;
  }
  h->dev.rank = rank;
  h->dev.world = world;
  h->allreduce = allreduce;
  h->allreduceUser = user;
  return GFS_OK;
}

int gfs_nccl_unique_id(void* out128) {
  GFS_REQUIRE(out128, GFS_ERR_INVALID, "null pointer");
  ncclrt::Api* n = ncclrt::api();
  GFS_REQUIRE(n, GFS_ERR_INVALID, "libnccl.so.2 not found (dlopen)");
  ncclrt::UniqueId id;
  const int rc = n->getUniqueId(&id);
  if (rc != 0) { gfs::set_error("ncclGetUniqueId -> %s", n->errStr ? n->errStr(rc) : "error"); return GFS_ERR_CUDA; }
  memcpy(out128, &id, sizeof(id));
  return GFS_OK;
}

int gfs_ba_set_partition_nccl(GfsBa* h, int rank, int world, const void* unique_id128) {
  GFS_REQUIRE(h, GFS_ERR_INVALID, "null handle");
  GFS_REQUIRE(world >= 1 && rank >= 0 && rank < world, GFS_ERR_INVALID, "bad rank/world");
  ncclrt::Api* n = ncclrt::api();
  if (h->ncclComm && n) { n->commDestroy(h->ncclComm); h->ncclComm = nullptr; }
  h->dev.rank = rank;
  h->dev.world = world;
  h->allreduce = nullptr;
  h->allreduceUser = nullptr;
  if (world == 1) return GFS_OK;
  GFS_REQUIRE(unique_id128, GFS_ERR_INVALID, "null unique id");
  GFS_REQUIRE(n, GFS_ERR_INVALID, "libnccl.so.2 not found (dlopen)");
  ncclrt::UniqueId id;
  memcpy(&id, unique_id128, sizeof(id));
  const int rc = n->commInitRank(&h->ncclComm, world, id, rank);
  if (rc != 0) {
    h->ncclComm = nullptr; h->dev.rank = 0; h->dev.world = 1;
    gfs::set_error("ncclCommInitRank -> %s", n->errStr ? n->errStr(rc) : "error");
    return GFS_ERR_CUDA;
  }
  return GFS_OK;
}

int gfs_ba_last_nccl_calls(const GfsBa* h) { return h ? h->ncclCalls : GFS_ERR_INVALID; }

int gfs_ba_upload(GfsBa* h, void* stream, const GfsBaProblem* problems, int batch) {
  GFS_REQUIRE(h && problems, GFS_ERR_INVALID, "null pointer");
  GFS_REQUIRE(batch > 0 && batch <= h->maxBatch, GFS_ERR_CAPACITY, "batch exceeds the handle's max_batch");
  cudaStream_t st = (cudaStream_t)stream;
  BaDev& D = h->dev;
  const size_t B = batch;
  h->calib.assign(B, BaCalib());
  h->h_kf.assign(B * D.maxKf * KF_STRIDE, 0.0); h->h_pt.assign(B * D.maxPt * 3, 0.0);
  h->h_kfImu.assign(B * D.maxKf, 0); h->h_ptClose.assign(B * D.maxPt, 0);
  h->h_obsKf.assign(B * D.maxObs, 0); h->h_obsPt.assign(B * D.maxObs, 0); h->h_uvr.assign(B * D.maxObs * 3, 0.0);
  h->h_obsW.assign(B * D.maxObs, 0.f);
  h->h_inKf1.assign(B * D.maxIn, 0); h->h_inKf2.assign(B * D.maxIn, 0); h->h_inPre.assign(B * D.maxIn * GFS_BA_PRE_STRIDE, 0.f);
  h->h_infoIn.assign(B * D.maxIn * 81, 0.0); h->h_infoG.assign(B * D.maxIn * 9, 0.0); h->h_infoA.assign(B * D.maxIn * 9, 0.0);
  h->h_icpKf1.assign(B * D.maxIn, 0); h->h_icpKf2.assign(B * D.maxIn, 0); h->h_icpRt.assign(B * D.maxIn * 12, 0.0);
  h->h_ptStart.assign(B * (D.maxPt + 1), 0); h->h_ptEdges.assign(B * D.maxObs, 0);
  h->h_kfStart.assign(B * (D.maxKf + 1), 0); h->h_kfEdges.assign(B * D.maxObs, 0);
  h->h_ptKfEdge.assign(B * D.maxPt * D.maxOpt, -1);
  for (size_t b = 0; b < B; b++) {
    const GfsBaProblem& P = problems[b];
    const int nKf = P.n_opt_kf + P.n_fixed_kf;
    GFS_REQUIRE(P.n_opt_kf <= D.maxOpt, GFS_ERR_CAPACITY, "more than 21 optimizable keyframes");
    GFS_REQUIRE(P.n_opt_kf >= 0 && P.n_fixed_kf >= 0 && nKf <= D.maxKf && P.n_points >= 0 && P.n_points <= D.maxPt &&
                    P.n_obs >= 0 && P.n_obs <= D.maxObs && P.n_inertial >= 0 && P.n_inertial <= D.maxIn,
                GFS_ERR_CAPACITY, "problem exceeds the handle's capacity");
    BaCalib& C = h->calib[b];
    memcpy(C.Rcb, P.Rcb, 72); memcpy(C.tcb, P.tcb, 24); memcpy(C.Rbc, P.Rbc, 72); memcpy(C.tbc, P.tbc, 24);
    C.fx = P.fx; C.fy = P.fy; C.cx = P.cx; C.cy = P.cy; C.bf = P.bf; C.lambda_init = P.lambda_init;
    C.deltaMono = (double)(float)std::sqrt(5.991);
    C.deltaStereo = (double)(float)std::sqrt(7.815);
    C.nOpt = P.n_opt_kf; C.nKf = nKf; C.nPt = P.n_points; C.nObs = P.n_obs; C.nIn = P.n_inertial; C.nIcp = P.n_icp;
    GFS_REQUIRE(P.n_icp >= 0 && P.n_icp <= D.maxIn, GFS_ERR_CAPACITY, "too many ICP edges");
    for (int e = 0; e < P.n_icp; e++) {
      GFS_REQUIRE(P.icp_kf1[e] >= 0 && P.icp_kf1[e] < nKf && P.icp_kf2[e] >= 0 && P.icp_kf2[e] < nKf, GFS_ERR_INVALID, "ICP edge index out of range");
      h->h_icpKf1[b * D.maxIn + e] = P.icp_kf1[e]; h->h_icpKf2[b * D.maxIn + e] = P.icp_kf2[e];
      memcpy(&h->h_icpRt[(b * D.maxIn + e) * 12], P.icp_Rt + 12 * (size_t)e, 96);
    }
    C.iterations = P.iterations; C.bLarge = P.b_large; C.dimP = 15 * P.n_opt_kf;
    C.se3 = P.vertex_se3 ? 1 : 0;
    GFS_REQUIRE(!C.se3 || (P.n_inertial == 0 && P.n_icp == 0), GFS_ERR_INVALID, "vertex_se3 problems take no inertial / ICP edges");
    for (int k = 0; k < nKf; k++) {
      double* s = &h->h_kf[(b * D.maxKf + k) * KF_STRIDE];
      memcpy(s + K_RWB, P.kf_Rwb + 9 * (size_t)k, 72); memcpy(s + K_TWB, P.kf_twb + 3 * (size_t)k, 24);
      memcpy(s + K_RCW, P.kf_Rcw + 9 * (size_t)k, 72); memcpy(s + K_TCW, P.kf_tcw + 3 * (size_t)k, 24);
      memcpy(s + K_VEL, P.kf_vel + 3 * (size_t)k, 24); memcpy(s + K_BG, P.kf_bg + 3 * (size_t)k, 24);
      memcpy(s + K_BA, P.kf_ba + 3 * (size_t)k, 24);
      h->h_kfImu[b * D.maxKf + k] = P.kf_has_imu ? P.kf_has_imu[k] : 1;
    }
    if (P.n_points) {
      memcpy(&h->h_pt[b * D.maxPt * 3], P.pt_xyz, (size_t)P.n_points * 24);
      for (int j = 0; j < P.n_points; j++) h->h_ptClose[b * D.maxPt + j] = P.pt_close ? P.pt_close[j] : 1;
    }
    std::vector<int> pc(P.n_points + 1, 0), kc(D.maxKf + 1, 0);
    for (int e = 0; e < P.n_obs; e++) {
      const int k = P.obs_kf[e], j = P.obs_pt[e];
      GFS_REQUIRE(k >= 0 && k < nKf && j >= 0 && j < P.n_points, GFS_ERR_INVALID, "observation index out of range");
      h->h_obsKf[b * D.maxObs + e] = k; h->h_obsPt[b * D.maxObs + e] = j;
      memcpy(&h->h_uvr[(b * D.maxObs + e) * 3], P.obs_uvr + 3 * (size_t)e, 24);
      h->h_obsW[b * D.maxObs + e] = P.obs_inv_sigma2[e];
      pc[j + 1]++;
      if (k < P.n_opt_kf) {
        kc[k + 1]++;
        int& slot = h->h_ptKfEdge[(b * D.maxPt + j) * D.maxOpt + k];
        GFS_REQUIRE(slot < 0, GFS_ERR_INVALID, "two observations of one landmark in one keyframe are not supported");
        slot = e;
      }
    }
    for (int j = 0; j < P.n_points; j++) pc[j + 1] += pc[j];
    for (int k = 0; k < D.maxKf; k++) kc[k + 1] += kc[k];
    for (int j = 0; j <= P.n_points; j++) h->h_ptStart[b * (D.maxPt + 1) + j] = pc[j];
    for (int k = 0; k <= D.maxKf; k++) h->h_kfStart[b * (D.maxKf + 1) + k] = kc[k];
    std::vector<int> pcur(pc.begin(), pc.end() - 1), kcur(kc.begin(), kc.end() - 1);
    for (int e = 0; e < P.n_obs; e++) {  // edge order inside every list = creation order
      const int k = P.obs_kf[e], j = P.obs_pt[e];
      h->h_ptEdges[b * D.maxObs + pcur[j]++] = e;
      if (k < P.n_opt_kf) h->h_kfEdges[b * D.maxObs + kcur[k]++] = e;
    }
    for (int e = 0; e < P.n_inertial; e++) {
      GFS_REQUIRE(P.in_kf1[e] >= 0 && P.in_kf1[e] < nKf && P.in_kf2[e] >= 0 && P.in_kf2[e] < nKf, GFS_ERR_INVALID,
                  "inertial edge index out of range");
      h->h_inKf1[b * D.maxIn + e] = P.in_kf1[e]; h->h_inKf2[b * D.maxIn + e] = P.in_kf2[e];
      const float* rec = P.in_pre + (size_t)e * GFS_BA_PRE_STRIDE;
      memcpy(&h->h_inPre[(b * D.maxIn + e) * GFS_BA_PRE_STRIDE], rec, GFS_BA_PRE_STRIDE * 4);
      double* I = &h->h_infoIn[(b * D.maxIn + e) * 81];
      inertial_information(rec + 60, I);
      if (P.in_downweight && P.in_downweight[e])
        for (int i = 0; i < 81; i++) I[i] *= 1e-2;
      double G[9], A[9];
      for (int r = 0; r < 3; r++)
        for (int c = 0; c < 3; c++) { G[3 * r + c] = (double)rec[60 + 15 * (9 + r) + 9 + c]; A[3 * r + c] = (double)rec[60 + 15 * (12 + r) + 12 + c]; }
      inv3_host(G, &h->h_infoG[(b * D.maxIn + e) * 9]);
      inv3_host(A, &h->h_infoA[(b * D.maxIn + e) * 9]);
    }
  }
#define UP(dst, src) GFS_CUDA(cudaMemcpyAsync((void*)D.dst, h->src.data(), h->src.size() * sizeof(h->src[0]), cudaMemcpyHostToDevice, st))
  GFS_CUDA(cudaMemcpyAsync(D.calib, h->calib.data(), B * sizeof(BaCalib), cudaMemcpyHostToDevice, st));
  UP(kf, h_kf); UP(pt, h_pt); UP(kfImu, h_kfImu); UP(ptClose, h_ptClose);
  UP(obsKf, h_obsKf); UP(obsPt, h_obsPt); UP(obsUvr, h_uvr); UP(obsW, h_obsW);
  UP(icpKf1, h_icpKf1); UP(icpKf2, h_icpKf2); UP(icpRt, h_icpRt);
  UP(inKf1, h_inKf1); UP(inKf2, h_inKf2); UP(inPre, h_inPre); UP(infoIn, h_infoIn); UP(infoG, h_infoG); UP(infoA, h_infoA);
  UP(ptStart, h_ptStart); UP(ptEdges, h_ptEdges); UP(kfStart, h_kfStart); UP(kfEdges, h_kfEdges); UP(ptKfEdge, h_ptKfEdge);
#undef UP
  GFS_CUDA(gfs::stream_wait(st));  // the staging vectors may be reused by the next upload
  h->batch = batch;
  return GFS_OK;
}

// In-place SUM over the ranks of up to four fp64 device buffers.  With an NCCL communicator (gfs_ba_set_partition_nccl) they go
// out as ONE grouped launch on the solve stream, with no host synchronisation; the callback mode (gfs_ba_set_partition) syncs
// and calls out once per buffer.
struct RedBuf { double* p; int n; };
static int ba_allreduce(GfsBa* h, cudaStream_t st, std::initializer_list<RedBuf> bufs) {
  if (h->dev.world <= 1) return GFS_OK;
  if (h->ncclComm) {
    ncclrt::Api* n = ncclrt::api();
    int rc = n->groupStart();
    for (const RedBuf& b : bufs)
      if (rc == 0 && b.n > 0) rc = n->allReduce(b.p, b.p, (size_t)b.n, ncclrt::kDouble, ncclrt::kSum, h->ncclComm, st);
    const int rc2 = n->groupEnd();
    if (rc == 0) rc = rc2;
    if (rc != 0) { gfs::set_error("ncclAllReduce -> %s", n->errStr ? n->errStr(rc) : "error"); return GFS_ERR_CUDA; }
    h->ncclCalls++;
    return GFS_OK;
  }
  GFS_REQUIRE(h->allreduce, GFS_ERR_INVALID, "partitioned mode without a communicator or callback");
  GFS_CUDA(gfs::stream_wait(st));
  for (const RedBuf& b : bufs) {
    const int rc = h->allreduce(b.p, b.n, h->allreduceUser);
    GFS_REQUIRE(rc == 0, GFS_ERR_CUDA, "all-reduce callback failed");
  }
  return GFS_OK;
}

int gfs_ba_solve_uploaded(GfsBa* h, void* stream) {
  GFS_REQUIRE(h && h->batch > 0, GFS_ERR_INVALID, "nothing uploaded");
  cudaStream_t st = (cudaStream_t)stream;
  BaDev& D = h->dev;
  const int B = h->batch;
  const dim3 gObs(div_up(D.maxObs, ERR_THREADS), B), gIn(div_up(D.maxIn, 32), B), gPt(div_up(D.maxPt, 128), B);
  const dim3 gKfW(div_up(D.maxOpt, 4), B), gPairs(div_up(D.maxOpt * (D.maxOpt + 1) / 2, 4), B);
  const dim3 gUpd(div_up(D.maxPt + D.maxKf, 128), B);
  const dim3 gCopy(div_up(std::max(D.maxKf * KF_STRIDE, D.maxPt * 3), 256), B);
  const dim3 gHs(std::min(64, div_up(D.maxDim * D.maxDim, 256)), B);
  int* hc = (int*)h->h_counters.p;
  h->launches = 0;
  const bool part = D.world > 1;
  const size_t ldltSmem = (size_t)(2 * LDLT_MAXN * LDLT_PITCH + 32 + LDLT_MAXN) * sizeof(double);
  // the attribute belongs to the (device, function) pair: set it on every solve (a process may drive several GPUs, and host
  // threads may share the handle's device) -- one cheap driver call against ~200 launches
  GFS_CUDA(cudaFuncSetAttribute(k_ldlt_solve, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ldltSmem));
  k_ba_init<<<div_up(B, 128), 128, 0, st>>>(D, B);
  // reset the working state from the uploaded problem (re-solvable)
  GFS_CUDA(cudaMemcpyAsync(D.kf, h->h_kf.data(), h->h_kf.size() * 8, cudaMemcpyHostToDevice, st));
  GFS_CUDA(cudaMemcpyAsync(D.pt, h->h_pt.data(), h->h_pt.size() * 8, cudaMemcpyHostToDevice, st));
  auto errors = [&](int flag) {
    k_vis_error<<<gObs, ERR_THREADS, 0, st>>>(D, flag);
    k_in_error<<<gIn, 32, 0, st>>>(D, flag);
    h->launches += 2;
  };
  h->ncclCalls = 0;
  auto reduce_chi = [&](bool withScale) -> int {  // partitioned mode: sum the chi2 partials (and the gain-ratio scale) over ranks
    if (!part) return GFS_OK;
    if (withScale)
      return ba_allreduce(h, st, {{D.partChi, B * D.nblk}, {D.icpRho, B * D.maxIn}, {D.inRho, B * D.maxIn}, {D.partScale, B * D.nblk}});
    return ba_allreduce(h, st, {{D.partChi, B * D.nblk}, {D.icpRho, B * D.maxIn}, {D.inRho, B * D.maxIn}});
  };
  int rc;
  errors(J_ACTIVE);
  if ((rc = reduce_chi(false))) return rc;
  k_lm_control<<<div_up(B, 128), 128, 0, st>>>(D, B, 0, 0);
  h->launches += 2;
  int maxIt = 0;
  bool autoLambda = false;
  for (const BaCalib& c : h->calib) { maxIt = std::max(maxIt, c.iterations); autoLambda = autoLambda || !(c.lambda_init > 0); }
  for (int it = 0; it < maxIt; it++) {
    GFS_CUDA(cudaMemsetAsync(D.counters, 0, 8, st));
    errors(J_ACTIVE);  // computeActiveErrors at the current estimate
    if ((rc = reduce_chi(false))) return rc;
    k_lm_control<<<div_up(B, 128), 128, 0, st>>>(D, B, 1, it);
    GFS_CUDA(cudaMemsetAsync(D.Hpp, 0, (size_t)B * D.maxDim * D.maxDim * 8, st));
    GFS_CUDA(cudaMemsetAsync(D.bp, 0, (size_t)B * D.maxDim * 8, st));
    k_lin_points<<<gPt, 128, 0, st>>>(D);
    k_lin_kf<<<gKfW, 128, 0, st>>>(D);
    k_lin_inertial<<<B, INERTIAL_THREADS, 0, st>>>(D);
    if (it == 0 && autoLambda) { k_lambda_auto<<<B, 256, 0, st>>>(D); h->launches += 1; }
    k_backup<<<gCopy, 256, 0, st>>>(D, J_ACTIVE);
    h->launches += 5;
    for (int trial = 0; trial < 10; trial++) {
      k_schur_prep<<<gPt, 128, 0, st>>>(D);
      k_hs_init<<<gHs, 256, 0, st>>>(D);
      k_schur_pairs<<<gPairs, 128, 0, st>>>(D);
      k_schur_rhs<<<gKfW, 128, 0, st>>>(D);
      if (part) {  // sum of the per-shard reduced systems = the single-shard system
        if ((rc = ba_allreduce(h, st, {{D.Hs, B * D.maxDim * D.maxDim}, {D.bs, B * D.maxDim}}))) return rc;
      }
      k_ldlt_solve<<<B, LDLT_THREADS, ldltSmem, st>>>(D);
      k_backsub_update<<<gUpd, 128, 0, st>>>(D);
      errors(J_NEED);
      if ((rc = reduce_chi(true))) return rc;
      GFS_CUDA(cudaMemsetAsync(D.counters, 0, 4, st));
      k_lm_control<<<div_up(B, 128), 128, 0, st>>>(D, B, 2, it);
      k_restore<<<gCopy, 256, 0, st>>>(D);
      h->launches += 8;
      GFS_CUDA(cudaMemcpyAsync(hc, D.counters, 8, cudaMemcpyDeviceToHost, st));
      GFS_CUDA(gfs::stream_wait(st));
      if (hc[0] == 0) break;
    }
    if (hc[1] == 0) break;
  }
  k_ba_finish<<<gObs, 128, 0, st>>>(D, h->out);
  h->launches += 1;
  GFS_CUDA(cudaGetLastError());
  return GFS_OK;
}

int gfs_ba_download(GfsBa* h, void* stream, GfsBaResult* results, int batch) {
  GFS_REQUIRE(h && results && batch == h->batch, GFS_ERR_INVALID, "bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  BaDev& D = h->dev;
  const size_t B = batch;
  std::vector<double> kf(B * D.maxKf * KF_STRIDE), pt(B * D.maxPt * 3), chi2(B * D.maxObs), lam(B);
  std::vector<uint8_t> dpos(B * D.maxObs), outl(B * D.maxObs);
  std::vector<float> errs(B * 2);
  std::vector<int> info(B * 4);
  GFS_CUDA(cudaMemcpyAsync(kf.data(), D.kf, kf.size() * 8, cudaMemcpyDeviceToHost, st));
  GFS_CUDA(cudaMemcpyAsync(pt.data(), D.pt, pt.size() * 8, cudaMemcpyDeviceToHost, st));
  GFS_CUDA(cudaMemcpyAsync(chi2.data(), D.chi2, chi2.size() * 8, cudaMemcpyDeviceToHost, st));
  GFS_CUDA(cudaMemcpyAsync(dpos.data(), h->out.depthPos, dpos.size(), cudaMemcpyDeviceToHost, st));
  GFS_CUDA(cudaMemcpyAsync(outl.data(), h->out.outlier, outl.size(), cudaMemcpyDeviceToHost, st));
  GFS_CUDA(cudaMemcpyAsync(errs.data(), h->out.errs, errs.size() * 4, cudaMemcpyDeviceToHost, st));
  GFS_CUDA(cudaMemcpyAsync(info.data(), h->out.info, info.size() * 4, cudaMemcpyDeviceToHost, st));
  GFS_CUDA(cudaMemcpyAsync(lam.data(), h->out.lambda, lam.size() * 8, cudaMemcpyDeviceToHost, st));
  GFS_CUDA(gfs::stream_wait(st));
  for (size_t b = 0; b < B; b++) {
    const BaCalib& C = h->calib[b];
    GfsBaResult& R = results[b];
    for (int k = 0; k < C.nKf; k++) {
      const double* s = &kf[(b * D.maxKf + k) * KF_STRIDE];
      if (R.kf_Rwb) memcpy(R.kf_Rwb + 9 * (size_t)k, s + K_RWB, 72);
      if (R.kf_twb) memcpy(R.kf_twb + 3 * (size_t)k, s + K_TWB, 24);
      if (R.kf_Rcw) memcpy(R.kf_Rcw + 9 * (size_t)k, s + K_RCW, 72);
      if (R.kf_tcw) memcpy(R.kf_tcw + 3 * (size_t)k, s + K_TCW, 24);
      if (R.kf_vel) memcpy(R.kf_vel + 3 * (size_t)k, s + K_VEL, 24);
      if (R.kf_bg) memcpy(R.kf_bg + 3 * (size_t)k, s + K_BG, 24);
      if (R.kf_ba) memcpy(R.kf_ba + 3 * (size_t)k, s + K_BA, 24);
    }
    if (R.pt_xyz && C.nPt) memcpy(R.pt_xyz, &pt[b * D.maxPt * 3], (size_t)C.nPt * 24);
    if (R.obs_chi2 && C.nObs) memcpy(R.obs_chi2, &chi2[b * D.maxObs], (size_t)C.nObs * 8);
    if (R.obs_depth_positive && C.nObs) memcpy(R.obs_depth_positive, &dpos[b * D.maxObs], C.nObs);
    if (R.obs_outlier && C.nObs) memcpy(R.obs_outlier, &outl[b * D.maxObs], C.nObs);
    R.err = errs[2 * b]; R.err_end = errs[2 * b + 1];
    R.failed = info[4 * b]; R.iterations_done = info[4 * b + 1]; R.lm_trials = info[4 * b + 2];
    R.lambda_final = lam[b];
  }
  return GFS_OK;
}

int gfs_ba_solve_batch(GfsBa* h, void* stream, const GfsBaProblem* problems, GfsBaResult* results, int batch) {
  int rc = gfs_ba_upload(h, stream, problems, batch);
  if (rc) return rc;
  if ((rc = gfs_ba_solve_uploaded(h, stream))) return rc;
  return gfs_ba_download(h, stream, results, batch);
}

int gfs_ba_solve(GfsBa* h, void* stream, const GfsBaProblem* problem, GfsBaResult* result) {
  return gfs_ba_solve_batch(h, stream, problem, result, 1);
}
}

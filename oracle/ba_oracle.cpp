// ORACLE -- TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of the numerical core of Optimizer::LocalInertialBA (reference
// src/Optimizer.cc:3056-3702) on a flattened problem (include/gfs_b200.h GfsBaProblem):
//   vertices        ImuCamPose / VertexPose::oplusImpl        src/G2oTypes.cc:28-74,191-217
//                   VertexVelocity / GyroBias / AccBias        include/G2oTypes.h:181-236
//   EdgeMono/Stereo computeError, linearizeOplus              include/G2oTypes.h:317-332,399-405; src/G2oTypes.cc:335-360,385-416
//   EdgeInertial    ctor, computeError, linearizeOplus        src/G2oTypes.cc:473-719
//   EdgeGyroRW / EdgeAccRW                                     include/G2oTypes.h:782-852
//   IMU::Preintegrated::GetDelta{Rotation,Velocity,Position}  src/ImuTypes.cc:283-313 (float32)
//   Pinhole::project / projectJac                              src/CameraModels/Pinhole.cpp:36-42,73-83
//   SO3 helpers                                                src/G2oTypes.cc:1007-1082
//   g2o: Huber (robust_kernel_impl.cpp:77-91), constructQuadraticForm (base_binary_edge.hpp:56-118,
//        base_multi_edge.hpp:36-48), BlockSolver Schur solve (block_solver.hpp:354-486),
//        Levenberg (optimization_algorithm_levenberg.cpp:59-190), optimize loop (sparse_optimizer.cpp:354-420)
// written without Eigen/g2o.  The reference has no tests for this path ("parity unpinned");
// tests/test_oracle_ba.py pins the Jacobians by finite differences, the Schur solve against a dense
// numpy solve, and the optimisation against ground truth.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/gfs_b200.h"

namespace gfo {
namespace ba {

typedef double M3[9];  // row-major 3x3

static inline void mm3(const double* A, const double* B, double* C) {
  double t[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) t[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
  memcpy(C, t, sizeof(t));
}
static inline void mt3(const double* A, double* T) {
  double t[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) t[3 * r + c] = A[3 * c + r];
  memcpy(T, t, sizeof(t));
}
static inline void mv3(const double* A, const double* v, double* o) {
  double t[3];
  for (int r = 0; r < 3; r++) t[r] = A[3 * r] * v[0] + A[3 * r + 1] * v[1] + A[3 * r + 2] * v[2];
  memcpy(o, t, sizeof(t));
}
static inline void mtv3(const double* A, const double* v, double* o) {  // A^T v
  double t[3];
  for (int r = 0; r < 3; r++) t[r] = A[r] * v[0] + A[3 + r] * v[1] + A[6 + r] * v[2];
  memcpy(o, t, sizeof(t));
}
static bool inv3(const double* A, double* I) {
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  const double id = 1.0 / det;
  double t[9];
  t[0] = c00 * id; t[3] = c01 * id; t[6] = c02 * id;
  t[1] = (A[2] * A[7] - A[1] * A[8]) * id; t[4] = (A[0] * A[8] - A[2] * A[6]) * id; t[7] = (A[1] * A[6] - A[0] * A[7]) * id;
  t[2] = (A[1] * A[5] - A[2] * A[4]) * id; t[5] = (A[2] * A[3] - A[0] * A[5]) * id; t[8] = (A[0] * A[4] - A[1] * A[3]) * id;
  memcpy(I, t, sizeof(t));
  return det != 0 && std::isfinite(id);
}
static inline void skew(const double* w, double* W) {
  W[0] = 0; W[1] = -w[2]; W[2] = w[1]; W[3] = w[2]; W[4] = 0; W[5] = -w[0]; W[6] = -w[1]; W[7] = w[0]; W[8] = 0;
}
// NormalizeRotation (G2oTypes.h:69-74): U V^T of the SVD = orthogonal polar factor.  Computed by the
// Newton iteration X <- (X + X^-T)/2, which converges quadratically to the same matrix.
template <class T>
static void normalize_rotation(T* R) {
  for (int it = 0; it < 20; it++) {
    double A[9], I[9];
    for (int i = 0; i < 9; i++) A[i] = (double)R[i];
    if (!inv3(A, I)) return;
    double diff = 0;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        const T n = (T)(0.5 * (A[3 * r + c] + I[3 * c + r]));
        diff = std::max(diff, std::fabs((double)n - (double)R[3 * r + c]));
        R[3 * r + c] = n;
      }
    if (diff < (sizeof(T) == 4 ? 1e-7 : 1e-15)) break;
  }
}
// ExpSO3 (G2oTypes.cc:1011-1025)
static void exp_so3(const double* w, double* R) {
  const double d2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], d = std::sqrt(d2);
  double W[9], W2[9];
  skew(w, W);
  mm3(W, W, W2);
  if (d < 1e-5) {
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + W[i] + 0.5 * W2[i];
  } else {
    const double a = std::sin(d) / d, b = (1.0 - std::cos(d)) / d2;
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + W[i] * a + W2[i] * b;
  }
  normalize_rotation(R);
}
// LogSO3 (:1027-1041) -- note the float literal 0.5f and the early returns
static void log_so3(const double* R, double* w) {
  const double tr = R[0] + R[4] + R[8];
  w[0] = (R[7] - R[5]) / 2; w[1] = (R[2] - R[6]) / 2; w[2] = (R[3] - R[1]) / 2;
  const double costheta = (tr - 1.0) * 0.5f;
  if (costheta > 1 || costheta < -1) return;
  const double theta = std::acos(costheta), s = std::sin(theta);
  if (std::fabs(s) < 1e-5) return;
  for (int i = 0; i < 3; i++) w[i] = theta * w[i] / s;
}
// InverseRightJacobianSO3 (:1047-1061), RightJacobianSO3 (:1067-1082)
static void inv_right_jac(const double* v, double* J) {
  const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = std::sqrt(d2);
  double W[9], W2[9];
  skew(v, W);
  mm3(W, W, W2);
  if (d < 1e-5) { for (int i = 0; i < 9; i++) J[i] = (i % 4 == 0); return; }
  const double k = 1.0 / d2 - (1.0 + std::cos(d)) / (2.0 * d * std::sin(d));
  for (int i = 0; i < 9; i++) J[i] = (i % 4 == 0 ? 1.0 : 0.0) + W[i] / 2 + W2[i] * k;
}
static void right_jac(const double* v, double* J) {
  const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = std::sqrt(d2);
  double W[9], W2[9];
  skew(v, W);
  mm3(W, W, W2);
  if (d < 1e-5) { for (int i = 0; i < 9; i++) J[i] = (i % 4 == 0); return; }
  const double a = (1.0 - std::cos(d)) / d2, b = (d - std::sin(d)) / (d2 * d);
  for (int i = 0; i < 9; i++) J[i] = (i % 4 == 0 ? 1.0 : 0.0) - W[i] * a + W2[i] * b;
}

// ---- float32 preintegration read side (ImuTypes.cc:283-313)
struct Pre {
  const float *dR, *dV, *dP, *JRg, *JVg, *JVa, *JPg, *JPa, *C;
  float dT;
  const float* b;  // bax bay baz bwx bwy bwz
  explicit Pre(const float* r) {
    dR = r; dV = r + 9; dP = r + 12; JRg = r + 15; JVg = r + 24; JVa = r + 33; JPg = r + 42; JPa = r + 51; C = r + 60;
    dT = r[285]; b = r + 286;
  }
};
static void so3f_exp(const float* w, float* R) {  // Sophus::SO3f::exp(w).matrix()
  const float th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  float imag, real;
  if (th2 < 1e-5f * 1e-5f) {
    const float th4 = th2 * th2;
    imag = 0.5f - (1.0f / 48.0f) * th2 + (1.0f / 3840.0f) * th4;
    real = 1.0f - (1.0f / 8.0f) * th2 + (1.0f / 384.0f) * th4;
  } else {
    const float th = std::sqrt(th2), half = 0.5f * th;
    imag = std::sin(half) / th;
    real = std::cos(half);
  }
  const float qw = real, qx = imag * w[0], qy = imag * w[1], qz = imag * w[2];
  const float tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const float twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy,
              tyz = tz * qy, tzz = tz * qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
// corrected deltas for bias estimate (bg, ba) given as doubles, narrowed to float as IMU::Bias does
static void delta_for_bias(const Pre& p, const double* bg, const double* ba, double* dR, double* dV, double* dP, double* dbg_out) {
  const float dbg[3] = {(float)bg[0] - p.b[3], (float)bg[1] - p.b[4], (float)bg[2] - p.b[5]};
  const float dba[3] = {(float)ba[0] - p.b[0], (float)ba[1] - p.b[1], (float)ba[2] - p.b[2]};
  float w[3], E[9], R[9];
  for (int r = 0; r < 3; r++) w[r] = p.JRg[3 * r] * dbg[0] + p.JRg[3 * r + 1] * dbg[1] + p.JRg[3 * r + 2] * dbg[2];
  so3f_exp(w, E);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R[3 * r + c] = p.dR[3 * r] * E[c] + p.dR[3 * r + 1] * E[3 + c] + p.dR[3 * r + 2] * E[6 + c];
  normalize_rotation(R);
  for (int i = 0; i < 9; i++) dR[i] = (double)R[i];
  for (int r = 0; r < 3; r++) {
    const float v = p.dV[r] + (p.JVg[3 * r] * dbg[0] + p.JVg[3 * r + 1] * dbg[1] + p.JVg[3 * r + 2] * dbg[2]) +
                    (p.JVa[3 * r] * dba[0] + p.JVa[3 * r + 1] * dba[1] + p.JVa[3 * r + 2] * dba[2]);
    const float q = p.dP[r] + (p.JPg[3 * r] * dbg[0] + p.JPg[3 * r + 1] * dbg[1] + p.JPg[3 * r + 2] * dbg[2]) +
                    (p.JPa[3 * r] * dba[0] + p.JPa[3 * r + 1] * dba[1] + p.JPa[3 * r + 2] * dba[2]);
    dV[r] = (double)v;
    dP[r] = (double)q;
  }
  if (dbg_out) for (int i = 0; i < 3; i++) dbg_out[i] = (double)dbg[i];
}

// ---- general dense helpers
static bool invert_n(std::vector<double> A, int n, std::vector<double>& inv) {  // Gauss-Jordan, partial pivoting
  inv.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) inv[(size_t)i * n + i] = 1.0;
  for (int c = 0; c < n; c++) {
    int piv = c;
    for (int r = c + 1; r < n; r++)
      if (std::fabs(A[(size_t)r * n + c]) > std::fabs(A[(size_t)piv * n + c])) piv = r;
    if (A[(size_t)piv * n + c] == 0.0) return false;
    if (piv != c)
      for (int k = 0; k < n; k++) { std::swap(A[(size_t)c * n + k], A[(size_t)piv * n + k]); std::swap(inv[(size_t)c * n + k], inv[(size_t)piv * n + k]); }
    const double d = 1.0 / A[(size_t)c * n + c];
    for (int k = 0; k < n; k++) { A[(size_t)c * n + k] *= d; inv[(size_t)c * n + k] *= d; }
    for (int r = 0; r < n; r++) {
      if (r == c) continue;
      const double f = A[(size_t)r * n + c];
      if (f == 0.0) continue;
      for (int k = 0; k < n; k++) { A[(size_t)r * n + k] -= f * A[(size_t)c * n + k]; inv[(size_t)r * n + k] -= f * inv[(size_t)c * n + k]; }
    }
  }
  return true;
}
// symmetric eigen-decomposition by cyclic Jacobi: A = V diag(e) V^T
static void jacobi_eig(std::vector<double> A, int n, std::vector<double>& e, std::vector<double>& V) {
  V.assign((size_t)n * n, 0.0);
  for (int i = 0; i < n; i++) V[(size_t)i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0, diag = 0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) (i == j ? diag : off) += A[(size_t)i * n + j] * A[(size_t)i * n + j];
    if (off <= 1e-32 * diag) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = A[(size_t)p * n + q];
        if (apq == 0.0) continue;
        const double theta = (A[(size_t)q * n + q] - A[(size_t)p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (std::fabs(theta) + std::sqrt(theta * theta + 1.0));
        const double c = 1.0 / std::sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = A[(size_t)k * n + p], akq = A[(size_t)k * n + q];
          A[(size_t)k * n + p] = c * akp - s * akq;
          A[(size_t)k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = A[(size_t)p * n + k], aqk = A[(size_t)q * n + k];
          A[(size_t)p * n + k] = c * apk - s * aqk;
          A[(size_t)q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[(size_t)k * n + p], vkq = V[(size_t)k * n + q];
          V[(size_t)k * n + p] = c * vkp - s * vkq;
          V[(size_t)k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  e.resize(n);
  for (int i = 0; i < n; i++) e[i] = A[(size_t)i * n + i];
}
// EdgeInertial information (G2oTypes.cc:487-494): inverse of C[0:9,0:9], symmetrised, eigenvalues < 1e-12 zeroed
static void inertial_information(const float* C15, double* Info81) {
  std::vector<double> C(81), Ci, e, V;
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) C[9 * r + c] = (double)C15[15 * r + c];
  invert_n(C, 9, Ci);
  std::vector<double> S(81);
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) S[9 * r + c] = (Ci[9 * r + c] + Ci[9 * c + r]) / 2;
  jacobi_eig(S, 9, e, V);
  for (int i = 0; i < 9; i++)
    if (e[i] < 1e-12) e[i] = 0;
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) {
      double s = 0;
      for (int k = 0; k < 9; k++) s += V[9 * r + k] * e[k] * V[9 * c + k];
      Info81[9 * r + c] = s;
    }
}

struct KF {
  double Rwb[9], twb[3], Rcw[9], tcw[3], vel[3], bg[3], ba[3];
};
struct State {
  std::vector<KF> kf;
  std::vector<double> pt;
};

struct Problem {
  const GfsBaProblem* P;
  int nOpt, nKf, nPt, nObs, nIn, dimP;
  std::vector<double> infoIn;  // [nIn][81]
  std::vector<double> infoG, infoA;  // [nIn][9]
  double deltaMono, deltaStereo, deltaIn, deltaIcp;
};

static void se3_exp_update(KF& k, const double* u);  // VertexSE3Expmap::oplusImpl, defined after the SE3Quat helpers
// ImuCamPose::Update (G2oTypes.cc:191-217)
static void kf_update(const GfsBaProblem& P, KF& k, const double* pu) {
  if (P.vertex_se3) { se3_exp_update(k, pu); return; }
  double t[3], E[9];
  mv3(k.Rwb, pu + 3, t);
  for (int i = 0; i < 3; i++) k.twb[i] += t[i];
  exp_so3(pu, E);
  mm3(k.Rwb, E, k.Rwb);
  // (NormalizeRotation(Rwb) every third update: its result is discarded in the reference, :203-207)
  double Rbw[9], tbw[3];
  mt3(k.Rwb, Rbw);
  mv3(Rbw, k.twb, tbw);
  for (int i = 0; i < 3; i++) tbw[i] = -tbw[i];
  mm3(P.Rcb, Rbw, k.Rcw);
  mv3(P.Rcb, tbw, k.tcw);
  for (int i = 0; i < 3; i++) k.tcw[i] += P.tcb[i];
}

// visual edge error: obs - Project[Stereo](Xw)  (G2oTypes.cc:172-188)
static int vis_error(const GfsBaProblem& P, const KF& k, const double* Xw, const double* obs, double* err, double* Xc_out) {
  double Xc[3];
  mv3(k.Rcw, Xw, Xc);
  for (int i = 0; i < 3; i++) Xc[i] += k.tcw[i];
  const double u = P.fx * Xc[0] / Xc[2] + P.cx, v = P.fy * Xc[1] / Xc[2] + P.cy;
  if (Xc_out) memcpy(Xc_out, Xc, sizeof(Xc));
  err[0] = obs[0] - u;
  err[1] = obs[1] - v;
  if (obs[2] < 0) return 2;
  if (P.vertex_se3) {  // g2o::EdgeStereoSE3ProjectXYZ::cam_project: `const float invz = 1.0f/trans_xyz[2];` (types_six_dof_expmap.cpp:213-221)
    const float invz = (float)(1.0 / Xc[2]);
    const double us = Xc[0] * (double)invz * P.fx + P.cx;
    err[0] = obs[0] - us;
    err[1] = obs[1] - (Xc[1] * (double)invz * P.fy + P.cy);
    err[2] = obs[2] - (us - P.bf * (double)invz);
    return 3;
  }
  const double invZ = 1 / Xc[2];
  err[2] = obs[2] - (u - P.bf * invZ);
  return 3;
}
static inline void huber(double e, double delta, double* rho) {  // robust_kernel_impl.cpp:77-91
  const double dsqr = delta * delta;
  if (e <= dsqr) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
  else { const double sq = std::sqrt(e); rho[0] = 2 * sq * delta - dsqr; rho[1] = delta / sq; rho[2] = -0.5 * rho[1] / e; }
}

// EdgeInertial::computeError (G2oTypes.cc:495-522)
static void inertial_error_core(const Pre& pre, const KF& k1, const KF& k2, double* err9);
static void inertial_error(const Problem& pr, const State& s, int e, double* err9) {
  const GfsBaProblem& P = *pr.P;
  inertial_error_core(Pre(P.in_pre + (size_t)e * GFS_BA_PRE_STRIDE), s.kf[P.in_kf1[e]], s.kf[P.in_kf2[e]], err9);
}
static void inertial_error_core(const Pre& pre, const KF& k1, const KF& k2, double* err9) {
  double dR[9], dV[3], dP[3];
  delta_for_bias(pre, k1.bg, k1.ba, dR, dV, dP, nullptr);
  const double dt = (double)pre.dT;
  const double g[3] = {0, 0, -(double)9.81f};
  double dRt[9], Rbw1[9], A[9], eR[9];
  mt3(dR, dRt);
  mt3(k1.Rwb, Rbw1);
  mm3(dRt, Rbw1, A);
  mm3(A, k2.Rwb, eR);
  log_so3(eR, err9);
  double tv[3], tp[3];
  for (int i = 0; i < 3; i++) {
    tv[i] = k2.vel[i] - k1.vel[i] - g[i] * dt;
    tp[i] = k2.twb[i] - k1.twb[i] - k1.vel[i] * dt - g[i] * dt * dt / 2;
  }
  double rv[3], rp[3];
  mv3(Rbw1, tv, rv);
  mv3(Rbw1, tp, rp);
  for (int i = 0; i < 3; i++) { err9[3 + i] = rv[i] - dV[i]; err9[6 + i] = rp[i] - dP[i]; }
}
// EdgeInertial::linearizeOplus (:524-719): J is 9 x 24, columns [pose1(6) vel1(3) bg1(3) ba1(3) pose2(6) vel2(3)]
static void inertial_jacobian_core(const Pre& pre, const KF& k1, const KF& k2, double* J /*9x24*/);
static void inertial_jacobian(const Problem& pr, const State& s, int e, double* J /*9x24*/) {
  const GfsBaProblem& P = *pr.P;
  inertial_jacobian_core(Pre(P.in_pre + (size_t)e * GFS_BA_PRE_STRIDE), s.kf[P.in_kf1[e]], s.kf[P.in_kf2[e]], J);
}
static void inertial_jacobian_core(const Pre& pre, const KF& k1, const KF& k2, double* J /*9x24*/) {
  double dR[9], dV[3], dP[3], dbg[3];
  delta_for_bias(pre, k1.bg, k1.ba, dR, dV, dP, dbg);
  const double dt = (double)pre.dT;
  const double g[3] = {0, 0, -(double)9.81f};
  double JRg[9], JVg[9], JPg[9], JVa[9], JPa[9];
  for (int i = 0; i < 9; i++) { JRg[i] = pre.JRg[i]; JVg[i] = pre.JVg[i]; JPg[i] = pre.JPg[i]; JVa[i] = pre.JVa[i]; JPa[i] = pre.JPa[i]; }
  double Rbw1[9], dRt[9], A[9], eR[9], er[3], invJr[9];
  mt3(k1.Rwb, Rbw1);
  mt3(dR, dRt);
  mm3(dRt, Rbw1, A);
  mm3(A, k2.Rwb, eR);
  log_so3(eR, er);
  inv_right_jac(er, invJr);
  memset(J, 0, sizeof(double) * 9 * 24);
  auto put = [&](int r0, int c0, const double* M, double sgn) {
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) J[(r0 + r) * 24 + c0 + c] = sgn * M[3 * r + c];
  };
  double Rwb2t[9], T1[9], T2[9];
  mt3(k2.Rwb, Rwb2t);
  mm3(invJr, Rwb2t, T1);
  mm3(T1, k1.Rwb, T2);
  put(0, 0, T2, -1.0);  // -invJr * Rwb2^T * Rwb1
  double tv[3], tp[3], v[3], W[9];
  for (int i = 0; i < 3; i++) {
    tv[i] = k2.vel[i] - k1.vel[i] - g[i] * dt;
    tp[i] = k2.twb[i] - k1.twb[i] - k1.vel[i] * dt - 0.5 * g[i] * dt * dt;
  }
  mv3(Rbw1, tv, v); skew(v, W); put(3, 0, W, 1.0);
  mv3(Rbw1, tp, v); skew(v, W); put(6, 0, W, 1.0);
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  put(6, 3, I3, -1.0);
  // velocity 1
  put(3, 6, Rbw1, -1.0);
  { double M[9]; for (int i = 0; i < 9; i++) M[i] = Rbw1[i] * dt; put(6, 6, M, -1.0); }
  // gyro bias 1
  {
    double w[3], Jr[9], eRt[9], M1[9], M2[9], M3_[9];
    mv3(JRg, dbg, w);
    right_jac(w, Jr);
    mt3(eR, eRt);
    mm3(invJr, eRt, M1);
    mm3(M1, Jr, M2);
    mm3(M2, JRg, M3_);
    put(0, 9, M3_, -1.0);
    put(3, 9, JVg, -1.0);
    put(6, 9, JPg, -1.0);
  }
  // acc bias 1
  put(3, 12, JVa, -1.0);
  put(6, 12, JPa, -1.0);
  // pose 2
  put(0, 15, invJr, 1.0);
  { double M[9]; mm3(Rbw1, k2.Rwb, M); put(6, 18, M, 1.0); }
  // velocity 2
  put(3, 21, Rbw1, 1.0);
}

// ---- g2o::SE3Quat restated (Thirdparty/g2o/g2o/types/se3quat.h:40-215, se3_ops.hpp) for EdgeICP
struct Quat { double w, x, y, z; };
struct SE3Q { Quat r; double t[3]; };
static Quat quat_from_R(const double* m) {  // Eigen::Quaterniond(Matrix3d)
  Quat q;
  double t = m[0] + m[4] + m[8];
  if (t > 0) {
    t = std::sqrt(t + 1.0);
    q.w = 0.5 * t;
    t = 0.5 / t;
    q.x = (m[7] - m[5]) * t; q.y = (m[2] - m[6]) * t; q.z = (m[3] - m[1]) * t;
  } else {
    int i = 0;
    if (m[4] > m[0]) i = 1;
    if (m[8] > m[4 * i]) i = 2;
    const int j = (i + 1) % 3, k = (j + 1) % 3;
    t = std::sqrt(m[4 * i] - m[4 * j] - m[4 * k] + 1.0);
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
static void quat_normalize_rot(Quat& q) {  // SE3Quat::normalizeRotation
  if (q.w < 0) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  const double n = std::sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  q.w /= n; q.x /= n; q.y /= n; q.z /= n;
}
static Quat quat_mul(const Quat& a, const Quat& b) {
  return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
static void quat_rot(const Quat& q, const double* v, double* o) {  // Eigen _transformVector
  const double ux = 2 * (q.y * v[2] - q.z * v[1]), uy = 2 * (q.z * v[0] - q.x * v[2]), uz = 2 * (q.x * v[1] - q.y * v[0]);
  o[0] = v[0] + q.w * ux + (q.y * uz - q.z * uy);
  o[1] = v[1] + q.w * uy + (q.z * ux - q.x * uz);
  o[2] = v[2] + q.w * uz + (q.x * uy - q.y * ux);
}
static void quat_to_R(const Quat& q, double* R) {
  const double tx = 2 * q.x, ty = 2 * q.y, tz = 2 * q.z;
  const double twx = tx * q.w, twy = ty * q.w, twz = tz * q.w, txx = tx * q.x, txy = ty * q.x, txz = tz * q.x, tyy = ty * q.y,
               tyz = tz * q.y, tzz = tz * q.z;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
static SE3Q se3q_make(const double* R, const double* t) {
  SE3Q s; s.r = quat_from_R(R); quat_normalize_rot(s.r);
  s.t[0] = t[0]; s.t[1] = t[1]; s.t[2] = t[2];
  return s;
}
static SE3Q se3q_mul(const SE3Q& a, const SE3Q& b) {
  SE3Q r = a;
  double rt[3];
  quat_rot(a.r, b.t, rt);
  for (int i = 0; i < 3; i++) r.t[i] += rt[i];
  r.r = quat_mul(a.r, b.r);
  quat_normalize_rot(r.r);
  return r;
}
static SE3Q se3q_inv(const SE3Q& a) {
  SE3Q r;
  r.r = Quat{a.r.w, -a.r.x, -a.r.y, -a.r.z};
  const double nt[3] = {-a.t[0], -a.t[1], -a.t[2]};
  quat_rot(r.r, nt, r.t);
  return r;
}
// SE3Quat::exp (Thirdparty/g2o/g2o/types/se3quat.h:223-257)
static SE3Q se3q_exp(const double* u) {
  const double w[3] = {u[0], u[1], u[2]}, ups[3] = {u[3], u[4], u[5]};
  const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  double O[9], O2[9], R[9], V[9];
  skew(w, O);
  mm3(O, O, O2);
  if (theta < 0.00001) {
    for (int i = 0; i < 9; i++) { R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
  } else {
    const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta), c = (theta - std::sin(theta)) / std::pow(theta, 3);
    for (int i = 0; i < 9; i++) {
      const double I = (i % 4 == 0) ? 1.0 : 0.0;
      R[i] = I + a * O[i] + b * O2[i];
      V[i] = I + b * O[i] + c * O2[i];
    }
  }
  SE3Q s;
  s.r = quat_from_R(R);
  quat_normalize_rot(s.r);
  mv3(V, ups, s.t);
  return s;
}
// g2o::VertexSE3Expmap::oplusImpl (types_six_dof_expmap.h:55-70): estimate = SE3Quat::exp(update) * estimate; the
// keyframe record keeps Tcw as (Rcw, tcw)
static void se3_exp_update(KF& k, const double* u) {
  const SE3Q T = se3q_mul(se3q_exp(u), se3q_make(k.Rcw, k.tcw));
  quat_to_R(T.r, k.Rcw);
  memcpy(k.tcw, T.t, sizeof(T.t));
}
static void se3q_log(const SE3Q& a, double* res) {
  double R[9];
  quat_to_R(a.r, R);
  const double d = 0.5 * (R[0] + R[4] + R[8] - 1);
  const double dR[3] = {R[7] - R[5], R[2] - R[6], R[3] - R[1]};
  double omega[3], Om[9], Om2[9], Vi[9];
  if (d > 0.99999) {
    for (int i = 0; i < 3; i++) omega[i] = 0.5 * dR[i];
    skew(omega, Om);
    mm3(Om, Om, Om2);
    for (int i = 0; i < 9; i++) Vi[i] = (i % 4 == 0 ? 1.0 : 0.0) - 0.5 * Om[i] + (1. / 12.) * Om2[i];
  } else {
    const double theta = std::acos(d);
    const double k = theta / (2 * std::sqrt(1 - d * d));
    for (int i = 0; i < 3; i++) omega[i] = k * dR[i];
    skew(omega, Om);
    mm3(Om, Om, Om2);
    const double c = (1 - theta / (2 * std::tan(theta / 2))) / (theta * theta);
    for (int i = 0; i < 9; i++) Vi[i] = (i % 4 == 0 ? 1.0 : 0.0) - 0.5 * Om[i] + c * Om2[i];
  }
  double ups[3];
  mv3(Vi, a.t, ups);
  for (int i = 0; i < 3; i++) { res[i] = omega[i]; res[3 + i] = ups[i]; }
}
// EdgeICP::computeError (include/G2oTypes.h:524-540): log(T_c1c2^-1 * T_c1w * T_c2w^-1)
static void icp_error_kf(const double* Rt, const KF& k1, const KF& k2, double* err6) {
  const SE3Q Tm = se3q_make(Rt, Rt + 9), T1 = se3q_make(k1.Rcw, k1.tcw), T2 = se3q_make(k2.Rcw, k2.tcw);
  se3q_log(se3q_mul(se3q_mul(se3q_inv(Tm), T1), se3q_inv(T2)), err6);
}

// pose-side layout: keyframe k (< nOpt) owns dofs [15k, 15k+15): pose 6, vel 3, bg 3, ba 3
struct System {
  int dimP, nPt;
  std::vector<double> Hpp, bp;       // dimP x dimP, dimP
  std::vector<double> Hll, bl;       // nPt x 9, nPt x 3
  std::vector<double> Hpl;           // nObs x 18 (6x3, row-major), zero for fixed keyframes
  std::vector<double> x;             // dimP + 3 nPt
};

static double robust_chi2(const Problem& pr, const State& s, std::vector<double>* chi2_vis) {
  const GfsBaProblem& P = *pr.P;
  double chi = 0, rho[3];
  for (int e = 0; e < pr.nIn; e++) {
    double r[9];
    inertial_error(pr, s, e, r);
    const double* I = &pr.infoIn[(size_t)e * 81];
    double c2 = 0;
    for (int a = 0; a < 9; a++)
      for (int b = 0; b < 9; b++) c2 += r[a] * I[9 * a + b] * r[b];
    huber(c2, pr.deltaIn, rho);
    chi += rho[0];
    const KF& k1 = s.kf[P.in_kf1[e]];
    const KF& k2 = s.kf[P.in_kf2[e]];
    double rg[3], ra[3];
    for (int i = 0; i < 3; i++) { rg[i] = k2.bg[i] - k1.bg[i]; ra[i] = k2.ba[i] - k1.ba[i]; }
    const double *G = &pr.infoG[(size_t)e * 9], *A = &pr.infoA[(size_t)e * 9];
    double cg = 0, ca = 0;
    for (int a = 0; a < 3; a++)
      for (int b = 0; b < 3; b++) { cg += rg[a] * G[3 * a + b] * rg[b]; ca += ra[a] * A[3 * a + b] * ra[b]; }
    chi += cg + ca;
  }
  for (int e = 0; e < P.n_icp; e++) {  // EdgeICP: information 1e2 * I, Huber sqrt(0.4) (Optimizer.cc:3258, 3306-3314)
    double r[6];
    icp_error_kf(P.icp_Rt + 12 * (size_t)e, s.kf[P.icp_kf1[e]], s.kf[P.icp_kf2[e]], r);
    double c2 = 0;
    for (int a = 0; a < 6; a++) c2 += r[a] * 1e2 * r[a];
    huber(c2, pr.deltaIcp, rho);
    chi += rho[0];
  }
  for (int e = 0; e < pr.nObs; e++) {
    double r[3];
    const int d = vis_error(P, s.kf[P.obs_kf[e]], &s.pt[3 * (size_t)P.obs_pt[e]], P.obs_uvr + 3 * (size_t)e, r, nullptr);
    const double w = (double)P.obs_inv_sigma2[e];
    double c2 = 0;
    for (int a = 0; a < d; a++) c2 += r[a] * w * r[a];
    if (chi2_vis) (*chi2_vis)[e] = c2;
    huber(c2, d == 2 ? pr.deltaMono : pr.deltaStereo, rho);
    chi += rho[0];
  }
  return chi;
}

static void build_system(const Problem& pr, const State& s, System& S) {
  const GfsBaProblem& P = *pr.P;
  const int dimP = pr.dimP;
  S.dimP = dimP; S.nPt = pr.nPt;
  S.Hpp.assign((size_t)dimP * dimP, 0.0); S.bp.assign(dimP, 0.0);
  S.Hll.assign((size_t)pr.nPt * 9, 0.0); S.bl.assign((size_t)pr.nPt * 3, 0.0);
  S.Hpl.assign((size_t)pr.nObs * 18, 0.0);
  double rho[3];
  // inertial + random-walk edges
  for (int e = 0; e < pr.nIn; e++) {
    const int k1 = P.in_kf1[e], k2 = P.in_kf2[e];
    double J[9 * 24], r[9];
    inertial_error(pr, s, e, r);
    inertial_jacobian(pr, s, e, J);
    const double* I = &pr.infoIn[(size_t)e * 81];
    double c2 = 0;
    for (int a = 0; a < 9; a++)
      for (int b = 0; b < 9; b++) c2 += r[a] * I[9 * a + b] * r[b];
    huber(c2, pr.deltaIn, rho);
    // global dof of each of the 24 columns (-1: fixed)
    int col[24];
    for (int c = 0; c < 15; c++) col[c] = (k1 < pr.nOpt) ? 15 * k1 + c : -1;
    for (int c = 0; c < 9; c++) col[15 + c] = (k2 < pr.nOpt) ? 15 * k2 + c : -1;
    double WJ[9 * 24], Wr[9];
    for (int a = 0; a < 9; a++) {
      for (int c = 0; c < 24; c++) {
        double t = 0;
        for (int b = 0; b < 9; b++) t += I[9 * a + b] * J[b * 24 + c];
        WJ[a * 24 + c] = rho[1] * t;
      }
      double t = 0;
      for (int b = 0; b < 9; b++) t += I[9 * a + b] * r[b];
      Wr[a] = -rho[1] * t;
    }
    for (int c1 = 0; c1 < 24; c1++) {
      if (col[c1] < 0) continue;
      double t = 0;
      for (int a = 0; a < 9; a++) t += J[a * 24 + c1] * Wr[a];
      S.bp[col[c1]] += t;
      for (int c2i = 0; c2i < 24; c2i++) {
        if (col[c2i] < 0) continue;
        double h = 0;
        for (int a = 0; a < 9; a++) h += J[a * 24 + c1] * WJ[a * 24 + c2i];
        S.Hpp[(size_t)col[c1] * dimP + col[c2i]] += h;
      }
    }
    // EdgeGyroRW / EdgeAccRW: r = b2 - b1, J = [-I, I], no robust kernel
    const KF& a1 = s.kf[k1];
    const KF& a2 = s.kf[k2];
    for (int which = 0; which < 2; which++) {
      const double* Om = which == 0 ? &pr.infoG[(size_t)e * 9] : &pr.infoA[(size_t)e * 9];
      double rr[3];
      for (int i = 0; i < 3; i++) rr[i] = which == 0 ? a2.bg[i] - a1.bg[i] : a2.ba[i] - a1.ba[i];
      const int o1 = (k1 < pr.nOpt) ? 15 * k1 + 9 + 3 * which : -1, o2 = (k2 < pr.nOpt) ? 15 * k2 + 9 + 3 * which : -1;
      double Or[3];
      for (int a = 0; a < 3; a++) Or[a] = Om[3 * a] * rr[0] + Om[3 * a + 1] * rr[1] + Om[3 * a + 2] * rr[2];
      for (int a = 0; a < 3; a++) {
        if (o1 >= 0) S.bp[o1 + a] += Or[a];   // -J1^T Om r with J1 = -I
        if (o2 >= 0) S.bp[o2 + a] += -Or[a];
        for (int b = 0; b < 3; b++) {
          if (o1 >= 0) S.Hpp[(size_t)(o1 + a) * dimP + o1 + b] += Om[3 * a + b];
          if (o2 >= 0) S.Hpp[(size_t)(o2 + a) * dimP + o2 + b] += Om[3 * a + b];
          if (o1 >= 0 && o2 >= 0) {
            S.Hpp[(size_t)(o1 + a) * dimP + o2 + b] += -Om[3 * a + b];
            S.Hpp[(size_t)(o2 + a) * dimP + o1 + b] += -Om[3 * b + a];
          }
        }
      }
    }
  }
  // EdgeICP: numeric Jacobians (BaseBinaryEdge::linearizeOplus, base_binary_edge.hpp:124-190) + quadratic form
  for (int e = 0; e < P.n_icp; e++) {
    const int k1 = P.icp_kf1[e], k2 = P.icp_kf2[e];
    const double* Rt = P.icp_Rt + 12 * (size_t)e;
    double r[6], J[6 * 12] = {0};
    icp_error_kf(Rt, s.kf[k1], s.kf[k2], r);
    const double delta = 1e-9, scalar = 1.0 / (2 * delta);
    for (int v = 0; v < 2; v++) {
      const int k = v == 0 ? k1 : k2;
      if (k >= pr.nOpt) continue;  // fixed vertex
      for (int d = 0; d < 6; d++) {
        double add[6] = {0, 0, 0, 0, 0, 0}, ep[6], em[6];
        KF kp = s.kf[k], km = s.kf[k];
        add[d] = delta; kf_update(P, kp, add);
        add[d] = -delta; kf_update(P, km, add);
        icp_error_kf(Rt, v == 0 ? kp : s.kf[k1], v == 0 ? s.kf[k2] : kp, ep);
        icp_error_kf(Rt, v == 0 ? km : s.kf[k1], v == 0 ? s.kf[k2] : km, em);
        for (int a = 0; a < 6; a++) J[a * 12 + 6 * v + d] = scalar * (ep[a] - em[a]);
      }
    }
    double c2 = 0;
    for (int a = 0; a < 6; a++) c2 += r[a] * 1e2 * r[a];
    huber(c2, pr.deltaIcp, rho);
    const double w = rho[1] * 1e2;
    int col[12];
    for (int c = 0; c < 6; c++) { col[c] = k1 < pr.nOpt ? 15 * k1 + c : -1; col[6 + c] = k2 < pr.nOpt ? 15 * k2 + c : -1; }
    for (int c1 = 0; c1 < 12; c1++) {
      if (col[c1] < 0) continue;
      double t = 0;
      for (int a = 0; a < 6; a++) t += J[a * 12 + c1] * (-w * r[a]);
      S.bp[col[c1]] += t;
      for (int c2i = 0; c2i < 12; c2i++) {
        if (col[c2i] < 0) continue;
        double h = 0;
        for (int a = 0; a < 6; a++) h += J[a * 12 + c1] * w * J[a * 12 + c2i];
        S.Hpp[(size_t)col[c1] * dimP + col[c2i]] += h;
      }
    }
  }
  // visual edges (EdgeMono / EdgeStereo linearizeOplus + BaseBinaryEdge::constructQuadraticForm)
  for (int e = 0; e < pr.nObs; e++) {
    const int k = P.obs_kf[e], j = P.obs_pt[e];
    const KF& kf = s.kf[k];
    double r[3], Xc[3];
    const int d = vis_error(P, kf, &s.pt[3 * (size_t)j], P.obs_uvr + 3 * (size_t)e, r, Xc);
    const double w0 = (double)P.obs_inv_sigma2[e];
    double c2 = 0;
    for (int a = 0; a < d; a++) c2 += r[a] * w0 * r[a];
    huber(c2, d == 2 ? pr.deltaMono : pr.deltaStereo, rho);
    const double w = rho[1] * w0;
    double pj[9] = {0};
    pj[0] = P.fx / Xc[2]; pj[1] = 0.f; pj[2] = -P.fx * Xc[0] / (Xc[2] * Xc[2]);
    pj[3] = 0.f; pj[4] = P.fy / Xc[2]; pj[5] = -P.fy * Xc[1] / (Xc[2] * Xc[2]);
    if (d == 3) { pj[6] = pj[0]; pj[7] = pj[1]; pj[8] = pj[2] + P.bf * (1.0 / (Xc[2] * Xc[2])); }
    double Jl[9] = {0};  // d x 3 = -proj_jac * Rcw
    for (int a = 0; a < d; a++)
      for (int c = 0; c < 3; c++) Jl[3 * a + c] = -(pj[3 * a] * kf.Rcw[c] + pj[3 * a + 1] * kf.Rcw[3 + c] + pj[3 * a + 2] * kf.Rcw[6 + c]);
    // VertexPose: d x 6 = proj_jac * Rcb * SE3deriv(Xb).  VertexSE3Expmap (EdgeSE3ProjectXYZ, OptimizableTypes.cpp;
    // g2o::EdgeStereoSE3ProjectXYZ, types_six_dof_expmap.cpp:228-274): d x 6 = -proj_jac * SE3deriv(Xc)
    double Xb[3];
    if (P.vertex_se3) {
      memcpy(Xb, Xc, sizeof(Xb));
    } else {
      mv3(P.Rbc, Xc, Xb);
      for (int i = 0; i < 3; i++) Xb[i] += P.tbc[i];
    }
    const double D[18] = {0.0, Xb[2], -Xb[1], 1.0, 0.0, 0.0, -Xb[2], 0.0, Xb[0], 0.0, 1.0, 0.0, Xb[1], -Xb[0], 0.0, 0.0, 0.0, 1.0};
    double PR[9] = {0}, Jp[18] = {0};
    for (int a = 0; a < d; a++)
      for (int c = 0; c < 3; c++)
        PR[3 * a + c] = P.vertex_se3 ? -pj[3 * a + c] : pj[3 * a] * P.Rcb[c] + pj[3 * a + 1] * P.Rcb[3 + c] + pj[3 * a + 2] * P.Rcb[6 + c];
    for (int a = 0; a < d; a++)
      for (int c = 0; c < 6; c++) Jp[6 * a + c] = PR[3 * a] * D[c] + PR[3 * a + 1] * D[6 + c] + PR[3 * a + 2] * D[12 + c];
    // point block
    for (int a = 0; a < 3; a++) {
      double t = 0;
      for (int q = 0; q < d; q++) t += Jl[3 * q + a] * (-w * r[q]);
      S.bl[3 * (size_t)j + a] += t;
      for (int b = 0; b < 3; b++) {
        double h = 0;
        for (int q = 0; q < d; q++) h += Jl[3 * q + a] * w * Jl[3 * q + b];
        S.Hll[9 * (size_t)j + 3 * a + b] += h;
      }
    }
    if (k < pr.nOpt) {
      const int o = 15 * k;
      for (int a = 0; a < 6; a++) {
        double t = 0;
        for (int q = 0; q < d; q++) t += Jp[6 * q + a] * (-w * r[q]);
        S.bp[o + a] += t;
        for (int b = 0; b < 6; b++) {
          double h = 0;
          for (int q = 0; q < d; q++) h += Jp[6 * q + a] * w * Jp[6 * q + b];
          S.Hpp[(size_t)(o + a) * dimP + o + b] += h;
        }
        for (int b = 0; b < 3; b++) {
          double h = 0;
          for (int q = 0; q < d; q++) h += Jp[6 * q + a] * w * Jl[3 * q + b];
          S.Hpl[18 * (size_t)e + 3 * a + b] = h;
        }
      }
    }
  }
}

// dense LDL^T without pivoting, in place on the lower triangle; false on a zero / non-finite pivot
static bool ldlt_solve(std::vector<double>& A, int n, std::vector<double>& b) {
  for (int j = 0; j < n; j++) {
    double d = A[(size_t)j * n + j];
    for (int k = 0; k < j; k++) d -= A[(size_t)j * n + k] * A[(size_t)j * n + k] * A[(size_t)k * n + k];
    if (d == 0.0 || !std::isfinite(d)) return false;
    A[(size_t)j * n + j] = d;
    for (int i = j + 1; i < n; i++) {
      double v = A[(size_t)i * n + j];
      for (int k = 0; k < j; k++) v -= A[(size_t)i * n + k] * A[(size_t)j * n + k] * A[(size_t)k * n + k];
      A[(size_t)i * n + j] = v / d;
    }
  }
  for (int i = 0; i < n; i++) {
    double v = b[i];
    for (int k = 0; k < i; k++) v -= A[(size_t)i * n + k] * b[k];
    b[i] = v;
  }
  for (int i = 0; i < n; i++) b[i] /= A[(size_t)i * n + i];
  for (int i = n - 1; i >= 0; i--) {
    double v = b[i];
    for (int k = i + 1; k < n; k++) v -= A[(size_t)k * n + i] * b[k];
    b[i] = v;
  }
  return true;
}

// BlockSolver::solve with Schur complement (block_solver.hpp:354-486) for damping lambda
static bool schur_solve(const Problem& pr, const System& S, double lambda, std::vector<double>& x) {
  const GfsBaProblem& P = *pr.P;
  const int n = S.dimP;
  std::vector<double> Hs = S.Hpp, bs = S.bp;
  std::vector<uint8_t> used(n, 0);
  for (int k = 0; k < pr.nOpt; k++) {
    const int nd = P.kf_has_imu[k] ? 15 : 6;
    for (int d = 0; d < nd; d++) used[15 * k + d] = 1;
  }
  for (int i = 0; i < n; i++) {
    if (used[i]) Hs[(size_t)i * n + i] += lambda;
    else { Hs[(size_t)i * n + i] = 1.0; bs[i] = 0.0; }  // dofs that are not vertices (keyframe without IMU)
  }
  std::vector<double> Dinv((size_t)pr.nPt * 9), db((size_t)pr.nPt * 3);
  std::vector<std::vector<int>> byPt(pr.nPt);
  for (int e = 0; e < pr.nObs; e++)
    if (P.obs_kf[e] < pr.nOpt) byPt[P.obs_pt[e]].push_back(e);
  for (int j = 0; j < pr.nPt; j++) {
    double D[9];
    memcpy(D, &S.Hll[9 * (size_t)j], sizeof(D));
    D[0] += lambda; D[4] += lambda; D[8] += lambda;
    inv3(D, &Dinv[9 * (size_t)j]);
    mv3(&Dinv[9 * (size_t)j], &S.bl[3 * (size_t)j], &db[3 * (size_t)j]);
    for (int e1 : byPt[j]) {
      const double* B1 = &S.Hpl[18 * (size_t)e1];
      const int o1 = 15 * P.obs_kf[e1];
      double BD[18];
      for (int a = 0; a < 6; a++)
        for (int c = 0; c < 3; c++)
          BD[3 * a + c] = B1[3 * a] * Dinv[9 * (size_t)j + c] + B1[3 * a + 1] * Dinv[9 * (size_t)j + 3 + c] + B1[3 * a + 2] * Dinv[9 * (size_t)j + 6 + c];
      for (int a = 0; a < 6; a++) bs[o1 + a] -= B1[3 * a] * db[3 * (size_t)j] + B1[3 * a + 1] * db[3 * (size_t)j + 1] + B1[3 * a + 2] * db[3 * (size_t)j + 2];
      for (int e2 : byPt[j]) {
        const double* B2 = &S.Hpl[18 * (size_t)e2];
        const int o2 = 15 * P.obs_kf[e2];
        for (int a = 0; a < 6; a++)
          for (int b = 0; b < 6; b++)
            Hs[(size_t)(o1 + a) * n + o2 + b] -= BD[3 * a] * B2[3 * b] + BD[3 * a + 1] * B2[3 * b + 1] + BD[3 * a + 2] * B2[3 * b + 2];
      }
    }
  }
  if (!ldlt_solve(Hs, n, bs)) return false;
  x.assign((size_t)n + 3 * (size_t)pr.nPt, 0.0);
  for (int i = 0; i < n; i++) x[i] = used[i] ? bs[i] : 0.0;
  for (int j = 0; j < pr.nPt; j++) {
    double c[3] = {S.bl[3 * (size_t)j], S.bl[3 * (size_t)j + 1], S.bl[3 * (size_t)j + 2]};
    for (int e : byPt[j]) {
      const double* B = &S.Hpl[18 * (size_t)e];
      const double* xp = &x[15 * (size_t)P.obs_kf[e]];
      for (int b = 0; b < 3; b++)
        for (int a = 0; a < 6; a++) c[b] -= B[3 * a + b] * xp[a];
    }
    mv3(&Dinv[9 * (size_t)j], c, &x[(size_t)n + 3 * (size_t)j]);
  }
  return true;
}

static void apply_update(const Problem& pr, State& s, const std::vector<double>& x) {
  const GfsBaProblem& P = *pr.P;
  for (int k = 0; k < pr.nOpt; k++) {
    const double* u = &x[15 * (size_t)k];
    kf_update(P, s.kf[k], u);
    if (P.kf_has_imu[k])
      for (int i = 0; i < 3; i++) { s.kf[k].vel[i] += u[6 + i]; s.kf[k].bg[i] += u[9 + i]; s.kf[k].ba[i] += u[12 + i]; }
  }
  for (size_t i = 0; i < s.pt.size(); i++) s.pt[i] += x[(size_t)pr.dimP + i];
}

static void setup(const GfsBaProblem* P, Problem& pr, State& s) {
  pr.P = P;
  pr.nOpt = P->n_opt_kf; pr.nKf = P->n_opt_kf + P->n_fixed_kf; pr.nPt = P->n_points; pr.nObs = P->n_obs; pr.nIn = P->n_inertial;
  pr.dimP = 15 * pr.nOpt;
  pr.deltaMono = (double)(float)std::sqrt(5.991);    // const float thHuberMono = sqrt(5.991), Optimizer.cc:3427
  pr.deltaStereo = (double)(float)std::sqrt(7.815);  // :3429
  pr.deltaIn = std::sqrt(16.0);                      // :3372
  pr.deltaIcp = (double)(float)std::sqrt(0.4);       // const float thHuberICP = sqrt(0.4), :3258
  pr.infoIn.resize((size_t)pr.nIn * 81); pr.infoG.resize((size_t)pr.nIn * 9); pr.infoA.resize((size_t)pr.nIn * 9);
  for (int e = 0; e < pr.nIn; e++) {
    const float* C = P->in_pre + (size_t)e * GFS_BA_PRE_STRIDE + 60;
    inertial_information(C, &pr.infoIn[(size_t)e * 81]);
    if (P->in_downweight[e])
      for (int i = 0; i < 81; i++) pr.infoIn[(size_t)e * 81 + i] *= 1e-2;
    double G[9], A[9];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) { G[3 * r + c] = (double)C[15 * (9 + r) + 9 + c]; A[3 * r + c] = (double)C[15 * (12 + r) + 12 + c]; }
    inv3(G, &pr.infoG[(size_t)e * 9]);
    inv3(A, &pr.infoA[(size_t)e * 9]);
  }
  s.kf.resize(pr.nKf);
  for (int k = 0; k < pr.nKf; k++) {
    memcpy(s.kf[k].Rwb, P->kf_Rwb + 9 * (size_t)k, 72); memcpy(s.kf[k].twb, P->kf_twb + 3 * (size_t)k, 24);
    memcpy(s.kf[k].Rcw, P->kf_Rcw + 9 * (size_t)k, 72); memcpy(s.kf[k].tcw, P->kf_tcw + 3 * (size_t)k, 24);
    memcpy(s.kf[k].vel, P->kf_vel + 3 * (size_t)k, 24); memcpy(s.kf[k].bg, P->kf_bg + 3 * (size_t)k, 24);
    memcpy(s.kf[k].ba, P->kf_ba + 3 * (size_t)k, 24);
  }
  s.pt.assign(P->pt_xyz, P->pt_xyz + 3 * (size_t)pr.nPt);
}

static void solve(const GfsBaProblem* P, GfsBaResult* R) {
  Problem pr;
  State s;
  setup(P, pr, s);
  std::vector<double> chi2v(pr.nObs, 0.0);
  // optimizer.computeActiveErrors(); err = activeRobustChi2()  (Optimizer.cc:3589-3590)
  double lastChi = robust_chi2(pr, s, &chi2v);
  R->err = (float)lastChi;
  // optimizer.optimize(opt_it): sparse_optimizer.cpp:354-420 + optimization_algorithm_levenberg.cpp:59-164
  double lambda = 0, ni = 2;
  int nBad = 0, done = 0, trials = 0;
  System S;
  std::vector<double> x;
  for (int it = 0; it < P->iterations; it++) {
    double currentChi = robust_chi2(pr, s, &chi2v);
    lastChi = currentChi;
    const double iniChi = currentChi;
    double tempChi = currentChi;
    build_system(pr, s, S);
    if (it == 0) {
      lambda = P->lambda_init;
      if (!(lambda > 0)) {  // computeLambdaInit: tau * max |H_jj| over every vertex (optimization_algorithm_levenberg.cpp:166-178)
        double md = 0;
        for (int j = 0; j < pr.dimP; j++) md = std::max(std::fabs(S.Hpp[(size_t)j * pr.dimP + j]), md);
        for (int j = 0; j < pr.nPt; j++)
          for (int a = 0; a < 3; a++) md = std::max(std::fabs(S.Hll[9 * (size_t)j + 4 * a]), md);
        lambda = 1e-5 * md;
      }
      ni = 2; nBad = 0;
    }
    double rho = 0;
    int qmax = 0;
    do {
      State backup = s;
      const bool ok2 = schur_solve(pr, S, lambda, x);
      if (ok2) apply_update(pr, s, x);
      else x.assign((size_t)pr.dimP + 3 * (size_t)pr.nPt, 0.0);
      tempChi = robust_chi2(pr, s, &chi2v);
      lastChi = tempChi;
      if (!ok2) tempChi = std::numeric_limits<double>::max();
      rho = currentChi - tempChi;
      double scale = 0;
      for (int j = 0; j < pr.dimP; j++) scale += x[j] * (lambda * x[j] + S.bp[j]);
      for (size_t j = 0; j < 3 * (size_t)pr.nPt; j++) scale += x[(size_t)pr.dimP + j] * (lambda * x[(size_t)pr.dimP + j] + S.bl[j]);
      scale += 1e-3;
      rho /= scale;
      if (rho > 0 && std::isfinite(tempChi)) {
        double alpha = 1. - std::pow((2 * rho - 1), 3);
        alpha = std::min(alpha, 2. / 3.);
        const double sf = std::max(1. / 3., alpha);
        lambda *= sf;
        ni = 2;
        currentChi = tempChi;
      } else {
        lambda *= ni;
        ni *= 2;
        s = backup;  // pop(): the per-edge errors keep the values of the rejected trial
      }
      qmax++;
      trials++;
    } while (rho < 0 && qmax < 10);
    done++;
    if (qmax == 10 || rho == 0) break;  // Terminate
    if ((iniChi - currentChi) * 1e3 < iniChi) nBad++;
    else nBad = 0;
    if (nBad >= 3) break;
  }
  // err_end = activeRobustChi2() over the stored edge errors, i.e. those of the last evaluated LM
  // trial, accepted or not (g2o does not recompute errors after pop(); Optimizer.cc:3592)
  R->err_end = (float)lastChi;
  R->iterations_done = done;
  R->lm_trials = trials;
  R->lambda_final = lambda;
  R->failed = (!P->vertex_se3 && (2 * R->err < R->err_end || std::isnan(R->err) || std::isnan(R->err_end)) && !P->b_large) ? 1 : 0;
  const float chi2Mono2 = 5.991f, chi2Stereo2 = 7.815f;
  for (int e = 0; e < pr.nObs; e++) {
    const int k = P->obs_kf[e], j = P->obs_pt[e];
    const bool mono = P->obs_uvr[3 * (size_t)e + 2] < 0;
    const double c2 = chi2v[e];
    R->obs_chi2[e] = c2;
    const KF& kf = s.kf[k];
    const double z = kf.Rcw[6] * s.pt[3 * (size_t)j] + kf.Rcw[7] * s.pt[3 * (size_t)j + 1] + kf.Rcw[8] * s.pt[3 * (size_t)j + 2] + kf.tcw[2];
    const bool dpos = (mono || P->vertex_se3) ? (z > 0.0) : true;
    R->obs_depth_positive[e] = dpos;
    bool out;
    if (P->vertex_se3) {  // LocalBundleAdjustment (Optimizer.cc:1972, 1996)
      out = c2 > (mono ? 5.991 : 7.815) || !dpos;
    } else if (mono) {
      const bool close = P->pt_close[j] != 0;
      out = (c2 > chi2Mono2 && !close) || (c2 > 1.5f * chi2Mono2 && close) || !dpos;
    } else {
      out = c2 > chi2Stereo2;
    }
    R->obs_outlier[e] = out;
  }
  for (int k = 0; k < pr.nKf; k++) {
    memcpy(R->kf_Rwb + 9 * (size_t)k, s.kf[k].Rwb, 72); memcpy(R->kf_twb + 3 * (size_t)k, s.kf[k].twb, 24);
    memcpy(R->kf_Rcw + 9 * (size_t)k, s.kf[k].Rcw, 72); memcpy(R->kf_tcw + 3 * (size_t)k, s.kf[k].tcw, 24);
    memcpy(R->kf_vel + 3 * (size_t)k, s.kf[k].vel, 24); memcpy(R->kf_bg + 3 * (size_t)k, s.kf[k].bg, 24);
    memcpy(R->kf_ba + 3 * (size_t)k, s.kf[k].ba, 24);
  }
  memcpy(R->pt_xyz, s.pt.data(), sizeof(double) * s.pt.size());
}


// =================================================================================================
// Inertial pose-only optimisers of the tracking thread:
// Optimizer::PoseInertialOptimizationLastKeyFrame (reference src/Optimizer.cc:5899-6284) and
// Optimizer::PoseInertialOptimizationLastFrame (:6762-7172) on the flattened GfsPoseInertialProblem.
//   vertices  VertexPose (ImuCamPose::Update, G2oTypes.cc:193-217), VertexVelocity / GyroBias / AccBias
//   edges     EdgeMonoOnlyPose / EdgeStereoOnlyPose (G2oTypes.h:354-456, G2oTypes.cc:362-383,418-443),
//             EdgeInertial (cores above), EdgeGyroRW / EdgeAccRW (G2oTypes.h:782-852),
//             EdgePriorPoseImu (G2oTypes.cc:927-995)
//   solver    OptimizationAlgorithmGaussNewton (optimization_algorithm_gauss_newton.cpp:49-93): errors,
//             linearise, dense Eigen LDLT over all unknowns in vertex-id order (linear_solver_dense.h:66-114),
//             update; a failed factorisation re-applies the previous x and ends the round
//   hand-over GetHessian / GetHessian2 blocks, Optimizer::Marginalize (:4408-4487; the JacobiSVD
//             pseudo-inverse of a symmetric matrix is restated through its eigen-decomposition),
//             ConstraintPoseImu's eigenvalue clamp (G2oTypes.h:857-868)
// The reference has no tests for this path: parity unpinned; tests/test_oracle_pose_inertial.py checks the
// Jacobians by finite differences, the marginalisation against numpy and the result against ground truth.
// =================================================================================================
extern "C" int gfo_eigen_ldlt_solve(const double* H, const double* b, int n, double* x);  // pose_oracle.cpp

namespace pin {

struct Cam { double fx, fy, cx, cy, bf; const double *Rcb, *tcb, *tbc; };

static void cam_update(const Cam& C, KF& k, const double* pu) {  // ImuCamPose::Update
  double t[3], E[9];
  mv3(k.Rwb, pu + 3, t);
  for (int i = 0; i < 3; i++) k.twb[i] += t[i];
  exp_so3(pu, E);
  mm3(k.Rwb, E, k.Rwb);
  double Rbw[9], tbw[3];
  mt3(k.Rwb, Rbw);
  mv3(Rbw, k.twb, tbw);
  for (int i = 0; i < 3; i++) tbw[i] = -tbw[i];
  mm3(C.Rcb, Rbw, k.Rcw);
  mv3(C.Rcb, tbw, k.tcw);
  for (int i = 0; i < 3; i++) k.tcw[i] += C.tcb[i];
}
// obs - Project[Stereo](Xw); returns the dimension
static int vis_err(const Cam& C, const KF& k, const double* Xw, const float* o, double* err, double* Xc_out) {
  double Xc[3];
  mv3(k.Rcw, Xw, Xc);
  for (int i = 0; i < 3; i++) Xc[i] += k.tcw[i];
  if (Xc_out) memcpy(Xc_out, Xc, sizeof(Xc));
  const double u = C.fx * Xc[0] / Xc[2] + C.cx, v = C.fy * Xc[1] / Xc[2] + C.cy;
  err[0] = (double)o[0] - u;
  err[1] = (double)o[1] - v;
  err[2] = 0;
  if (o[2] < 0) return 2;
  const double invZ = 1 / Xc[2];
  err[2] = (double)o[2] - (u - C.bf * invZ);
  return 3;
}
// Edge{Mono,Stereo}OnlyPose::linearizeOplus: J = proj_jac * Rcb * SE3deriv(Xb), d x 6 (unused rows zero)
static void vis_jac(const Cam& C, const KF& k, const double* Xw, bool mono, double* J /*3x6*/) {
  double Xc[3], Xb[3];
  mv3(k.Rcw, Xw, Xc);
  for (int i = 0; i < 3; i++) Xc[i] += k.tcw[i];
  mtv3(C.Rcb, Xc, Xb);  // Rbc = Rcb^T
  for (int i = 0; i < 3; i++) Xb[i] += C.tbc[i];
  double pj[9] = {C.fx / Xc[2], 0.0, -C.fx * Xc[0] / (Xc[2] * Xc[2]), 0.0, C.fy / Xc[2], -C.fy * Xc[1] / (Xc[2] * Xc[2]), 0, 0, 0};
  if (!mono) {
    const double inv_z2 = 1.0 / (Xc[2] * Xc[2]);
    pj[6] = pj[0]; pj[7] = pj[1]; pj[8] = pj[2] + C.bf * inv_z2;
  }
  double A[9];
  mm3(pj, C.Rcb, A);
  const double x = Xb[0], y = Xb[1], z = Xb[2];
  const double D[18] = {0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1};
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 6; c++) J[6 * r + c] = A[3 * r] * D[c] + A[3 * r + 1] * D[6 + c] + A[3 * r + 2] * D[12 + c];
  if (mono) for (int c = 0; c < 6; c++) J[12 + c] = 0;
}
// EdgePriorPoseImu::computeError / linearizeOplus for the state k (pose, v, bg, ba)
static void prior_err(const GfsPoseInertialProblem& P, const KF& k, double* e15) {
  double Rt[9], M[9], d[3];
  mt3(P.c_Rwb, Rt);
  mm3(Rt, k.Rwb, M);
  log_so3(M, e15);
  for (int i = 0; i < 3; i++) d[i] = k.twb[i] - P.c_twb[i];
  mv3(Rt, d, e15 + 3);
  for (int i = 0; i < 3; i++) { e15[6 + i] = k.vel[i] - P.c_vwb[i]; e15[9 + i] = k.bg[i] - P.c_bg[i]; e15[12 + i] = k.ba[i] - P.c_ba[i]; }
}
static void prior_jac(const GfsPoseInertialProblem& P, const KF& k, double* J /*15x15*/) {
  double Rt[9], M[9], er[3], iJ[9];
  mt3(P.c_Rwb, Rt);
  mm3(Rt, k.Rwb, M);
  log_so3(M, er);
  inv_right_jac(er, iJ);
  memset(J, 0, sizeof(double) * 225);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) { J[15 * r + c] = iJ[3 * r + c]; J[15 * (3 + r) + 3 + c] = M[3 * r + c]; }
  for (int i = 6; i < 15; i++) J[15 * i + i] = 1.0;
}
// Optimizer::Marginalize(H, 0, 14) on a 30x30 matrix -> the lower-right 15x15 block of the result
static void marginalize_first15(const double* H30, double* out15) {
  std::vector<double> Hb(225), e, V;
  for (int r = 0; r < 15; r++)
    for (int c = 0; c < 15; c++) Hb[15 * r + c] = H30[30 * r + c];
  // JacobiSVD of the (symmetric) block: singular values |lambda|, U = V sign(lambda)  ->  V diag(1/lambda) V^T
  jacobi_eig(Hb, 15, e, V);
  std::vector<double> inv(225, 0.0);
  for (int k = 0; k < 15; k++) {
    if (!(std::fabs(e[k]) > 1e-6)) continue;
    const double w = 1.0 / e[k];
    for (int r = 0; r < 15; r++)
      for (int c = 0; c < 15; c++) inv[15 * r + c] += V[15 * r + k] * w * V[15 * c + k];
  }
  // c* = Hcc - Hcb * invHb * Hbc   (Hcb = H[15:30, 0:15], Hbc = H[0:15, 15:30])
  std::vector<double> T(225);
  for (int r = 0; r < 15; r++)
    for (int c = 0; c < 15; c++) {
      double a = 0;
      for (int k = 0; k < 15; k++) a += H30[30 * (15 + r) + k] * inv[15 * k + c];
      T[15 * r + c] = a;
    }
  for (int r = 0; r < 15; r++)
    for (int c = 0; c < 15; c++) {
      double a = 0;
      for (int k = 0; k < 15; k++) a += T[15 * r + k] * H30[30 * k + 15 + c];
      out15[15 * r + c] = H30[30 * (15 + r) + 15 + c] - a;
    }
}
// ConstraintPoseImu constructor: H = (H + H) / 2 (sic), eigenvalues < 1e-12 zeroed
static void constraint_clamp(double* H15) {
  std::vector<double> A(H15, H15 + 225), e, V;
  for (auto& v : A) v = (v + v) / 2;
  // SelfAdjointEigenSolver reads the lower triangle only
  for (int r = 0; r < 15; r++)
    for (int c = r + 1; c < 15; c++) A[15 * r + c] = A[15 * c + r];
  jacobi_eig(A, 15, e, V);
  for (int i = 0; i < 15; i++)
    if (e[i] < 1e-12) e[i] = 0;
  for (int r = 0; r < 15; r++)
    for (int c = 0; c < 15; c++) {
      double a = 0;
      for (int k = 0; k < 15; k++) a += V[15 * r + k] * e[k] * V[15 * c + k];
      H15[15 * r + c] = a;
    }
}

struct Ctx {
  const GfsPoseInertialProblem* P;
  Cam C;
  int n, dim;
  bool free_prev;
  KF cur, prev;
  double infoI[81], infoG[9], infoA[9];
  double deltaI, deltaMono, deltaStereo;
  std::vector<double> err;     // [n][3] stored _error of the visual edges
  std::vector<int> level;
  std::vector<char> robust;
  double errI[9], errG[3], errA[3], errP[15];
};
static void compute_active_errors(Ctx& X) {
  const GfsPoseInertialProblem& P = *X.P;
  for (int e = 0; e < X.n; e++)
    if (X.level[e] == 0) vis_err(X.C, X.cur, P.Xw + 3 * (size_t)e, P.uvr + 3 * (size_t)e, &X.err[3 * (size_t)e], nullptr);
  inertial_error_core(Pre(P.pre), X.prev, X.cur, X.errI);
  for (int i = 0; i < 3; i++) { X.errG[i] = X.cur.bg[i] - X.prev.bg[i]; X.errA[i] = X.cur.ba[i] - X.prev.ba[i]; }
  if (X.free_prev) prior_err(P, X.prev, X.errP);
}
static double quad(const double* e, const double* Om, int n) {  // e^T Om e as g2o's chi2(): _error.dot(information()*_error)
  double s = 0;
  for (int r = 0; r < n; r++) {
    double a = 0;
    for (int c = 0; c < n; c++) a += Om[n * r + c] * e[c];
    s += e[r] * a;
  }
  return s;
}
// unknown order = vertex ids: VP 0-5, VV 6-8, VG 9-11, VA 12-14, then (LastFrame) VPk 15-20, VVk 21-23, VGk 24-26, VAk 27-29
static void add_block(std::vector<double>& H, std::vector<double>& b, int dim, const double* J, int rows, int jcols, const int* map,
                      const double* Om, double w, const double* e) {
  // H += J^T (w Om) J,  b -= w J^T Om e   over the mapped (non-fixed) columns
  std::vector<double> WJ((size_t)rows * jcols), We(rows);
  for (int r = 0; r < rows; r++) {
    for (int c = 0; c < jcols; c++) {
      double a = 0;
      for (int k = 0; k < rows; k++) a += (w * Om[rows * r + k]) * J[jcols * k + c];
      WJ[(size_t)r * jcols + c] = a;
    }
    double a = 0;
    for (int k = 0; k < rows; k++) a += Om[rows * r + k] * e[k];
    We[r] = w * a;
  }
  for (int a = 0; a < jcols; a++) {
    if (map[a] < 0) continue;
    double g = 0;
    for (int r = 0; r < rows; r++) g += J[jcols * r + a] * We[r];
    b[map[a]] -= g;
    for (int c = 0; c < jcols; c++) {
      if (map[c] < 0) continue;
      double h = 0;
      for (int r = 0; r < rows; r++) h += J[jcols * r + a] * WJ[(size_t)r * jcols + c];
      H[(size_t)map[a] * dim + map[c]] += h;
    }
  }
}
static void build_system(Ctx& X, std::vector<double>& H, std::vector<double>& b) {
  const GfsPoseInertialProblem& P = *X.P;
  const int dim = X.dim;
  H.assign((size_t)dim * dim, 0.0);
  b.assign(dim, 0.0);
  for (int e = 0; e < X.n; e++) {
    if (X.level[e] != 0) continue;
    const bool mono = P.uvr[3 * (size_t)e + 2] < 0;
    const int d = mono ? 2 : 3;
    double J[18];
    vis_jac(X.C, X.cur, P.Xw + 3 * (size_t)e, mono, J);
    const double om = (double)P.inv_sigma2[e];
    const double* er = &X.err[3 * (size_t)e];
    double w = 1.0;
    if (X.robust[e]) {
      double c2 = 0;
      for (int i = 0; i < d; i++) c2 += er[i] * (om * er[i]);
      double rho[3];
      huber(c2, mono ? X.deltaMono : X.deltaStereo, rho);
      w = rho[1];
    }
    for (int a = 0; a < 6; a++) {
      double g = 0;
      for (int i = 0; i < d; i++) g += J[6 * i + a] * (om * er[i]);
      b[a] -= w * g;
      for (int c = 0; c < 6; c++) {
        double h = 0;
        for (int i = 0; i < d; i++) h += J[6 * i + a] * ((w * om) * J[6 * i + c]);
        H[(size_t)a * dim + c] += h;
      }
    }
  }
  {  // EdgeInertial: J columns [p1 v1 bg1 ba1 p2 v2]
    double J[9 * 24];
    inertial_jacobian_core(Pre(P.pre), X.prev, X.cur, J);
    int map[24];
    for (int c = 0; c < 15; c++) map[c] = X.free_prev ? 15 + c : -1;
    for (int c = 0; c < 9; c++) map[15 + c] = c;
    double rho[3];
    huber(quad(X.errI, X.infoI, 9), X.deltaI, rho);
    add_block(H, b, dim, J, 9, 24, map, X.infoI, rho[1], X.errI);
  }
  {  // EdgeGyroRW / EdgeAccRW: J = [-I, I] over (prev bias, cur bias), no kernel
    double J[18] = {-1, 0, 0, 1, 0, 0, 0, -1, 0, 0, 1, 0, 0, 0, -1, 0, 0, 1};
    int mg[6], ma[6];
    for (int c = 0; c < 3; c++) { mg[c] = X.free_prev ? 24 + c : -1; mg[3 + c] = 9 + c; ma[c] = X.free_prev ? 27 + c : -1; ma[3 + c] = 12 + c; }
    add_block(H, b, dim, J, 3, 6, mg, X.infoG, 1.0, X.errG);
    add_block(H, b, dim, J, 3, 6, ma, X.infoA, 1.0, X.errA);
  }
  if (X.free_prev) {  // EdgePriorPoseImu, Huber 5
    double J[225];
    prior_jac(P, X.prev, J);
    int map[15];
    for (int c = 0; c < 15; c++) map[c] = 15 + c;
    double rho[3];
    huber(quad(X.errP, P.c_H, 15), 5.0, rho);
    add_block(H, b, dim, J, 15, 15, map, P.c_H, rho[1], X.errP);
  }
}
static void apply_update(Ctx& X, const double* x) {
  cam_update(X.C, X.cur, x);
  for (int i = 0; i < 3; i++) { X.cur.vel[i] += x[6 + i]; X.cur.bg[i] += x[9 + i]; X.cur.ba[i] += x[12 + i]; }
  if (X.free_prev) {
    cam_update(X.C, X.prev, x + 15);
    for (int i = 0; i < 3; i++) { X.prev.vel[i] += x[21 + i]; X.prev.bg[i] += x[24 + i]; X.prev.ba[i] += x[27 + i]; }
  }
}

static void optimize(const GfsPoseInertialProblem* Pp, GfsPoseInertialResult* R) {
  const GfsPoseInertialProblem& P = *Pp;
  Ctx X;
  X.P = Pp;
  X.C = Cam{(double)P.fx, (double)P.fy, (double)P.cx, (double)P.cy, (double)P.bf, P.Rcb, P.tcb, P.tbc};
  X.n = P.n_obs;
  X.free_prev = P.mode == GFS_PIN_LAST_FRAME;
  X.dim = X.free_prev ? 30 : 15;
  memcpy(X.cur.Rwb, P.Rwb, 72); memcpy(X.cur.twb, P.twb, 24); memcpy(X.cur.Rcw, P.Rcw, 72); memcpy(X.cur.tcw, P.tcw, 24);
  memcpy(X.cur.vel, P.vel, 24); memcpy(X.cur.bg, P.bg, 24); memcpy(X.cur.ba, P.ba, 24);
  memset(&X.prev, 0, sizeof(KF));
  memcpy(X.prev.Rwb, P.p_Rwb, 72); memcpy(X.prev.twb, P.p_twb, 24);
  memcpy(X.prev.vel, P.p_vel, 24); memcpy(X.prev.bg, P.p_bg, 24); memcpy(X.prev.ba, P.p_ba, 24);
  inertial_information(Pre(P.pre).C, X.infoI);
  {
    double Cg[9], Ca[9];
    for (int i = 0; i < 9; i++) { Cg[i] = (double)P.rw_Cg[i]; Ca[i] = (double)P.rw_Ca[i]; }
    inv3(Cg, X.infoG);
    inv3(Ca, X.infoA);
  }
  X.deltaI = 6.0;
  X.deltaMono = (double)(float)std::sqrt(5.991);
  X.deltaStereo = (double)(float)std::sqrt(7.815);
  const int N = X.n;
  X.err.assign((size_t)3 * N + 3, 0.0);
  X.level.assign(N, 0);
  X.robust.assign(N, 1);
  for (int e = 0; e < N; e++) { R->outlier[e] = 0; if (R->chi2) R->chi2[e] = 0.f; }
  memset(R->gn_iterations, 0, sizeof(R->gn_iterations));
  R->rounds_done = 0;
  const bool lastKF = !X.free_prev;
  const float chi2MonoKF[4] = {12, 7.5, 5.991, 5.991}, chi2MonoF[4] = {5.991, 5.991, 5.991, 5.991};
  const float chi2Stereo[4] = {15.6f, 9.8f, 7.815f, 7.815f};
  const float* chi2Mono = lastKF ? chi2MonoKF : chi2MonoF;
  int nBad = 0, nInliers = 0;
  float avg = 0.f;
  std::vector<double> H, b, x(X.dim, 0.0);
  const int rounds = std::min(std::max(P.n_rounds, 0), 4);
  for (int it = 0; it < rounds; it++) {
    // optimizer.initializeOptimization(0); optimizer.optimize(10)
    for (int iter = 0; iter < 10; iter++) {
      compute_active_errors(X);
      build_system(X, H, b);
      const bool ok = gfo_eigen_ldlt_solve(H.data(), b.data(), X.dim, x.data()) != 0;  // a failure leaves x as it was
      apply_update(X, x.data());
      R->gn_iterations[it]++;
      if (!ok) break;
    }
    nBad = 0;
    int nInMono = 0, nInStereo = 0;
    const float chi2close = 1.5 * chi2Mono[it];
    avg = 0.0f;
    if (lastKF) {
      const float chi2IMU = (float)quad(X.errI, X.infoI, 9);
      if (chi2IMU > 15.0) {
        X.deltaI = 2.0;
        for (int i = 0; i < 81; i++) X.infoI[i] *= 1e-2;
      }
    }
    for (int pass = 0; pass < 2; pass++)
      for (int e = 0; e < N; e++) {
        const float* o = P.uvr + 3 * (size_t)e;
        const bool mono = o[2] < 0;
        if (mono != (pass == 0)) continue;
        double* er = &X.err[3 * (size_t)e];
        if (R->outlier[e]) vis_err(X.C, X.cur, P.Xw + 3 * (size_t)e, o, er, nullptr);
        const int d = mono ? 2 : 3;
        const double om = (double)P.inv_sigma2[e];
        double c2 = 0;
        for (int i = 0; i < d; i++) c2 += er[i] * (om * er[i]);
        const float chi2 = (float)c2;
        if (R->chi2) R->chi2[e] = chi2;
        bool bad;
        if (mono) {
          const bool bClose = P.close[e] != 0;
          const double* Xw = P.Xw + 3 * (size_t)e;
          const bool depthPos = (X.cur.Rcw[6] * Xw[0] + X.cur.Rcw[7] * Xw[1] + X.cur.Rcw[8] * Xw[2] + X.cur.tcw[2]) > 0.0;
          bad = (chi2 > chi2Mono[it] && !bClose) || (bClose && chi2 > chi2close) || !depthPos;
        } else {
          bad = chi2 > chi2Stereo[it];
        }
        if (bad) { R->outlier[e] = 1; X.level[e] = 1; nBad++; }
        else { avg += chi2; R->outlier[e] = 0; X.level[e] = 0; (mono ? nInMono : nInStereo)++; }
        if (it == 2) X.robust[e] = 0;
      }
    nInliers = nInMono + nInStereo;
    avg /= nInliers;
    R->rounds_done = it + 1;
    if (N + (X.free_prev ? 4 : 3) < 10) break;  // optimizer.edges().size() < 10
  }
  // If not too much tracks, recover not too bad points
  if (nInliers < 30 && !P.rec_init) {
    nBad = 0;
    for (int pass = 0; pass < 2; pass++)
      for (int e = 0; e < N; e++) {
        const float* o = P.uvr + 3 * (size_t)e;
        const bool mono = o[2] < 0;
        if (mono != (pass == 0)) continue;
        double* er = &X.err[3 * (size_t)e];
        vis_err(X.C, X.cur, P.Xw + 3 * (size_t)e, o, er, nullptr);
        const int d = mono ? 2 : 3;
        const double om = (double)P.inv_sigma2[e];
        double c2 = 0;
        for (int i = 0; i < d; i++) c2 += er[i] * (om * er[i]);
        if (c2 < (mono ? 18.0 : 24.0)) R->outlier[e] = 0;
        else nBad++;
      }
  }
  memcpy(R->Rwb, X.cur.Rwb, 72); memcpy(R->twb, X.cur.twb, 24); memcpy(R->vel, X.cur.vel, 24);
  memcpy(R->bg, X.cur.bg, 24); memcpy(R->ba, X.cur.ba, 24);
  // ---- Hessian hand-over (fresh linearisation at the final estimates, raw information, inlier edges only)
  double Hv[36] = {0};
  for (int e = 0; e < N; e++) {
    if (R->outlier[e]) continue;
    const bool mono = P.uvr[3 * (size_t)e + 2] < 0;
    const int d = mono ? 2 : 3;
    double J[18];
    vis_jac(X.C, X.cur, P.Xw + 3 * (size_t)e, mono, J);
    const double om = (double)P.inv_sigma2[e];
    for (int a = 0; a < 6; a++)
      for (int c = 0; c < 6; c++) {
        double h = 0;
        for (int i = 0; i < d; i++) h += J[6 * i + a] * (om * J[6 * i + c]);
        Hv[6 * a + c] += h;
      }
  }
  double JI[9 * 24];
  inertial_jacobian_core(Pre(P.pre), X.prev, X.cur, JI);
  std::vector<double> HI(24 * 24, 0.0), dummyb(24, 0.0);
  {
    int map[24];
    for (int c = 0; c < 24; c++) map[c] = c;
    double zero[9] = {0};
    add_block(HI, dummyb, 24, JI, 9, 24, map, X.infoI, 1.0, zero);  // J^T information() J
  }
  double Hout[225];
  if (lastKF) {
    memset(Hout, 0, sizeof(Hout));
    for (int r = 0; r < 9; r++)
      for (int c = 0; c < 9; c++) Hout[15 * r + c] += HI[24 * (15 + r) + 15 + c];  // GetHessian2: [p2 v2]
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) { Hout[15 * (9 + r) + 9 + c] += X.infoG[3 * r + c]; Hout[15 * (12 + r) + 12 + c] += X.infoA[3 * r + c]; }
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < 6; c++) Hout[15 * r + c] += Hv[6 * r + c];
  } else {
    std::vector<double> H30(900, 0.0);
    for (int r = 0; r < 24; r++)
      for (int c = 0; c < 24; c++) H30[30 * r + c] += HI[24 * r + c];
    const int gi[2] = {9, 24}, ai[2] = {12, 27};
    const double sg[2] = {-1.0, 1.0};
    for (int A = 0; A < 2; A++)
      for (int B = 0; B < 2; B++)
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) {
            H30[30 * (gi[A] + r) + gi[B] + c] += sg[A] * sg[B] * X.infoG[3 * r + c];
            H30[30 * (ai[A] + r) + ai[B] + c] += sg[A] * sg[B] * X.infoA[3 * r + c];
          }
    {
      double J[225];
      prior_jac(P, X.prev, J);
      std::vector<double> HP(225, 0.0), db(15, 0.0);
      int map[15];
      for (int c = 0; c < 15; c++) map[c] = c;
      double zero[15] = {0};
      add_block(HP, db, 15, J, 15, 15, map, P.c_H, 1.0, zero);
      for (int r = 0; r < 15; r++)
        for (int c = 0; c < 15; c++) H30[30 * r + c] += HP[15 * r + c];
    }
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < 6; c++) H30[30 * (15 + r) + 15 + c] += Hv[6 * r + c];
    marginalize_first15(H30.data(), Hout);
  }
  constraint_clamp(Hout);
  memcpy(R->H, Hout, sizeof(Hout));
  R->n_bad = nBad;
  R->n_inliers_last = nInliers;
  R->avg_reproj_error = avg;
  R->n_inliers = N - nBad;
}

}  // namespace pin
}  // namespace ba
}  // namespace gfo

extern "C" {

void gfo_ba_solve(const GfsBaProblem* P, GfsBaResult* R) { gfo::ba::solve(P, R); }

// ---- stage hooks for the oracle's own validation
// robust chi2 of the problem at its initial state
double gfo_ba_chi2(const GfsBaProblem* P) {
  gfo::ba::Problem pr;
  gfo::ba::State s;
  gfo::ba::setup(P, pr, s);
  return gfo::ba::robust_chi2(pr, s, nullptr);
}
// 9-D inertial error and 9x24 Jacobian of edge e at the problem's initial state
void gfo_ba_inertial(const GfsBaProblem* P, int e, double* err9, double* J216, double* info81) {
  gfo::ba::Problem pr;
  gfo::ba::State s;
  gfo::ba::setup(P, pr, s);
  gfo::ba::inertial_error(pr, s, e, err9);
  gfo::ba::inertial_jacobian(pr, s, e, J216);
  memcpy(info81, &pr.infoIn[(size_t)e * 81], 81 * sizeof(double));
}
// normal equations at the initial state: Hpp (dimP x dimP), bp, Hll (nPt x 9), bl (nPt x 3), Hpl (nObs x 18)
void gfo_ba_system(const GfsBaProblem* P, double* Hpp, double* bp, double* Hll, double* bl, double* Hpl) {
  gfo::ba::Problem pr;
  gfo::ba::State s;
  gfo::ba::System S;
  gfo::ba::setup(P, pr, s);
  gfo::ba::build_system(pr, s, S);
  memcpy(Hpp, S.Hpp.data(), S.Hpp.size() * 8); memcpy(bp, S.bp.data(), S.bp.size() * 8);
  memcpy(Hll, S.Hll.data(), S.Hll.size() * 8); memcpy(bl, S.bl.data(), S.bl.size() * 8);
  memcpy(Hpl, S.Hpl.data(), S.Hpl.size() * 8);
}
// one damped Schur solve at the initial state: x (dimP + 3 nPt); returns 1 on success
int gfo_ba_step(const GfsBaProblem* P, double lambda, double* x) {
  gfo::ba::Problem pr;
  gfo::ba::State s;
  gfo::ba::System S;
  gfo::ba::setup(P, pr, s);
  gfo::ba::build_system(pr, s, S);
  std::vector<double> xv;
  const bool ok = gfo::ba::schur_solve(pr, S, lambda, xv);
  if (ok) memcpy(x, xv.data(), xv.size() * 8);
  return ok;
}
// apply an update vector to the initial state and return the robust chi2 there (for finite differences)
double gfo_ba_chi2_at(const GfsBaProblem* P, const double* x, int robust) {
  gfo::ba::Problem pr;
  gfo::ba::State s;
  gfo::ba::setup(P, pr, s);
  std::vector<double> xv(x, x + pr.dimP + 3 * (size_t)pr.nPt);
  gfo::ba::apply_update(pr, s, xv);
  if (!robust) { pr.deltaMono = pr.deltaStereo = pr.deltaIn = 1e300; }
  return gfo::ba::robust_chi2(pr, s, nullptr);
}
// SE3Quat(R, t).log() and EdgeICP error for two camera poses (validation hooks)
void gfo_se3quat_log(const double* R, const double* t, double* out6) { gfo::ba::se3q_log(gfo::ba::se3q_make(R, t), out6); }
void gfo_icp_error(const double* Rt12, const double* Rc1w, const double* tc1w, const double* Rc2w, const double* tc2w, double* out6) {
  gfo::ba::KF a, b;
  memcpy(a.Rcw, Rc1w, 72); memcpy(a.tcw, tc1w, 24); memcpy(b.Rcw, Rc2w, 72); memcpy(b.tcw, tc2w, 24);
  gfo::ba::icp_error_kf(Rt12, a, b, out6);
}
void gfo_so3(const double* w, double* R_exp, double* log_of_exp, double* Jr, double* Jrinv) {
  gfo::ba::exp_so3(w, R_exp);
  gfo::ba::log_so3(R_exp, log_of_exp);
  gfo::ba::right_jac(w, Jr);
  gfo::ba::inv_right_jac(w, Jrinv);
}

void gfo_pose_inertial_optimize(const GfsPoseInertialProblem* P, GfsPoseInertialResult* R) { gfo::ba::pin::optimize(P, R); }
// test hooks: visual edge (error, 3x6 Jacobian) at a body pose; prior edge; marginalisation; clamp
int gfo_pin_vis_edge(const GfsPoseInertialProblem* P, const double* Rwb, const double* twb, int e, double* err3, double* J18) {
  using namespace gfo::ba;
  pin::Cam C{(double)P->fx, (double)P->fy, (double)P->cx, (double)P->cy, (double)P->bf, P->Rcb, P->tcb, P->tbc};
  KF k;
  memset(&k, 0, sizeof(k));
  memcpy(k.Rwb, Rwb, 72); memcpy(k.twb, twb, 24);
  const double zero[6] = {0, 0, 0, 0, 0, 0};
  pin::cam_update(C, k, zero);  // Rcw / tcw from Rwb / twb
  const bool mono = P->uvr[3 * (size_t)e + 2] < 0;
  const int d = pin::vis_err(C, k, P->Xw + 3 * (size_t)e, P->uvr + 3 * (size_t)e, err3, nullptr);
  pin::vis_jac(C, k, P->Xw + 3 * (size_t)e, mono, J18);
  return d;
}
void gfo_pin_pose_update(const GfsPoseInertialProblem* P, const double* Rwb, const double* twb, const double* u6, double* Rwb_out, double* twb_out) {
  using namespace gfo::ba;
  pin::Cam C{(double)P->fx, (double)P->fy, (double)P->cx, (double)P->cy, (double)P->bf, P->Rcb, P->tcb, P->tbc};
  KF k;
  memset(&k, 0, sizeof(k));
  memcpy(k.Rwb, Rwb, 72); memcpy(k.twb, twb, 24);
  pin::cam_update(C, k, u6);
  memcpy(Rwb_out, k.Rwb, 72); memcpy(twb_out, k.twb, 24);
}
void gfo_pin_prior_edge(const GfsPoseInertialProblem* P, const double* Rwb, const double* twb, const double* v, const double* bg,
                        const double* ba, double* err15, double* J225) {
  using namespace gfo::ba;
  KF k;
  memset(&k, 0, sizeof(k));
  memcpy(k.Rwb, Rwb, 72); memcpy(k.twb, twb, 24); memcpy(k.vel, v, 24); memcpy(k.bg, bg, 24); memcpy(k.ba, ba, 24);
  pin::prior_err(*P, k, err15);
  pin::prior_jac(*P, k, J225);
}
void gfo_pin_marginalize(const double* H30, double* out15) { gfo::ba::pin::marginalize_first15(H30, out15); }
void gfo_pin_clamp(double* H15) { gfo::ba::pin::constraint_clamp(H15); }
}

// ORACLE -- TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of Optimizer::PoseOptimization (reference src/Optimizer.cc:763-1099) on a
// flattened problem (include/gfs_b200.h GfsPoseProblem): the motion-only bundle adjustment the
// tracking thread runs every frame.
//   vertex          g2o::VertexSE3Expmap::oplusImpl   Thirdparty/g2o/g2o/types/types_six_dof_expmap.h:55-70
//                   SE3Quat ctor / operator* / map / exp  Thirdparty/g2o/g2o/types/se3quat.h:65-75,101-110,223-257
//   mono edge       EdgeSE3ProjectXYZOnlyPose          include/OptimizableTypes.h:30-60, src/OptimizableTypes.cpp:27-41
//                   Pinhole::project / projectJac       src/CameraModels/Pinhole.cpp:35-41,71-81
//   stereo edge     g2o::EdgeStereoSE3ProjectXYZOnlyPose Thirdparty/g2o/g2o/types/types_six_dof_expmap.cpp:339-346,375-404
//                   (cam_project narrows 1/z to float, as the source does)
//   g2o             Huber (robust_kernel_impl.cpp:77-91), constructQuadraticForm (base_unary_edge.hpp),
//                   Levenberg with tau = 1e-5 (optimization_algorithm_levenberg.cpp:59-190), optimize loop
//                   (sparse_optimizer.cpp:354-420), LinearSolverDense = Eigen::LDLT (linear_solver_dense.h:60-115)
//   outer logic     4 rounds x 10 iterations from the frame's pose, chi2 classification with float
//                   thresholds, levels, kernel removal after round 2, cumulative nGood (Optimizer.cc:955-1075)
// written without Eigen/g2o.  The fork does not write the optimised pose back to the frame (the
// SetPose calls are commented out, :1086-1097): the outputs are mvbOutlier, the reprojection-error
// average and the inlier count; the pose estimate of the last round is returned for inspection.
// The reference has no tests for this path: parity unpinned; tests/test_oracle_pose.py checks the
// Jacobians by finite differences and the optimisation against ground truth.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <vector>

#include "../include/gfs_b200.h"

namespace gfo {
namespace pose {

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
// SE3Quat::exp (se3quat.h:223-257)
static SE3Q se3q_exp(const double* u) {
  const double w[3] = {u[0], u[1], u[2]}, ups[3] = {u[3], u[4], u[5]};
  const double theta = std::sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double O2[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) O2[3 * r + c] = O[3 * r] * O[c] + O[3 * r + 1] * O[3 + c] + O[3 * r + 2] * O[6 + c];
  double R[9], V[9];
  if (theta < 0.00001) {
    for (int i = 0; i < 9; i++) { R[i] = ((i % 4 == 0) ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
  } else {
    const double a = std::sin(theta) / theta, b = (1 - std::cos(theta)) / (theta * theta),
                 c = (theta - std::sin(theta)) / std::pow(theta, 3);
    for (int i = 0; i < 9; i++) {
      const double I = (i % 4 == 0) ? 1.0 : 0.0;
      R[i] = I + a * O[i] + b * O2[i];
      V[i] = I + b * O[i] + c * O2[i];
    }
  }
  SE3Q s;
  s.r = quat_from_R(R);
  quat_normalize_rot(s.r);
  for (int r = 0; r < 3; r++) s.t[r] = V[3 * r] * ups[0] + V[3 * r + 1] * ups[1] + V[3 * r + 2] * ups[2];
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

static inline void huber(double e, double delta, double* rho) {
  const double dsqr = delta * delta;
  if (e <= dsqr) { rho[0] = e; rho[1] = 1.; rho[2] = 0.; }
  else { const double sq = std::sqrt(e); rho[0] = 2 * sq * delta - dsqr; rho[1] = delta / sq; rho[2] = -0.5 * rho[1] / e; }
}

struct Prob {
  int n;
  double fx, fy, cx, cy, bf;  // float members widened
  const double* Xw;
  const float* uvr;
  const float* is2;
  double deltaMono, deltaStereo;
};

// error of edge e at pose T; returns the dimension (2 mono, 3 stereo)
static int edge_error(const Prob& P, const SE3Q& T, int e, double* err, double* xc_out) {
  double xc[3];
  quat_rot(T.r, P.Xw + 3 * (size_t)e, xc);
  for (int i = 0; i < 3; i++) xc[i] += T.t[i];
  if (xc_out) memcpy(xc_out, xc, 24);
  const float* o = P.uvr + 3 * (size_t)e;
  if (o[2] < 0) {
    err[0] = (double)o[0] - (P.fx * xc[0] / xc[2] + P.cx);
    err[1] = (double)o[1] - (P.fy * xc[1] / xc[2] + P.cy);
    return 2;
  }
  const float invz = (float)(1.0 / xc[2]);  // `const float invz = 1.0f/trans_xyz[2];`
  const double u = xc[0] * (double)invz * P.fx + P.cx;
  err[0] = (double)o[0] - u;
  err[1] = (double)o[1] - (xc[1] * (double)invz * P.fy + P.cy);
  err[2] = (double)o[2] - (u - P.bf * (double)invz);
  return 3;
}
static void edge_jacobian(const Prob& P, const SE3Q& T, int e, double* J /*dim x 6*/) {
  double xc[3];
  quat_rot(T.r, P.Xw + 3 * (size_t)e, xc);
  for (int i = 0; i < 3; i++) xc[i] += T.t[i];
  const double x = xc[0], y = xc[1], z = xc[2];
  if (P.uvr[3 * (size_t)e + 2] < 0) {
    // -projectJac(xyz) * SE3deriv
    const double pj[6] = {P.fx / z, 0.0, -P.fx * x / (z * z), 0.0, P.fy / z, -P.fy * y / (z * z)};
    const double D[18] = {0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1};
    for (int r = 0; r < 2; r++)
      for (int c = 0; c < 6; c++) J[6 * r + c] = -(pj[3 * r] * D[c] + pj[3 * r + 1] * D[6 + c] + pj[3 * r + 2] * D[12 + c]);
    return;
  }
  const double invz = 1.0 / z, invz_2 = invz * invz;
  J[0] = x * y * invz_2 * P.fx; J[1] = -(1 + (x * x * invz_2)) * P.fx; J[2] = y * invz * P.fx;
  J[3] = -invz * P.fx; J[4] = 0; J[5] = x * invz_2 * P.fx;
  J[6] = (1 + y * y * invz_2) * P.fy; J[7] = -x * y * invz_2 * P.fy; J[8] = -x * invz * P.fy;
  J[9] = 0; J[10] = -invz * P.fy; J[11] = y * invz_2 * P.fy;
  J[12] = J[0] - P.bf * y * invz_2; J[13] = J[1] + P.bf * x * invz_2; J[14] = J[2];
  J[15] = J[3]; J[16] = 0; J[17] = J[5] - P.bf * invz_2;
}

// Eigen::LDLT<MatrixXd>::compute + solve on a dense symmetric n x n system (unblocked, pivoted; n = 6)
static bool eigen_ldlt_solve(const double* Hin, const double* b, int n, double* x) {
  std::vector<double> m(Hin, Hin + (size_t)n * n);
  std::vector<int> tr(n);
  std::vector<double> temp(n);
  enum { PosSemi, NegSemi, Zero, Indef } sign = Zero;
  bool found_zero_pivot = false, ret = true;
  auto M = [&](int r, int c) -> double& { return m[(size_t)r * n + c]; };
  for (int k = 0; k < n; k++) {
    int big = k;
    double bv = std::fabs(M(k, k));
    for (int i = k + 1; i < n; i++)
      if (std::fabs(M(i, i)) > bv) { bv = std::fabs(M(i, i)); big = i; }
    tr[k] = big;
    if (k != big) {
      const int s = n - big - 1;
      for (int c = 0; c < k; c++) std::swap(M(k, c), M(big, c));
      for (int r = 0; r < s; r++) std::swap(M(big + 1 + r, k), M(big + 1 + r, big));
      std::swap(M(k, k), M(big, big));
      for (int i = k + 1; i < big; i++) std::swap(M(i, k), M(big, i));
    }
    const int rs = n - k - 1;
    if (k > 0) {
      for (int c = 0; c < k; c++) temp[c] = M(c, c) * M(k, c);
      double acc = 0;
      for (int c = 0; c < k; c++) acc += M(k, c) * temp[c];
      M(k, k) -= acc;
      for (int r = 0; r < rs; r++) {
        double a2 = 0;
        for (int c = 0; c < k; c++) a2 += M(k + 1 + r, c) * temp[c];
        M(k + 1 + r, k) -= a2;
      }
    }
    const double akk = M(k, k);
    const bool valid = std::fabs(akk) > 0;
    if (k == 0 && !valid) { sign = Zero; for (int j = 0; j < n; j++) tr[j] = j; break; }
    if (rs > 0 && valid) for (int r = 0; r < rs; r++) M(k + 1 + r, k) /= akk;
    else if (rs > 0) for (int r = 0; r < rs; r++) ret = ret && (M(k + 1 + r, k) == 0);
    if (found_zero_pivot && valid) ret = false;
    else if (!valid) found_zero_pivot = true;
    if (sign == PosSemi) { if (akk < 0) sign = Indef; }
    else if (sign == NegSemi) { if (akk > 0) sign = Indef; }
    else if (sign == Zero) { if (akk > 0) sign = PosSemi; else if (akk < 0) sign = NegSemi; }
  }
  (void)ret;
  if (!(sign == PosSemi || sign == Zero)) return false;  // _cholesky.isPositive()
  std::vector<double> y(b, b + n);
  for (int k = 0; k < n; k++) std::swap(y[k], y[tr[k]]);
  for (int i = 0; i < n; i++)
    for (int c = 0; c < i; c++) y[i] -= M(i, c) * y[c];
  const double tol = 1.0 / std::numeric_limits<double>::max();
  for (int i = 0; i < n; i++) y[i] = (std::fabs(M(i, i)) > tol) ? y[i] / M(i, i) : 0.0;
  for (int i = n - 1; i >= 0; i--)
    for (int c = i + 1; c < n; c++) y[i] -= M(c, i) * y[c];
  for (int k = n - 1; k >= 0; k--) std::swap(y[k], y[tr[k]]);
  memcpy(x, y.data(), sizeof(double) * n);
  return true;
}

struct Edge {
  double err[3];
  int level;
  bool robust;
};

static double robust_chi2_active(const Prob& P, const SE3Q& T, std::vector<Edge>& E, bool recompute) {
  double sum = 0;
  for (int e = 0; e < P.n; e++) {
    if (E[e].level != 0) continue;
    const bool mono = P.uvr[3 * (size_t)e + 2] < 0;
    if (recompute) edge_error(P, T, e, E[e].err, nullptr);
    const int d = mono ? 2 : 3;
    double c2 = 0;
    for (int i = 0; i < d; i++) c2 += E[e].err[i] * ((double)P.is2[e] * E[e].err[i]);
    if (E[e].robust) {
      double rho[3];
      huber(c2, mono ? P.deltaMono : P.deltaStereo, rho);
      sum += rho[0];
    } else {
      sum += c2;
    }
  }
  return sum;
}
static void build_system(const Prob& P, const SE3Q& T, const std::vector<Edge>& E, double* H, double* b) {
  memset(H, 0, 36 * 8);
  memset(b, 0, 6 * 8);
  for (int e = 0; e < P.n; e++) {
    if (E[e].level != 0) continue;
    const bool mono = P.uvr[3 * (size_t)e + 2] < 0;
    const int d = mono ? 2 : 3;
    double J[18];
    edge_jacobian(P, T, e, J);
    const double om = (double)P.is2[e];
    double w = 1.0;  // rho[1]
    if (E[e].robust) {
      double c2 = 0;
      for (int i = 0; i < d; i++) c2 += E[e].err[i] * (om * E[e].err[i]);
      double rho[3];
      huber(c2, mono ? P.deltaMono : P.deltaStereo, rho);
      w = rho[1];
    }
    for (int a = 0; a < 6; a++) {
      double g = 0;
      for (int i = 0; i < d; i++) g += J[6 * i + a] * (om * E[e].err[i]);
      b[a] -= w * g;
      for (int c = 0; c < 6; c++) {
        double h = 0;
        for (int i = 0; i < d; i++) h += J[6 * i + a] * ((w * om) * J[6 * i + c]);
        H[6 * a + c] += h;
      }
    }
  }
}

static void optimize(const GfsPoseProblem* pb, GfsPoseResult* R) {
  Prob P;
  P.n = pb->n_obs;
  P.fx = pb->fx; P.fy = pb->fy; P.cx = pb->cx; P.cy = pb->cy; P.bf = pb->bf;
  P.Xw = pb->Xw; P.uvr = pb->uvr; P.is2 = pb->inv_sigma2;
  P.deltaMono = (double)(float)std::sqrt(5.991);
  P.deltaStereo = (double)(float)std::sqrt(7.815);
  memset(R->lm_iterations, 0, sizeof(R->lm_iterations));
  R->rounds_done = 0; R->n_bad = 0; R->n_good = 0; R->avg_reproj_error = 0.f; R->n_inliers = 0;
  SE3Q T0;
  T0.r = Quat{(double)pb->q_wxyz[0], (double)pb->q_wxyz[1], (double)pb->q_wxyz[2], (double)pb->q_wxyz[3]};
  quat_normalize_rot(T0.r);
  for (int i = 0; i < 3; i++) T0.t[i] = (double)pb->t[i];
  SE3Q T = T0;
  auto write_pose = [&]() {
    R->q_wxyz[0] = T.r.w; R->q_wxyz[1] = T.r.x; R->q_wxyz[2] = T.r.y; R->q_wxyz[3] = T.r.z;
    memcpy(R->t, T.t, 24);
  };
  write_pose();
  const int N = P.n;
  for (int e = 0; e < N; e++) { R->outlier[e] = 0; if (R->chi2) R->chi2[e] = 0.f; }
  if (N < 3) return;  // nInitialCorrespondences < 3 -> return 0
  std::vector<Edge> E(N);
  for (int e = 0; e < N; e++) { E[e].level = 0; E[e].robust = true; E[e].err[0] = E[e].err[1] = E[e].err[2] = 0; }
  const float chi2Mono[4] = {5.991f, 5.991f, 5.991f, 5.991f}, chi2Stereo[4] = {7.815f, 7.815f, 7.815f, 7.815f};
  int nBad = 0, nGood = 0;
  for (int it = 0; it < 4; it++) {
    T = T0;
    int nActive = 0;
    for (int e = 0; e < N; e++) nActive += E[e].level == 0;
    if (nActive > 0) {
      // optimizer.optimize(10)
      double lambda = 0, ni = 2;
      int nBadLM = 0;
      double H[36], b[6], x[6];
      for (int iter = 0; iter < 10; iter++) {
        double currentChi = robust_chi2_active(P, T, E, true);
        const double iniChi = currentChi;
        double tempChi = currentChi;
        build_system(P, T, E, H, b);
        if (iter == 0) {
          double md = 0;
          for (int j = 0; j < 6; j++) md = std::max(std::fabs(H[7 * j]), md);
          lambda = 1e-5 * md;
          ni = 2;
          nBadLM = 0;
        }
        double rho = 0;
        int qmax = 0;
        do {
          const SE3Q backup = T;
          double Hl[36];
          memcpy(Hl, H, sizeof(Hl));
          for (int j = 0; j < 6; j++) Hl[7 * j] += lambda;
          const bool ok2 = eigen_ldlt_solve(Hl, b, 6, x);
          if (!ok2) memset(x, 0, sizeof(x));  // LinearSolverDense leaves x untouched; block solver zeroed it before
          T = se3q_mul(se3q_exp(x), T);
          tempChi = robust_chi2_active(P, T, E, true);
          if (!ok2) tempChi = std::numeric_limits<double>::max();
          rho = currentChi - tempChi;
          double scale = 0;
          for (int j = 0; j < 6; j++) scale += x[j] * (lambda * x[j] + b[j]);
          scale += 1e-3;
          rho /= scale;
          if (rho > 0 && std::isfinite(tempChi)) {
            double alpha = 1. - std::pow((2 * rho - 1), 3);
            alpha = std::min(alpha, 2. / 3.);
            lambda *= std::max(1. / 3., alpha);
            ni = 2;
            currentChi = tempChi;
          } else {
            lambda *= ni;
            ni *= 2;
            T = backup;  // pop(): the edges keep the errors of the rejected trial
          }
          qmax++;
        } while (rho < 0 && qmax < 10);
        R->lm_iterations[it]++;
        if (qmax == 10 || rho == 0) break;
        if ((iniChi - currentChi) * 1e3 < iniChi) nBadLM++;
        else nBadLM = 0;
        if (nBadLM >= 3) break;
      }
    }
    // classification (Optimizer.cc:968-1062): mono edges first, then stereo, each in frame order
    nBad = 0;
    float avg = 0.0f;
    for (int pass = 0; pass < 2; pass++)
      for (int e = 0; e < N; e++) {
        const bool mono = P.uvr[3 * (size_t)e + 2] < 0;
        if (mono != (pass == 0)) continue;
        if (R->outlier[e]) edge_error(P, T, e, E[e].err, nullptr);
        const int d = mono ? 2 : 3;
        double c2d = 0;
        for (int i = 0; i < d; i++) c2d += E[e].err[i] * ((double)P.is2[e] * E[e].err[i]);
        const float chi2 = (float)c2d;
        if (R->chi2) R->chi2[e] = chi2;
        if (chi2 > (mono ? chi2Mono[it] : chi2Stereo[it])) {
          R->outlier[e] = 1;
          E[e].level = 1;
          nBad++;
        } else {
          avg += chi2;
          R->outlier[e] = 0;
          E[e].level = 0;
          nGood++;
        }
        if (it == 2) E[e].robust = false;
      }
    avg /= (float)nGood;
    R->avg_reproj_error = avg;
    R->rounds_done = it + 1;
    write_pose();
    if (N < 10) break;
  }
  R->n_bad = nBad;
  R->n_good = nGood;
  R->n_inliers = N - nBad;
}

}  // namespace pose
}  // namespace gfo

extern "C" {
void gfo_pose_optimize(const GfsPoseProblem* P, GfsPoseResult* R) { gfo::pose::optimize(P, R); }
// test hooks: error (returns dim) and Jacobian of one observation at pose (q_wxyz, t)
int gfo_pose_edge(const GfsPoseProblem* pb, const double* q_wxyz, const double* t, int e, double* err3, double* J18) {
  gfo::pose::Prob P;
  P.n = pb->n_obs;
  P.fx = pb->fx; P.fy = pb->fy; P.cx = pb->cx; P.cy = pb->cy; P.bf = pb->bf;
  P.Xw = pb->Xw; P.uvr = pb->uvr; P.is2 = pb->inv_sigma2;
  gfo::pose::SE3Q T;
  T.r = gfo::pose::Quat{q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]};
  memcpy(T.t, t, 24);
  const int d = gfo::pose::edge_error(P, T, e, err3, nullptr);
  gfo::pose::edge_jacobian(P, T, e, J18);
  return d;
}
// exp(u) * (q, t) -> (q_out, t_out): the vertex update
void gfo_pose_oplus(const double* q_wxyz, const double* t, const double* u6, double* q_out, double* t_out) {
  gfo::pose::SE3Q T;
  T.r = gfo::pose::Quat{q_wxyz[0], q_wxyz[1], q_wxyz[2], q_wxyz[3]};
  memcpy(T.t, t, 24);
  const gfo::pose::SE3Q r = gfo::pose::se3q_mul(gfo::pose::se3q_exp(u6), T);
  q_out[0] = r.r.w; q_out[1] = r.r.x; q_out[2] = r.r.y; q_out[3] = r.r.z;
  memcpy(t_out, r.t, 24);
}
int gfo_eigen_ldlt_solve(const double* H, const double* b, int n, double* x) { return gfo::pose::eigen_ldlt_solve(H, b, n, x) ? 1 : 0; }
}

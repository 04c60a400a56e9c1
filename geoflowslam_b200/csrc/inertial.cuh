// Device and host helpers shared by the inertial optimisers (ba.cu: LocalInertialBA, pose_inertial.cu:
// PoseInertialOptimizationLast{KeyFrame,Frame}): SO(3) maps as src/G2oTypes.cc:1011-1082 writes them, the float32
// preintegration read side (src/ImuTypes.cc:283-313), EdgeInertial::computeError / linearizeOplus
// (src/G2oTypes.cc:495-719) on two body-state records, and the EdgeInertial information matrix
// (G2oTypes.cc:487-494) on the host.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <vector>

#include "common.cuh"

namespace gfs {

// body-state record: Rwb9 twb3 Rcw9 tcw3 vel3 bg3 ba3
static const int KF_STRIDE = 33;
enum { K_RWB = 0, K_TWB = 9, K_RCW = 12, K_TCW = 21, K_VEL = 24, K_BG = 27, K_BA = 30 };

__device__ __forceinline__ void mm3(const double* A, const double* B, double* C) {
  double t[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) t[3 * r + c] = A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c] + A[3 * r + 2] * B[6 + c];
#pragma unroll
  for (int i = 0; i < 9; i++) C[i] = t[i];
}
__device__ __forceinline__ void mt3(const double* A, double* T) {
  double t[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) t[3 * r + c] = A[3 * c + r];
#pragma unroll
  for (int i = 0; i < 9; i++) T[i] = t[i];
}
__device__ __forceinline__ void mv3(const double* A, const double* v, double* o) {
  double t[3];
#pragma unroll
  for (int r = 0; r < 3; r++) t[r] = A[3 * r] * v[0] + A[3 * r + 1] * v[1] + A[3 * r + 2] * v[2];
  o[0] = t[0]; o[1] = t[1]; o[2] = t[2];
}
__device__ __forceinline__ bool inv3(const double* A, double* I) {
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
  const double id = 1.0 / det;
  double t[9];
  t[0] = c00 * id; t[3] = c01 * id; t[6] = c02 * id;
  t[1] = (A[2] * A[7] - A[1] * A[8]) * id; t[4] = (A[0] * A[8] - A[2] * A[6]) * id; t[7] = (A[1] * A[6] - A[0] * A[7]) * id;
  t[2] = (A[1] * A[5] - A[2] * A[4]) * id; t[5] = (A[2] * A[3] - A[0] * A[5]) * id; t[8] = (A[0] * A[4] - A[1] * A[3]) * id;
#pragma unroll
  for (int i = 0; i < 9; i++) I[i] = t[i];
  return det != 0 && isfinite(id);
}
__device__ __forceinline__ void skew3(const double* w, double* W) {
  W[0] = 0; W[1] = -w[2]; W[2] = w[1]; W[3] = w[2]; W[4] = 0; W[5] = -w[0]; W[6] = -w[1]; W[7] = w[0]; W[8] = 0;
}
// NormalizeRotation (G2oTypes.h:69-74) = orthogonal polar factor, by Newton iteration
template <class T>
__device__ void normalize_rotation(T* R) {
  for (int it = 0; it < 20; it++) {
    double A[9], I[9];
    for (int i = 0; i < 9; i++) A[i] = (double)R[i];
    if (!inv3(A, I)) return;
    double diff = 0;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        const T n = (T)(0.5 * (A[3 * r + c] + I[3 * c + r]));
        diff = fmax(diff, fabs((double)n - (double)R[3 * r + c]));
        R[3 * r + c] = n;
      }
    if (diff < (sizeof(T) == 4 ? 1e-7 : 1e-15)) break;
  }
}
__device__ inline void exp_so3(const double* w, double* R) {  // G2oTypes.cc:1011-1025
  const double d2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], d = sqrt(d2);
  double W[9], W2[9];
  skew3(w, W);
  mm3(W, W, W2);
  if (d < 1e-5) {
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + W[i] + 0.5 * W2[i];
  } else {
    const double a = sin(d) / d, b = (1.0 - cos(d)) / d2;
    for (int i = 0; i < 9; i++) R[i] = (i % 4 == 0 ? 1.0 : 0.0) + W[i] * a + W2[i] * b;
  }
  normalize_rotation(R);
}
__device__ inline void log_so3(const double* R, double* w) {  // :1027-1041
  const double tr = R[0] + R[4] + R[8];
  w[0] = (R[7] - R[5]) / 2; w[1] = (R[2] - R[6]) / 2; w[2] = (R[3] - R[1]) / 2;
  const double costheta = (tr - 1.0) * 0.5f;
  if (costheta > 1 || costheta < -1) return;
  const double theta = acos(costheta), s = sin(theta);
  if (fabs(s) < 1e-5) return;
  for (int i = 0; i < 3; i++) w[i] = theta * w[i] / s;
}
__device__ inline void inv_right_jac(const double* v, double* J) {  // :1047-1061
  const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = sqrt(d2);
  double W[9], W2[9];
  skew3(v, W);
  mm3(W, W, W2);
  if (d < 1e-5) { for (int i = 0; i < 9; i++) J[i] = (i % 4 == 0); return; }
  const double k = 1.0 / d2 - (1.0 + cos(d)) / (2.0 * d * sin(d));
  for (int i = 0; i < 9; i++) J[i] = (i % 4 == 0 ? 1.0 : 0.0) + W[i] / 2 + W2[i] * k;
}
__device__ inline void right_jac(const double* v, double* J) {  // :1067-1082
  const double d2 = v[0] * v[0] + v[1] * v[1] + v[2] * v[2], d = sqrt(d2);
  double W[9], W2[9];
  skew3(v, W);
  mm3(W, W, W2);
  if (d < 1e-5) { for (int i = 0; i < 9; i++) J[i] = (i % 4 == 0); return; }
  const double a = (1.0 - cos(d)) / d2, b = (d - sin(d)) / (d2 * d);
  for (int i = 0; i < 9; i++) J[i] = (i % 4 == 0 ? 1.0 : 0.0) - W[i] * a + W2[i] * b;
}
__device__ __forceinline__ void huber(double e, double delta, double* rho) {  // robust_kernel_impl.cpp:77-91
  const double dsqr = delta * delta;
  if (e <= dsqr) { rho[0] = e; rho[1] = 1.; }
  else { const double sq = sqrt(e); rho[0] = 2 * sq * delta - dsqr; rho[1] = delta / sq; }
}

// float32 preintegration read side (ImuTypes.cc:283-313); record layout GFS_BA_PRE_STRIDE
__device__ inline void so3f_exp(const float* w, float* R) {
  const float th2 = __fadd_rn(__fadd_rn(__fmul_rn(w[0], w[0]), __fmul_rn(w[1], w[1])), __fmul_rn(w[2], w[2]));
  float imag, real;
  if (th2 < 1e-5f * 1e-5f) {
    const float th4 = th2 * th2;
    imag = 0.5f - (1.0f / 48.0f) * th2 + (1.0f / 3840.0f) * th4;
    real = 1.0f - (1.0f / 8.0f) * th2 + (1.0f / 384.0f) * th4;
  } else {
    const float th = sqrtf(th2), half = 0.5f * th;
    imag = sinf(half) / th;
    real = cosf(half);
  }
  const float qw = real, qx = imag * w[0], qy = imag * w[1], qz = imag * w[2];
  const float tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const float twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy,
              tyz = tz * qy, tzz = tz * qz;
  R[0] = 1 - (tyy + tzz); R[1] = txy - twz; R[2] = txz + twy;
  R[3] = txy + twz; R[4] = 1 - (txx + tzz); R[5] = tyz - twx;
  R[6] = txz - twy; R[7] = tyz + twx; R[8] = 1 - (txx + tyy);
}
__device__ inline void delta_for_bias(const float* pre, const double* bg, const double* ba, double* dR, double* dV, double* dP,
                               double* dbg_out) {
  const float *pdR = pre, *pdV = pre + 9, *pdP = pre + 12, *JRg = pre + 15, *JVg = pre + 24, *JVa = pre + 33, *JPg = pre + 42,
              *JPa = pre + 51, *b = pre + 286;
  const float dbg[3] = {(float)bg[0] - b[3], (float)bg[1] - b[4], (float)bg[2] - b[5]};
  const float dba[3] = {(float)ba[0] - b[0], (float)ba[1] - b[1], (float)ba[2] - b[2]};
  float w[3], E[9], R[9];
  for (int r = 0; r < 3; r++) w[r] = JRg[3 * r] * dbg[0] + JRg[3 * r + 1] * dbg[1] + JRg[3 * r + 2] * dbg[2];
  so3f_exp(w, E);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R[3 * r + c] = pdR[3 * r] * E[c] + pdR[3 * r + 1] * E[3 + c] + pdR[3 * r + 2] * E[6 + c];
  normalize_rotation(R);
  for (int i = 0; i < 9; i++) dR[i] = (double)R[i];
  for (int r = 0; r < 3; r++) {
    const float v = pdV[r] + (JVg[3 * r] * dbg[0] + JVg[3 * r + 1] * dbg[1] + JVg[3 * r + 2] * dbg[2]) +
                    (JVa[3 * r] * dba[0] + JVa[3 * r + 1] * dba[1] + JVa[3 * r + 2] * dba[2]);
    const float q = pdP[r] + (JPg[3 * r] * dbg[0] + JPg[3 * r + 1] * dbg[1] + JPg[3 * r + 2] * dbg[2]) +
                    (JPa[3 * r] * dba[0] + JPa[3 * r + 1] * dba[1] + JPa[3 * r + 2] * dba[2]);
    dV[r] = (double)v;
    dP[r] = (double)q;
  }
  if (dbg_out) for (int i = 0; i < 3; i++) dbg_out[i] = (double)dbg[i];
}


// EdgeInertial::computeError (G2oTypes.cc:495-522) between the body states s1 (previous) and s2; err9 out
__device__ inline void inertial_error_core(const double* s1, const double* s2, const float* pre, double* err9) {
  double dR[9], dV[3], dP[3];
  delta_for_bias(pre, s1 + K_BG, s1 + K_BA, dR, dV, dP, nullptr);
  const double dt = (double)pre[285];
  const double g[3] = {0, 0, -(double)9.81f};
  double dRt[9], Rbw1[9], A[9], eR[9];
  mt3(dR, dRt);
  mt3(s1 + K_RWB, Rbw1);
  mm3(dRt, Rbw1, A);
  mm3(A, s2 + K_RWB, eR);
  log_so3(eR, err9);
  double tv[3], tp[3], rv[3], rp[3];
  for (int i = 0; i < 3; i++) {
    tv[i] = s2[K_VEL + i] - s1[K_VEL + i] - g[i] * dt;
    tp[i] = s2[K_TWB + i] - s1[K_TWB + i] - s1[K_VEL + i] * dt - g[i] * dt * dt / 2;
  }
  mv3(Rbw1, tv, rv);
  mv3(Rbw1, tp, rp);
  for (int i = 0; i < 3; i++) { err9[3 + i] = rv[i] - dV[i]; err9[6 + i] = rp[i] - dP[i]; }
}
// EdgeInertial::linearizeOplus (G2oTypes.cc:524-719): J is 9 x 24, columns [pose1(6) vel1(3) bg1(3) ba1(3) pose2(6) vel2(3)];
// r = the residual at the same states
__device__ inline void inertial_jacobian_core(const double* s1, const double* s2, const float* pre, double* J, double* r) {
  double dR[9], dV[3], dP[3], dbg[3];
  delta_for_bias(pre, s1 + K_BG, s1 + K_BA, dR, dV, dP, dbg);
  const double dt = (double)pre[285];
  const double g[3] = {0, 0, -(double)9.81f};
  double JRg[9], JVg[9], JVa[9], JPg[9], JPa[9];
  for (int i = 0; i < 9; i++) { JRg[i] = pre[15 + i]; JVg[i] = pre[24 + i]; JVa[i] = pre[33 + i]; JPg[i] = pre[42 + i]; JPa[i] = pre[51 + i]; }
  double Rbw1[9], dRt[9], A[9], eR[9], er[3], invJr[9];
  mt3(s1 + K_RWB, Rbw1);
  mt3(dR, dRt);
  mm3(dRt, Rbw1, A);
  mm3(A, s2 + K_RWB, eR);
  log_so3(eR, er);
  inv_right_jac(er, invJr);
  for (int i = 0; i < 216; i++) J[i] = 0;
  auto put = [&](int r0, int c0, const double* M, double sgn) {
    for (int rr = 0; rr < 3; rr++)
      for (int c = 0; c < 3; c++) J[(r0 + rr) * 24 + c0 + c] = sgn * M[3 * rr + c];
  };
  double Rwb2t[9], T1[9], T2[9];
  mt3(s2 + K_RWB, Rwb2t);
  mm3(invJr, Rwb2t, T1);
  mm3(T1, s1 + K_RWB, T2);
  put(0, 0, T2, -1.0);
  double tv[3], tp[3], v[3], W[9];
  for (int i = 0; i < 3; i++) {
    tv[i] = s2[K_VEL + i] - s1[K_VEL + i] - g[i] * dt;
    tp[i] = s2[K_TWB + i] - s1[K_TWB + i] - s1[K_VEL + i] * dt - 0.5 * g[i] * dt * dt;
  }
  mv3(Rbw1, tv, v); skew3(v, W); put(3, 0, W, 1.0);
  mv3(Rbw1, tp, v); skew3(v, W); put(6, 0, W, 1.0);
  const double I3[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
  put(6, 3, I3, -1.0);
  put(3, 6, Rbw1, -1.0);
  { double M[9]; for (int i = 0; i < 9; i++) M[i] = Rbw1[i] * dt; put(6, 6, M, -1.0); }
  {
    double w[3], Jr[9], eRt[9], M1[9], M2[9], M3[9];
    mv3(JRg, dbg, w);
    right_jac(w, Jr);
    mt3(eR, eRt);
    mm3(invJr, eRt, M1);
    mm3(M1, Jr, M2);
    mm3(M2, JRg, M3);
    put(0, 9, M3, -1.0);
    put(3, 9, JVg, -1.0);
    put(6, 9, JPg, -1.0);
  }
  put(3, 12, JVa, -1.0);
  put(6, 12, JPa, -1.0);
  put(0, 15, invJr, 1.0);
  { double M[9]; mm3(Rbw1, s2 + K_RWB, M); put(6, 18, M, 1.0); }
  put(3, 21, Rbw1, 1.0);
  // residual (same expressions as computeError)
  r[0] = er[0]; r[1] = er[1]; r[2] = er[2];
  double rv[3], rp[3];
  for (int i = 0; i < 3; i++) tp[i] = s2[K_TWB + i] - s1[K_TWB + i] - s1[K_VEL + i] * dt - g[i] * dt * dt / 2;
  mv3(Rbw1, tv, rv);
  mv3(Rbw1, tp, rp);
  for (int i = 0; i < 3; i++) { r[3 + i] = rv[i] - dV[i]; r[6 + i] = rp[i] - dP[i]; }
}

// ---- host side
inline bool invert_n(std::vector<double> A, int n, std::vector<double>& inv) {
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
inline void jacobi_eig(std::vector<double> A, int n, std::vector<double>& e, std::vector<double>& V) {
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
// EdgeInertial ctor (G2oTypes.cc:487-494)
inline void inertial_information(const float* C15, double* Info81) {
  std::vector<double> C(81), Ci, e, V, S(81);
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) C[9 * r + c] = (double)C15[15 * r + c];
  invert_n(C, 9, Ci);
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
inline void inv3_host(const double* A, double* I) {
  const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
  const double id = 1.0 / (A[0] * c00 + A[1] * c01 + A[2] * c02);
  I[0] = c00 * id; I[3] = c01 * id; I[6] = c02 * id;
  I[1] = (A[2] * A[7] - A[1] * A[8]) * id; I[4] = (A[0] * A[8] - A[2] * A[6]) * id; I[7] = (A[1] * A[6] - A[0] * A[7]) * id;
  I[2] = (A[1] * A[5] - A[2] * A[4]) * id; I[5] = (A[2] * A[3] - A[0] * A[5]) * id; I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
}


}  // namespace gfs

// Motion-only bundle adjustment of the tracking thread, batched over frames, sm_100a.
//
// Replaces the g2o block of Optimizer::PoseOptimization (reference src/Optimizer.cc:763-1099):
//   vertex   g2o::VertexSE3Expmap (SE3Quat exp / operator* / map, Thirdparty/g2o/g2o/types/se3quat.h)
//   edges    EdgeSE3ProjectXYZOnlyPose (src/OptimizableTypes.cpp:27-41, Pinhole.cpp:35-41,71-81) and
//            g2o::EdgeStereoSE3ProjectXYZOnlyPose (types_six_dof_expmap.cpp:339-346,375-404), Huber kernels
//   solver   OptimizationAlgorithmLevenberg, tau = 1e-5 (optimization_algorithm_levenberg.cpp:59-190),
//            LinearSolverDense = Eigen::LDLT on the 6x6 system (linear_solver_dense.h:60-115)
//   outer    4 rounds x 10 iterations from the frame's pose, chi2 classification, levels, kernels off
//            after the third round, cumulative nGood (Optimizer.cc:955-1075)
//
// One CTA per frame runs the whole optimisation in ONE launch: threads stride over the frame's
// observations (errors, Jacobians, the 21 + 6 entries of H and b in registers), fixed-order block sums,
// thread 0 does the 6x6 pivoted LDLT, the SE3 update and g2o's lambda / rho bookkeeping.  No fp atomics:
// results do not depend on scheduling.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "common.cuh"

namespace gfs {
namespace po {

static const int PO_THREADS = 128;

struct PoseHdr {
  int n, off;
  float q[4], t[3];
  float fx, fy, cx, cy, bf;
};
struct PoseOut {
  int n_inliers, n_bad, n_good;
  float avg;
  int rounds_done, lm_iterations[4];
  double q[4], t[3];
};

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
__device__ void quat_normalize_rot(Quat& q) {  // SE3Quat::normalizeRotation
  if (q.w < 0) { q.w = -q.w; q.x = -q.x; q.y = -q.y; q.z = -q.z; }
  const double n = sqrt(q.w * q.w + q.x * q.x + q.y * q.y + q.z * q.z);
  q.w /= n; q.x /= n; q.y /= n; q.z /= n;
}
__device__ Quat quat_mul(const Quat& a, const Quat& b) {
  return Quat{a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z, a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y,
              a.w * b.y + a.y * b.w + a.z * b.x - a.x * b.z, a.w * b.z + a.z * b.w + a.x * b.y - a.y * b.x};
}
__device__ __forceinline__ void quat_rot(const Quat& q, const double* v, double* o) {  // Eigen _transformVector
  const double ux = 2 * (q.y * v[2] - q.z * v[1]), uy = 2 * (q.z * v[0] - q.x * v[2]), uz = 2 * (q.x * v[1] - q.y * v[0]);
  o[0] = v[0] + q.w * ux + (q.y * uz - q.z * uy);
  o[1] = v[1] + q.w * uy + (q.z * ux - q.x * uz);
  o[2] = v[2] + q.w * uz + (q.x * uy - q.y * ux);
}
__device__ SE3Q se3q_exp(const double* u) {  // SE3Quat::exp, se3quat.h:223-257
  const double w[3] = {u[0], u[1], u[2]}, ups[3] = {u[3], u[4], u[5]};
  const double theta = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
  const double O[9] = {0, -w[2], w[1], w[2], 0, -w[0], -w[1], w[0], 0};
  double O2[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) O2[3 * r + c] = O[3 * r] * O[c] + O[3 * r + 1] * O[3 + c] + O[3 * r + 2] * O[6 + c];
  double R[9], V[9];
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
  for (int r = 0; r < 3; r++) s.t[r] = V[3 * r] * ups[0] + V[3 * r + 1] * ups[1] + V[3 * r + 2] * ups[2];
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
__device__ __forceinline__ void huber(double e, double delta, double* rho) {  // robust_kernel_impl.cpp:77-91
  const double dsqr = delta * delta;
  if (e <= dsqr) { rho[0] = e; rho[1] = 1.; }
  else { const double sq = sqrt(e); rho[0] = 2 * sq * delta - dsqr; rho[1] = delta / sq; }
}

struct PoseCam { double fx, fy, cx, cy, bf; };

// error of one observation at pose T -> dimension (2 monocular, 3 stereo)
__device__ __forceinline__ int edge_error(const PoseCam& C, const SE3Q& T, const double* Xw, const float* o, double* err,
                                          double* xc) {
  quat_rot(T.r, Xw, xc);
  for (int i = 0; i < 3; i++) xc[i] += T.t[i];
  if (o[2] < 0) {
    err[0] = (double)o[0] - (C.fx * xc[0] / xc[2] + C.cx);
    err[1] = (double)o[1] - (C.fy * xc[1] / xc[2] + C.cy);
    err[2] = 0;
    return 2;
  }
  const float invz = (float)(1.0 / xc[2]);  // the reference narrows: `const float invz = 1.0f/trans_xyz[2];`
  const double u = xc[0] * (double)invz * C.fx + C.cx;
  err[0] = (double)o[0] - u;
  err[1] = (double)o[1] - (xc[1] * (double)invz * C.fy + C.cy);
  err[2] = (double)o[2] - (u - C.bf * (double)invz);
  return 3;
}
__device__ __forceinline__ void edge_jacobian(const PoseCam& C, bool mono, const double* xc, double* J) {
  const double x = xc[0], y = xc[1], z = xc[2];
  if (mono) {
    const double pj[6] = {C.fx / z, 0.0, -C.fx * x / (z * z), 0.0, C.fy / z, -C.fy * y / (z * z)};
    const double D[18] = {0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1};
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
      for (int c = 0; c < 6; c++) J[6 * r + c] = -(pj[3 * r] * D[c] + pj[3 * r + 1] * D[6 + c] + pj[3 * r + 2] * D[12 + c]);
#pragma unroll
    for (int c = 0; c < 6; c++) J[12 + c] = 0;
    return;
  }
  const double invz = 1.0 / z, invz_2 = invz * invz;
  J[0] = x * y * invz_2 * C.fx; J[1] = -(1 + (x * x * invz_2)) * C.fx; J[2] = y * invz * C.fx;
  J[3] = -invz * C.fx; J[4] = 0; J[5] = x * invz_2 * C.fx;
  J[6] = (1 + y * y * invz_2) * C.fy; J[7] = -x * y * invz_2 * C.fy; J[8] = -x * invz * C.fy;
  J[9] = 0; J[10] = -invz * C.fy; J[11] = y * invz_2 * C.fy;
  J[12] = J[0] - C.bf * y * invz_2; J[13] = J[1] + C.bf * x * invz_2; J[14] = J[2];
  J[15] = J[3]; J[16] = 0; J[17] = J[5] - C.bf * invz_2;
}

// Eigen::LDLT<MatrixXd>::compute + solve, unblocked pivoted, n = 6
__device__ bool eigen_ldlt_solve6(const double* Hin, const double* b, double* x) {
  const int n = 6;
  double m[36], temp[6], y[6];
  int tr[6];
  for (int i = 0; i < 36; i++) m[i] = Hin[i];
  int sign = 2;  // 0 PosSemi, 1 NegSemi, 2 Zero, 3 Indefinite
#define M_(r, c) m[(r) * n + (c)]
  for (int k = 0; k < n; k++) {
    int big = k;
    double bv = fabs(M_(k, k));
    for (int i = k + 1; i < n; i++)
      if (fabs(M_(i, i)) > bv) { bv = fabs(M_(i, i)); big = i; }
    tr[k] = big;
    if (k != big) {
      const int s = n - big - 1;
      for (int c = 0; c < k; c++) { const double t = M_(k, c); M_(k, c) = M_(big, c); M_(big, c) = t; }
      for (int r = 0; r < s; r++) { const double t = M_(big + 1 + r, k); M_(big + 1 + r, k) = M_(big + 1 + r, big); M_(big + 1 + r, big) = t; }
      { const double t = M_(k, k); M_(k, k) = M_(big, big); M_(big, big) = t; }
      for (int i = k + 1; i < big; i++) { const double t = M_(i, k); M_(i, k) = M_(big, i); M_(big, i) = t; }
    }
    const int rs = n - k - 1;
    if (k > 0) {
      for (int c = 0; c < k; c++) temp[c] = M_(c, c) * M_(k, c);
      double acc = 0;
      for (int c = 0; c < k; c++) acc += M_(k, c) * temp[c];
      M_(k, k) -= acc;
      for (int r = 0; r < rs; r++) {
        double a2 = 0;
        for (int c = 0; c < k; c++) a2 += M_(k + 1 + r, c) * temp[c];
        M_(k + 1 + r, k) -= a2;
      }
    }
    const double akk = M_(k, k);
    const bool valid = fabs(akk) > 0;
    if (k == 0 && !valid) { sign = 2; for (int j = 0; j < n; j++) tr[j] = j; break; }
    if (rs > 0 && valid) for (int r = 0; r < rs; r++) M_(k + 1 + r, k) /= akk;
    if (sign == 0) { if (akk < 0) sign = 3; }
    else if (sign == 1) { if (akk > 0) sign = 3; }
    else if (sign == 2) { if (akk > 0) sign = 0; else if (akk < 0) sign = 1; }
  }
  if (!(sign == 0 || sign == 2)) return false;  // _cholesky.isPositive()
  for (int i = 0; i < n; i++) y[i] = b[i];
  for (int k = 0; k < n; k++) { const double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
  for (int i = 0; i < n; i++)
    for (int c = 0; c < i; c++) y[i] -= M_(i, c) * y[c];
  const double tol = 1.0 / DBL_MAX;
  for (int i = 0; i < n; i++) y[i] = (fabs(M_(i, i)) > tol) ? y[i] / M_(i, i) : 0.0;
  for (int i = n - 1; i >= 0; i--)
    for (int c = i + 1; c < n; c++) y[i] -= M_(c, i) * y[c];
  for (int k = n - 1; k >= 0; k--) { const double t = y[k]; y[k] = y[tr[k]]; y[tr[k]] = t; }
  for (int i = 0; i < n; i++) x[i] = y[i];
#undef M_
  return true;
}

// fixed-order block sum of NV doubles per thread: lanes (shuffle tree), then warps in order; result in out[]
// (shared), valid for every thread after the trailing barrier
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* s_part, double* out) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double x = v[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) s_part[warp * NV + k] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) {
    double x = 0;
#pragma unroll
    for (int w = 0; w < PO_THREADS / 32; w++) x += s_part[w * NV + threadIdx.x];
    out[threadIdx.x] = x;
  }
  __syncthreads();
}

struct PoseShared {
  SE3Q T, Tnew;
  double H[21], b[6], x[6];
  double sums[28];
  double part[(PO_THREADS / 32) * 28];
  double currentChi, lambda, ni;
  int cont, brk, ok2, nBadLM, qmax;
  int counts[2];
};

__global__ void __launch_bounds__(PO_THREADS) k_pose_opt(const PoseHdr* __restrict__ hdr, const double* __restrict__ gXw,
                                                         const float* __restrict__ guvr, const float* __restrict__ gis2,
                                                         double* __restrict__ gerr, uint8_t* __restrict__ glevel,
                                                         uint8_t* __restrict__ goutlier, float* __restrict__ gchi2,
                                                         PoseOut* __restrict__ gout) {
  __shared__ PoseShared S;
  const int p = blockIdx.x, tid = threadIdx.x;
  const PoseHdr h = hdr[p];
  const int n = h.n;
  const double* Xw = gXw + (size_t)h.off * 3;
  const float* uvr = guvr + (size_t)h.off * 3;
  const float* is2 = gis2 + h.off;
  double* err = gerr + (size_t)h.off * 3;
  uint8_t* level = glevel + h.off;
  uint8_t* outlier = goutlier + h.off;
  float* chi2 = gchi2 + h.off;
  PoseOut* out = gout + p;
  PoseCam C;
  C.fx = h.fx; C.fy = h.fy; C.cx = h.cx; C.cy = h.cy; C.bf = h.bf;
  const double deltaMono = (double)(float)sqrt(5.991), deltaStereo = (double)(float)sqrt(7.815);
  SE3Q T0;
  T0.r = Quat{(double)h.q[0], (double)h.q[1], (double)h.q[2], (double)h.q[3]};
  quat_normalize_rot(T0.r);
  for (int i = 0; i < 3; i++) T0.t[i] = (double)h.t[i];
  for (int e = tid; e < n; e += PO_THREADS) { level[e] = 0; outlier[e] = 0; chi2[e] = 0.f; }
  if (tid == 0) {
    out->n_inliers = 0; out->n_bad = 0; out->n_good = 0; out->avg = 0.f; out->rounds_done = 0;
    for (int i = 0; i < 4; i++) out->lm_iterations[i] = 0;
    out->q[0] = T0.r.w; out->q[1] = T0.r.x; out->q[2] = T0.r.y; out->q[3] = T0.r.z;
    for (int i = 0; i < 3; i++) out->t[i] = T0.t[i];
  }
  if (n < 3) return;  // nInitialCorrespondences < 3 -> return 0
  __syncthreads();
  int nGood = 0, nBad = 0;

  // sum of the (robustified) chi2 of the active edges at pose T; errors are stored (g2o keeps _error)
  auto active_chi = [&](const SE3Q& T, bool robust) {
    double v[1] = {0.0};
    for (int e = tid; e < n; e += PO_THREADS) {
      if (level[e] != 0) continue;
      double er[3], xc[3];
      const int d = edge_error(C, T, Xw + 3 * (size_t)e, uvr + 3 * (size_t)e, er, xc);
      err[3 * (size_t)e] = er[0]; err[3 * (size_t)e + 1] = er[1]; err[3 * (size_t)e + 2] = er[2];
      const double om = (double)is2[e];
      double c2 = er[0] * (om * er[0]) + er[1] * (om * er[1]);
      if (d == 3) c2 += er[2] * (om * er[2]);
      if (robust) {
        double rho[2];
        huber(c2, d == 2 ? deltaMono : deltaStereo, rho);
        v[0] += rho[0];
      } else {
        v[0] += c2;
      }
    }
    block_sum<1>(v, S.part, S.sums);
    return S.sums[0];
  };

  for (int it = 0; it < 4; it++) {
    const bool robust = it < 3;  // setRobustKernel(0) on every edge at the end of round 2
    if (tid == 0) S.T = T0;
    int act = 0;
    for (int e = tid; e < n; e += PO_THREADS) act += level[e] == 0;
    const int nActive = __syncthreads_count(act > 0) > 0 ? 1 : 0;  // also publishes S.T
    int lmIters = 0;
    if (nActive) {
      for (int iter = 0; iter < 10; iter++) {
        const SE3Q T = S.T;
        const double currentChi0 = active_chi(T, robust);
        // ---- buildSystem: H (upper 21) and b
        double acc[27];
#pragma unroll
        for (int k = 0; k < 27; k++) acc[k] = 0.0;
        for (int e = tid; e < n; e += PO_THREADS) {
          if (level[e] != 0) continue;
          const float* o = uvr + 3 * (size_t)e;
          const bool mono = o[2] < 0;
          double xc[3];
          quat_rot(T.r, Xw + 3 * (size_t)e, xc);
          for (int i = 0; i < 3; i++) xc[i] += T.t[i];
          double J[18];
          edge_jacobian(C, mono, xc, J);
          const double er[3] = {err[3 * (size_t)e], err[3 * (size_t)e + 1], mono ? 0.0 : err[3 * (size_t)e + 2]};
          const double om = (double)is2[e];
          double w = 1.0;
          if (robust) {
            double c2 = er[0] * (om * er[0]) + er[1] * (om * er[1]);
            if (!mono) c2 += er[2] * (om * er[2]);
            double rho[2];
            huber(c2, mono ? deltaMono : deltaStereo, rho);
            w = rho[1];
          }
          const double wo = w * om;
          int k = 0;
#pragma unroll
          for (int a = 0; a < 6; a++) {
#pragma unroll
            for (int c = a; c < 6; c++) acc[k++] += J[a] * (wo * J[c]) + J[6 + a] * (wo * J[6 + c]) + J[12 + a] * (wo * J[12 + c]);
          }
#pragma unroll
          for (int a = 0; a < 6; a++) acc[21 + a] -= w * (J[a] * (om * er[0]) + J[6 + a] * (om * er[1]) + J[12 + a] * (om * er[2]));
        }
        block_sum<27>(acc, S.part, S.sums);
        if (tid == 0) {
          for (int k = 0; k < 21; k++) S.H[k] = S.sums[k];
          for (int k = 0; k < 6; k++) S.b[k] = S.sums[21 + k];
          S.currentChi = currentChi0;
          if (iter == 0) {
            double md = 0;
            int k = 0;
            for (int a = 0; a < 6; a++) { md = fmax(fabs(S.H[k]), md); k += 6 - a; }
            S.lambda = 1e-5 * md;  // computeLambdaInit: tau * max |H_jj|
            S.ni = 2;
            S.nBadLM = 0;
          }
          S.qmax = 0;
        }
        __syncthreads();
        const double iniChi = currentChi0;
        // ---- Levenberg trials
        do {
          if (tid == 0) {
            double Hl[36];
            int k = 0;
            for (int a = 0; a < 6; a++)
              for (int c = a; c < 6; c++) { Hl[6 * a + c] = S.H[k]; Hl[6 * c + a] = S.H[k]; k++; }
            for (int j = 0; j < 6; j++) Hl[7 * j] += S.lambda;
            double x[6];
            const bool ok2 = eigen_ldlt_solve6(Hl, S.b, x);
            if (!ok2) for (int j = 0; j < 6; j++) x[j] = 0;
            for (int j = 0; j < 6; j++) S.x[j] = x[j];
            S.ok2 = ok2 ? 1 : 0;
            S.Tnew = se3q_mul(se3q_exp(x), S.T);
          }
          __syncthreads();
          const SE3Q Tn = S.Tnew;
          double tempChi = active_chi(Tn, robust);
          if (tid == 0) {
            const bool ok2 = S.ok2 != 0;
            if (!ok2) tempChi = DBL_MAX;
            double rho = S.currentChi - tempChi;
            double scale = 0;
            for (int j = 0; j < 6; j++) scale += S.x[j] * (S.lambda * S.x[j] + S.b[j]);
            scale += 1e-3;
            rho /= scale;
            if (rho > 0 && isfinite(tempChi)) {
              double alpha = 1. - pow((2 * rho - 1), 3.0);
              alpha = fmin(alpha, 2. / 3.);
              S.lambda *= fmax(1. / 3., alpha);
              S.ni = 2;
              S.currentChi = tempChi;
              S.T = S.Tnew;
            } else {
              S.lambda *= S.ni;
              S.ni *= 2;  // pop(): the pose is restored, the edges keep the errors of the rejected trial
            }
            S.qmax++;
            S.cont = (rho < 0 && S.qmax < 10) ? 1 : 0;
            // after the loop: Terminate / nBad bookkeeping (decided here, where rho is known)
            S.brk = 0;
            if (!S.cont) {
              if (S.qmax == 10 || rho == 0) S.brk = 1;
              else {
                if ((iniChi - S.currentChi) * 1e3 < iniChi) S.nBadLM++;
                else S.nBadLM = 0;
                if (S.nBadLM >= 3) S.brk = 1;
              }
            }
          }
          __syncthreads();
        } while (S.cont);
        lmIters++;
        if (S.brk) break;
      }
    }
    __syncthreads();
    // ---- classification (Optimizer.cc:968-1062)
    const SE3Q T = S.T;
    const float thMono = 5.991f, thStereo = 7.815f;
    int bad = 0, good = 0;
    for (int e = tid; e < n; e += PO_THREADS) {
      const float* o = uvr + 3 * (size_t)e;
      const bool mono = o[2] < 0;
      double er[3];
      if (outlier[e]) {
        double xc[3];
        edge_error(C, T, Xw + 3 * (size_t)e, o, er, xc);
        err[3 * (size_t)e] = er[0]; err[3 * (size_t)e + 1] = er[1]; err[3 * (size_t)e + 2] = er[2];
      } else {
        er[0] = err[3 * (size_t)e]; er[1] = err[3 * (size_t)e + 1]; er[2] = err[3 * (size_t)e + 2];
      }
      const double om = (double)is2[e];
      double c2 = er[0] * (om * er[0]) + er[1] * (om * er[1]);
      if (!mono) c2 += er[2] * (om * er[2]);
      const float c2f = (float)c2;
      chi2[e] = c2f;
      if (c2f > (mono ? thMono : thStereo)) { outlier[e] = 1; level[e] = 1; bad++; }
      else { outlier[e] = 0; level[e] = 0; good++; }
    }
    // block counts (integers: order-free)
    if (tid == 0) { S.counts[0] = 0; S.counts[1] = 0; }
    __syncthreads();
    atomicAdd(&S.counts[0], bad);
    atomicAdd(&S.counts[1], good);
    __syncthreads();
    nBad = S.counts[0];
    nGood += S.counts[1];
    if (tid == 0) {
      // avgReprojectionError: float sum, monocular edges first, then stereo, each in frame order
      float avg = 0.0f;
      for (int pass = 0; pass < 2; pass++)
        for (int e = 0; e < n; e++) {
          const bool mono = uvr[3 * (size_t)e + 2] < 0;
          if (mono != (pass == 0) || outlier[e]) continue;
          avg += chi2[e];
        }
      avg /= (float)nGood;
      out->avg = avg;
      out->rounds_done = it + 1;
      out->lm_iterations[it] = lmIters;
      out->q[0] = T.r.w; out->q[1] = T.r.x; out->q[2] = T.r.y; out->q[3] = T.r.z;
      for (int i = 0; i < 3; i++) out->t[i] = T.t[i];
      out->n_bad = nBad; out->n_good = nGood; out->n_inliers = n - nBad;
    }
    __syncthreads();
    if (n < 10) break;  // optimizer.edges().size() < 10
  }
}

}  // namespace po
}  // namespace gfs

using namespace gfs;
using namespace gfs::po;

struct GfsPose {
  int maxObs = 0, maxBatch = 0;
  DevBuf d_hdr, d_Xw, d_uvr, d_is2, d_err, d_level, d_outlier, d_chi2, d_out;
  PinnedBuf h_in, h_out;
  int launches = 0;
};

extern "C" {

int gfs_pose_create(int max_obs, int max_batch, GfsPose** out) {
  GFS_REQUIRE(out, GFS_ERR_INVALID, "out is null");
  *out = nullptr;
  GFS_REQUIRE(max_obs > 0 && max_batch > 0, GFS_ERR_INVALID, "bad capacity");
  int rc = gfs_device_check();
  if (rc) return rc;
  GfsPose* h = new GfsPose();
  h->maxObs = max_obs;
  h->maxBatch = max_batch;
  const size_t N = (size_t)max_obs * max_batch, B = max_batch;
  if ((rc = h->d_hdr.reserve(B * sizeof(PoseHdr))) || (rc = h->d_Xw.reserve(N * 24)) || (rc = h->d_uvr.reserve(N * 12)) ||
      (rc = h->d_is2.reserve(N * 4)) || (rc = h->d_err.reserve(N * 24)) || (rc = h->d_level.reserve(N)) ||
      (rc = h->d_outlier.reserve(N)) || (rc = h->d_chi2.reserve(N * 4)) || (rc = h->d_out.reserve(B * sizeof(PoseOut))) ||
      (rc = h->h_in.reserve(B * sizeof(PoseHdr) + N * (24 + 12 + 4))) || (rc = h->h_out.reserve(B * sizeof(PoseOut) + N * 5 + 16))) {
    gfs_pose_destroy(h);
    return rc;
  }
  *out = h;
  return GFS_OK;
}

int gfs_pose_destroy(GfsPose* h) {
  if (!h) return GFS_OK;
  DevBuf* d[] = {&h->d_hdr, &h->d_Xw, &h->d_uvr, &h->d_is2, &h->d_err, &h->d_level, &h->d_outlier, &h->d_chi2, &h->d_out};
  for (DevBuf* b : d) b->release();
  h->h_in.release();
  h->h_out.release();
  delete h;
  return GFS_OK;
}

int gfs_pose_last_launches(const GfsPose* h) { return h ? h->launches : GFS_ERR_INVALID; }

int gfs_pose_optimize_batch(GfsPose* h, void* stream, const GfsPoseProblem* problems, int batch, GfsPoseResult* results) {
  GFS_REQUIRE(h && problems && results, GFS_ERR_INVALID, "null argument");
  GFS_REQUIRE(batch > 0 && batch <= h->maxBatch, GFS_ERR_CAPACITY, "batch exceeds the handle's max_batch");
  size_t total = 0;
  for (int p = 0; p < batch; p++) {
    const GfsPoseProblem& P = problems[p];
    GFS_REQUIRE(P.n_obs >= 0 && P.n_obs <= h->maxObs, GFS_ERR_CAPACITY, "n_obs exceeds the handle's max_obs");
    GFS_REQUIRE(P.n_obs == 0 || (P.Xw && P.uvr && P.inv_sigma2), GFS_ERR_INVALID, "null observation arrays");
    GFS_REQUIRE(P.n_obs == 0 || results[p].outlier, GFS_ERR_INVALID, "null outlier output");
    total += (size_t)P.n_obs;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // ---- pack: headers | Xw | uvr | inv_sigma2 (one pinned staging buffer, four copies)
  uint8_t* hp = (uint8_t*)h->h_in.p;
  PoseHdr* hh = (PoseHdr*)hp;
  double* hX = (double*)(hp + (size_t)h->maxBatch * sizeof(PoseHdr));
  float* hU = (float*)((uint8_t*)hX + (size_t)h->maxObs * h->maxBatch * 24);
  float* hS = hU + (size_t)h->maxObs * h->maxBatch * 3;
  size_t off = 0;
  for (int p = 0; p < batch; p++) {
    const GfsPoseProblem& P = problems[p];
    PoseHdr& H = hh[p];
    H.n = P.n_obs;
    H.off = (int)off;
    memcpy(H.q, P.q_wxyz, 16);
    memcpy(H.t, P.t, 12);
    H.fx = P.fx; H.fy = P.fy; H.cx = P.cx; H.cy = P.cy; H.bf = P.bf;
    if (P.n_obs) {
      memcpy(hX + off * 3, P.Xw, (size_t)P.n_obs * 24);
      memcpy(hU + off * 3, P.uvr, (size_t)P.n_obs * 12);
      memcpy(hS + off, P.inv_sigma2, (size_t)P.n_obs * 4);
    }
    off += (size_t)P.n_obs;
  }
  GFS_CUDA(cudaMemcpyAsync(h->d_hdr.p, hh, (size_t)batch * sizeof(PoseHdr), cudaMemcpyHostToDevice, st));
  if (total) {
    GFS_CUDA(cudaMemcpyAsync(h->d_Xw.p, hX, total * 24, cudaMemcpyHostToDevice, st));
    GFS_CUDA(cudaMemcpyAsync(h->d_uvr.p, hU, total * 12, cudaMemcpyHostToDevice, st));
    GFS_CUDA(cudaMemcpyAsync(h->d_is2.p, hS, total * 4, cudaMemcpyHostToDevice, st));
  }
  k_pose_opt<<<batch, PO_THREADS, 0, st>>>((const PoseHdr*)h->d_hdr.p, (const double*)h->d_Xw.p, (const float*)h->d_uvr.p,
                                           (const float*)h->d_is2.p, (double*)h->d_err.p, (uint8_t*)h->d_level.p,
                                           (uint8_t*)h->d_outlier.p, (float*)h->d_chi2.p, (PoseOut*)h->d_out.p);
  GFS_CUDA(cudaGetLastError());
  h->launches = 1;
  uint8_t* op = (uint8_t*)h->h_out.p;
  PoseOut* ho = (PoseOut*)op;
  uint8_t* hOut = op + (size_t)h->maxBatch * sizeof(PoseOut);
  float* hChi = (float*)(hOut + (((size_t)h->maxObs * h->maxBatch + 3) & ~(size_t)3));
  GFS_CUDA(cudaMemcpyAsync(ho, h->d_out.p, (size_t)batch * sizeof(PoseOut), cudaMemcpyDeviceToHost, st));
  if (total) {
    GFS_CUDA(cudaMemcpyAsync(hOut, h->d_outlier.p, total, cudaMemcpyDeviceToHost, st));
    GFS_CUDA(cudaMemcpyAsync(hChi, h->d_chi2.p, total * 4, cudaMemcpyDeviceToHost, st));
  }
  GFS_CUDA(gfs::stream_wait(st));
  off = 0;
  for (int p = 0; p < batch; p++) {
    GfsPoseResult& R = results[p];
    const PoseOut& O = ho[p];
    R.n_inliers = O.n_inliers; R.n_bad = O.n_bad; R.n_good = O.n_good; R.avg_reproj_error = O.avg;
    R.rounds_done = O.rounds_done;
    memcpy(R.lm_iterations, O.lm_iterations, sizeof(R.lm_iterations));
    memcpy(R.q_wxyz, O.q, 32);
    memcpy(R.t, O.t, 24);
    const int n = problems[p].n_obs;
    if (n) {
      memcpy(R.outlier, hOut + off, (size_t)n);
      if (R.chi2) memcpy(R.chi2, hChi + off, (size_t)n * 4);
    }
    off += (size_t)n;
  }
  return GFS_OK;
}

int gfs_pose_optimize(GfsPose* h, void* stream, const GfsPoseProblem* problem, GfsPoseResult* result) {
  return gfs_pose_optimize_batch(h, stream, problem, 1, result);
}

}  // extern "C"

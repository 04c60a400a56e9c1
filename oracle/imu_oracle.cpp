// ORACLE -- TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of IMU::Preintegrated::Initialize + IntegrateNewMeasurement (reference src/ImuTypes.cc:163-246)
// with IMU::IntegratedRotation (:87-112), IMU::Calib::Set (:399-412) and NormalizeRotation (:35-39), float32 as the
// reference.  Output: the packed record the inertial edges read (GFS_BA_PRE_STRIDE floats:
// dR9 dV3 dP3 JRg9 JVg9 JVa9 JPg9 JPa9 C225 dT1 b6).  Eigen evaluates the expressions left to right
// ((0.5f*dR)*acc)*dt*dt ...; sin / cos / sqrt are the C (double) functions here -- src/ImuTypes.cc has no
// `using namespace std`, so `sin(d)` on a float promotes -- and their results narrow to float where they scale a
// float matrix.  NormalizeRotation is a JacobiSVD<Matrix3f> U V^T in the reference; the polar factor is unique, it
// is computed by Newton iteration here (agrees to float rounding).
// The reference has no tests for this path: parity unpinned; tests/test_oracle_imu.py checks it against the
// independent numpy restatement geoflowslam_b200.synth.preintegrate and against closed-form integrals.
#include <cmath>
#include <cstring>

namespace gfo {
namespace imu {

typedef float f;
static void mm(const f* A, const f* B, f* C) {  // 3x3 product, each entry (a0*b0 + a1*b1) + a2*b2
  f t[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) t[3 * r + c] = (A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c]) + A[3 * r + 2] * B[6 + c];
  memcpy(C, t, sizeof(t));
}
static void hat(const f* v, f* W) {
  W[0] = 0; W[1] = -v[2]; W[2] = v[1]; W[3] = v[2]; W[4] = 0; W[5] = -v[0]; W[6] = -v[1]; W[7] = v[0]; W[8] = 0;
}
static void polar(f* R) {  // NormalizeRotation: orthogonal polar factor
  for (int it = 0; it < 20; it++) {
    double A[9], I[9];
    for (int i = 0; i < 9; i++) A[i] = (double)R[i];
    const double c00 = A[4] * A[8] - A[5] * A[7], c01 = A[5] * A[6] - A[3] * A[8], c02 = A[3] * A[7] - A[4] * A[6];
    const double det = A[0] * c00 + A[1] * c01 + A[2] * c02;
    if (det == 0) return;
    const double id = 1.0 / det;
    I[0] = c00 * id; I[3] = c01 * id; I[6] = c02 * id;
    I[1] = (A[2] * A[7] - A[1] * A[8]) * id; I[4] = (A[0] * A[8] - A[2] * A[6]) * id; I[7] = (A[1] * A[6] - A[0] * A[7]) * id;
    I[2] = (A[1] * A[5] - A[2] * A[4]) * id; I[5] = (A[2] * A[3] - A[0] * A[5]) * id; I[8] = (A[0] * A[4] - A[1] * A[3]) * id;
    double diff = 0;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        const f n = (f)(0.5 * (A[3 * r + c] + I[3 * c + r]));
        diff = std::fmax(diff, std::fabs((double)n - (double)R[3 * r + c]));
        R[3 * r + c] = n;
      }
    if (diff < 1e-7) break;
  }
}

struct State {
  f dR[9], dV[3], dP[3], JRg[9], JVg[9], JVa[9], JPg[9], JPa[9], C[225], dT, b[6];
};

static void integrate(State& S, const f* acc_m, const f* w_m, f dt, const f* Nga, const f* NgaWalk) {
  const f acc[3] = {acc_m[0] - S.b[0], acc_m[1] - S.b[1], acc_m[2] - S.b[2]};
  f A[81], B[54];
  for (int i = 0; i < 81; i++) A[i] = (i % 10 == 0) ? 1.f : 0.f;
  for (int i = 0; i < 54; i++) B[i] = 0.f;
  // dP = dP + dV*dt + 0.5f*dR*acc*dt*dt ; dV = dV + dR*acc*dt
  f hR[9], hRa[3], Ra[3];
  for (int i = 0; i < 9; i++) hR[i] = 0.5f * S.dR[i];
  for (int r = 0; r < 3; r++) {
    hRa[r] = (hR[3 * r] * acc[0] + hR[3 * r + 1] * acc[1]) + hR[3 * r + 2] * acc[2];
    Ra[r] = (S.dR[3 * r] * acc[0] + S.dR[3 * r + 1] * acc[1]) + S.dR[3 * r + 2] * acc[2];
  }
  for (int r = 0; r < 3; r++) {
    S.dP[r] = (S.dP[r] + S.dV[r] * dt) + (hRa[r] * dt) * dt;
    S.dV[r] = S.dV[r] + Ra[r] * dt;
  }
  f Wacc[9];
  hat(acc, Wacc);
  f Rdt[9], nRdt[9], hRdt2[9], nhRdt2[9], T[9], T2[9];
  for (int i = 0; i < 9; i++) {
    Rdt[i] = S.dR[i] * dt;
    nRdt[i] = (-S.dR[i]) * dt;
    hRdt2[i] = (hR[i] * dt) * dt;
    nhRdt2[i] = ((-0.5f * S.dR[i]) * dt) * dt;
  }
  mm(nRdt, Wacc, T);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) A[9 * (3 + r) + c] = T[3 * r + c];
  mm(nhRdt2, Wacc, T);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) A[9 * (6 + r) + c] = T[3 * r + c];
  for (int r = 0; r < 3; r++) A[9 * (6 + r) + 3 + r] = dt;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) { B[6 * (3 + r) + 3 + c] = Rdt[3 * r + c]; B[6 * (6 + r) + 3 + c] = hRdt2[3 * r + c]; }
  // bias Jacobians (with the non-updated dR)
  for (int i = 0; i < 9; i++) S.JPa[i] = (S.JPa[i] + S.JVa[i] * dt) - hRdt2[i];
  mm(hRdt2, Wacc, T); mm(T, S.JRg, T2);
  for (int i = 0; i < 9; i++) S.JPg[i] = (S.JPg[i] + S.JVg[i] * dt) - T2[i];
  for (int i = 0; i < 9; i++) S.JVa[i] = S.JVa[i] - Rdt[i];
  mm(Rdt, Wacc, T); mm(T, S.JRg, T2);
  for (int i = 0; i < 9; i++) S.JVg[i] = S.JVg[i] - T2[i];
  // IntegratedRotation
  const f x = (w_m[0] - S.b[3]) * dt, y = (w_m[1] - S.b[4]) * dt, z = (w_m[2] - S.b[5]) * dt;
  const f d2 = (x * x + y * y) + z * z;
  const f d = (f)std::sqrt((double)d2);
  const f v[3] = {x, y, z};
  f W[9], WW[9], dRi[9], rJ[9];
  hat(v, W);
  mm(W, W, WW);
  if (d < 1e-4f) {
    for (int i = 0; i < 9; i++) { dRi[i] = ((i % 4 == 0) ? 1.f : 0.f) + W[i]; rJ[i] = (i % 4 == 0) ? 1.f : 0.f; }
  } else {
    const f s = (f)std::sin((double)d), omc = (f)(1.0 - std::cos((double)d)), dms = (f)((double)d - std::sin((double)d));
    const f d3 = d2 * d;
    for (int i = 0; i < 9; i++) {
      const f I = (i % 4 == 0) ? 1.f : 0.f;
      dRi[i] = (I + (W[i] * s) / d) + (WW[i] * omc) / d2;
      rJ[i] = (I - (W[i] * omc) / d2) + (WW[i] * dms) / d3;
    }
  }
  mm(S.dR, dRi, S.dR);
  polar(S.dR);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) { A[9 * r + c] = dRi[3 * c + r]; B[6 * r + c] = rJ[3 * r + c] * dt; }
  // C[0:9,0:9] = A C A^T + B Nga B^T (Nga diagonal)
  f AC[81], N[81];
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) {
      f a = 0;
      for (int k = 0; k < 9; k++) a += A[9 * r + k] * S.C[15 * k + c];
      AC[9 * r + c] = a;
    }
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) {
      f a = 0, bq = 0;
      for (int k = 0; k < 9; k++) a += AC[9 * r + k] * A[9 * c + k];
      for (int k = 0; k < 6; k++) bq += (B[6 * r + k] * Nga[k]) * B[6 * c + k];
      N[9 * r + c] = a + bq;
    }
  for (int r = 0; r < 9; r++)
    for (int c = 0; c < 9; c++) S.C[15 * r + c] = N[9 * r + c];
  for (int k = 0; k < 6; k++) S.C[15 * (9 + k) + 9 + k] += NgaWalk[k];
  // JRg = dRi^T * JRg - rightJ*dt
  f dRiT[9];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) dRiT[3 * r + c] = dRi[3 * c + r];
  mm(dRiT, S.JRg, T);
  for (int i = 0; i < 9; i++) S.JRg[i] = T[i] - rJ[i] * dt;
  S.dT += dt;
}

}  // namespace imu
}  // namespace gfo

extern "C" {
// meas: n rows of (ax, ay, az, wx, wy, wz, dt); bias: bax bay baz bwx bwy bwz; Calib::Set(ng, na, ngw, naw)
void gfo_imu_preintegrate(const float* meas, int n, const float* bias, float ng, float na, float ngw, float naw, float* out292) {
  using namespace gfo::imu;
  State S;
  memset(&S, 0, sizeof(S));
  S.dR[0] = S.dR[4] = S.dR[8] = 1.f;
  memcpy(S.b, bias, sizeof(S.b));
  const f Nga[6] = {ng * ng, ng * ng, ng * ng, na * na, na * na, na * na};
  const f Nw[6] = {ngw * ngw, ngw * ngw, ngw * ngw, naw * naw, naw * naw, naw * naw};
  for (int i = 0; i < n; i++) integrate(S, meas + 7 * i, meas + 7 * i + 3, meas[7 * i + 6], Nga, Nw);
  f* o = out292;
  memcpy(o, S.dR, 36); o += 9; memcpy(o, S.dV, 12); o += 3; memcpy(o, S.dP, 12); o += 3;
  memcpy(o, S.JRg, 36); o += 9; memcpy(o, S.JVg, 36); o += 9; memcpy(o, S.JVa, 36); o += 9;
  memcpy(o, S.JPg, 36); o += 9; memcpy(o, S.JPa, 36); o += 9; memcpy(o, S.C, 900); o += 225;
  *o++ = S.dT; memcpy(o, S.b, 24);
}
}

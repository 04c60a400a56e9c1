// Batched IMU preintegration on sm_100a.
//
// Replaces IMU::Preintegrated::Initialize + IntegrateNewMeasurement over an interval's measurements (reference
// src/ImuTypes.cc:163-246, with IMU::IntegratedRotation :87-112, IMU::Calib::Set :399-412, NormalizeRotation
// :35-39) for a batch of independent intervals -- what Tracking::PreintegrateIMU, Preintegrated::Reintegrate (after a
// bias update, one call per keyframe of the map) and MergePrevious run one interval at a time.  float32 as the
// reference; the expression order follows Eigen's left-to-right evaluation; sin / cos are the double functions whose
// results narrow to float (the reference calls the C functions on floats).  NormalizeRotation's JacobiSVD U V^T is
// the orthogonal polar factor, computed by Newton iteration (csrc/inertial.cuh normalize_rotation).
// One thread per interval: the recurrence is sequential in time and an interval is 7 (frame) to ~100 (keyframe)
// samples; the batch is what runs in parallel.  Output: the packed record the inertial edges read
// (GFS_BA_PRE_STRIDE floats: dR9 dV3 dP3 JRg9 JVg9 JVa9 JPg9 JPa9 C225 dT1 b6).
#include "common.cuh"
#include "inertial.cuh"

namespace gfs {
namespace imu {

__device__ __forceinline__ void mm3f(const float* A, const float* B, float* C) {
  float t[9];
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 3; c++) t[3 * r + c] = (A[3 * r] * B[c] + A[3 * r + 1] * B[3 + c]) + A[3 * r + 2] * B[6 + c];
#pragma unroll
  for (int i = 0; i < 9; i++) C[i] = t[i];
}
__device__ __forceinline__ void hat3f(const float* v, float* W) {
  W[0] = 0; W[1] = -v[2]; W[2] = v[1]; W[3] = v[2]; W[4] = 0; W[5] = -v[0]; W[6] = -v[1]; W[7] = v[0]; W[8] = 0;
}

__global__ void __launch_bounds__(64) k_imu_preintegrate(const float* __restrict__ meas, const int* __restrict__ offsets,
                                                         const float* __restrict__ bias, int n, float ng2, float na2, float ngw2,
                                                         float naw2, float* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  float dR[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1}, dV[3] = {0, 0, 0}, dP[3] = {0, 0, 0};
  float JRg[9], JVg[9], JVa[9], JPg[9], JPa[9];
#pragma unroll
  for (int i = 0; i < 9; i++) { JRg[i] = 0; JVg[i] = 0; JVa[i] = 0; JPg[i] = 0; JPa[i] = 0; }
  float* C = out + (size_t)p * GFS_BA_PRE_STRIDE + 60;  // the 15x15 covariance lives in the output record (L1 / L2 resident)
  for (int i = 0; i < 225; i++) C[i] = 0.f;
  float b[6];
#pragma unroll
  for (int i = 0; i < 6; i++) b[i] = bias[(size_t)p * 6 + i];
  const float Nga[6] = {ng2, ng2, ng2, na2, na2, na2};
  float dT = 0.f;
  for (int m = offsets[p]; m < offsets[p + 1]; m++) {
    const float* M = meas + (size_t)m * 7;
    const float dt = M[6];
    const float acc[3] = {M[0] - b[0], M[1] - b[1], M[2] - b[2]};
    float A[81], B[54];
    for (int i = 0; i < 81; i++) A[i] = (i % 10 == 0) ? 1.f : 0.f;
    for (int i = 0; i < 54; i++) B[i] = 0.f;
    float hR[9], hRa[3], Ra[3];
#pragma unroll
    for (int i = 0; i < 9; i++) hR[i] = 0.5f * dR[i];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      hRa[r] = (hR[3 * r] * acc[0] + hR[3 * r + 1] * acc[1]) + hR[3 * r + 2] * acc[2];
      Ra[r] = (dR[3 * r] * acc[0] + dR[3 * r + 1] * acc[1]) + dR[3 * r + 2] * acc[2];
    }
#pragma unroll
    for (int r = 0; r < 3; r++) {
      dP[r] = (dP[r] + dV[r] * dt) + (hRa[r] * dt) * dt;
      dV[r] = dV[r] + Ra[r] * dt;
    }
    float Wacc[9], Rdt[9], nRdt[9], hRdt2[9], nhRdt2[9], T[9], T2[9];
    hat3f(acc, Wacc);
#pragma unroll
    for (int i = 0; i < 9; i++) {
      Rdt[i] = dR[i] * dt;
      nRdt[i] = (-dR[i]) * dt;
      hRdt2[i] = (hR[i] * dt) * dt;
      nhRdt2[i] = ((-0.5f * dR[i]) * dt) * dt;
    }
    mm3f(nRdt, Wacc, T);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) A[9 * (3 + r) + c] = T[3 * r + c];
    mm3f(nhRdt2, Wacc, T);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) A[9 * (6 + r) + c] = T[3 * r + c];
    for (int r = 0; r < 3; r++) A[9 * (6 + r) + 3 + r] = dt;
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) { B[6 * (3 + r) + 3 + c] = Rdt[3 * r + c]; B[6 * (6 + r) + 3 + c] = hRdt2[3 * r + c]; }
#pragma unroll
    for (int i = 0; i < 9; i++) JPa[i] = (JPa[i] + JVa[i] * dt) - hRdt2[i];
    mm3f(hRdt2, Wacc, T); mm3f(T, JRg, T2);
#pragma unroll
    for (int i = 0; i < 9; i++) JPg[i] = (JPg[i] + JVg[i] * dt) - T2[i];
#pragma unroll
    for (int i = 0; i < 9; i++) JVa[i] = JVa[i] - Rdt[i];
    mm3f(Rdt, Wacc, T); mm3f(T, JRg, T2);
#pragma unroll
    for (int i = 0; i < 9; i++) JVg[i] = JVg[i] - T2[i];
    // IntegratedRotation
    const float x = (M[3] - b[3]) * dt, y = (M[4] - b[4]) * dt, z = (M[5] - b[5]) * dt;
    const float d2 = (x * x + y * y) + z * z;
    const float d = (float)sqrt((double)d2);
    const float v[3] = {x, y, z};
    float W[9], WW[9], dRi[9], rJ[9];
    hat3f(v, W);
    mm3f(W, W, WW);
    if (d < 1e-4f) {
#pragma unroll
      for (int i = 0; i < 9; i++) { dRi[i] = ((i % 4 == 0) ? 1.f : 0.f) + W[i]; rJ[i] = (i % 4 == 0) ? 1.f : 0.f; }
    } else {
      const float s = (float)sin((double)d), omc = (float)(1.0 - cos((double)d)), dms = (float)((double)d - sin((double)d));
      const float d3 = d2 * d;
#pragma unroll
      for (int i = 0; i < 9; i++) {
        const float I = (i % 4 == 0) ? 1.f : 0.f;
        dRi[i] = (I + (W[i] * s) / d) + (WW[i] * omc) / d2;
        rJ[i] = (I - (W[i] * omc) / d2) + (WW[i] * dms) / d3;
      }
    }
    mm3f(dR, dRi, dR);
    normalize_rotation(dR);
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) { A[9 * r + c] = dRi[3 * c + r]; B[6 * r + c] = rJ[3 * r + c] * dt; }
    // C[0:9,0:9] = A C A^T + B Nga B^T
    float AC[81];
    for (int r = 0; r < 9; r++)
      for (int c = 0; c < 9; c++) {
        float a = 0;
        for (int k = 0; k < 9; k++) a += A[9 * r + k] * C[15 * k + c];
        AC[9 * r + c] = a;
      }
    for (int r = 0; r < 9; r++)
      for (int c = 0; c < 9; c++) {
        float a = 0, bq = 0;
        for (int k = 0; k < 9; k++) a += AC[9 * r + k] * A[9 * c + k];
        for (int k = 0; k < 6; k++) bq += (B[6 * r + k] * Nga[k]) * B[6 * c + k];
        C[15 * r + c] = a + bq;
      }
    for (int k = 0; k < 3; k++) { C[15 * (9 + k) + 9 + k] += ngw2; C[15 * (12 + k) + 12 + k] += naw2; }
    float dRiT[9];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int c = 0; c < 3; c++) dRiT[3 * r + c] = dRi[3 * c + r];
    mm3f(dRiT, JRg, T);
#pragma unroll
    for (int i = 0; i < 9; i++) JRg[i] = T[i] - rJ[i] * dt;
    dT += dt;
  }
  float* o = out + (size_t)p * GFS_BA_PRE_STRIDE;
  for (int i = 0; i < 9; i++) { o[i] = dR[i]; o[15 + i] = JRg[i]; o[24 + i] = JVg[i]; o[33 + i] = JVa[i]; o[42 + i] = JPg[i]; o[51 + i] = JPa[i]; }
  for (int i = 0; i < 3; i++) { o[9 + i] = dV[i]; o[12 + i] = dP[i]; }
  o[285] = dT;
  for (int i = 0; i < 6; i++) o[286 + i] = b[i];
}

}  // namespace imu
}  // namespace gfs

using namespace gfs;

extern "C" {

int gfs_imu_preintegrate_batch_device(void* stream, const float* d_meas, const int* d_offsets, const float* d_bias, int n, float ng,
                                      float na, float ngw, float naw, float* d_out) {
  GFS_REQUIRE(d_meas && d_offsets && d_bias && d_out, GFS_ERR_INVALID, "null pointer");
  GFS_REQUIRE(n > 0, GFS_ERR_INVALID, "n must be positive");
  int rc = gfs_device_check();
  if (rc) return rc;
  imu::k_imu_preintegrate<<<div_up(n, 64), 64, 0, (cudaStream_t)stream>>>(d_meas, d_offsets, d_bias, n, ng * ng, na * na, ngw * ngw,
                                                                       naw * naw, d_out);
  GFS_CUDA(cudaGetLastError());
  return GFS_OK;
}

int gfs_imu_preintegrate_batch(void* stream, const float* meas, const int* offsets, const float* bias, int n, float ng, float na, float ngw,
                               float naw, float* out) {
  GFS_REQUIRE(meas && offsets && bias && out, GFS_ERR_INVALID, "null pointer");
  GFS_REQUIRE(n > 0, GFS_ERR_INVALID, "n must be positive");
  for (int i = 0; i < n; i++) GFS_REQUIRE(offsets[i + 1] >= offsets[i] && offsets[0] >= 0, GFS_ERR_INVALID, "offsets must be non-decreasing");
  int rc = gfs_device_check();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t total = (size_t)offsets[n];
  float *dm = nullptr, *db = nullptr, *dout = nullptr;
  int* doff = nullptr;
  // every exit path frees whatever was allocated (stream-ordered)
  auto step = [&](cudaError_t e, const char* what) {
    if (rc == GFS_OK && e != cudaSuccess) { gfs::set_error("%s -> %s", what, cudaGetErrorString(e)); rc = GFS_ERR_CUDA; }
    return rc == GFS_OK;
  };
  step(cudaMallocAsync((void**)&dm, std::max<size_t>(total, 1) * 28, st), "cudaMallocAsync(meas)") &&
      step(cudaMallocAsync((void**)&doff, (size_t)(n + 1) * 4, st), "cudaMallocAsync(offsets)") &&
      step(cudaMallocAsync((void**)&db, (size_t)n * 24, st), "cudaMallocAsync(bias)") &&
      step(cudaMallocAsync((void**)&dout, (size_t)n * GFS_BA_PRE_STRIDE * 4, st), "cudaMallocAsync(out)") &&
      (!total || step(cudaMemcpyAsync(dm, meas, total * 28, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(meas)")) &&
      step(cudaMemcpyAsync(doff, offsets, (size_t)(n + 1) * 4, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(offsets)") &&
      step(cudaMemcpyAsync(db, bias, (size_t)n * 24, cudaMemcpyHostToDevice, st), "cudaMemcpyAsync(bias)");
  if (rc == GFS_OK) rc = gfs_imu_preintegrate_batch_device(stream, dm, doff, db, n, ng, na, ngw, naw, dout);
  if (rc == GFS_OK) step(cudaMemcpyAsync(out, dout, (size_t)n * GFS_BA_PRE_STRIDE * 4, cudaMemcpyDeviceToHost, st), "cudaMemcpyAsync(out)");
  if (dm) cudaFreeAsync(dm, st);
  if (doff) cudaFreeAsync(doff, st);
  if (db) cudaFreeAsync(db, st);
  if (dout) cudaFreeAsync(dout, st);
  if (rc) return rc;
  GFS_CUDA(gfs::stream_wait(st));
  return GFS_OK;
}

}  // extern "C"

// Visual-inertial pose-only optimisation of the tracking thread, batched over frames, sm_100a.
//
// Replaces the g2o blocks of Optimizer::PoseInertialOptimizationLastKeyFrame (reference
// src/Optimizer.cc:5899-6284) and Optimizer::PoseInertialOptimizationLastFrame (:6762-7172):
//   vertices  VertexPose (ImuCamPose::Update, src/G2oTypes.cc:193-217), VertexVelocity / GyroBias / AccBias of the
//             frame and, LastFrame, of the previous frame
//   edges     EdgeMonoOnlyPose / EdgeStereoOnlyPose with Huber (G2oTypes.cc:362-383,418-443), EdgeInertial with
//             Huber 6.0 (:473-719), EdgeGyroRW / EdgeAccRW (include/G2oTypes.h:782-852), EdgePriorPoseImu with
//             Huber 5 (G2oTypes.cc:927-995)
//   solver    OptimizationAlgorithmGaussNewton (optimization_algorithm_gauss_newton.cpp:49-93) with
//             LinearSolverDense = pivoted Eigen::LDLT over all 15 / 30 unknowns (linear_solver_dense.h:66-114)
//   outer     nIterations rounds of 10 iterations, chi2 classification with the stale _error g2o keeps, levels,
//             kernels off after the third round, the "recover not too bad points" pass
//   hand-over GetHessian / GetHessian2 blocks at the final estimates, Optimizer::Marginalize (:4408-4487) over the
//             previous frame, the ConstraintPoseImu eigenvalue clamp (G2oTypes.h:857-868)
//
// One CTA per frame runs everything in ONE launch.  Threads stride over the frame's observations (error,
// Jacobian, the 21 + 6 visual entries of H and b in registers, fixed-order block sums); thread 0 linearises the
// inertial edge and thread 32 the prior / random-walk edges meanwhile; the 15 / 30-dim system is assembled entry-
// parallel in shared memory, already in Eigen's pivot order (which follows from the diagonal alone), and
// factorised by the whole CTA (rank-1 trailing updates); the eigen-decompositions of the hand-over run on the whole
// CTA too (seven disjoint Jacobi rotations per step).  No fp atomics: results do not depend on scheduling.
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "common.cuh"
#include "inertial.cuh"

namespace gfs {
namespace pin {

static const int PIN_THREADS = 128;
static const int MAXD = 30;

struct PinHdr {
  int mode, n, off, rounds, rec_init, pad;
  double fx, fy, cx, cy, bf;
  double Rcb[9], tcb[3], tbc[3];
  double cur[KF_STRIDE], prev[KF_STRIDE];
  double infoG[9], infoA[9];
  double c_Rwb[9], c_twb[3], c_vwb[3], c_bg[3], c_ba[3], c_H[225];
  float pre[GFS_BA_PRE_STRIDE];
};
struct PinOut {
  int n_inliers, n_bad, n_inliers_last, rounds_done, gn_iterations[4];
  float avg;
  int pad;
  double Rwb[9], twb[3], vel[3], bg[3], ba[3], H[225];
};
struct PinCam { double fx, fy, cx, cy, bf; const double *Rcb, *tcb, *tbc; };

// ImuCamPose::Update on a body-state record
__device__ void state_oplus(const PinCam& C, double* st, const double* u) {
  double t[3], E[9];
  mv3(st + K_RWB, u + 3, t);
  st[K_TWB] += t[0]; st[K_TWB + 1] += t[1]; st[K_TWB + 2] += t[2];
  exp_so3(u, E);
  mm3(st + K_RWB, E, st + K_RWB);
  double Rbw[9], tbw[3];
  mt3(st + K_RWB, Rbw);
  mv3(Rbw, st + K_TWB, tbw);
  tbw[0] = -tbw[0]; tbw[1] = -tbw[1]; tbw[2] = -tbw[2];
  mm3(C.Rcb, Rbw, st + K_RCW);
  mv3(C.Rcb, tbw, st + K_TCW);
  st[K_TCW] += C.tcb[0]; st[K_TCW + 1] += C.tcb[1]; st[K_TCW + 2] += C.tcb[2];
}
// obs - Project[Stereo](Xw) (G2oTypes.cc:172-188); returns the dimension
__device__ __forceinline__ int vis_err(const PinCam& C, const double* st, const double* Xw, const float* o, double* err, double* Xc) {
  mv3(st + K_RCW, Xw, Xc);
  Xc[0] += st[K_TCW]; Xc[1] += st[K_TCW + 1]; Xc[2] += st[K_TCW + 2];
  const double u = C.fx * Xc[0] / Xc[2] + C.cx, v = C.fy * Xc[1] / Xc[2] + C.cy;
  err[0] = (double)o[0] - u;
  err[1] = (double)o[1] - v;
  err[2] = 0;
  if (o[2] < 0) return 2;
  const double invZ = 1 / Xc[2];
  err[2] = (double)o[2] - (u - C.bf * invZ);
  return 3;
}
// Edge{Mono,Stereo}OnlyPose::linearizeOplus: J = proj_jac * Rcb * SE3deriv(Xb); 3 x 6, unused row zero
__device__ __forceinline__ void vis_jac(const PinCam& C, const double* Xc, bool mono, double* J) {
  double Xb[3];
#pragma unroll
  for (int r = 0; r < 3; r++) Xb[r] = (C.Rcb[r] * Xc[0] + C.Rcb[3 + r] * Xc[1] + C.Rcb[6 + r] * Xc[2]) + C.tbc[r];  // Rbc = Rcb^T
  double pj[9] = {C.fx / Xc[2], 0.0, -C.fx * Xc[0] / (Xc[2] * Xc[2]), 0.0, C.fy / Xc[2], -C.fy * Xc[1] / (Xc[2] * Xc[2]), 0, 0, 0};
  if (!mono) {
    const double inv_z2 = 1.0 / (Xc[2] * Xc[2]);
    pj[6] = pj[0]; pj[7] = pj[1]; pj[8] = pj[2] + C.bf * inv_z2;
  }
  double A[9];
  mm3(pj, C.Rcb, A);
  const double x = Xb[0], y = Xb[1], z = Xb[2];
  const double D[18] = {0, z, -y, 1, 0, 0, -z, 0, x, 0, 1, 0, y, -x, 0, 0, 0, 1};
#pragma unroll
  for (int r = 0; r < 3; r++)
#pragma unroll
    for (int c = 0; c < 6; c++) J[6 * r + c] = A[3 * r] * D[c] + A[3 * r + 1] * D[6 + c] + A[3 * r + 2] * D[12 + c];
  if (mono) {
#pragma unroll
    for (int c = 0; c < 6; c++) J[12 + c] = 0;
  }
}

// fixed-order block sum of NV doubles per thread: lanes (shuffle tree), then warps in order; out[] valid after return
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
    for (int w = 0; w < PIN_THREADS / 32; w++) x += s_part[w * NV + threadIdx.x];
    out[threadIdx.x] = x;
  }
  __syncthreads();
}

// Eigen::LDLT<MatrixXd> pivots on the largest |diagonal| of the not-yet-eliminated part; its unblocked algorithm is
// left-looking, so those diagonal entries are still the ORIGINAL ones: the pivot order is the diagonal sorted by
// decreasing magnitude, known before the factorisation starts.  perm[k] = original index that ends up at position k.
// Exactly equal diagonal entries (the three axes of a bias random walk) are ordered by index here; Eigen orders
// them by its swap history -- either order eliminates the same unknowns, the results differ by rounding only.
__device__ __forceinline__ void ldlt_pivot_order(const double* diag, int n, int* perm) {
  const int i = threadIdx.x;
  if (i < n) {
    const double di = fabs(diag[i]);
    int rank = 0;
    for (int j = 0; j < n; j++) {
      const double dj = fabs(diag[j]);
      rank += (dj > di || (dj == di && j < i)) ? 1 : 0;
    }
    perm[rank] = i;
  }
}
// LDL^T of the symmetrically permuted matrix (P H P^T, already in M: n <= 32, row-major, leading dimension ld, lower
// triangle used, destroyed) by the whole CTA, then the two triangular solves by warp 0.  Right-looking: step k applies
// the rank-1 update of column k to the trailing lower triangle, one entry per thread and pass
// (the factors are those of Eigen's left-looking kernel up to rounding).  Returns isPositive() (Eigen's sign bookkeeping); x (in
// the ORIGINAL ordering) is written only then (LinearSolverDense leaves it untouched otherwise).  bp = P b.
__device__ bool block_ldlt_solve_permuted(double* M, int ld, int n, const double* bp, const int* perm, double* x, double* col) {
  const int tid = threadIdx.x, lane = tid & 31;
  int sign = 2;  // 0 PosSemi, 1 NegSemi, 2 Zero, 3 Indefinite
#define M_(r, c) M[(r) * ld + (c)]
  for (int k = 0; k < n; k++) {
    const double akk = M_(k, k);
    const bool valid = fabs(akk) > 0;
    if (k == 0 && !valid) break;  // ZeroSign: the matrix is treated as zero (uniform)
    if (sign == 0) { if (akk < 0) sign = 3; }
    else if (sign == 1) { if (akk > 0) sign = 3; }
    else if (sign == 2) { if (akk > 0) sign = 0; else if (akk < 0) sign = 1; }
    const int m = n - k - 1;
    if (m == 0) break;
    // l_j = M(j,k) / d_k once per row (double-buffered across steps), then the trailing update with the unscaled
    // column: M(i,j) -= M(i,k) * l_j; warp w takes rows w, w+4, ...
    double* cl = col + (k & 1) * 32;
    if (tid < m) cl[tid] = valid ? M_(k + 1 + tid, k) / akk : M_(k + 1 + tid, k);
    __syncthreads();
    for (int i = tid >> 5; i < m; i += PIN_THREADS / 32) {
      const int j = tid & 31;
      if (j <= i) M_(k + 1 + i, k + 1 + j) -= M_(k + 1 + i, k) * cl[j];
    }
    __syncthreads();
  }
  // L(i,k) = M(i,k) / d_k, all columns at once
  for (int e = tid; e < n * 32; e += PIN_THREADS) {
    const int i = e >> 5, k = e & 31;
    if (k < i) {
      const double d = M_(k, k);
      if (fabs(d) > 0) M_(i, k) /= d;
    }
  }
  __syncthreads();
  const bool positive = sign == 0 || sign == 2;
  if (positive && tid < 32) {
    double y = lane < n ? bp[lane] : 0.0;
    for (int c = 0; c < n; c++) {  // L y = P b (unit lower)
      const double yc = __shfl_sync(0xffffffffu, y, c);
      if (lane > c && lane < n) y -= M_(lane, c) * yc;
    }
    const double tol = 1.0 / DBL_MAX;
    if (lane < n) y = (fabs(M_(lane, lane)) > tol) ? y / M_(lane, lane) : 0.0;
    for (int c = n - 1; c >= 1; c--) {  // L^T z = y
      const double yc = __shfl_sync(0xffffffffu, y, c);
      if (lane < c) y -= M_(c, lane) * yc;
    }
    if (lane < n) x[perm[lane]] = y;
  }
#undef M_
  return positive;
}

// Jacobi eigen-decomposition A = V diag(e) V^T of a symmetric N x N matrix (N odd, <= 15) by the whole CTA (128
// threads): every round applies (N-1)/2 disjoint rotations at once (round-robin pairing, one 16-thread group per
// pair, thread k of a group owning row / column k), N rounds per sweep.  A is destroyed (its diagonal holds the eigenvalues).  The
// oracle runs the cyclic-by-row order; both converge to the same decomposition, and only V f(e) V^T is used.
template <int N>
__device__ void block_jacobi_eig(double* A, double* V, double* red) {
  const int n = N;
  const int tid = threadIdx.x, g = tid >> 4, k = tid & 15;
  for (int i = tid; i < n * n; i += PIN_THREADS) V[i] = (i / n == i % n) ? 1.0 : 0.0;
  __syncthreads();
  for (int sweep = 0; sweep < 100; sweep++) {
    if (tid < n) {
      double off = 0, diag = 0;
      for (int j = 0; j < n; j++) { const double a = A[tid * n + j]; if (j == tid) diag += a * a; else off += a * a; }
      red[tid] = off; red[16 + tid] = diag;
    }
    __syncthreads();
    double off = 0, diag = 0;
    for (int j = 0; j < n; j++) { off += red[j]; diag += red[16 + j]; }
    if (off <= 1e-32 * diag) break;  // uniform
    for (int r = 0; r < n; r++) {
      // circle method over 16 players (15 = bye): pair g+1 of round r
      int p = (r + g + 1) % n, q = (r + n - (g + 1)) % n;
      if (p > q) { const int t = p; p = q; q = t; }
      double c = 1.0, s = 0.0;
      const bool on = g < (N - 1) / 2 && k < n;
      if (on) {
        const double apq = A[p * n + q];
        if (apq != 0.0) {
          const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
          const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
          c = 1.0 / sqrt(t * t + 1.0);
          s = t * c;
        }
      }
      __syncwarp();  // the pair's (p,p), (q,q), (p,q) are read by its whole group before anyone overwrites them
      if (on) {
        const double akp = A[k * n + p], akq = A[k * n + q];
        A[k * n + p] = c * akp - s * akq;
        A[k * n + q] = s * akp + c * akq;
        const double vkp = V[k * n + p], vkq = V[k * n + q];
        V[k * n + p] = c * vkp - s * vkq;
        V[k * n + q] = s * vkp + c * vkq;
      }
      __syncthreads();
      if (on) {
        const double apk = A[p * n + k], aqk = A[q * n + k];
        A[p * n + k] = c * apk - s * aqk;
        A[q * n + k] = s * apk + c * aqk;
      }
      __syncthreads();
    }
  }
  __syncthreads();
}

// EdgeInertial constructor (G2oTypes.cc:487-494): Info = C[0:9,0:9]^-1, symmetrised, eigenvalues < 1e-12 zeroed.
// The inverse is the Gauss-Jordan elimination with partial pivoting the host / oracle restatement runs (the same
// operations entry by entry), spread over the CTA; scratch: aug[9 x 18], E[81], V[81].
__device__ void block_inertial_information(const float* C15, double* info81, double* aug, double* E, double* V, double* red) {
  const int tid = threadIdx.x;
  for (int i = tid; i < 162; i += PIN_THREADS) {
    const int r = i / 18, k = i % 18;
    aug[i] = k < 9 ? (double)C15[15 * r + k] : (k - 9 == r ? 1.0 : 0.0);
  }
  __syncthreads();
  for (int c = 0; c < 9; c++) {
    int piv = c;
    for (int r = c + 1; r < 9; r++)
      if (fabs(aug[r * 18 + c]) > fabs(aug[piv * 18 + c])) piv = r;
    const bool singular = aug[piv * 18 + c] == 0.0;
    __syncthreads();
    if (singular) break;  // uniform
    if (piv != c) {
      if (tid < 18) { const double t = aug[c * 18 + tid]; aug[c * 18 + tid] = aug[piv * 18 + tid]; aug[piv * 18 + tid] = t; }
      __syncthreads();
    }
    const double d = 1.0 / aug[c * 18 + c];
    __syncthreads();
    if (tid < 18) aug[c * 18 + tid] *= d;
    __syncthreads();
    double f[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int i = tid + u * PIN_THREADS;
      f[u] = i < 162 ? aug[(i / 18) * 18 + c] : 0.0;
    }
    __syncthreads();
#pragma unroll
    for (int u = 0; u < 2; u++) {
      const int i = tid + u * PIN_THREADS;
      if (i < 162 && i / 18 != c && f[u] != 0.0) aug[i] -= f[u] * aug[c * 18 + i % 18];
    }
    __syncthreads();
  }
  for (int i = tid; i < 81; i += PIN_THREADS) {
    const int r = i / 9, c = i % 9;
    E[i] = (aug[r * 18 + 9 + c] + aug[c * 18 + 9 + r]) / 2;
  }
  __syncthreads();
  block_jacobi_eig<9>(E, V, red);
  for (int i = tid; i < 81; i += PIN_THREADS) {
    const int r = i / 9, c = i % 9;
    double a = 0;
    for (int k = 0; k < 9; k++) {
      double ev = E[10 * k];
      if (ev < 1e-12) ev = 0;
      a += V[9 * r + k] * ev * V[9 * c + k];
    }
    info81[i] = a;
  }
  __syncthreads();
}

struct PinShared {
  double cur[KF_STRIDE], prev[KF_STRIDE];
  double H[MAXD * (MAXD + 1)], b[MAXD], x[MAXD], tmp[32], hd[32], col[64];  // H rows padded to an odd length (bank spread)
  double J[225], WJ[225], We[16];   // inertial edge: 9 x 24; prior edge: 15 x 15 (second slot)
  double Jp[225], WJp[225], Wep[16];
  double infoI[81];
  double errI[9], errG[3], errA[3], errP[15];
  double wI, wP, deltaI;
  double sums[28], part[(PIN_THREADS / 32) * 28];
  double V[225], E[225];
  int perm[32];
  int ok, counts[3];
};

// EdgePriorPoseImu::computeError / linearizeOplus
__device__ void prior_err(const PinHdr& h, const double* st, double* e15) {
  double Rt[9], M[9], d[3];
  mt3(h.c_Rwb, Rt);
  mm3(Rt, st + K_RWB, M);
  log_so3(M, e15);
  for (int i = 0; i < 3; i++) d[i] = st[K_TWB + i] - h.c_twb[i];
  mv3(Rt, d, e15 + 3);
  for (int i = 0; i < 3; i++) { e15[6 + i] = st[K_VEL + i] - h.c_vwb[i]; e15[9 + i] = st[K_BG + i] - h.c_bg[i]; e15[12 + i] = st[K_BA + i] - h.c_ba[i]; }
}
__device__ void prior_jac(const PinHdr& h, const double* st, double* J) {
  double Rt[9], M[9], er[3], iJ[9];
  mt3(h.c_Rwb, Rt);
  mm3(Rt, st + K_RWB, M);
  log_so3(M, er);
  inv_right_jac(er, iJ);
  for (int i = 0; i < 225; i++) J[i] = 0;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) { J[15 * r + c] = iJ[3 * r + c]; J[15 * (3 + r) + 3 + c] = M[3 * r + c]; }
  for (int i = 6; i < 15; i++) J[15 * i + i] = 1.0;
}
__device__ double quad_form(const double* e, const double* Om, int n) {
  double s = 0;
  for (int r = 0; r < n; r++) {
    double a = 0;
    for (int c = 0; c < n; c++) a += Om[n * r + c] * e[c];
    s += e[r] * a;
  }
  return s;
}

__global__ void __launch_bounds__(PIN_THREADS, 4) k_pose_inertial(const PinHdr* __restrict__ hdr, const double* __restrict__ gXw,
                                                               const float* __restrict__ guvr, const float* __restrict__ gis2,
                                                               const uint8_t* __restrict__ gclose, double* __restrict__ gerr,
                                                               uint8_t* __restrict__ glevel, uint8_t* __restrict__ goutlier,
                                                               float* __restrict__ gchi2, PinOut* __restrict__ gout) {
  __shared__ PinShared S;
  const int p = blockIdx.x, tid = threadIdx.x;
  const PinHdr& h = hdr[p];
  const int n = h.n;
  const bool freePrev = h.mode == GFS_PIN_LAST_FRAME;
  const int dim = freePrev ? 30 : 15;
  const double* Xw = gXw + (size_t)h.off * 3;
  const float* uvr = guvr + (size_t)h.off * 3;
  const float* is2 = gis2 + h.off;
  const uint8_t* close = gclose + h.off;
  double* err = gerr + (size_t)h.off * 3;
  uint8_t* level = glevel + h.off;
  uint8_t* outlier = goutlier + h.off;
  float* chi2 = gchi2 + h.off;
  PinOut* out = gout + p;
  PinCam C;
  C.fx = h.fx; C.fy = h.fy; C.cx = h.cx; C.cy = h.cy; C.bf = h.bf;
  C.Rcb = h.Rcb; C.tcb = h.tcb; C.tbc = h.tbc;
  const double deltaMono = (double)(float)sqrt(5.991), deltaStereo = (double)(float)sqrt(7.815);
  for (int i = tid; i < KF_STRIDE; i += PIN_THREADS) { S.cur[i] = h.cur[i]; S.prev[i] = h.prev[i]; }
  for (int i = tid; i < MAXD; i += PIN_THREADS) S.x[i] = 0.0;
  for (int e = tid; e < n; e += PIN_THREADS) { level[e] = 0; outlier[e] = 0; chi2[e] = 0.f; }
  if (tid == 0) { S.deltaI = 6.0; S.ok = 1; }
  __syncthreads();
  block_inertial_information(h.pre + 60, S.infoI, S.H, S.E, S.V, S.tmp);
  const bool lastKF = !freePrev;
  const float chi2MonoKF[4] = {12, 7.5, 5.991, 5.991}, chi2MonoF[4] = {5.991, 5.991, 5.991, 5.991};
  const float chi2Stereo[4] = {15.6f, 9.8f, 7.815f, 7.815f};
  int nBad = 0, nInliers = 0, roundsDone = 0;
  int gnIters[4] = {0, 0, 0, 0};
  float avgOut = 0.f;
  const int rounds = min(max(h.rounds, 0), 4);

  for (int it = 0; it < rounds; it++) {
    const bool robust = it < 3;  // setRobustKernel(0) on every visual edge at the end of round 2
    for (int iter = 0; iter < 10; iter++) {
      // ---- serial edges: thread 0 the inertial edge, thread 32 the random-walk and prior edges
      if (tid == 0) {
        double r9[9];
        inertial_jacobian_core(S.prev, S.cur, h.pre, S.J, r9);
        for (int i = 0; i < 9; i++) S.errI[i] = r9[i];
        double rho[2];
        huber(quad_form(S.errI, S.infoI, 9), S.deltaI, rho);
        S.wI = rho[1];
      } else if (tid == 32) {
        for (int i = 0; i < 3; i++) { S.errG[i] = S.cur[K_BG + i] - S.prev[K_BG + i]; S.errA[i] = S.cur[K_BA + i] - S.prev[K_BA + i]; }
        if (freePrev) {
          prior_err(h, S.prev, S.errP);
          prior_jac(h, S.prev, S.Jp);
          double rho[2];
          huber(quad_form(S.errP, h.c_H, 15), 5.0, rho);
          S.wP = rho[1];
        }
      }
      // ---- visual edges: computeActiveErrors + linearizeOplus + constructQuadraticForm
      double acc[27];
#pragma unroll
      for (int k = 0; k < 27; k++) acc[k] = 0.0;
      for (int e = tid; e < n; e += PIN_THREADS) {
        if (level[e] != 0) continue;
        const float* o = uvr + 3 * (size_t)e;
        const bool mono = o[2] < 0;
        double er[3], Xc[3], J[18];
        vis_err(C, S.cur, Xw + 3 * (size_t)e, o, er, Xc);
        err[3 * (size_t)e] = er[0]; err[3 * (size_t)e + 1] = er[1]; err[3 * (size_t)e + 2] = er[2];
        vis_jac(C, Xc, mono, J);
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
      block_sum<27>(acc, S.part, S.sums);  // barriers inside: the serial edges are published too
      // ---- assembly, entry-parallel
      for (int i = tid; i < 9 * 24; i += PIN_THREADS) {  // WJ = (w Omega) J
        const int r = i / 24, c = i % 24;
        double a = 0;
        for (int k = 0; k < 9; k++) a += (S.wI * S.infoI[9 * r + k]) * S.J[24 * k + c];
        S.WJ[i] = a;
      }
      if (tid < 9) {
        double a = 0;
        for (int k = 0; k < 9; k++) a += S.infoI[9 * tid + k] * S.errI[k];
        S.We[tid] = S.wI * a;
      }
      if (freePrev) {
        for (int i = tid; i < 225; i += PIN_THREADS) {
          const int r = i / 15, c = i % 15;
          double a = 0;
          for (int k = 0; k < 15; k++) a += (S.wP * h.c_H[15 * r + k]) * S.Jp[15 * k + c];
          S.WJp[i] = a;
        }
        if (tid >= 32 && tid < 47) {
          const int r = tid - 32;
          double a = 0;
          for (int k = 0; k < 15; k++) a += h.c_H[15 * r + k] * S.errP[k];
          S.Wep[r] = S.wP * a;
        }
      }
      __syncthreads();
      // entry (a, c) of H: the edges are added in a fixed order: visual, inertial, random walks, prior
      auto h_entry = [&](int a, int c) {
        double v = 0.0;
        if (a < 6 && c < 6) {
          const int lo = min(a, c), hi = max(a, c);
          v = S.sums[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)];
        }
        // inertial: unknown a -> Jacobian column (cur pose/vel 0..8 -> 15..23, prev 15..29 -> 0..14)
        const int ja = a < 9 ? 15 + a : (a >= 15 ? a - 15 : -1), jc = c < 9 ? 15 + c : (c >= 15 ? c - 15 : -1);
        if (ja >= 0 && jc >= 0) {
          double t = 0;
#pragma unroll
          for (int r = 0; r < 9; r++) t += S.J[24 * r + ja] * S.WJ[24 * r + jc];
          v += t;
        }
        // random walks: J = [-I, I] over (prev bias, cur bias)
        const int ga = (a >= 9 && a < 12) ? a - 9 : ((a >= 24 && a < 27) ? a - 24 : -1), gc = (c >= 9 && c < 12) ? c - 9 : ((c >= 24 && c < 27) ? c - 24 : -1);
        if (ga >= 0 && gc >= 0) v += (((a >= 24) != (c >= 24)) ? -1.0 : 1.0) * h.infoG[3 * ga + gc];
        const int aa = (a >= 12 && a < 15) ? a - 12 : ((a >= 27 && a < 30) ? a - 27 : -1), ac = (c >= 12 && c < 15) ? c - 12 : ((c >= 27 && c < 30) ? c - 27 : -1);
        if (aa >= 0 && ac >= 0) v += (((a >= 27) != (c >= 27)) ? -1.0 : 1.0) * h.infoA[3 * aa + ac];
        if (freePrev && a >= 15 && c >= 15) {
          double t = 0;
#pragma unroll
          for (int r = 0; r < 15; r++) t += S.Jp[15 * r + (a - 15)] * S.WJp[15 * r + (c - 15)];
          v += t;
        }
        return v;
      };
      auto b_entry = [&](int a) {
        double v = a < 6 ? S.sums[21 + a] : 0.0;
        const int ja = a < 9 ? 15 + a : (a >= 15 ? a - 15 : -1);
        if (ja >= 0) {
          double t = 0;
          for (int r = 0; r < 9; r++) t += S.J[24 * r + ja] * S.We[r];
          v -= t;
        }
        if ((a >= 9 && a < 12) || (a >= 24 && a < 27)) {
          const int r = a >= 24 ? a - 24 : a - 9;
          const double g = h.infoG[3 * r] * S.errG[0] + h.infoG[3 * r + 1] * S.errG[1] + h.infoG[3 * r + 2] * S.errG[2];
          v -= (a >= 24 ? -1.0 : 1.0) * g;
        }
        if ((a >= 12 && a < 15) || (a >= 27 && a < 30)) {
          const int r = a >= 27 ? a - 27 : a - 12;
          const double g = h.infoA[3 * r] * S.errA[0] + h.infoA[3 * r + 1] * S.errA[1] + h.infoA[3 * r + 2] * S.errA[2];
          v -= (a >= 27 ? -1.0 : 1.0) * g;
        }
        if (freePrev && a >= 15) {
          double t = 0;
          for (int r = 0; r < 15; r++) t += S.Jp[15 * r + (a - 15)] * S.Wep[r];
          v -= t;
        }
        return v;
      };
      // the diagonal first: it fixes Eigen's pivot order, H is then assembled already permuted
      if (tid < dim) S.hd[tid] = h_entry(tid, tid);
      __syncthreads();
      ldlt_pivot_order(S.hd, dim, S.perm);
      __syncthreads();
      const int ld = dim + 1;
      for (int i = tid; i < dim * dim; i += PIN_THREADS) {
        const int r = i / dim, c = i % dim;
        if (c <= r) S.H[r * ld + c] = h_entry(S.perm[r], S.perm[c]);
      }
      if (tid < dim) S.b[tid] = b_entry(S.perm[tid]);
      __syncthreads();
      // ---- solve + update
      {
        const bool ok = block_ldlt_solve_permuted(S.H, ld, dim, S.b, S.perm, S.x, S.col);
        if (tid == 0) S.ok = ok ? 1 : 0;
      }
      __syncthreads();
      if (tid == 0) {
        state_oplus(C, S.cur, S.x);
        for (int i = 0; i < 3; i++) { S.cur[K_VEL + i] += S.x[6 + i]; S.cur[K_BG + i] += S.x[9 + i]; S.cur[K_BA + i] += S.x[12 + i]; }
      } else if (tid == 32 && freePrev) {
        state_oplus(C, S.prev, S.x + 15);
        for (int i = 0; i < 3; i++) { S.prev[K_VEL + i] += S.x[21 + i]; S.prev[K_BG + i] += S.x[24 + i]; S.prev[K_BA + i] += S.x[27 + i]; }
      }
      __syncthreads();
      gnIters[it]++;
      if (!S.ok) break;
    }
    // ---- classification
    const float chi2close = 1.5 * (lastKF ? chi2MonoKF[it] : chi2MonoF[it]);
    const float thMono = lastKF ? chi2MonoKF[it] : chi2MonoF[it], thStereo = chi2Stereo[it];
    if (tid == 0 && lastKF) {
      const float chi2IMU = (float)quad_form(S.errI, S.infoI, 9);
      if (chi2IMU > 15.0) {
        S.deltaI = 2.0;
        for (int i = 0; i < 81; i++) S.infoI[i] *= 1e-2;
      }
    }
    int bad = 0, good = 0;
    for (int e = tid; e < n; e += PIN_THREADS) {
      const float* o = uvr + 3 * (size_t)e;
      const bool mono = o[2] < 0;
      double er[3], Xc[3];
      if (outlier[e]) {
        vis_err(C, S.cur, Xw + 3 * (size_t)e, o, er, Xc);
        err[3 * (size_t)e] = er[0]; err[3 * (size_t)e + 1] = er[1]; err[3 * (size_t)e + 2] = er[2];
      } else {
        er[0] = err[3 * (size_t)e]; er[1] = err[3 * (size_t)e + 1]; er[2] = err[3 * (size_t)e + 2];
      }
      const double om = (double)is2[e];
      double c2 = er[0] * (om * er[0]) + er[1] * (om * er[1]);
      if (!mono) c2 += er[2] * (om * er[2]);
      const float c2f = (float)c2;
      chi2[e] = c2f;
      bool isBad;
      if (mono) {
        const bool bClose = close[e] != 0;
        const double* X = Xw + 3 * (size_t)e;
        const bool depthPos = (S.cur[K_RCW + 6] * X[0] + S.cur[K_RCW + 7] * X[1] + S.cur[K_RCW + 8] * X[2] + S.cur[K_TCW + 2]) > 0.0;
        isBad = (c2f > thMono && !bClose) || (bClose && c2f > chi2close) || !depthPos;
      } else {
        isBad = c2f > thStereo;
      }
      if (isBad) { outlier[e] = 1; level[e] = 1; bad++; }
      else { outlier[e] = 0; level[e] = 0; good++; }
    }
    // counts (integers: order-free) and the chi2 sum of the inliers.  The reference adds the chi2 in float, monocular
    // edges first, then stereo, in frame order; here both groups are summed in double in the block's fixed tree order
    // and narrowed once (the two agree to float rounding, ~1e-7 relative)
    double cs[2] = {0.0, 0.0};
    for (int e = tid; e < n; e += PIN_THREADS)
      if (!outlier[e]) cs[uvr[3 * (size_t)e + 2] < 0 ? 0 : 1] += (double)chi2[e];
    if (tid == 0) { S.counts[0] = 0; S.counts[1] = 0; }
    block_sum<2>(cs, S.part, S.sums);
    atomicAdd(&S.counts[0], bad);
    atomicAdd(&S.counts[1], good);
    __syncthreads();
    nBad = S.counts[0];
    nInliers = S.counts[1];
    {
      float avg = (float)S.sums[0] + (float)S.sums[1];
      avg /= nInliers;
      avgOut = avg;
    }
    __syncthreads();
    roundsDone = it + 1;
    if (n + (freePrev ? 4 : 3) < 10) break;  // optimizer.edges().size() < 10
  }
  // ---- "If not too much tracks, recover not too bad points"
  if (nInliers < 30 && !h.rec_init) {
    int bad = 0;
    for (int e = tid; e < n; e += PIN_THREADS) {
      const float* o = uvr + 3 * (size_t)e;
      const bool mono = o[2] < 0;
      double er[3], Xc[3];
      vis_err(C, S.cur, Xw + 3 * (size_t)e, o, er, Xc);
      const double om = (double)is2[e];
      double c2 = er[0] * (om * er[0]) + er[1] * (om * er[1]);
      if (!mono) c2 += er[2] * (om * er[2]);
      if (c2 < (mono ? 18.0 : 24.0)) outlier[e] = 0;
      else bad++;
    }
    __syncthreads();
    if (tid == 0) S.counts[0] = 0;
    __syncthreads();
    atomicAdd(&S.counts[0], bad);
    __syncthreads();
    nBad = S.counts[0];
  }
  __syncthreads();
  // ---- Hessian hand-over: fresh linearisation at the final estimates, raw information, inlier edges only
  {
    double acc[21];
#pragma unroll
    for (int k = 0; k < 21; k++) acc[k] = 0.0;
    for (int e = tid; e < n; e += PIN_THREADS) {
      if (outlier[e]) continue;
      const float* o = uvr + 3 * (size_t)e;
      const bool mono = o[2] < 0;
      double er[3], Xc[3], J[18];
      vis_err(C, S.cur, Xw + 3 * (size_t)e, o, er, Xc);
      vis_jac(C, Xc, mono, J);
      const double om = (double)is2[e];
      int k = 0;
#pragma unroll
      for (int a = 0; a < 6; a++) {
#pragma unroll
        for (int c = a; c < 6; c++) acc[k++] += J[a] * (om * J[c]) + J[6 + a] * (om * J[6 + c]) + J[12 + a] * (om * J[12 + c]);
      }
    }
    if (tid == 0) {
      double r9[9];
      inertial_jacobian_core(S.prev, S.cur, h.pre, S.J, r9);
    } else if (tid == 32 && freePrev) {
      prior_jac(h, S.prev, S.Jp);
    }
    block_sum<21>(acc, S.part, S.sums);
  }
  for (int i = tid; i < 9 * 24; i += PIN_THREADS) {  // information() * J
    const int r = i / 24, c = i % 24;
    double a = 0;
    for (int k = 0; k < 9; k++) a += S.infoI[9 * r + k] * S.J[24 * k + c];
    S.WJ[i] = a;
  }
  if (freePrev)
    for (int i = tid; i < 225; i += PIN_THREADS) {
      const int r = i / 15, c = i % 15;
      double a = 0;
      for (int k = 0; k < 15; k++) a += h.c_H[15 * r + k] * S.Jp[15 * k + c];
      S.WJp[i] = a;
    }
  __syncthreads();
  // H30 in the reference's order [prev(15) | cur(15)] (LastFrame) or H15 over cur (LastKeyFrame)
  for (int i = tid; i < dim * dim; i += PIN_THREADS) {
    const int a = i / dim, c = i % dim;
    double v = 0.0;
    if (lastKF) {
      if (a < 9 && c < 9) {  // GetHessian2: [pose2 vel2]
        double s = 0;
        for (int r = 0; r < 9; r++) s += S.J[24 * r + 15 + a] * S.WJ[24 * r + 15 + c];
        v += s;
      }
      if (a >= 9 && a < 12 && c >= 9 && c < 12) v += h.infoG[3 * (a - 9) + (c - 9)];
      if (a >= 12 && c >= 12) v += h.infoA[3 * (a - 12) + (c - 12)];
      if (a < 6 && c < 6) { const int lo = min(a, c), hi = max(a, c); v += S.sums[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)]; }
    } else {
      if (a < 24 && c < 24) {
        double s = 0;
        for (int r = 0; r < 9; r++) s += S.J[24 * r + a] * S.WJ[24 * r + c];
        v += s;
      }
      const int ga = (a >= 9 && a < 12) ? a - 9 : ((a >= 24 && a < 27) ? a - 24 : -1), gc = (c >= 9 && c < 12) ? c - 9 : ((c >= 24 && c < 27) ? c - 24 : -1);
      if (ga >= 0 && gc >= 0) v += (((a >= 24) != (c >= 24)) ? -1.0 : 1.0) * h.infoG[3 * ga + gc];
      const int aa = (a >= 12 && a < 15) ? a - 12 : ((a >= 27 && a < 30) ? a - 27 : -1), ac = (c >= 12 && c < 15) ? c - 12 : ((c >= 27 && c < 30) ? c - 27 : -1);
      if (aa >= 0 && ac >= 0) v += (((a >= 27) != (c >= 27)) ? -1.0 : 1.0) * h.infoA[3 * aa + ac];
      if (a < 15 && c < 15) {
        double s = 0;
        for (int r = 0; r < 15; r++) s += S.Jp[15 * r + a] * S.WJp[15 * r + c];
        v += s;
      }
      if (a >= 15 && a < 21 && c >= 15 && c < 21) { const int lo = min(a, c) - 15, hi = max(a, c) - 15; v += S.sums[lo * 6 - lo * (lo - 1) / 2 + (hi - lo)]; }
    }
    S.H[i] = v;
  }
  __syncthreads();
  double* Hout = S.WJ;  // 15 x 15 result
  if (lastKF) {
    for (int i = tid; i < 225; i += PIN_THREADS) Hout[i] = S.H[i];
    __syncthreads();
  } else {
    // Optimizer::Marginalize(H, 0, 14): pseudo-inverse of the previous frame's block through its eigen-decomposition
    for (int i = tid; i < 225; i += PIN_THREADS) S.E[i] = S.H[30 * (i / 15) + (i % 15)];
    __syncthreads();
    block_jacobi_eig<15>(S.E, S.V, S.tmp);
    for (int i = tid; i < 225; i += PIN_THREADS) {  // inv = V diag(1/lambda, |lambda| > 1e-6) V^T
      const int r = i / 15, c = i % 15;
      double a = 0;
      for (int k = 0; k < 15; k++) {
        const double ev = S.E[16 * k];
        if (fabs(ev) > 1e-6) a += S.V[15 * r + k] * (1.0 / ev) * S.V[15 * c + k];
      }
      S.Jp[i] = a;
    }
    __syncthreads();
    for (int i = tid; i < 225; i += PIN_THREADS) {  // T = Hcb inv
      const int r = i / 15, c = i % 15;
      double a = 0;
      for (int k = 0; k < 15; k++) a += S.H[30 * (15 + r) + k] * S.Jp[15 * k + c];
      S.WJp[i] = a;
    }
    __syncthreads();
    for (int i = tid; i < 225; i += PIN_THREADS) {
      const int r = i / 15, c = i % 15;
      double a = 0;
      for (int k = 0; k < 15; k++) a += S.WJp[15 * r + k] * S.H[30 * k + 15 + c];
      Hout[i] = S.H[30 * (15 + r) + 15 + c] - a;
    }
    __syncthreads();
  }
  // ConstraintPoseImu: H = (H + H) / 2 (sic), SelfAdjointEigenSolver on the lower triangle, eigenvalues < 1e-12 -> 0
  for (int i = tid; i < 225; i += PIN_THREADS) {
    const int r = i / 15, c = i % 15;
    const double v = c <= r ? Hout[i] : Hout[15 * c + r];
    S.E[i] = (v + v) / 2;
  }
  __syncthreads();
  block_jacobi_eig<15>(S.E, S.V, S.tmp);
  for (int i = tid; i < 225; i += PIN_THREADS) {
    const int r = i / 15, c = i % 15;
    double a = 0;
    for (int k = 0; k < 15; k++) {
      double ev = S.E[16 * k];
      if (ev < 1e-12) ev = 0;
      a += S.V[15 * r + k] * ev * S.V[15 * c + k];
    }
    out->H[i] = a;
  }
  if (tid == 0) {
    out->n_bad = nBad; out->n_inliers_last = nInliers; out->n_inliers = n - nBad; out->avg = avgOut;
    out->rounds_done = roundsDone;
    for (int i = 0; i < 4; i++) out->gn_iterations[i] = gnIters[i];
    for (int i = 0; i < 9; i++) out->Rwb[i] = S.cur[K_RWB + i];
    for (int i = 0; i < 3; i++) { out->twb[i] = S.cur[K_TWB + i]; out->vel[i] = S.cur[K_VEL + i]; out->bg[i] = S.cur[K_BG + i]; out->ba[i] = S.cur[K_BA + i]; }
  }
}

}  // namespace pin
}  // namespace gfs

using namespace gfs;
using namespace gfs::pin;

struct GfsPoseInertial {
  int maxObs = 0, maxBatch = 0;
  DevBuf d_hdr, d_Xw, d_uvr, d_is2, d_close, d_err, d_level, d_outlier, d_chi2, d_out;
  PinnedBuf h_in, h_out;
  int launches = 0;
};

extern "C" {

int gfs_pose_inertial_create(int max_obs, int max_batch, GfsPoseInertial** out) {
  GFS_REQUIRE(out, GFS_ERR_INVALID, "out is null");
  *out = nullptr;
  GFS_REQUIRE(max_obs > 0 && max_batch > 0, GFS_ERR_INVALID, "bad capacity");
  int rc = gfs_device_check();
  if (rc) return rc;
  GfsPoseInertial* h = new GfsPoseInertial();
  h->maxObs = max_obs;
  h->maxBatch = max_batch;
  const size_t N = (size_t)max_obs * max_batch, B = max_batch;
  if ((rc = h->d_hdr.reserve(B * sizeof(PinHdr))) || (rc = h->d_Xw.reserve(N * 24)) || (rc = h->d_uvr.reserve(N * 12)) ||
      (rc = h->d_is2.reserve(N * 4)) || (rc = h->d_close.reserve(N)) || (rc = h->d_err.reserve(N * 24)) ||
      (rc = h->d_level.reserve(N)) || (rc = h->d_outlier.reserve(N)) || (rc = h->d_chi2.reserve(N * 4)) ||
      (rc = h->d_out.reserve(B * sizeof(PinOut))) || (rc = h->h_in.reserve(B * sizeof(PinHdr) + N * (24 + 12 + 4 + 1) + 64)) ||
      (rc = h->h_out.reserve(B * sizeof(PinOut) + N * 5 + 64))) {
    gfs_pose_inertial_destroy(h);
    return rc;
  }
  *out = h;
  return GFS_OK;
}

int gfs_pose_inertial_destroy(GfsPoseInertial* h) {
  if (!h) return GFS_OK;
  DevBuf* d[] = {&h->d_hdr, &h->d_Xw, &h->d_uvr, &h->d_is2, &h->d_close, &h->d_err, &h->d_level, &h->d_outlier, &h->d_chi2, &h->d_out};
  for (DevBuf* b : d) b->release();
  h->h_in.release();
  h->h_out.release();
  delete h;
  return GFS_OK;
}

int gfs_pose_inertial_last_launches(const GfsPoseInertial* h) { return h ? h->launches : GFS_ERR_INVALID; }

int gfs_pose_inertial_optimize_batch(GfsPoseInertial* h, void* stream, const GfsPoseInertialProblem* problems, int batch,
                                     GfsPoseInertialResult* results) {
  GFS_REQUIRE(h && problems && results, GFS_ERR_INVALID, "null argument");
  GFS_REQUIRE(batch > 0 && batch <= h->maxBatch, GFS_ERR_CAPACITY, "batch exceeds the handle's max_batch");
  size_t total = 0;
  for (int p = 0; p < batch; p++) {
    const GfsPoseInertialProblem& P = problems[p];
    GFS_REQUIRE(P.mode == GFS_PIN_LAST_KEYFRAME || P.mode == GFS_PIN_LAST_FRAME, GFS_ERR_INVALID, "bad mode");
    GFS_REQUIRE(P.n_obs >= 0 && P.n_obs <= h->maxObs, GFS_ERR_CAPACITY, "n_obs exceeds the handle's max_obs");
    GFS_REQUIRE(P.n_rounds >= 0 && P.n_rounds <= 4, GFS_ERR_INVALID, "n_rounds must be 0..4 (the reference's threshold tables hold 4 entries)");
    GFS_REQUIRE(P.pre, GFS_ERR_INVALID, "null preintegration record");
    GFS_REQUIRE(P.n_obs == 0 || (P.Xw && P.uvr && P.inv_sigma2 && P.close), GFS_ERR_INVALID, "null observation arrays");
    GFS_REQUIRE(P.n_obs == 0 || results[p].outlier, GFS_ERR_INVALID, "null outlier output");
    total += (size_t)P.n_obs;
  }
  cudaStream_t st = (cudaStream_t)stream;
  const size_t NB = (size_t)h->maxObs * h->maxBatch;
  uint8_t* hp = (uint8_t*)h->h_in.p;
  PinHdr* hh = (PinHdr*)hp;
  double* hX = (double*)(hp + align_up((size_t)h->maxBatch * sizeof(PinHdr), 16));
  float* hU = (float*)((uint8_t*)hX + NB * 24);
  float* hS = hU + NB * 3;
  uint8_t* hC = (uint8_t*)(hS + NB);
  size_t off = 0;
  for (int p = 0; p < batch; p++) {
    const GfsPoseInertialProblem& P = problems[p];
    PinHdr& H = hh[p];
    memset(&H, 0, sizeof(H));
    H.mode = P.mode; H.n = P.n_obs; H.off = (int)off; H.rounds = P.n_rounds; H.rec_init = P.rec_init;
    H.fx = P.fx; H.fy = P.fy; H.cx = P.cx; H.cy = P.cy; H.bf = P.bf;
    memcpy(H.Rcb, P.Rcb, 72); memcpy(H.tcb, P.tcb, 24); memcpy(H.tbc, P.tbc, 24);
    memcpy(H.cur + K_RWB, P.Rwb, 72); memcpy(H.cur + K_TWB, P.twb, 24); memcpy(H.cur + K_RCW, P.Rcw, 72);
    memcpy(H.cur + K_TCW, P.tcw, 24); memcpy(H.cur + K_VEL, P.vel, 24); memcpy(H.cur + K_BG, P.bg, 24); memcpy(H.cur + K_BA, P.ba, 24);
    memcpy(H.prev + K_RWB, P.p_Rwb, 72); memcpy(H.prev + K_TWB, P.p_twb, 24); memcpy(H.prev + K_VEL, P.p_vel, 24);
    memcpy(H.prev + K_BG, P.p_bg, 24); memcpy(H.prev + K_BA, P.p_ba, 24);
    memcpy(H.pre, P.pre, sizeof(H.pre));
    double Cg[9], Ca[9];
    for (int i = 0; i < 9; i++) { Cg[i] = (double)P.rw_Cg[i]; Ca[i] = (double)P.rw_Ca[i]; }
    inv3_host(Cg, H.infoG);
    inv3_host(Ca, H.infoA);
    memcpy(H.c_Rwb, P.c_Rwb, 72); memcpy(H.c_twb, P.c_twb, 24); memcpy(H.c_vwb, P.c_vwb, 24); memcpy(H.c_bg, P.c_bg, 24);
    memcpy(H.c_ba, P.c_ba, 24); memcpy(H.c_H, P.c_H, sizeof(H.c_H));
    if (P.n_obs) {
      memcpy(hX + off * 3, P.Xw, (size_t)P.n_obs * 24);
      memcpy(hU + off * 3, P.uvr, (size_t)P.n_obs * 12);
      memcpy(hS + off, P.inv_sigma2, (size_t)P.n_obs * 4);
      memcpy(hC + off, P.close, (size_t)P.n_obs);
    }
    off += (size_t)P.n_obs;
  }
  GFS_CUDA(cudaMemcpyAsync(h->d_hdr.p, hh, (size_t)batch * sizeof(PinHdr), cudaMemcpyHostToDevice, st));
  if (total) {
    GFS_CUDA(cudaMemcpyAsync(h->d_Xw.p, hX, total * 24, cudaMemcpyHostToDevice, st));
    GFS_CUDA(cudaMemcpyAsync(h->d_uvr.p, hU, total * 12, cudaMemcpyHostToDevice, st));
    GFS_CUDA(cudaMemcpyAsync(h->d_is2.p, hS, total * 4, cudaMemcpyHostToDevice, st));
    GFS_CUDA(cudaMemcpyAsync(h->d_close.p, hC, total, cudaMemcpyHostToDevice, st));
  }
  k_pose_inertial<<<batch, PIN_THREADS, 0, st>>>((const PinHdr*)h->d_hdr.p, (const double*)h->d_Xw.p, (const float*)h->d_uvr.p,
                                                 (const float*)h->d_is2.p, (const uint8_t*)h->d_close.p, (double*)h->d_err.p,
                                                 (uint8_t*)h->d_level.p, (uint8_t*)h->d_outlier.p, (float*)h->d_chi2.p,
                                                 (PinOut*)h->d_out.p);
  GFS_CUDA(cudaGetLastError());
  h->launches = 1;
  uint8_t* op = (uint8_t*)h->h_out.p;
  PinOut* ho = (PinOut*)op;
  uint8_t* hOut = op + align_up((size_t)h->maxBatch * sizeof(PinOut), 16);
  float* hChi = (float*)(hOut + align_up(NB, 16));
  GFS_CUDA(cudaMemcpyAsync(ho, h->d_out.p, (size_t)batch * sizeof(PinOut), cudaMemcpyDeviceToHost, st));
  if (total) {
    GFS_CUDA(cudaMemcpyAsync(hOut, h->d_outlier.p, total, cudaMemcpyDeviceToHost, st));
    GFS_CUDA(cudaMemcpyAsync(hChi, h->d_chi2.p, total * 4, cudaMemcpyDeviceToHost, st));
  }
  GFS_CUDA(gfs::stream_wait(st));
  off = 0;
  for (int p = 0; p < batch; p++) {
    GfsPoseInertialResult& R = results[p];
    const PinOut& O = ho[p];
    R.n_inliers = O.n_inliers; R.n_bad = O.n_bad; R.n_inliers_last = O.n_inliers_last; R.avg_reproj_error = O.avg;
    R.rounds_done = O.rounds_done;
    memcpy(R.gn_iterations, O.gn_iterations, sizeof(R.gn_iterations));
    memcpy(R.Rwb, O.Rwb, 72); memcpy(R.twb, O.twb, 24); memcpy(R.vel, O.vel, 24); memcpy(R.bg, O.bg, 24); memcpy(R.ba, O.ba, 24);
    memcpy(R.H, O.H, sizeof(R.H));
    const int n = problems[p].n_obs;
    if (n) {
      memcpy(R.outlier, hOut + off, (size_t)n);
      if (R.chi2) memcpy(R.chi2, hChi + off, (size_t)n * 4);
    }
    off += (size_t)n;
  }
  return GFS_OK;
}

int gfs_pose_inertial_optimize(GfsPoseInertial* h, void* stream, const GfsPoseInertialProblem* problem,
                               GfsPoseInertialResult* result) {
  return gfs_pose_inertial_optimize_batch(h, stream, problem, 1, result);
}

}  // extern "C"

// ORACLE -- TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of RegistrationGICP::RegisterPointClouds (reference src/RegistrationGICP.cc:5-20)
// and of the small_gicp code it runs (Thirdparty/small_gicp/include/small_gicp/...):
//   points/point_cloud.hpp:25-31          float4 -> double4, w = 1
//   util/downsampling.hpp:23-78           voxelgrid_sampling (single-thread variant: the OMP one the
//                                         reference runs is documented non-deterministic, :17-18)
//   ann/kdtree.hpp:74-131,161-233         KdTreeBuilder / knn_search, ann/knn_result.hpp:29-105,
//                                         ann/projection.hpp:18-53
//   util/normal_estimation.hpp:66-92      10-NN covariance, SelfAdjointEigenSolver::computeDirect
//   factors/gicp_factor.hpp:34-89         GICPFactor::linearize / error
//   registration/reduction_omp.hpp:21-66  sum of H, b, e
//   registration/optimizer.hpp:83-148     LevenbergMarquardtOptimizer::optimize
//   registration/termination_criteria.hpp, rejector.hpp, util/lie.hpp
// written without Eigen.  Parity status: small_gicp's own tests need fixtures that are not vendored
//  -> "parity unpinned" by the reference; pinned here by brute-force kNN, numpy
// eigen-decomposition / dense solves and finite differences in tests/test_oracle_gicp.py.
//
// Deliberate pins (reference behaviour is not deterministic to the last ulp): points of one voxel
// are summed in input order; H, b, e are summed in source-point order.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <numeric>
#include <vector>

namespace gfo {

struct P3 {
  double x, y, z;
};
struct Cov {  // symmetric 3x3: xx xy xz yy yz zz
  double c[6];
};

// ---- voxelgrid_sampling (util/downsampling.hpp:23-78, fast_floor.hpp:12-15)
static inline int fast_floor(double v) {
  int n = (int)v;
  return n - (v < (double)n);
}
static void voxelgrid(const float* pts4, int n, double leaf, std::vector<P3>& out) {
  out.clear();
  if (n == 0) return;
  const double inv = 1.0 / leaf;
  const int bits = 21, offset = 1 << (bits - 1);
  const int64_t mask = (1 << 21) - 1;
  std::vector<std::pair<uint64_t, size_t>> cp;
  cp.reserve(n);
  for (int i = 0; i < n; i++) {
    const double p[3] = {(double)pts4[4 * i], (double)pts4[4 * i + 1], (double)pts4[4 * i + 2]};
    int64_t c[3];
    bool ok = true;
    for (int a = 0; a < 3; a++) {
      c[a] = (int64_t)fast_floor(p[a] * inv) + offset;
      if (c[a] < 0 || c[a] > mask) ok = false;
    }
    if (!ok) continue;  // out-of-range points are ignored (:45-49)
    cp.push_back({(uint64_t)c[0] | ((uint64_t)c[1] << bits) | ((uint64_t)c[2] << (2 * bits)), (size_t)i});
  }
  std::sort(cp.begin(), cp.end());  // by (key, index): in-voxel order pinned to input order
  size_t i = 0;
  while (i < cp.size()) {
    double s[3] = {0, 0, 0}, w = 0;
    size_t j = i;
    for (; j < cp.size() && cp[j].first == cp[i].first; j++) {
      const float* p = pts4 + 4 * cp[j].second;
      s[0] += (double)p[0]; s[1] += (double)p[1]; s[2] += (double)p[2];
      w += 1.0;
    }
    out.push_back({s[0] / w, s[1] / w, s[2] / w});
    i = j;
  }
}

// ---- KdTree (ann/kdtree.hpp)
struct KdNode {
  uint32_t first, last;  // leaf
  int axis;
  double thresh;
  uint32_t left = 0xffffffffu, right = 0xffffffffu;
};
struct KdTree {
  const std::vector<P3>* pts = nullptr;
  std::vector<size_t> idx;
  std::vector<KdNode> nodes;
  static inline double coord(const P3& p, int a) { return a == 0 ? p.x : (a == 1 ? p.y : p.z); }

  int find_axis(size_t first, size_t last) const {  // AxisAlignedProjection::find_axis, projection.hpp:29-48
    const size_t N = last - first;
    double s[3] = {0, 0, 0}, q[3] = {0, 0, 0}, w = 0;
    const size_t step = N < 128 ? 1 : N / 128;
    const size_t ns = N / step;
    for (size_t i = 0; i < ns; i++) {
      const P3& p = (*pts)[idx[first + step * i]];
      s[0] += p.x; s[1] += p.y; s[2] += p.z; w += 1.0;
      q[0] += p.x * p.x; q[1] += p.y * p.y; q[2] += p.z * p.z;
    }
    double var[3];
    for (int a = 0; a < 3; a++) var[a] = q[a] - (s[a] / w) * s[a];
    return var[0] > var[1] ? (var[0] > var[2] ? 0 : 2) : (var[1] > var[2] ? 1 : 2);
  }
  uint32_t create(size_t first, size_t last) {
    const size_t N = last - first;
    const uint32_t ni = (uint32_t)nodes.size();
    nodes.push_back(KdNode());
    if (N <= 20) {
      nodes[ni].first = (uint32_t)first;
      nodes[ni].last = (uint32_t)last;
      return ni;
    }
    const int ax = find_axis(first, last);
    const size_t med = first + N / 2;
    std::nth_element(idx.begin() + first, idx.begin() + med, idx.begin() + last,
                     [&](size_t i, size_t j) { return coord((*pts)[i], ax) < coord((*pts)[j], ax); });
    nodes[ni].axis = ax;
    nodes[ni].thresh = coord((*pts)[idx[med]], ax);
    const uint32_t l = create(first, med);
    nodes[ni].left = l;
    const uint32_t r = create(med, last);
    nodes[ni].right = r;
    return ni;
  }
  void build(const std::vector<P3>& p) {
    pts = &p;
    idx.resize(p.size());
    std::iota(idx.begin(), idx.end(), 0);
    nodes.clear();
    nodes.reserve(p.size());
    if (!p.empty()) create(0, p.size());
  }
  // KnnResult<-1>::push (knn_result.hpp:79-98): ascending, ties keep the earlier entry first
  struct Result {
    int k, found = 0;
    size_t* ind;
    double* d;
    double worst() const { return d[k - 1]; }
    void push(size_t index, double dist) {
      if (dist >= worst()) return;
      if (k == 1) {
        ind[0] = index; d[0] = dist;
      } else {
        int loc = std::min(found, k - 1);
        for (; loc > 0 && dist < d[loc - 1]; loc--) { ind[loc] = ind[loc - 1]; d[loc] = d[loc - 1]; }
        ind[loc] = index; d[loc] = dist;
      }
      found = std::min(found + 1, k);
    }
  };
  static inline double sqdist(const P3& a, const P3& q) {
    // Eigen Vector4d squaredNorm with SSE2 packets: (dx^2 + dz^2) + (dy^2 + dw^2), dw = 0
    const double dx = a.x - q.x, dy = a.y - q.y, dz = a.z - q.z;
    return (dx * dx + dz * dz) + dy * dy;
  }
  bool search(const P3& q, uint32_t ni, Result& r) const {  // kdtree.hpp:194-233 (epsilon = 0)
    const KdNode& n = nodes[ni];
    if (n.left == 0xffffffffu) {
      for (size_t i = n.first; i < n.last; i++) r.push(idx[i], sqdist((*pts)[idx[i]], q));
      return !(r.worst() < 0.0);
    }
    const double diff = coord(q, n.axis) - n.thresh;
    const double cut = diff * diff;
    const uint32_t best = diff < 0.0 ? n.left : n.right, other = diff < 0.0 ? n.right : n.left;
    if (!search(q, best, r)) return false;
    if (r.worst() > cut) return search(q, other, r);
    return true;
  }
  int knn(const P3& q, int k, size_t* ind, double* d) const {
    for (int i = 0; i < k; i++) { ind[i] = std::numeric_limits<size_t>::max(); d[i] = std::numeric_limits<double>::max(); }
    Result r{k, 0, ind, d};
    if (!nodes.empty()) search(q, 0, r);
    return r.found;
  }
};

// ---- Eigen::SelfAdjointEigenSolver<Matrix3d>::computeDirect restated (closed-form roots + kernel
// extraction by cross products), eigenvalues ascending.  V: columns are eigenvectors (V[r][c]).
static void sym3_roots(const double m[3][3], double roots[3]) {
  const double s_inv3 = 1.0 / 3.0, s_sqrt3 = std::sqrt(3.0);
  const double c0 = m[0][0] * m[1][1] * m[2][2] + 2.0 * m[1][0] * m[2][0] * m[2][1] - m[0][0] * m[2][1] * m[2][1] -
                    m[1][1] * m[2][0] * m[2][0] - m[2][2] * m[1][0] * m[1][0];
  const double c1 = m[0][0] * m[1][1] - m[1][0] * m[1][0] + m[0][0] * m[2][2] - m[2][0] * m[2][0] + m[1][1] * m[2][2] -
                    m[2][1] * m[2][1];
  const double c2 = m[0][0] + m[1][1] + m[2][2];
  const double c2_over_3 = c2 * s_inv3;
  double a_over_3 = (c2 * c2_over_3 - c1) * s_inv3;
  a_over_3 = std::max(a_over_3, 0.0);
  const double half_b = 0.5 * (c0 + c2_over_3 * (2.0 * c2_over_3 * c2_over_3 - c1));
  double q = a_over_3 * a_over_3 * a_over_3 - half_b * half_b;
  q = std::max(q, 0.0);
  const double rho = std::sqrt(a_over_3);
  const double theta = std::atan2(std::sqrt(q), half_b) * s_inv3;
  const double ct = std::cos(theta), st = std::sin(theta);
  roots[0] = c2_over_3 - rho * (ct + s_sqrt3 * st);
  roots[1] = c2_over_3 - rho * (ct - s_sqrt3 * st);
  roots[2] = c2_over_3 + 2.0 * rho * ct;
}
static void cross3(const double a[3], const double b[3], double c[3]) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
static void extract_kernel(const double m[3][3], double res[3], double rep[3]) {
  int i0 = 0;
  double best = std::fabs(m[0][0]);
  for (int i = 1; i < 3; i++)
    if (std::fabs(m[i][i]) > best) { best = std::fabs(m[i][i]); i0 = i; }
  double col[3][3];
  for (int c = 0; c < 3; c++)
    for (int r = 0; r < 3; r++) col[c][r] = m[r][c];
  for (int r = 0; r < 3; r++) rep[r] = col[i0][r];
  double c0[3], c1[3];
  cross3(rep, col[(i0 + 1) % 3], c0);
  cross3(rep, col[(i0 + 2) % 3], c1);
  const double n0 = c0[0] * c0[0] + c0[1] * c0[1] + c0[2] * c0[2];
  const double n1 = c1[0] * c1[0] + c1[1] * c1[1] + c1[2] * c1[2];
  if (n0 > n1) { const double s = std::sqrt(n0); for (int r = 0; r < 3; r++) res[r] = c0[r] / s; }
  else { const double s = std::sqrt(n1); for (int r = 0; r < 3; r++) res[r] = c1[r] / s; }
}
static void eig3_direct(const double A[3][3], double evals[3], double V[3][3]) {
  const double eps = std::numeric_limits<double>::epsilon();
  double m[3][3];
  const double shift = (A[0][0] + A[1][1] + A[2][2]) / 3.0;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) m[r][c] = (r >= c) ? A[r][c] : A[c][r];  // selfadjointView<Lower>
  for (int i = 0; i < 3; i++) m[i][i] -= shift;
  double scale = 0;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) scale = std::max(scale, std::fabs(m[r][c]));
  if (scale > 0)
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) m[r][c] /= scale;
  sym3_roots(m, evals);
  double v[3][3];  // v[k] = eigenvector k
  if ((evals[2] - evals[0]) <= eps) {
    for (int k = 0; k < 3; k++)
      for (int r = 0; r < 3; r++) v[k][r] = (k == r) ? 1.0 : 0.0;
  } else {
    double tmp[3][3];
    double d0 = evals[2] - evals[1], d1 = evals[1] - evals[0];
    int k = 0, l = 2;
    if (d0 > d1) { std::swap(k, l); d0 = d1; }
    memcpy(tmp, m, sizeof(tmp));
    for (int i = 0; i < 3; i++) tmp[i][i] -= evals[k];
    extract_kernel(tmp, v[k], v[l]);
    if (d0 <= 2 * eps * d1) {
      const double dot = v[k][0] * v[l][0] + v[k][1] * v[l][1] + v[k][2] * v[l][2];
      for (int r = 0; r < 3; r++) v[l][r] -= dot * v[l][r];
      const double nn = std::sqrt(v[l][0] * v[l][0] + v[l][1] * v[l][1] + v[l][2] * v[l][2]);
      for (int r = 0; r < 3; r++) v[l][r] /= nn;
    } else {
      memcpy(tmp, m, sizeof(tmp));
      for (int i = 0; i < 3; i++) tmp[i][i] -= evals[l];
      double dummy[3];
      extract_kernel(tmp, v[l], dummy);
    }
    double c[3];
    cross3(v[2], v[0], c);
    const double nn = std::sqrt(c[0] * c[0] + c[1] * c[1] + c[2] * c[2]);
    for (int r = 0; r < 3; r++) v[1][r] = c[r] / nn;
  }
  for (int k = 0; k < 3; k++) evals[k] = evals[k] * scale + shift;
  for (int k = 0; k < 3; k++)
    for (int r = 0; r < 3; r++) V[r][k] = v[k][r];
}

// ---- estimate_local_features<CovarianceSetter> (util/normal_estimation.hpp:28-44,66-92)
static void covariance_from_neighbors(const std::vector<P3>& pts, const size_t* ind, int n, Cov& out) {
  if (n < 5) {
    out = Cov{{1, 0, 0, 1, 0, 1}};
    return;
  }
  double s[3] = {0, 0, 0}, cr[3][3] = {{0}};
  for (int i = 0; i < n; i++) {
    const P3& p = pts[ind[i]];
    const double v[3] = {p.x, p.y, p.z};
    for (int a = 0; a < 3; a++) {
      s[a] += v[a];
      for (int b = 0; b < 3; b++) cr[a][b] += v[a] * v[b];
    }
  }
  double C[3][3];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) C[a][b] = (cr[a][b] - (s[a] / n) * s[b]) / n;
  double ev[3], V[3][3];
  eig3_direct(C, ev, V);
  const double val[3] = {1e-3, 1.0, 1.0};
  double R[3][3];
  for (int a = 0; a < 3; a++)
    for (int b = 0; b < 3; b++) {
      // (V * diag) * V^T, evaluated left to right
      double acc = 0;
      for (int k = 0; k < 3; k++) acc += (V[a][k] * val[k]) * V[b][k];
      R[a][b] = acc;
    }
  out = Cov{{R[0][0], R[0][1], R[0][2], R[1][1], R[1][2], R[2][2]}};
}

struct Cloud {
  std::vector<P3> pts;
  std::vector<Cov> cov;
  KdTree tree;
};

static void preprocess(const float* pts4, int n, double leaf, int k, int threads, Cloud& c) {
  voxelgrid(pts4, n, leaf, c.pts);
  c.tree.build(c.pts);
  c.cov.resize(c.pts.size());
#pragma omp parallel for num_threads(threads) schedule(static)
  for (int64_t i = 0; i < (int64_t)c.pts.size(); i++) {
    size_t ind[32];
    double d[32];
    const int found = c.tree.knn(c.pts[i], k, ind, d);
    covariance_from_neighbors(c.pts, ind, found, c.cov[i]);
  }
}

// ---- small fixed-size algebra
struct T44 {
  double R[3][3], t[3];
};
static inline void transform(const T44& T, const P3& p, double o[3]) {
  for (int r = 0; r < 3; r++) o[r] = ((T.R[r][0] * p.x + T.R[r][1] * p.y) + T.R[r][2] * p.z) + T.t[r];
}
static bool inv3(const double A[3][3], double I[3][3]) {
  const double c00 = A[1][1] * A[2][2] - A[1][2] * A[2][1], c01 = A[1][2] * A[2][0] - A[1][0] * A[2][2],
               c02 = A[1][0] * A[2][1] - A[1][1] * A[2][0];
  const double det = A[0][0] * c00 + A[0][1] * c01 + A[0][2] * c02;
  const double id = 1.0 / det;
  I[0][0] = c00 * id; I[1][0] = c01 * id; I[2][0] = c02 * id;
  I[0][1] = (A[0][2] * A[2][1] - A[0][1] * A[2][2]) * id;
  I[1][1] = (A[0][0] * A[2][2] - A[0][2] * A[2][0]) * id;
  I[2][1] = (A[0][1] * A[2][0] - A[0][0] * A[2][1]) * id;
  I[0][2] = (A[0][1] * A[1][2] - A[0][2] * A[1][1]) * id;
  I[1][2] = (A[0][2] * A[1][0] - A[0][0] * A[1][2]) * id;
  I[2][2] = (A[0][0] * A[1][1] - A[0][1] * A[1][0]) * id;
  return det != 0;
}
// solve (H) x = rhs, H 6x6 SPD, LDL^T without pivoting
static void ldlt_solve6(const double Hin[6][6], const double rhs[6], double x[6]) {
  double L[6][6] = {{0}}, D[6];
  for (int j = 0; j < 6; j++) {
    double d = Hin[j][j];
    for (int k = 0; k < j; k++) d -= L[j][k] * L[j][k] * D[k];
    D[j] = d;
    L[j][j] = 1;
    for (int i = j + 1; i < 6; i++) {
      double v = Hin[i][j];
      for (int k = 0; k < j; k++) v -= L[i][k] * L[j][k] * D[k];
      L[i][j] = v / d;
    }
  }
  double y[6];
  for (int i = 0; i < 6; i++) {
    double v = rhs[i];
    for (int k = 0; k < i; k++) v -= L[i][k] * y[k];
    y[i] = v;
  }
  for (int i = 0; i < 6; i++) y[i] /= D[i];
  for (int i = 5; i >= 0; i--) {
    double v = y[i];
    for (int k = i + 1; k < 6; k++) v -= L[k][i] * x[k];
    x[i] = v;
  }
}
// se3_exp (util/lie.hpp:52-96): rotation via quaternion so3_exp, translation via V
static void se3_exp(const double a[6], T44& out) {
  const double w[3] = {a[0], a[1], a[2]};
  const double theta_sq = w[0] * w[0] + w[1] * w[1] + w[2] * w[2];
  double imag, real;
  if (theta_sq < 1e-10) {
    const double tq = theta_sq * theta_sq;
    imag = 0.5 - 1.0 / 48.0 * theta_sq + 1.0 / 3840.0 * tq;
    real = 1.0 - 1.0 / 8.0 * theta_sq + 1.0 / 384.0 * tq;
  } else {
    const double theta = std::sqrt(theta_sq), half = 0.5 * theta;
    imag = std::sin(half) / theta;
    real = std::cos(half);
  }
  const double qw = real, qx = imag * w[0], qy = imag * w[1], qz = imag * w[2];
  // Eigen::Quaternion::toRotationMatrix
  const double tx = 2 * qx, ty = 2 * qy, tz = 2 * qz;
  const double twx = tx * qw, twy = ty * qw, twz = tz * qw, txx = tx * qx, txy = ty * qx, txz = tz * qx, tyy = ty * qy,
               tyz = tz * qy, tzz = tz * qz;
  out.R[0][0] = 1 - (tyy + tzz); out.R[0][1] = txy - twz; out.R[0][2] = txz + twy;
  out.R[1][0] = txy + twz; out.R[1][1] = 1 - (txx + tzz); out.R[1][2] = tyz - twx;
  out.R[2][0] = txz - twy; out.R[2][1] = tyz + twx; out.R[2][2] = 1 - (txx + tyy);
  const double theta = std::sqrt(theta_sq);
  const double tr[3] = {a[3], a[4], a[5]};
  if (theta < 1e-10) {
    for (int r = 0; r < 3; r++) out.t[r] = out.R[r][0] * tr[0] + out.R[r][1] * tr[1] + out.R[r][2] * tr[2];
  } else {
    const double O[3][3] = {{0, -w[2], w[1]}, {w[2], 0, -w[0]}, {-w[1], w[0], 0}};
    const double k1 = (1.0 - std::cos(theta)) / theta_sq, k2 = (theta - std::sin(theta)) / (theta_sq * theta);
    double V[3][3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        double oo = 0;
        for (int k = 0; k < 3; k++) oo += O[r][k] * O[k][c];
        V[r][c] = (r == c ? 1.0 : 0.0) + k1 * O[r][c] + k2 * oo;
      }
    for (int r = 0; r < 3; r++) out.t[r] = V[r][0] * tr[0] + V[r][1] * tr[1] + V[r][2] * tr[2];
  }
}
static void compose(const T44& A, const T44& B, T44& C) {  // C = A * B
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) C.R[r][c] = A.R[r][0] * B.R[0][c] + A.R[r][1] * B.R[1][c] + A.R[r][2] * B.R[2][c];
    C.t[r] = A.R[r][0] * B.t[0] + A.R[r][1] * B.t[1] + A.R[r][2] * B.t[2] + A.t[r];
  }
}

struct Factor {
  size_t target = std::numeric_limits<size_t>::max();
  double M[3][3] = {{0}};
};
struct Contribution {
  double H[21], b[6], e;
  int inlier;
};

// GICPFactor::linearize (factors/gicp_factor.hpp:34-73)
static void linearize_one(const Cloud& tgt, const Cloud& src, const T44& T, size_t i, double max_d2, Factor& f,
                          Contribution& c) {
  f.target = std::numeric_limits<size_t>::max();
  c.inlier = 0;
  double q[3];
  transform(T, src.pts[i], q);
  size_t ki;
  double kd;
  if (tgt.tree.knn(P3{q[0], q[1], q[2]}, 1, &ki, &kd) == 0 || kd > max_d2) return;
  f.target = ki;
  const double* cs = src.cov[i].c;
  const double* ct = tgt.cov[ki].c;
  const double Cs[3][3] = {{cs[0], cs[1], cs[2]}, {cs[1], cs[3], cs[4]}, {cs[2], cs[4], cs[5]}};
  const double Ct[3][3] = {{ct[0], ct[1], ct[2]}, {ct[1], ct[3], ct[4]}, {ct[2], ct[4], ct[5]}};
  double RC[3][3], RCR[3][3];
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < 3; k++) RC[r][k] = T.R[r][0] * Cs[0][k] + T.R[r][1] * Cs[1][k] + T.R[r][2] * Cs[2][k];
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < 3; k++) RCR[r][k] = Ct[r][k] + (RC[r][0] * T.R[k][0] + RC[r][1] * T.R[k][1] + RC[r][2] * T.R[k][2]);
  inv3(RCR, f.M);
  const P3& pt = tgt.pts[ki];
  const double res[3] = {pt.x - q[0], pt.y - q[1], pt.z - q[2]};
  const P3& ps = src.pts[i];
  const double S[3][3] = {{0, -ps.z, ps.y}, {ps.z, 0, -ps.x}, {-ps.y, ps.x, 0}};
  double J[3][6];
  for (int r = 0; r < 3; r++)
    for (int k = 0; k < 3; k++) {
      J[r][k] = T.R[r][0] * S[0][k] + T.R[r][1] * S[1][k] + T.R[r][2] * S[2][k];
      J[r][3 + k] = -T.R[r][k];
    }
  double MJ[3][6], Mr[3];
  for (int r = 0; r < 3; r++) {
    for (int k = 0; k < 6; k++) MJ[r][k] = f.M[r][0] * J[0][k] + f.M[r][1] * J[1][k] + f.M[r][2] * J[2][k];
    Mr[r] = f.M[r][0] * res[0] + f.M[r][1] * res[1] + f.M[r][2] * res[2];
  }
  int n = 0;
  for (int a = 0; a < 6; a++)
    for (int b2 = a; b2 < 6; b2++) c.H[n++] = J[0][a] * MJ[0][b2] + J[1][a] * MJ[1][b2] + J[2][a] * MJ[2][b2];
  for (int a = 0; a < 6; a++) c.b[a] = J[0][a] * Mr[0] + J[1][a] * Mr[1] + J[2][a] * Mr[2];
  c.e = 0.5 * (res[0] * Mr[0] + res[1] * Mr[1] + res[2] * Mr[2]);
  c.inlier = 1;
}
// GICPFactor::error (:82-89)
static double error_one(const Cloud& tgt, const Cloud& src, const T44& T, size_t i, const Factor& f) {
  if (f.target == std::numeric_limits<size_t>::max()) return 0.0;
  double q[3];
  transform(T, src.pts[i], q);
  const P3& pt = tgt.pts[f.target];
  const double res[3] = {pt.x - q[0], pt.y - q[1], pt.z - q[2]};
  double Mr[3];
  for (int r = 0; r < 3; r++) Mr[r] = f.M[r][0] * res[0] + f.M[r][1] * res[1] + f.M[r][2] * res[2];
  return 0.5 * (res[0] * Mr[0] + res[1] * Mr[1] + res[2] * Mr[2]);
}

struct Result {
  double T[16], H[36], b[6], error;
  int iterations, num_inliers, converged;
  int n_target, n_source, inner_evals;
};

// LevenbergMarquardtOptimizer::optimize (registration/optimizer.hpp:83-148)
static void align(const Cloud& tgt, const Cloud& src, const double T0[16], double max_dist, int max_iter, double rot_eps,
                  double trans_eps, int threads, Result& out) {
  T44 T;
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) T.R[r][c] = T0[4 * r + c];
    T.t[r] = T0[4 * r + 3];
  }
  const size_t M = src.pts.size();
  std::vector<Factor> factors(M);
  std::vector<Contribution> contrib(M);
  std::vector<double> errs(M);
  double lambda = 1e-3;
  const double max_d2 = max_dist * max_dist;
  bool converged = false;
  double H[6][6] = {{0}}, b[6] = {0}, e = 0;
  int iterations = 0, inner = 0;
  for (int i = 0; i < max_iter && !converged; i++) {
#pragma omp parallel for num_threads(threads) schedule(guided, 8)
    for (int64_t k = 0; k < (int64_t)M; k++) linearize_one(tgt, src, T, k, max_d2, factors[k], contrib[k]);
    double Hs[21] = {0};
    for (int a = 0; a < 6; a++) b[a] = 0;
    e = 0;
    for (size_t k = 0; k < M; k++) {
      if (!contrib[k].inlier) continue;
      for (int a = 0; a < 21; a++) Hs[a] += contrib[k].H[a];
      for (int a = 0; a < 6; a++) b[a] += contrib[k].b[a];
      e += contrib[k].e;
    }
    int n = 0;
    for (int a = 0; a < 6; a++)
      for (int c = a; c < 6; c++) H[a][c] = H[c][a] = Hs[n++];
    bool success = false;
    for (int j = 0; j < 10; j++) {
      double Hd[6][6], nb[6], delta[6];
      for (int a = 0; a < 6; a++) {
        for (int c = 0; c < 6; c++) Hd[a][c] = H[a][c] + (a == c ? lambda : 0.0);
        nb[a] = -b[a];
      }
      ldlt_solve6(Hd, nb, delta);
      T44 dT, newT;
      se3_exp(delta, dT);
      compose(T, dT, newT);
#pragma omp parallel for num_threads(threads) schedule(guided, 8)
      for (int64_t k = 0; k < (int64_t)M; k++) errs[k] = error_one(tgt, src, newT, k, factors[k]);
      double new_e = 0;
      for (size_t k = 0; k < M; k++) new_e += errs[k];
      inner++;
      if (new_e <= e) {
        const double dr = std::sqrt(delta[0] * delta[0] + delta[1] * delta[1] + delta[2] * delta[2]);
        const double dt = std::sqrt(delta[3] * delta[3] + delta[4] * delta[4] + delta[5] * delta[5]);
        converged = dr <= rot_eps && dt <= trans_eps;
        T = newT;
        lambda /= 10.0;
        success = true;
        break;
      } else {
        lambda *= 10.0;
      }
    }
    iterations = i;
    if (!success) break;
  }
  for (int r = 0; r < 3; r++) {
    for (int c = 0; c < 3; c++) out.T[4 * r + c] = T.R[r][c];
    out.T[4 * r + 3] = T.t[r];
  }
  out.T[12] = out.T[13] = out.T[14] = 0; out.T[15] = 1;
  for (int a = 0; a < 6; a++) {
    for (int c = 0; c < 6; c++) out.H[6 * a + c] = H[a][c];
    out.b[a] = b[a];
  }
  out.error = e;
  out.iterations = iterations;
  out.converged = converged;
  int ninl = 0;
  for (size_t k = 0; k < M; k++) ninl += factors[k].target != std::numeric_limits<size_t>::max();
  out.num_inliers = ninl;
  out.n_target = (int)tgt.pts.size();
  out.n_source = (int)M;
  out.inner_evals = inner;
}

}  // namespace gfo

extern "C" {

// RegistrationGICP::RegisterPointClouds(target, source, init_T): points are float4 (x, y, z, 1).
// T row-major 4x4.  setting: voxel 0.02, max corr. dist 0.1, k = 10, 20 iterations, rot eps 0.1 deg,
// trans eps 1e-3 (RegistrationGICP.cc:9-15; registration_helper.hpp:41-49).
void gfo_gicp_align(const float* target4, int nt, const float* source4, int ns, const double* T0, double voxel,
                    double max_dist, int k, int max_iter, double rot_eps, double trans_eps, int threads,
                    gfo::Result* out) {
  gfo::Cloud tgt, src;
  gfo::preprocess(target4, nt, voxel, k, threads, tgt);
  gfo::preprocess(source4, ns, voxel, k, threads, src);
  gfo::align(tgt, src, T0, max_dist, max_iter, rot_eps, trans_eps, threads, *out);
}

// stage hooks
int gfo_voxelgrid(const float* pts4, int n, double leaf, double* out_xyz, int cap) {
  std::vector<gfo::P3> o;
  gfo::voxelgrid(pts4, n, leaf, o);
  const int m = (int)std::min<size_t>(o.size(), (size_t)cap);
  for (int i = 0; i < m; i++) { out_xyz[3 * i] = o[i].x; out_xyz[3 * i + 1] = o[i].y; out_xyz[3 * i + 2] = o[i].z; }
  return (int)o.size();
}
// kNN through the restated KdTree; queries/points are double xyz
void gfo_kdtree_knn(const double* pts_xyz, int n, const double* q_xyz, int nq, int k, int64_t* out_idx, double* out_d2,
                    int* out_found) {
  std::vector<gfo::P3> p(n);
  for (int i = 0; i < n; i++) p[i] = {pts_xyz[3 * i], pts_xyz[3 * i + 1], pts_xyz[3 * i + 2]};
  gfo::KdTree t;
  t.build(p);
  std::vector<size_t> ind(k);
  for (int i = 0; i < nq; i++) {
    const int f = t.knn(gfo::P3{q_xyz[3 * i], q_xyz[3 * i + 1], q_xyz[3 * i + 2]}, k, ind.data(), out_d2 + (size_t)i * k);
    for (int j = 0; j < k; j++) out_idx[(size_t)i * k + j] = j < f ? (int64_t)ind[j] : -1;
    out_found[i] = f;
  }
}
// covariances of a (downsampled) cloud: out 6 doubles per point (xx xy xz yy yz zz)
void gfo_covariances(const double* pts_xyz, int n, int k, double* out_cov6) {
  std::vector<gfo::P3> p(n);
  for (int i = 0; i < n; i++) p[i] = {pts_xyz[3 * i], pts_xyz[3 * i + 1], pts_xyz[3 * i + 2]};
  gfo::KdTree t;
  t.build(p);
  for (int i = 0; i < n; i++) {
    size_t ind[32];
    double d[32];
    const int f = t.knn(p[i], k, ind, d);
    gfo::Cov c;
    gfo::covariance_from_neighbors(p, ind, f, c);
    memcpy(out_cov6 + 6 * (size_t)i, c.c, sizeof(c.c));
  }
}
void gfo_eig3(const double* A9, double* evals3, double* V9) {
  double A[3][3], V[3][3];
  memcpy(A, A9, sizeof(A));
  gfo::eig3_direct(A, evals3, V);
  memcpy(V9, V, sizeof(V));
}
void gfo_se3_exp(const double* a6, double* T16) {
  gfo::T44 T;
  gfo::se3_exp(a6, T);
  for (int r = 0; r < 3; r++) { for (int c = 0; c < 3; c++) T16[4 * r + c] = T.R[r][c]; T16[4 * r + 3] = T.t[r]; }
  T16[12] = T16[13] = T16[14] = 0; T16[15] = 1;
}
}

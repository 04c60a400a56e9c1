// ORACLE -- TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of the reference's descriptor matching path:
//   * ORBmatcher::DescriptorDistance                      src/ORBmatcher.cc:2536-2550
//   * cv::BFMatcher(NORM_HAMMING).match(d1, d2)           call sites src/ORBmatcher.cc:755-756,805-806,888-889
//   * gms_matcher (GetInlierMask(false,false) -> run(1))  Thirdparty/GMS/include/gms_matcher.h
// Parity status: the reference has no tests/golden vectors for this path ("parity unpinned" by the
// reference itself); BF matching is pinned against cv2.BFMatcher in tests/test_oracle_match.py.
#include <cmath>
#include <cstdint>
#include <cstring>
#include <utility>
#include <vector>

namespace gfo {

// ORBmatcher::DescriptorDistance, src/ORBmatcher.cc:2536-2550 (SWAR popcount over 8 words)
static inline int descriptor_distance(const uint8_t* a, const uint8_t* b) {
  int dist = 0;
  for (int i = 0; i < 8; i++) {
    uint32_t x, y;
    memcpy(&x, a + 4 * i, 4);
    memcpy(&y, b + 4 * i, 4);
    uint32_t v = x ^ y;
    v = v - ((v >> 1) & 0x55555555);
    v = (v & 0x33333333) + ((v >> 2) & 0x33333333);
    dist += (((v + (v >> 4)) & 0xF0F0F0F) * 0x1010101) >> 24;
  }
  return dist;
}

// BFMatcher::match without cross-check: one match per query row, argmin over train rows,
// ties resolved to the lowest train index.
static void bf_match(const uint8_t* dq, int nq, const uint8_t* dt, int nt, int* idx, int* dist, int threads) {
#pragma omp parallel for num_threads(threads)
  for (int q = 0; q < nq; q++) {
    int best = 1 << 30, bi = -1;
    for (int t = 0; t < nt; t++) {
      int d = descriptor_distance(dq + 32 * (size_t)q, dt + 32 * (size_t)t);
      if (d < best) { best = d; bi = t; }
    }
    idx[q] = bi;
    dist[q] = bi < 0 ? -1 : best;
  }
}

// gms_matcher, Thirdparty/GMS/include/gms_matcher.h:49-64 (ctor), :236-246 (GetInlierMask no
// scale / no rotation), :305-331 (AssignMatchPairs), :334-383 (VerifyCellPairs), :385-419 (run).
// 20x20 left grid, right grid = left grid (scale index 0), rotation pattern 1 (identity).
struct Gms {
  static const int G = 20, NG = 400;
  std::vector<float> p1, p2;  // normalised points
  std::vector<std::pair<int, int>> matches;

  static int nb9(int idx, int k) {  // GetNB9, :166-186 (k = 0..8 row-major 3x3, -1 outside)
    int x = idx % G, y = idx / G;
    int xi = k % 3 - 1, yi = k / 3 - 1;
    int xx = x + xi, yy = y + yi;
    if (xx < 0 || xx >= G || yy < 0 || yy >= G) return -1;
    return xx + yy * G;
  }

  static int grid_left(float px, float py, int type) {  // GetGridIndexLeft, :125-153
    int x = 0, y = 0;
    if (type == 1) { x = (int)std::floor(px * G); y = (int)std::floor(py * G); }
    if (type == 2) { x = (int)std::floor(px * G + 0.5); y = (int)std::floor(py * G); }
    if (type == 3) { x = (int)std::floor(px * G); y = (int)std::floor(py * G + 0.5); }
    if (type == 4) { x = (int)std::floor(px * G + 0.5); y = (int)std::floor(py * G + 0.5); }
    if (x >= G || y >= G) return -1;
    return x + y * G;
  }
  static int grid_right(float px, float py) {  // GetGridIndexRight, :155-160
    int x = (int)std::floor(px * G);
    int y = (int)std::floor(py * G);
    return x + y * G;
  }

  int run(std::vector<uint8_t>& mask) {
    const size_t nm = matches.size();
    mask.assign(nm, 0);
    std::vector<int> stat((size_t)NG * NG);
    std::vector<std::pair<int, int>> mp(nm, std::make_pair(0, 0));
    for (int type = 1; type <= 4; type++) {
      std::fill(stat.begin(), stat.end(), 0);
      std::vector<int> cellPairs(NG, -1), nLeft(NG, 0);
      for (size_t i = 0; i < nm; i++) {  // AssignMatchPairs
        const float* lp = &p1[2 * matches[i].first];
        const float* rp = &p2[2 * matches[i].second];
        int l = mp[i].first = grid_left(lp[0], lp[1], type);
        int r;
        if (type == 1) r = mp[i].second = grid_right(rp[0], rp[1]);
        else r = mp[i].second;
        if (l < 0 || r < 0) continue;
        if (l >= NG || r >= NG) continue;
        stat[(size_t)l * NG + r]++;
        nLeft[l]++;
      }
      for (int i = 0; i < NG; i++) {  // VerifyCellPairs(RotationType = 1)
        const int* row = &stat[(size_t)i * NG];
        long s = 0;
        for (int j = 0; j < NG; j++) s += row[j];
        if (s == 0) { cellPairs[i] = -1; continue; }
        int mx = 0;
        for (int j = 0; j < NG; j++)
          if (row[j] > mx) { cellPairs[i] = j; mx = row[j]; }
        int rt = cellPairs[i];
        int score = 0, numpair = 0;
        double thresh = 0;
        for (int k = 0; k < 9; k++) {
          int ll = nb9(i, k), rr = nb9(rt, k);
          if (ll == -1 || rr == -1) continue;
          score += stat[(size_t)ll * NG + rr];
          thresh += nLeft[ll];
          numpair++;
        }
        thresh = 6 * std::sqrt(thresh / numpair);  // THRESH_FACTOR
        if (score < thresh) cellPairs[i] = -2;
      }
      for (size_t i = 0; i < nm; i++) {
        // The reference indexes mCellPairs[-1] when the left point fell outside a shifted grid
        // (undefined behaviour, :406).  ORB keypoints never get there (x/w, y/h < 0.975); the
        // oracle defines that case as "not an inlier for this grid type".
        if (mp[i].first < 0) continue;
        if (cellPairs[mp[i].first] == mp[i].second) mask[i] = 1;
      }
    }
    int n = 0;
    for (size_t i = 0; i < nm; i++) n += mask[i];
    return n;
  }
};

}  // namespace gfo

extern "C" {

int gfo_descriptor_distance(const uint8_t* a, const uint8_t* b) { return gfo::descriptor_distance(a, b); }

void gfo_bf_match(const uint8_t* dq, int nq, const uint8_t* dt, int nt, int* idx, int* dist, int threads) {
  gfo::bf_match(dq, nq, dt, nt, idx, dist, threads < 1 ? 1 : threads);
}

// pts: interleaved float x,y (pixels); matches: interleaved int (queryIdx, trainIdx).
int gfo_gms_filter(const float* pts1, int n1, int w1, int h1, const float* pts2, int n2, int w2, int h2,
                   const int* matches, int nm, uint8_t* inlier) {
  gfo::Gms g;
  g.p1.resize(2 * (size_t)n1);
  g.p2.resize(2 * (size_t)n2);
  for (int i = 0; i < n1; i++) { g.p1[2 * i] = pts1[2 * i] / w1; g.p1[2 * i + 1] = pts1[2 * i + 1] / h1; }  // NormalizePoints :104-115
  for (int i = 0; i < n2; i++) { g.p2[2 * i] = pts2[2 * i] / w2; g.p2[2 * i + 1] = pts2[2 * i + 1] / h2; }
  g.matches.resize(nm);
  for (int i = 0; i < nm; i++) g.matches[i] = std::make_pair(matches[2 * i], matches[2 * i + 1]);
  std::vector<uint8_t> mask;
  int n = g.run(mask);
  if (nm) memcpy(inlier, mask.data(), nm);
  return n;
}
}

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

// =================================================================================================
// Projection-window search (ORBmatcher::SearchByProjection, both overloads) on flattened inputs,
// with the Frame grid helpers it relies on.  RGB-D / monocular layout (Frame::Nleft == -1).
//   Frame::AssignFeaturesToGrid / PosInGrid        src/Frame.cc:734-761, 1073-1084 (64 x 48 cells)
//   Frame::GetFeaturesInArea                       src/Frame.cc:1007-1071
//   SearchByProjection(Frame&, vector<MapPoint*>&) src/ORBmatcher.cc:43-207   (mode 1)
//   SearchByProjection(Frame&, const Frame&, ...)  src/ORBmatcher.cc:1853-2063 (mode 0)
//   ComputeThreeMaxima                             src/ORBmatcher.cc:2500-2532
// =================================================================================================
#include <algorithm>
namespace gfo {

struct ProjQuery {  // == GfsProjQuery (include/gfs_b200.h)
  float u, v, radius, ur, angle;
  int min_level, max_level, blocks;
  uint8_t desc[32];
};
struct KeyPt {
  float x, y, size, angle, response;
  int octave;
};
static const int GRID_COLS = 64, GRID_ROWS = 48, TH_HIGH = 100, HISTO_LENGTH = 30;

struct FrameGrid {
  std::vector<int> cells[GRID_COLS][GRID_ROWS];
  float minX, minY, invW, invH;
  void build(const KeyPt* k, int n) {
    for (int i = 0; i < n; i++) {
      const int px = (int)std::round((k[i].x - minX) * invW), py = (int)std::round((k[i].y - minY) * invH);
      if (px < 0 || px >= GRID_COLS || py < 0 || py >= GRID_ROWS) continue;
      cells[px][py].push_back(i);
    }
  }
  void in_area(const KeyPt* k, float x, float y, float r, int minLevel, int maxLevel, std::vector<int>& out) const {
    out.clear();
    const int nMinCellX = std::max(0, (int)std::floor((x - minX - r) * invW));
    if (nMinCellX >= GRID_COLS) return;
    const int nMaxCellX = std::min(GRID_COLS - 1, (int)std::ceil((x - minX + r) * invW));
    if (nMaxCellX < 0) return;
    const int nMinCellY = std::max(0, (int)std::floor((y - minY - r) * invH));
    if (nMinCellY >= GRID_ROWS) return;
    const int nMaxCellY = std::min(GRID_ROWS - 1, (int)std::ceil((y - minY + r) * invH));
    if (nMaxCellY < 0) return;
    const bool check = (minLevel > 0) || (maxLevel >= 0);
    for (int ix = nMinCellX; ix <= nMaxCellX; ix++)
      for (int iy = nMinCellY; iy <= nMaxCellY; iy++)
        for (int j : cells[ix][iy]) {
          if (check) {
            if (k[j].octave < minLevel) continue;
            if (maxLevel >= 0 && k[j].octave > maxLevel) continue;
          }
          const float dx = k[j].x - x, dy = k[j].y - y;
          if (std::fabs(dx) < r && std::fabs(dy) < r) out.push_back(j);
        }
  }
};

static int search_by_projection(int mode, float nnratio, int checkOri, const ProjQuery* q, int nq, const KeyPt* kps,
                                const float* uRight, const uint8_t* desc, const uint8_t* occupied, int n, float minX,
                                float minY, float invW, float invH, int* assign) {
  FrameGrid* g = new FrameGrid();
  g->minX = minX; g->minY = minY; g->invW = invW; g->invH = invH;
  g->build(kps, n);
  std::vector<int> blocked(n), cand;
  for (int i = 0; i < n; i++) { assign[i] = -1; blocked[i] = occupied ? occupied[i] != 0 : 0; }
  std::vector<std::vector<int>> rotHist(HISTO_LENGTH);
  const float factor = 1.0f / HISTO_LENGTH;
  int nmatches = 0;
  for (int i = 0; i < nq; i++) {
    const ProjQuery& Q = q[i];
    if (Q.radius < 0) continue;  // query not in view / filtered by the caller
    g->in_area(kps, Q.u, Q.v, Q.radius, Q.min_level, Q.max_level, cand);
    if (cand.empty()) continue;
    int bestDist = 256, bestLevel = -1, bestDist2 = 256, bestLevel2 = -1, bestIdx = -1;
    for (int idx : cand) {
      if (blocked[idx]) continue;
      if (uRight[idx] > 0) {
        const float er = std::fabs(Q.ur - uRight[idx]);
        if (er > Q.radius) continue;
      }
      const int dist = descriptor_distance(Q.desc, desc + 32 * (size_t)idx);
      if (dist < bestDist) {
        bestDist2 = bestDist; bestDist = dist; bestLevel2 = bestLevel; bestLevel = kps[idx].octave; bestIdx = idx;
      } else if (mode == 1 && dist < bestDist2) {
        bestLevel2 = kps[idx].octave; bestDist2 = dist;
      }
    }
    if (bestDist <= TH_HIGH) {
      if (mode == 1) {
        if (bestLevel == bestLevel2 && bestDist > nnratio * bestDist2) continue;
        if (bestLevel != bestLevel2 || bestDist <= nnratio * bestDist2) {
          assign[bestIdx] = i;
          blocked[bestIdx] = Q.blocks != 0;
          nmatches++;
        }
      } else {
        assign[bestIdx] = i;
        blocked[bestIdx] = Q.blocks != 0;
        nmatches++;
        if (checkOri) {
          float rot = Q.angle - kps[bestIdx].angle;
          if (rot < 0.0) rot += 360.0f;
          int bin = (int)std::round(rot * factor);
          if (bin == HISTO_LENGTH) bin = 0;
          rotHist[bin].push_back(bestIdx);
        }
      }
    }
  }
  if (mode == 0 && checkOri) {
    int max1 = 0, max2 = 0, max3 = 0, ind1 = -1, ind2 = -1, ind3 = -1;
    for (int i = 0; i < HISTO_LENGTH; i++) {
      const int s = (int)rotHist[i].size();
      if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
      else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
      else if (s > max3) { max3 = s; ind3 = i; }
    }
    if (max2 < 0.1f * (float)max1) { ind2 = -1; ind3 = -1; }
    else if (max3 < 0.1f * (float)max1) { ind3 = -1; }
    for (int i = 0; i < HISTO_LENGTH; i++)
      if (i != ind1 && i != ind2 && i != ind3)
        for (int idx : rotHist[i]) { assign[idx] = -1; nmatches--; }
  }
  delete g;
  return nmatches;
}

}  // namespace gfo

extern "C" {
int gfo_search_by_projection(int mode, float nnratio, int checkOri, const void* queries, int nq, const void* kps,
                             const float* uRight, const uint8_t* desc, const uint8_t* occupied, int n, float minX, float minY,
                             float invW, float invH, int* assign) {
  return gfo::search_by_projection(mode, nnratio, checkOri, (const gfo::ProjQuery*)queries, nq, (const gfo::KeyPt*)kps, uRight,
                                   desc, occupied, n, minX, minY, invW, invH, assign);
}

// Frame::ConvertDepthToPointCloud (src/Frame.cc:590-623): every `stride`-th pixel with 0 < d < 10,
// scan order, float32 back-projection.  Returns the number of points.
int gfo_depth_to_cloud(const float* depth, int w, int h, int stride, float fx, float fy, float cx, float cy, float* out4,
                       int cap) {
  int n = 0;
  for (int v = 0; v < h; v += stride)
    for (int u = 0; u < w; u += stride) {
      const float d = depth[(size_t)v * w + u];
      if (d > 0.0 && d < 10.0) {
        if (n < cap) {
          out4[4 * n] = (u - cx) * d / fx;
          out4[4 * n + 1] = (v - cy) * d / fy;
          out4[4 * n + 2] = d;
          out4[4 * n + 3] = 1.0f;
        }
        n++;
      }
    }
  return n;
}
}

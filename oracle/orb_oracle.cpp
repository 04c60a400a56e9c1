// ORACLE -- TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of the reference's ORB extractor (GeoFlow-SLAM src/ORBextractor.cc) and of
// the OpenCV primitives it calls (cv::resize INTER_AREA, cv::FAST, cv::GaussianBlur 7x7,
// cv::fastAtan2), written without OpenCV/Eigen.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library.
//
// Parity status: the reference ships NO golden vectors for this path and
// cannot be compiled here (needs OpenCV C++ headers).  The OpenCV-owned stages are pinned
// against the cv2 4.13.0 wheel in tests/test_oracle_orb.py (bit-exact); the reference-owned
// logic (cell grid, quadtree, orientation, rBRIEF, packing) follows the cited lines.
//
// One deliberate pin: the reference evaluates cosf/sinf through the platform libm, whose
// result can differ by 1 ulp between glibc builds/ifunc variants.  The oracle uses the
// correctly-rounded value (double libm, rounded once to float) so that results are
// machine-independent; see DESIGN.md "rBRIEF trig".
#include <algorithm>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <list>
#include <utility>
#include <vector>

#include "../include/gfs_orb_pattern.h"

namespace gfo {

static const int PATCH_SIZE = 31;       // ORBextractor.cc (file-scope constants)
static const int HALF_PATCH_SIZE = 15;  // include/ORBextractor.h:30
static const int EDGE_THRESHOLD = 19;   // include/ORBextractor.h:31

struct KeyPoint {
  float x, y, size, angle, response;
  int octave;
};

// cvRound on x86-64: cvtsd2si / lrint -> round-half-to-even.
static inline int cv_round(double v) { return (int)std::nearbyint(v); }
static inline int cv_floor(double v) { return (int)std::floor(v); }
static inline int cv_ceil(double v) { return (int)std::ceil(v); }

struct Image {
  int w = 0, h = 0;
  std::vector<uint8_t> d;
  Image() {}
  Image(int w_, int h_) : w(w_), h(h_), d((size_t)w_ * h_) {}
  inline uint8_t at(int y, int x) const { return d[(size_t)y * w + x]; }
  inline const uint8_t* row(int y) const { return &d[(size_t)y * w]; }
  inline uint8_t* row(int y) { return &d[(size_t)y * w]; }
};

// ---------------------------------------------------------------------------------------------
// cv::resize(src, dst, dsize, 0, 0, INTER_AREA) for a non-integer shrink factor
// (call site ORBextractor.cc:1240).  Restatement of OpenCV's area table + ResizeArea loops,
// SURVEY.md Appendix A.  Weights float32; accumulation order preserved; no FMA.
// ---------------------------------------------------------------------------------------------
struct AreaTab {
  int si, di;
  float alpha;
};

static void area_tab(int ssize, int dsize, std::vector<AreaTab>& tab) {
  const double scale = 1.0 / ((double)dsize / (double)ssize);  // cv::resize: scale = 1/inv_scale
  tab.clear();
  for (int dx = 0; dx < dsize; dx++) {
    double fsx1 = dx * scale;
    double fsx2 = fsx1 + scale;
    double cellWidth = std::min(scale, ssize - fsx1);
    int sx1 = cv_ceil(fsx1), sx2 = cv_floor(fsx2);
    sx2 = std::min(sx2, ssize - 1);
    sx1 = std::min(sx1, sx2);
    if (sx1 - fsx1 > 1e-3) tab.push_back({sx1 - 1, dx, (float)((sx1 - fsx1) / cellWidth)});
    for (int sx = sx1; sx < sx2; sx++) tab.push_back({sx, dx, (float)(1.0 / cellWidth)});
    if (fsx2 - sx2 > 1e-3)
      tab.push_back({sx2, dx, (float)(std::min(std::min(fsx2 - sx2, 1.), cellWidth) / cellWidth)});
  }
}

static void resize_area(const Image& src, Image& dst) {
  std::vector<AreaTab> xtab, ytab;
  area_tab(src.w, dst.w, xtab);
  area_tab(src.h, dst.h, ytab);
  std::vector<float> buf(dst.w), sum(dst.w);
  // rows of ytab are grouped by destination row, ascending
  size_t k = 0;
  while (k < ytab.size()) {
    int dy = ytab[k].di;
    bool first = true;
    for (; k < ytab.size() && ytab[k].di == dy; k++) {
      const uint8_t* S = src.row(ytab[k].si);
      const float beta = ytab[k].alpha;
      std::fill(buf.begin(), buf.end(), 0.f);
      for (const AreaTab& t : xtab) {
        const float prod = (float)S[t.si] * t.alpha;  // built with -ffp-contract=off
        buf[t.di] = buf[t.di] + prod;
      }
      if (first) {
        for (int dx = 0; dx < dst.w; dx++) sum[dx] = beta * buf[dx];
        first = false;
      } else {
        for (int dx = 0; dx < dst.w; dx++) {
          const float prod = beta * buf[dx];
          sum[dx] = sum[dx] + prod;
        }
      }
    }
    uint8_t* D = dst.row(dy);
    for (int dx = 0; dx < dst.w; dx++) {
      int v = (int)std::nearbyintf(sum[dx]);  // saturate_cast<uchar>(float) = cvRound + clamp
      D[dx] = (uint8_t)std::min(255, std::max(0, v));
    }
  }
}

// ---------------------------------------------------------------------------------------------
// cv::GaussianBlur(u8, Size(7,7), 2, 2, BORDER_REFLECT_101)  (ORBextractor.cc:1189).
// OpenCV's 8-bit path is fixed point: Q8 kernel [18,34,48,56,48,34,18], horizontal exact,
// vertical exact in Q16, (v + 32768) >> 16.  SURVEY.md Appendix A.
// ---------------------------------------------------------------------------------------------
static inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) {
    if (p < 0) p = -p;
    else p = 2 * (n - 1) - p;
  }
  return p;
}

static void blur7(const Image& src, Image& dst) {
  static const int K[7] = {18, 34, 48, 56, 48, 34, 18};
  const int w = src.w, h = src.h;
  std::vector<uint16_t> tmp((size_t)w * h);
  for (int y = 0; y < h; y++) {
    const uint8_t* S = src.row(y);
    for (int x = 0; x < w; x++) {
      int acc = 0;
      for (int k = 0; k < 7; k++) acc += K[k] * S[reflect101(x + k - 3, w)];
      tmp[(size_t)y * w + x] = (uint16_t)acc;  // <= 256*255 fits
    }
  }
  for (int y = 0; y < h; y++) {
    uint8_t* D = dst.row(y);
    for (int x = 0; x < w; x++) {
      uint32_t acc = 0;
      for (int k = 0; k < 7; k++) acc += (uint32_t)K[k] * tmp[(size_t)reflect101(y + k - 3, h) * w + x];
      D[x] = (uint8_t)((acc + 32768u) >> 16);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// cv::fastAtan2(y, x) in degrees (ORBextractor.cc:94).  float32, no FMA.  Appendix A.
// ---------------------------------------------------------------------------------------------
static float fast_atan2(float y, float x) {
  static const float scale = (float)(180.0 / M_PI);
  static const float p1 = 0.9997878412794807f * scale;
  static const float p3 = -0.3258083974640975f * scale;
  static const float p5 = 0.1555786518463281f * scale;
  static const float p7 = -0.04432655554792128f * scale;
  const float ax = std::fabs(x), ay = std::fabs(y);
  float a, c, c2, t;
  if (ax >= ay) {
    c = ay / (ax + (float)DBL_EPSILON);
    c2 = c * c;
    t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1;
    a = t * c;
  } else {
    c = ax / (ay + (float)DBL_EPSILON);
    c2 = c * c;
    t = p7 * c2; t = t + p5; t = t * c2; t = t + p3; t = t * c2; t = t + p1;
    t = t * c;
    a = 90.f - t;
  }
  if (x < 0) a = 180.f - a;
  if (y < 0) a = 360.f - a;
  return a;
}

// ---------------------------------------------------------------------------------------------
// cv::FAST(roi, kps, threshold, true) == FastFeatureDetector TYPE_9_16 with NMS
// (ORBextractor.cc:809,826).  The ROI is [x0,x1) x [y0,y1) of img; only ROI pixels are read;
// corners are tested >= 3 px inside the ROI; NMS sees score 0 outside the tested region.
// Output row-major, coordinates relative to the ROI origin.  Appendix A.
// ---------------------------------------------------------------------------------------------
static const int FAST_DX[16] = {0, 1, 2, 3, 3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1};
static const int FAST_DY[16] = {3, 3, 2, 1, 0, -1, -2, -3, -3, -3, -2, -1, 0, 1, 2, 3};

// best = max over the 16 arcs of 9 of max(min d, min -d); -1 when the antipodal-pair
// pre-test (any 9-arc contains one pixel of every pair (k, k+8)) already rules a corner out.
static inline int fast_best(const uint8_t* p, int pitch, int thr) {
  const int c = p[0];
  const int lo = c - thr, hi = c + thr;
  auto flag = [&](int k) -> int {
    int v = p[FAST_DY[k] * pitch + FAST_DX[k]];
    return (v < lo ? 1 : 0) | (v > hi ? 2 : 0);
  };
  int f = flag(0) | flag(8);
  if (!f) return -1;
  f &= flag(2) | flag(10);
  f &= flag(4) | flag(12);
  f &= flag(6) | flag(14);
  if (!f) return -1;
  f &= flag(1) | flag(9);
  f &= flag(3) | flag(11);
  f &= flag(5) | flag(13);
  f &= flag(7) | flag(15);
  if (!f) return -1;
  int d[25];
  for (int k = 0; k < 16; k++) d[k] = c - p[FAST_DY[k] * pitch + FAST_DX[k]];
  for (int k = 16; k < 25; k++) d[k] = d[k - 16];
  int best = -1000;
  for (int s = 0; s < 16; s++) {
    int mn = d[s], mx = d[s];
    for (int k = 1; k < 9; k++) {
      mn = std::min(mn, d[s + k]);
      mx = std::max(mx, d[s + k]);
    }
    best = std::max(best, std::max(mn, -mx));
  }
  return best;
}

struct FastPt {
  int x, y, score;
};

static void fast_roi(const Image& img, int x0, int y0, int x1, int y1, int thr, std::vector<FastPt>& out) {
  out.clear();
  thr = std::min(std::max(thr, 0), 255);
  const int w = x1 - x0, h = y1 - y0;
  if (w < 7 || h < 7) return;
  std::vector<int> sc((size_t)w * h, 0);
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      int best = fast_best(img.row(y0 + y) + x0 + x, img.w, thr);
      if (best > thr) sc[(size_t)y * w + x] = best - 1;
    }
  // 3x3 NMS, strict '>' (thr >= 0, so a kept corner always has score > 0)
  for (int y = 3; y < h - 3; y++)
    for (int x = 3; x < w - 3; x++) {
      const int s = sc[(size_t)y * w + x];
      if (s <= 0) continue;
      const int* r0 = &sc[(size_t)(y - 1) * w + x];
      const int* r1 = r0 + w;
      const int* r2 = r1 + w;
      if (s > r0[-1] && s > r0[0] && s > r0[1] && s > r1[-1] && s > r1[1] && s > r2[-1] && s > r2[0] && s > r2[1])
        out.push_back({x, y, s});
    }
}

// ---------------------------------------------------------------------------------------------
// ORBextractor (ORBextractor.cc:421-479 ctor, 1145-1225 operator(), 1227-1251 ComputePyramid)
// ---------------------------------------------------------------------------------------------
struct Node {  // ExtractorNode, include/ORBextractor.h:32-44
  std::vector<KeyPoint> keys;
  int ULx, ULy, URx, URy, BLx, BLy, BRx, BRy;
  std::list<Node>::iterator lit;
  bool noMore = false;
};

// ExtractorNode::DivideNode, ORBextractor.cc:502-550
static void divide(const Node& n, Node& n1, Node& n2, Node& n3, Node& n4) {
  const int halfX = (int)std::ceil(static_cast<float>(n.URx - n.ULx) / 2);
  const int halfY = (int)std::ceil(static_cast<float>(n.BRy - n.ULy) / 2);
  n1.ULx = n.ULx; n1.ULy = n.ULy;
  n1.URx = n.ULx + halfX; n1.URy = n.ULy;
  n1.BLx = n.ULx; n1.BLy = n.ULy + halfY;
  n1.BRx = n.ULx + halfX; n1.BRy = n.ULy + halfY;
  n2.ULx = n1.URx; n2.ULy = n1.URy;
  n2.URx = n.URx; n2.URy = n.URy;
  n2.BLx = n1.BRx; n2.BLy = n1.BRy;
  n2.BRx = n.URx; n2.BRy = n.ULy + halfY;
  n3.ULx = n1.BLx; n3.ULy = n1.BLy;
  n3.URx = n1.BRx; n3.URy = n1.BRy;
  n3.BLx = n.BLx; n3.BLy = n.BLy;
  n3.BRx = n1.BRx; n3.BRy = n.BLy;
  n4.ULx = n3.URx; n4.ULy = n3.URy;
  n4.URx = n2.BRx; n4.URy = n2.BRy;
  n4.BLx = n3.BRx; n4.BLy = n3.BRy;
  n4.BRx = n.BRx; n4.BRy = n.BRy;
  for (const KeyPoint& kp : n.keys) {
    if (kp.x < n1.URx) {
      if (kp.y < n1.BRy) n1.keys.push_back(kp);
      else n3.keys.push_back(kp);
    } else if (kp.y < n1.BRy)
      n2.keys.push_back(kp);
    else
      n4.keys.push_back(kp);
  }
  if (n1.keys.size() == 1) n1.noMore = true;
  if (n2.keys.size() == 1) n2.noMore = true;
  if (n3.keys.size() == 1) n3.noMore = true;
  if (n4.keys.size() == 1) n4.noMore = true;
}

typedef std::pair<int, Node*> SzNode;
// compareNodes, ORBextractor.cc:552-565.  std::sort (libstdc++ introsort) is not stable; the
// reference is built with GCC, so the oracle calls the same std::sort.
static bool compare_nodes(SzNode& a, SzNode& b) {
  if (a.first < b.first) return true;
  if (a.first > b.first) return false;
  return a.second->ULx < b.second->ULx;
}

// ORBextractor::DistributeOctTree, ORBextractor.cc:567-768
static std::vector<KeyPoint> distribute_octtree(const std::vector<KeyPoint>& in, int minX, int maxX, int minY,
                                                int maxY, int N) {
  int nIni = (int)std::round(static_cast<float>(maxX - minX) / (maxY - minY));
  if (nIni == 0) nIni = 1;
  const float hX = static_cast<float>(maxX - minX) / nIni;
  std::list<Node> L;
  std::vector<Node*> ini(nIni);
  for (int i = 0; i < nIni; i++) {
    Node ni;
    ni.ULx = (int)(hX * static_cast<float>(i)); ni.ULy = 0;
    ni.URx = (int)(hX * static_cast<float>(i + 1)); ni.URy = 0;
    ni.BLx = ni.ULx; ni.BLy = maxY - minY;
    ni.BRx = ni.URx; ni.BRy = maxY - minY;
    L.push_back(ni);
    ini[i] = &L.back();
  }
  for (const KeyPoint& kp : in) ini[(int)(kp.x / hX)]->keys.push_back(kp);
  auto lit = L.begin();
  while (lit != L.end()) {
    if (lit->keys.size() == 1) { lit->noMore = true; lit++; }
    else if (lit->keys.empty()) lit = L.erase(lit);
    else lit++;
  }
  bool finish = false;
  std::vector<SzNode> vSz;
  auto push_child = [&](Node& c, std::vector<SzNode>& v, int* nToExpand) {
    if (c.keys.size() > 0) {
      L.push_front(c);
      if (c.keys.size() > 1) {
        if (nToExpand) (*nToExpand)++;
        v.push_back(std::make_pair((int)c.keys.size(), &L.front()));
        L.front().lit = L.begin();
      }
    }
  };
  while (!finish) {
    int prevSize = (int)L.size();
    lit = L.begin();
    int nToExpand = 0;
    vSz.clear();
    while (lit != L.end()) {
      if (lit->noMore) { lit++; continue; }
      Node n1, n2, n3, n4;
      divide(*lit, n1, n2, n3, n4);
      push_child(n1, vSz, &nToExpand);
      push_child(n2, vSz, &nToExpand);
      push_child(n3, vSz, &nToExpand);
      push_child(n4, vSz, &nToExpand);
      lit = L.erase(lit);
    }
    if ((int)L.size() >= N || (int)L.size() == prevSize) {
      finish = true;
    } else if (((int)L.size() + nToExpand * 3) > N) {
      while (!finish) {
        prevSize = (int)L.size();
        std::vector<SzNode> prev = vSz;
        vSz.clear();
        std::sort(prev.begin(), prev.end(), compare_nodes);
        for (int j = (int)prev.size() - 1; j >= 0; j--) {
          Node n1, n2, n3, n4;
          divide(*prev[j].second, n1, n2, n3, n4);
          push_child(n1, vSz, nullptr);
          push_child(n2, vSz, nullptr);
          push_child(n3, vSz, nullptr);
          push_child(n4, vSz, nullptr);
          L.erase(prev[j].second->lit);
          if ((int)L.size() >= N) break;
        }
        if ((int)L.size() >= N || (int)L.size() == prevSize) finish = true;
      }
    }
  }
  std::vector<KeyPoint> res;
  for (auto it = L.begin(); it != L.end(); it++) {
    const std::vector<KeyPoint>& v = it->keys;
    const KeyPoint* p = &v[0];
    float maxR = p->response;
    for (size_t k = 1; k < v.size(); k++)
      if (v[k].response > maxR) { p = &v[k]; maxR = v[k].response; }
    res.push_back(*p);
  }
  return res;
}

struct Extractor {
  int nfeatures, nlevels, iniTh, minTh;
  double scaleFactor;
  std::vector<float> sf, isf;
  std::vector<int> nPerLevel, umax;
  std::vector<Image> pyr, blurred;
  std::vector<std::vector<KeyPoint>> lastCandidates;  // test hook: FAST candidates per level
  int threads = 1;

  Extractor(int nf, float s, int nl, int ini, int mn) : nfeatures(nf), nlevels(nl), iniTh(ini), minTh(mn), scaleFactor(s) {
    sf.resize(nl); isf.resize(nl);
    sf[0] = 1.0f;
    for (int i = 1; i < nl; i++) sf[i] = (float)(sf[i - 1] * scaleFactor);
    for (int i = 0; i < nl; i++) isf[i] = 1.0f / sf[i];
    nPerLevel.resize(nl);
    float factor = (float)(1.0f / scaleFactor);
    float nDesired = nfeatures * (1 - factor) / (1 - (float)std::pow((double)factor, (double)nlevels));
    int sum = 0;
    for (int l = 0; l < nl - 1; l++) {
      nPerLevel[l] = cv_round(nDesired);
      sum += nPerLevel[l];
      nDesired *= factor;
    }
    nPerLevel[nl - 1] = std::max(nfeatures - sum, 0);
    umax.resize(HALF_PATCH_SIZE + 1);
    int v, v0, vmax = cv_floor(HALF_PATCH_SIZE * std::sqrt(2.f) / 2 + 1);
    int vmin = cv_ceil(HALF_PATCH_SIZE * std::sqrt(2.f) / 2);
    const double hp2 = HALF_PATCH_SIZE * HALF_PATCH_SIZE;
    for (v = 0; v <= vmax; ++v) umax[v] = cv_round(std::sqrt(hp2 - v * v));
    for (v = HALF_PATCH_SIZE, v0 = 0; v >= vmin; --v) {
      while (umax[v0] == umax[v0 + 1]) ++v0;
      umax[v] = v0;
      ++v0;
    }
  }

  void level_size(int w, int h, int l, int* lw, int* lh) const {
    float s = isf[l];
    *lw = cv_round((float)w * s);
    *lh = cv_round((float)h * s);
  }

  // ComputePyramid, ORBextractor.cc:1227-1251.  The 19-px REFLECT_101 border the reference
  // adds around every level is never read by the extractor itself (FAST cells start 16 px
  // inside, the patch radius is 15, the blur reflects on its own clone), so the oracle keeps
  // the un-bordered views.
  void compute_pyramid(const uint8_t* img, int w, int h, int pitch) {
    pyr.resize(nlevels);
    for (int l = 0; l < nlevels; l++) {
      int lw, lh;
      level_size(w, h, l, &lw, &lh);
      pyr[l] = Image(lw, lh);
      if (l == 0) {
        for (int y = 0; y < h; y++) memcpy(pyr[0].row(y), img + (size_t)y * pitch, w);
      } else {
        resize_area(pyr[l - 1], pyr[l]);
      }
    }
  }

  // IC_Angle, ORBextractor.cc:71-95
  float ic_angle(const Image& im, float px, float py) const {
    int m_01 = 0, m_10 = 0;
    const int cx = cv_round(px), cy = cv_round(py);
    for (int u = -HALF_PATCH_SIZE; u <= HALF_PATCH_SIZE; ++u) m_10 += u * im.at(cy, cx + u);
    for (int v = 1; v <= HALF_PATCH_SIZE; ++v) {
      int v_sum = 0;
      int d = umax[v];
      for (int u = -d; u <= d; ++u) {
        int val_plus = im.at(cy + v, cx + u), val_minus = im.at(cy - v, cx + u);
        v_sum += (val_plus - val_minus);
        m_10 += u * (val_plus + val_minus);
      }
      m_01 += v * v_sum;
    }
    return fast_atan2((float)m_01, (float)m_10);
  }

  // ComputeKeyPointsOctTree for one level, ORBextractor.cc:770-871 (cell FAST + quadtree)
  void keypoints_level(int level, std::vector<KeyPoint>& kps, std::vector<KeyPoint>* cand_out) const {
    const Image& im = pyr[level];
    const float W = 35;
    const int minBorderX = EDGE_THRESHOLD - 3;
    const int minBorderY = minBorderX;
    const int maxBorderX = im.w - EDGE_THRESHOLD + 3;
    const int maxBorderY = im.h - EDGE_THRESHOLD + 3;
    std::vector<KeyPoint> cand;
    const float width = (float)(maxBorderX - minBorderX);
    const float height = (float)(maxBorderY - minBorderY);
    const int nCols = (int)(width / W);
    const int nRows = (int)(height / W);
    const int wCell = (int)std::ceil(width / nCols);
    const int hCell = (int)std::ceil(height / nRows);
    std::vector<FastPt> cell;
    for (int i = 0; i < nRows; i++) {
      const float iniY = (float)(minBorderY + i * hCell);
      float maxY = iniY + hCell + 6;
      if (iniY >= maxBorderY - 3) continue;
      if (maxY > maxBorderY) maxY = (float)maxBorderY;
      for (int j = 0; j < nCols; j++) {
        const float iniX = (float)(minBorderX + j * wCell);
        float maxX = iniX + wCell + 6;
        if (iniX >= maxBorderX - 6) continue;
        if (maxX > maxBorderX) maxX = (float)maxBorderX;
        fast_roi(im, (int)iniX, (int)iniY, (int)maxX, (int)maxY, iniTh, cell);
        if (cell.empty()) fast_roi(im, (int)iniX, (int)iniY, (int)maxX, (int)maxY, minTh, cell);
        for (const FastPt& p : cell) {
          KeyPoint kp;
          kp.x = (float)p.x + j * wCell;
          kp.y = (float)p.y + i * hCell;
          kp.size = 7.f;  // cv::FAST sets size 7
          kp.angle = -1.f;
          kp.response = (float)p.score;
          kp.octave = 0;
          cand.push_back(kp);
        }
      }
    }
    if (cand_out) *cand_out = cand;
    kps = distribute_octtree(cand, minBorderX, maxBorderX, minBorderY, maxBorderY, nPerLevel[level]);
    const int scaledPatchSize = (int)(PATCH_SIZE * sf[level]);
    for (KeyPoint& k : kps) {
      k.x += minBorderX;
      k.y += minBorderY;
      k.octave = level;
      k.size = (float)scaledPatchSize;
    }
    for (KeyPoint& k : kps) k.angle = ic_angle(im, k.x, k.y);  // computeOrientation :481-500
  }

  // computeOrbDescriptor, ORBextractor.cc:99-160
  void descriptor(const KeyPoint& kpt, const Image& img, uint8_t* desc) const {
    const float factorPI = (float)(M_PI / 180.f);
    const float angle = (float)kpt.angle * factorPI;
    // reference: cosf/sinf (float overloads under `using namespace std`); oracle pins the
    // correctly rounded value -- see file header.
    const float a = (float)std::cos((double)angle), b = (float)std::sin((double)angle);
    const int cy = (int)std::round(kpt.y), cx = (int)std::round(kpt.x);
    const signed char* pat = GFS_ORB_PATTERN;
    auto get = [&](int idx) -> int {
      const float px = (float)pat[2 * idx], py = (float)pat[2 * idx + 1];
      const float r1 = px * b + py * a;  // -ffp-contract=off: two roundings + add, as SSE2 -O3
      const float r2 = px * a - py * b;
      const int ry = (int)std::round(r1), rx = (int)std::round(r2);
      return img.at(cy + ry, cx + rx);
    };
    for (int i = 0; i < 32; ++i, pat += 32) {
      int val = 0;
      for (int k = 0; k < 8; k++) {
        int t0 = get(2 * k), t1 = get(2 * k + 1);
        val |= (t0 < t1) << k;
      }
      desc[i] = (uint8_t)val;
    }
  }

  // operator(), ORBextractor.cc:1145-1225.  Returns monoIndex (or -1 on empty image).
  int extract(const uint8_t* img, int w, int h, int pitch, int lap0, int lap1, std::vector<KeyPoint>& out,
              std::vector<uint8_t>& desc) {
    if (!img || w <= 0 || h <= 0) return -1;
    compute_pyramid(img, w, h, pitch);
    std::vector<std::vector<KeyPoint>> all(nlevels);
    lastCandidates.assign(nlevels, {});
    blurred.assign(nlevels, Image());
#pragma omp parallel for schedule(dynamic) num_threads(threads)
    for (int l = 0; l < nlevels; l++) keypoints_level(l, all[l], &lastCandidates[l]);
    int n = 0;
    for (int l = 0; l < nlevels; l++) n += (int)all[l].size();
    out.assign(n, KeyPoint());
    desc.assign((size_t)n * 32, 0);
    int mono = 0, stereo = n - 1;
    for (int l = 0; l < nlevels; l++) {
      std::vector<KeyPoint>& kps = all[l];
      if (kps.empty()) continue;
      blurred[l] = Image(pyr[l].w, pyr[l].h);
      blur7(pyr[l], blurred[l]);
      std::vector<uint8_t> d((size_t)kps.size() * 32);
#pragma omp parallel for num_threads(threads)
      for (int i = 0; i < (int)kps.size(); i++) descriptor(kps[i], blurred[l], &d[(size_t)i * 32]);
      const float scale = sf[l];
      for (size_t i = 0; i < kps.size(); i++) {
        KeyPoint& k = kps[i];
        if (l != 0) { k.x *= scale; k.y *= scale; }
        int dst;
        if (k.x >= lap0 && k.x <= lap1) dst = stereo--;
        else dst = mono++;
        out[dst] = k;
        memcpy(&desc[(size_t)dst * 32], &d[i * 32], 32);
      }
    }
    return mono;
  }
};

}  // namespace gfo

// ---------------------------------------------------------------------------------------------
// C entry points for ctypes (tests / bench cpu_baseline only)
// ---------------------------------------------------------------------------------------------
extern "C" {

void* gfo_orb_create(int nfeatures, float scale, int nlevels, int iniTh, int minTh) {
  return new gfo::Extractor(nfeatures, scale, nlevels, iniTh, minTh);
}
void gfo_orb_destroy(void* h) { delete (gfo::Extractor*)h; }
void gfo_orb_set_threads(void* h, int t) { ((gfo::Extractor*)h)->threads = t; }

void gfo_orb_tables(void* h, float* sf, int* nPerLevel, int* umax16) {
  gfo::Extractor* e = (gfo::Extractor*)h;
  for (int i = 0; i < e->nlevels; i++) { sf[i] = e->sf[i]; nPerLevel[i] = e->nPerLevel[i]; }
  for (int i = 0; i < 16; i++) umax16[i] = e->umax[i];
}
void gfo_orb_level_size(void* h, int w, int ht, int l, int* lw, int* lh) {
  ((gfo::Extractor*)h)->level_size(w, ht, l, lw, lh);
}

// returns n keypoints; *mono = monoIndex.  kps capacity cap (6 x 4 bytes each), desc cap*32.
int gfo_orb_extract(void* h, const uint8_t* img, int w, int ht, int pitch, int lap0, int lap1, void* kps,
                    uint8_t* desc, int cap, int* mono) {
  gfo::Extractor* e = (gfo::Extractor*)h;
  std::vector<gfo::KeyPoint> out;
  std::vector<uint8_t> d;
  int m = e->extract(img, w, ht, pitch, lap0, lap1, out, d);
  if (mono) *mono = m;
  if (m < 0) return -1;
  int n = (int)std::min<size_t>(out.size(), (size_t)cap);
  memcpy(kps, out.data(), (size_t)n * sizeof(gfo::KeyPoint));
  memcpy(desc, d.data(), (size_t)n * 32);
  return (int)out.size();
}

// stage hooks (valid after gfo_orb_extract)
int gfo_orb_get_level(void* h, int l, int blurred, uint8_t* dst) {
  gfo::Extractor* e = (gfo::Extractor*)h;
  const gfo::Image& im = blurred ? e->blurred[l] : e->pyr[l];
  if (im.d.empty()) return 0;
  memcpy(dst, im.d.data(), im.d.size());
  return 1;
}
int gfo_orb_get_candidates(void* h, int l, float* xyr, int cap) {
  gfo::Extractor* e = (gfo::Extractor*)h;
  const auto& c = e->lastCandidates[l];
  int n = (int)std::min<size_t>(c.size(), (size_t)cap);
  for (int i = 0; i < n; i++) { xyr[3 * i] = c[i].x; xyr[3 * i + 1] = c[i].y; xyr[3 * i + 2] = c[i].response; }
  return (int)c.size();
}

void gfo_resize_area(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh) {
  gfo::Image s(sw, sh), d(dw, dh);
  memcpy(s.d.data(), src, s.d.size());
  gfo::resize_area(s, d);
  memcpy(dst, d.d.data(), d.d.size());
}
void gfo_blur7(const uint8_t* src, int w, int h, uint8_t* dst) {
  gfo::Image s(w, h), d(w, h);
  memcpy(s.d.data(), src, s.d.size());
  gfo::blur7(s, d);
  memcpy(dst, d.d.data(), d.d.size());
}
float gfo_fast_atan2(float y, float x) { return gfo::fast_atan2(y, x); }
int gfo_fast(const uint8_t* img, int w, int h, int thr, int* xys, int cap) {
  gfo::Image s(w, h);
  memcpy(s.d.data(), img, s.d.size());
  std::vector<gfo::FastPt> out;
  gfo::fast_roi(s, 0, 0, w, h, thr, out);
  int n = (int)std::min<size_t>(out.size(), (size_t)cap);
  for (int i = 0; i < n; i++) { xys[3 * i] = out[i].x; xys[3 * i + 1] = out[i].y; xys[3 * i + 2] = out[i].score; }
  return (int)out.size();
}
// quadtree alone (for kernel-level parity tests): in = n x (x,y,response) floats relative to minX/minY
int gfo_distribute_octtree(const float* xyr, int n, int minX, int maxX, int minY, int maxY, int N, float* out_xyr,
                           int cap) {
  std::vector<gfo::KeyPoint> in(n);
  for (int i = 0; i < n; i++) { in[i].x = xyr[3 * i]; in[i].y = xyr[3 * i + 1]; in[i].response = xyr[3 * i + 2]; }
  std::vector<gfo::KeyPoint> r = gfo::distribute_octtree(in, minX, maxX, minY, maxY, N);
  int m = (int)std::min<size_t>(r.size(), (size_t)cap);
  for (int i = 0; i < m; i++) { out_xyr[3 * i] = r[i].x; out_xyr[3 * i + 1] = r[i].y; out_xyr[3 * i + 2] = r[i].response; }
  return (int)r.size();
}
}

// std::sort of (first, second) pairs with the compareNodes ordering (ORBextractor.cc:552-565),
// key for ties = ulx[second].  Reference permutation for tests of csrc/gcc_sort.h.
extern "C" void gfo_std_sort_pairs(int* first, int* second, const int* ulx, int n) {
  std::vector<std::pair<int, int>> v(n);
  for (int i = 0; i < n; i++) v[i] = std::make_pair(first[i], second[i]);
  std::sort(v.begin(), v.end(), [ulx](std::pair<int, int>& a, std::pair<int, int>& b) {
    if (a.first < b.first) return true;
    if (a.first > b.first) return false;
    return ulx[a.second] < ulx[b.second];
  });
  for (int i = 0; i < n; i++) { first[i] = v[i].first; second[i] = v[i].second; }
}

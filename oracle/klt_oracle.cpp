// ORACLE -- TEST INFRASTRUCTURE ONLY. Not part of the product path.
//
// CPU restatement of the optical-flow front end:
//   Frame::Frame            cv::buildOpticalFlowPyramid(image, mImGray, winSize, 3)      reference src/Frame.cc:370-373
//   ORBmatcher::fbKltTracking / Tracking::fbKltTracking                                   src/ORBmatcher.cc:2186-2293,
//                            forward cv::calcOpticalFlowPyrLK over nbpyrlvl levels with                src/Tracking.cc:3262-3360
//                            OPTFLOW_USE_INITIAL_FLOW | OPTFLOW_LK_GET_MIN_EIGENVALS, 30 iterations,
//                            eps 0.01, status / min-eigenvalue / inBorder filtering, backward pass at
//                            level 0, forward-backward distance check
// The arithmetic lives in OpenCV (unpinned system package, CMakeLists.txt:66; restated from the published
// algorithm of modules/video/src/lkpyramid.cpp and modules/imgproc/src/pyramids.cpp, 4.x):
//   pyrDown (8U)            separable [1 4 6 4 1], BORDER_REFLECT_101, (sum + 128) >> 8, size (w+1)/2 x (h+1)/2
//   calcScharrDeriv         int16 (dI/dx, dI/dy) with the 3-10-3 Scharr kernel, reflect-101 inside the image;
//                           outside the image the derivative is 0 (derivBorder = BORDER_CONSTANT), the image is
//                           reflect-101 (pyrBorder), exactly what the padded pyramid holds
//   LKTrackerInvoker        fixed-point bilinear weights (W_BITS = 14), int16 patches (5 fractional bits),
//                           G = sum of derivative products, 2x2 solve, eps / oscillation termination
// OpenCV accumulates the integer products in float SIMD lanes (order-dependent rounding above 2^24); here they
// are summed exactly in int64 and narrowed once, so the CUDA path can agree with this oracle bit for bit.
// Pinning: tests/test_oracle_klt.py checks pyrDown and the Scharr derivative bit-exact against the cv2 4.13 wheel
// and the tracker against cv2.calcOpticalFlowPyrLK within a stated pixel tolerance.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

namespace gfo {
namespace klt {

static inline int reflect101(int p, int len) {  // cv::borderInterpolate(BORDER_REFLECT_101)
  if (len == 1) return 0;
  while (p < 0 || p >= len) {
    if (p < 0) p = -p;
    else p = 2 * (len - 1) - p;
  }
  return p;
}
static inline int cv_round(float v) { return (int)lrintf(v); }   // cvRound: round half to even (SSE cvtss2si)
static inline int cv_floor(float v) { return (int)std::floor(v); }

// cv::CLAHE::apply for CV_8UC1 (modules/imgproc/src/clahe.cpp: CLAHE_CalcLut_Body, CLAHE_Interpolation_Body), as
// Frame::Frame applies it in place with createCLAHE(3.0, Size(8, 8)) (reference src/Frame.cc:366-368, 498-500)
static void clahe(const uint8_t* src, int w, int h, double clipLimit_, int tilesX, int tilesY, uint8_t* dst) {
  const int histSize = 256;
  int ew = w, eh = h;  // extended size: copyMakeBorder(BORDER_REFLECT_101) on the right / bottom when not divisible
  if (!(w % tilesX == 0 && h % tilesY == 0)) { ew = w + (tilesX - (w % tilesX)); eh = h + (tilesY - (h % tilesY)); }
  const int tw = ew / tilesX, th = eh / tilesY;
  const int tileSizeTotal = tw * th;
  const float lutScale = static_cast<float>(histSize - 1) / tileSizeTotal;
  int clipLimit = 0;
  if (clipLimit_ > 0.0) {
    clipLimit = static_cast<int>(clipLimit_ * tileSizeTotal / histSize);
    clipLimit = std::max(clipLimit, 1);
  }
  std::vector<uint8_t> lut((size_t)tilesX * tilesY * histSize);
  for (int k = 0; k < tilesX * tilesY; k++) {
    const int ty = k / tilesX, tx = k % tilesX;
    int hist[256] = {0};
    for (int y = 0; y < th; y++)
      for (int x = 0; x < tw; x++) hist[src[(size_t)reflect101(ty * th + y, h) * w + reflect101(tx * tw + x, w)]]++;
    if (clipLimit > 0) {
      int clipped = 0;
      for (int i = 0; i < histSize; i++)
        if (hist[i] > clipLimit) { clipped += hist[i] - clipLimit; hist[i] = clipLimit; }
      const int redistBatch = clipped / histSize;
      int residual = clipped - redistBatch * histSize;
      for (int i = 0; i < histSize; i++) hist[i] += redistBatch;
      if (residual != 0) {
        const int residualStep = std::max(histSize / residual, 1);
        for (int i = 0; i < histSize && residual > 0; i += residualStep, residual--) hist[i]++;
      }
    }
    int sum = 0;
    for (int i = 0; i < histSize; i++) {
      sum += hist[i];
      const int v = cv_round(sum * lutScale);
      lut[(size_t)k * histSize + i] = (uint8_t)std::min(std::max(v, 0), 255);
    }
  }
  const float inv_tw = 1.0f / tw, inv_th = 1.0f / th;
  for (int y = 0; y < h; y++) {
    const float tyf = y * inv_th - 0.5f;
    int ty1 = cv_floor(tyf), ty2 = ty1 + 1;
    const float ya = tyf - ty1, ya1 = 1.0f - ya;
    ty1 = std::max(ty1, 0); ty2 = std::min(ty2, tilesY - 1);
    const uint8_t* p1 = lut.data() + (size_t)ty1 * tilesX * histSize;
    const uint8_t* p2 = lut.data() + (size_t)ty2 * tilesX * histSize;
    for (int x = 0; x < w; x++) {
      const float txf = x * inv_tw - 0.5f;
      int tx1 = cv_floor(txf), tx2 = tx1 + 1;
      const float xa = txf - tx1, xa1 = 1.0f - xa;
      tx1 = std::max(tx1, 0); tx2 = std::min(tx2, tilesX - 1);
      const int v = src[(size_t)y * w + x];
      const int i1 = tx1 * histSize + v, i2 = tx2 * histSize + v;
      const float res = (p1[i1] * xa1 + p1[i2] * xa) * ya1 + (p2[i1] * xa1 + p2[i2] * xa) * ya;
      dst[(size_t)y * w + x] = (uint8_t)std::min(std::max(cv_round(res), 0), 255);
    }
  }
}

static void pyr_down(const uint8_t* src, int w, int h, uint8_t* dst) {
  const int dw = (w + 1) / 2, dh = (h + 1) / 2;
  std::vector<int> row((size_t)5 * dw);
  auto hrow = [&](int sy, int* out) {
    const uint8_t* s = src + (size_t)reflect101(sy, h) * w;
    for (int x = 0; x < dw; x++) {
      const int c = 2 * x;
      out[x] = s[reflect101(c - 2, w)] + s[reflect101(c + 2, w)] + 4 * (s[reflect101(c - 1, w)] + s[reflect101(c + 1, w)]) + 6 * s[reflect101(c, w)];
    }
  };
  for (int y = 0; y < dh; y++) {
    for (int k = 0; k < 5; k++) hrow(2 * y - 2 + k, row.data() + (size_t)k * dw);
    for (int x = 0; x < dw; x++) {
      const int v = row[x] + row[4 * (size_t)dw + x] + 4 * (row[(size_t)dw + x] + row[3 * (size_t)dw + x]) + 6 * row[2 * (size_t)dw + x];
      dst[(size_t)y * dw + x] = (uint8_t)((v + 128) >> 8);
    }
  }
}
// calcScharrDeriv: dst[2*(y*w+x)] = dI/dx, +1 = dI/dy
static void scharr(const uint8_t* src, int w, int h, int16_t* dst) {
  auto px = [&](int x, int y) { return (int)src[(size_t)reflect101(y, h) * w + reflect101(x, w)]; };
  for (int y = 0; y < h; y++)
    for (int x = 0; x < w; x++) {
      // t0 = (s(y-1) + s(y+1)) * 3 + s(y) * 10 ; t1 = s(y+1) - s(y-1)
      int t0[3], t1[3];
      for (int k = -1; k <= 1; k++) {
        t0[k + 1] = (px(x + k, y - 1) + px(x + k, y + 1)) * 3 + px(x + k, y) * 10;
        t1[k + 1] = px(x + k, y + 1) - px(x + k, y - 1);
      }
      dst[2 * ((size_t)y * w + x)] = (int16_t)(t0[2] - t0[0]);
      dst[2 * ((size_t)y * w + x) + 1] = (int16_t)((t1[2] + t1[0]) * 3 + t1[1] * 10);
    }
}

struct Level { int w, h; const uint8_t* img; const int16_t* der; };
struct Pyr { std::vector<Level> lv; };

static inline int img_at(const Level& L, int x, int y) { return L.img[(size_t)reflect101(y, L.h) * L.w + reflect101(x, L.w)]; }
static inline int der_at(const Level& L, int x, int y, int c) {
  if (x < 0 || y < 0 || x >= L.w || y >= L.h) return 0;
  return L.der[2 * ((size_t)y * L.w + x) + c];
}
static inline int descale(int v, int n) { return (v + (1 << (n - 1))) >> n; }

// one level of LKTrackerInvoker for one point; returns false when the point is skipped (`continue`) at this level
static void lk_level(const Level& I, const Level& J, int level, int maxLevel, int win, int maxCount, float eps2, float minEigThr,
                     bool useInitial, float px, float py, float* nx, float* ny, uint8_t* status, float* err) {
  const float halfWin = (win - 1) * 0.5f;
  float prevx = px * (float)(1. / (1 << level)), prevy = py * (float)(1. / (1 << level));
  float nextx, nexty;
  if (level == maxLevel) {
    if (useInitial) { nextx = *nx * (float)(1. / (1 << level)); nexty = *ny * (float)(1. / (1 << level)); }
    else { nextx = prevx; nexty = prevy; }
  } else {
    nextx = *nx * 2.f; nexty = *ny * 2.f;
  }
  *nx = nextx; *ny = nexty;
  prevx -= halfWin; prevy -= halfWin;
  const int ipx = cv_floor(prevx), ipy = cv_floor(prevy);
  if (ipx < -win || ipx >= I.w || ipy < -win || ipy >= I.h) {
    if (level == 0) { *status = 0; *err = 0; }
    return;
  }
  float a = prevx - ipx, b = prevy - ipy;
  const int W_BITS = 14;
  const float FLT_SCALE = 1.f / (1 << 20);
  int iw00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
  int iw01 = cv_round(a * (1.f - b) * (1 << W_BITS));
  int iw10 = cv_round((1.f - a) * b * (1 << W_BITS));
  int iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
  std::vector<int16_t> Iw((size_t)win * win), dIw((size_t)2 * win * win);
  long long sA11 = 0, sA12 = 0, sA22 = 0;
  for (int y = 0; y < win; y++)
    for (int x = 0; x < win; x++) {
      const int X = ipx + x, Y = ipy + y;
      const int ival = descale(img_at(I, X, Y) * iw00 + img_at(I, X + 1, Y) * iw01 + img_at(I, X, Y + 1) * iw10 + img_at(I, X + 1, Y + 1) * iw11, W_BITS - 5);
      const int ixval = descale(der_at(I, X, Y, 0) * iw00 + der_at(I, X + 1, Y, 0) * iw01 + der_at(I, X, Y + 1, 0) * iw10 + der_at(I, X + 1, Y + 1, 0) * iw11, W_BITS);
      const int iyval = descale(der_at(I, X, Y, 1) * iw00 + der_at(I, X + 1, Y, 1) * iw01 + der_at(I, X, Y + 1, 1) * iw10 + der_at(I, X + 1, Y + 1, 1) * iw11, W_BITS);
      Iw[(size_t)y * win + x] = (int16_t)ival;
      dIw[2 * ((size_t)y * win + x)] = (int16_t)ixval;
      dIw[2 * ((size_t)y * win + x) + 1] = (int16_t)iyval;
      sA11 += (long long)ixval * ixval; sA12 += (long long)ixval * iyval; sA22 += (long long)iyval * iyval;
    }
  const float A11 = (float)sA11 * FLT_SCALE, A12 = (float)sA12 * FLT_SCALE, A22 = (float)sA22 * FLT_SCALE;
  float D = A11 * A22 - A12 * A12;
  const float minEig = (A22 + A11 - std::sqrt((A11 - A22) * (A11 - A22) + 4.f * A12 * A12)) / (2 * win * win);
  *err = minEig;  // OPTFLOW_LK_GET_MIN_EIGENVALS
  if (minEig < minEigThr || D < 1.1920929e-07f) {
    if (level == 0) *status = 0;
    return;
  }
  D = 1.f / D;
  nextx -= halfWin; nexty -= halfWin;
  float pdx = 0, pdy = 0;
  for (int j = 0; j < maxCount; j++) {
    const int inx = cv_floor(nextx), iny = cv_floor(nexty);
    if (inx < -win || inx >= J.w || iny < -win || iny >= J.h) {
      if (level == 0) *status = 0;
      break;
    }
    a = nextx - inx; b = nexty - iny;
    iw00 = cv_round((1.f - a) * (1.f - b) * (1 << W_BITS));
    iw01 = cv_round(a * (1.f - b) * (1 << W_BITS));
    iw10 = cv_round((1.f - a) * b * (1 << W_BITS));
    iw11 = (1 << W_BITS) - iw00 - iw01 - iw10;
    long long sb1 = 0, sb2 = 0;
    for (int y = 0; y < win; y++)
      for (int x = 0; x < win; x++) {
        const int X = inx + x, Y = iny + y;
        const int diff = descale(img_at(J, X, Y) * iw00 + img_at(J, X + 1, Y) * iw01 + img_at(J, X, Y + 1) * iw10 + img_at(J, X + 1, Y + 1) * iw11, W_BITS - 5) -
                         Iw[(size_t)y * win + x];
        sb1 += (long long)diff * dIw[2 * ((size_t)y * win + x)];
        sb2 += (long long)diff * dIw[2 * ((size_t)y * win + x) + 1];
      }
    const float b1 = (float)sb1 * FLT_SCALE, b2 = (float)sb2 * FLT_SCALE;
    const float dx = (float)((A12 * b2 - A22 * b1) * D), dy = (float)((A12 * b1 - A11 * b2) * D);
    nextx += dx; nexty += dy;
    *nx = nextx + halfWin; *ny = nexty + halfWin;
    if ((double)dx * dx + (double)dy * dy <= (double)eps2) break;  // Point2f::ddot
    if (j > 0 && std::abs(dx + pdx) < 0.01 && std::abs(dy + pdy) < 0.01) {
      *nx -= dx * 0.5f; *ny -= dy * 0.5f;
      break;
    }
    pdx = dx; pdy = dy;
  }
}

// cv::calcOpticalFlowPyrLK on prebuilt pyramids with OPTFLOW_LK_GET_MIN_EIGENVALS (+ optional initial flow)
static void calc_lk(const Pyr& P, const Pyr& C, const float* pts, float* next, int n, int win, int maxLevel, int maxCount, float eps,
                    bool useInitial, uint8_t* status, float* err) {
  maxLevel = std::min(maxLevel, (int)std::min(P.lv.size(), C.lv.size()) - 1);
  // cv::buildOpticalFlowPyramid stops once the next level would be <= the window in either dimension and returns the
  // last level it built; calcOpticalFlowPyrLK clamps maxLevel to that (lkpyramid.cpp)
  {
    int w = P.lv[0].w, h = P.lv[0].h, eff = 0;
    while (eff < maxLevel) {
      const int nw = (w + 1) / 2, nh = (h + 1) / 2;
      if (nw <= win || nh <= win) break;
      w = nw; h = nh; eff++;
    }
    maxLevel = eff;
  }
  maxCount = std::min(std::max(maxCount, 0), 100);
  eps = std::min(std::max(eps, 0.f), 10.f);
  const float eps2 = eps * eps;
  for (int i = 0; i < n; i++) { status[i] = 1; err[i] = 0; }
  for (int level = maxLevel; level >= 0; level--)
    for (int i = 0; i < n; i++)
      lk_level(P.lv[level], C.lv[level], level, maxLevel, win, maxCount, eps2, 1e-4f, useInitial, pts[2 * i], pts[2 * i + 1], &next[2 * i],
               &next[2 * i + 1], &status[i], &err[i]);
}

}  // namespace klt
}  // namespace gfo

using namespace gfo::klt;

extern "C" {

void gfo_pyr_down(const uint8_t* src, int w, int h, uint8_t* dst) { pyr_down(src, w, h, dst); }
void gfo_clahe(const uint8_t* src, int w, int h, double clip_limit, int tiles_x, int tiles_y, uint8_t* dst) { clahe(src, w, h, clip_limit, tiles_x, tiles_y, dst); }
void gfo_scharr(const uint8_t* src, int w, int h, int16_t* dst) { scharr(src, w, h, dst); }

// level sizes and offsets (in pixels) of the packed pyramid: images [sum w*h] u8, derivatives [2 * sum w*h] int16
int gfo_klt_pyramid_pixels(int w, int h, int levels) {
  int total = 0;
  for (int l = 0; l <= levels; l++) { total += w * h; w = (w + 1) / 2; h = (h + 1) / 2; }
  return total;
}
// cv::buildOpticalFlowPyramid(img, pyr, winSize, levels, withDerivatives = true)
void gfo_klt_build_pyramid(const uint8_t* img, int w, int h, int levels, uint8_t* out_img, int16_t* out_der) {
  size_t off = 0;
  int cw = w, ch = h;
  for (int l = 0; l <= levels; l++) {
    uint8_t* dst = out_img + off;
    if (l == 0) memcpy(dst, img, (size_t)cw * ch);
    scharr(dst, cw, ch, out_der + 2 * off);
    off += (size_t)cw * ch;
    if (l < levels) {
      pyr_down(dst, cw, ch, out_img + off);
      cw = (cw + 1) / 2; ch = (ch + 1) / 2;
    }
  }
}

static Pyr make_pyr(const uint8_t* img, const int16_t* der, int w, int h, int levels) {
  Pyr P;
  size_t off = 0;
  for (int l = 0; l <= levels; l++) {
    P.lv.push_back(Level{w, h, img + off, der + 2 * off});
    off += (size_t)w * h;
    w = (w + 1) / 2; h = (h + 1) / 2;
  }
  return P;
}

void gfo_klt_calc(const uint8_t* pimg, const int16_t* pder, const uint8_t* cimg, const int16_t* cder, int w, int h, int levels,
                  const float* pts, float* next, int n, int win, int maxLevel, int maxCount, float eps, int useInitial,
                  uint8_t* status, float* err) {
  const Pyr P = make_pyr(pimg, pder, w, h, levels), C = make_pyr(cimg, cder, w, h, levels);
  calc_lk(P, C, pts, next, n, win, maxLevel, maxCount, eps, useInitial != 0, status, err);
}

// ORBmatcher::fbKltTracking (src/ORBmatcher.cc:2186-2293): prior in/out, status out (1 = tracked)
void gfo_fb_klt(const uint8_t* pimg, const int16_t* pder, const uint8_t* cimg, const int16_t* cder, int w, int h, int levels,
                const float* kps, float* prior, int n, int win, int nbpyrlvl, float ferr, float fmax_fbklt_dist, uint8_t* kpstatus) {
  if (n <= 0) return;
  const Pyr P = make_pyr(pimg, pder, w, h, levels), C = make_pyr(cimg, cder, w, h, levels);
  std::vector<uint8_t> st(n);
  std::vector<float> er(n);
  calc_lk(P, C, kps, prior, n, win, nbpyrlvl, 30, 0.01f, true, st.data(), er.data());
  std::vector<float> newk, backk;
  std::vector<int> idx;
  for (int i = 0; i < n; i++) {
    kpstatus[i] = 0;
    if (!st[i]) continue;
    if (er[i] > ferr) continue;
    // inBorder (ORBmatcher.cc:2552-2557, BORDER_SIZE = 1.f) against the level-0 image
    const float x = prior[2 * i], y = prior[2 * i + 1];
    if (!(1.f <= x && x < (float)w - 1.f && 1.f <= y && y < (float)h - 1.f)) continue;
    kpstatus[i] = 1;
    newk.push_back(prior[2 * i]); newk.push_back(prior[2 * i + 1]);
    backk.push_back(kps[2 * i]); backk.push_back(kps[2 * i + 1]);
    idx.push_back(i);
  }
  if (idx.empty()) return;
  const int m = (int)idx.size();
  std::vector<uint8_t> st2(m);
  std::vector<float> er2(m);
  calc_lk(C, P, newk.data(), backk.data(), m, win, 0, 30, 0.01f, true, st2.data(), er2.data());
  for (int k = 0; k < m; k++) {
    const int i = idx[k];
    if (!st2[k]) { kpstatus[i] = 0; continue; }
    const double dx = (double)kps[2 * i] - (double)backk[2 * k], dy = (double)kps[2 * i + 1] - (double)backk[2 * k + 1];
    if (std::sqrt(dx * dx + dy * dy) > fmax_fbklt_dist) kpstatus[i] = 0;
  }
}
}

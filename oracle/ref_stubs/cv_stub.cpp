// The five OpenCV image primitives the reference's ORBextractor.cc calls, forwarded to the restatements in
// oracle/orb_oracle.cpp (bit-exact against the cv2 4.13 wheel: tests/test_oracle_orb.py, tests/golden/*.npz), plus the two
// cv::sum overloads gms_matcher.h uses.  TEST INFRASTRUCTURE (oracle/_ref).
#include <opencv2/opencv.hpp>

extern "C" {
void gfo_resize_area(const uint8_t* src, int sw, int sh, uint8_t* dst, int dw, int dh);
void gfo_blur7(const uint8_t* src, int w, int h, uint8_t* dst);
float gfo_fast_atan2(float y, float x);
int gfo_fast(const uint8_t* img, int w, int h, int thr, int* xys, int cap);
}

namespace cv {

static std::vector<uint8_t> packed(const Mat& m) {
  assert(m.type() == CV_8UC1);
  std::vector<uint8_t> v((size_t)m.rows * m.cols);
  for (int y = 0; y < m.rows; y++) memcpy(v.data() + (size_t)y * m.cols, m.ptr(y), (size_t)m.cols);
  return v;
}
static void unpack(const std::vector<uint8_t>& v, Mat& m) {
  for (int y = 0; y < m.rows; y++) memcpy(m.ptr(y), v.data() + (size_t)y * m.cols, (size_t)m.cols);
}

// cv::FAST(image, keypoints, threshold, nonmaxSuppression = true): TYPE_9_16, keypoints in row-major order,
// KeyPoint(x, y, 7.f, -1, score)
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression) {
  assert(nonmaxSuppression && "the reference always passes true");
  keypoints.clear();
  if (image.rows < 7 || image.cols < 7) return;
  const std::vector<uint8_t> img = packed(image);
  std::vector<int> xys(3 * ((size_t)image.rows * image.cols / 4 + 16));
  const int n = gfo_fast(img.data(), image.cols, image.rows, threshold, xys.data(), (int)(xys.size() / 3));
  assert(n <= (int)(xys.size() / 3));
  keypoints.reserve(n);
  for (int i = 0; i < n; i++) keypoints.push_back(KeyPoint((float)xys[3 * i], (float)xys[3 * i + 1], 7.f, -1, (float)xys[3 * i + 2]));
}

void resize(const Mat& src, Mat& dst, Size dsize, double, double, int interpolation) {
  assert(interpolation == INTER_AREA && "only the pyramid's INTER_AREA resize is pinned");
  dst.create(dsize.height, dsize.width, src.type());   // no-op when dst already is the ROI of `temp` (ORBextractor.cc:1236-1241)
  const std::vector<uint8_t> s = packed(src);
  std::vector<uint8_t> d((size_t)dsize.width * dsize.height);
  gfo_resize_area(s.data(), src.cols, src.rows, d.data(), dsize.width, dsize.height);
  unpack(d, dst);
}

static inline int reflect101(int p, int n) {
  if (n == 1) return 0;
  while (p < 0 || p >= n) p = p < 0 ? -p : 2 * (n - 1) - p;
  return p;
}
// BORDER_REFLECT_101 (+ BORDER_ISOLATED: the source ROI's own pixels are reflected -- which is also all this stub can see)
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType) {
  assert((borderType & ~BORDER_ISOLATED) == BORDER_REFLECT_101);
  const int h = src.rows + top + bottom, w = src.cols + left + right;
  dst.create(h, w, src.type());
  const std::vector<uint8_t> s = packed(src);            // the source may be the interior of dst (ORBextractor.cc:1243)
  for (int y = 0; y < h; y++) {
    const uint8_t* row = s.data() + (size_t)reflect101(y - top, src.rows) * src.cols;
    uchar* o = dst.ptr(y);
    for (int x = 0; x < w; x++) o[x] = row[reflect101(x - left, src.cols)];
  }
}

void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY, int borderType) {
  assert(ksize.width == 7 && ksize.height == 7 && sigmaX == 2 && sigmaY == 2 && borderType == BORDER_REFLECT_101);
  const std::vector<uint8_t> s = packed(src);
  std::vector<uint8_t> d(s.size());
  gfo_blur7(s.data(), src.cols, src.rows, d.data());
  dst.create(src.rows, src.cols, src.type());
  unpack(d, dst);
}

float fastAtan2(float y, float x) { return gfo_fast_atan2(y, x); }

void KeyPointsFilter::retainBest(std::vector<KeyPoint>&, int) {
  std::cerr << "KeyPointsFilter::retainBest: not provided by the stand-in (dead code in the reference)" << std::endl;
  abort();
}

Scalar sum(const Mat& m) {
  double s = 0;
  for (int y = 0; y < m.rows; y++) {
    if ((m.type() & 7) == CV_32S) { const int* p = m.ptr<int>(y); for (int x = 0; x < m.cols; x++) s += p[x]; }
    else if ((m.type() & 7) == CV_8U) { const uchar* p = m.ptr(y); for (int x = 0; x < m.cols; x++) s += p[x]; }
    else { const float* p = m.ptr<float>(y); for (int x = 0; x < m.cols; x++) s += p[x]; }
  }
  return Scalar(s);
}
Scalar sum(const std::vector<bool>& v) {
  double s = 0;
  for (bool b : v) s += b ? 1 : 0;
  return Scalar(s);
}

}  // namespace cv

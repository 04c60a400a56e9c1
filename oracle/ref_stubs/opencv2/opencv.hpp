// Minimal stand-in for <opencv2/opencv.hpp> -- TEST INFRASTRUCTURE (oracle/_ref), never part of the product.
//
// The reference cannot be built here (no OpenCV / Eigen / PCL headers).  Two of its translation units,
// however, use OpenCV only as a container library plus five image primitives:
//     /root/reference/src/ORBextractor.cc                      (the whole extractor: cell grid, quadtree, IC_Angle, rBRIEF)
//     /root/reference/Thirdparty/GMS/include/gms_matcher.h     (the whole GMS filter)
// This header declares just enough of the cv:: types for those two files to compile UNMODIFIED from where they lie
// (oracle/Makefile, target _ref).  Containers (Mat, Point_, Size, Rect, KeyPoint, DMatch, InputArray ...) are written here
// from the documented OpenCV semantics.  The image primitives -- cv::FAST, cv::resize(INTER_AREA), cv::GaussianBlur,
// cv::copyMakeBorder, cv::fastAtan2 -- are forwarded (cv_stub.cpp) to the restatements in oracle/orb_oracle.cpp that
// tests/test_oracle_orb.py and tests/golden pin bit-exact to the cv2 4.13 wheel.  So in oracle/_ref/libgfs_ref.so every
// line of reference-owned logic is the reference's own compiled source and every OpenCV-owned pixel operation is pinned
// to OpenCV itself.
#pragma once
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <memory>
#include <vector>

typedef unsigned char uchar;

#define CV_PI 3.1415926535897932384626433832795
#define CV_8U 0
#define CV_32S 4
#define CV_32F 5
#define CV_MAKETYPE(depth, cn) ((depth) + (((cn)-1) << 3))
#define CV_8UC1 CV_MAKETYPE(CV_8U, 1)
#define CV_8UC3 CV_MAKETYPE(CV_8U, 3)
#define CV_32SC1 CV_MAKETYPE(CV_32S, 1)

// cvRound: round half to even (SSE2 cvtsd2si under the default rounding mode); cvFloor / cvCeil as documented
static inline int cvRound(double v) { return (int)std::nearbyint(v); }
static inline int cvRound(float v) { return (int)std::nearbyint((double)v); }
static inline int cvRound(int v) { return v; }
static inline int cvFloor(double v) { int i = (int)v; return i - (i > v); }
static inline int cvCeil(double v) { int i = (int)v; return i + (i < v); }

namespace cv {

template <typename T>
struct Point_ {
  T x, y;
  Point_() : x(0), y(0) {}
  Point_(T _x, T _y) : x(_x), y(_y) {}
  Point_& operator*=(float s) { x = (T)(x * s); y = (T)(y * s); return *this; }
  Point_& operator*=(double s) { x = (T)(x * s); y = (T)(y * s); return *this; }
  Point_ operator+(const Point_& o) const { return Point_((T)(x + o.x), (T)(y + o.y)); }
  Point_ operator-(const Point_& o) const { return Point_((T)(x - o.x), (T)(y - o.y)); }
};
typedef Point_<int> Point2i;
typedef Point_<int> Point;
typedef Point_<float> Point2f;

struct Size {
  int width, height;
  Size() : width(0), height(0) {}
  Size(int w, int h) : width(w), height(h) {}
};
struct Rect {
  int x, y, width, height;
  Rect() : x(0), y(0), width(0), height(0) {}
  Rect(int _x, int _y, int w, int h) : x(_x), y(_y), width(w), height(h) {}
};
struct Range {
  int start, end;
  Range(int s, int e) : start(s), end(e) {}
};
struct Scalar {
  double val[4];
  Scalar(double a = 0, double b = 0, double c = 0, double d = 0) { val[0] = a; val[1] = b; val[2] = c; val[3] = d; }
  double operator[](int i) const { return val[i]; }
};

struct KeyPoint {
  Point2f pt;
  float size, angle, response;
  int octave, class_id;
  KeyPoint() : pt(0, 0), size(0), angle(-1), response(0), octave(0), class_id(-1) {}
  KeyPoint(float x, float y, float _size, float _angle = -1, float _response = 0, int _octave = 0, int _class_id = -1)
      : pt(x, y), size(_size), angle(_angle), response(_response), octave(_octave), class_id(_class_id) {}
};
struct DMatch {
  int queryIdx, trainIdx, imgIdx;
  float distance;
  DMatch() : queryIdx(-1), trainIdx(-1), imgIdx(-1), distance(3.402823466e+38f) {}
  DMatch(int q, int t, float d) : queryIdx(q), trainIdx(t), imgIdx(-1), distance(d) {}
};

// reference-counted 2-D array with ROI views (the subset of cv::Mat the two files use)
class Mat {
 public:
  int rows, cols;
  size_t step;  // bytes per row
  uchar* data;

  Mat() : rows(0), cols(0), step(0), data(nullptr), type_(CV_8UC1) {}
  Mat(int r, int c, int type) { alloc(r, c, type); }
  Mat(Size s, int type) { alloc(s.height, s.width, type); }
  Mat(int r, int c, int type, const Scalar& v) { alloc(r, c, type); setTo(v); }
  static Mat zeros(int r, int c, int type) { Mat m(r, c, type); m.setTo(Scalar(0)); return m; }

  int type() const { return type_; }
  bool empty() const { return data == nullptr || rows == 0 || cols == 0; }
  size_t elemSize() const { const int d = type_ & 7, cn = (type_ >> 3) + 1; return (size_t)(d == CV_8U ? 1 : 4) * cn; }
  size_t step1() const { const int d = type_ & 7; return step / (d == CV_8U ? 1 : 4); }
  bool isContinuous() const { return step == (size_t)cols * elemSize(); }
  void create(int r, int c, int type) { if (r != rows || c != cols || type != type_ || !data) alloc(r, c, type); }
  void release() { buf_.reset(); data = nullptr; rows = cols = 0; step = 0; }

  Mat operator()(const Rect& r) const { return view(r.y, r.x, r.height, r.width); }
  Mat rowRange(int a, int b) const { return view(a, 0, b - a, cols); }
  Mat colRange(int a, int b) const { return view(0, a, rows, b - a); }
  Mat row(int i) const { return view(i, 0, 1, cols); }
  Mat clone() const {
    Mat m(rows, cols, type_);
    for (int y = 0; y < rows; y++) memcpy(m.data + y * m.step, data + y * step, (size_t)cols * elemSize());
    return m;
  }
  // copyTo onto an existing array of the same size and type writes in place (cv::Mat::copyTo -> create() is a no-op)
  void copyTo(Mat dst) const {
    assert(dst.rows == rows && dst.cols == cols && dst.type_ == type_ && "stub copyTo: destination must be preallocated");
    for (int y = 0; y < rows; y++) memcpy(dst.data + y * dst.step, data + y * step, (size_t)cols * elemSize());
  }
  Mat& setTo(const Scalar& v) {
    const int d = type_ & 7;
    for (int y = 0; y < rows; y++) {
      if (d == CV_8U) memset(data + y * step, (int)v.val[0], (size_t)cols * elemSize());
      else if (d == CV_32S) { int* p = (int*)(data + y * step); for (int x = 0; x < cols; x++) p[x] = (int)v.val[0]; }
      else { float* p = (float*)(data + y * step); for (int x = 0; x < cols; x++) p[x] = (float)v.val[0]; }
    }
    return *this;
  }
  Mat& setTo(int v) { return setTo(Scalar(v)); }

  template <typename T> T& at(int y, int x) { return *(T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> const T& at(int y, int x) const { return *(const T*)(data + (size_t)y * step + (size_t)x * sizeof(T)); }
  template <typename T> T* ptr(int y = 0) { return (T*)(data + (size_t)y * step); }
  template <typename T> const T* ptr(int y = 0) const { return (const T*)(data + (size_t)y * step); }
  uchar* ptr(int y = 0) { return data + (size_t)y * step; }
  const uchar* ptr(int y = 0) const { return data + (size_t)y * step; }

 private:
  int type_;
  std::shared_ptr<uchar> buf_;
  void alloc(int r, int c, int type) {
    type_ = type; rows = r; cols = c;
    step = (size_t)c * elemSize();
    const size_t n = std::max<size_t>(step * (size_t)r, 1);
    buf_ = std::shared_ptr<uchar>(new uchar[n], std::default_delete<uchar[]>());
    data = buf_.get();
  }
  Mat view(int y, int x, int h, int w) const {
    Mat m;
    m.type_ = type_; m.buf_ = buf_; m.rows = h; m.cols = w; m.step = step;
    m.data = data + (size_t)y * step + (size_t)x * elemSize();
    return m;
  }
};

// InputArray / OutputArray: thin proxies over Mat (what operator()'s signature needs)
class _InputArray {
 public:
  _InputArray() {}
  _InputArray(const Mat& m) : m_(m) {}
  Mat getMat() const { return m_; }
  bool empty() const { return m_.empty(); }
 private:
  Mat m_;
};
class _OutputArray {
 public:
  _OutputArray(Mat& m) : p_(&m) {}
  void create(int r, int c, int type) const { p_->create(r, c, type); }
  void release() const { p_->release(); }
  Mat getMat() const { return *p_; }
 private:
  Mat* p_;
};
typedef const _InputArray& InputArray;
typedef const _OutputArray& OutputArray;
inline _InputArray noArray() { return _InputArray(); }

enum { BORDER_CONSTANT = 0, BORDER_REPLICATE = 1, BORDER_REFLECT = 2, BORDER_WRAP = 3, BORDER_REFLECT_101 = 4, BORDER_ISOLATED = 16 };
enum { INTER_NEAREST = 0, INTER_LINEAR = 1, INTER_CUBIC = 2, INTER_AREA = 3 };

// ---- the image primitives: forwarded to the cv2-pinned restatements (cv_stub.cpp)
void FAST(const Mat& image, std::vector<KeyPoint>& keypoints, int threshold, bool nonmaxSuppression = true);
void resize(const Mat& src, Mat& dst, Size dsize, double fx = 0, double fy = 0, int interpolation = INTER_LINEAR);
void copyMakeBorder(const Mat& src, Mat& dst, int top, int bottom, int left, int right, int borderType);
void GaussianBlur(const Mat& src, Mat& dst, Size ksize, double sigmaX, double sigmaY = 0, int borderType = BORDER_REFLECT_101);
float fastAtan2(float y, float x);

struct KeyPointsFilter {  // only ComputeKeyPointsOld (dead code in the reference, ORBextractor.cc:1160) calls it
  static void retainBest(std::vector<KeyPoint>& keypoints, int npoints);
};

Scalar sum(const Mat& m);
Scalar sum(const std::vector<bool>& v);

// drawing helpers gms_matcher.h's demo utilities name (never called by the matcher itself)
inline void line(Mat&, Point2f, Point2f, const Scalar&) {}
inline void circle(Mat&, Point2f, int, const Scalar&, int = 1) {}

}  // namespace cv

// extern-C entry points into the reference's own code, compiled from /root/reference (oracle/Makefile target _ref):
//   ORB_SLAM3::ORBextractor (src/ORBextractor.cc, include/ORBextractor.h) and gms_matcher (Thirdparty/GMS/include).
// TEST INFRASTRUCTURE: tests compare oracle/*.cpp (the restatement) against these bit for bit.
#include "ORBextractor.h"
#include "gms_matcher.h"

extern "C" {

// ORBextractor(nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST)(image, noArray, keypoints, descriptors, vLappingArea)
// out_kp rows: x y size angle response octave (6 floats); returns the number of keypoints, *mono = the return value
int ref_orb_extract(const uint8_t* img, int w, int h, int nfeatures, float scale, int nlevels, int ini_th, int min_th, int lap0,
                    int lap1, float* out_kp, uint8_t* out_desc, int cap, int* mono) {
  ORB_SLAM3::ORBextractor ex(nfeatures, scale, nlevels, ini_th, min_th);
  cv::Mat image(h, w, CV_8UC1);
  for (int y = 0; y < h; y++) memcpy(image.ptr(y), img + (size_t)y * w, (size_t)w);
  std::vector<cv::KeyPoint> kps;
  cv::Mat desc;
  std::vector<int> lap = {lap0, lap1};
  const int m = ex(cv::_InputArray(image), cv::_InputArray(), kps, cv::_OutputArray(desc), lap);
  if (mono) *mono = m;
  const int n = (int)std::min<size_t>(kps.size(), (size_t)cap);
  for (int i = 0; i < n; i++) {
    out_kp[6 * i] = kps[i].pt.x; out_kp[6 * i + 1] = kps[i].pt.y; out_kp[6 * i + 2] = kps[i].size; out_kp[6 * i + 3] = kps[i].angle;
    out_kp[6 * i + 4] = kps[i].response; out_kp[6 * i + 5] = (float)kps[i].octave;
    memcpy(out_desc + 32 * (size_t)i, desc.ptr(i), 32);
  }
  return (int)kps.size();
}

// level `level` of mvImagePyramid after extracting `img` (tightly packed lw x lh bytes); returns lw | lh << 16
int ref_orb_pyramid_level(const uint8_t* img, int w, int h, float scale, int nlevels, int level, uint8_t* out, int cap) {
  ORB_SLAM3::ORBextractor ex(1000, scale, nlevels, 20, 7);
  cv::Mat image(h, w, CV_8UC1);
  for (int y = 0; y < h; y++) memcpy(image.ptr(y), img + (size_t)y * w, (size_t)w);
  std::vector<cv::KeyPoint> kps;
  cv::Mat desc;
  std::vector<int> lap = {0, 0};
  ex(cv::_InputArray(image), cv::_InputArray(), kps, cv::_OutputArray(desc), lap);
  const cv::Mat& L = ex.mvImagePyramid[level];
  if ((size_t)L.rows * L.cols <= (size_t)cap)
    for (int y = 0; y < L.rows; y++) memcpy(out + (size_t)y * L.cols, L.ptr(y), (size_t)L.cols);
  return L.cols | (L.rows << 16);
}

// gms_matcher(kp1, size1, kp2, size2, matches).GetInlierMask(mask, with_scale, with_rotation); returns the inlier count
int ref_gms(const float* xy1, int n1, int w1, int h1, const float* xy2, int n2, int w2, int h2, const int* matches, int nm,
            int with_scale, int with_rotation, uint8_t* out_mask) {
  std::vector<cv::KeyPoint> k1(n1), k2(n2);
  for (int i = 0; i < n1; i++) k1[i].pt = cv::Point2f(xy1[2 * i], xy1[2 * i + 1]);
  for (int i = 0; i < n2; i++) k2[i].pt = cv::Point2f(xy2[2 * i], xy2[2 * i + 1]);
  std::vector<cv::DMatch> dm(nm);
  for (int i = 0; i < nm; i++) { dm[i].queryIdx = matches[2 * i]; dm[i].trainIdx = matches[2 * i + 1]; dm[i].distance = 0; }
  gms_matcher gms(k1, cv::Size(w1, h1), k2, cv::Size(w2, h2), dm);
  std::vector<bool> mask;
  const int n = gms.GetInlierMask(mask, with_scale != 0, with_rotation != 0);
  for (size_t i = 0; i < mask.size() && i < (size_t)nm; i++) out_mask[i] = mask[i] ? 1 : 0;
  return n;
}

}  // extern "C"

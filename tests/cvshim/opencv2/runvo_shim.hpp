// TEST SHIM — not OpenCV.  What the reference's run_vo.cpp (run_vo.cpp:61-300) uses from OpenCV beyond the value types of
// opencv2/core.hpp, so that the UNMODIFIED file compiles and links against my_slam_adapter/ + libmvo.so in an image without
// OpenCV (tests/test_run_vo_reference_source.py).  Display calls are no-ops; cv::imread decodes PNG through the product's
// apps/png_reader.cpp; cv::Rodrigues is the rotation-matrix -> vector direction run_vo.cpp needs.
#pragma once
#include <cmath>
#include <string>
#include <vector>
#include "opencv2/core.hpp"

namespace mvo_app { bool read_png_bgr(const std::string &path, std::vector<unsigned char> *bgr, int *rows, int *cols, std::string *err); }

namespace cv {
enum { WINDOW_AUTOSIZE = 1 };
struct Scalar { double v[4]; Scalar(double a = 0, double b = 0, double c = 0, double d = 0) : v{a, b, c, d} {} };
inline void namedWindow(const std::string &, int = WINDOW_AUTOSIZE) {}
inline void moveWindow(const std::string &, int, int) {}
inline void imshow(const std::string &, const Mat &) {}
inline int waitKey(int = 0) { return -1; }
inline void destroyAllWindows() {}
inline bool imwrite(const std::string &, const Mat &) { return true; }
inline Mat imread(const std::string &path) {
  std::vector<unsigned char> bgr;
  int rows = 0, cols = 0;
  std::string err;
  if (!mvo_app::read_png_bgr(path, &bgr, &rows, &cols, &err)) return Mat();
  Mat m(rows, cols, CV_8UC3);
  std::memcpy(m.data, bgr.data(), bgr.size());
  return m;
}
inline void drawMatches(const Mat &, const std::vector<KeyPoint> &, const Mat &img2, const std::vector<KeyPoint> &, const std::vector<DMatch> &, Mat &out) { out = img2.clone(); }
inline void drawKeypoints(const Mat &img, const std::vector<KeyPoint> &, Mat &out, const Scalar & = Scalar()) { if (out.data != img.data) out = img.clone(); }
// cv::Rodrigues, rotation matrix (3 x 3, CV_64F) -> rotation vector (3 x 1)
inline void Rodrigues(const Mat &R, Mat &rvec) {
  const double r[9] = {R.at<double>(0, 0), R.at<double>(0, 1), R.at<double>(0, 2), R.at<double>(1, 0), R.at<double>(1, 1), R.at<double>(1, 2),
                       R.at<double>(2, 0), R.at<double>(2, 1), R.at<double>(2, 2)};
  const double c = std::fmin(1.0, std::fmax(-1.0, (r[0] + r[4] + r[8] - 1) / 2)), th = std::acos(c);
  rvec.create(3, 1, CV_64FC1);
  const double s = 2 * std::sin(th), k = th < 1e-12 ? 0.5 : th / s;
  rvec.at<double>(0, 0) = k * (r[7] - r[5]);
  rvec.at<double>(1, 0) = k * (r[2] - r[6]);
  rvec.at<double>(2, 0) = k * (r[3] - r[1]);
}
inline Mat &operator/=(Mat &m, double s) {                        // CV_64F matrices only (run_vo.cpp:262: truth_t /= scale)
  for (int r = 0; r < m.rows; ++r) for (int c = 0; c < m.cols; ++c) m.at<double>(r, c) /= s;
  return m;
}
}  // namespace cv

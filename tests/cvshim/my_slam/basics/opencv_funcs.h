// TEST SHIM of the reference's include/my_slam/basics/opencv_funcs.h: the three helpers run_vo.cpp's display code calls.
#pragma once
#include <cmath>
#include <opencv2/core.hpp>
namespace my_slam {
namespace basics {
inline void getRtFromT(const cv::Mat &T, cv::Mat &R, cv::Mat &t) {
  R.create(3, 3, CV_64FC1);
  t.create(3, 1, CV_64FC1);
  for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R.at<double>(i, j) = T.at<double>(i, j); t.at<double>(i, 0) = T.at<double>(i, 3); }
}
inline double calcMatNorm(const cv::Mat &m) {
  double s = 0;
  for (int r = 0; r < m.rows; ++r) for (int c = 0; c < m.cols; ++c) s += m.at<double>(r, c) * m.at<double>(r, c);
  return std::sqrt(s);
}
inline cv::Point3f transCoord(const cv::Point3f &p, const cv::Mat &R, const cv::Mat &t) {
  const double q[3] = {p.x, p.y, p.z};
  float o[3];
  for (int i = 0; i < 3; ++i) o[i] = (float)(R.at<double>(i, 0) * q[0] + R.at<double>(i, 1) * q[1] + R.at<double>(i, 2) * q[2] + t.at<double>(i, 0));
  return cv::Point3f(o[0], o[1], o[2]);
}
}  // namespace basics
}  // namespace my_slam

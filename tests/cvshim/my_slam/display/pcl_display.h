// TEST SHIM of the reference's include/my_slam/display/pcl_display.h: a viewer that shows nothing (PCL is out of scope,
// SURVEY.md section 2) and reports itself closed, so that run_vo.cpp's "wait for the user to close the window" loop ends.
#pragma once
#include "my_slam/common_include.h"
namespace my_slam {
namespace display {
class PclViewer {
 public:
  typedef std::shared_ptr<PclViewer> Ptr;
  PclViewer(double = 1.0, double = -1.0, double = -1.0, double = -0.5, double = 0, double = 0) {}
  void updateMapPoints(const vector<cv::Point3f> &, const vector<vector<unsigned char>> &) {}
  void updateCurrPoints(const vector<cv::Point3f> &, const vector<vector<unsigned char>> &) {}
  void updatePointsInView(const vector<cv::Point3f> &, const vector<vector<unsigned char>> &) {}
  void updateCameraPose(const cv::Mat &, const cv::Mat &, int) {}
  void updateCameraTruthPose(const cv::Mat &, const cv::Mat &) {}
  void update() {}
  void spinOnce(unsigned int) {}
  bool isStopped() { return true; }
  bool isKeyPressed() { return true; }
};
}  // namespace display
}  // namespace my_slam

// TEST SHIM of the reference's include/my_slam/geometry/camera.h: the Camera value class the VO adapter is handed
// (fx_, fy_, cx_, cy_, K_), nothing else of that header.
#pragma once
#include <memory>
#include <opencv2/core.hpp>
namespace my_slam {
namespace geometry {
class Camera {
 public:
  typedef std::shared_ptr<Camera> Ptr;
  double fx_, fy_, cx_, cy_;
  cv::Mat K_;
  explicit Camera(cv::Mat K) : fx_(K.at<double>(0, 0)), fy_(K.at<double>(1, 1)), cx_(K.at<double>(0, 2)), cy_(K.at<double>(1, 2)), K_(K) {}
};
}  // namespace geometry
}  // namespace my_slam

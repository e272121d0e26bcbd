// TEST SHIM of the reference's include/my_slam/vo/vo_io.h (src/vo/vo_io.cpp) over the product's C ABI for the same formats
// (mvo_image_path, mvo_write_pose_file, mvo_read_pose_file; csrc/vo_io.cpp).
#pragma once
#include <stdexcept>
#include "my_slam/common_include.h"
#include "my_slam/basics/yaml.h"
#include "mvo.h"
namespace my_slam {
namespace vo {
inline vector<string> readImagePaths(const string &dataset_dir, int num_images, const string &image_formatting, bool) {
  vector<string> out;
  for (int i = 0; i < num_images; ++i) {
    char buf[4096];
    if (mvo_image_path(dataset_dir.c_str(), image_formatting.c_str(), i, buf, sizeof buf) != MVO_OK) throw std::runtime_error("readImagePaths: bad format");
    out.push_back(buf);
  }
  return out;
}
inline cv::Mat readCameraIntrinsics(const basics::Yaml &config, bool = true) {
  cv::Mat K(3, 3, CV_64FC1);
  K.at<double>(0, 0) = config.get<double>("camera_info.fx"); K.at<double>(1, 1) = config.get<double>("camera_info.fy");
  K.at<double>(0, 2) = config.get<double>("camera_info.cx"); K.at<double>(1, 2) = config.get<double>("camera_info.cy");
  K.at<double>(2, 2) = 1;
  return K;
}
inline void writePoseToFile(const string filename, vector<cv::Mat> list_T) {
  vector<double> flat;
  for (const cv::Mat &T : list_T) for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) flat.push_back(T.at<double>(i, j));
  if (mvo_write_pose_file(filename.c_str(), flat.data(), (int)list_T.size()) != MVO_OK) throw std::runtime_error("writePoseToFile: " + filename);
}
inline vector<cv::Mat> readPoseFromFile(const string filename) {
  int n = 0;
  mvo_read_pose_file(filename.c_str(), nullptr, 0, &n);
  vector<double> flat((size_t)(n > 0 ? n : 1) * 16);
  if (mvo_read_pose_file(filename.c_str(), flat.data(), n, &n) != MVO_OK) throw std::runtime_error("readPoseFromFile: " + filename);
  vector<cv::Mat> out;
  for (int k = 0; k < n; ++k) {
    cv::Mat T(4, 4, CV_64FC1);
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T.at<double>(i, j) = flat[(size_t)k * 16 + i * 4 + j];
    out.push_back(T);
  }
  return out;
}
}  // namespace vo
}  // namespace my_slam

// TEST SHIM of my_slam::basics::Config (reference include/my_slam/basics/config.h:18-46): same static interface,
// backed by a map that holds the values of the shipped config/config.yaml instead of a cv::FileStorage.
#pragma once
#include <cmath>
#include <map>
#include <stdexcept>
#include <string>
namespace my_slam {
namespace basics {
class Config {
 public:
  static std::map<std::string, double> &table() {
    static std::map<std::string, double> t = {
        {"number_of_keypoints_to_extract", 8000}, {"max_number_of_keypoints", 1500}, {"scale_factor", 1.2},
        {"level_pyramid", 4}, {"score_threshold", 20}, {"kpts_uniform_selection_grid_size", 16},
        {"kpts_uniform_selection_max_pts_per_grid", 8}, {"xiang_gao_method_match_ratio", 2},
        {"lowe_method_dist_ratio", 0.8}, {"findEssentialMat_prob", 0.999}, {"findEssentialMat_threshold", 1.0}};
    return t;
  }
  template <typename T> static T get(const std::string &key) {
    auto it = table().find(key);
    if (it == table().end()) throw std::runtime_error("Key " + key + " doesn't exist");     // config.cpp:35
    return convert<T>(it->second);
  }

 private:
  // cv::FileNode -> int goes through cvRound (0.8 -> 1: what the reference's get<int>("lowe_method_dist_ratio") yields)
  template <typename T> static T convert(double v) { return static_cast<T>(v); }
};
template <> inline int Config::convert<int>(double v) { return (int)std::nearbyint(v); }
}  // namespace basics
}  // namespace my_slam

// TEST SHIM of my_slam::basics::Config (reference include/my_slam/basics/config.h:18-46): same static interface,
// backed by a map that holds the values of the shipped config/config.yaml instead of a cv::FileStorage.
#pragma once
#include <cmath>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include "mvo.h"
namespace my_slam {
namespace basics {
class Config {
 public:
  // Config::setParameterFile (config.cpp:12-24): from then on the keys come from that file (the product's reader of the
  // reference's config dialect, mvo_config_*); without it the built-in table of the shipped values below answers
  static std::shared_ptr<mvo_config> &file() { static std::shared_ptr<mvo_config> f; return f; }
  static void setParameterFile(const std::string &filename) {
    mvo_config *c = nullptr;
    if (mvo_config_load(filename.c_str(), &c) != MVO_OK) throw std::runtime_error("parameter file " + filename + " does not exist.");
    file() = std::shared_ptr<mvo_config>(c, mvo_config_free);
  }
  static std::map<std::string, double> &table() {
    static std::map<std::string, double> t = {
        {"number_of_keypoints_to_extract", 8000}, {"max_number_of_keypoints", 1500}, {"scale_factor", 1.2},
        {"level_pyramid", 4}, {"score_threshold", 20}, {"kpts_uniform_selection_grid_size", 16},
        {"kpts_uniform_selection_max_pts_per_grid", 8}, {"xiang_gao_method_match_ratio", 2},
        {"lowe_method_dist_ratio", 0.8}, {"findEssentialMat_prob", 0.999}, {"findEssentialMat_threshold", 1.0},
        {"feature_match_method_index_initialization", 1}, {"feature_match_method_index_pnp", 1},
        {"max_matching_pixel_dist_in_initialization", 100}, {"max_matching_pixel_dist_in_triangulation", 100},
        {"max_matching_pixel_dist_in_pnp", 50}, {"min_triang_angle", 1.0}, {"max_ratio_between_max_angle_and_median_angle", 20},
        {"min_inlier_matches", 15}, {"min_pixel_dist", 50}, {"min_median_triangulation_angle", 2.0},
        {"assumed_mean_pts_depth_during_vo_init", 0.8}, {"min_dist_between_two_keyframes", 0.03},
        {"max_possible_dist_to_prev_keyframe", 0.3}, {"num_prev_frames_to_opti_by_ba", 5}};
    return t;
  }
  static std::map<std::string, std::string> &strings() {           // the string-valued keys of the shipped file
    static std::map<std::string, std::string> t = {
        {"is_enable_ba", "true"}, {"is_ba_fix_map_points", "true"}, {"information_matrix", "1.0 0.0 0.0 1.0"}};
    return t;
  }
  static bool getBool(const std::string &key) {                    // config.h: "true" / "True"
    if (file()) {
      int b = 0;
      if (mvo_config_get_bool(file().get(), key.c_str(), &b) != MVO_OK) throw std::runtime_error("Key " + key + " doesn't exist");
      return b != 0;
    }
    auto it = strings().find(key);
    if (it == strings().end()) throw std::runtime_error("Key " + key + " doesn't exist");
    return it->second == "true" || it->second == "True";
  }
  template <typename T> static T get(const std::string &key) {
    if (file()) {
      double v = 0;
      if (mvo_config_get_double(file().get(), key.c_str(), &v) != MVO_OK) throw std::runtime_error("Key " + key + " doesn't exist");
      return convert<T>(v);
    }
    auto it = table().find(key);
    if (it == table().end()) throw std::runtime_error("Key " + key + " doesn't exist");     // config.cpp:35
    return convert<T>(it->second);
  }

 private:
  // cv::FileNode -> int goes through cvRound (0.8 -> 1: what the reference's get<int>("lowe_method_dist_ratio") yields)
  template <typename T> static T convert(double v) { return static_cast<T>(v); }
};
template <> inline int Config::convert<int>(double v) { return (int)std::nearbyint(v); }
template <> inline std::string Config::get<std::string>(const std::string &key) {
  if (file()) {
    char buf[4096];
    if (mvo_config_get_string(file().get(), key.c_str(), buf, sizeof buf) != MVO_OK) throw std::runtime_error("Key " + key + " doesn't exist");
    return buf;
  }
  auto it = strings().find(key);
  if (it == strings().end()) throw std::runtime_error("Key " + key + " doesn't exist");
  return it->second;
}
}  // namespace basics
}  // namespace my_slam

// TEST SHIM of my_slam::basics::Yaml (reference include/my_slam/basics/yaml.h): the interface run_vo.cpp uses, backed by the
// product's reader of the reference's config dialect (mvo_config_*, csrc/vo_io.cpp) instead of cv::FileStorage.
#pragma once
#include <assert.h>
#include <memory>
#include <stdexcept>
#include <string>
#include "mvo.h"
namespace my_slam {
namespace basics {
class Yaml {
 public:
  explicit Yaml(const std::string &filename) {
    mvo_config *c = nullptr;
    if (mvo_config_load(filename.c_str(), &c) != MVO_OK) throw std::runtime_error("Yaml: cannot read " + filename);
    cfg_ = std::shared_ptr<mvo_config>(c, mvo_config_free);
  }
  Yaml get(const std::string &key) const { Yaml y(*this); y.prefix_ = prefix_ + key + "/"; return y; }        // a dataset section
  template <typename T> T get(const std::string &key) const;
  bool getBool(const std::string &key) const;
 private:
  std::shared_ptr<mvo_config> cfg_;
  std::string prefix_;
  std::string path(const std::string &key) const { return prefix_ + key; }
};
template <> inline std::string Yaml::get<std::string>(const std::string &key) const {
  char buf[4096];
  if (mvo_config_get_string(cfg_.get(), path(key).c_str(), buf, sizeof buf) != MVO_OK) throw std::runtime_error("Key " + key + " doesn't exist");
  return buf;
}
template <> inline double Yaml::get<double>(const std::string &key) const {
  double v = 0;
  if (mvo_config_get_double(cfg_.get(), path(key).c_str(), &v) != MVO_OK) throw std::runtime_error("Key " + key + " doesn't exist");
  return v;
}
template <> inline int Yaml::get<int>(const std::string &key) const {
  int v = 0;
  if (mvo_config_get_int(cfg_.get(), path(key).c_str(), &v) != MVO_OK) throw std::runtime_error("Key " + key + " doesn't exist");
  return v;
}
inline bool Yaml::getBool(const std::string &key) const { const std::string v = get<std::string>(key); return v == "true" || v == "True"; }
}  // namespace basics
}  // namespace my_slam

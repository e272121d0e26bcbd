// TEST SHIM of the reference's include/my_slam/basics/basics.h: int2str and makedirs, which run_vo.cpp calls.
#pragma once
#include <sys/stat.h>
#include <string>
namespace my_slam {
namespace basics {
inline std::string int2str(int num, int width, char char_to_fill = '0') {
  std::string s = std::to_string(num);
  if ((int)s.size() < width) s.insert(0, (size_t)width - s.size(), char_to_fill);
  return s;
}
inline bool makedirs(const std::string &dir) {                    // like os.makedirs
  for (size_t i = 1; i <= dir.size(); ++i)
    if (i == dir.size() || dir[i] == '/') mkdir(dir.substr(0, i).c_str(), 0777);
  struct stat st;
  return stat(dir.c_str(), &st) == 0;
}
}  // namespace basics
}  // namespace my_slam
#include "my_slam/basics/opencv_funcs.h"

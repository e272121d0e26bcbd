// TEST SHIM of the reference's include/my_slam/common_include.h: what the adapter sources need from it.
#pragma once
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include <opencv2/core.hpp>
using namespace std;

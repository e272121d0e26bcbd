// TEST SHIM of the reference's include/my_slam/common_include.h: what the adapter sources and run_vo.cpp need from it.
#pragma once
#include <iostream>
#include <memory>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>
#include <opencv2/core.hpp>
#include <opencv2/runvo_shim.hpp>
using namespace std;

// TEST SHIM: run_vo.cpp includes the reference's geometry/motion_estimation.h without calling anything from it.
#pragma once
#include "my_slam/common_include.h"

// TEST SHIM — not OpenCV: <opencv2/features2d/features2d.hpp> for run_vo.cpp (drawMatches / drawKeypoints are no-ops).
#pragma once
#include "opencv2/core.hpp"
#include "opencv2/runvo_shim.hpp"

// TEST SHIM — not OpenCV: the highgui calls of the reference's run_vo.cpp (window functions are no-ops, cv::imread reads PNG
// files through the product's own decoder apps/png_reader.cpp, cv::imwrite does nothing).
#pragma once
#include "opencv2/core.hpp"
#include "opencv2/runvo_shim.hpp"

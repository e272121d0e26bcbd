// TEST SHIM — not OpenCV: <opencv2/core/core.hpp> as the reference's run_vo.cpp includes it (see opencv2/core.hpp here).
#pragma once
#include "opencv2/core.hpp"
#include "opencv2/runvo_shim.hpp"

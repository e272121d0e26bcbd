// Takes the place of the reference's include/my_slam/vo/frame.h: the class lives in my_slam_adapter/vo_mvo.h.
#pragma once
#include "vo_mvo.h"

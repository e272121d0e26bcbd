// One libmvo context per process, the way the reference keeps its ORB objects, matchers and latched config values
// in function-local statics (reference src/geometry/feature_match.cpp:16-23,42-45,56-62,137-141).
#pragma once
#include <stdexcept>
#include <string>
#include "mvo.h"

namespace my_slam {
namespace mvo_adapter {

// Created on first use from my_slam::basics::Config (the keys the reference reads in feature_match.cpp).
// Throws std::runtime_error when no sm_100 GPU is usable: libmvo has no CPU path.
mvo_ctx *context();
const mvo_params &params();
// Reference-style error behaviour: libmvo status codes become the exceptions the reference throws.
void check(int rc, const char *where);

}  // namespace mvo_adapter
}  // namespace my_slam

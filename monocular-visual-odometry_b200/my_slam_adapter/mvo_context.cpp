#include "mvo_context.h"
#include "my_slam/basics/config.h"

namespace my_slam {
namespace mvo_adapter {

namespace {
mvo_params g_params;
mvo_ctx *create() {
  using basics::Config;
  mvo_default_params(&g_params);
  // feature_match.cpp:16-21 (ORB), :56-58 (grid selection), :137-139 (match ratios; read with get<int> like the reference)
  g_params.orb_nfeatures = Config::get<int>("number_of_keypoints_to_extract");
  g_params.orb_scale_factor = (float)Config::get<double>("scale_factor");
  g_params.orb_nlevels = Config::get<int>("level_pyramid");
  g_params.orb_fast_threshold = Config::get<int>("score_threshold");
  g_params.max_keypoints = Config::get<int>("max_number_of_keypoints");
  g_params.grid_size = Config::get<int>("kpts_uniform_selection_grid_size");
  g_params.max_pts_per_grid = Config::get<int>("kpts_uniform_selection_max_pts_per_grid");
  g_params.xiang_gao_ratio = Config::get<int>("xiang_gao_method_match_ratio");
  g_params.lowe_ratio = Config::get<int>("lowe_method_dist_ratio");
  mvo_ctx *c = nullptr;
  const int rc = mvo_create(&c, /*device*/ 0, &g_params);
  if (rc != MVO_OK) throw std::runtime_error("libmvo: mvo_create failed (" + std::to_string(rc) + "): no usable sm_100 GPU, and there is no CPU path");
  return c;
}
}  // namespace

mvo_ctx *context() {
  static mvo_ctx *ctx = create();
  return ctx;
}

const mvo_params &params() {
  context();
  return g_params;
}

void check(int rc, const char *where) {
  if (rc == MVO_OK) return;
  throw std::runtime_error(std::string(where) + ": " + mvo_last_error(context()));
}

}  // namespace mvo_adapter
}  // namespace my_slam

// Replacement for the one OpenCV call the VO layer makes inline, reference src/vo/vo.cpp:318-320:
//   cv::solvePnPRansac(pts_3d, pts_2d, K, cv::Mat(), R_vec, t, useExtrinsicGuess=false, iterationsCount=100,
//                      reprojectionError=2.0, confidence=0.999, pnp_inliers_mask);
// Same argument meaning and outputs: R_vec, t are 3x1 CV_64F, inliers is K x 1 CV_32SC1 holding ascending INDICES
// (vo.cpp:308,333-337).  iterationsCount / confidence do not apply: all hypotheses (mvo_params::pnp_hypotheses) are
// scored in one batch.  Returns false (and leaves inliers empty) when no model reaches 4 inliers.
#pragma once
#include <vector>
#include <opencv2/core.hpp>
#include "mvo_context.h"

namespace my_slam {
namespace mvo_adapter {

inline bool solvePnPRansac(const std::vector<cv::Point3f> &pts_3d, const std::vector<cv::Point2f> &pts_2d, const cv::Mat &K, cv::Mat &R_vec,
                           cv::Mat &t, float reprojectionError, cv::Mat &inliers) {
  static_assert(sizeof(cv::Point3f) == 12 && sizeof(cv::Point2f) == 8, "contiguous float triples / pairs");
  mvo_ctx *ctx = context();
  mvo_params p;
  mvo_get_params(ctx, &p);
  if (p.pnp_reproj_error != reprojectionError) { p.pnp_reproj_error = reprojectionError; check(mvo_set_params(ctx, &p), "solvePnPRansac"); }
  const int n = (int)pts_3d.size();
  double rvec[3], tvec[3];
  std::vector<int32_t> inl((size_t)(n > 0 ? n : 1));
  int n_inl = n;
  const double Kf[9] = {K.at<double>(0, 0), K.at<double>(0, 1), K.at<double>(0, 2), K.at<double>(1, 0), K.at<double>(1, 1),
                        K.at<double>(1, 2), K.at<double>(2, 0), K.at<double>(2, 1), K.at<double>(2, 2)};
  const int rc = mvo_solve_pnp_ransac(ctx, n ? &pts_3d[0].x : nullptr, n ? &pts_2d[0].x : nullptr, n, Kf, rvec, tvec, inl.data(), &n_inl);
  R_vec.create(3, 1, CV_64FC1);
  t.create(3, 1, CV_64FC1);
  if (rc == MVO_ERR_DEGENERATE) { inliers = cv::Mat(); return false; }
  check(rc, "solvePnPRansac");
  for (int i = 0; i < 3; ++i) { R_vec.at<double>(i, 0) = rvec[i]; t.at<double>(i, 0) = tvec[i]; }
  inliers.create(n_inl, 1, CV_32SC1);
  for (int i = 0; i < n_inl; ++i) inliers.at<int>(i, 0) = inl[(size_t)i];
  return true;
}

}  // namespace mvo_adapter
}  // namespace my_slam

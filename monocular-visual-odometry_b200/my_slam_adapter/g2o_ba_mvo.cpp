// Drop-in bodies for the reference's src/optimization/g2o_ba.cpp (declarations: include/my_slam/optimization/g2o_ba.h:16-30).
// The reference hands g2o raw pointers into live objects (cv::Point2f* into Frame::keypoints_, cv::Point3f* into
// MapPoint::pos_, cv::Mat* = Frame::T_w_c_) and g2o's results are written back through them (g2o_ba.cpp:298-316).
// Here the graph is flattened into the arrays of mvo_bundle_adjustment, solved on the GPU, and scattered back through
// the same pointers.
#include <cstring>
#include "my_slam/optimization/g2o_ba.h"
#include "mvo_context.h"

namespace my_slam {
namespace optimization {

using mvo_adapter::check;
using mvo_adapter::context;

// reference g2o_ba.cpp:172-317
void bundleAdjustment(const vector<vector<cv::Point2f *>> &v_pts_2d, const vector<vector<int>> &v_pts_2d_to_3d_idx, const cv::Mat &K,
                      std::unordered_map<int, cv::Point3f *> &pts_3d, vector<cv::Mat *> &v_camera_g2o_poses,
                      const cv::Mat &information_matrix, bool is_fix_map_pts, bool is_update_map_pts) {
  const int F = (int)v_camera_g2o_poses.size();
  vector<double> poses((size_t)F * 16);
  vector<float> pts, obs;
  vector<int32_t> ef, ep;
  vector<int> ids;                                   // vertex order = first appearance in the edge list
  std::unordered_map<int, int> id2idx;
  for (int f = 0; f < F; ++f) {
    for (int r = 0; r < 4; ++r) std::memcpy(&poses[(size_t)f * 16 + 4 * r], v_camera_g2o_poses[f]->ptr<double>(r), 4 * sizeof(double));
    for (size_t j = 0; j < v_pts_2d[f].size(); ++j) {
      const int id = v_pts_2d_to_3d_idx[f][j];
      auto it = id2idx.find(id);
      if (it == id2idx.end()) {
        it = id2idx.emplace(id, (int)ids.size()).first;
        ids.push_back(id);
        const cv::Point3f *p = pts_3d.at(id);
        pts.push_back(p->x); pts.push_back(p->y); pts.push_back(p->z);
      }
      ef.push_back(f);
      ep.push_back(it->second);
      obs.push_back(v_pts_2d[f][j]->x);
      obs.push_back(v_pts_2d[f][j]->y);
    }
  }
  const double Kf[9] = {K.at<double>(0, 0), K.at<double>(0, 1), K.at<double>(0, 2), K.at<double>(1, 0), K.at<double>(1, 1),
                        K.at<double>(1, 2), K.at<double>(2, 0), K.at<double>(2, 1), K.at<double>(2, 2)};
  const double info[4] = {information_matrix.at<double>(0, 0), information_matrix.at<double>(0, 1),
                          information_matrix.at<double>(1, 0), information_matrix.at<double>(1, 1)};
  check(mvo_bundle_adjustment(context(), poses.data(), F, pts.data(), (int)ids.size(), ef.data(), ep.data(), obs.data(), (int)ef.size(), Kf,
                              info, is_fix_map_pts ? 1 : 0, is_update_map_pts ? 1 : 0, nullptr),
        "bundleAdjustment");
  for (int f = 0; f < F; ++f)                        // g2o_ba.cpp:298-305: poses back in place
    for (int r = 0; r < 4; ++r) std::memcpy(v_camera_g2o_poses[f]->ptr<double>(r), &poses[(size_t)f * 16 + 4 * r], 4 * sizeof(double));
  if (is_update_map_pts)                             // g2o_ba.cpp:308-316
    for (size_t k = 0; k < ids.size(); ++k) {
      cv::Point3f *p = pts_3d.at(ids[k]);
      p->x = pts[3 * k]; p->y = pts[3 * k + 1]; p->z = pts[3 * k + 2];
    }
}

// reference g2o_ba.cpp:34-145 (dead code in the shipped pipeline, vo.cpp:456-470, exported for completeness)
void optimizeSingleFrame(const vector<cv::Point2f *> &points_2d, const cv::Mat &K, vector<cv::Point3f *> &points_3d,
                         cv::Mat &cam_pose_in_world, bool is_fix_map_pts, bool is_update_map_pts) {
  const int n = (int)points_3d.size();
  vector<float> pts((size_t)n * 3), obs((size_t)n * 2);
  for (int i = 0; i < n; ++i) {
    pts[3 * i] = points_3d[i]->x; pts[3 * i + 1] = points_3d[i]->y; pts[3 * i + 2] = points_3d[i]->z;
    obs[2 * i] = points_2d[i]->x; obs[2 * i + 1] = points_2d[i]->y;
  }
  double pose[16];
  for (int r = 0; r < 4; ++r) std::memcpy(pose + 4 * r, cam_pose_in_world.ptr<double>(r), 4 * sizeof(double));
  const double Kf[9] = {K.at<double>(0, 0), K.at<double>(0, 1), K.at<double>(0, 2), K.at<double>(1, 0), K.at<double>(1, 1),
                        K.at<double>(1, 2), K.at<double>(2, 0), K.at<double>(2, 1), K.at<double>(2, 2)};
  check(mvo_optimize_single_frame(context(), pose, pts.data(), obs.data(), n, Kf, is_fix_map_pts ? 1 : 0, is_update_map_pts ? 1 : 0),
        "optimizeSingleFrame");
  for (int r = 0; r < 4; ++r) std::memcpy(cam_pose_in_world.ptr<double>(r), pose + 4 * r, 4 * sizeof(double));
  if (is_update_map_pts)
    for (int i = 0; i < n; ++i) { points_3d[i]->x = pts[3 * i]; points_3d[i]->y = pts[3 * i + 1]; points_3d[i]->z = pts[3 * i + 2]; }
}

}  // namespace optimization
}  // namespace my_slam

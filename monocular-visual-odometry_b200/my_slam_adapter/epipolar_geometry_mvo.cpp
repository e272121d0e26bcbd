// Drop-in bodies for the four estimation functions of the reference's src/geometry/epipolar_geometry.cpp (declarations:
// include/my_slam/geometry/epipolar_geometry.h:21-27, :32-39, :50-53, :58-63).  The error helpers of that file
// (computeEpipolarConsError, ...) are plain host arithmetic and stay as they are: in an integration these bodies replace
// the originals inside epipolar_geometry.cpp.  (The homography pair forwards to kernels that have not run on hardware
// yet, see tests/test_homography_gpu.py.)
#include "my_slam/geometry/epipolar_geometry.h"
#include "my_slam/basics/config.h"
#include "mvo_context.h"

namespace my_slam {
namespace geometry {

using mvo_adapter::check;
using mvo_adapter::context;

static void K9_of(const cv::Mat &K, double *k) {
  for (int i = 0; i < 9; ++i) k[i] = K.at<double>(i / 3, i % 3);
}
static cv::Mat mat_of(const double *v, int rows, int cols) {
  cv::Mat m(rows, cols, CV_64FC1);
  for (int i = 0; i < rows * cols; ++i) m.at<double>(i / cols, i % cols) = v[i];
  return m;
}

// reference epipolar_geometry.cpp:17-57: findEssentialMat (RANSAC) + E /= E(2,2) + inliers from the mask + recoverPose + t /= |t|
void estiMotionByEssential(const vector<cv::Point2f> &pts_in_img1, const vector<cv::Point2f> &pts_in_img2, const cv::Mat &camera_intrinsics,
                           cv::Mat &essential_matrix, cv::Mat &R, cv::Mat &t, vector<int> &inliers_index) {
  inliers_index.clear();
  static const double threshold = basics::Config::get<double>("findEssentialMat_threshold");     // :31 (prob, :30, has no counterpart)
  const int n = (int)pts_in_img1.size();
  double K[9], E[9], Rm[9], tv[3];
  K9_of(camera_intrinsics, K);
  std::vector<int32_t> inl((size_t)(n > 0 ? n : 1));
  int n_inl = n;
  check(mvo_esti_motion_by_essential(context(), n ? &pts_in_img1[0].x : nullptr, n ? &pts_in_img2[0].x : nullptr, n, K, threshold, E, Rm, tv,
                                     inl.data(), &n_inl), "estiMotionByEssential");
  essential_matrix = mat_of(E, 3, 3);
  R = mat_of(Rm, 3, 3);
  t = mat_of(tv, 3, 1);
  inliers_index.assign(inl.begin(), inl.begin() + n_inl);
}

// reference epipolar_geometry.cpp:90-128: findHomography (RANSAC, 3 px) + H /= H(2,2) + inliers + decomposeHomographyMat + t /= |t|
void estiMotionByHomography(const vector<cv::Point2f> &pts_in_img1, const vector<cv::Point2f> &pts_in_img2, const cv::Mat &camera_intrinsics,
                            cv::Mat &homography_matrix, vector<cv::Mat> &Rs, vector<cv::Mat> &ts, vector<cv::Mat> &normals,
                            vector<int> &inliers_index) {
  Rs.clear(); ts.clear(); normals.clear(); inliers_index.clear();                       // :97-100
  const double ransacReprojThreshold = 3;                                                 // :103
  const int n = (int)pts_in_img1.size();
  double K[9], H[9], R4[36], t4[12], n4[12];
  K9_of(camera_intrinsics, K);
  std::vector<int32_t> inl((size_t)(n > 0 ? n : 1));
  int n_inl = n, n_sol = 0;
  check(mvo_esti_motion_by_homography(context(), n ? &pts_in_img1[0].x : nullptr, n ? &pts_in_img2[0].x : nullptr, n, K, ransacReprojThreshold, H,
                                      R4, t4, n4, &n_sol, inl.data(), &n_inl), "estiMotionByHomography");
  homography_matrix = mat_of(H, 3, 3);
  for (int s = 0; s < n_sol; ++s) { Rs.push_back(mat_of(R4 + 9 * s, 3, 3)); ts.push_back(mat_of(t4 + 3 * s, 3, 1)); normals.push_back(mat_of(n4 + 3 * s, 3, 1)); }
  inliers_index.assign(inl.begin(), inl.begin() + n_inl);
}

// reference epipolar_geometry.cpp:59-88: cv::filterHomographyDecompByVisibleRefpoints on the inlier points
void removeWrongRtOfHomography(const vector<cv::Point2f> &pts_on_np1, const vector<cv::Point2f> &pts_on_np2, const vector<int> &inliers,
                               vector<cv::Mat> &Rs, vector<cv::Mat> &ts, vector<cv::Mat> &normals) {
  int n_sol = (int)Rs.size();
  if (n_sol > 4) throw std::runtime_error("removeWrongRtOfHomography: more than 4 solutions");
  double R4[36], t4[12], n4[12];
  for (int s = 0; s < n_sol; ++s) {
    K9_of(Rs[s], R4 + 9 * s);
    for (int i = 0; i < 3; ++i) { t4[3 * s + i] = ts[s].at<double>(i, 0); n4[3 * s + i] = normals[s].at<double>(i, 0); }
  }
  const int n = (int)pts_on_np1.size(), ni = (int)inliers.size();
  check(mvo_remove_wrong_rt_of_homography(context(), n ? &pts_on_np1[0].x : nullptr, n ? &pts_on_np2[0].x : nullptr, n,
                                          ni ? reinterpret_cast<const int32_t *>(inliers.data()) : nullptr, ni, R4, t4, n4, &n_sol),
        "removeWrongRtOfHomography");
  Rs.clear(); ts.clear(); normals.clear();
  for (int s = 0; s < n_sol; ++s) { Rs.push_back(mat_of(R4 + 9 * s, 3, 3)); ts.push_back(mat_of(t4 + 3 * s, 3, 1)); normals.push_back(mat_of(n4 + 3 * s, 3, 1)); }
}

// reference epipolar_geometry.cpp:130-175: cv::triangulatePoints([I|0], [R|t], inlier points) and the division by w
void doTriangulation(const vector<cv::Point2f> &pts_on_np1, const vector<cv::Point2f> &pts_on_np2, const cv::Mat &R_cam2_to_cam1,
                     const cv::Mat &t_cam2_to_cam1, const vector<int> &inliers, vector<cv::Point3f> &pts3d_in_cam1) {
  static_assert(sizeof(cv::Point3f) == 12 && sizeof(cv::Point2f) == 8 && sizeof(int) == sizeof(int32_t), "flat float / int32 arrays");
  double Rm[9], tv[3];
  K9_of(R_cam2_to_cam1, Rm);
  for (int i = 0; i < 3; ++i) tv[i] = t_cam2_to_cam1.at<double>(i, 0);
  const int n = (int)pts_on_np1.size(), ni = (int)inliers.size();
  pts3d_in_cam1.assign((size_t)ni, cv::Point3f());
  check(mvo_do_triangulation(context(), n ? &pts_on_np1[0].x : nullptr, n ? &pts_on_np2[0].x : nullptr, n, Rm, tv,
                             ni ? reinterpret_cast<const int32_t *>(inliers.data()) : nullptr, ni, ni ? &pts3d_in_cam1[0].x : nullptr),
        "doTriangulation");
}

}  // namespace geometry
}  // namespace my_slam

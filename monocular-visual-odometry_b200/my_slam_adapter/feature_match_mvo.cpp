// Drop-in bodies for the reference's src/geometry/feature_match.cpp: every function keeps the signature declared in
// include/my_slam/geometry/feature_match.h:12-54 and forwards to the C ABI of libmvo.so (include/mvo.h).  Build the
// reference's `geometry` library from this file instead of src/geometry/feature_match.cpp and link it against libmvo.
//   cv::KeyPoint (pt, size, angle, response, octave, class_id: 28 bytes) == mvo_keypoint
//   cv::DMatch   (queryIdx, trainIdx, imgIdx, distance: 16 bytes)        == mvo_dmatch
// so vectors are passed by pointer without conversion.
#include <algorithm>
#include <cassert>
#include <cmath>
#include <cstdio>
#include "my_slam/geometry/feature_match.h"
#include "mvo_context.h"

static_assert(sizeof(cv::KeyPoint) == sizeof(mvo_keypoint), "cv::KeyPoint layout");
static_assert(sizeof(cv::DMatch) == sizeof(mvo_dmatch), "cv::DMatch layout");

namespace my_slam {
namespace geometry {

using mvo_adapter::check;
using mvo_adapter::context;

// reference feature_match.cpp:11-36: cv::ORB::detect + selectUniformKptsByGrid
void calcKeyPoints(const cv::Mat &image, vector<cv::KeyPoint> &keypoints) {
  keypoints.resize((size_t)mvo_adapter::params().max_keypoints + 1);     // the grid selection keeps up to max + 1 (:77)
  int n = (int)keypoints.size();
  check(mvo_calc_keypoints(context(), image.data, image.rows, image.cols, image.channels(), image.step,
                           reinterpret_cast<mvo_keypoint *>(keypoints.data()), &n), "calcKeyPoints");
  keypoints.resize((size_t)n);
}

// reference feature_match.cpp:38-49: cv::ORB::compute (keypoints level-sorted, as calcKeyPoints returns them)
void calcDescriptors(const cv::Mat &image, vector<cv::KeyPoint> &keypoints, cv::Mat &descriptors) {
  descriptors.create((int)keypoints.size(), 32, CV_8UC1);
  check(mvo_calc_descriptors(context(), image.data, image.rows, image.cols, image.channels(), image.step,
                             reinterpret_cast<const mvo_keypoint *>(keypoints.data()), (int)keypoints.size(), descriptors.data),
        "calcDescriptors");
}

// reference feature_match.cpp:51-84
void selectUniformKptsByGrid(vector<cv::KeyPoint> &keypoints, int image_rows, int image_cols) {
  int n = (int)keypoints.size();
  check(mvo_select_uniform_kpts_by_grid(context(), reinterpret_cast<mvo_keypoint *>(keypoints.data()), &n, image_rows, image_cols),
        "selectUniformKptsByGrid");
  keypoints.resize((size_t)n);
}

static void kpts_xy(const vector<cv::KeyPoint> &k, vector<float> &xy) {
  xy.resize(2 * k.size());
  for (size_t i = 0; i < k.size(); ++i) { xy[2 * i] = k[i].pt.x; xy[2 * i + 1] = k[i].pt.y; }
}

// reference feature_match.cpp:86-124
vector<cv::DMatch> matchByRadiusAndBruteForce(const vector<cv::KeyPoint> &keypoints_1, const vector<cv::KeyPoint> &keypoints_2,
                                              const cv::Mat1b &descriptors_1, const cv::Mat1b &descriptors_2,
                                              float max_matching_pixel_dist) {
  assert(descriptors_1.rows == (int)keypoints_1.size() && descriptors_2.rows == (int)keypoints_2.size());     // :94
  vector<float> xy1, xy2;
  kpts_xy(keypoints_1, xy1);
  kpts_xy(keypoints_2, xy2);
  vector<cv::DMatch> matches((size_t)std::max(descriptors_1.rows, 1));
  int n = 0;
  check(mvo_match_radius_sad(context(), descriptors_1.data, xy1.data(), descriptors_1.rows, descriptors_2.data, xy2.data(),
                             descriptors_2.rows, max_matching_pixel_dist, reinterpret_cast<mvo_dmatch *>(matches.data()), &n),
        "matchByRadiusAndBruteForce");
  matches.resize((size_t)n);
  return matches;
}

// reference feature_match.cpp:126-239 (method 1: the exact nearest-neighbour search that the reference's FLANN-LSH
// approximates; 2: knn2 + ratio; 3: radius-gated SAD), thresholds and duplicate removal included
void matchFeatures(const cv::Mat1b &descriptors_1, const cv::Mat1b &descriptors_2, vector<cv::DMatch> &matches, int method_index,
                   bool is_print_res, const vector<cv::KeyPoint> &keypoints_1, const vector<cv::KeyPoint> &keypoints_2,
                   float max_matching_pixel_dist) {
  if (method_index < 1 || method_index > 3)
    throw std::runtime_error("feature_match.cpp::matchFeatures: wrong method index.");     // :225
  vector<float> xy1, xy2;
  if (method_index == 3) { kpts_xy(keypoints_1, xy1); kpts_xy(keypoints_2, xy2); }
  matches.assign((size_t)std::max(descriptors_1.rows, 1), cv::DMatch());
  int n = 0;
  check(mvo_match_features(context(), descriptors_1.data, descriptors_1.rows, descriptors_2.data, descriptors_2.rows, method_index,
                           xy1.data(), xy2.data(), max_matching_pixel_dist, reinterpret_cast<mvo_dmatch *>(matches.data()), &n),
        "matchFeatures");
  matches.resize((size_t)n);
  if (is_print_res) {                                                                       // :231-238
    double mn = 1e30, mx = 0;
    for (const cv::DMatch &m : matches) { mn = std::min(mn, (double)m.distance); mx = std::max(mx, (double)m.distance); }
    printf("Matching features:\n");
    printf("Using method %d, number of matches: %d\n", method_index, (int)matches.size());
    printf("-- Max dist : %f \n", matches.empty() ? 0.0 : mx);
    printf("-- Min dist : %f \n", matches.empty() ? 0.0 : mn);
  }
}

// reference feature_match.cpp:241-260 (libstdc++ std::sort by trainIdx, first of every run survives)
void removeDuplicatedMatches(vector<cv::DMatch> &matches) {
  int n = (int)matches.size();
  mvo_remove_duplicated_matches(reinterpret_cast<mvo_dmatch *>(matches.data()), &n);
  matches.resize((size_t)n);
}

// ---- host helpers of the same translation unit (feature_match.cpp:262-300): no GPU work, same results ----
double computeMeanDistBetweenKeypoints(const vector<cv::KeyPoint> &kpts1, const vector<cv::KeyPoint> &kpts2,
                                       const vector<cv::DMatch> &matches) {
  double sum = 0;
  for (const cv::DMatch &d : matches) {
    const cv::Point2f p1 = kpts1[d.queryIdx].pt, p2 = kpts2[d.trainIdx].pt;
    const double dx = p1.x - p2.x, dy = p1.y - p2.y;
    sum += std::sqrt(dx * dx + dy * dy);
  }
  return sum / (double)matches.size();
}

vector<cv::DMatch> inliers2DMatches(const vector<int> inliers) {
  vector<cv::DMatch> matches;
  for (int idx : inliers) matches.push_back(cv::DMatch(idx, idx, 0.0));
  return matches;
}

vector<cv::KeyPoint> pts2Keypts(const vector<cv::Point2f> pts) {
  vector<cv::KeyPoint> keypts;
  for (const cv::Point2f &pt : pts) keypts.push_back(cv::KeyPoint(pt, 10));
  return keypts;
}

}  // namespace geometry
}  // namespace my_slam

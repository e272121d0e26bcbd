// TEST SHIM: the declarations of the reference's include/my_slam/geometry/feature_match.h:12-54 (the drop-in
// boundary) so that the adapter sources compile without the reference tree.
#pragma once
#include "my_slam/common_include.h"
namespace my_slam {
namespace geometry {
void calcKeyPoints(const cv::Mat &image, vector<cv::KeyPoint> &keypoints);
void calcDescriptors(const cv::Mat &image, vector<cv::KeyPoint> &keypoints, cv::Mat &descriptors);
void matchFeatures(const cv::Mat1b &descriptors_1, const cv::Mat1b &descriptors_2, vector<cv::DMatch> &matches, int method_index = 1,
                   bool is_print_res = false, const vector<cv::KeyPoint> &keypoints_1 = vector<cv::KeyPoint>(),
                   const vector<cv::KeyPoint> &keypoints_2 = vector<cv::KeyPoint>(), float max_matching_pixel_dist = 0.0);
vector<cv::DMatch> matchByRadiusAndBruteForce(const vector<cv::KeyPoint> &keypoints_1, const vector<cv::KeyPoint> &keypoints_2,
                                              const cv::Mat1b &descriptors_1, const cv::Mat1b &descriptors_2,
                                              float max_matching_pixel_dist);
void removeDuplicatedMatches(vector<cv::DMatch> &matches);
void selectUniformKptsByGrid(vector<cv::KeyPoint> &keypoints, int image_rows, int image_cols);
double computeMeanDistBetweenKeypoints(const vector<cv::KeyPoint> &kpts1, const vector<cv::KeyPoint> &kpts2,
                                       const vector<cv::DMatch> &matches);
vector<cv::DMatch> inliers2DMatches(const vector<int> inliers);
vector<cv::KeyPoint> pts2Keypts(const vector<cv::Point2f> pts);
}  // namespace geometry
}  // namespace my_slam

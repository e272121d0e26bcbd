// TEST SHIM: the declarations of the reference's include/my_slam/geometry/epipolar_geometry.h (:21-27, :32-39, :50-53,
// :58-63) that the adapter layer implements.
#pragma once
#include "my_slam/common_include.h"
namespace my_slam {
namespace geometry {
void estiMotionByEssential(const vector<cv::Point2f> &pts_in_img1, const vector<cv::Point2f> &pts_in_img2, const cv::Mat &camera_intrinsics,
                           cv::Mat &essential_matrix, cv::Mat &R, cv::Mat &t, vector<int> &inliers_index);
void estiMotionByHomography(const vector<cv::Point2f> &pts_in_img1, const vector<cv::Point2f> &pts_in_img2, const cv::Mat &camera_intrinsics,
                            cv::Mat &homography_matrix, vector<cv::Mat> &Rs, vector<cv::Mat> &ts, vector<cv::Mat> &normals,
                            vector<int> &inliers_index);
void removeWrongRtOfHomography(const vector<cv::Point2f> &pts_on_np1, const vector<cv::Point2f> &pts_on_np2, const vector<int> &inliers,
                               vector<cv::Mat> &Rs, vector<cv::Mat> &ts, vector<cv::Mat> &normals);
void doTriangulation(const vector<cv::Point2f> &pts_on_np1, const vector<cv::Point2f> &pts_on_np2, const cv::Mat &R_cam2_to_cam1,
                     const cv::Mat &t_cam2_to_cam1, const vector<int> &inliers, vector<cv::Point3f> &pts3d_in_cam1);
}  // namespace geometry
}  // namespace my_slam

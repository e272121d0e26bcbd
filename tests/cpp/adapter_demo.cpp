// Exercises the my_slam adapter layer the way the reference's own callers do:
//   test/test_epipolor_geometry.cpp:91-98   calcKeyPoints / calcDescriptors / matchFeatures on an image pair
//   src/vo/vo.cpp:293-337                   solvePnPRansac on 3d-2d pairs, inliers as a K x 1 CV_32SC1 index column
//   src/vo/vo.cpp:428-462                   bundleAdjustment through raw pointers into keypoints / map points / poses
// Inputs are raw little-endian arrays in <dir>, outputs are written next to them; tests/test_adapters.py compares
// them with the ctypes path and the oracles.  Built against tests/cvshim (no OpenCV C++ in this image).
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include "my_slam/geometry/feature_match.h"
#include "my_slam/optimization/g2o_ba.h"
#include "my_slam/basics/config.h"
#include "pnp_mvo.h"

template <class T> static std::vector<T> rd(const std::string &f) {
  std::ifstream in(f, std::ios::binary | std::ios::ate);
  if (!in) throw std::runtime_error("cannot open " + f);
  const size_t bytes = (size_t)in.tellg();
  in.seekg(0);
  std::vector<T> v(bytes / sizeof(T));
  in.read((char *)v.data(), (std::streamsize)(v.size() * sizeof(T)));
  return v;
}
template <class T> static void wr(const std::string &f, const T *p, size_t n) {
  std::ofstream out(f, std::ios::binary);
  out.write((const char *)p, (std::streamsize)(n * sizeof(T)));
}

int main(int argc, char **argv) {
  if (argc < 2) { fprintf(stderr, "usage: adapter_demo <dir>\n"); return 2; }
  const std::string d = std::string(argv[1]) + "/";
  try {
    using namespace my_slam;
    basics::Config::table()["max_number_of_keypoints"] = 2000;
    // ---- feature extraction and matching (test_epipolor_geometry.cpp:91-98) ----
    std::vector<unsigned char> b1 = rd<unsigned char>(d + "img1.bin"), b2 = rd<unsigned char>(d + "img2.bin");
    cv::Mat img1(480, 640, CV_8UC3, b1.data()), img2(480, 640, CV_8UC3, b2.data());
    vector<cv::KeyPoint> kp1, kp2;
    cv::Mat desc1, desc2;
    geometry::calcKeyPoints(img1, kp1);
    geometry::calcKeyPoints(img2, kp2);
    geometry::calcDescriptors(img1, kp1, desc1);
    geometry::calcDescriptors(img2, kp2, desc2);
    wr(d + "kp1.bin", kp1.data(), kp1.size());
    wr(d + "kp2.bin", kp2.data(), kp2.size());
    wr(d + "desc1.bin", desc1.data, (size_t)desc1.rows * 32);
    wr(d + "desc2.bin", desc2.data, (size_t)desc2.rows * 32);
    for (int method = 1; method <= 3; ++method) {
      vector<cv::DMatch> matches;
      geometry::matchFeatures(cv::Mat1b(desc1), cv::Mat1b(desc2), matches, method, method == 3, kp1, kp2, 50.f);
      wr(d + "matches" + std::to_string(method) + ".bin", matches.data(), matches.size());
      if (method == 3) printf("mean keypoint distance %.3f\n", geometry::computeMeanDistBetweenKeypoints(kp1, kp2, matches));
    }
    bool threw = false;
    try { vector<cv::DMatch> m; geometry::matchFeatures(cv::Mat1b(desc1), cv::Mat1b(desc2), m, 7); }
    catch (const std::runtime_error &) { threw = true; }                    // feature_match.cpp:225
    if (!threw) { fprintf(stderr, "matchFeatures(method 7) did not throw\n"); return 1; }

    // ---- PnP (vo.cpp:293-337) ----
    std::vector<float> p3 = rd<float>(d + "pnp_p3.bin"), p2 = rd<float>(d + "pnp_p2.bin");
    std::vector<double> Kv = rd<double>(d + "K.bin");
    cv::Mat K(3, 3, CV_64FC1);
    for (int i = 0; i < 9; ++i) K.at<double>(i / 3, i % 3) = Kv[(size_t)i];
    vector<cv::Point3f> pts_3d(p3.size() / 3);
    vector<cv::Point2f> pts_2d(p2.size() / 2);
    for (size_t i = 0; i < pts_3d.size(); ++i) { pts_3d[i] = cv::Point3f(p3[3 * i], p3[3 * i + 1], p3[3 * i + 2]); pts_2d[i] = cv::Point2f(p2[2 * i], p2[2 * i + 1]); }
    cv::Mat R_vec, t, inliers;
    const bool ok = mvo_adapter::solvePnPRansac(pts_3d, pts_2d, K, R_vec, t, 2.0f, inliers);
    double rt[6] = {R_vec.at<double>(0, 0), R_vec.at<double>(1, 0), R_vec.at<double>(2, 0), t.at<double>(0, 0), t.at<double>(1, 0), t.at<double>(2, 0)};
    wr(d + "pnp_rt.bin", rt, 6);
    wr(d + "pnp_inliers.bin", inliers.ptr<int>(0), (size_t)inliers.rows);
    printf("pnp ok %d inliers %d\n", (int)ok, inliers.rows);

    // ---- bundle adjustment through live pointers (vo.cpp:428-462) ----
    std::vector<double> poses = rd<double>(d + "ba_poses.bin");
    std::vector<float> bpts = rd<float>(d + "ba_points.bin"), bobs = rd<float>(d + "ba_obs.bin");
    std::vector<int32_t> bef = rd<int32_t>(d + "ba_edge_frame.bin"), bep = rd<int32_t>(d + "ba_edge_point.bin");
    const int F = (int)poses.size() / 16, P = (int)bpts.size() / 3, E = (int)bef.size();
    vector<cv::Mat> T((size_t)F);
    vector<cv::Mat *> Tp;
    for (int f = 0; f < F; ++f) { T[f].create(4, 4, CV_64FC1); for (int i = 0; i < 16; ++i) T[f].at<double>(i / 4, i % 4) = poses[(size_t)f * 16 + i]; Tp.push_back(&T[f]); }
    vector<cv::Point3f> map_pos((size_t)P);
    std::unordered_map<int, cv::Point3f *> um;
    for (int i = 0; i < P; ++i) { map_pos[i] = cv::Point3f(bpts[3 * i], bpts[3 * i + 1], bpts[3 * i + 2]); um[1000 + i] = &map_pos[i]; }   // ids need not be 0..P-1
    vector<cv::Point2f> kp_pt((size_t)E);
    vector<vector<cv::Point2f *>> v2d((size_t)F);
    vector<vector<int>> vidx((size_t)F);
    for (int e = 0; e < E; ++e) { kp_pt[e] = cv::Point2f(bobs[2 * e], bobs[2 * e + 1]); v2d[bef[e]].push_back(&kp_pt[e]); vidx[bef[e]].push_back(1000 + bep[e]); }
    cv::Mat info(2, 2, CV_64FC1);
    info.at<double>(0, 0) = 1; info.at<double>(1, 1) = 1;
    for (int fix = 1; fix >= 0; --fix) {
      vector<cv::Mat> Tc;
      for (auto &m : T) Tc.push_back(m.clone());
      vector<cv::Mat *> Tcp;
      for (auto &m : Tc) Tcp.push_back(&m);
      vector<cv::Point3f> pos = map_pos;
      std::unordered_map<int, cv::Point3f *> um2;
      for (int i = 0; i < P; ++i) um2[1000 + i] = &pos[i];
      optimization::bundleAdjustment(v2d, vidx, K, um2, Tcp, info, fix != 0, fix == 0);
      std::vector<double> out;
      for (auto &m : Tc) for (int i = 0; i < 16; ++i) out.push_back(m.at<double>(i / 4, i % 4));
      wr(d + (fix ? "ba_out_poses_fixed.bin" : "ba_out_poses_free.bin"), out.data(), out.size());
      wr(d + (fix ? "ba_out_points_fixed.bin" : "ba_out_points_free.bin"), &pos[0].x, (size_t)P * 3);
    }
    printf("adapter demo done\n");
  } catch (const std::exception &e) {
    fprintf(stderr, "adapter_demo: %s\n", e.what());
    return 1;
  }
  return 0;
}

// Drives my_slam::vo::VisualOdometry (my_slam_adapter/vo_mvo.h, reached through the forwarding header my_slam/vo/vo.h) the
// way the reference's run_vo.cpp main loop does (run_vo.cpp:118-137, 149-151): createFrame / addFrame per image, the pose
// history written with writePoseToFile's format, plus the members the display code reads (:184-232, :286-300).
// usage: adapter_vo_demo <dir> <n_frames>   — <dir>/frames.bin holds n_frames x 480 x 640 x 3 bytes (BGR);
// outputs: <dir>/traj.txt, <dir>/summary.txt (one line per frame).  Built against tests/cvshim (no OpenCV C++ here).
#include <cstdio>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "my_slam/basics/config.h"
#include "my_slam/vo/frame.h"
#include "my_slam/vo/vo.h"

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: adapter_vo_demo <dir> <n_frames>\n"); return 2; }
  const std::string d = std::string(argv[1]) + "/";
  const int n = atoi(argv[2]);
  try {
    using namespace my_slam;
    basics::Config::table()["max_number_of_keypoints"] = 2000;
    std::ifstream in(d + "frames.bin", std::ios::binary);
    if (!in) throw std::runtime_error("cannot open frames.bin");
    cv::Mat K(3, 3, CV_64FC1);
    K.at<double>(0, 0) = 615; K.at<double>(1, 1) = 615; K.at<double>(0, 2) = 320; K.at<double>(1, 2) = 240; K.at<double>(2, 2) = 1;
    geometry::Camera::Ptr camera(new geometry::Camera(K));
    vo::VisualOdometry::Ptr vo(new vo::VisualOdometry);
    std::vector<double> history;
    std::ofstream summary(d + "summary.txt");
    for (int img_id = 0; img_id < n; ++img_id) {
      cv::Mat rgb_img(480, 640, CV_8UC3);
      in.read((char *)rgb_img.data, 480 * 640 * 3);
      if (!in) throw std::runtime_error("frames.bin is too short");
      vo::Frame::Ptr frame = vo::Frame::createFrame(rgb_img, camera);
      vo->addFrame(frame);
      for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) history.push_back(frame->T_w_c_.at<double>(i, j));      // cam_pose_history (:137)
      summary << frame->id_ << " " << (vo->isInitialized() ? 1 : 0) << " " << (vo->getMap()->hasKeyFrame(frame->id_) ? 1 : 0) << " "
              << frame->keypoints_.size() << " " << frame->matches_with_ref_.size() << " " << frame->matches_with_map_.size() << " "
              << frame->inliers_pts3d_.size() << " " << vo->getMap()->map_points_.size() << " "
              << (vo->getPrevRef() ? vo->getPrevRef()->id_ : -1) << "\n";
      frame->clearNoUsed();
    }
    if (mvo_write_pose_file((d + "traj.txt").c_str(), history.data(), n) != MVO_OK) throw std::runtime_error("cannot write traj.txt");
  } catch (const std::exception &e) {
    fprintf(stderr, "adapter_vo_demo: %s\n", e.what());
    return 1;
  }
  return 0;
}

// my_slam::vo::{Frame, MapPoint, Map, VisualOdometry} over libmvo's mvo_vo_* — the public surface that the reference's
// run_vo.cpp uses (run_vo.cpp:118-137 the main loop, :184-232 drawResultByOpenCV, :236-300 drawResultByPcl), with the same
// class, member and method names as include/my_slam/vo/{frame,mappoint,map,vo}.h, so that run_vo.cpp compiles against
// these headers unchanged (put my_slam_adapter/include in front of the reference's include directory).
//
// The state machine itself runs inside libmvo (csrc/vo_pipeline.cpp).  After every addFrame the members run_vo.cpp reads
// are refreshed from it: the frame's keypoints_ / descriptors_ / matches_with_ref_ / matches_with_map_ / inliers_pts3d_ /
// T_w_c_, the poses of the buffered frames (bundle adjustment moves them, and the reference's Frame::Ptr aliases show it),
// Map::keyframes_ and Map::map_points_ (id, pos_, color_, descriptor_).  Members that only the reference's own
// implementation files touch (the per-frame connection graph, matched_times_ ...) stay inside libmvo.
#pragma once
#include <deque>
#include <memory>
#include <unordered_map>
#include <vector>
#include <opencv2/core.hpp>
#include "my_slam/geometry/camera.h"
#include "mvo.h"

namespace my_slam {
namespace vo {

class Frame {                                                     // include/my_slam/vo/frame.h:20-98
 public:
  typedef std::shared_ptr<Frame> Ptr;
  static int factory_id_;
  int id_ = 0;
  double time_stamp_ = -1;
  cv::Mat rgb_img_;
  std::vector<cv::KeyPoint> keypoints_;
  cv::Mat descriptors_;
  std::vector<cv::DMatch> matches_with_ref_;
  std::vector<cv::DMatch> matches_with_map_;
  std::vector<cv::Point3f> inliers_pts3d_;
  geometry::Camera::Ptr camera_;
  cv::Mat T_w_c_;

  static Frame::Ptr createFrame(cv::Mat rgb_img, geometry::Camera::Ptr camera, double time_stamp = -1);
  void clearNoUsed() {                                            // frame.h:61-69
    matches_with_ref_.clear();
    matches_with_map_.clear();
  }
  cv::Mat getCamCenter() const;                                   // frame.cpp:45-48
};

class MapPoint {                                                  // include/my_slam/vo/mappoint.h:15-37
 public:
  typedef std::shared_ptr<MapPoint> Ptr;
  int id_ = 0;
  cv::Point3f pos_;
  std::vector<unsigned char> color_;                              // r, g, b
  cv::Mat descriptor_;
};

class Map {                                                       // include/my_slam/vo/map.h:15-29
 public:
  typedef std::shared_ptr<Map> Ptr;
  std::unordered_map<int, Frame::Ptr> keyframes_;
  std::unordered_map<int, MapPoint::Ptr> map_points_;
  Frame::Ptr findKeyFrame(int frame_id) { auto it = keyframes_.find(frame_id); return it == keyframes_.end() ? nullptr : it->second; }
  bool hasKeyFrame(int frame_id) { return keyframes_.find(frame_id) != keyframes_.end(); }
};

class VisualOdometry {                                            // include/my_slam/vo/vo.h:24-118
 public:
  typedef std::shared_ptr<VisualOdometry> Ptr;
  VisualOdometry();
  ~VisualOdometry();
  VisualOdometry(const VisualOdometry &) = delete;
  VisualOdometry &operator=(const VisualOdometry &) = delete;

  void addFrame(Frame::Ptr frame);                                // vo_addFrame.cpp:10-142
  bool isInitialized();                                           // vo.cpp:174-177
  Frame::Ptr getPrevRef() { return prev_ref_; }
  Map::Ptr getMap() { return map_; }

 private:
  void create(const Frame &first);
  mvo_vo *vo_ = nullptr;
  int buffer_size_ = 20;                                          // kBuffSize_ (include/my_slam/vo/vo.h:77)
  Map::Ptr map_;
  Frame::Ptr prev_ref_;
  std::deque<Frame::Ptr> frames_buff_;                            // the same 20 newest frames libmvo keeps
  std::unordered_map<int, Frame::Ptr> by_id_;                     // libmvo frame id -> the caller's Frame
  std::unordered_map<int, int> lib_id_;                           // caller's Frame::id_ -> libmvo frame id
  std::unordered_map<int, Frame::Ptr> kf_by_lib_id_;              // keyframes by libmvo frame id (getPrevRef)
};

}  // namespace vo
}  // namespace my_slam

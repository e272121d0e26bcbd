#include "vo_mvo.h"
#include <cstring>
#include <sstream>
#include <stdexcept>
#include "mvo_context.h"
#include "my_slam/basics/config.h"

namespace my_slam {
namespace vo {

int Frame::factory_id_ = 0;

Frame::Ptr Frame::createFrame(cv::Mat rgb_img, geometry::Camera::Ptr camera, double time_stamp) {      // frame.cpp:12-20
  Frame::Ptr frame(new Frame());
  frame->rgb_img_ = rgb_img;
  frame->id_ = factory_id_++;
  frame->time_stamp_ = time_stamp;
  frame->camera_ = camera;
  return frame;
}

cv::Mat Frame::getCamCenter() const {
  cv::Mat c(3, 1, CV_64FC1);
  for (int i = 0; i < 3; ++i) c.at<double>(i, 0) = T_w_c_.at<double>(i, 3);
  return c;
}

VisualOdometry::VisualOdometry() : map_(new Map) {}
VisualOdometry::~VisualOdometry() { mvo_vo_destroy(vo_); }

// The reference latches its parameters from basics::Config inside the functions that use them (function-local statics in
// vo_addFrame.cpp:38-41,95-98 and vo.cpp:103,123-125,183-185,280-281,305,388-393); here they are read once, when the first
// frame fixes the image size.
void VisualOdometry::create(const Frame &first) {
  using basics::Config;
  mvo_vo_params p;
  mvo_vo_default_params(&p);
  p.match_method_init = (int)Config::get<float>("feature_match_method_index_initialization");
  p.max_match_dist_init = Config::get<float>("max_matching_pixel_dist_in_initialization");
  p.max_match_dist_triangulation = Config::get<float>("max_matching_pixel_dist_in_triangulation");
  p.essential_threshold = Config::get<double>("findEssentialMat_threshold");
  p.min_triang_angle = Config::get<double>("min_triang_angle");
  p.max_ratio_angle_to_median = Config::get<double>("max_ratio_between_max_angle_and_median_angle");
  p.min_inlier_matches = Config::get<int>("min_inlier_matches");
  p.min_pixel_dist = Config::get<double>("min_pixel_dist");
  p.min_median_triangulation_angle = Config::get<double>("min_median_triangulation_angle");
  p.assumed_mean_depth_init = Config::get<double>("assumed_mean_pts_depth_during_vo_init");
  p.track.match_method = (int)Config::get<float>("feature_match_method_index_pnp");
  p.track.match_radius = Config::get<float>("max_matching_pixel_dist_in_pnp");
  p.track.max_dist_to_prev = Config::get<double>("max_possible_dist_to_prev_keyframe");
  p.track.min_dist_keyframe = Config::get<double>("min_dist_between_two_keyframes");
  p.track.ba_enable = Config::getBool("is_enable_ba") ? 1 : 0;
  p.track.ba_window = Config::get<int>("num_prev_frames_to_opti_by_ba");
  p.track.ba_fix_points = Config::getBool("is_ba_fix_map_points") ? 1 : 0;
  std::istringstream im(Config::get<std::string>("information_matrix"));                   // vo.cpp:390-391
  for (int i = 0; i < 4; ++i) if (!(im >> p.track.information[i])) throw std::runtime_error("information_matrix: four numbers expected");
  const cv::Mat &K = first.camera_->K_;
  double Kf[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) Kf[i * 3 + j] = K.at<double>(i, j);
  buffer_size_ = p.track.buffer_size;
  mvo_adapter::check(mvo_vo_create(mvo_adapter::context(), Kf, first.rgb_img_.rows, first.rgb_img_.cols, &p, &vo_), "VisualOdometry");
}

bool VisualOdometry::isInitialized() { return vo_ && mvo_vo_is_initialized(vo_); }

namespace {
template <class T> std::vector<T> fetch(mvo_vo *vo, int which, int what) {
  static_assert(sizeof(cv::KeyPoint) == sizeof(mvo_keypoint) && sizeof(cv::DMatch) == sizeof(mvo_dmatch) && sizeof(cv::Point3f) == 12,
                "cv value types are layout-compatible with the C ABI structs");
  int n = 0;
  int rc = mvo_vo_frame_data(vo, which, what, nullptr, 0, &n);
  if (rc != MVO_OK && rc != MVO_ERR_CAPACITY) mvo_adapter::check(rc, "VisualOdometry::addFrame");
  std::vector<T> v((size_t)n);
  if (n > 0) mvo_adapter::check(mvo_vo_frame_data(vo, which, what, v.data(), n, &n), "VisualOdometry::addFrame");
  return v;
}
void set_pose(cv::Mat &T, const double *src) {
  if (T.empty()) T.create(4, 4, CV_64FC1);
  for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) T.at<double>(i, j) = src[i * 4 + j];
}
}  // namespace

void VisualOdometry::addFrame(Frame::Ptr frame) {
  if (!frame || frame->rgb_img_.empty() || !frame->camera_) throw std::runtime_error("VisualOdometry::addFrame: frame without image or camera");
  if (!vo_) create(*frame);
  const cv::Mat &img = frame->rgb_img_;
  double T[16];
  mvo_vo_frame_info info;
  mvo_adapter::check(mvo_vo_add_frame(vo_, img.data, img.channels(), img.step, T, &info), "VisualOdometry::addFrame");
  by_id_[info.frame_id] = frame;
  lib_id_[frame->id_] = info.frame_id;
  frames_buff_.push_back(frame);                                  // pushFrameToBuff_ (vo.h:81-86)
  if ((int)frames_buff_.size() > buffer_size_) {                  // kBuffSize_ (vo.h:77) = mvo_track_params::buffer_size
    auto it = lib_id_.find(frames_buff_.front()->id_);
    if (it != lib_id_.end()) { by_id_.erase(it->second); lib_id_.erase(it); }
    frames_buff_.pop_front();
  }

  // ---- the members run_vo.cpp reads from the frame ----
  frame->keypoints_ = fetch<cv::KeyPoint>(vo_, 0, MVO_VO_KEYPOINTS);
  const int nk = (int)frame->keypoints_.size();
  frame->descriptors_.create(nk, 32, CV_8UC1);
  if (nk > 0) {
    int n = 0;
    std::vector<unsigned char> d((size_t)nk * 32);
    mvo_adapter::check(mvo_vo_frame_data(vo_, 0, MVO_VO_DESCRIPTORS, d.data(), nk, &n), "VisualOdometry::addFrame");
    for (int r = 0; r < nk; ++r) std::memcpy(frame->descriptors_.ptr<unsigned char>(r), &d[(size_t)r * 32], 32);
  }
  frame->matches_with_ref_ = fetch<cv::DMatch>(vo_, 0, MVO_VO_MATCHES_WITH_REF);
  frame->matches_with_map_ = fetch<cv::DMatch>(vo_, 0, MVO_VO_MATCHES_WITH_MAP);
  frame->inliers_pts3d_ = fetch<cv::Point3f>(vo_, 0, MVO_VO_INLIERS_PTS3D);
  // poses of every buffered frame: bundle adjustment has moved the newest ones
  for (size_t k = 0; k < frames_buff_.size(); ++k) {
    double Tk[16];
    mvo_adapter::check(mvo_vo_frame_pose(vo_, (int)k, Tk), "VisualOdometry::addFrame");
    set_pose(frames_buff_[frames_buff_.size() - 1 - k]->T_w_c_, Tk);
  }
  // prev_ref_ = ref_ at the start of addFrame (vo_addFrame.cpp:26)
  {
    int32_t rid = -1;
    int n = 0;
    if (mvo_vo_frame_data(vo_, -1, MVO_VO_FRAME_ID, &rid, 1, &n) == MVO_OK) {
      auto it = by_id_.find(rid);
      auto kf = kf_by_lib_id_.find(rid);                          // a reference keyframe may be older than the 20-frame buffer
      prev_ref_ = it != by_id_.end() ? it->second : (kf != kf_by_lib_id_.end() ? kf->second : prev_ref_);
    } else {
      prev_ref_ = nullptr;
    }
  }
  // ---- the map ----
  if (info.keyframe) {
    map_->keyframes_[frame->id_] = frame;                         // Map::insertKeyFrame, keyed by the caller's frame id like the reference
    kf_by_lib_id_[info.frame_id] = frame;
  }
  const int nmap = mvo_vo_map_size(vo_);
  std::vector<int32_t> ids((size_t)(nmap > 0 ? nmap : 1));
  std::vector<float> pos((size_t)(nmap > 0 ? nmap : 1) * 3);
  std::vector<unsigned char> desc((size_t)(nmap > 0 ? nmap : 1) * 32), rgb((size_t)(nmap > 0 ? nmap : 1) * 3);
  int n = 0;
  mvo_adapter::check(mvo_vo_get_map(vo_, ids.data(), pos.data(), desc.data(), rgb.data(), (int)ids.size(), &n), "VisualOdometry::addFrame");
  std::unordered_map<int, MapPoint::Ptr> fresh;
  fresh.reserve((size_t)n);
  for (int i = 0; i < n; ++i) {
    auto old = map_->map_points_.find(ids[(size_t)i]);
    MapPoint::Ptr mp = old != map_->map_points_.end() ? old->second : MapPoint::Ptr(new MapPoint());
    if (old == map_->map_points_.end()) {
      mp->id_ = ids[(size_t)i];
      mp->color_.assign(&rgb[(size_t)i * 3], &rgb[(size_t)i * 3] + 3);
      mp->descriptor_.create(1, 32, CV_8UC1);
      std::memcpy(mp->descriptor_.data, &desc[(size_t)i * 32], 32);
    }
    mp->pos_ = cv::Point3f(pos[(size_t)i * 3], pos[(size_t)i * 3 + 1], pos[(size_t)i * 3 + 2]);
    fresh.emplace(mp->id_, mp);
  }
  map_->map_points_.swap(fresh);
}

}  // namespace vo
}  // namespace my_slam

// Per-frame tracking step over flat arrays: the DOING_TRACKING branch of
// vo::VisualOdometry::addFrame (reference src/vo/vo_addFrame.cpp:71-91) with its callees
// getMappointsInCurrentView_ (src/vo/vo.cpp:16-49), poseEstimationPnP_ (:267-381) and
// callBundleAdjustment_ (:384-478).  Host logic in C++ like the reference; every numeric stage
// goes through the C ABI of this library (no CPU fallback anywhere).
#include <deque>
#include <math.h>
#include <string.h>
#include <vector>
#include "mvo_internal.h"

struct TrackedFrame {
  double T_w_c[16];
  // inliers_to_mappt_connections_: keypoint pixel + map point index, in insertion order
  std::vector<float> obs_xy;
  std::vector<int32_t> map_idx;
};

struct mvo_tracker {
  mvo_ctx *ctx = nullptr;
  double K[9];
  int rows = 0, cols = 0;
  mvo_track_params prm;
  std::vector<float> map_pts;        // n x 3
  std::vector<uint8_t> map_desc;     // n x 32
  std::deque<TrackedFrame> frames;   // frames_buff_ (oldest first)
  double T_ref[16];                  // reference keyframe pose (initial guess for the next frame)
  bool has_prev = false;
  double T_prev[16];
  // scratch
  std::vector<mvo_keypoint> kpts;
  std::vector<uint8_t> desc;
  std::vector<uint8_t> cand_desc;
  std::vector<float> cand_xy, kp_xy, p3, p2;
  std::vector<int32_t> cand_idx, inliers;
  std::vector<mvo_dmatch> matches;
};

namespace {

void inv_rigid(const double *T, double *Ti) {     // [R t; 0 1]^-1 = [R^T, -R^T t]
  for (int i = 0; i < 3; ++i) {
    for (int j = 0; j < 3; ++j) Ti[i * 4 + j] = T[j * 4 + i];
    Ti[i * 4 + 3] = -(T[0 * 4 + i] * T[3] + T[1 * 4 + i] * T[7] + T[2 * 4 + i] * T[11]);
  }
  Ti[12] = Ti[13] = Ti[14] = 0;
  Ti[15] = 1;
}

void rvec_to_R(const double *w, double *R) {      // cv::Rodrigues
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
  double A, B;
  if (th < 1e-12) { A = 1; B = 0.5; }
  else { A = sin(th) / th; B = (1 - cos(th)) / th2; }
  const double x = w[0], y = w[1], z = w[2];
  R[0] = 1 - B * (y * y + z * z); R[1] = -A * z + B * x * y;      R[2] = A * y + B * x * z;
  R[3] = A * z + B * x * y;       R[4] = 1 - B * (x * x + z * z); R[5] = -A * x + B * y * z;
  R[6] = -A * y + B * x * z;      R[7] = A * x + B * y * z;       R[8] = 1 - B * (x * x + y * y);
}

double trans_dist(const double *Ta, const double *Tb) {
  const double dx = Ta[3] - Tb[3], dy = Ta[7] - Tb[7], dz = Ta[11] - Tb[11];
  return sqrt(dx * dx + dy * dy + dz * dz);
}

}  // namespace

extern "C" {

void mvo_default_track_params(mvo_track_params *p) {
  if (!p) return;
  memset(p, 0, sizeof *p);
  p->match_method = 1;            // config/config.yaml:75
  p->match_radius = 50.f;         // :91
  p->min_pnp_points = 5;          // src/vo/vo.cpp:304
  p->max_dist_to_prev = 0.3;      // config.yaml:117
  p->min_dist_keyframe = 0.03;    // :116
  p->ba_enable = 1;               // :120
  p->ba_window = 5;               // :121
  p->ba_fix_points = 1;           // :123
  p->information[0] = 1; p->information[1] = 0; p->information[2] = 0; p->information[3] = 1;   // :122
  p->buffer_size = 20;            // include/my_slam/vo/vo.h:77
}

int mvo_tracker_create(mvo_ctx *ctx, const double *K, int rows, int cols, const mvo_track_params *params,
                       mvo_tracker **out) {
  if (!ctx || !out) return MVO_ERR_INVALID_ARG;
  *out = nullptr;
  if (!K || rows <= 0 || cols <= 0) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "tracker: bad K / image size");
  mvo_tracker *t = new mvo_tracker();
  t->ctx = ctx;
  memcpy(t->K, K, sizeof t->K);
  t->rows = rows;
  t->cols = cols;
  if (params) t->prm = *params;
  else mvo_default_track_params(&t->prm);
  if (t->prm.match_method < 1 || t->prm.match_method > 3 || t->prm.ba_window < 1 || t->prm.buffer_size < 2) {
    delete t;
    return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "tracker: bad parameters");
  }
  memset(t->T_ref, 0, sizeof t->T_ref);
  t->T_ref[0] = t->T_ref[5] = t->T_ref[10] = t->T_ref[15] = 1;
  *out = t;
  return MVO_OK;
}

void mvo_tracker_destroy(mvo_tracker *t) { delete t; }

int mvo_tracker_set_map(mvo_tracker *t, const float *pts3d, const uint8_t *desc, int n) {
  if (!t) return MVO_ERR_INVALID_ARG;
  if (n < 0 || (n > 0 && (!pts3d || !desc))) return mvo_fail(t->ctx, MVO_ERR_INVALID_ARG, "tracker: null map");
  t->map_pts.assign(pts3d, pts3d + (size_t)n * 3);
  t->map_desc.assign(desc, desc + (size_t)n * 32);
  return MVO_OK;
}

int mvo_tracker_reset(mvo_tracker *t, const double *T_w_c_ref) {
  if (!t || !T_w_c_ref) return MVO_ERR_INVALID_ARG;
  memcpy(t->T_ref, T_w_c_ref, sizeof t->T_ref);
  t->frames.clear();
  t->has_prev = false;
  return MVO_OK;
}

int mvo_tracker_frame_pose(const mvo_tracker *t, int k, double *T_w_c) {
  if (!t || !T_w_c || k < 0 || k >= (int)t->frames.size()) return MVO_ERR_INVALID_ARG;
  memcpy(T_w_c, t->frames[t->frames.size() - 1 - k].T_w_c, 16 * sizeof(double));
  return MVO_OK;
}

int mvo_tracker_track(mvo_tracker *t, const uint8_t *image, int channels, size_t stride, int image_on_device,
                      double *T_w_c_out, mvo_track_result *res) {
  if (!t) return MVO_ERR_INVALID_ARG;
  mvo_ctx *ctx = t->ctx;
  if (!image || !T_w_c_out) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "tracker: null pointer");
  mvo_track_result r;
  memset(&r, 0, sizeof r);
  const int cap = ctx->prm.max_keypoints + 1;

  // pushFrameToBuff_ (vo.h:81-86) + Frame::calcKeyPoints / calcDescriptors
  t->frames.emplace_back();
  if ((int)t->frames.size() > t->prm.buffer_size) t->frames.pop_front();
  TrackedFrame &cur = t->frames.back();
  t->kpts.resize(cap);
  t->desc.resize((size_t)cap * 32);
  int nk = cap;
  int rc = mvo_orb_extract_ex(ctx, image, t->rows, t->cols, channels, stride, image_on_device, t->kpts.data(), &nk,
                              t->desc.data());
  if (rc != MVO_OK) { t->frames.pop_back(); return rc; }
  r.n_keypoints = nk;

  // curr_->T_w_c_ = ref_->T_w_c_.clone()  (vo_addFrame.cpp:74): initial guess = reference keyframe
  memcpy(cur.T_w_c, t->T_ref, sizeof cur.T_w_c);

  // ---- getMappointsInCurrentView_ (vo.cpp:16-49) ----
  double Tcw[16];
  inv_rigid(cur.T_w_c, Tcw);
  const int nmap = (int)(t->map_pts.size() / 3);
  t->cand_idx.clear(); t->cand_xy.clear(); t->cand_desc.clear();
  for (int i = 0; i < nmap; ++i) {
    // basics::preTranslatePoint3f: double accumulation of T(row, j) * p[j], result narrowed to float
    const double p[4] = {t->map_pts[3 * i], t->map_pts[3 * i + 1], t->map_pts[3 * i + 2], 1};
    double q[3] = {0, 0, 0};
    for (int row = 0; row < 3; ++row)
      for (int j = 0; j < 4; ++j) q[row] += Tcw[row * 4 + j] * p[j];
    const float cx = (float)q[0], cy = (float)q[1], cz = (float)q[2];
    bool in_frame = !(cz < 0);
    // geometry::cam2pixel (camera.cpp): K(0,0) * p.x / p.z + K(0,2) in double, narrowed to Point2f
    const float u = (float)(t->K[0] * cx / cz + t->K[2]), v = (float)(t->K[4] * cy / cz + t->K[5]);
    if (!(u > 0 && v > 0 && u < t->cols && v < t->rows)) in_frame = false;
    if (in_frame) {
      t->cand_idx.push_back(i);
      t->cand_xy.push_back(u);
      t->cand_xy.push_back(v);
      t->cand_desc.insert(t->cand_desc.end(), t->map_desc.begin() + (size_t)i * 32, t->map_desc.begin() + (size_t)i * 32 + 32);
    }
  }
  const int nc = (int)t->cand_idx.size();
  r.n_candidates = nc;

  // ---- matchFeatures(map descriptors, frame descriptors) (vo.cpp:283-289) ----
  t->kp_xy.resize((size_t)nk * 2);
  for (int i = 0; i < nk; ++i) { t->kp_xy[2 * i] = t->kpts[i].x; t->kp_xy[2 * i + 1] = t->kpts[i].y; }
  t->matches.resize(nc > 0 ? nc : 1);
  int nm = 0;
  if (nc > 0 && nk > 0 && !(t->prm.match_method == 2 && nk < 2)) {
    rc = mvo_match_features(ctx, t->cand_desc.data(), nc, t->desc.data(), nk, t->prm.match_method, t->cand_xy.data(),
                            t->kp_xy.data(), t->prm.match_radius, t->matches.data(), &nm);
    if (rc != MVO_OK) { t->frames.pop_back(); return rc; }
  }
  r.n_matches = nm;
  t->p3.resize((size_t)nm * 3);
  t->p2.resize((size_t)nm * 2);
  for (int i = 0; i < nm; ++i) {            // vo.cpp:293-301
    const int mi = t->cand_idx[t->matches[i].query_idx], ki = t->matches[i].train_idx;
    memcpy(&t->p3[3 * i], &t->map_pts[3 * mi], 12);
    t->p2[2 * i] = t->kpts[ki].x;
    t->p2[2 * i + 1] = t->kpts[ki].y;
  }

  // ---- poseEstimationPnP_ (vo.cpp:304-381) ----
  bool pnp_ok = nm >= t->prm.min_pnp_points;
  if (pnp_ok) {
    double rvec[3], tvec[3];
    t->inliers.resize(nm);
    int ni = nm;
    rc = mvo_solve_pnp_ransac(ctx, t->p3.data(), t->p2.data(), nm, t->K, rvec, tvec, t->inliers.data(), &ni);
    if (rc == MVO_ERR_DEGENERATE) { pnp_ok = false; ni = 0; }
    else if (rc != MVO_OK) { t->frames.pop_back(); return rc; }
    if (pnp_ok) {
      r.n_inliers = ni;
      for (int i = 0; i < ni; ++i) {        // vo.cpp:333-354
        const mvo_dmatch &m = t->matches[t->inliers[i]];
        cur.obs_xy.push_back(t->kpts[m.train_idx].x);
        cur.obs_xy.push_back(t->kpts[m.train_idx].y);
        cur.map_idx.push_back(t->cand_idx[m.query_idx]);
      }
      double Tc[16], R[9];
      rvec_to_R(rvec, R);
      for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) Tc[i * 4 + j] = R[i * 3 + j]; Tc[i * 4 + 3] = tvec[i]; }
      Tc[12] = Tc[13] = Tc[14] = 0; Tc[15] = 1;
      inv_rigid(Tc, cur.T_w_c);             // vo.cpp:357: T_w_c = [R|t]^-1
      // vo.cpp:360-369: reject jumps relative to the previous frame
      if (t->has_prev && trans_dist(cur.T_w_c, t->T_prev) >= t->prm.max_dist_to_prev) pnp_ok = false;
    }
  }
  if (!pnp_ok && t->has_prev) memcpy(cur.T_w_c, t->T_prev, sizeof cur.T_w_c);   // vo.cpp:376-379
  r.pnp_ok = pnp_ok;
  memcpy(r.T_w_c_pnp, cur.T_w_c, sizeof r.T_w_c_pnp);

  // ---- callBundleAdjustment_ (vo.cpp:384-478) ----
  if (pnp_ok && t->prm.ba_enable) {
    const int total = (int)t->frames.size();
    const int nba = std::min(t->prm.ba_window, total - 1);
    std::vector<double> poses;
    std::vector<int> which;
    std::vector<int32_t> ef, ep;
    std::vector<float> ob;
    for (int b = total - 1; b >= total - nba; --b) {          // newest first (vo.cpp:417-419)
      TrackedFrame &f = t->frames[b];
      if ((int)f.map_idx.size() < 3) continue;                // vo.cpp:423-426
      const int fi = (int)which.size();
      which.push_back(b);
      poses.insert(poses.end(), f.T_w_c, f.T_w_c + 16);
      for (size_t k = 0; k < f.map_idx.size(); ++k) {
        ef.push_back(fi);
        ep.push_back(f.map_idx[k]);
        ob.push_back(f.obs_xy[2 * k]);
        ob.push_back(f.obs_xy[2 * k + 1]);
      }
    }
    if (!which.empty()) {
      // only the map points that appear in the graph become vertices (um_pts_3d_in_prev_frames)
      std::vector<int32_t> remap(nmap, -1), used;
      for (int32_t &e : ep) {
        if (remap[e] < 0) { remap[e] = (int32_t)used.size(); used.push_back(e); }
        e = remap[e];
      }
      std::vector<float> pts((size_t)used.size() * 3);
      for (size_t k = 0; k < used.size(); ++k) memcpy(&pts[3 * k], &t->map_pts[3 * (size_t)used[k]], 12);
      const int fix = t->prm.ba_fix_points ? 1 : 0;
      rc = mvo_bundle_adjustment(ctx, poses.data(), (int)which.size(), pts.data(), (int)used.size(), ef.data(),
                                 ep.data(), ob.data(), (int)ef.size(), t->K, t->prm.information, fix, !fix, nullptr);
      if (rc != MVO_OK) return rc;
      for (size_t k = 0; k < which.size(); ++k) memcpy(t->frames[which[k]].T_w_c, &poses[16 * k], 16 * sizeof(double));
      if (!fix)
        for (size_t k = 0; k < used.size(); ++k) memcpy(&t->map_pts[3 * (size_t)used[k]], &pts[3 * k], 12);
      r.ba_frames = (int)which.size();
      r.ba_edges = (int)ef.size();
    }
  }
  // checkLargeMoveForAddKeyFrame_ (vo.cpp:247-265), translation part: the guess pose follows the camera
  if (pnp_ok && trans_dist(t->frames.back().T_w_c, t->T_ref) > t->prm.min_dist_keyframe)
    memcpy(t->T_ref, t->frames.back().T_w_c, sizeof t->T_ref);
  memcpy(t->T_prev, t->frames.back().T_w_c, sizeof t->T_prev);
  t->has_prev = true;
  memcpy(T_w_c_out, t->frames.back().T_w_c, 16 * sizeof(double));
  if (res) *res = r;
  return MVO_OK;
}

}  // extern "C"

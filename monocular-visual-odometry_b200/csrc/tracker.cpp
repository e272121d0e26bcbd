// Per-frame tracking step over flat arrays: the DOING_TRACKING branch of
// vo::VisualOdometry::addFrame (reference src/vo/vo_addFrame.cpp:71-91) with its callees
// getMappointsInCurrentView_ (src/vo/vo.cpp:16-49), poseEstimationPnP_ (:267-381) and
// callBundleAdjustment_ (:384-478).  Host logic in C++ like the reference; every numeric stage
// goes through the C ABI of this library (no CPU fallback anywhere).
#include <chrono>
#include <deque>
#include <stdlib.h>
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
  // BA assembly scratch (capacity reused across frames)
  std::vector<double> ba_poses;
  std::vector<int> ba_which;
  std::vector<int32_t> ba_ef, ba_ep, ba_used, ba_remap, ba_stamp;
  std::vector<float> ba_ob, ba_pts;
  int32_t ba_gen = 0;
  // extraction runs on two alternating contexts (own stream + workspace each) so that frame i+1 can be
  // extracted while frame i is being tracked: extraction does not depend on the VO state
  mvo_ctx *xctx[2] = {nullptr, nullptr};
  const uint8_t *pend_img[2] = {nullptr, nullptr};
  bool pend[2] = {false, false};
  unsigned n_submit = 0, n_consume = 0;
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
  for (int k = 0; k < 2; ++k) {
    const int rc = mvo_create(&t->xctx[k], ctx->device, &ctx->prm);
    if (rc != MVO_OK) {
      mvo_tracker_destroy(t);
      return mvo_fail(ctx, rc, "tracker: cannot create the extraction context");
    }
  }
  *out = t;
  return MVO_OK;
}

void mvo_tracker_destroy(mvo_tracker *t) {
  if (!t) return;
  for (int k = 0; k < 2; ++k)
    if (t->xctx[k]) mvo_destroy(t->xctx[k]);
  delete t;
}

int mvo_tracker_prefetch(mvo_tracker *t, const uint8_t *image, int channels, size_t stride, int image_on_device) {
  if (!t) return MVO_ERR_INVALID_ARG;
  mvo_ctx *ctx = t->ctx;
  if (!image) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "tracker: null image");
  const int slot = t->n_submit & 1;
  if (t->pend[slot]) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "tracker: two frames are already in flight");
  mvo_ctx *x = t->xctx[slot];
  MVO_TRY(mvo_set_params(x, &ctx->prm));          // follow parameter changes made on the main context
  const int rc = mvo_orb_extract_begin(x, image, t->rows, t->cols, channels, stride, image_on_device);
  if (rc != MVO_OK) return mvo_fail(ctx, rc, "tracker: %s", mvo_last_error(x));
  t->pend[slot] = true;
  t->pend_img[slot] = image;
  ++t->n_submit;
  return MVO_OK;
}

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
  // drop frames that were prefetched but never tracked
  for (int k = 0; k < 2; ++k)
    if (t->pend[k]) {
      const int cap = t->ctx->prm.max_keypoints + 1;
      t->kpts.resize(cap);
      t->desc.resize((size_t)cap * 32);
      int nk = cap;
      mvo_orb_extract_end(t->xctx[k], t->kpts.data(), &nk, t->desc.data(), nullptr);
      t->pend[k] = false;
    }
  t->n_submit = t->n_consume = 0;
  return MVO_OK;
}

int mvo_tracker_timing_enable(mvo_tracker *t, uint32_t mask) {
  if (!t) return MVO_ERR_INVALID_ARG;
  MVO_TRY(mvo_timing_enable(t->ctx, mask));
  for (int k = 0; k < 2; ++k) MVO_TRY(mvo_timing_enable(t->xctx[k], mask));
  return MVO_OK;
}

int mvo_tracker_timing_read(mvo_tracker *t, double *ms, uint64_t *counts) {
  if (!t) return MVO_ERR_INVALID_ARG;
  MVO_TRY(mvo_timing_read(t->ctx, ms, counts));
  for (int k = 0; k < 2; ++k) MVO_TRY(mvo_timing_read(t->xctx[k], ms, counts));
  return MVO_OK;
}

uint64_t mvo_tracker_kernel_launches(const mvo_tracker *t) {
  if (!t) return 0;
  return mvo_kernel_launches(t->ctx) + mvo_kernel_launches(t->xctx[0]) + mvo_kernel_launches(t->xctx[1]);
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
  static const bool dbg = getenv("MVO_TRACK_DEBUG") != nullptr;
  static double acc[8] = {0};
  static int nacc = 0;
  auto tnow = []() { return std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
  double t0 = dbg ? tnow() : 0;
#define TMARK(i) do { if (dbg) { const double t_ = tnow(); acc[i] += t_ - t0; t0 = t_; } } while (0)
  const int cap = ctx->prm.max_keypoints + 1;

  // pushFrameToBuff_ (vo.h:81-86) + Frame::calcKeyPoints / calcDescriptors
  t->frames.emplace_back();
  if ((int)t->frames.size() > t->prm.buffer_size) t->frames.pop_front();
  TrackedFrame &cur = t->frames.back();
  t->kpts.resize(cap);
  t->desc.resize((size_t)cap * 32);
  int nk = cap;
  // take the frame from the prefetch queue, or extract it now
  const int slot = t->n_consume & 1;
  if (t->pend[slot] && t->pend_img[slot] != image) { t->frames.pop_back(); return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "tracker: frames must be tracked in the order they were prefetched"); }
  if (!t->pend[slot]) {
    const int rc0 = mvo_tracker_prefetch(t, image, channels, stride, image_on_device);
    if (rc0 != MVO_OK) { t->frames.pop_back(); return rc0; }
  }
  mvo_ctx *x = t->xctx[slot];
  const uint8_t *d_desc = nullptr;
  t->pend[slot] = false;
  ++t->n_consume;
  int rc = mvo_orb_extract_end(x, t->kpts.data(), &nk, t->desc.data(), &d_desc);
  if (rc != MVO_OK) { t->frames.pop_back(); return mvo_fail(ctx, rc, "tracker: %s", mvo_last_error(x)); }
  r.n_keypoints = nk;
  TMARK(0);

  // curr_->T_w_c_ = ref_->T_w_c_.clone()  (vo_addFrame.cpp:74): initial guess = reference keyframe
  memcpy(cur.T_w_c, t->T_ref, sizeof cur.T_w_c);

  // ---- getMappointsInCurrentView_ (vo.cpp:16-49) ----
  double Tcw[16];
  inv_rigid(cur.T_w_c, Tcw);
  const int nmap = (int)(t->map_pts.size() / 3);
  t->cand_idx.resize(nmap); t->cand_xy.resize((size_t)nmap * 2); t->cand_desc.resize((size_t)nmap * 32);
  int ncand = 0;
  {
    const float *mp = t->map_pts.data();
    const uint8_t *md = t->map_desc.data();
    uint8_t *cd = t->cand_desc.data();
    const float fcols = (float)t->cols, frows = (float)t->rows;
    for (int i = 0; i < nmap; ++i) {
      // basics::preTranslatePoint3f: double accumulation of T(row, j) * p[j] (j = 0..3), narrowed to float
      const double p0 = mp[3 * i], p1 = mp[3 * i + 1], p2 = mp[3 * i + 2];
      double q[3];
      for (int row = 0; row < 3; ++row) {
        double acc = 0;
        acc += Tcw[row * 4] * p0; acc += Tcw[row * 4 + 1] * p1; acc += Tcw[row * 4 + 2] * p2; acc += Tcw[row * 4 + 3] * 1.0;
        q[row] = acc;
      }
      const float cx = (float)q[0], cy = (float)q[1], cz = (float)q[2];
      if (cz < 0) continue;
      // geometry::cam2pixel (camera.cpp): K(0,0) * p.x / p.z + K(0,2) in double, narrowed to Point2f
      const float u = (float)(t->K[0] * cx / cz + t->K[2]), v = (float)(t->K[4] * cy / cz + t->K[5]);
      if (!(u > 0 && v > 0 && u < fcols && v < frows)) continue;
      t->cand_idx[ncand] = i;
      t->cand_xy[2 * ncand] = u;
      t->cand_xy[2 * ncand + 1] = v;
      memcpy(cd + (size_t)ncand * 32, md + (size_t)i * 32, 32);
      ++ncand;
    }
  }
  t->cand_idx.resize(ncand);
  const int nc = (int)t->cand_idx.size();
  r.n_candidates = nc;
  TMARK(1);

  // ---- matchFeatures(map descriptors, frame descriptors) (vo.cpp:283-289) ----
  t->kp_xy.resize((size_t)nk * 2);
  for (int i = 0; i < nk; ++i) { t->kp_xy[2 * i] = t->kpts[i].x; t->kp_xy[2 * i + 1] = t->kpts[i].y; }
  t->matches.resize(nc > 0 ? nc : 1);
  int nm = 0;
  if (nc > 0 && nk > 0 && !(t->prm.match_method == 2 && nk < 2)) {
    rc = mvo_match_features_ex(ctx, t->cand_desc.data(), nc, d_desc, nk, 1, t->prm.match_method, t->cand_xy.data(),
                               t->kp_xy.data(), t->prm.match_radius, t->matches.data(), &nm);
    if (rc != MVO_OK) { t->frames.pop_back(); return rc; }
  }
  r.n_matches = nm;
  TMARK(2);
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
  TMARK(3);
  memcpy(r.T_w_c_pnp, cur.T_w_c, sizeof r.T_w_c_pnp);

  // ---- callBundleAdjustment_ (vo.cpp:384-478) ----
  if (pnp_ok && t->prm.ba_enable) {
    const int total = (int)t->frames.size();
    const int nba = std::min(t->prm.ba_window, total - 1);
    std::vector<double> &poses = t->ba_poses;
    std::vector<int> &which = t->ba_which;
    std::vector<int32_t> &ef = t->ba_ef, &ep = t->ba_ep, &used = t->ba_used;
    std::vector<float> &ob = t->ba_ob, &pts = t->ba_pts;
    poses.clear(); which.clear(); ef.clear(); ep.clear(); ob.clear(); used.clear();
    for (int b = total - 1; b >= total - nba; --b) {          // newest first (vo.cpp:417-419)
      TrackedFrame &f = t->frames[b];
      if ((int)f.map_idx.size() < 3) continue;                // vo.cpp:423-426
      const int fi = (int)which.size();
      which.push_back(b);
      poses.insert(poses.end(), f.T_w_c, f.T_w_c + 16);
      ef.insert(ef.end(), f.map_idx.size(), fi);
      ep.insert(ep.end(), f.map_idx.begin(), f.map_idx.end());
      ob.insert(ob.end(), f.obs_xy.begin(), f.obs_xy.end());
    }
    if (!which.empty()) {
      // only the map points that appear in the graph become vertices (um_pts_3d_in_prev_frames)
      if ((int)t->ba_remap.size() != nmap) { t->ba_remap.assign(nmap, 0); t->ba_stamp.assign(nmap, 0); t->ba_gen = 0; }
      const int32_t gen = ++t->ba_gen;
      for (int32_t &e : ep) {
        if (t->ba_stamp[e] != gen) { t->ba_stamp[e] = gen; t->ba_remap[e] = (int32_t)used.size(); used.push_back(e); }
        e = t->ba_remap[e];
      }
      pts.resize((size_t)used.size() * 3);
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
  TMARK(4);
  if (dbg && ++nacc % 50 == 0) {
    fprintf(stderr, "tracker us/frame: extract %.1f candidates %.1f match %.1f pnp %.1f ba %.1f\n", acc[0] / 50, acc[1] / 50, acc[2] / 50, acc[3] / 50, acc[4] / 50);
    for (double &a : acc) a = 0;
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

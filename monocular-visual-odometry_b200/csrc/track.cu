// Glue kernels of the device-resident tracking step (tracker.cpp): the pieces of
// vo::VisualOdometry::addFrame's tracking branch (reference src/vo/vo_addFrame.cpp:71-91) that the
// reference does with host loops over cv::Mat / std::vector, kept on the GPU so that the map, the frame
// buffer and the BA graph never cross PCIe:
//   k_project_map   getMappointsInCurrentView_ (src/vo/vo.cpp:16-49): project every map point with the
//                   guess pose, flag the ones in front of the camera and inside the image
//   k_kpt_xy        keypoint pt pairs for the radius-gated matcher (method 3)
//   k_gather_pairs  the 3d-2d pairs of poseEstimationPnP_ (vo.cpp:293-301) from the match list
//   k_track_glue    vo.cpp:333-379: inlier connections of the new frame (-> its slot of the frame buffer),
//                   T_w_c = [R|t]^-1 jump test against the previous frame, pose fallback, BA gate
// All tiny (<= 2001 elements): one or a few CTAs each; they exist to remove host round trips, not for FLOPs.
#include "mvo_internal.h"

namespace {

struct Rt12 { double v[12]; };   // R row-major (9) + t (3), world->camera

__global__ void __launch_bounds__(256)
k_project_map(const float *__restrict__ map_pts, int nmap, Rt12 Tcw, double fx, double fy, double cx, double cy,
              float fcols, float frows, uint8_t *__restrict__ vis, float2 *__restrict__ cxy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nmap) return;
  // basics::preTranslatePoint3f (opencv_funcs.cpp:67-78): double accumulation of T(row, j) * p[j], j = 0..3,
  // narrowed to float.  Explicit _rn intrinsics: no FMA contraction, so the visibility decision is the one the
  // host arithmetic of the reference takes.
  const double p0 = map_pts[3 * i], p1 = map_pts[3 * i + 1], p2 = map_pts[3 * i + 2];
  double q[3];
#pragma unroll
  for (int r = 0; r < 3; ++r) {
    double acc = __dmul_rn(Tcw.v[3 * r], p0);
    acc = __dadd_rn(acc, __dmul_rn(Tcw.v[3 * r + 1], p1));
    acc = __dadd_rn(acc, __dmul_rn(Tcw.v[3 * r + 2], p2));
    acc = __dadd_rn(acc, Tcw.v[9 + r]);
    q[r] = acc;
  }
  const float xc = (float)q[0], yc = (float)q[1], zc = (float)q[2];
  bool ok = !(zc < 0);
  // geometry::cam2pixel: K(0,0) * p.x / p.z + K(0,2) in double, narrowed to Point2f
  const float u = (float)__dadd_rn(__ddiv_rn(__dmul_rn(fx, (double)xc), (double)zc), cx);
  const float v = (float)__dadd_rn(__ddiv_rn(__dmul_rn(fy, (double)yc), (double)zc), cy);
  ok = ok && (u > 0 && v > 0 && u < fcols && v < frows);
  vis[i] = ok ? 1 : 0;
  cxy[i] = make_float2(u, v);
}

__global__ void __launch_bounds__(256)
k_kpt_xy(const mvo_keypoint *__restrict__ kpts, int n, float2 *__restrict__ xy) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) xy[i] = make_float2(kpts[i].x, kpts[i].y);
}

__global__ void __launch_bounds__(256)
k_gather_pairs(const int2 *__restrict__ pairs, int n, const float *__restrict__ map_pts,
               const mvo_keypoint *__restrict__ kpts, float *__restrict__ p3, float *__restrict__ p2) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int2 pr = pairs[i];
  p3[3 * i] = map_pts[3 * pr.x]; p3[3 * i + 1] = map_pts[3 * pr.x + 1]; p3[3 * i + 2] = map_pts[3 * pr.x + 2];
  p2[2 * i] = kpts[pr.y].x; p2[2 * i + 1] = kpts[pr.y].y;
}

struct GlueArgs {
  int mode;                 // 1: a PnP ran (pose_io / out_i / inl valid); 0: too few pairs, only the fallback applies
  int slot, cap, ba_enable, has_prev;
  double max_dist;          // max_possible_dist_to_prev_keyframe
  double prev_twc[3];       // translation of the previous frame's T_w_c
  Rt12 fallback;            // world->camera pose the frame keeps when PnP fails (vo.cpp:376-379)
  const double *pose_io;
  const int32_t *out_i, *inl;
  const int2 *pairs;
  const mvo_keypoint *kpts;
  int32_t *edge_map;        // [ring][cap]
  float2 *edge_obs;
  int32_t *cnt;             // [ring]
  double *pose;             // [ring][12]
  int32_t *skip_flag;
  int32_t *res_i;           // [0] model found, [1] pnp_ok, [2] consensus-set size
  double *res_d;            // [0..11] world->camera pose of the frame before BA
};

__global__ void __launch_bounds__(256) k_track_glue(GlueArgs a) {
  __shared__ int s_n;
  const int tid = threadIdx.x;
  if (tid == 0) {
    int n_in = 0, model = 0, ok = 0;
    double P[12];
    for (int q = 0; q < 12; ++q) P[q] = a.fallback.v[q];
    if (a.mode == 1) {
      n_in = a.out_i[0];
      model = n_in >= 4;
      if (model) {
        ok = 1;
        for (int q = 0; q < 12; ++q) P[q] = a.pose_io[q];
        // T_w_c = [R|t]^-1 (vo.cpp:357): translation -R^T t; reject jumps relative to the previous frame (:360-369)
        const double tx = -(P[0] * P[9] + P[3] * P[10] + P[6] * P[11]);
        const double ty = -(P[1] * P[9] + P[4] * P[10] + P[7] * P[11]);
        const double tz = -(P[2] * P[9] + P[5] * P[10] + P[8] * P[11]);
        const double dx = tx - a.prev_twc[0], dy = ty - a.prev_twc[1], dz = tz - a.prev_twc[2];
        if (a.has_prev && sqrt(dx * dx + dy * dy + dz * dz) >= a.max_dist) ok = 0;
        if (!ok) for (int q = 0; q < 12; ++q) P[q] = a.fallback.v[q];
      } else {
        n_in = 0;
      }
    }
    for (int q = 0; q < 12; ++q) { a.pose[(size_t)a.slot * 12 + q] = P[q]; a.res_d[q] = P[q]; }
    a.cnt[a.slot] = n_in;                       // the connections are recorded before the jump test (vo.cpp:333-354)
    a.skip_flag[0] = (ok && a.ba_enable) ? 0 : 1;
    a.res_i[0] = model; a.res_i[1] = ok; a.res_i[2] = n_in;
    s_n = n_in;
  }
  __syncthreads();
  const int n_in = min(s_n, a.cap);
  for (int j = tid; j < n_in; j += blockDim.x) {
    const int2 pr = a.pairs[a.inl[j]];
    a.edge_map[(size_t)a.slot * a.cap + j] = pr.x;
    a.edge_obs[(size_t)a.slot * a.cap + j] = make_float2(a.kpts[pr.y].x, a.kpts[pr.y].y);
  }
}

}  // namespace

// ---- launchers (asynchronous on ctx->stream) ------------------------------------------------------------
int mvo_track_project_map(mvo_ctx *ctx, const float *d_map_pts, int nmap, const double *Tcw12, const double *K,
                          int rows, int cols, uint8_t *d_vis, float *d_cxy) {
  if (nmap <= 0) return MVO_OK;
  Rt12 T;
  for (int q = 0; q < 12; ++q) T.v[q] = Tcw12[q];
  KTimer kt(ctx, KC_TRACK);
  k_project_map<<<(nmap + 255) / 256, 256, 0, ctx->stream>>>(d_map_pts, nmap, T, K[0], K[4], K[2], K[5], (float)cols, (float)rows,
                                                               d_vis, (float2 *)d_cxy);
  MVO_CHECK_LAUNCH(ctx);
  return MVO_OK;
}

int mvo_track_kpt_xy(mvo_ctx *ctx, const mvo_keypoint *d_kpts, int n, float *d_xy) {
  if (n <= 0) return MVO_OK;
  KTimer kt(ctx, KC_TRACK);
  k_kpt_xy<<<(n + 255) / 256, 256, 0, ctx->stream>>>(d_kpts, n, (float2 *)d_xy);
  MVO_CHECK_LAUNCH(ctx);
  return MVO_OK;
}

int mvo_track_gather_pairs(mvo_ctx *ctx, const int32_t *d_pairs, int n, const float *d_map_pts, const mvo_keypoint *d_kpts,
                           float *d_p3, float *d_p2) {
  if (n <= 0) return MVO_OK;
  KTimer kt(ctx, KC_TRACK);
  k_gather_pairs<<<(n + 255) / 256, 256, 0, ctx->stream>>>((const int2 *)d_pairs, n, d_map_pts, d_kpts, d_p3, d_p2);
  MVO_CHECK_LAUNCH(ctx);
  return MVO_OK;
}

int mvo_track_glue(mvo_ctx *ctx, const MvoTrackGlue &g) {
  GlueArgs a;
  a.mode = g.mode; a.slot = g.slot; a.cap = g.cap; a.ba_enable = g.ba_enable; a.has_prev = g.has_prev;
  a.max_dist = g.max_dist;
  for (int q = 0; q < 3; ++q) a.prev_twc[q] = g.prev_twc[q];
  for (int q = 0; q < 12; ++q) a.fallback.v[q] = g.fallback[q];
  a.pose_io = g.pose_io; a.out_i = g.out_i; a.inl = g.inl; a.pairs = (const int2 *)g.pairs; a.kpts = g.kpts;
  a.edge_map = g.edge_map; a.edge_obs = (float2 *)g.edge_obs; a.cnt = g.cnt; a.pose = g.pose;
  a.skip_flag = g.skip_flag; a.res_i = g.res_i; a.res_d = g.res_d;
  KTimer kt(ctx, KC_TRACK);
  k_track_glue<<<1, 256, 0, ctx->stream>>>(a);
  MVO_CHECK_LAUNCH(ctx);
  return MVO_OK;
}

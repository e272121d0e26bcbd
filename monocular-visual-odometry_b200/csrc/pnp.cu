// Batched RANSAC PnP for sm_100a — replaces the inline cv::solvePnPRansac call at reference
// src/vo/vo.cpp:318-320 (iterationsCount 100, reprojectionError 2.0, confidence 0.999,
// useExtrinsicGuess false, default SOLVEPNP_ITERATIVE).  Structure kept from OpenCV (SURVEY.md
// App. C): hypotheses from minimal sets -> consensus set of the best hypothesis (returned as
// ascending indices) -> iterative least-squares refit on exactly those inliers.  What differs by
// design: all H hypotheses (default 4096, not <=100 adaptive) are generated and scored in one
// batch; the minimal solver is a P3P (3 points + 1 to disambiguate) instead of 5-point EPnP.
//
//   k_pnp_hypotheses  one thread per hypothesis: counter-based sampling (splitmix64), P3P by
//                     intersecting the two distance-ratio conics through their degenerate
//                     pencil member, pose from the three depths, 4th point picks the root.
//   k_pnp_score       persistent grid, one warp per hypothesis at a time, the N correspondences
//                     staged once per CTA in shared memory; fp64 reprojection, count(err^2<=thr^2).
//   k_pnp_finish      one CTA: arg-max (ties -> lowest hypothesis), ordered inlier compaction,
//                     damped Gauss-Newton on SE(3) over the inliers (fp64).
// All arithmetic fp64: the stage is latency-bound (N*H = 8.2e6 reprojections), not FLOP-bound.
#include <string.h>
#include "mvo_internal.h"

namespace {

#define PNP_DYN_SMEM(type, name) extern __shared__ type name[]
#include "launch_pdl.cuh"
#include "pnp_kernels.cuh"

}  // namespace

// ---------------------------------------------------------------------------- host entry
static void rotation_to_rvec(const double *R, double *rvec) {
  // via the unit quaternion: robust for every angle in [0, pi]
  double q[4];
  const double tr = R[0] + R[4] + R[8];
  if (tr > 0) { double s = sqrt(tr + 1.0) * 2; q[0] = 0.25 * s; q[1] = (R[7] - R[5]) / s; q[2] = (R[2] - R[6]) / s; q[3] = (R[3] - R[1]) / s; }
  else if (R[0] > R[4] && R[0] > R[8]) { double s = sqrt(1.0 + R[0] - R[4] - R[8]) * 2; q[0] = (R[7] - R[5]) / s; q[1] = 0.25 * s; q[2] = (R[1] + R[3]) / s; q[3] = (R[2] + R[6]) / s; }
  else if (R[4] > R[8]) { double s = sqrt(1.0 + R[4] - R[0] - R[8]) * 2; q[0] = (R[2] - R[6]) / s; q[1] = (R[1] + R[3]) / s; q[2] = 0.25 * s; q[3] = (R[5] + R[7]) / s; }
  else { double s = sqrt(1.0 + R[8] - R[0] - R[4]) * 2; q[0] = (R[3] - R[1]) / s; q[1] = (R[2] + R[6]) / s; q[2] = (R[5] + R[7]) / s; q[3] = 0.25 * s; }
  if (q[0] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
  const double vn = sqrt(q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  if (vn < 1e-15) { rvec[0] = rvec[1] = rvec[2] = 0; return; }
  const double th = 2 * atan2(vn, q[0]);
  for (int i = 0; i < 3; ++i) rvec[i] = th * q[1 + i] / vn;
}

static void rvec_to_rotation(const double *w, double *R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
  double A, B;
  if (th < 1e-8) { A = 1 - th2 / 6; B = 0.5 - th2 / 24; }
  else { A = sin(th) / th; B = (1 - cos(th)) / th2; }
  const double wx = w[0], wy = w[1], wz = w[2];
  R[0] = 1 - B * (wy * wy + wz * wz); R[1] = -A * wz + B * wx * wy;       R[2] = A * wy + B * wx * wz;
  R[3] = A * wz + B * wx * wy;        R[4] = 1 - B * (wx * wx + wz * wz); R[5] = -A * wx + B * wy * wz;
  R[6] = -A * wy + B * wx * wz;       R[7] = A * wx + B * wy * wz;        R[8] = 1 - B * (wx * wx + wy * wy);
}

static int pnp_hyp_count(const mvo_ctx *ctx);
struct PnpWs { float *p3, *p2; double *poses, *pose_io, *ex, *eo, *stats; int32_t *valid, *counts, *out_i, *inl, *ef; };

// ba.cu
int mvo_ba_pose_launch(mvo_ctx *ctx, int F, int E, int chunk, const int32_t *d_eframe, const double *d_X, const double *d_obs, double fx, double fy,
                       double cx, double cy, const double *info, int iters, int use_huber, double huber, int fix_first,
                       double step_tol, double *d_poses, double *d_stats);

static int pnp_ws(mvo_ctx *ctx, int n, int H, PnpWs *w) {
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_p3 = 0, o_p2 = al(o_p3 + (size_t)n * 12), o_inl = al(o_p2 + (size_t)n * 8), o_ex = al(o_inl + (size_t)n * 4);
  const size_t o_eo = al(o_ex + (size_t)n * 24), o_ef = al(o_eo + (size_t)n * 16), o_end = al(o_ef + (size_t)n * 4);
  MVO_TRY(mvo_reserve(ctx, ctx->pnp_pts, o_end));
  MVO_TRY(mvo_reserve(ctx, ctx->pnp_hyp, (size_t)H * 12 * 8 + 256));
  MVO_TRY(mvo_reserve(ctx, ctx->pnp_cnt, (size_t)H * 8 + 512));
  MVO_TRY(mvo_reserve(ctx, ctx->pnp_out, 1024));
  uint8_t *b = (uint8_t *)ctx->pnp_pts.p;
  w->p3 = (float *)(b + o_p3); w->p2 = (float *)(b + o_p2); w->inl = (int32_t *)(b + o_inl);
  w->ex = (double *)(b + o_ex); w->eo = (double *)(b + o_eo); w->ef = (int32_t *)(b + o_ef);
  w->poses = (double *)ctx->pnp_hyp.p;
  w->valid = (int32_t *)ctx->pnp_cnt.p;
  w->counts = w->valid + ((H + 63) & ~63);
  w->pose_io = (double *)ctx->pnp_out.p;
  w->out_i = (int32_t *)((uint8_t *)ctx->pnp_out.p + 256);
  w->stats = (double *)((uint8_t *)ctx->pnp_out.p + 512);
  return MVO_OK;
}

static int pnp_cam(mvo_ctx *ctx, const double *K, PnpCam *cam) {
  if (!K) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "null K");
  cam->fx = K[0]; cam->fy = K[4]; cam->cx = K[2]; cam->cy = K[5];
  if (!(cam->fx > 0 && cam->fy > 0)) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "K: focal lengths must be positive");
  return MVO_OK;
}

// Least-squares refit of the pose on the one-frame edge list written by k_pnp_finish: the same
// pose-only LM as the fixed-points BA (no robust kernel, identity information, fx/fy), stopping
// once an accepted step is below 1e-8 (quadratic convergence: the next step would be ~1e-16).  E = n is an upper bound: unused edge slots carry frame -1.
static int pnp_refit(mvo_ctx *ctx, const PnpWs &w, int n, const PnpCam &cam) {
  static const double I2[4] = {1, 0, 0, 1};
  return mvo_ba_pose_launch(ctx, 1, n, 1, w.ef, w.ex, w.eo, cam.fx, cam.fy, cam.cx, cam.cy, I2, ctx->prm.pnp_refine_iters, 0, 1.0, 0,
                            1e-8, w.pose_io, w.stats);
}

// hypotheses -> scores -> consensus set of the best hypothesis -> least-squares refit, all on ctx->stream;
// w.p3 / w.p2 hold the n correspondences on the device
static int pnp_hyp_count(const mvo_ctx *ctx) { return ctx->prm.pnp_mode == 1 ? 100 : ctx->prm.pnp_hypotheses; }   // vo.cpp:315: iterationsCount = 100

// pnp_mode = 1: cv::solvePnPRansac's own flow (pnp_cv_kernels.cuh) — 100 EPnP iterations evaluated in parallel, the adaptive loop replayed
static int pnp_enqueue_cv(mvo_ctx *ctx, const PnpWs &w, int n, const PnpCam &cam, const int32_t *n_dev) {
  const int H = pnp_hyp_count(ctx);
  const double thr2 = (double)ctx->prm.pnp_reproj_error * (double)ctx->prm.pnp_reproj_error;
  { KTimer kt(ctx, KC_PNP_HYP);
  k_pnp_epnp<<<(H + CVP_WARPS - 1) / CVP_WARPS, CVP_WARPS * 32, 0, ctx->stream>>>(w.p3, w.p2, n, n_dev, cam, H, w.poses, w.valid); }
  MVO_CHECK_LAUNCH(ctx);
  { KTimer kt(ctx, KC_PNP_SCORE);
  k_pnp_score_cv<<<(H + 7) / 8, 256, 0, ctx->stream>>>(w.p3, w.p2, n, n_dev, cam, (float)thr2, H, w.poses, w.valid, w.counts); }
  MVO_CHECK_LAUNCH(ctx);
  { KTimer kt(ctx, KC_PNP_FINISH);
  k_pnp_finish<<<1, FIN_T, 0, ctx->stream>>>(w.p3, w.p2, n, n_dev, cam, thr2, H, w.poses, w.counts, 2, ctx->prm.pnp_refine_iters,
                                             w.pose_io, w.out_i, w.inl, w.ex, w.eo, w.ef); }
  MVO_CHECK_LAUNCH(ctx);
  ctx->pnp_last_h = H;
  return pnp_refit(ctx, w, n, cam);
}

static int pnp_enqueue(mvo_ctx *ctx, const PnpWs &w, int n, const PnpCam &cam, const int32_t *n_dev = nullptr) {
  if (ctx->prm.pnp_mode == 1) return pnp_enqueue_cv(ctx, w, n, cam, n_dev);
  const int H = ctx->prm.pnp_hypotheses;
  const size_t smem = (size_t)n * 5 * sizeof(float);
  if (smem > 200 * 1024) return mvo_fail(ctx, MVO_ERR_UNSUPPORTED, "solvePnPRansac: more than %d correspondences", 200 * 1024 / 20);
  const double thr2 = (double)ctx->prm.pnp_reproj_error * (double)ctx->prm.pnp_reproj_error;
  { KTimer kt(ctx, KC_PNP_HYP);
  MVO_CUDA(ctx, launch_pdl(ctx->stream, (unsigned)((H + 127) / 128), 128, 0, 1, k_pnp_hypotheses, (const float *)w.p3, (const float *)w.p2, n, n_dev, cam,
                           (uint64_t)ctx->prm.pnp_seed, H, w.poses, w.valid)); }
  ctx->launches++;
  if (smem > 48 * 1024) MVO_CUDA(ctx, cudaFuncSetAttribute(k_pnp_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (H + 7) / 8;                                   // one hypothesis per warp while the CTAs of one wave allow it
  if (grid > 4 * ctx->sm_count) grid = 4 * ctx->sm_count;
  { KTimer kt(ctx, KC_PNP_SCORE);
  MVO_CUDA(ctx, launch_pdl(ctx->stream, (unsigned)grid, 256, smem, 1, k_pnp_score, (const float *)w.p3, (const float *)w.p2, n, n_dev, cam, thr2, H,
                           (const double *)w.poses, (const int32_t *)w.valid, w.counts)); }
  ctx->launches++;
  { KTimer kt(ctx, KC_PNP_FINISH);
  MVO_CUDA(ctx, launch_pdl(ctx->stream, 1, FIN_T, 0, 1, k_pnp_finish, (const float *)w.p3, (const float *)w.p2, n, n_dev, cam, thr2, H,
                           (const double *)w.poses, (const int32_t *)w.counts, 0, (int)ctx->prm.pnp_refine_iters, w.pose_io, w.out_i, w.inl, w.ex, w.eo, w.ef)); }
  ctx->launches++;
  ctx->pnp_last_h = H;
  // the refit runs on the consensus set whose size only the device knows: launch it for the
  // worst case E = n with the real count read on the device (edges beyond n_in are masked out)
  return pnp_refit(ctx, w, n, cam);
}

// Device-resident entry points for the tracker (tracker.cpp): reserve the workspace for n correspondences
// (the caller gathers them into *p3 / *p2 on ctx->stream), then enqueue the whole solvePnPRansac replacement.
// Results stay on the device: pose_io = [R|t] world->camera after the refit, out_i[0] = consensus-set size
// (< 4: no model), inl = its ascending indices.
int mvo_pnp_dev_buffers(mvo_ctx *ctx, int n, float **p3, float **p2, double **pose_io, int32_t **out_i, int32_t **inl) {
  PnpWs w;
  MVO_TRY(pnp_ws(ctx, n, pnp_hyp_count(ctx), &w));
  *p3 = w.p3; *p2 = w.p2; *pose_io = w.pose_io; *out_i = w.out_i; *inl = w.inl;
  return MVO_OK;
}

// n = number of correspondences, or their upper bound when d_n (device counter) is given
int mvo_pnp_dev_run(mvo_ctx *ctx, int n, const double *K, const int32_t *d_n) {
  if (n < 4) return mvo_fail(ctx, MVO_ERR_DEGENERATE, "solvePnPRansac: %d correspondences (< 4)", n);
  PnpCam cam;
  MVO_TRY(pnp_cam(ctx, K, &cam));
  PnpWs w;
  MVO_TRY(pnp_ws(ctx, n, pnp_hyp_count(ctx), &w));
  return pnp_enqueue(ctx, w, n, cam, d_n);
}

extern "C" {

int mvo_solve_pnp_ransac(mvo_ctx *ctx, const float *pts3d, const float *pts2d, int n, const double *K,
                         double *rvec, double *tvec, int32_t *inliers, int *n_inliers) {
  if (!ctx) return MVO_ERR_INVALID_ARG;
  if (!pts3d || !pts2d || !rvec || !tvec || !n_inliers || (*n_inliers > 0 && !inliers))
    return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "solvePnPRansac: null pointer");
  if (n < 4) return mvo_fail(ctx, MVO_ERR_DEGENERATE, "solvePnPRansac: %d correspondences (< 4)", n);
  PnpCam cam;
  MVO_TRY(pnp_cam(ctx, K, &cam));
  MVO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int H = pnp_hyp_count(ctx);
  PnpWs w;
  MVO_TRY(pnp_ws(ctx, n, H, &w));
  if ((size_t)n * 20 > 200 * 1024) return mvo_fail(ctx, MVO_ERR_UNSUPPORTED, "solvePnPRansac: more than %d correspondences", 200 * 1024 / 20);
  const size_t hb = (size_t)n * 20 + (size_t)n * 4 + 1024;
  MVO_TRY(mvo_reserve_pinned(ctx, ctx->h_b, hb));
  float *h3 = (float *)ctx->h_b.p, *h2 = h3 + (size_t)n * 3;
  memcpy(h3, pts3d, (size_t)n * 12);
  memcpy(h2, pts2d, (size_t)n * 8);
  MVO_CUDA(ctx, cudaMemcpyAsync(w.p3, h3, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
  MVO_CUDA(ctx, cudaMemcpyAsync(w.p2, h2, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  MVO_TRY(pnp_enqueue(ctx, w, n, cam));
  uint8_t *hout = (uint8_t *)ctx->h_b.p + (size_t)n * 20;
  double *h_pose = (double *)hout;
  int32_t *h_i = (int32_t *)(hout + 128), *h_inl = (int32_t *)(hout + 256);
  MVO_CUDA(ctx, cudaMemcpyAsync(h_pose, w.pose_io, 96, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaMemcpyAsync(h_i, w.out_i, 16, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaMemcpyAsync(h_inl, w.inl, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int ni = h_i[0];
  if (ni < 4) { *n_inliers = 0; return mvo_fail(ctx, MVO_ERR_DEGENERATE, "solvePnPRansac: no hypothesis reached 4 inliers"); }
  if (ni > *n_inliers) return mvo_fail(ctx, MVO_ERR_CAPACITY, "inlier capacity %d < %d", *n_inliers, ni);
  rotation_to_rvec(h_pose, rvec);
  tvec[0] = h_pose[9]; tvec[1] = h_pose[10]; tvec[2] = h_pose[11];
  memcpy(inliers, h_inl, (size_t)ni * 4);
  *n_inliers = ni;
  return MVO_OK;
}

int mvo_pnp_last_hypotheses(mvo_ctx *ctx, double *poses, int32_t *counts, int cap, int *n_hyp) {
  if (!ctx) return MVO_ERR_INVALID_ARG;
  if (!poses || !counts || !n_hyp) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "null pointer");
  const int H = ctx->pnp_last_h;
  if (H <= 0) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "no PnP call yet");
  if (cap < H) return mvo_fail(ctx, MVO_ERR_CAPACITY, "capacity %d < %d", cap, H);
  MVO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int32_t *d_counts = (int32_t *)ctx->pnp_cnt.p + ((H + 63) & ~63);
  MVO_CUDA(ctx, cudaMemcpyAsync(poses, ctx->pnp_hyp.p, (size_t)H * 96, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaMemcpyAsync(counts, d_counts, (size_t)H * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  *n_hyp = H;
  return MVO_OK;
}

int mvo_pnp_refine(mvo_ctx *ctx, const float *pts3d, const float *pts2d, int n, const double *K, double *rvec,
                   double *tvec) {
  if (!ctx) return MVO_ERR_INVALID_ARG;
  if (!pts3d || !pts2d || !rvec || !tvec) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "pnp_refine: null pointer");
  if (n < 3) return mvo_fail(ctx, MVO_ERR_DEGENERATE, "pnp_refine: %d correspondences (< 3)", n);
  PnpCam cam;
  MVO_TRY(pnp_cam(ctx, K, &cam));
  MVO_CUDA(ctx, cudaSetDevice(ctx->device));
  PnpWs w;
  MVO_TRY(pnp_ws(ctx, n, pnp_hyp_count(ctx), &w));
  MVO_TRY(mvo_reserve_pinned(ctx, ctx->h_b, (size_t)n * 20 + 1024));
  float *h3 = (float *)ctx->h_b.p, *h2 = h3 + (size_t)n * 3;
  double *h_pose = (double *)((uint8_t *)ctx->h_b.p + (size_t)n * 20);
  memcpy(h3, pts3d, (size_t)n * 12);
  memcpy(h2, pts2d, (size_t)n * 8);
  rvec_to_rotation(rvec, h_pose);
  h_pose[9] = tvec[0]; h_pose[10] = tvec[1]; h_pose[11] = tvec[2];
  MVO_CUDA(ctx, cudaMemcpyAsync(w.p3, h3, (size_t)n * 12, cudaMemcpyHostToDevice, ctx->stream));
  MVO_CUDA(ctx, cudaMemcpyAsync(w.p2, h2, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  MVO_CUDA(ctx, cudaMemcpyAsync(w.pose_io, h_pose, 96, cudaMemcpyHostToDevice, ctx->stream));
  { KTimer kt(ctx, KC_PNP_FINISH);
  k_pnp_finish<<<1, FIN_T, 0, ctx->stream>>>(w.p3, w.p2, n, nullptr, cam, 0.0, 0, nullptr, nullptr, 1, ctx->prm.pnp_refine_iters,
                                             w.pose_io, w.out_i, w.inl, w.ex, w.eo, w.ef); }
  MVO_CHECK_LAUNCH(ctx);
  MVO_TRY(pnp_refit(ctx, w, n, cam));
  MVO_CUDA(ctx, cudaMemcpyAsync(h_pose, w.pose_io, 96, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  rotation_to_rvec(h_pose, rvec);
  tvec[0] = h_pose[9]; tvec[1] = h_pose[10]; tvec[2] = h_pose[11];
  return MVO_OK;
}

}  // extern "C"

// Two-view geometry for sm_100a (first part) — replaces
//   geometry::estiMotionByEssential  reference src/geometry/epipolar_geometry.cpp:17-57:
//       cv::findEssentialMat(pts1, pts2, focal, pp, RANSAC, prob 0.999, threshold 1.0, mask)  +  E /= E(2,2)
//       + inliers from the mask  +  cv::recoverPose(E, pts1, pts2, R, t, focal, pp, mask)  +  t /= |t|
//   geometry::doTriangulation        :130-175:  cv::triangulatePoints([I|0], [R|t], inlier points on the normalised
//       plane) and the division by the fourth coordinate
// Structure kept from OpenCV: hypotheses from minimal samples -> Sampson-distance consensus -> decomposeEssentialMat ->
// the cheirality vote of recoverPose over its four (R, t) candidates (triangulation of every inlier, depth in (0, 50)
// in both cameras, OpenCV's tie order).  What differs by design, as for PnP: all H hypotheses
// (mvo_params::epi_hypotheses, default 4096 — not the adaptive <= ~1000 of prob 0.999) are generated and scored in one
// batch; the minimal solver is the linear eight-point method projected onto the essential manifold instead of
// Nister's five-point solver; and the best minimal model is locally optimised on its consensus set (k_epi_finish).
//   k_epi_hypotheses  one thread per hypothesis: counter-based sampling, epi::essential_from_8
//   k_epi_score       one warp per hypothesis, points staged once per CTA in shared memory, count(Sampson <= thr^2)
//   k_epi_finish      one CTA: arg-max (ties -> lowest hypothesis), ordered inlier list, recoverPose vote
//   k_triangulate     one thread per point: epi::triangulate_dlt
// The numerics live in epipolar_math.cuh and are unit-tested on the host (tests/test_epipolar_math.py).
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "epipolar_math.cuh"
#include "mvo_internal.h"

namespace {

__device__ __forceinline__ uint64_t splitmix64e(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct EpiCam { double f, cx, cy; };      // findEssentialMat(focal, pp): one focal length (epipolar_geometry.cpp:26-27)

__global__ void __launch_bounds__(128)
k_epi_hypotheses(const float *__restrict__ p1, const float *__restrict__ p2, int n, EpiCam cam, uint64_t seed, int H,
                 double *__restrict__ Es, int32_t *__restrict__ valid) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  int idx[8];
  uint64_t ctr = 0;
  for (int k = 0; k < 8; ++k)
    for (int attempt = 0; attempt < 64; ++attempt) {
      const uint64_t r = splitmix64e(seed ^ splitmix64e(((uint64_t)h << 20) ^ ctr++));
      const int cand = (int)(r % (uint64_t)n);
      bool dup = false;
      for (int q = 0; q < k; ++q) dup |= idx[q] == cand;
      idx[k] = cand;
      if (!dup) break;
    }
  double a[16], b[16];
  const double inv = 1.0 / cam.f;
  for (int k = 0; k < 8; ++k) {
    a[2 * k] = ((double)p1[2 * idx[k]] - cam.cx) * inv; a[2 * k + 1] = ((double)p1[2 * idx[k] + 1] - cam.cy) * inv;
    b[2 * k] = ((double)p2[2 * idx[k]] - cam.cx) * inv; b[2 * k + 1] = ((double)p2[2 * idx[k] + 1] - cam.cy) * inv;
  }
  double E[9];
  int reason = 0;
  const bool ok = epi::essential_from_8(a, b, E, &reason);
  for (int q = 0; q < 9; ++q) Es[(size_t)h * 9 + q] = ok ? E[q] : 0.0;
  valid[h] = ok ? 1 : -reason;             // <= 0: rejected sample (the reason is kept for MVO_EPI_DEBUG)
}

__global__ void __launch_bounds__(256)
k_epi_score(const float *__restrict__ p1, const float *__restrict__ p2, int n, EpiCam cam, double thr2, int H,
            const double *__restrict__ Es, const int32_t *__restrict__ valid, int32_t *__restrict__ counts) {
  extern __shared__ double s_pt[];         // [n][4] calibrated coordinates x1 y1 x2 y2
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s_pt[4 * i] = ((double)p1[2 * i] - cam.cx) / cam.f; s_pt[4 * i + 1] = ((double)p1[2 * i + 1] - cam.cy) / cam.f;
    s_pt[4 * i + 2] = ((double)p2[2 * i] - cam.cx) / cam.f; s_pt[4 * i + 3] = ((double)p2[2 * i + 1] - cam.cy) / cam.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int h = blockIdx.x * wpb + warp; h < H; h += gridDim.x * wpb) {
    if (valid[h] <= 0) { if (lane == 0) counts[h] = -1; continue; }
    double E[9];
    for (int q = 0; q < 9; ++q) E[q] = Es[(size_t)h * 9 + q];
    int c = 0;
    for (int i = lane; i < n; i += 32) c += epi::sampson_err(E, s_pt[4 * i], s_pt[4 * i + 1], s_pt[4 * i + 2], s_pt[4 * i + 3]) <= thr2;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane == 0) counts[h] = c;
  }
}

constexpr int EFIN_T = 1024;
constexpr int EPI_LO_ROUNDS = 3, EPI_GN_ITERS = 8;

__device__ __forceinline__ void skew_times(const double *t, const double *R, double *E) {      // E = [t]x R
  for (int j = 0; j < 3; ++j) {
    E[0 * 3 + j] = -t[2] * R[1 * 3 + j] + t[1] * R[2 * 3 + j];
    E[1 * 3 + j] = t[2] * R[0 * 3 + j] - t[0] * R[2 * 3 + j];
    E[2 * 3 + j] = -t[1] * R[0 * 3 + j] + t[0] * R[1 * 3 + j];
  }
}

__device__ void so3_exp(const double *w, double *R) {
  const double th2 = w[0] * w[0] + w[1] * w[1] + w[2] * w[2], th = sqrt(th2);
  double A, B;
  if (th < 1e-8) { A = 1 - th2 / 6; B = 0.5 - th2 / 24; }
  else { A = sin(th) / th; B = (1 - cos(th)) / th2; }
  const double x = w[0], y = w[1], z = w[2];
  R[0] = 1 - B * (y * y + z * z); R[1] = -A * z + B * x * y;      R[2] = A * y + B * x * z;
  R[3] = A * z + B * x * y;       R[4] = 1 - B * (x * x + z * z); R[5] = -A * x + B * y * z;
  R[6] = -A * y + B * x * z;      R[7] = A * x + B * y * z;       R[8] = 1 - B * (x * x + y * y);
}

// (R, t) moved by d = (rotation increment on the right, two tangent-plane components of the unit translation)
__device__ void epi_retract(const double *R, const double *t, const double *d, double *Ro, double *to) {
  double dR[9];
  so3_exp(d, dR);
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) Ro[i * 3 + j] = R[i * 3] * dR[j] + R[i * 3 + 1] * dR[3 + j] + R[i * 3 + 2] * dR[6 + j];
  double a[3] = {1, 0, 0};
  if (fabs(t[0]) >= 0.9) { a[0] = 0; a[1] = 1; }
  double b1[3], b2[3];
  epi::cross3(t, a, b1);
  const double n1 = sqrt(b1[0] * b1[0] + b1[1] * b1[1] + b1[2] * b1[2]);
  for (int i = 0; i < 3; ++i) b1[i] /= n1;
  epi::cross3(t, b1, b2);
  double n = 0;
  for (int i = 0; i < 3; ++i) { to[i] = t[i] + d[3] * b1[i] + d[4] * b2[i]; n += to[i] * to[i]; }
  n = sqrt(n);
  for (int i = 0; i < 3; ++i) to[i] /= n;
}

__device__ __forceinline__ double sampson_signed(const double *E, double x1, double y1, double x2, double y2) {
  const double Ex0 = E[0] * x1 + E[1] * y1 + E[2], Ex1 = E[3] * x1 + E[4] * y1 + E[5], Ex2 = E[6] * x1 + E[7] * y1 + E[8];
  const double Et0 = E[0] * x2 + E[3] * y2 + E[6], Et1 = E[1] * x2 + E[4] * y2 + E[7];
  return (x2 * Ex0 + y2 * Ex1 + Ex2) * rsqrt(Ex0 * Ex0 + Ex1 * Ex1 + Et0 * Et0 + Et1 * Et1);
}

// out_d: [0..8] E (scaled so that E[8] = 1, epipolar_geometry.cpp:37), [9..17] R, [18..20] t (unit); out_i: [0] inliers,
// [1] best hypothesis, [2] cheirality votes of the chosen (R, t), [3] consensus of the best minimal model
//
// After the arg-max the best minimal model is locally optimised (LO-RANSAC style): EPI_LO_ROUNDS times its consensus
// set is re-selected and (R, t/|t|) is re-estimated on it by Gauss-Newton on the signed Sampson residual (5 degrees of
// freedom, forward-difference Jacobian).  OpenCV returns the minimal five-point model as it is; the eight-point
// minimal models used here are noisier, and the local optimisation more than makes up for it (pose errors against
// synthetic truth ~10x below cv2's on the test scenes).  The reported inliers are the consensus set of the final model.
__global__ void __launch_bounds__(EFIN_T)
k_epi_finish(const float *__restrict__ p1, const float *__restrict__ p2, int n, EpiCam cam, double thr2, int H,
             const double *__restrict__ Es, const int32_t *__restrict__ counts, double *__restrict__ out_d,
             int32_t *__restrict__ out_i, int32_t *__restrict__ inl) {
  __shared__ long long s_k[32];
  __shared__ int s_best, s_cnt[32], s_good[4][32], s_stop;
  __shared__ double s_E[9], s_R1[9], s_R2[9], s_t[3], s_R[9], s_E0[9];
  __shared__ double s_Ek[6][9], s_red[32][20], s_sum[20];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long best = -1;
  for (int h = tid; h < H; h += EFIN_T) {
    const long long key = ((long long)counts[h] << 20) | (long long)(0xFFFFF - h);
    best = key > best ? key : best;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { const long long o = __shfl_xor_sync(0xffffffffu, best, d); best = o > best ? o : best; }
  if (lane == 0) s_k[warp] = best;
  __syncthreads();
  if (tid == 0) {
    long long b = -1;
    for (int w = 0; w < 32; ++w) b = s_k[w] > b ? s_k[w] : b;
    s_best = (int)(b >> 20) >= 8 ? (int)(0xFFFFF - (b & 0xFFFFF)) : -1;
    out_i[3] = (int)(b >> 20);
  }
  __syncthreads();
  if (s_best < 0) {
    if (tid == 0) { out_i[0] = 0; out_i[1] = -1; out_i[2] = 0; }
    return;
  }
  if (tid == 0) {
    for (int q = 0; q < 9; ++q) s_E[q] = Es[(size_t)s_best * 9 + q];
    epi::decompose_essential(s_E, s_R1, s_R2, s_t);
    for (int q = 0; q < 9; ++q) { s_R[q] = s_R1[q]; s_E0[q] = s_E[q]; }      // either rotation of the twisted pair spans the same E = [t]x R (up to sign)
    skew_times(s_t, s_R, s_E);
  }
  __syncthreads();
  // this thread's points (contiguous chunk), calibrated coordinates
  const int per = (n + EFIN_T - 1) / EFIN_T, b0 = tid * per, e0 = min(b0 + per, n);
  // ---- local optimisation ----
  for (int round = 0; round < EPI_LO_ROUNDS; ++round) {
    double Esel[9];
    for (int q = 0; q < 9; ++q) Esel[q] = s_E[q];          // consensus set of this round: fixed during its GN iterations
    for (int it = 0; it < EPI_GN_ITERS; ++it) {
      if (tid < 6) {                                         // E at the current estimate and at its 5 forward perturbations
        double d[5] = {0, 0, 0, 0, 0}, Rk[9], tk[3];
        if (tid > 0) d[tid - 1] = 1e-6;
        epi_retract(s_R, s_t, d, Rk, tk);
        skew_times(tk, Rk, s_Ek[tid]);
      }
      if (tid == 0) s_stop = 0;
      __syncthreads();
      double acc[20];
#pragma unroll
      for (int q = 0; q < 20; ++q) acc[q] = 0;
      for (int i = b0; i < e0; ++i) {
        const double x1 = ((double)p1[2 * i] - cam.cx) / cam.f, y1 = ((double)p1[2 * i + 1] - cam.cy) / cam.f;
        const double x2 = ((double)p2[2 * i] - cam.cx) / cam.f, y2 = ((double)p2[2 * i + 1] - cam.cy) / cam.f;
        if (!(epi::sampson_err(Esel, x1, y1, x2, y2) <= thr2)) continue;
        const double r0 = sampson_signed(s_Ek[0], x1, y1, x2, y2);
        double J[5];
#pragma unroll
        for (int k = 0; k < 5; ++k) J[k] = (sampson_signed(s_Ek[k + 1], x1, y1, x2, y2) - r0) * 1e6;
        int q = 0;
#pragma unroll
        for (int r = 0; r < 5; ++r)
#pragma unroll
          for (int c = r; c < 5; ++c) acc[q++] += J[r] * J[c];
#pragma unroll
        for (int r = 0; r < 5; ++r) acc[15 + r] += J[r] * r0;
      }
#pragma unroll
      for (int q = 0; q < 20; ++q) {
        double v = acc[q];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
        if (lane == 0) s_red[warp][q] = v;
      }
      __syncthreads();
      if (tid < 20) { double v = 0; for (int w = 0; w < 32; ++w) v += s_red[w][tid]; s_sum[tid] = v; }
      __syncthreads();
      if (tid == 0) {
        // (J^T J + eps I) dx = -J^T r by Gaussian elimination with partial pivoting
        double A[5][6];
        int q = 0;
        for (int r = 0; r < 5; ++r) for (int c = r; c < 5; ++c) { A[r][c] = s_sum[q]; A[c][r] = s_sum[q]; ++q; }
        double tr = 0;
        for (int r = 0; r < 5; ++r) tr += A[r][r];
        for (int r = 0; r < 5; ++r) { A[r][r] += 1e-12 * tr + 1e-300; A[r][5] = -s_sum[15 + r]; }
        bool ok = true;
        for (int k = 0; k < 5 && ok; ++k) {
          int pr = k;
          for (int r = k + 1; r < 5; ++r) if (fabs(A[r][k]) > fabs(A[pr][k])) pr = r;
          if (!(fabs(A[pr][k]) > 0)) { ok = false; break; }
          if (pr != k) for (int c = 0; c < 6; ++c) { const double t_ = A[k][c]; A[k][c] = A[pr][c]; A[pr][c] = t_; }
          for (int r = k + 1; r < 5; ++r) { const double f = A[r][k] / A[k][k]; for (int c = k; c < 6; ++c) A[r][c] -= f * A[k][c]; }
        }
        double dx[5] = {0, 0, 0, 0, 0}, mx = 0;
        if (ok) {
          for (int r = 4; r >= 0; --r) { double v = A[r][5]; for (int c = r + 1; c < 5; ++c) v -= A[r][c] * dx[c]; dx[r] = v / A[r][r]; }
          for (int r = 0; r < 5; ++r) { if (!isfinite(dx[r])) ok = false; mx = fmax(mx, fabs(dx[r])); }
        }
        if (ok && mx < 0.5) {                                 // a Gauss-Newton step of half a radian is not a refinement: keep the estimate
          double Rn[9], tn[3];
          epi_retract(s_R, s_t, dx, Rn, tn);
          for (int r = 0; r < 9; ++r) s_R[r] = Rn[r];
          for (int r = 0; r < 3; ++r) s_t[r] = tn[r];
        }
        if (!ok || mx < 1e-10 || mx >= 0.5) s_stop = 1;
      }
      __syncthreads();
      if (s_stop) break;
    }
    if (tid == 0) skew_times(s_t, s_R, s_E);
    __syncthreads();
  }
  // the local optimisation must not lose support: otherwise the minimal model stands
  {
    int c = 0;
    for (int i = b0; i < e0; ++i) {
      const double x1 = ((double)p1[2 * i] - cam.cx) / cam.f, y1 = ((double)p1[2 * i + 1] - cam.cy) / cam.f;
      const double x2 = ((double)p2[2 * i] - cam.cx) / cam.f, y2 = ((double)p2[2 * i + 1] - cam.cy) / cam.f;
      c += epi::sampson_err(s_E, x1, y1, x2, y2) <= thr2;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane == 0) s_cnt[warp] = c;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += s_cnt[w];
      out_i[4] = tot;
      if (!(tot >= out_i[3])) for (int q = 0; q < 9; ++q) s_E[q] = s_E0[q];
    }
    __syncthreads();
  }
  // ---- consensus set of the final model, ascending (the mask of findEssentialMat, epipolar_geometry.cpp:40-47) ----
  if (tid == 0) epi::decompose_essential(s_E, s_R1, s_R2, s_t);
  double E[9];
  for (int q = 0; q < 9; ++q) E[q] = s_E[q];
  int mine = 0;
  for (int i = b0; i < e0; ++i) {
    const double x1 = ((double)p1[2 * i] - cam.cx) / cam.f, y1 = ((double)p1[2 * i + 1] - cam.cy) / cam.f;
    const double x2 = ((double)p2[2 * i] - cam.cx) / cam.f, y2 = ((double)p2[2 * i + 1] - cam.cy) / cam.f;
    mine += epi::sampson_err(E, x1, y1, x2, y2) <= thr2;
  }
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
  if (lane == 31) s_cnt[warp] = incl;
  __syncthreads();                                   // also publishes R1, R2, t
  int off = incl - mine, n_in = 0;
  for (int w = 0; w < 32; ++w) { if (w < warp) off += s_cnt[w]; n_in += s_cnt[w]; }
  // recoverPose (calib3d five-point.cpp): triangulate every inlier with [I|0] and each of (R1,t) (R2,t) (R1,-t) (R2,-t);
  // a point votes for a candidate when its depth is in (0, 50) in both cameras
  int good[4] = {0, 0, 0, 0};
  const double P0[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  for (int i = b0; i < e0; ++i) {
    const double x1 = ((double)p1[2 * i] - cam.cx) / cam.f, y1 = ((double)p1[2 * i + 1] - cam.cy) / cam.f;
    const double x2 = ((double)p2[2 * i] - cam.cx) / cam.f, y2 = ((double)p2[2 * i + 1] - cam.cy) / cam.f;
    if (!(epi::sampson_err(E, x1, y1, x2, y2) <= thr2)) continue;
    inl[off++] = i;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const double *R = (c & 1) ? s_R2 : s_R1;
      const double sg = c < 2 ? 1.0 : -1.0;
      double P[12];
      for (int r = 0; r < 3; ++r) { P[4 * r] = R[3 * r]; P[4 * r + 1] = R[3 * r + 1]; P[4 * r + 2] = R[3 * r + 2]; P[4 * r + 3] = sg * s_t[r]; }
      double X[4];
      epi::triangulate_dlt(P0, P, x1, y1, x2, y2, X);
      bool ok = X[2] * X[3] > 0;                     // mask = Q.z * Q.w > 0
      const double z1 = X[2] / X[3];
      ok = ok && z1 < 50.0;                          // distanceThresh
      const double z2 = (P[8] * X[0] + P[9] * X[1] + P[10] * X[2] + P[11] * X[3]) / X[3];
      ok = ok && z2 > 0 && z2 < 50.0;
      good[c] += ok;
    }
  }
#pragma unroll
  for (int c = 0; c < 4; ++c) {
    int g = good[c];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) g += __shfl_xor_sync(0xffffffffu, g, d);
    if (lane == 0) s_good[c][warp] = g;
  }
  __syncthreads();
  if (tid == 0) {
    int g[4] = {0, 0, 0, 0};
    for (int c = 0; c < 4; ++c) for (int w = 0; w < 32; ++w) g[c] += s_good[c][w];
    // OpenCV's order: (R1,t) if good1 is a maximum, else (R2,t), else (R1,-t), else (R2,-t)
    int pick = 3;
    if (g[0] >= g[1] && g[0] >= g[2] && g[0] >= g[3]) pick = 0;
    else if (g[1] >= g[0] && g[1] >= g[2] && g[1] >= g[3]) pick = 1;
    else if (g[2] >= g[0] && g[2] >= g[1] && g[2] >= g[3]) pick = 2;
    const double *R = (pick & 1) ? s_R2 : s_R1;
    const double sg = pick < 2 ? 1.0 : -1.0;
    const double e22 = s_E[8];
    for (int q = 0; q < 9; ++q) { out_d[q] = s_E[q] / e22; out_d[9 + q] = R[q]; }       // E /= E(2,2) (:37)
    const double nt = sqrt(s_t[0] * s_t[0] + s_t[1] * s_t[1] + s_t[2] * s_t[2]);          // t /= |t| (:54-55)
    for (int q = 0; q < 3; ++q) out_d[18 + q] = sg * s_t[q] / nt;
    out_i[0] = n_in; out_i[1] = s_best; out_i[2] = g[pick];
  }
}

__global__ void __launch_bounds__(128)
k_triangulate(const float *__restrict__ np1, const float *__restrict__ np2, const int32_t *__restrict__ inl, int n_in,
              const double *__restrict__ Rt /* 9 + 3 */, float *__restrict__ out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n_in) return;
  const int i = inl[j];
  const double P0[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
  double P[12];
  for (int r = 0; r < 3; ++r) { P[4 * r] = Rt[3 * r]; P[4 * r + 1] = Rt[3 * r + 1]; P[4 * r + 2] = Rt[3 * r + 2]; P[4 * r + 3] = Rt[9 + r]; }
  double X[4];
  epi::triangulate_dlt(P0, P, np1[2 * i], np1[2 * i + 1], np2[2 * i], np2[2 * i + 1], X);
  // pts4d is CV_32F for float input points; the reference divides by the fourth coordinate in float (:161-168)
  const float x = (float)X[0], y = (float)X[1], z = (float)X[2], w = (float)X[3];
  out[3 * j] = x / w; out[3 * j + 1] = y / w; out[3 * j + 2] = z / w;
}

// ---------------------------------------------------------------------------------------------------- homography
// estiMotionByHomography (reference src/geometry/epipolar_geometry.cpp:90-128): cv::findHomography(pts1, pts2, RANSAC,
// 3.0, mask) + H /= H(2,2) + inliers from the mask + cv::decomposeHomographyMat(H, K) + t /= |t|.  Same batched design
// as the essential-matrix path: H four-point hypotheses, forward transfer error consensus (the error OpenCV
// thresholds), local optimisation of the best one by Gauss-Newton on the transfer error over its consensus set
// (OpenCV refines its RANSAC result with LM on the same cost).  All three kernels work in isotropically scaled pixel
// coordinates ((u - cx) / f, (v - cy) / f, f = mean focal length): errors scale by 1 / f, the conditioning of the
// 4-point systems and of the normal equations does not depend on the image size.  The decomposition itself is a few
// hundred flops and runs on the host (the same epipolar_math.cuh routine).
struct HomoCam { double f, cx, cy; };

__global__ void __launch_bounds__(128)
k_homo_hypotheses(const float *__restrict__ p1, const float *__restrict__ p2, int n, HomoCam cam, uint64_t seed, int H,
                  double *__restrict__ Hs, int32_t *__restrict__ valid) {
  const int h = blockIdx.x * blockDim.x + threadIdx.x;
  if (h >= H) return;
  int idx[4];
  uint64_t ctr = 0;
  for (int k = 0; k < 4; ++k)
    for (int attempt = 0; attempt < 64; ++attempt) {
      const uint64_t r = splitmix64e(seed ^ splitmix64e(((uint64_t)h << 20) ^ (0x5bd1e995ull + ctr++)));
      const int cand = (int)(r % (uint64_t)n);
      bool dup = false;
      for (int q = 0; q < k; ++q) dup |= idx[q] == cand;
      idx[k] = cand;
      if (!dup) break;
    }
  double a[8], b[8];
  const double inv = 1.0 / cam.f;
  for (int k = 0; k < 4; ++k) {
    a[2 * k] = ((double)p1[2 * idx[k]] - cam.cx) * inv; a[2 * k + 1] = ((double)p1[2 * idx[k] + 1] - cam.cy) * inv;
    b[2 * k] = ((double)p2[2 * idx[k]] - cam.cx) * inv; b[2 * k + 1] = ((double)p2[2 * idx[k] + 1] - cam.cy) * inv;
  }
  double Hm[9];
  const bool ok = epi::homography_from_4(a, b, Hm);
  for (int q = 0; q < 9; ++q) Hs[(size_t)h * 9 + q] = ok ? Hm[q] : 0.0;
  valid[h] = ok ? 1 : 0;
}

__global__ void __launch_bounds__(256)
k_homo_score(const float *__restrict__ p1, const float *__restrict__ p2, int n, HomoCam cam, double thr2, int H,
             const double *__restrict__ Hs, const int32_t *__restrict__ valid, int32_t *__restrict__ counts) {
  extern __shared__ double s_pt[];         // [n][4] scaled coordinates x1 y1 x2 y2
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    s_pt[4 * i] = ((double)p1[2 * i] - cam.cx) / cam.f; s_pt[4 * i + 1] = ((double)p1[2 * i + 1] - cam.cy) / cam.f;
    s_pt[4 * i + 2] = ((double)p2[2 * i] - cam.cx) / cam.f; s_pt[4 * i + 3] = ((double)p2[2 * i + 1] - cam.cy) / cam.f;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, wpb = blockDim.x >> 5;
  for (int h = blockIdx.x * wpb + warp; h < H; h += gridDim.x * wpb) {
    if (valid[h] <= 0) { if (lane == 0) counts[h] = -1; continue; }
    double Hm[9];
    for (int q = 0; q < 9; ++q) Hm[q] = Hs[(size_t)h * 9 + q];
    int c = 0;
    for (int i = lane; i < n; i += 32) c += epi::homography_transfer_err(Hm, s_pt[4 * i], s_pt[4 * i + 1], s_pt[4 * i + 2], s_pt[4 * i + 3]) <= thr2;
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane == 0) counts[h] = c;
  }
}

constexpr int HOMO_NV = epi::HOMO_GN_NV;    // 45 entries of J^T J (upper triangle) + 9 of J^T r

// out_d: [0..8] H in pixel coordinates scaled so that H[8] = 1; out_i: [0] inliers, [1] best hypothesis,
// [3] consensus of the best minimal model, [4] consensus after the local optimisation
__global__ void __launch_bounds__(EFIN_T, 1)
k_homo_finish(const float *__restrict__ p1, const float *__restrict__ p2, int n, HomoCam cam, double thr2, int H,
              const double *__restrict__ Hs, const int32_t *__restrict__ counts, double *__restrict__ out_d,
              int32_t *__restrict__ out_i, int32_t *__restrict__ inl) {
  __shared__ long long s_k[32];
  __shared__ int s_best, s_cnt[32], s_stop;
  __shared__ double s_H[9], s_H0[9], s_red[32][HOMO_NV], s_sum[HOMO_NV];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  long long best = -1;
  for (int h = tid; h < H; h += EFIN_T) {
    const long long key = ((long long)counts[h] << 20) | (long long)(0xFFFFF - h);
    best = key > best ? key : best;
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) { const long long o = __shfl_xor_sync(0xffffffffu, best, d); best = o > best ? o : best; }
  if (lane == 0) s_k[warp] = best;
  __syncthreads();
  if (tid == 0) {
    long long b = -1;
    for (int w = 0; w < 32; ++w) b = s_k[w] > b ? s_k[w] : b;
    s_best = (int)(b >> 20) >= 4 ? (int)(0xFFFFF - (b & 0xFFFFF)) : -1;
    out_i[3] = (int)(b >> 20);
  }
  __syncthreads();
  if (s_best < 0) {
    if (tid == 0) { out_i[0] = 0; out_i[1] = -1; out_i[4] = 0; }
    return;
  }
  if (tid < 9) { const double v = Hs[(size_t)s_best * 9 + tid]; s_H[tid] = v; s_H0[tid] = v; }
  __syncthreads();
  const int per = (n + EFIN_T - 1) / EFIN_T, b0 = tid * per, e0 = min(b0 + per, n);
  // ---- local optimisation: Gauss-Newton on the transfer error, additive update of the 9 entries (the scale of H is a
  // null direction of the normal equations: a small damping fixes the gauge, H is renormalised after every step) ----
  for (int round = 0; round < EPI_LO_ROUNDS; ++round) {
    double Hsel[9];
    for (int q = 0; q < 9; ++q) Hsel[q] = s_H[q];
    for (int it = 0; it < EPI_GN_ITERS; ++it) {
      double Hc[9];
      for (int q = 0; q < 9; ++q) Hc[q] = s_H[q];
      if (tid == 0) s_stop = 0;
      double acc[HOMO_NV];
#pragma unroll
      for (int q = 0; q < HOMO_NV; ++q) acc[q] = 0;
      for (int i = b0; i < e0; ++i) {
        const double x1 = ((double)p1[2 * i] - cam.cx) / cam.f, y1 = ((double)p1[2 * i + 1] - cam.cy) / cam.f;
        const double x2 = ((double)p2[2 * i] - cam.cx) / cam.f, y2 = ((double)p2[2 * i + 1] - cam.cy) / cam.f;
        if (!(epi::homography_transfer_err(Hsel, x1, y1, x2, y2) <= thr2)) continue;
        epi::homography_gn_accumulate(Hc, x1, y1, x2, y2, acc);
      }
#pragma unroll
      for (int q = 0; q < HOMO_NV; ++q) {
        double vq = acc[q];
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) vq += __shfl_xor_sync(0xffffffffu, vq, d);
        if (lane == 0) s_red[warp][q] = vq;
      }
      __syncthreads();
      if (tid < HOMO_NV) { double vq = 0; for (int w = 0; w < 32; ++w) vq += s_red[w][tid]; s_sum[tid] = vq; }
      __syncthreads();
      if (tid == 0) {
        double Hn[9];
        for (int q = 0; q < 9; ++q) Hn[q] = s_H[q];
        if (epi::homography_gn_step(s_sum, Hn)) s_stop = 1;
        for (int q = 0; q < 9; ++q) s_H[q] = Hn[q];
      }
      __syncthreads();
      if (s_stop) break;
    }
    __syncthreads();
  }
  // the local optimisation must not lose support: otherwise the minimal model stands
  {
    int c = 0;
    for (int i = b0; i < e0; ++i) {
      const double x1 = ((double)p1[2 * i] - cam.cx) / cam.f, y1 = ((double)p1[2 * i + 1] - cam.cy) / cam.f;
      const double x2 = ((double)p2[2 * i] - cam.cx) / cam.f, y2 = ((double)p2[2 * i + 1] - cam.cy) / cam.f;
      c += epi::homography_transfer_err(s_H, x1, y1, x2, y2) <= thr2;
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
    if (lane == 0) s_cnt[warp] = c;
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int w = 0; w < 32; ++w) tot += s_cnt[w];
      out_i[4] = tot;
      if (!(tot >= out_i[3])) for (int q = 0; q < 9; ++q) s_H[q] = s_H0[q];
    }
    __syncthreads();
  }
  // ---- consensus set of the final model, ascending (the mask of findHomography, epipolar_geometry.cpp:108-116) ----
  double Hf[9];
  for (int q = 0; q < 9; ++q) Hf[q] = s_H[q];
  int mine = 0;
  for (int i = b0; i < e0; ++i) {
    const double x1 = ((double)p1[2 * i] - cam.cx) / cam.f, y1 = ((double)p1[2 * i + 1] - cam.cy) / cam.f;
    const double x2 = ((double)p2[2 * i] - cam.cx) / cam.f, y2 = ((double)p2[2 * i + 1] - cam.cy) / cam.f;
    mine += epi::homography_transfer_err(Hf, x1, y1, x2, y2) <= thr2;
  }
  int incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(0xffffffffu, incl, o); if (lane >= o) incl += u; }
  __syncthreads();
  if (lane == 31) s_cnt[warp] = incl;
  __syncthreads();
  int off = incl - mine, n_in = 0;
  for (int w = 0; w < 32; ++w) { if (w < warp) off += s_cnt[w]; n_in += s_cnt[w]; }
  for (int i = b0; i < e0; ++i) {
    const double x1 = ((double)p1[2 * i] - cam.cx) / cam.f, y1 = ((double)p1[2 * i + 1] - cam.cy) / cam.f;
    const double x2 = ((double)p2[2 * i] - cam.cx) / cam.f, y2 = ((double)p2[2 * i + 1] - cam.cy) / cam.f;
    if (epi::homography_transfer_err(Hf, x1, y1, x2, y2) <= thr2) inl[off++] = i;
  }
  if (tid == 0) {
    // back to pixel coordinates: x_scaled = S x_pix with S = [1/f 0 -cx/f; 0 1/f -cy/f; 0 0 1]  =>  H_pix = S^-1 H S
    const double f = cam.f, cx = cam.cx, cy = cam.cy;
    const double S[9] = {1 / f, 0, -cx / f, 0, 1 / f, -cy / f, 0, 0, 1}, Si[9] = {f, 0, cx, 0, f, cy, 0, 0, 1};
    double T[9], Hp[9];
    epi::mat3_mul(Hf, S, T);
    epi::mat3_mul(Si, T, Hp);
    for (int q = 0; q < 9; ++q) out_d[q] = Hp[q] / Hp[8];                       // H /= H(2,2) (:107)
    out_i[0] = n_in; out_i[1] = s_best;
  }
}

// device-side unit test of epipolar_math.cuh (test hook: tests compare with the host build of the same header)
__global__ void k_epi_math_test(const double *__restrict__ in, double *__restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  // in: [0..15] xy1 (8x2), [16..31] xy2, [32..40] a 3x3 matrix, [41..112] an 8x9 matrix
  double E[9];
  int reason = -1;
  const bool ok = epi::essential_from_8(in, in + 16, E, &reason);
  out[0] = ok ? 1 : 0; out[1] = reason;
  for (int q = 0; q < 9; ++q) out[2 + q] = E[q];
  double U[9], sv[3], V[9];
  epi::svd3(in + 32, U, sv, V);
  for (int q = 0; q < 9; ++q) { out[11 + q] = U[q]; out[23 + q] = V[q]; }
  for (int q = 0; q < 3; ++q) out[20 + q] = sv[q];
  double M[72], x[9];
  for (int q = 0; q < 72; ++q) M[q] = in[41 + q];
  out[32] = epi::null_vector_8x9(M, x) ? 1 : 0;
  for (int q = 0; q < 9; ++q) out[33 + q] = x[q];
  double A[9], w[3], Q[9];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) A[i * 3 + j] = in[32 + i * 3 + j] + in[32 + j * 3 + i];      // a symmetric matrix
  epi::sym_eigen_jacobi<3>(A, w, Q);
  for (int q = 0; q < 3; ++q) out[42 + q] = w[q];
  for (int q = 0; q < 9; ++q) out[45 + q] = Q[q];
  // the inside of essential_from_8, step by step
  double M8[72], f[9];
  for (int k = 0; k < 8; ++k) {
    const double x1 = in[2 * k], y1 = in[2 * k + 1], x2 = in[16 + 2 * k], y2 = in[16 + 2 * k + 1];
    double *r = M8 + 9 * k;
    r[0] = x2 * x1; r[1] = x2 * y1; r[2] = x2; r[3] = y2 * x1; r[4] = y2 * y1; r[5] = y2; r[6] = x1; r[7] = y1; r[8] = 1.0;
  }
  out[54] = epi::null_vector_8x9(M8, f) ? 1 : 0;
  epi::svd3(f, U, sv, V);
  for (int q = 0; q < 3; ++q) out[55 + q] = sv[q];
  double A2[9], w2[3], Q2[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A2[i * 3 + j] = f[0 * 3 + i] * f[0 * 3 + j] + f[1 * 3 + i] * f[1 * 3 + j] + f[2 * 3 + i] * f[2 * 3 + j];
  for (int q = 0; q < 3; ++q) out[58 + q] = A2[q * 4];          // diagonal of f^T f
  epi::sym_eigen_jacobi<3>(A2, w2, Q2);
  for (int q = 0; q < 3; ++q) out[61 + q] = w2[q];
}

}  // namespace

extern "C" int mvo_test_epi_math(mvo_ctx *ctx, const double *in /* 113 */, double *out /* 64 */) {
  if (!ctx || !in || !out) return MVO_ERR_INVALID_ARG;
  MVO_CUDA(ctx, cudaSetDevice(ctx->device));
  MVO_TRY(mvo_reserve(ctx, ctx->d_c, 4096));
  double *d = (double *)ctx->d_c.p;
  MVO_CUDA(ctx, cudaMemcpyAsync(d, in, 113 * 8, cudaMemcpyHostToDevice, ctx->stream));
  k_epi_math_test<<<1, 32, 0, ctx->stream>>>(d, d + 128);
  MVO_CHECK_LAUNCH(ctx);
  MVO_CUDA(ctx, cudaMemcpyAsync(out, d + 128, 64 * 8, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  return MVO_OK;
}

extern "C" {

int mvo_esti_motion_by_essential(mvo_ctx *ctx, const float *pts1, const float *pts2, int n, const double *K, double threshold,
                                 double *E, double *R, double *t, int32_t *inliers, int *n_inliers) {
  if (!ctx) return MVO_ERR_INVALID_ARG;
  if (!pts1 || !pts2 || !K || !E || !R || !t || !n_inliers || (*n_inliers > 0 && !inliers))
    return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "estiMotionByEssential: null pointer");
  if (n < 8) return mvo_fail(ctx, MVO_ERR_DEGENERATE, "estiMotionByEssential: %d correspondences (< 8)", n);
  if (!(threshold > 0)) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "estiMotionByEssential: threshold must be positive");
  EpiCam cam;
  cam.f = (K[0] + K[4]) / 2; cam.cx = K[2]; cam.cy = K[5];               // epipolar_geometry.cpp:26-27
  if (!(cam.f > 0)) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "K: focal length must be positive");
  const size_t smem = (size_t)n * 4 * sizeof(double);
  if (smem > 200 * 1024) return mvo_fail(ctx, MVO_ERR_UNSUPPORTED, "estiMotionByEssential: more than %d correspondences", 200 * 1024 / 32);
  MVO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int H = ctx->prm.epi_hypotheses;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_p1 = 0, o_p2 = al(o_p1 + (size_t)n * 8), o_inl = al(o_p2 + (size_t)n * 8), o_E = al(o_inl + (size_t)n * 4);
  const size_t o_valid = al(o_E + (size_t)H * 72), o_cnt = al(o_valid + (size_t)H * 4), o_out = al(o_cnt + (size_t)H * 4), o_end = o_out + 512;
  MVO_TRY(mvo_reserve(ctx, ctx->d_a, o_end));
  MVO_TRY(mvo_reserve_pinned(ctx, ctx->h_a, al((size_t)n * 16) + (size_t)n * 4 + 1024));
  uint8_t *d = (uint8_t *)ctx->d_a.p, *h = (uint8_t *)ctx->h_a.p;
  memcpy(h, pts1, (size_t)n * 8);
  memcpy(h + (size_t)n * 8, pts2, (size_t)n * 8);
  MVO_CUDA(ctx, cudaMemcpyAsync(d + o_p1, h, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  MVO_CUDA(ctx, cudaMemcpyAsync(d + o_p2, h + (size_t)n * 8, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  const float *d1 = (const float *)(d + o_p1), *d2 = (const float *)(d + o_p2);
  double *dE = (double *)(d + o_E), *dout = (double *)(d + o_out);
  int32_t *dvalid = (int32_t *)(d + o_valid), *dcnt = (int32_t *)(d + o_cnt), *dinl = (int32_t *)(d + o_inl), *dout_i = (int32_t *)(d + o_out + 256);
  const double thr2 = (threshold / cam.f) * (threshold / cam.f);          // findEssentialMat: threshold /= focal
  { KTimer kt(ctx, KC_EPI);
  k_epi_hypotheses<<<(H + 127) / 128, 128, 0, ctx->stream>>>(d1, d2, n, cam, ctx->prm.pnp_seed, H, dE, dvalid); }
  MVO_CHECK_LAUNCH(ctx);
  if (smem > 48 * 1024) MVO_CUDA(ctx, cudaFuncSetAttribute(k_epi_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (H + 7) / 8;
  if (grid > 4 * ctx->sm_count) grid = 4 * ctx->sm_count;
  { KTimer kt(ctx, KC_EPI);
  k_epi_score<<<grid, 256, smem, ctx->stream>>>(d1, d2, n, cam, thr2, H, dE, dvalid, dcnt); }
  MVO_CHECK_LAUNCH(ctx);
  { KTimer kt(ctx, KC_EPI);
  k_epi_finish<<<1, EFIN_T, 0, ctx->stream>>>(d1, d2, n, cam, thr2, H, dE, dcnt, dout, dout_i, dinl); }
  MVO_CHECK_LAUNCH(ctx);
  double *h_out = (double *)(h + al((size_t)n * 16));
  int32_t *h_i = (int32_t *)((uint8_t *)h_out + 256), *h_inl = (int32_t *)((uint8_t *)h_out + 512);
  MVO_CUDA(ctx, cudaMemcpyAsync(h_out, dout, 512, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaMemcpyAsync(h_inl, dinl, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int ni = h_i[0];
  if (getenv("MVO_EPI_DEBUG")) {
    std::vector<int32_t> hv(H), hc(H);
    std::vector<double> hE(9);
    cudaMemcpy(hv.data(), dvalid, (size_t)H * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc.data(), dcnt, (size_t)H * 4, cudaMemcpyDeviceToHost);
    cudaMemcpy(hE.data(), dE, 72, cudaMemcpyDeviceToHost);
    int nv = 0, mx = -2, why[4] = {0, 0, 0, 0};
    for (int h2 = 0; h2 < H; ++h2) { nv += hv[h2] > 0; if (hv[h2] <= 0 && hv[h2] > -4) why[-hv[h2]]++; if (hc[h2] > mx) mx = hc[h2]; }
    fprintf(stderr, "epi debug: rejected samples by reason 0..3: %d %d %d %d\n", why[0], why[1], why[2], why[3]);
    fprintf(stderr, "epi debug: H=%d valid=%d max count=%d | finish: inliers %d best %d votes %d minimal %d after LO %d | thr2 %.3e f %.1f | E[0]= %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f %.4f\n",
            H, nv, mx, h_i[0], h_i[1], h_i[2], h_i[3], h_i[4], thr2, cam.f, hE[0], hE[1], hE[2], hE[3], hE[4], hE[5], hE[6], hE[7], hE[8]);
  }
  if (ni < 8) {
    *n_inliers = 0;
    return mvo_fail(ctx, MVO_ERR_DEGENERATE, "estiMotionByEssential: no model reached 8 inliers (best minimal model %d, after local optimisation %d, final %d)",
                    h_i[3], h_i[4], ni);
  }
  if (ni > *n_inliers) return mvo_fail(ctx, MVO_ERR_CAPACITY, "inlier capacity %d < %d", *n_inliers, ni);
  memcpy(E, h_out, 72); memcpy(R, h_out + 9, 72); memcpy(t, h_out + 18, 24);
  memcpy(inliers, h_inl, (size_t)ni * 4);
  *n_inliers = ni;
  return MVO_OK;
}

int mvo_esti_motion_by_homography(mvo_ctx *ctx, const float *pts1, const float *pts2, int n, const double *K, double threshold,
                                  double *Hout, double *Rs, double *ts, double *normals, int *n_solutions, int32_t *inliers,
                                  int *n_inliers) {
  if (!ctx) return MVO_ERR_INVALID_ARG;
  if (!pts1 || !pts2 || !K || !Hout || !Rs || !ts || !normals || !n_solutions || !n_inliers || (*n_inliers > 0 && !inliers))
    return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "estiMotionByHomography: null pointer");
  *n_solutions = 0;
  if (n < 4) return mvo_fail(ctx, MVO_ERR_DEGENERATE, "estiMotionByHomography: %d correspondences (< 4)", n);
  if (!(threshold > 0)) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "estiMotionByHomography: threshold must be positive");
  HomoCam cam;
  cam.f = (K[0] + K[4]) / 2; cam.cx = K[2]; cam.cy = K[5];
  if (!(K[0] > 0 && K[4] > 0)) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "K: focal lengths must be positive");
  const size_t smem = (size_t)n * 4 * sizeof(double);
  if (smem > 200 * 1024) return mvo_fail(ctx, MVO_ERR_UNSUPPORTED, "estiMotionByHomography: more than %d correspondences", 200 * 1024 / 32);
  MVO_CUDA(ctx, cudaSetDevice(ctx->device));
  const int H = ctx->prm.epi_hypotheses;
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_p1 = 0, o_p2 = al(o_p1 + (size_t)n * 8), o_inl = al(o_p2 + (size_t)n * 8), o_H = al(o_inl + (size_t)n * 4);
  const size_t o_valid = al(o_H + (size_t)H * 72), o_cnt = al(o_valid + (size_t)H * 4), o_out = al(o_cnt + (size_t)H * 4), o_end = o_out + 512;
  MVO_TRY(mvo_reserve(ctx, ctx->d_a, o_end));
  MVO_TRY(mvo_reserve_pinned(ctx, ctx->h_a, al((size_t)n * 16) + (size_t)n * 4 + 1024));
  uint8_t *d = (uint8_t *)ctx->d_a.p, *h = (uint8_t *)ctx->h_a.p;
  memcpy(h, pts1, (size_t)n * 8);
  memcpy(h + (size_t)n * 8, pts2, (size_t)n * 8);
  MVO_CUDA(ctx, cudaMemcpyAsync(d + o_p1, h, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  MVO_CUDA(ctx, cudaMemcpyAsync(d + o_p2, h + (size_t)n * 8, (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
  const float *d1 = (const float *)(d + o_p1), *d2 = (const float *)(d + o_p2);
  double *dH = (double *)(d + o_H), *dout = (double *)(d + o_out);
  int32_t *dvalid = (int32_t *)(d + o_valid), *dcnt = (int32_t *)(d + o_cnt), *dinl = (int32_t *)(d + o_inl), *dout_i = (int32_t *)(d + o_out + 256);
  const double thr2 = (threshold / cam.f) * (threshold / cam.f);          // pixels -> scaled coordinates
  MVO_CUDA(ctx, cudaMemsetAsync(dout, 0, 512, ctx->stream));
  { KTimer kt(ctx, KC_EPI);
  k_homo_hypotheses<<<(H + 127) / 128, 128, 0, ctx->stream>>>(d1, d2, n, cam, ctx->prm.pnp_seed, H, dH, dvalid); }
  MVO_CHECK_LAUNCH(ctx);
  if (smem > 48 * 1024) MVO_CUDA(ctx, cudaFuncSetAttribute(k_homo_score, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int grid = (H + 7) / 8;
  if (grid > 4 * ctx->sm_count) grid = 4 * ctx->sm_count;
  { KTimer kt(ctx, KC_EPI);
  k_homo_score<<<grid, 256, smem, ctx->stream>>>(d1, d2, n, cam, thr2, H, dH, dvalid, dcnt); }
  MVO_CHECK_LAUNCH(ctx);
  { KTimer kt(ctx, KC_EPI);
  k_homo_finish<<<1, EFIN_T, 0, ctx->stream>>>(d1, d2, n, cam, thr2, H, dH, dcnt, dout, dout_i, dinl); }
  MVO_CHECK_LAUNCH(ctx);
  double *h_out = (double *)(h + al((size_t)n * 16));
  int32_t *h_i = (int32_t *)((uint8_t *)h_out + 256), *h_inl = (int32_t *)((uint8_t *)h_out + 512);
  MVO_CUDA(ctx, cudaMemcpyAsync(h_out, dout, 512, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaMemcpyAsync(h_inl, dinl, (size_t)n * 4, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  const int ni = h_i[0];
  if (ni < 4) {
    *n_inliers = 0;
    return mvo_fail(ctx, MVO_ERR_DEGENERATE, "estiMotionByHomography: no model reached 4 inliers (best minimal model %d, after local optimisation %d)",
                    h_i[3], h_i[4]);
  }
  if (ni > *n_inliers) return mvo_fail(ctx, MVO_ERR_CAPACITY, "inlier capacity %d < %d", *n_inliers, ni);
  memcpy(Hout, h_out, 72);
  memcpy(inliers, h_inl, (size_t)ni * 4);
  *n_inliers = ni;
  // cv::decomposeHomographyMat(H, K) (:119-120) on the host: Hn = K^-1 H K, then t /= |t| (:122-126)
  const double fx = K[0], fy = K[4], cx = K[2], cy = K[5];
  const double Km[9] = {fx, 0, cx, 0, fy, cy, 0, 0, 1}, Ki[9] = {1 / fx, 0, -cx / fx, 0, 1 / fy, -cy / fy, 0, 0, 1};
  double T[9], Hn[9];
  epi::mat3_mul(Hout, Km, T);
  epi::mat3_mul(Ki, T, Hn);
  const int ns = epi::decompose_homography(Hn, Rs, ts, normals);
  for (int s = 0; s < ns; ++s) {
    const double nt = sqrt(ts[3 * s] * ts[3 * s] + ts[3 * s + 1] * ts[3 * s + 1] + ts[3 * s + 2] * ts[3 * s + 2]);
    if (nt > 0) for (int q = 0; q < 3; ++q) ts[3 * s + q] /= nt;      // a pure rotation keeps t = 0 (the reference divides by zero here)
  }
  *n_solutions = ns;
  return MVO_OK;
}

int mvo_remove_wrong_rt_of_homography(mvo_ctx *ctx, const float *pts_np1, const float *pts_np2, int n, const int32_t *inliers, int n_inliers,
                                      double *Rs, double *ts, double *normals, int *n_solutions) {
  if (!ctx) return MVO_ERR_INVALID_ARG;
  if (!n_solutions || *n_solutions < 0 || *n_solutions > 4 || !Rs || !ts || !normals || (n_inliers > 0 && (!inliers || !pts_np1 || !pts_np2)) || n_inliers < 0)
    return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "removeWrongRtOfHomography: bad arguments");
  for (int j = 0; j < n_inliers; ++j)
    if (inliers[j] < 0 || inliers[j] >= n) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "removeWrongRtOfHomography: inlier index %d outside [0,%d)", inliers[j], n);
  int keep[4] = {0, 0, 0, 0};
  static_assert(sizeof(int) == sizeof(int32_t), "int32 indices");
  epi::filter_homography_solutions(Rs, normals, *n_solutions, pts_np1, pts_np2, (const int *)inliers, n_inliers, keep);
  int w = 0;
  for (int s = 0; s < *n_solutions; ++s)
    if (keep[s]) {
      if (w != s) { memmove(Rs + 9 * w, Rs + 9 * s, 72); memmove(ts + 3 * w, ts + 3 * s, 24); memmove(normals + 3 * w, normals + 3 * s, 24); }
      ++w;
    }
  *n_solutions = w;
  return MVO_OK;
}

int mvo_do_triangulation(mvo_ctx *ctx, const float *pts_np1, const float *pts_np2, int n, const double *R, const double *t,
                         const int32_t *inliers, int n_inliers, float *pts3d) {
  if (!ctx) return MVO_ERR_INVALID_ARG;
  if (n < 0 || n_inliers < 0 || !R || !t || (n > 0 && (!pts_np1 || !pts_np2)) || (n_inliers > 0 && (!inliers || !pts3d)))
    return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "doTriangulation: null pointer");
  for (int j = 0; j < n_inliers; ++j)
    if (inliers[j] < 0 || inliers[j] >= n) return mvo_fail(ctx, MVO_ERR_INVALID_ARG, "doTriangulation: inlier index %d outside [0,%d)", inliers[j], n);
  if (n_inliers == 0) return MVO_OK;
  MVO_CUDA(ctx, cudaSetDevice(ctx->device));
  auto al = [](size_t v) { return (v + 255) & ~(size_t)255; };
  const size_t o_p1 = 0, o_p2 = al((size_t)n * 8), o_inl = al(o_p2 + (size_t)n * 8), o_rt = al(o_inl + (size_t)n_inliers * 4);
  const size_t o_out = al(o_rt + 96), o_end = o_out + (size_t)n_inliers * 12;
  MVO_TRY(mvo_reserve(ctx, ctx->d_b, o_end + 256));
  MVO_TRY(mvo_reserve_pinned(ctx, ctx->h_a, o_end + 256));
  uint8_t *d = (uint8_t *)ctx->d_b.p, *h = (uint8_t *)ctx->h_a.p;
  memcpy(h + o_p1, pts_np1, (size_t)n * 8);
  memcpy(h + o_p2, pts_np2, (size_t)n * 8);
  memcpy(h + o_inl, inliers, (size_t)n_inliers * 4);
  memcpy(h + o_rt, R, 72);
  memcpy(h + o_rt + 72, t, 24);
  MVO_CUDA(ctx, cudaMemcpyAsync(d, h, o_out, cudaMemcpyHostToDevice, ctx->stream));
  { KTimer kt(ctx, KC_EPI);
  k_triangulate<<<(n_inliers + 127) / 128, 128, 0, ctx->stream>>>((const float *)(d + o_p1), (const float *)(d + o_p2), (const int32_t *)(d + o_inl),
                                                                  n_inliers, (const double *)(d + o_rt), (float *)(d + o_out)); }
  MVO_CHECK_LAUNCH(ctx);
  MVO_CUDA(ctx, cudaMemcpyAsync(h + o_out, d + o_out, (size_t)n_inliers * 12, cudaMemcpyDeviceToHost, ctx->stream));
  MVO_CUDA(ctx, cudaStreamSynchronize(ctx->stream));
  memcpy(pts3d, h + o_out, (size_t)n_inliers * 12);
  return MVO_OK;
}

}  // extern "C"

// mvo_params::pnp_mode = 1: the control flow of cv::solvePnPRansac as the reference calls it (src/vo/vo.cpp:314-320:
// useExtrinsicGuess = false, 100 iterations, 2.0 px, confidence 0.999, SOLVEPNP_ITERATIVE), restated from OpenCV 4.13
// (calib3d ptsetreg.cpp RANSACPointSetRegistrator::run, solvepnp.cpp PnPRansacCallback, epnp.cpp, core lapack.cpp
// JacobiSVDImpl_) — CPU restatement and pins: oracle/pnp_cv_oracle.py, tests/test_pnp_oracle.py.
//   * sampler: cv::RNG((uint64)-1), 5 distinct indices per iteration from rng.uniform(0, count) (getSubset)
//   * minimal solver: EPnP on the 5 points (control points from a PCA, barycentric coordinates, M 10 x 12, SVD of M^T M by
//     one-sided Jacobi in OpenCV's sweep order, betas of the N = 1, 2, 3 approximations + 5 Gauss-Newton steps, the
//     candidate with the smallest reprojection error)
//   * scoring: reprojection in double, narrowed to float, squared distance in float, err <= (float)(thr * thr)
//   * the model with the most inliers so far wins (strictly more than max(best, 4)); after every improvement
//     niters = RANSACUpdateNumIters(0.999, outlier ratio, 5, niters); the loop ends at niters
//   * consensus set of the best model -> least-squares refit from that model (k_ba_pose, as in the batched mode)
// All max_iters hypotheses are evaluated in parallel (one warp each); the sequential part of the loop — which hypotheses
// count and where it stops — is replayed afterwards over their inlier counts by one thread, which gives the same result
// as evaluating them one after the other.
// What cannot be reproduced bit for bit by ANY independent implementation is documented in oracle/pnp_cv_oracle.py: on
// samples that contain an outlier the EPnP pose is an artefact of the rounding noise of the SVD.  Samples made of inliers —
// the ones that win — agree with OpenCV to ~1e-6, so the consensus sets are equal up to points within ~1e-3 px of the threshold.
#pragma once

constexpr int CVP_WARPS = 4;          // hypotheses (warps) per CTA
constexpr int CVP_MAXH = 1024;        // iteration cap of this mode (the reference asks for 100)

struct CvRngDev {
  unsigned long long state;
  __device__ unsigned next() {
    state = (unsigned long long)(unsigned)state * 4164903690ull + (unsigned)(state >> 32);
    return (unsigned)state;
  }
  __device__ int uniform(int a, int b) { return a == b ? a : (int)(next() % (unsigned)(b - a) + a); }
};

// cvSolve(A, b, x, CV_SVD) for a 6 x n system (n <= 5): cv::SVD of A by one-sided Jacobi on the rows of A^T (OpenCV's sweep
// order), then SVD::backSubst — x = sum over the singular values above 2 eps sum(w) of v_k (u_k . b) / w_k, i.e. the
// minimum-norm least-squares solution (rank-deficient systems occur with coplanar points).
__device__ void cvp_solve_svd(const double a[6][5], const double *b, int n, double *x) {
  double At[5][6], Vt[5][5], W[5];
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int k = 0; k < 6; ++k) { At[i][k] = a[k][i]; s += a[k][i] * a[k][i]; }
    W[i] = s;
    for (int k = 0; k < n; ++k) Vt[i][k] = i == k ? 1.0 : 0.0;
  }
  const double eps = 2.220446049250313e-16 * 10;
  for (int iter = 0; iter < 30; ++iter) {
    bool changed = false;
    for (int i = 0; i < n - 1; ++i)
      for (int j = i + 1; j < n; ++j) {
        const double aa = W[i], bb = W[j];
        double p = 0;
        for (int k = 0; k < 6; ++k) p += At[i][k] * At[j][k];
        if (fabs(p) <= eps * sqrt(aa * bb)) continue;
        p *= 2;
        const double beta = aa - bb, gamma = hypot(p, beta);
        double c, s;
        if (beta < 0) { s = sqrt((gamma - beta) * 0.5 / gamma); c = p / (gamma * s * 2); }
        else { c = sqrt((gamma + beta) / (gamma * 2)); s = p / (gamma * c * 2); }
        double na = 0, nb = 0;
        for (int k = 0; k < 6; ++k) {
          const double t0 = c * At[i][k] + s * At[j][k], t1 = -s * At[i][k] + c * At[j][k];
          At[i][k] = t0; At[j][k] = t1;
          na += t0 * t0; nb += t1 * t1;
        }
        for (int k = 0; k < n; ++k) {
          const double v0 = c * Vt[i][k] + s * Vt[j][k], v1 = -s * Vt[i][k] + c * Vt[j][k];
          Vt[i][k] = v0; Vt[j][k] = v1;
        }
        W[i] = na; W[j] = nb;
        changed = true;
      }
    if (!changed) break;
  }
  double thr = 0;
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int k = 0; k < 6; ++k) s += At[i][k] * At[i][k];
    W[i] = sqrt(s);
    thr += W[i];
  }
  thr *= 2.220446049250313e-16 * 2;
  for (int q = 0; q < n; ++q) x[q] = 0;
  for (int i = 0; i < n; ++i) {
    if (fabs(W[i]) <= thr) continue;
    double ub = 0;
    for (int k = 0; k < 6; ++k) ub += At[i][k] * b[k];          // (w_i u_i) . b
    const double f = ub / (W[i] * W[i]);
    for (int q = 0; q < n; ++q) x[q] += f * Vt[i][q];
  }
}

// JacobiSVDImpl_'s treatment of zero singular values (exactly singular input: coplanar points): the row is replaced by a
// pseudo-random +-1/m vector (cv::RNG 0x12345678, bit 8 of every draw), orthogonalised against the rows before it in two
// Gram-Schmidt rounds with an L1 renormalisation, then normalised.  W = the sorted singular values; rows = U^T.
template <int N>
__device__ void cvp_fill_zero_rows(double (*At)[N], const double *W) {
  const double tiny = 2.2250738585072014e-308, eps = 2.220446049250313e-16 * 10;
  CvRngDev rng;
  rng.state = 0x12345678ull;
  for (int i = 0; i < N; ++i) {
    double sd = W[i];
    for (int ii = 0; ii < 100 && sd <= tiny; ++ii) {
      const double val0 = 1.0 / N;
      for (int k = 0; k < N; ++k) At[i][k] = (rng.next() & 256) != 0 ? val0 : -val0;
      for (int iter = 0; iter < 2; ++iter)
        for (int j = 0; j < i; ++j) {
          double d = 0;
          for (int k = 0; k < N; ++k) d += At[i][k] * At[j][k];
          double asum = 0;
          for (int k = 0; k < N; ++k) { const double t = At[i][k] - d * At[j][k]; At[i][k] = t; asum += fabs(t); }
          asum = asum > eps * 100 ? 1.0 / asum : 0.0;
          for (int k = 0; k < N; ++k) At[i][k] *= asum;
        }
      sd = 0;
      for (int k = 0; k < N; ++k) sd += At[i][k] * At[i][k];
      sd = sqrt(sd);
    }
    const double s = sd > tiny ? 1.0 / sd : 0.0;
    for (int k = 0; k < N; ++k) At[i][k] *= s;
  }
}

// cv::SVD of a 3 x 3 matrix (rows of At = columns of A on entry): one-sided Jacobi, OpenCV's sweep order; on return the
// rows of At are the left singular vectors (U^T), Vt the right ones, w descending.
__device__ void cvp_jacobi3(double At[3][3], double Vt[3][3], double *w) {
  double W[3];
  for (int i = 0; i < 3; ++i) {
    W[i] = At[i][0] * At[i][0] + At[i][1] * At[i][1] + At[i][2] * At[i][2];
    for (int k = 0; k < 3; ++k) Vt[i][k] = i == k ? 1.0 : 0.0;
  }
  const double eps = 2.220446049250313e-16 * 10;
  for (int iter = 0; iter < 30; ++iter) {
    bool changed = false;
    for (int i = 0; i < 2; ++i)
      for (int j = i + 1; j < 3; ++j) {
        const double a = W[i], b = W[j];
        double p = At[i][0] * At[j][0] + At[i][1] * At[j][1] + At[i][2] * At[j][2];
        if (fabs(p) <= eps * sqrt(a * b)) continue;
        p *= 2;
        const double beta = a - b, gamma = hypot(p, beta);
        double c, s;
        if (beta < 0) { s = sqrt((gamma - beta) * 0.5 / gamma); c = p / (gamma * s * 2); }
        else { c = sqrt((gamma + beta) / (gamma * 2)); s = p / (gamma * c * 2); }
        double na = 0, nb = 0;
        for (int k = 0; k < 3; ++k) {
          const double t0 = c * At[i][k] + s * At[j][k], t1 = -s * At[i][k] + c * At[j][k];
          At[i][k] = t0; At[j][k] = t1;
          na += t0 * t0; nb += t1 * t1;
          const double v0 = c * Vt[i][k] + s * Vt[j][k], v1 = -s * Vt[i][k] + c * Vt[j][k];
          Vt[i][k] = v0; Vt[j][k] = v1;
        }
        W[i] = na; W[j] = nb;
        changed = true;
      }
    if (!changed) break;
  }
  for (int i = 0; i < 3; ++i) W[i] = sqrt(At[i][0] * At[i][0] + At[i][1] * At[i][1] + At[i][2] * At[i][2]);
  for (int i = 0; i < 2; ++i) {
    int j = i;
    for (int k = i + 1; k < 3; ++k) if (W[j] < W[k]) j = k;
    if (i != j) {
      double t = W[i]; W[i] = W[j]; W[j] = t;
      for (int k = 0; k < 3; ++k) { t = At[i][k]; At[i][k] = At[j][k]; At[j][k] = t; t = Vt[i][k]; Vt[i][k] = Vt[j][k]; Vt[j][k] = t; }
    }
  }
  for (int i = 0; i < 3; ++i) w[i] = W[i];
  cvp_fill_zero_rows<3>(At, W);
}

__device__ __forceinline__ double cvp_sum16(double v) {          // sum over lanes 0..15 (12 used), result in every lane of the half warp
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// One warp: one-sided Jacobi of the symmetric 12 x 12 matrix in At (shared, row-major); on return rows of At = U^T, sorted
// by descending singular value.  Lanes 0..11 own column k of every row.
__device__ void cvp_jacobi12(double (*At)[12], double *W, int lane) {
  const int k = lane < 12 ? lane : 0;
  const bool act = lane < 12;
  for (int i = lane; i < 12; i += 32) {
    double s = 0;
    for (int q = 0; q < 12; ++q) s += At[i][q] * At[i][q];
    W[i] = s;
  }
  __syncwarp();
  const double eps = 2.220446049250313e-16 * 10;
  for (int iter = 0; iter < 30; ++iter) {
    bool changed = false;
    for (int i = 0; i < 11; ++i)
      for (int j = i + 1; j < 12; ++j) {
        const double ai = act ? At[i][k] : 0.0, aj = act ? At[j][k] : 0.0;
        double p = cvp_sum16(ai * aj);
        p = __shfl_sync(0xffffffffu, p, 0);
        const double a = W[i], b = W[j];
        if (fabs(p) <= eps * sqrt(a * b)) continue;            // uniform across the warp
        p *= 2;
        const double beta = a - b, gamma = hypot(p, beta);
        double c, s;
        if (beta < 0) { s = sqrt((gamma - beta) * 0.5 / gamma); c = p / (gamma * s * 2); }
        else { c = sqrt((gamma + beta) / (gamma * 2)); s = p / (gamma * c * 2); }
        const double t0 = c * ai + s * aj, t1 = -s * ai + c * aj;
        double na = cvp_sum16(act ? t0 * t0 : 0.0), nb = cvp_sum16(act ? t1 * t1 : 0.0);
        na = __shfl_sync(0xffffffffu, na, 0);
        nb = __shfl_sync(0xffffffffu, nb, 0);
        __syncwarp();
        if (act) { At[i][k] = t0; At[j][k] = t1; }
        if (lane == 0) { W[i] = na; W[j] = nb; }
        __syncwarp();
        changed = true;
      }
    if (!changed) break;
  }
  if (lane == 0) {
    for (int i = 0; i < 12; ++i) {
      double s = 0;
      for (int q = 0; q < 12; ++q) s += At[i][q] * At[i][q];
      W[i] = sqrt(s);
    }
    for (int i = 0; i < 11; ++i) {
      int j = i;
      for (int q = i + 1; q < 12; ++q) if (W[j] < W[q]) j = q;
      if (i != j) {
        double t = W[i]; W[i] = W[j]; W[j] = t;
        for (int q = 0; q < 12; ++q) { t = At[i][q]; At[i][q] = At[j][q]; At[j][q] = t; }
      }
    }
    cvp_fill_zero_rows<12>(At, W);
  }
  __syncwarp();
}

struct CvpShared {
  double At[12][12];
  double W[12];
  double P[5][3], us[5][2], al[5][4], cws[4][3];
};

// epnp::compute_R_and_t for one set of betas (lane 0): pose + mean reprojection error
__device__ double cvp_r_and_t(const CvpShared &S, const double *be, const PnpCam &cam, double *R, double *t) {
  double ccs[4][3] = {};
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j)
      for (int q = 0; q < 3; ++q) ccs[j][q] += be[i] * S.At[11 - i][3 * j + q];
  double pcs[5][3];
  for (int p = 0; p < 5; ++p)
    for (int q = 0; q < 3; ++q) pcs[p][q] = S.al[p][0] * ccs[0][q] + S.al[p][1] * ccs[1][q] + S.al[p][2] * ccs[2][q] + S.al[p][3] * ccs[3][q];
  if (pcs[0][2] < 0)                                  // solve_for_sign
    for (int p = 0; p < 5; ++p) for (int q = 0; q < 3; ++q) pcs[p][q] = -pcs[p][q];
  double pc0[3] = {0, 0, 0}, pw0[3] = {0, 0, 0};
  for (int p = 0; p < 5; ++p) for (int q = 0; q < 3; ++q) { pc0[q] += pcs[p][q]; pw0[q] += S.P[p][q]; }
  for (int q = 0; q < 3; ++q) { pc0[q] /= 5; pw0[q] /= 5; }
  double abt[3][3] = {};
  for (int p = 0; p < 5; ++p)
    for (int j = 0; j < 3; ++j)
      for (int q = 0; q < 3; ++q) abt[j][q] += (pcs[p][j] - pc0[j]) * (S.P[p][q] - pw0[q]);
  double At3[3][3], Vt[3][3], w3[3];
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) At3[i][j] = abt[j][i];
  cvp_jacobi3(At3, Vt, w3);
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) R[3 * i + j] = At3[0][i] * Vt[0][j] + At3[1][i] * Vt[1][j] + At3[2][i] * Vt[2][j];   // U V^T
  const double det = R[0] * R[4] * R[8] + R[1] * R[5] * R[6] + R[2] * R[3] * R[7] - R[2] * R[4] * R[6] - R[1] * R[3] * R[8] - R[0] * R[5] * R[7];
  if (det < 0) { R[6] = -R[6]; R[7] = -R[7]; R[8] = -R[8]; }
  for (int q = 0; q < 3; ++q) t[q] = pc0[q] - (R[3 * q] * pw0[0] + R[3 * q + 1] * pw0[1] + R[3 * q + 2] * pw0[2]);
  double sum = 0;
  for (int p = 0; p < 5; ++p) {
    const double Xc = R[0] * S.P[p][0] + R[1] * S.P[p][1] + R[2] * S.P[p][2] + t[0];
    const double Yc = R[3] * S.P[p][0] + R[4] * S.P[p][1] + R[5] * S.P[p][2] + t[1];
    const double iz = 1.0 / (R[6] * S.P[p][0] + R[7] * S.P[p][1] + R[8] * S.P[p][2] + t[2]);
    const double du = S.us[p][0] - (cam.cx + cam.fx * Xc * iz), dv = S.us[p][1] - (cam.cy + cam.fy * Yc * iz);
    sum += sqrt(du * du + dv * dv);
  }
  return sum / 5;
}

// One warp per RANSAC iteration h: its 5-point subset (replaying the generator from the start) and the EPnP pose.
__global__ void __launch_bounds__(CVP_WARPS * 32)
k_pnp_epnp(const float *__restrict__ p3, const float *__restrict__ p2, int n, const int32_t *__restrict__ n_dev, PnpCam cam, int H,
           double *__restrict__ poses, int32_t *__restrict__ valid) {
  __shared__ CvpShared sh[CVP_WARPS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int h = blockIdx.x * CVP_WARPS + warp;
  if (n_dev) n = min(n, *n_dev);
  if (h >= H) return;
  CvpShared &S = sh[warp];
  if (n < 6) {                         // count == modelPoints / below: no RANSAC in OpenCV either; reported as no model
    if (lane == 0) valid[h] = 0;
    return;
  }
  int bad = 0;
  if (lane == 0) {
    // getSubset for iterations 0..h (the generator is sequential): only the last subset is kept
    CvRngDev rng;
    rng.state = 0xFFFFFFFFFFFFFFFFull;
    int idx[5];
    for (int it = 0; it <= h; ++it)
      for (int i = 0; i < 5;) {
        const int c = rng.uniform(0, n);
        int j = 0;
        for (; j < i; ++j) if (idx[j] == c) break;
        if (j == i) idx[i++] = c;
      }
    const double ifx = 1.0 / cam.fx, ify = 1.0 / cam.fy;
    for (int p = 0; p < 5; ++p) {
      for (int q = 0; q < 3; ++q) S.P[p][q] = (double)p3[3 * idx[p] + q];
      // solvePnP(EPNP): undistortPoints (no distortion) stores (u - cx) / fx as float; epnp::init_points maps it back
      const float xn = (float)(((double)p2[2 * idx[p]] - cam.cx) * ifx), yn = (float)(((double)p2[2 * idx[p] + 1] - cam.cy) * ify);
      S.us[p][0] = (double)xn * cam.fx + cam.cx;
      S.us[p][1] = (double)yn * cam.fy + cam.cy;
    }
    // choose_control_points: centroid + PCA axes scaled by sqrt(eigenvalue / n)
    double c0[3] = {0, 0, 0};
    for (int p = 0; p < 5; ++p) for (int q = 0; q < 3; ++q) c0[q] += S.P[p][q];
    for (int q = 0; q < 3; ++q) c0[q] /= 5;
    double A3[3][3] = {}, Vt[3][3], dc[3];
    for (int p = 0; p < 5; ++p)
      for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) A3[i][j] += (S.P[p][i] - c0[i]) * (S.P[p][j] - c0[j]);
    cvp_jacobi3(A3, Vt, dc);           // symmetric: rows of A3 = U^T afterwards
    for (int q = 0; q < 3; ++q) S.cws[0][q] = c0[q];
    for (int i = 1; i < 4; ++i) {
      const double kk = sqrt(dc[i - 1] / 5);
      for (int q = 0; q < 3; ++q) S.cws[i][q] = c0[q] + kk * A3[i - 1][q];
    }
    // compute_barycentric_coordinates: cvInvert(&CC, &CC_inv, CV_SVD) of the 3 x 3 matrix of control-point offsets — the
    // pseudo-inverse, singular values <= 2 eps sum(w) dropped (coplanar points leave the third control point on the centroid)
    double cct[3][3], cvt[3][3], cw[3], ci[3][3] = {};
    for (int i = 0; i < 3; ++i) for (int j = 1; j < 4; ++j) cct[j - 1][i] = S.cws[j][i] - S.cws[0][i];      // rows of cct = columns of CC
    cvp_jacobi3(cct, cvt, cw);
    const double cthr = 2.220446049250313e-16 * 2 * (cw[0] + cw[1] + cw[2]);
    for (int q = 0; q < 3; ++q)
      if (fabs(cw[q]) > cthr)
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) ci[i][j] += cvt[q][i] * cct[q][j] / cw[q];
    if (!(cw[0] > 0)) bad = 1;                       // all five points coincide
    for (int p = 0; p < 5; ++p) {
      for (int j = 0; j < 3; ++j)
        S.al[p][1 + j] = ci[j][0] * (S.P[p][0] - c0[0]) + ci[j][1] * (S.P[p][1] - c0[1]) + ci[j][2] * (S.P[p][2] - c0[2]);
      S.al[p][0] = 1.0 - S.al[p][1] - S.al[p][2] - S.al[p][3];
    }
    // M^T M (12 x 12) from the 10 rows of fill_M
    for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) S.At[i][j] = 0;
    for (int p = 0; p < 5; ++p) {
      double r1[12], r2[12];
      for (int j = 0; j < 4; ++j) {
        r1[3 * j] = S.al[p][j] * cam.fx; r1[3 * j + 1] = 0; r1[3 * j + 2] = S.al[p][j] * (cam.cx - S.us[p][0]);
        r2[3 * j] = 0; r2[3 * j + 1] = S.al[p][j] * cam.fy; r2[3 * j + 2] = S.al[p][j] * (cam.cy - S.us[p][1]);
      }
      for (int i = 0; i < 12; ++i) for (int j = 0; j < 12; ++j) S.At[i][j] += r1[i] * r1[j] + r2[i] * r2[j];
    }
  }
  bad = __shfl_sync(0xffffffffu, bad, 0);
  __syncwarp();
  if (bad) { if (lane == 0) valid[h] = 0; return; }
  cvp_jacobi12(S.At, S.W, lane);
  if (lane != 0) return;
  // compute_L_6x10 / compute_rho
  double dv[4][6][3];
  for (int i = 0; i < 4; ++i) {
    const double *v = S.At[11 - i];
    int a = 0, b = 1;
    for (int j = 0; j < 6; ++j) {
      for (int q = 0; q < 3; ++q) dv[i][j][q] = v[3 * a + q] - v[3 * b + q];
      if (++b > 3) { ++a; b = a + 1; }
    }
  }
  auto dot3 = [](const double *x, const double *y) { return x[0] * y[0] + x[1] * y[1] + x[2] * y[2]; };
  double L[6][10], rho[6];
  for (int i = 0; i < 6; ++i) {
    L[i][0] = dot3(dv[0][i], dv[0][i]); L[i][1] = 2 * dot3(dv[0][i], dv[1][i]); L[i][2] = dot3(dv[1][i], dv[1][i]);
    L[i][3] = 2 * dot3(dv[0][i], dv[2][i]); L[i][4] = 2 * dot3(dv[1][i], dv[2][i]); L[i][5] = dot3(dv[2][i], dv[2][i]);
    L[i][6] = 2 * dot3(dv[0][i], dv[3][i]); L[i][7] = 2 * dot3(dv[1][i], dv[3][i]); L[i][8] = 2 * dot3(dv[2][i], dv[3][i]);
    L[i][9] = dot3(dv[3][i], dv[3][i]);
  }
  {
    const int pa[6] = {0, 0, 0, 1, 1, 2}, pb[6] = {1, 2, 3, 2, 3, 3};
    for (int i = 0; i < 6; ++i) {
      double s = 0;
      for (int q = 0; q < 3; ++q) { const double d = S.cws[pa[i]][q] - S.cws[pb[i]][q]; s += d * d; }
      rho[i] = s;
    }
  }
  double bestR[9], bestT[3], best_err = 0;
  int have = 0;
  for (int N = 1; N <= 3; ++N) {
    // find_betas_approx_N: least squares on a column subset of L
    const int cols1[4] = {0, 1, 3, 6}, cols2[3] = {0, 1, 2}, cols3[5] = {0, 1, 2, 3, 4};
    const int nc = N == 1 ? 4 : (N == 2 ? 3 : 5);
    const int *cols = N == 1 ? cols1 : (N == 2 ? cols2 : cols3);
    double a[6][5], b[6], x[5] = {0, 0, 0, 0, 0};
    for (int i = 0; i < 6; ++i) { for (int j = 0; j < nc; ++j) a[i][j] = L[i][cols[j]]; b[i] = rho[i]; }
    cvp_solve_svd(a, b, nc, x);
    double be[4] = {0, 0, 0, 0};
    if (N == 1) {
      if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = -x[1] / be[0]; be[2] = -x[2] / be[0]; be[3] = -x[3] / be[0]; }
      else { be[0] = sqrt(x[0]); be[1] = x[1] / be[0]; be[2] = x[2] / be[0]; be[3] = x[3] / be[0]; }
    } else {
      if (x[0] < 0) { be[0] = sqrt(-x[0]); be[1] = x[2] < 0 ? sqrt(-x[2]) : 0.0; }
      else { be[0] = sqrt(x[0]); be[1] = x[2] > 0 ? sqrt(x[2]) : 0.0; }
      if (x[1] < 0) be[0] = -be[0];
      if (N == 3) be[2] = x[3] / be[0];
    }
    // gauss_newton: 5 steps on the 6 distance constraints
    for (int it = 0; it < 5; ++it) {
      double dx[5] = {0, 0, 0, 0, 0};
      for (int i = 0; i < 6; ++i) {
        const double *r = L[i];
        a[i][0] = 2 * r[0] * be[0] + r[1] * be[1] + r[3] * be[2] + r[6] * be[3];
        a[i][1] = r[1] * be[0] + 2 * r[2] * be[1] + r[4] * be[2] + r[7] * be[3];
        a[i][2] = r[3] * be[0] + r[4] * be[1] + 2 * r[5] * be[2] + r[8] * be[3];
        a[i][3] = r[6] * be[0] + r[7] * be[1] + r[8] * be[2] + 2 * r[9] * be[3];
        b[i] = rho[i] - (r[0] * be[0] * be[0] + r[1] * be[0] * be[1] + r[2] * be[1] * be[1] + r[3] * be[0] * be[2] + r[4] * be[1] * be[2] +
                         r[5] * be[2] * be[2] + r[6] * be[0] * be[3] + r[7] * be[1] * be[3] + r[8] * be[2] * be[3] + r[9] * be[3] * be[3]);
      }
      cvp_solve_svd(a, b, 4, dx);
      for (int q = 0; q < 4; ++q) be[q] += dx[q];
    }
    double R[9], t[3];
    const double err = cvp_r_and_t(S, be, cam, R, t);
    // int N = 1; if (rep_errors[2] < rep_errors[1]) N = 2; if (rep_errors[3] < rep_errors[N]) N = 3;  (NaN never wins)
    if (!have || err < best_err) {
      if (!have || err == err) { for (int q = 0; q < 9; ++q) bestR[q] = R[q]; for (int q = 0; q < 3; ++q) bestT[q] = t[q]; best_err = err; }
      have = 1;
    }
  }
  bool fin = true;
  for (int q = 0; q < 9; ++q) fin = fin && isfinite(bestR[q]);
  for (int q = 0; q < 3; ++q) fin = fin && isfinite(bestT[q]);
  for (int q = 0; q < 9; ++q) poses[(size_t)h * 12 + q] = fin ? bestR[q] : (q % 4 == 0 ? 1.0 : 0.0);
  for (int q = 0; q < 3; ++q) poses[(size_t)h * 12 + 9 + q] = fin ? bestT[q] : 0.0;
  valid[h] = fin ? 1 : 0;
}

// PnPRansacCallback::computeError + findInliers: projectPoints in double, narrowed to float; squared distance in float
__device__ __forceinline__ bool cvp_is_inlier(const Pose &P, const PnpCam &cam, V3 X, float u, float v, float thr2f) {
  const double x = P.R[0] * X.x + P.R[1] * X.y + P.R[2] * X.z + P.t[0];
  const double y = P.R[3] * X.x + P.R[4] * X.y + P.R[5] * X.z + P.t[1];
  const double z = P.R[6] * X.x + P.R[7] * X.y + P.R[8] * X.z + P.t[2];
  const double iz = z != 0 ? 1.0 / z : 1.0;                          // projectPoints: z = z ? 1./z : 1
  const float pu = (float)(x * iz * cam.fx + cam.cx), pv = (float)(y * iz * cam.fy + cam.cy);
  const float du = __fsub_rn(u, pu), dv = __fsub_rn(v, pv);
  const float e = __fadd_rn(__fmul_rn(du, du), __fmul_rn(dv, dv));
  return e <= thr2f;
}

// one warp per hypothesis: inlier count under the float rule
__global__ void __launch_bounds__(256)
k_pnp_score_cv(const float *__restrict__ p3, const float *__restrict__ p2, int n, const int32_t *__restrict__ n_dev, PnpCam cam, float thr2f, int H,
               const double *__restrict__ poses, const int32_t *__restrict__ valid, int32_t *__restrict__ counts) {
  const int lane = threadIdx.x & 31, h = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (n_dev) n = min(n, *n_dev);
  if (h >= H) return;
  int c = 0;
  if (valid[h]) {
    Pose P;
    for (int q = 0; q < 9; ++q) P.R[q] = poses[(size_t)h * 12 + q];
    for (int q = 0; q < 3; ++q) P.t[q] = poses[(size_t)h * 12 + 9 + q];
    for (int i = lane; i < n; i += 32)
      c += cvp_is_inlier(P, cam, v3(p3[3 * i], p3[3 * i + 1], p3[3 * i + 2]), p2[2 * i], p2[2 * i + 1], thr2f);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
  if (lane == 0) counts[h] = valid[h] ? c : -1;
}

// RANSACUpdateNumIters (ptsetreg.cpp)
__device__ int cvp_update_num_iters(double p, double ep, int model_points, int max_iters) {
  p = fmin(fmax(p, 0.0), 1.0);
  ep = fmin(fmax(ep, 0.0), 1.0);
  double num = fmax(1.0 - p, 2.2250738585072014e-308);
  double denom = 1.0 - pow(1.0 - ep, (double)model_points);
  if (denom < 2.2250738585072014e-308) return 0;
  num = log(num);
  denom = log(denom);
  return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)rint(num / denom);
}

// The sequential part of RANSACPointSetRegistrator::run replayed over the counts of all H evaluated iterations: the index
// of the model the loop ends with (-1: none), the number of iterations it would have run.
__device__ int cvp_replay(const int32_t *__restrict__ counts, int H, int n, double confidence, int *iters_run) {
  int niters = H, best = -1, max_good = 0, it = 0;
  for (; it < niters; ++it) {
    const int good = counts[it];
    if (good > max(max_good, 4)) {
      best = it;
      max_good = good;
      niters = cvp_update_num_iters(confidence, (double)(n - good) / n, 5, niters);
    }
  }
  *iters_run = it;
  return best;
}

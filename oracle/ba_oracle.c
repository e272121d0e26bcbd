/* TEST ORACLE — CPU restatement (plain C, fp64) of the bundle adjustment the reference runs
 * through g2o.  Not product code: only tests/, smoke() and bench.py's cpu_baseline use it.
 *
 * Reference driver:  src/optimization/g2o_ba.cpp:172-317 (bundleAdjustment) and :34-145
 * (optimizeSingleFrame).  g2o itself ("last version in year 2017", README.md:184) is a
 * third-party dependency ABSENT from the reference tree and from this container, so
 *   >>> PARITY WITH THE REAL g2o IS UNPINNED <<<
 * What is restated here is g2o's published algorithm for exactly the objects the reference
 * instantiates:
 *   VertexSE3Expmap        pose = SE3Quat (unit quaternion + t), world->camera; update
 *                          T <- exp(delta) * T with delta = (omega, upsilon)
 *   VertexSBAPointXYZ      point in R^3, additive update, marginalised (Schur)
 *   EdgeProjectXYZ2UV      e = obs - (f * (x/z, y/z) + c), single focal f = K(0,0)
 *   RobustKernelHuber      delta = 1: rho = e2 (e2 <= d^2) else 2 sqrt(e2) d - d^2
 *   OptimizationAlgorithmLevenberg  tau 1e-5, good-step scale in [1/3, 2/3], <= 10 trials
 *   BlockSolver<6,3> + LinearSolverDense  Schur complement on the points, dense LDLT
 * The optimum is cross-checked against scipy.optimize.least_squares(loss="huber") in
 * tests/test_ba_oracle.py.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct { double q[4]; double t[3]; } Se3;   /* q = (x, y, z, w), unit */

static void quat_normalize(double *q) {
  if (q[3] < 0) { q[0] = -q[0]; q[1] = -q[1]; q[2] = -q[2]; q[3] = -q[3]; }
  double n = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
  for (int i = 0; i < 4; ++i) q[i] /= n;
}

static void quat_to_R(const double *q, double *R) {
  const double x = q[0], y = q[1], z = q[2], w = q[3];
  R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - z * w);     R[2] = 2 * (x * z + y * w);
  R[3] = 2 * (x * y + z * w);     R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - x * w);
  R[6] = 2 * (x * z - y * w);     R[7] = 2 * (y * z + x * w);     R[8] = 1 - 2 * (x * x + y * y);
}

static void R_to_quat(const double *R, double *q) {   /* Eigen::Quaterniond(Matrix3d) */
  double t = R[0] + R[4] + R[8];
  if (t > 0) {
    t = sqrt(t + 1.0);
    q[3] = 0.5 * t;
    t = 0.5 / t;
    q[0] = (R[7] - R[5]) * t; q[1] = (R[2] - R[6]) * t; q[2] = (R[3] - R[1]) * t;
  } else {
    int i = 0;
    if (R[4] > R[0]) i = 1;
    if (R[8] > R[i * 4]) i = 2;
    int j = (i + 1) % 3, k = (j + 1) % 3;
    t = sqrt(R[i * 4] - R[j * 4] - R[k * 4] + 1.0);
    q[i] = 0.5 * t;
    t = 0.5 / t;
    q[3] = (R[k * 3 + j] - R[j * 3 + k]) * t;
    q[j] = (R[j * 3 + i] + R[i * 3 + j]) * t;
    q[k] = (R[k * 3 + i] + R[i * 3 + k]) * t;
  }
}

static void quat_mul(const double *a, const double *b, double *o) {
  o[3] = a[3] * b[3] - a[0] * b[0] - a[1] * b[1] - a[2] * b[2];
  o[0] = a[3] * b[0] + a[0] * b[3] + a[1] * b[2] - a[2] * b[1];
  o[1] = a[3] * b[1] - a[0] * b[2] + a[1] * b[3] + a[2] * b[0];
  o[2] = a[3] * b[2] + a[0] * b[1] - a[1] * b[0] + a[2] * b[3];
}

static void mat3_mul(const double *A, const double *B, double *C) {
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) C[i * 3 + j] = A[i * 3] * B[j] + A[i * 3 + 1] * B[3 + j] + A[i * 3 + 2] * B[6 + j];
}

/* g2o SE3Quat::exp(update), update = (omega, upsilon) — including its small-angle branch */
static void se3_exp(const double *u, Se3 *out) {
  const double om[3] = {u[0], u[1], u[2]}, up[3] = {u[3], u[4], u[5]};
  const double theta = sqrt(om[0] * om[0] + om[1] * om[1] + om[2] * om[2]);
  const double O[9] = {0, -om[2], om[1], om[2], 0, -om[0], -om[1], om[0], 0};
  double O2[9], R[9], V[9];
  mat3_mul(O, O, O2);
  if (theta < 0.00001) {
    for (int i = 0; i < 9; ++i) { R[i] = (i % 4 == 0 ? 1.0 : 0.0) + O[i] + O2[i]; V[i] = R[i]; }
  } else {
    const double a = sin(theta) / theta, b = (1 - cos(theta)) / (theta * theta), c = (theta - sin(theta)) / (theta * theta * theta);
    for (int i = 0; i < 9; ++i) {
      const double I = (i % 4 == 0 ? 1.0 : 0.0);
      R[i] = I + a * O[i] + b * O2[i];
      V[i] = I + b * O[i] + c * O2[i];
    }
  }
  R_to_quat(R, out->q);
  quat_normalize(out->q);
  for (int i = 0; i < 3; ++i) out->t[i] = V[i * 3] * up[0] + V[i * 3 + 1] * up[1] + V[i * 3 + 2] * up[2];
}

/* SE3Quat::operator*: r = a * b */
static void se3_mul(const Se3 *a, const Se3 *b, Se3 *r) {
  double Ra[9];
  quat_to_R(a->q, Ra);
  Se3 o;
  for (int i = 0; i < 3; ++i) o.t[i] = a->t[i] + Ra[i * 3] * b->t[0] + Ra[i * 3 + 1] * b->t[1] + Ra[i * 3 + 2] * b->t[2];
  quat_mul(a->q, b->q, o.q);
  quat_normalize(o.q);
  *r = o;
}

typedef struct {
  int F, P, E;
  const int32_t *ef, *ep;
  const float *obs;
  double f, cx, cy, info[4], huber;
  int fix_points, fix_first;
} Problem;

/* robust chi2 of the whole graph; optionally per-edge (e, rho') */
static double robust_chi2(const Problem *pb, const Se3 *poses, const double *pts) {
  double sum = 0;
  double R[16][9];
  for (int f = 0; f < pb->F; ++f) quat_to_R(poses[f].q, R[f]);
  for (int k = 0; k < pb->E; ++k) {
    const int f = pb->ef[k];
    const double *X = pts + 3 * pb->ep[k], *Rf = R[f];
    const double x = Rf[0] * X[0] + Rf[1] * X[1] + Rf[2] * X[2] + poses[f].t[0];
    const double y = Rf[3] * X[0] + Rf[4] * X[1] + Rf[5] * X[2] + poses[f].t[1];
    const double z = Rf[6] * X[0] + Rf[7] * X[1] + Rf[8] * X[2] + poses[f].t[2];
    const double e0 = pb->obs[2 * k] - (pb->f * x / z + pb->cx), e1 = pb->obs[2 * k + 1] - (pb->f * y / z + pb->cy);
    const double chi = e0 * (pb->info[0] * e0 + pb->info[1] * e1) + e1 * (pb->info[2] * e0 + pb->info[3] * e1);
    const double d2 = pb->huber * pb->huber;
    sum += pb->huber > 0 ? (chi <= d2 ? chi : 2 * sqrt(chi) * pb->huber - d2) : chi;
  }
  return sum;
}

/* dense LDLT solve (no pivoting; the damped system is SPD).  A is n x n row-major, overwritten. */
static int ldlt_solve(double *A, double *b, int n) {
  for (int j = 0; j < n; ++j) {
    double d = A[j * n + j];
    for (int k = 0; k < j; ++k) d -= A[j * n + k] * A[j * n + k] * A[k * n + k];
    if (!(fabs(d) > 0) || !isfinite(d)) return 0;
    A[j * n + j] = d;
    for (int i = j + 1; i < n; ++i) {
      double s = A[i * n + j];
      for (int k = 0; k < j; ++k) s -= A[i * n + k] * A[j * n + k] * A[k * n + k];
      A[i * n + j] = s / d;
    }
  }
  for (int i = 0; i < n; ++i) for (int k = 0; k < i; ++k) b[i] -= A[i * n + k] * b[k];
  for (int i = 0; i < n; ++i) b[i] /= A[i * n + i];
  for (int i = n - 1; i >= 0; --i) for (int k = i + 1; k < n; ++k) b[i] -= A[k * n + i] * b[k];
  return 1;
}

static int inv3(const double *m, double *o) {
  const double c0 = m[4] * m[8] - m[5] * m[7], c1 = m[5] * m[6] - m[3] * m[8], c2 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c0 + m[1] * c1 + m[2] * c2;
  if (!(fabs(det) > 0)) return 0;
  const double id = 1.0 / det;
  o[0] = c0 * id; o[1] = (m[2] * m[7] - m[1] * m[8]) * id; o[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  o[3] = c1 * id; o[4] = (m[0] * m[8] - m[2] * m[6]) * id; o[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  o[6] = c2 * id; o[7] = (m[1] * m[6] - m[0] * m[7]) * id; o[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return 1;
}

/* Optional trace of the Levenberg-Marquardt control flow (tests/test_ba_oracle.py compares it with the independent numpy
 * implementation oracle/ba_g2o_trace.py): one record of 5 doubles per TRIAL — outer iteration, lambda used by the trial,
 * robust chi2 of the trial state, gain ratio rho, accepted (1/0). */
static double *g_trace = 0;
static int g_trace_cap = 0, *g_trace_n = 0;
void orc_ba_set_trace(double *buf, int cap_records, int *count) { g_trace = buf; g_trace_cap = cap_records; g_trace_n = count; if (count) *count = 0; }

/* poses_T_w_c: F x 16 row-major camera->world (in/out); points: P x 3 float (in/out if update_points)
 * stats[4]: initial robust chi2, final robust chi2, outer iterations run, final lambda.
 * use_huber = 0 reproduces optimizeSingleFrame (no robust kernel). Returns 0 on success. */
int orc_bundle_adjustment(double *poses_T_w_c, int F, float *points, int P, const int32_t *edge_frame,
                          const int32_t *edge_point, const float *obs, int E, const double *K,
                          const double *information, int fix_points, int update_points, int iterations,
                          double huber_delta, int fix_first_pose, double *stats) {
  if (F < 1 || F > 16 || P < 0 || E < 0) return -1;
  Problem pb = {F, P, E, edge_frame, edge_point, obs, K[0], K[2], K[5], {information[0], information[1], information[2], information[3]},
                huber_delta, fix_points, fix_first_pose};
  Se3 *poses = (Se3 *)malloc(sizeof(Se3) * F), *trial = (Se3 *)malloc(sizeof(Se3) * F);
  double *pts = (double *)malloc(sizeof(double) * 3 * (P + 1)), *pts_try = (double *)malloc(sizeof(double) * 3 * (P + 1));
  /* g2o_ba.cpp:183-190: T_cw = (T_w_c)^-1 as SE3Quat */
  for (int f = 0; f < F; ++f) {
    const double *T = poses_T_w_c + 16 * f;
    double Rt[9] = {T[0], T[4], T[8], T[1], T[5], T[9], T[2], T[6], T[10]};   /* R^T */
    R_to_quat(Rt, poses[f].q);
    quat_normalize(poses[f].q);
    for (int i = 0; i < 3; ++i) poses[f].t[i] = -(Rt[i * 3] * T[3] + Rt[i * 3 + 1] * T[7] + Rt[i * 3 + 2] * T[11]);
  }
  for (int i = 0; i < 3 * P; ++i) pts[i] = points[i];

  /* index mapping: active pose blocks */
  int pose_idx[16], npose = 0;
  for (int f = 0; f < F; ++f) pose_idx[f] = (fix_first_pose && f == 0) ? -1 : npose++;
  const int n = 6 * npose;
  const int free_pts = !fix_points;
  double *Hpp = (double *)calloc((size_t)F * 36, sizeof(double)), *bp = (double *)calloc((size_t)F * 6, sizeof(double));
  double *Hll = (double *)calloc((size_t)(P + 1) * 9, sizeof(double)), *bl = (double *)calloc((size_t)(P + 1) * 3, sizeof(double));
  double *W = (double *)calloc((size_t)(E + 1) * 18, sizeof(double));      /* Hpl block per edge: 6 x 3 */
  double *S = (double *)malloc(sizeof(double) * (n * n + 1)), *rhs = (double *)malloc(sizeof(double) * (n + 1));
  double *dl = (double *)calloc((size_t)(P + 1) * 3, sizeof(double)), *Hinv = (double *)malloc(sizeof(double) * 9 * (P + 1));
  /* edges grouped by point (for the Schur products) */
  int *pstart = (int *)calloc((size_t)P + 2, sizeof(int)), *pedge = (int *)malloc(sizeof(int) * (E + 1));
  for (int k = 0; k < E; ++k) pstart[edge_point[k] + 1]++;
  for (int p = 0; p < P; ++p) pstart[p + 1] += pstart[p];
  { int *cur = (int *)malloc(sizeof(int) * (P + 1)); memcpy(cur, pstart, sizeof(int) * (P + 1));
    for (int k = 0; k < E; ++k) pedge[cur[edge_point[k]]++] = k; free(cur); }

  double lambda = 0, ni = 2;
  double chi_init = robust_chi2(&pb, poses, pts), current_chi = chi_init;
  int it = 0;
  for (; it < iterations; ++it) {
    /* computeActiveErrors + buildSystem */
    current_chi = robust_chi2(&pb, poses, pts);
    memset(Hpp, 0, sizeof(double) * F * 36); memset(bp, 0, sizeof(double) * F * 6);
    memset(Hll, 0, sizeof(double) * P * 9);  memset(bl, 0, sizeof(double) * P * 3);
    double R[16][9];
    for (int f = 0; f < F; ++f) quat_to_R(poses[f].q, R[f]);
    for (int k = 0; k < E; ++k) {
      const int f = edge_frame[k], l = edge_point[k];
      const double *X = pts + 3 * l, *Rf = R[f];
      const double x = Rf[0] * X[0] + Rf[1] * X[1] + Rf[2] * X[2] + poses[f].t[0];
      const double y = Rf[3] * X[0] + Rf[4] * X[1] + Rf[5] * X[2] + poses[f].t[1];
      const double z = Rf[6] * X[0] + Rf[7] * X[1] + Rf[8] * X[2] + poses[f].t[2];
      const double z2 = z * z, fl = pb.f;
      const double e[2] = {obs[2 * k] - (fl * x / z + pb.cx), obs[2 * k + 1] - (fl * y / z + pb.cy)};
      /* EdgeProjectXYZ2UV::linearizeOplus */
      const double tmp[6] = {fl, 0, -x / z * fl, 0, fl, -y / z * fl};
      double A[6];   /* 2x3 wrt point: -1/z * tmp * R */
      for (int r = 0; r < 2; ++r)
        for (int c = 0; c < 3; ++c) A[r * 3 + c] = -1.0 / z * (tmp[r * 3] * Rf[c] + tmp[r * 3 + 1] * Rf[3 + c] + tmp[r * 3 + 2] * Rf[6 + c]);
      const double B[12] = {x * y / z2 * fl, -(1 + (x * x / z2)) * fl, y / z * fl, -1.0 / z * fl, 0, x / z2 * fl,
                            (1 + y * y / z2) * fl, -x * y / z2 * fl, -x / z * fl, 0, -1.0 / z * fl, y / z2 * fl};
      /* robustify */
      const double Oe[2] = {pb.info[0] * e[0] + pb.info[1] * e[1], pb.info[2] * e[0] + pb.info[3] * e[1]};
      const double chi = e[0] * Oe[0] + e[1] * Oe[1];
      double w = 1.0;
      if (pb.huber > 0 && chi > pb.huber * pb.huber) w = pb.huber / sqrt(chi);
      const double Om[4] = {w * pb.info[0], w * pb.info[1], w * pb.info[2], w * pb.info[3]};
      const double om_r[2] = {-w * Oe[0], -w * Oe[1]};
      const int pose_active = pose_idx[f] >= 0;
      if (free_pts) {
        double OA[6];
        for (int c = 0; c < 3; ++c) { OA[c] = Om[0] * A[c] + Om[1] * A[3 + c]; OA[3 + c] = Om[2] * A[c] + Om[3] * A[3 + c]; }
        for (int r = 0; r < 3; ++r) {
          bl[3 * l + r] += A[r] * om_r[0] + A[3 + r] * om_r[1];
          for (int c = 0; c < 3; ++c) Hll[9 * l + r * 3 + c] += A[r] * OA[c] + A[3 + r] * OA[3 + c];
        }
        if (pose_active)
          for (int r = 0; r < 6; ++r)
            for (int c = 0; c < 3; ++c) W[18 * k + r * 3 + c] = B[r] * OA[c] + B[6 + r] * OA[3 + c];
      }
      if (pose_active) {
        double OB[12];
        for (int c = 0; c < 6; ++c) { OB[c] = Om[0] * B[c] + Om[1] * B[6 + c]; OB[6 + c] = Om[2] * B[c] + Om[3] * B[6 + c]; }
        for (int r = 0; r < 6; ++r) {
          bp[6 * f + r] += B[r] * om_r[0] + B[6 + r] * om_r[1];
          for (int c = 0; c < 6; ++c) Hpp[36 * f + r * 6 + c] += B[r] * OB[c] + B[6 + r] * OB[6 + c];
        }
      }
    }
    if (it == 0) {   /* computeLambdaInit: tau * max diagonal over all non-fixed vertices */
      double md = 0;
      for (int f = 0; f < F; ++f) if (pose_idx[f] >= 0) for (int j = 0; j < 6; ++j) md = fmax(md, fabs(Hpp[36 * f + j * 7]));
      if (free_pts) for (int l = 0; l < P; ++l) if (pstart[l + 1] > pstart[l]) for (int j = 0; j < 3; ++j) md = fmax(md, fabs(Hll[9 * l + j * 4]));
      lambda = 1e-5 * md;
      ni = 2;
    }
    double rho = 0;
    int qmax = 0;
    do {
      /* solve (H + lambda I) dx = b with the Schur complement on the points */
      memset(S, 0, sizeof(double) * n * n);
      for (int f = 0; f < F; ++f) {
        const int a = pose_idx[f];
        if (a < 0) continue;
        for (int r = 0; r < 6; ++r) {
          rhs[6 * a + r] = bp[6 * f + r];
          for (int c = 0; c < 6; ++c) S[(6 * a + r) * n + 6 * a + c] = Hpp[36 * f + r * 6 + c] + (r == c ? lambda : 0);
        }
      }
      int ok = 1;
      if (free_pts) {
        for (int l = 0; l < P && ok; ++l) {
          if (pstart[l + 1] == pstart[l]) continue;
          double Hd[9];
          memcpy(Hd, Hll + 9 * l, sizeof Hd);
          Hd[0] += lambda; Hd[4] += lambda; Hd[8] += lambda;
          if (!inv3(Hd, Hinv + 9 * l)) { ok = 0; break; }
          const double *Hi = Hinv + 9 * l;
          for (int ia = pstart[l]; ia < pstart[l + 1]; ++ia) {
            const int ka = pedge[ia], a = pose_idx[edge_frame[ka]];
            if (a < 0) continue;
            double Y[18];   /* W_a Hinv : 6x3 */
            for (int r = 0; r < 6; ++r)
              for (int c = 0; c < 3; ++c) Y[r * 3 + c] = W[18 * ka + r * 3] * Hi[c] + W[18 * ka + r * 3 + 1] * Hi[3 + c] + W[18 * ka + r * 3 + 2] * Hi[6 + c];
            for (int r = 0; r < 6; ++r) rhs[6 * a + r] -= Y[r * 3] * bl[3 * l] + Y[r * 3 + 1] * bl[3 * l + 1] + Y[r * 3 + 2] * bl[3 * l + 2];
            for (int ib = pstart[l]; ib < pstart[l + 1]; ++ib) {
              const int kb = pedge[ib], b = pose_idx[edge_frame[kb]];
              if (b < 0) continue;
              for (int r = 0; r < 6; ++r)
                for (int c = 0; c < 6; ++c)
                  S[(6 * a + r) * n + 6 * b + c] -= Y[r * 3] * W[18 * kb + c * 3] + Y[r * 3 + 1] * W[18 * kb + c * 3 + 1] + Y[r * 3 + 2] * W[18 * kb + c * 3 + 2];
            }
          }
        }
      }
      if (ok && n > 0) ok = ldlt_solve(S, rhs, n);
      double scale = 0;
      memcpy(trial, poses, sizeof(Se3) * F);
      memcpy(pts_try, pts, sizeof(double) * 3 * P);
      if (ok) {
        for (int f = 0; f < F; ++f) {
          const int a = pose_idx[f];
          if (a < 0) continue;
          Se3 d;
          se3_exp(rhs + 6 * a, &d);
          se3_mul(&d, &poses[f], &trial[f]);
          for (int r = 0; r < 6; ++r) scale += rhs[6 * a + r] * (lambda * rhs[6 * a + r] + bp[6 * f + r]);
        }
        if (free_pts)
          for (int l = 0; l < P; ++l) {
            if (pstart[l + 1] == pstart[l]) continue;
            double c[3] = {bl[3 * l], bl[3 * l + 1], bl[3 * l + 2]};
            for (int ia = pstart[l]; ia < pstart[l + 1]; ++ia) {
              const int ka = pedge[ia], a = pose_idx[edge_frame[ka]];
              if (a < 0) continue;
              for (int r = 0; r < 3; ++r)
                for (int q = 0; q < 6; ++q) c[r] -= W[18 * ka + q * 3 + r] * rhs[6 * a + q];
            }
            const double *Hi = Hinv + 9 * l;
            for (int r = 0; r < 3; ++r) {
              dl[3 * l + r] = Hi[r * 3] * c[0] + Hi[r * 3 + 1] * c[1] + Hi[r * 3 + 2] * c[2];
              pts_try[3 * l + r] = pts[3 * l + r] + dl[3 * l + r];
              scale += dl[3 * l + r] * (lambda * dl[3 * l + r] + bl[3 * l + r]);
            }
          }
      }
      double temp_chi = ok ? robust_chi2(&pb, trial, pts_try) : 1.7976931348623157e308;
      rho = (current_chi - temp_chi) / (scale + 1e-3);
      if (g_trace && g_trace_n && *g_trace_n < g_trace_cap) {
        double *rec = g_trace + 5 * (*g_trace_n)++;
        rec[0] = it; rec[1] = lambda; rec[2] = temp_chi; rec[3] = rho; rec[4] = (rho > 0 && isfinite(temp_chi)) ? 1 : 0;
      }
      if (rho > 0 && isfinite(temp_chi)) {
        double alpha = 1. - pow(2 * rho - 1, 3);
        alpha = fmin(alpha, 2. / 3.);
        lambda *= fmax(1. / 3., alpha);
        ni = 2;
        current_chi = temp_chi;
        memcpy(poses, trial, sizeof(Se3) * F);
        memcpy(pts, pts_try, sizeof(double) * 3 * P);
      } else {
        lambda *= ni;
        ni *= 2;
      }
      qmax++;
    } while (rho < 0 && qmax < 10);
    if (qmax == 10 || rho == 0) { ++it; break; }   /* Terminate */
  }
  /* g2o_ba.cpp:298-316: write back */
  for (int f = 0; f < F; ++f) {
    double R[9];
    quat_to_R(poses[f].q, R);
    double *T = poses_T_w_c + 16 * f;
    for (int i = 0; i < 3; ++i) {
      for (int j = 0; j < 3; ++j) T[i * 4 + j] = R[j * 3 + i];
      T[i * 4 + 3] = -(R[i] * poses[f].t[0] + R[3 + i] * poses[f].t[1] + R[6 + i] * poses[f].t[2]);
    }
    T[12] = T[13] = T[14] = 0; T[15] = 1;
  }
  if (update_points) for (int i = 0; i < 3 * P; ++i) points[i] = (float)pts[i];
  if (stats) { stats[0] = chi_init; stats[1] = current_chi; stats[2] = it; stats[3] = lambda; }
  free(poses); free(trial); free(pts); free(pts_try); free(Hpp); free(bp); free(Hll); free(bl); free(W);
  free(S); free(rhs); free(dl); free(Hinv); free(pstart); free(pedge);
  return 0;
}

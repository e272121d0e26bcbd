// Small dense numerics of the two-view stage, written once for host and device: the CUDA kernels in
// epipolar.cu call these from one thread per hypothesis / per point, and tests/cpp/epipolar_math_host.cpp compiles the
// very same functions with g++ so that they are checked against numpy / cv2 on the CPU-only test tier.
//   sym_eigen_jacobi<N>      cyclic Jacobi eigen-decomposition of a symmetric NxN matrix (N = 3, 4)
//   null_vector_8x9          the 1-dimensional null space of an 8x9 system by Gaussian elimination with complete pivoting
//   svd3                     3x3 SVD from the eigen-decomposition of M^T M
//   essential_from_8         eight-point essential matrix on calibrated coordinates, projected onto the essential manifold
//   sampson_err              the squared Sampson distance OpenCV's findEssentialMat scores with (five-point.cpp)
//   decompose_essential      R1, R2, t of cv::decomposeEssentialMat
//   triangulate_dlt          one point of cv::triangulatePoints (the 4x4 homogeneous DLT, triangulate.cpp)
// fp64 throughout, no dynamic memory, fixed loop bounds.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define EPI_HD __host__ __device__ __forceinline__
// the three routines with nested loops over local arrays are real calls on the device: inlining them into each other
// (essential_from_8 -> null_vector_8x9 / svd3 -> sym_eigen_jacobi) bloats every caller and, with nvcc 12.9 for sm_100a,
// produced wrong singular values inside essential_from_8 although each routine is correct on its own
// (tests/test_epipolar_gpu.py::test_device_math_equals_host_build)
#define EPI_HD_CALL __host__ __device__ __noinline__
#else
#define EPI_HD inline
#define EPI_HD_CALL inline
#endif

namespace epi {

// A (symmetric, row-major, destroyed) -> eigenvalues w[N] (unsorted) and eigenvectors as COLUMNS of V
template <int N>
EPI_HD_CALL void sym_eigen_jacobi(double *A, double *w, double *V) {
  for (int i = 0; i < N; ++i)
    for (int j = 0; j < N; ++j) V[i * N + j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 12; ++sweep) {
    double off = 0;
    for (int i = 0; i < N; ++i)
      for (int j = i + 1; j < N; ++j) off += A[i * N + j] * A[i * N + j];
    double diag = 0;
    for (int i = 0; i < N; ++i) diag += A[i * N + i] * A[i * N + i];
    if (off <= 1e-32 * diag || off == 0) break;
    for (int p = 0; p < N; ++p)
      for (int q = p + 1; q < N; ++q) {
        const double apq = A[p * N + q];
        if (apq == 0) continue;
        const double theta = (A[q * N + q] - A[p * N + p]) / (2 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1));
        const double c = 1 / sqrt(t * t + 1), s = t * c;
        for (int k = 0; k < N; ++k) {               // columns p, q of A
          const double akp = A[k * N + p], akq = A[k * N + q];
          A[k * N + p] = c * akp - s * akq;
          A[k * N + q] = s * akp + c * akq;
        }
        for (int k = 0; k < N; ++k) {               // rows p, q of A
          const double apk = A[p * N + k], aqk = A[q * N + k];
          A[p * N + k] = c * apk - s * aqk;
          A[q * N + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < N; ++k) {
          const double vkp = V[k * N + p], vkq = V[k * N + q];
          V[k * N + p] = c * vkp - s * vkq;
          V[k * N + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < N; ++i) w[i] = A[i * N + i];
}

// Null vector x (unit norm) of the 8x9 system M x = 0 (M row-major, destroyed).  Returns false when the system is
// rank deficient beyond one dimension (degenerate sample).
EPI_HD_CALL bool null_vector_8x9(double *M, double *x) {
  int perm[9];
  for (int j = 0; j < 9; ++j) perm[j] = j;
  for (int k = 0; k < 8; ++k) {
    int pr = k, pc = k;                              // complete pivoting
    double best = 0;
    for (int i = k; i < 8; ++i)
      for (int j = k; j < 9; ++j) {
        const double a = fabs(M[i * 9 + j]);
        if (a > best) { best = a; pr = i; pc = j; }
      }
    if (!(best > 1e-12)) return false;
    if (pr != k) for (int j = 0; j < 9; ++j) { const double t = M[k * 9 + j]; M[k * 9 + j] = M[pr * 9 + j]; M[pr * 9 + j] = t; }
    if (pc != k) {
      for (int i = 0; i < 8; ++i) { const double t = M[i * 9 + k]; M[i * 9 + k] = M[i * 9 + pc]; M[i * 9 + pc] = t; }
      const int t = perm[k]; perm[k] = perm[pc]; perm[pc] = t;
    }
    const double inv = 1.0 / M[k * 9 + k];
    for (int i = k + 1; i < 8; ++i) {
      const double f = M[i * 9 + k] * inv;
      if (f == 0) continue;
      for (int j = k; j < 9; ++j) M[i * 9 + j] -= f * M[k * 9 + j];
    }
  }
  double y[9];
  y[8] = 1.0;                                        // the free variable
  for (int k = 7; k >= 0; --k) {
    double s = 0;
    for (int j = k + 1; j < 9; ++j) s += M[k * 9 + j] * y[j];
    y[k] = -s / M[k * 9 + k];
  }
  double nrm = 0;
  for (int j = 0; j < 9; ++j) nrm += y[j] * y[j];
  nrm = sqrt(nrm);
  if (!(nrm > 0) || !(nrm - nrm == 0.0)) return false;
  for (int j = 0; j < 9; ++j) x[perm[j]] = y[j] / nrm;
  return true;
}

EPI_HD void cross3(const double *a, const double *b, double *c) {
  c[0] = a[1] * b[2] - a[2] * b[1]; c[1] = a[2] * b[0] - a[0] * b[2]; c[2] = a[0] * b[1] - a[1] * b[0];
}
EPI_HD double det3(const double *M) {
  return M[0] * (M[4] * M[8] - M[5] * M[7]) - M[1] * (M[3] * M[8] - M[5] * M[6]) + M[2] * (M[3] * M[7] - M[4] * M[6]);
}

// M = U diag(s) V^T, s sorted descending, U and V orthogonal (columns), row-major 3x3.
EPI_HD_CALL void svd3(const double *M, double *U, double *s, double *V) {
  double A[9], w[3], Q[9];
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) A[i * 3 + j] = M[0 * 3 + i] * M[0 * 3 + j] + M[1 * 3 + i] * M[1 * 3 + j] + M[2 * 3 + i] * M[2 * 3 + j];
  sym_eigen_jacobi<3>(A, w, Q);
  int o[3] = {0, 1, 2};                               // order by eigenvalue, descending
  for (int a = 0; a < 3; ++a)
    for (int b = a + 1; b < 3; ++b)
      if (w[o[b]] > w[o[a]]) { const int t = o[a]; o[a] = o[b]; o[b] = t; }
  for (int k = 0; k < 3; ++k) {
    s[k] = sqrt(w[o[k]] > 0 ? w[o[k]] : 0.0);
    for (int i = 0; i < 3; ++i) V[i * 3 + k] = Q[i * 3 + o[k]];
  }
  // U columns: M v_k / s_k for the two leading directions, the third completes a right-handed frame
  double u[3][3];
  for (int k = 0; k < 2; ++k) {
    double n = 0;
    for (int i = 0; i < 3; ++i) { u[k][i] = M[i * 3] * V[0 * 3 + k] + M[i * 3 + 1] * V[1 * 3 + k] + M[i * 3 + 2] * V[2 * 3 + k]; n += u[k][i] * u[k][i]; }
    n = sqrt(n);
    if (n > 0) for (int i = 0; i < 3; ++i) u[k][i] /= n;
  }
  {                                                   // re-orthogonalise u1 against u0 (nearly equal singular values)
    const double d = u[0][0] * u[1][0] + u[0][1] * u[1][1] + u[0][2] * u[1][2];
    double n = 0;
    for (int i = 0; i < 3; ++i) { u[1][i] -= d * u[0][i]; n += u[1][i] * u[1][i]; }
    n = sqrt(n);
    if (n > 0) for (int i = 0; i < 3; ++i) u[1][i] /= n;
  }
  cross3(u[0], u[1], u[2]);
  for (int k = 0; k < 3; ++k)
    for (int i = 0; i < 3; ++i) U[i * 3 + k] = u[k][i];
  // keep M = U S V^T for the third direction: flip v2 if needed (s2 >= 0 by construction, the sign lives in V)
  double mv[3];
  for (int i = 0; i < 3; ++i) mv[i] = M[i * 3] * V[0 * 3 + 2] + M[i * 3 + 1] * V[1 * 3 + 2] + M[i * 3 + 2] * V[2 * 3 + 2];
  if (mv[0] * u[2][0] + mv[1] * u[2][1] + mv[2] * u[2][2] < 0)
    for (int i = 0; i < 3; ++i) V[i * 3 + 2] = -V[i * 3 + 2];
}

// Essential matrix from 8 correspondences in calibrated coordinates (x1 -> x2, x2^T E x1 = 0), projected onto the
// essential manifold (singular values 1, 1, 0).  xy1 / xy2: 8 x 2.  Returns false for degenerate samples.
EPI_HD bool finite_d(double v) { return v - v == 0.0; }      // false for NaN and +-inf, no library call

// reason (optional): 0 ok, 1 singular sample, 2 rank-one null vector, 3 non-finite result
EPI_HD_CALL bool essential_from_8(const double *xy1, const double *xy2, double *E, int *reason = nullptr) {
  if (reason) *reason = 0;
  double M[72];
  for (int k = 0; k < 8; ++k) {
    const double x1 = xy1[2 * k], y1 = xy1[2 * k + 1], x2 = xy2[2 * k], y2 = xy2[2 * k + 1];
    double *r = M + 9 * k;
    r[0] = x2 * x1; r[1] = x2 * y1; r[2] = x2; r[3] = y2 * x1; r[4] = y2 * y1; r[5] = y2; r[6] = x1; r[7] = y1; r[8] = 1.0;
  }
  double f[9];
  if (!null_vector_8x9(M, f)) { if (reason) *reason = 1; return false; }
  double U[9], s[3], V[9];
  svd3(f, U, s, V);
  if (!(s[1] > 1e-9 * s[0])) { if (reason) *reason = 2; return false; }            // rank one: not an essential matrix
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) E[i * 3 + j] = U[i * 3] * V[j * 3] + U[i * 3 + 1] * V[j * 3 + 1];      // U diag(1,1,0) V^T
  for (int i = 0; i < 9; ++i)
    if (!finite_d(E[i])) { if (reason) *reason = 3; return false; }
  return true;
}

// Squared Sampson distance of (x1, y1) <-> (x2, y2) under E (calibrated coordinates)
EPI_HD double sampson_err(const double *E, double x1, double y1, double x2, double y2) {
  const double Ex0 = E[0] * x1 + E[1] * y1 + E[2], Ex1 = E[3] * x1 + E[4] * y1 + E[5], Ex2 = E[6] * x1 + E[7] * y1 + E[8];
  const double Et0 = E[0] * x2 + E[3] * y2 + E[6], Et1 = E[1] * x2 + E[4] * y2 + E[7];
  const double x2tEx1 = x2 * Ex0 + y2 * Ex1 + Ex2;
  return x2tEx1 * x2tEx1 / (Ex0 * Ex0 + Ex1 * Ex1 + Et0 * Et0 + Et1 * Et1);
}

// cv::decomposeEssentialMat: E = U diag(1,1,0) V^T, det(U), det(V) forced positive, W = [0 1 0; -1 0 0; 0 0 1],
// R1 = U W V^T, R2 = U W^T V^T, t = U[:, 2]
EPI_HD_CALL void decompose_essential(const double *E, double *R1, double *R2, double *t) {
  double U[9], s[3], V[9];
  svd3(E, U, s, V);
  if (det3(U) < 0) for (int i = 0; i < 9; ++i) U[i] = -U[i];
  if (det3(V) < 0) for (int i = 0; i < 9; ++i) V[i] = -V[i];
  // U W = [-u1, u0, u2] (columns), U W^T = [u1, -u0, u2]
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) {
      const double a = -U[i * 3 + 1] * V[j * 3 + 0] + U[i * 3 + 0] * V[j * 3 + 1] + U[i * 3 + 2] * V[j * 3 + 2];
      const double b = U[i * 3 + 1] * V[j * 3 + 0] - U[i * 3 + 0] * V[j * 3 + 1] + U[i * 3 + 2] * V[j * 3 + 2];
      R1[i * 3 + j] = a;
      R2[i * 3 + j] = b;
    }
  for (int i = 0; i < 3; ++i) t[i] = U[i * 3 + 2];
}

// One point of cv::triangulatePoints: P1, P2 3x4 row-major, (x1, y1), (x2, y2) in the units of the projection
// matrices; X = right singular vector of the 4x4 DLT system for its smallest singular value (homogeneous, unit norm).
EPI_HD_CALL void triangulate_dlt(const double *P1, const double *P2, double x1, double y1, double x2, double y2, double *X) {
  double A[16];
  for (int j = 0; j < 4; ++j) {
    A[0 * 4 + j] = x1 * P1[2 * 4 + j] - P1[0 * 4 + j];
    A[1 * 4 + j] = y1 * P1[2 * 4 + j] - P1[1 * 4 + j];
    A[2 * 4 + j] = x2 * P2[2 * 4 + j] - P2[0 * 4 + j];
    A[3 * 4 + j] = y2 * P2[2 * 4 + j] - P2[1 * 4 + j];
  }
  double AtA[16], w[4], V[16];
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) AtA[i * 4 + j] = A[0 * 4 + i] * A[0 * 4 + j] + A[1 * 4 + i] * A[1 * 4 + j] + A[2 * 4 + i] * A[2 * 4 + j] + A[3 * 4 + i] * A[3 * 4 + j];
  sym_eigen_jacobi<4>(AtA, w, V);
  int m = 0;
  for (int k = 1; k < 4; ++k) if (w[k] < w[m]) m = k;
  for (int i = 0; i < 4; ++i) X[i] = V[i * 4 + m];
}

}  // namespace epi

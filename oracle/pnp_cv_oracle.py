"""TEST ORACLE — restatement of cv::solvePnPRansac's CONTROL FLOW as the reference calls it
(src/vo/vo.cpp:314-320: useExtrinsicGuess=false, 100 iterations, 2.0 px, confidence 0.999, SOLVEPNP_ITERATIVE).  Not product code.

OpenCV 4.13 (calib3d solvepnp.cpp / ptsetreg.cpp / epnp.cpp; an absent third-party dependency, pinned here against cv2 live):
  RANSACPointSetRegistrator::run   RNG rng((uint64)-1); <= maxIters iterations; per iteration getSubset (5 distinct
                                   indices from rng.uniform(0, count)), runKernel = solvePnP(EPNP) on the 5 points,
                                   findInliers (float reprojection error <= (float)(thr*thr)), keep the model with the
                                   most inliers, niters = RANSACUpdateNumIters(confidence, outlier ratio, 5, niters)
  epnp::compute_pose               control points from a PCA of the 5 points, barycentric coordinates, M (10 x 12),
                                   SVD of M^T M, the betas of the N = 1, 2, 3 approximations + 5 Gauss-Newton steps,
                                   the candidate with the smallest reprojection error
  final pose                       solvePnP(ITERATIVE, useExtrinsicGuess = the RANSAC model) over the consensus set

FINDING (tests/test_pnp_oracle.py::test_cv_ransac_minimal_solver_is_decided_by_rounding_noise): M is 10 x 12, so M^T M
has an exactly two-dimensional null space, and epnp takes v[0], v[1] = the last two rows of U^T from cv::SVD's one-sided
Jacobi.  Which orthonormal basis of that plane comes out is decided by the rounding noise of the Jacobi sweeps (the two
singular values are ~1e-10 against 1e6), and epnp's 5 Gauss-Newton steps do not converge to a basis-independent optimum:
the same restatement evaluated with np.linalg.svd instead of cv2.SVDecomp moves the 5-point pose by 1e-3..1e-1 on most
samples, and even with cv2.SVDecomp itself a last-bit difference in M^T M (numpy's summation order instead of
cvMulTransposed's) changes a quarter of the poses.  cv::solvePnPRansac's consensus set is therefore reproducible only by
the same binary on the same inputs: no independent implementation (CPU or GPU) can be inlier-index-exact with it, and the
reference's own trajectory is defined only up to this noise.  What IS pinned: the sampler (same subsets, checked through the
identical final result whenever the SVD noise does not interfere), the float scoring rule and the adaptive iteration count.
"""
from __future__ import annotations

import numpy as np

CV_RNG_COEFF = 4164903690


class CvRng:
    """cv::RNG (multiply-with-carry), modules/core/include/opencv2/core/operations.hpp."""

    def __init__(self, state=0xFFFFFFFFFFFFFFFF):
        self.state = state if state else 0xFFFFFFFF

    def next(self):
        self.state = ((self.state & 0xFFFFFFFF) * CV_RNG_COEFF + (self.state >> 32)) & 0xFFFFFFFFFFFFFFFF
        return self.state & 0xFFFFFFFF

    def uniform(self, a, b):
        return a if a == b else int(self.next() % (b - a) + a)


def get_subset(rng, count, model_points):
    """RANSACPointSetRegistrator::getSubset (ptsetreg.cpp): model_points distinct indices, redrawn on collision."""
    idx = []
    while len(idx) < model_points:
        while True:
            i = rng.uniform(0, count)
            if i not in idx:
                break
        idx.append(i)
    return idx


def ransac_update_num_iters(p, ep, model_points, max_iters):
    p = min(max(p, 0.0), 1.0)
    ep = min(max(ep, 0.0), 1.0)
    num = max(1.0 - p, np.finfo(np.float64).tiny)
    denom = 1.0 - (1.0 - ep) ** model_points
    if denom < np.finfo(np.float64).tiny:
        return 0
    num, denom = np.log(num), np.log(denom)
    if denom >= 0 or -num >= max_iters * (-denom):
        return max_iters
    return int(np.rint(num / denom))                       # cvRound: round half to even


def _np_svd(A):
    U, d, Vt = np.linalg.svd(A)
    return U, d, Vt


def jacobi_svd(A):
    """cv::SVD::compute on a square matrix without LAPACK (below 25 rows OpenCV's HAL declines LAPACK): the one-sided Jacobi
    of modules/core/src/lapack.cpp JacobiSVDImpl_<double> on the rows of A^T — cyclic (i, j) sweeps, rotation skipped when
    |p| <= 10 eps sqrt(a b), at most max(m, 30) sweeps, singular values sorted descending by selection, rows scaled by
    1/sigma.  Same algorithm as cv2.SVDecomp, NOT bit-identical with it (summation order / hypot): where the basis of a
    degenerate subspace is decided by rounding noise the two differ (module docstring).  The device kernel restates this."""
    import math
    A = np.array(A, np.float64)
    n = A.shape[0]
    At = A.T.copy()
    Vt = np.eye(n)
    W = (At * At).sum(1)
    eps = np.finfo(np.float64).eps * 10
    for _ in range(max(n, 30)):
        changed = False
        for i in range(n - 1):
            for j in range(i + 1, n):
                a, b = W[i], W[j]
                p = float(At[i] @ At[j])
                if abs(p) <= eps * math.sqrt(a * b):
                    continue
                p *= 2
                beta = a - b
                gamma = math.hypot(p, beta)
                if beta < 0:
                    s = math.sqrt((gamma - beta) * 0.5 / gamma)
                    c = p / (gamma * s * 2)
                else:
                    c = math.sqrt((gamma + beta) / (gamma * 2))
                    s = p / (gamma * c * 2)
                t0, t1 = c * At[i] + s * At[j], -s * At[i] + c * At[j]
                At[i], At[j] = t0, t1
                W[i], W[j] = float(t0 @ t0), float(t1 @ t1)
                Vt[i], Vt[j] = c * Vt[i] + s * Vt[j], -s * Vt[i] + c * Vt[j]
                changed = True
        if not changed:
            break
    W = np.sqrt((At * At).sum(1))
    for i in range(n - 1):
        j = i + int(np.argmax(W[i:]))
        if W[j] > W[i]:
            W[[i, j]] = W[[j, i]]
            At[[i, j]] = At[[j, i]]
            Vt[[i, j]] = Vt[[j, i]]
    # rows with a zero singular value: OpenCV fills them with pseudo-random +-1/m vectors (RNG 0x12345678), orthogonalised
    # against the rows before them (two Gram-Schmidt rounds with an L1 renormalisation), then normalised
    tiny = np.finfo(np.float64).tiny
    rng = CvRng(0x12345678)
    m = n
    for i in range(n):
        sd = W[i]
        for _ in range(100):
            if sd > tiny:
                break
            val0 = 1.0 / m
            At[i] = [val0 if (rng.next() & 256) != 0 else -val0 for _k in range(m)]
            for _it in range(2):
                for j in range(i):
                    dot = float(At[i] @ At[j])
                    At[i] = At[i] - dot * At[j]
                    asum = float(np.abs(At[i]).sum())
                    At[i] = At[i] * ((1.0 / asum) if asum > eps * 100 else 0.0)
            sd = math.sqrt(float(At[i] @ At[i]))
        At[i] *= (1.0 / sd) if sd > tiny else 0.0
    return At.T.copy(), W, Vt


def cv_svd(A):
    import cv2
    w, u, vt = cv2.SVDecomp(np.ascontiguousarray(A, np.float64))
    return u, w.ravel(), vt


def epnp(P, us, fu, fv, uc, vc, svd=None):
    """epnp::compute_pose for n points: P n x 3 (double), us n x 2 pixel coordinates as epnp::init_points builds them
    ((float)undistorted * f + c).  Returns (R, t, reprojection errors of the three candidates, chosen index)."""
    svd = svd or cv_svd
    n = len(P)
    c0 = P.sum(0) / n
    PW0 = P - c0
    U, dc, _ = svd(PW0.T @ PW0)
    uct = U.T
    cws = np.zeros((4, 3))
    cws[0] = c0
    for i in range(1, 4):
        cws[i] = c0 + np.sqrt(dc[i - 1] / n) * uct[i - 1]
    # cvInvert(&CC, &CC_inv, CV_SVD): pseudo-inverse, singular values <= 2 eps * sum(w) dropped (coplanar points)
    Uc, wc, Vtc = svd((cws[1:] - cws[0]).T)
    thr = np.finfo(np.float64).eps * 2 * wc.sum()
    CCi = np.zeros((3, 3))
    for k in range(3):
        if abs(wc[k]) > thr:
            CCi += np.outer(Vtc[k], Uc[:, k]) / wc[k]
    al = np.zeros((n, 4))
    al[:, 1:] = (CCi @ (P - c0).T).T
    al[:, 0] = 1.0 - al[:, 1] - al[:, 2] - al[:, 3]
    M = np.zeros((2 * n, 12))
    for j in range(4):
        M[0::2, 3 * j] = al[:, j] * fu
        M[0::2, 3 * j + 2] = al[:, j] * (uc - us[:, 0])
        M[1::2, 3 * j + 1] = al[:, j] * fv
        M[1::2, 3 * j + 2] = al[:, j] * (vc - us[:, 1])
    U, _, _ = svd(M.T @ M)
    ut = U.T
    v = [ut[11], ut[10], ut[9], ut[8]]
    dv = np.zeros((4, 6, 3))
    for i in range(4):
        a, b = 0, 1
        for j in range(6):
            dv[i, j] = v[i][3 * a:3 * a + 3] - v[i][3 * b:3 * b + 3]
            b += 1
            if b > 3:
                a += 1
                b = a + 1
    L = np.zeros((6, 10))
    for i in range(6):
        d0, d1, d2, d3 = dv[0, i], dv[1, i], dv[2, i], dv[3, i]
        L[i] = [d0 @ d0, 2 * d0 @ d1, d1 @ d1, 2 * d0 @ d2, 2 * d1 @ d2, d2 @ d2, 2 * d0 @ d3, 2 * d1 @ d3, 2 * d2 @ d3, d3 @ d3]
    pairs = [(0, 1), (0, 2), (0, 3), (1, 2), (1, 3), (2, 3)]
    rho = np.array([((cws[a] - cws[b]) ** 2).sum() for a, b in pairs])

    def lstsq(A, b):
        return np.linalg.lstsq(A, b, rcond=None)[0]

    def approx(cols, kind):
        bb = lstsq(L[:, cols], rho)
        be = np.zeros(4)
        if kind == 1:
            if bb[0] < 0:
                be[0] = np.sqrt(-bb[0]); be[1:] = -bb[1:] / be[0]
            else:
                be[0] = np.sqrt(bb[0]); be[1:] = bb[1:] / be[0]
            return be
        if bb[0] < 0:
            be[0] = np.sqrt(-bb[0]); be[1] = np.sqrt(-bb[2]) if bb[2] < 0 else 0.0
        else:
            be[0] = np.sqrt(bb[0]); be[1] = np.sqrt(bb[2]) if bb[2] > 0 else 0.0
        if bb[1] < 0:
            be[0] = -be[0]
        if kind == 3:
            be[2] = bb[3] / be[0]
        return be

    def gauss_newton(be):
        be = be.copy()
        for _ in range(5):
            A = np.zeros((6, 4))
            b = np.zeros(6)
            for i in range(6):
                r = L[i]
                A[i] = [2 * r[0] * be[0] + r[1] * be[1] + r[3] * be[2] + r[6] * be[3],
                        r[1] * be[0] + 2 * r[2] * be[1] + r[4] * be[2] + r[7] * be[3],
                        r[3] * be[0] + r[4] * be[1] + 2 * r[5] * be[2] + r[8] * be[3],
                        r[6] * be[0] + r[7] * be[1] + r[8] * be[2] + 2 * r[9] * be[3]]
                b[i] = rho[i] - (r[0] * be[0] * be[0] + r[1] * be[0] * be[1] + r[2] * be[1] * be[1] + r[3] * be[0] * be[2] + r[4] * be[1] * be[2] +
                                 r[5] * be[2] * be[2] + r[6] * be[0] * be[3] + r[7] * be[1] * be[3] + r[8] * be[2] * be[3] + r[9] * be[3] * be[3])
            be += lstsq(A, b)
        return be

    def r_and_t(be):
        ccs = np.zeros((4, 3))
        for i in range(4):
            for j in range(4):
                ccs[j] += be[i] * ut[11 - i][3 * j:3 * j + 3]
        pcs = al @ ccs
        if pcs[0, 2] < 0:
            ccs, pcs = -ccs, -pcs
        pc0, pw0 = pcs.sum(0) / n, P.sum(0) / n
        Ua, _, Vta = np.linalg.svd((pcs - pc0).T @ (P - pw0))
        R = Ua @ Vta
        if np.linalg.det(R) < 0:
            R[2] = -R[2]
        t = pc0 - R @ pw0
        Xc = P @ R.T + t
        with np.errstate(divide="ignore", invalid="ignore"):
            ue = uc + fu * Xc[:, 0] / Xc[:, 2]
            ve = vc + fv * Xc[:, 1] / Xc[:, 2]
        return R, t, np.sqrt((us[:, 0] - ue) ** 2 + (us[:, 1] - ve) ** 2).sum() / n

    cands = [r_and_t(gauss_newton(approx([0, 1, 3, 6], 1))), r_and_t(gauss_newton(approx([0, 1, 2], 2))),
             r_and_t(gauss_newton(approx([0, 1, 2, 3, 4], 3)))]
    N = 0
    if cands[1][2] < cands[0][2]:
        N = 1
    if cands[2][2] < cands[N][2]:
        N = 2
    return cands[N][0], cands[N][1], [c[2] for c in cands], N


def epnp_pixels(P32, uv32, K):
    """The image points as solvePnP(EPNP) hands them to epnp: undistortPoints (no distortion: (u - cx) * (1/fx) in double,
    stored as float) times f plus c in double."""
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    xn = ((uv32[:, 0].astype(np.float64) - cx) * (1.0 / fx)).astype(np.float32).astype(np.float64)
    yn = ((uv32[:, 1].astype(np.float64) - cy) * (1.0 / fy)).astype(np.float32).astype(np.float64)
    return np.stack([xn * fx + cx, yn * fy + cy], 1)


def float_errors(P32, uv32, K, R, t):
    """PnPRansacCallback::computeError: projectPoints in double, narrowed to float, squared distance in float."""
    pc = P32.astype(np.float64) @ R.T + t
    with np.errstate(divide="ignore", invalid="ignore"):
        z = 1.0 / pc[:, 2]
        u = (pc[:, 0] * z * K[0, 0] + K[0, 2]).astype(np.float32)
        v = (pc[:, 1] * z * K[1, 1] + K[1, 2]).astype(np.float32)
    du, dv = uv32[:, 0] - u, uv32[:, 1] - v
    return (du * du + dv * dv).astype(np.float32)


def solve_pnp_ransac_cv(P, uv, K, iterations=100, reproj_error=2.0, confidence=0.999, svd=None, refine=True):
    """cv::solvePnPRansac(P, uv, K, noArray(), rvec, tvec, false, iterations, reproj_error, confidence, inliers) restated.
    Returns (ok, rvec, tvec, inlier indices, trace) — trace = per-iteration (subset, inlier count, niters after it)."""
    import cv2
    P32 = np.ascontiguousarray(P, np.float32)
    uv32 = np.ascontiguousarray(uv, np.float32)
    count = len(P32)
    rng = CvRng()
    niters = max(iterations, 1)
    thr = np.float32(reproj_error * reproj_error)
    best_mask, best_model, max_good = None, None, 0
    trace = []
    it = 0
    while it < niters:
        sub = get_subset(rng, count, 5)
        us = epnp_pixels(P32[sub], uv32[sub], K)
        R, t, _, _ = epnp(P32[sub].astype(np.float64), us, K[0, 0], K[1, 1], K[0, 2], K[1, 2], svd=svd)
        if np.all(np.isfinite(R)) and np.all(np.isfinite(t)):
            rvec, _ = cv2.Rodrigues(R)                    # the model travels as (rvec, tvec): one Rodrigues round trip
            R2, _ = cv2.Rodrigues(rvec)
            err = float_errors(P32, uv32, K, R2, t)
            mask = err <= thr
            good = int(mask.sum())
            if good > max(max_good, 4):
                best_mask, best_model, max_good = mask, (rvec.ravel().copy(), t.copy()), good
                niters = ransac_update_num_iters(confidence, (count - good) / count, 5, niters)
        trace.append((sub, max_good, niters))
        it += 1
    if best_model is None:
        return False, None, None, np.zeros(0, np.int32), trace
    inl = np.nonzero(best_mask)[0].astype(np.int32)
    rvec, tvec = best_model
    if refine:
        ok, rvec, tvec = cv2.solvePnP(P32[inl].astype(np.float64), uv32[inl].astype(np.float64), K, None, rvec.reshape(3, 1).copy(),
                                      tvec.reshape(3, 1).copy(), True, cv2.SOLVEPNP_ITERATIVE)
        rvec, tvec = rvec.ravel(), tvec.ravel()
    return True, rvec, tvec, inl, trace

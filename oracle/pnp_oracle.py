"""TEST ORACLE — numpy restatement of the parts of cv::solvePnPRansac that are deterministic
functions of their input (reference call site: src/vo/vo.cpp:318-320; algorithm: OpenCV
calib3d solvepnp.cpp / ptsetreg.cpp).  Not product code.

  count_inliers   reprojection test ||proj - obs||^2 <= thr^2 for a list of hypothesis poses
  refine          the final pose: least-squares minimum of the reprojection error over the
                  inlier set (what solvePnP(SOLVEPNP_ITERATIVE) converges to)
  rodrigues / rvec_from_R

OpenCV is an absent third-party dependency; tests pin these against cv2.solvePnP /
cv2.projectPoints live and against tests/golden/pnp_config3.npz."""
from __future__ import annotations

import numpy as np


def rodrigues(rvec):
    rvec = np.asarray(rvec, np.float64).reshape(3)
    th = np.linalg.norm(rvec)
    if th < 1e-12:
        return np.eye(3)
    k = rvec / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def rvec_from_R(R):
    R = np.asarray(R, np.float64).reshape(3, 3)
    c = np.clip((np.trace(R) - 1) / 2, -1, 1)
    th = np.arccos(c)
    if th < 1e-12:
        return np.zeros(3)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / (2 * np.sin(th))
    return w * th


def project(P, R, t, K):
    pc = np.asarray(P, np.float64) @ np.asarray(R).T + np.asarray(t).reshape(3)
    z = pc[:, 2]
    uv = np.stack([K[0, 0] * pc[:, 0] / z + K[0, 2], K[1, 1] * pc[:, 1] / z + K[1, 2]], 1)
    return uv, z


def reproj_err2(P, uv, K, R, t):
    proj, z = project(P, R, t, K)
    e = ((proj - np.asarray(uv, np.float64)) ** 2).sum(1)
    e[~(z > 1e-9)] = np.inf
    return e


def count_inliers(P, uv, K, poses, thr):
    """poses: H x 12 (R row-major, t).  Returns (counts[H], margin[H]) where margin is the smallest
    |err^2 - thr^2| over the points (so tests can exclude numerically borderline hypotheses)."""
    counts = np.zeros(len(poses), np.int64)
    margin = np.zeros(len(poses))
    for h, p in enumerate(poses):
        e = reproj_err2(P, uv, K, p[:9].reshape(3, 3), p[9:])
        counts[h] = int((e <= thr * thr).sum())
        margin[h] = np.min(np.abs(e[np.isfinite(e)] - thr * thr)) if np.isfinite(e).any() else np.inf
    return counts, margin


def refine(P, uv, K, rvec, tvec):
    """Least-squares minimum of the reprojection error from (rvec, tvec) (scipy LM on 6 dof)."""
    from scipy.optimize import least_squares
    P = np.asarray(P, np.float64)
    uv = np.asarray(uv, np.float64)

    def res(x):
        proj, _ = project(P, rodrigues(x[:3]), x[3:], K)
        return (proj - uv).ravel()

    x0 = np.concatenate([np.asarray(rvec, np.float64).reshape(3), np.asarray(tvec, np.float64).reshape(3)])
    r = least_squares(res, x0, method="lm", xtol=1e-15, ftol=1e-15, gtol=1e-15)
    return r.x[:3], r.x[3:]

"""TEST INFRASTRUCTURE (never imported by the product): the reference's two-view functions restated over the
OpenCV routines they call — cv2 4.13 is the same calib3d code, so these ARE the reference's arithmetic.

  esti_motion_by_essential   reference src/geometry/epipolar_geometry.cpp:17-57
  do_triangulation           reference src/geometry/epipolar_geometry.cpp:130-175
"""
import numpy as np


def esti_motion_by_essential(pts1, pts2, K, prob=0.999, threshold=1.0):
    """Returns (E scaled so that E[2,2] = 1, R, t unit, inlier indices from findEssentialMat's mask)."""
    import cv2
    p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 1, 2)
    p2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 1, 2)
    pp = (float(K[0, 2]), float(K[1, 2]))                       # :26
    focal = float(K[0, 0] + K[1, 1]) / 2                        # :27
    E, mask = cv2.findEssentialMat(p1, p2, focal, pp, cv2.RANSAC, prob, threshold)          # :35-38
    E = E / E[2, 2]                                             # :39
    inliers = np.nonzero(mask.ravel() == 1)[0].astype(np.int32)  # :43-49 (taken BEFORE recoverPose edits the mask)
    _, R, t, _ = cv2.recoverPose(E, p1, p2, focal=focal, pp=pp, mask=mask.copy())           # :52
    t = t.ravel() / np.linalg.norm(t)                           # :54-55
    return E, R, t, inliers


def do_triangulation(pts_np1, pts_np2, R, t, inliers):
    """Points on the normalised plane of both cameras -> 3-D points in camera 1 (float32, n_inliers x 3)."""
    import cv2
    a = np.ascontiguousarray(pts_np1, np.float32)[inliers]       # :139-144
    b = np.ascontiguousarray(pts_np2, np.float32)[inliers]
    T1 = np.hstack([np.eye(3), np.zeros((3, 1))]).astype(np.float32)                         # :146-149
    T2 = np.hstack([np.asarray(R, np.float64), np.asarray(t, np.float64).reshape(3, 1)])     # convertRt2T_3x4 (:150)
    X = cv2.triangulatePoints(T1.astype(np.float64), T2, a.T.copy(), b.T.copy())             # :153-156
    X = X.astype(np.float32)
    return (X[:3] / X[3]).T.copy()                              # :160-168


def esti_motion_by_homography(pts1, pts2, K, threshold=3.0):
    """reference src/geometry/epipolar_geometry.cpp:90-128.  Returns (H with H[2,2] = 1, Rs, ts unit, normals, inliers)."""
    import cv2
    p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 1, 2)
    p2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 1, 2)
    H, mask = cv2.findHomography(p1, p2, cv2.RANSAC, threshold)                 # :106
    H = H / H[2, 2]                                                              # :107
    inliers = np.nonzero(mask.ravel() == 1)[0].astype(np.int32)                  # :110-117
    _, Rs, ts, ns = cv2.decomposeHomographyMat(H, np.asarray(K, np.float64))     # :120-121
    ts = [t.ravel() / np.linalg.norm(t) for t in ts]                             # :123-127
    return H, [np.asarray(R) for R in Rs], ts, [n.ravel() for n in ns], inliers


def remove_wrong_rt_of_homography(pts_np1, pts_np2, inliers, Rs, ts, normals):
    """reference src/geometry/epipolar_geometry.cpp:59-88: indices of the solutions that survive
    cv::filterHomographyDecompByVisibleRefpoints on the inlier points."""
    import cv2
    a = np.ascontiguousarray(pts_np1, np.float32)[inliers].reshape(-1, 1, 2)
    b = np.ascontiguousarray(pts_np2, np.float32)[inliers].reshape(-1, 1, 2)
    sol = cv2.filterHomographyDecompByVisibleRefpoints([np.asarray(R, np.float64) for R in Rs],
                                                       [np.asarray(n, np.float64).reshape(3, 1) for n in normals], a, b)
    return [] if sol is None else sol.ravel().tolist()

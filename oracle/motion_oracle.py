"""TEST INFRASTRUCTURE: numpy restatement of the E / H scoring and choice of the reference's two-view initialisation.

  check_essential_score   reference src/geometry/motion_estimation.cpp:501-581
  check_homography_score  reference src/geometry/motion_estimation.cpp:583-664
  choose_e_or_h           reference src/geometry/motion_estimation.cpp:134-154
"""
import numpy as np


def check_essential_score(E21, K, pts1, pts2, inliers, sigma=1.0):
    Kinv = np.linalg.inv(K)
    F = Kinv.T @ E21 @ Kinv                                         # :507-509
    th, th_score, inv_s2 = 3.841, 5.991, 1.0 / (sigma * sigma)
    score, keep = 0.0, []
    for i in inliers:
        u1, v1 = float(pts1[i][0]), float(pts1[i][1])
        u2, v2 = float(pts2[i][0]), float(pts2[i][1])
        good = True
        a2, b2, c2 = F[0] @ [u1, v1, 1], F[1] @ [u1, v1, 1], F[2] @ [u1, v1, 1]
        chi1 = (a2 * u2 + b2 * v2 + c2) ** 2 / (a2 * a2 + b2 * b2) * inv_s2
        if chi1 > th:
            good = False
        else:
            score += th_score - chi1
        a1, b1, c1 = F[:, 0] @ [u2, v2, 1], F[:, 1] @ [u2, v2, 1], F[:, 2] @ [u2, v2, 1]
        chi2 = (a1 * u1 + b1 * v1 + c1) ** 2 / (a1 * a1 + b1 * b1) * inv_s2
        if chi2 > th:
            good = False
        else:
            score += th_score - chi2
        if good:
            keep.append(int(i))
    return score, np.array(keep, np.int32)


def check_homography_score(H21, pts1, pts2, inliers, sigma=1.0):
    H12 = np.linalg.inv(H21)
    th, inv_s2 = 5.991, 1.0 / (sigma * sigma)
    score, keep = 0.0, []                                           # the reference's `score` starts uninitialised (:586)
    for i in inliers:
        u1, v1 = float(pts1[i][0]), float(pts1[i][1])
        u2, v2 = float(pts2[i][0]), float(pts2[i][1])
        good = True
        q = H12 @ [u2, v2, 1]
        chi1 = ((u1 - q[0] / q[2]) ** 2 + (v1 - q[1] / q[2]) ** 2) * inv_s2
        if chi1 > th:
            good = False
        else:
            score += th - chi1
        q = H21 @ [u1, v1, 1]
        chi2 = ((u2 - q[0] / q[2]) ** 2 + (v2 - q[1] / q[2]) ** 2) * inv_s2
        if chi2 > th:
            good = False
        else:
            score += th - chi2
        if good:
            keep.append(int(i))
    return score, np.array(keep, np.int32)


def choose_e_or_h(score_e, score_h, h_normals):
    ratio = score_h / (score_e + score_h)                           # :137
    best = 0
    if ratio > 0.5 and len(h_normals):
        best, largest = 1, abs(h_normals[0][2])
        for i in range(2, len(h_normals) + 1):
            if abs(h_normals[i - 1][2]) > largest:
                largest, best = abs(h_normals[i - 1][2]), i
    return best, ratio

"""TEST INFRASTRUCTURE: numpy restatement of the host-side initialisation / keyframe decisions of the reference VO.

  retain_good_triangulation  reference src/vo/vo.cpp:181-244 (+ basics::calcAngleBetweenTwoVectors, opencv_funcs.cpp:176-190)
  normalize_init_depth       reference src/vo/vo.cpp:96-110  (+ basics::calcMeanDepth / scalePointPos)
  is_vo_good_to_init         reference src/vo/vo.cpp:113-172 (+ computeMeanDistBetweenKeypoints, feature_match.cpp:262-279)
  check_large_move           reference src/vo/vo.cpp:247-265

Only tests/ import this file.
"""
import numpy as np

REF_PI = 3.1415926  # the constant the reference divides by (vo.cpp:213)


def retain_good_triangulation(pts3d_in_curr, T_w_c_curr, T_w_c_ref, min_triang_angle, max_ratio):
    pts = np.asarray(pts3d_in_curr, np.float32).reshape(-1, 3)
    if len(pts) == 0:
        return np.zeros(0, np.int32), np.zeros(0)
    Tc, Tr = np.asarray(T_w_c_curr, float).reshape(4, 4), np.asarray(T_w_c_ref, float).reshape(4, 4)
    pw = (Tc[:3, :3] @ pts.astype(float).T + Tc[:3, 3:4]).T.astype(np.float32).astype(float)   # Point3f in world (:208)
    v1, v2 = Tc[:3, 3] - pw, Tr[:3, 3] - pw
    cosv = (v1 * v2).sum(1) / (np.linalg.norm(v1, axis=1) * np.linalg.norm(v2, axis=1))
    ang = np.arccos(cosv) / REF_PI * 180.0
    med = np.sort(ang)[len(ang) // 2]                                                         # :220
    keep = ~((ang < min_triang_angle) | (ang / med > max_ratio))                              # :239-240
    return np.nonzero(keep)[0].astype(np.int32), ang[keep]


def normalize_init_depth(pts3d, t, assumed_mean_depth):
    pts = np.asarray(pts3d, np.float32).reshape(-1, 3)
    scale = assumed_mean_depth / (pts[:, 2].astype(float).sum() / len(pts))
    return (pts.astype(float) * scale).astype(np.float32), np.asarray(t, float) * scale, scale


def is_vo_good_to_init(k_ref, k_cur, angles, min_inlier_matches, min_pixel_dist, min_median_angle):
    k_ref, k_cur = np.asarray(k_ref, np.float32).reshape(-1, 2), np.asarray(k_cur, np.float32).reshape(-1, 2)
    c0 = len(k_ref) >= min_inlier_matches
    d = (k_ref - k_cur).astype(float)                      # Point2f difference is a float, then the norm in double
    mean_dist = np.sqrt((d * d).sum(1)).sum() / len(k_ref) if len(k_ref) else float("nan")
    c1 = mean_dist > min_pixel_dist
    if len(angles):
        med = np.sort(np.asarray(angles, float))[len(angles) // 2]
        c2 = med > min_median_angle
    else:
        med, c2 = 0.0, False
    return bool(c0 and c1 and c2), mean_dist, med


def check_large_move(T_w_c_curr, T_w_c_ref, min_dist):
    T = np.linalg.inv(np.asarray(T_w_c_ref, float).reshape(4, 4)) @ np.asarray(T_w_c_curr, float).reshape(4, 4)
    dist = np.linalg.norm(T[:3, 3])
    ang = np.arccos(np.clip((np.trace(T[:3, :3]) - 1) / 2, -1, 1))
    return bool(dist > min_dist), dist, ang

"""TEST ORACLE loader — ctypes view of oracle/build/liboracle.so (CPU restatements).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product (libmvo.so) never does.
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB = HERE / "build" / "liboracle.so"

KEYPOINT_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                           ("response", "<f4"), ("octave", "<i4"), ("class_id", "<i4")])
DMATCH_DTYPE = np.dtype([("query_idx", "<i4"), ("train_idx", "<i4"), ("img_idx", "<i4"), ("distance", "<f4")])

_lib = None


def build():
    subprocess.run(["make", "-C", str(HERE)], check=True, capture_output=True)


def lib():
    global _lib
    if _lib is None:
        if not LIB.exists():
            build()
        _lib = C.CDLL(str(LIB))
        _lib.orc_match_radius_bf.argtypes = [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_float, C.c_void_p]
        _lib.orc_match_features.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                            C.c_void_p, C.c_float, C.c_double, C.c_double, C.c_void_p]
        _lib.orc_retain_best.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    return _lib


def _p(a):
    return None if a is None else C.c_void_p(a.ctypes.data)


def _u8(a):
    return np.ascontiguousarray(a, np.uint8)


def hamming_nn(d1, d2):
    d1, d2 = _u8(d1), _u8(d2)
    out = np.zeros(len(d1), DMATCH_DTYPE)
    lib().orc_hamming_nn(_p(d1), len(d1), _p(d2), len(d2), _p(out))
    return out


def hamming_knn2(d1, d2):
    d1, d2 = _u8(d1), _u8(d2)
    out = np.zeros((len(d1), 2), DMATCH_DTYPE)
    lib().orc_hamming_knn2(_p(d1), len(d1), _p(d2), len(d2), _p(out))
    return out


def match_radius_bf(xy1, xy2, d1, d2, radius):
    d1, d2 = _u8(d1), _u8(d2)
    xy1, xy2 = np.ascontiguousarray(xy1, np.float32), np.ascontiguousarray(xy2, np.float32)
    out = np.zeros(max(len(d1), 1), DMATCH_DTYPE)
    n = lib().orc_match_radius_bf(_p(xy1), _p(xy2), _p(d1), _p(d2), len(d1), len(d2), radius, _p(out))
    return out[:n].copy()


def match_features(d1, d2, method_index, xy1=None, xy2=None, radius=0.0, xiang_gao_ratio=2.0, lowe_ratio=1.0):
    d1, d2 = _u8(d1), _u8(d2)
    if xy1 is not None:
        xy1, xy2 = np.ascontiguousarray(xy1, np.float32), np.ascontiguousarray(xy2, np.float32)
    out = np.zeros(max(len(d1), 1), DMATCH_DTYPE)
    n = lib().orc_match_features(_p(d1), len(d1), _p(d2), len(d2), method_index, _p(xy1), _p(xy2), radius,
                                 xiang_gao_ratio, lowe_ratio, _p(out))
    if n < 0:
        raise RuntimeError("feature_match.cpp::matchFeatures: wrong method index.")
    return out[:n].copy()


def threshold_and_dedup(all_matches, xiang_gao_ratio=2.0):
    m = np.ascontiguousarray(all_matches, DMATCH_DTYPE).copy()
    lib().orc_threshold_and_dedup.argtypes = [C.c_void_p, C.c_int, C.c_double]
    n = lib().orc_threshold_and_dedup(_p(m), len(m), xiang_gao_ratio)
    return m[:n].copy()


def remove_duplicated_matches(m):
    m = np.ascontiguousarray(m, DMATCH_DTYPE).copy()
    n = lib().orc_remove_duplicated_matches(_p(m), len(m))
    return m[:n].copy()


def select_uniform_kpts_by_grid(kpts, rows, cols, max_num_keypoints=1500, grid_size=16, max_pts_per_grid=8):
    kp = np.ascontiguousarray(kpts, KEYPOINT_DTYPE).copy()
    n = lib().orc_select_uniform_kpts_by_grid(_p(kp), len(kp), rows, cols, max_num_keypoints, grid_size, max_pts_per_grid)
    return kp[:n].copy()


def retain_best(response, n_points):
    """cv::KeyPointsFilter::retainBest; returns the surviving original indices in output order."""
    r = np.ascontiguousarray(response, np.float32).copy()
    idx = np.arange(len(r), dtype=np.int32)
    n = lib().orc_retain_best(_p(r), _p(idx), len(r), int(n_points))
    return idx[:n].copy()


def bundle_adjustment(poses_T_w_c, points, edge_frame, edge_point, obs, K, information=None, fix_points=False,
                      update_points=True, iterations=50, huber_delta=1.0, fix_first_pose=False):
    """oracle/ba_oracle.c — restated g2o LM (parity vs real g2o unpinned).  Returns (poses, points, stats)."""
    L = lib()
    L.orc_bundle_adjustment.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double,
                                        C.c_int, C.c_void_p]
    poses = np.ascontiguousarray(poses_T_w_c, np.float64).copy().reshape(-1, 16)
    pts = np.ascontiguousarray(points, np.float32).copy().reshape(-1, 3)
    ef = np.ascontiguousarray(edge_frame, np.int32)
    ep = np.ascontiguousarray(edge_point, np.int32)
    ob = np.ascontiguousarray(obs, np.float32).reshape(-1, 2)
    Kc = np.ascontiguousarray(K, np.float64)
    info = np.ascontiguousarray(np.eye(2) if information is None else information, np.float64)
    stats = np.zeros(4)
    rc = L.orc_bundle_adjustment(_p(poses), len(poses), _p(pts), len(pts), _p(ef), _p(ep), _p(ob), len(ef), _p(Kc),
                                 _p(info), int(fix_points), int(update_points), int(iterations), float(huber_delta),
                                 int(fix_first_pose), _p(stats))
    if rc != 0:
        raise RuntimeError(f"orc_bundle_adjustment failed ({rc})")
    return poses.reshape(-1, 4, 4), pts, stats


def bundle_adjustment_with_trace(*args, max_records=4096, **kw):
    """bundle_adjustment + the per-trial record of its Levenberg-Marquardt loop: rows (iteration, lambda, chi2 of the trial,
    gain ratio, accepted)."""
    L = lib()
    buf, cnt = np.zeros((max_records, 5)), C.c_int(0)
    L.orc_ba_set_trace.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int)]
    L.orc_ba_set_trace(_p(buf), max_records, C.byref(cnt))
    try:
        out = bundle_adjustment(*args, **kw)
    finally:
        L.orc_ba_set_trace(None, 0, None)
    return out + (buf[: cnt.value].copy(),)

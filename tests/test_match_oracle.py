"""Pins the matching oracle (oracle/match_oracle.cpp) against OpenCV's BFMatcher — the
third-party code the reference calls at src/geometry/feature_match.cpp:141,162,208 — live via
cv2 when importable, and against the committed golden vectors."""
import numpy as np
import pytest
from conftest import GOLDEN, have_cv2

import mvo_synth
from oracle import oracle_lib


def _cv_match(d1, d2):
    import cv2
    ms = cv2.BFMatcher(cv2.NORM_HAMMING).match(d1, d2)
    return np.array([(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in ms], oracle_lib.DMATCH_DTYPE)


def _cv_knn(d1, d2):
    import cv2
    ms = cv2.BFMatcher(cv2.NORM_HAMMING).knnMatch(d1, d2, k=2)
    return np.array([[(m.queryIdx, m.trainIdx, m.imgIdx, m.distance) for m in mm] for mm in ms], oracle_lib.DMATCH_DTYPE)


@pytest.mark.skipif(not have_cv2(), reason="cv2 not importable")
@pytest.mark.parametrize("n1,n2,dup", [(1, 2, 0), (37, 53, 0), (300, 257, 5), (2001, 2001, 11)])
def test_oracle_hamming_vs_cv2(n1, n2, dup):
    d1 = mvo_synth.random_descriptors(10 + n1, n1, dup_every=dup)
    d2 = mvo_synth.random_descriptors(20 + n2, n2, dup_every=dup)
    if dup:
        d1[::3] = d2[(np.arange(0, n1, 3) * 7) % n2]      # exact matches + ties through duplicates
    assert oracle_lib.hamming_nn(d1, d2).tobytes() == _cv_match(d1, d2).tobytes()
    assert oracle_lib.hamming_knn2(d1, d2).tobytes() == _cv_knn(d1, d2).tobytes()


def test_oracle_vs_golden():
    g = np.load(GOLDEN / "match_golden.npz")
    d1, d2 = g["d1"], g["d2"]
    nn = oracle_lib.hamming_nn(d1, d2)
    assert np.array_equal(nn["train_idx"], g["nn_train"]) and np.array_equal(nn["distance"], g["nn_dist"])
    kn = oracle_lib.hamming_knn2(d1, d2)
    assert np.array_equal(kn["train_idx"], g["knn_train"]) and np.array_equal(kn["distance"], g["knn_dist"])


def test_oracle_radius_and_features_properties():
    rng = np.random.default_rng(5)
    n1, n2 = 400, 380
    d1, d2 = mvo_synth.random_descriptors(1, n1), mvo_synth.random_descriptors(2, n2)
    xy1 = rng.uniform(0, 640, (n1, 2)).astype(np.float32)
    xy2 = rng.uniform(0, 480, (n2, 2)).astype(np.float32)
    m = oracle_lib.match_radius_bf(xy1, xy2, d1, d2, 50.0)
    # numpy cross-check of feature_match.cpp:86-124
    for mm in m[:50]:
        i = mm["query_idx"]
        dd = ((xy1[i, 0] - xy2[:, 0]) ** 2 + (xy1[i, 1] - xy2[:, 1]) ** 2) <= np.float32(50.0) ** 2
        sad = np.abs(d1[i].astype(int) - d2.astype(int)).sum(1) / 32.0
        sad[~dd] = np.inf
        assert mm["train_idx"] == int(np.argmin(sad)) and mm["distance"] == np.float32(sad.min())
        assert mm["img_idx"] == -1
    for method in (1, 2, 3):
        r = oracle_lib.match_features(d1, d2, method, xy1, xy2, 50.0)
        assert np.all(np.diff(r["train_idx"]) > 0)
    with pytest.raises(RuntimeError):
        oracle_lib.match_features(d1, d2, 4)


def test_select_uniform_grid_matches_python_restatement():
    rng = np.random.default_rng(0)
    n = 5000
    kp = np.zeros(n, oracle_lib.KEYPOINT_DTYPE)
    kp["x"] = rng.uniform(31, 608, n)
    kp["y"] = rng.uniform(31, 448, n)
    out = oracle_lib.select_uniform_kpts_by_grid(kp, 480, 640, 1500, 16, 8)
    # reference feature_match.cpp:68-81 restated in Python
    grid = np.zeros((30, 40), int)
    keep = []
    for i in range(n):
        r, c = int(kp["y"][i]) // 16, int(kp["x"][i]) // 16
        if grid[r, c] < 8:
            keep.append(i)
            grid[r, c] += 1
            if len(keep) > 1500:
                break
    assert len(out) == 1501 and out.tobytes() == kp[keep].tobytes()

"""Pins the ORB oracle (oracle/orb_oracle.py, numpy) against OpenCV's cv::ORB — the third-party
code the reference calls at src/geometry/feature_match.cpp:22-23,34,45,48 — live via cv2 when it
is importable, and against the committed golden vectors (generated from cv2 4.13)."""
import numpy as np
import pytest
from conftest import GOLDEN, have_cv2

import mvo_synth
from oracle import oracle_lib, orb_oracle as oo


def _cvkp(kps):
    return np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id) for k in kps], oo.KEYPOINT_DTYPE)


SCENES = {
    "rect0": lambda: mvo_synth.gray_to_bgr(mvo_synth.rect_scene(0)),
    "color1": lambda: mvo_synth.color_scene(1),              # above the per-level caps on levels 0-1
    "odd": lambda: mvo_synth.gray_to_bgr(mvo_synth.rect_scene(5, 517, 389)),
}


@pytest.mark.skipif(not have_cv2(), reason="cv2 not importable")
@pytest.mark.parametrize("name", list(SCENES))
def test_oracle_detect_compute_vs_cv2(name):
    import cv2
    img = SCENES[name]()
    ref = _cvkp(cv2.ORB_create(8000, 1.2, 4, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 20).detect(img, None))
    got = oo.detect(img)
    assert got.tobytes() == ref.tobytes()
    sel = oracle_lib.select_uniform_kpts_by_grid(ref, img.shape[0], img.shape[1], 2000, 16, 8)
    kps = [cv2.KeyPoint(float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]),
                        int(k["octave"]), int(k["class_id"])) for k in sel]
    kps2, dref = cv2.ORB_create(8000, 1.2, 4).compute(img, kps)
    assert len(kps2) == len(sel)
    assert np.array_equal(oo.compute(img, sel), dref)


@pytest.mark.skipif(not have_cv2(), reason="cv2 not importable")
def test_oracle_stages_vs_cv2():
    import cv2
    g = mvo_synth.rect_scene(3)
    bgr = mvo_synth.color_scene(2)
    assert np.array_equal(oo.bgr_to_gray(bgr), cv2.cvtColor(bgr, cv2.COLOR_BGR2GRAY))
    lv = oo.build_pyramid(g)
    prev = g
    for l in range(1, 4):
        prev = cv2.resize(prev, (lv[l].shape[1], lv[l].shape[0]), interpolation=cv2.INTER_LINEAR_EXACT)
        assert np.array_equal(prev, lv[l])
    fast = cv2.FastFeatureDetector_create(20, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(g, None)
    xs, ys, sc = oo.fast_detect(g, 20)
    assert [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in fast] == list(zip(xs.tolist(), ys.tolist(), sc.tolist()))
    for y, x in [(3.0, 4.0), (-1.0, 2.0), (0.0, -5.0), (7.0, 0.0), (-3.0, -3.0), (1e-3, 250.0)]:
        assert oo.fast_atan2(y, x) == np.float32(cv2.fastAtan2(y, x))


def test_oracle_vs_golden():
    for name in ("rect0", "color1"):
        g = np.load(GOLDEN / f"orb_{name}.npz")
        img = SCENES[name]()
        assert np.array_equal(img, g["image"])
        kp = oo.detect(img)
        assert kp.tobytes() == g["detect"].tobytes()
        sel = oracle_lib.select_uniform_kpts_by_grid(kp, img.shape[0], img.shape[1], 2000, 16, 8)
        assert sel.tobytes() == g["selected"].tobytes()
        assert np.array_equal(oo.compute(img, sel), g["descriptors"])


def test_level_geometry():
    assert oo.level_sizes(640, 480, 4, 1.2) == [(640, 480), (533, 400), (444, 333), (370, 278)]
    assert oo.features_per_level(8000, 4, 1.2) == [2575, 2146, 1788, 1491]

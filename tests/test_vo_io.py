"""On-disk formats either side of the hot path, host code in csrc/vo_io.cpp:
trajectory file (reference src/vo/vo_io.cpp:51-120), image naming (:12-37, run_vo.cpp:90) and the
config.yaml dialect read by my_slam::basics::Config (src/basics/config.cpp:12-47)."""
import ctypes as C
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden"


@pytest.fixture(scope="module")
def lib(built):
    import mvo_b200
    return mvo_b200.load_library()


def _fmt(x):
    """What `std::ofstream << double` prints with default precision (6 significant digits, %g style)."""
    return "%g" % x


def test_pose_file_format_and_roundtrip(lib, tmp_path):
    rng = np.random.default_rng(0)
    T = np.tile(np.eye(4), (5, 1, 1))
    T[:, :3, :] = rng.normal(0, 1, (5, 3, 4))
    T[0, :3, :] = np.arange(1, 13).reshape(3, 4) + 0.123456789        # recognisable entries
    f = tmp_path / "traj.txt"
    assert lib.mvo_write_pose_file(str(f).encode(), T.ctypes.data, len(T)) == 0
    lines = f.read_text().splitlines()
    assert len(lines) == 5
    for k, line in enumerate(lines):
        tok = line.split()
        # tx ty tz, then the rotation column by column: R00 R10 R20 R01 R11 R21 R02 R12 R22 (vo_io.cpp:63-73)
        want = [T[k, 0, 3], T[k, 1, 3], T[k, 2, 3]] + [T[k, j, i] for i in range(3) for j in range(3)]
        assert tok == [_fmt(w) for w in want]
        assert line.endswith(" ")                                     # every number is followed by a blank (:66,72)
    out = np.zeros((8, 16))
    n = C.c_int(0)
    assert lib.mvo_read_pose_file(str(f).encode(), out.ctypes.data, 8, C.byref(n)) == 0 and n.value == 5
    back = out[:5].reshape(5, 4, 4)
    assert np.allclose(back, T, rtol=1e-5, atol=1e-6) and np.array_equal(back[:, 3], np.tile([0, 0, 0, 1.0], (5, 1)))
    # the reader is a flat stream of doubles: line structure does not matter, a trailing partial pose is dropped (:95-112)
    g = tmp_path / "flat.txt"
    vals = np.arange(1, 30, dtype=float)
    g.write_text("\n".join(" ".join(str(v) for v in vals[i:i + 5]) for i in range(0, len(vals), 5)))
    assert lib.mvo_read_pose_file(str(g).encode(), out.ctypes.data, 8, C.byref(n)) == 0 and n.value == 2
    p = vals[:12]
    assert np.array_equal(out[0].reshape(4, 4), np.array([[p[3], p[6], p[9], p[0]], [p[4], p[7], p[10], p[1]],
                                                          [p[5], p[8], p[11], p[2]], [0, 0, 0, 1]]))
    # capacity and missing file
    assert lib.mvo_read_pose_file(str(f).encode(), out.ctypes.data, 2, C.byref(n)) == -4 and n.value == 5
    assert lib.mvo_read_pose_file(str(tmp_path / "nope.txt").encode(), out.ctypes.data, 8, C.byref(n)) == -1
    assert lib.mvo_write_pose_file(str(tmp_path / "no_dir" / "x.txt").encode(), T.ctypes.data, 1) == -1


def test_image_path(lib):
    buf = C.create_string_buffer(256)
    assert lib.mvo_image_path(b"data/dataset_images_matlab", b"/rgb_%05d.png", 7, buf, 256) == 0
    assert buf.value == b"data/dataset_images_matlab/rgb_00007.png"
    assert lib.mvo_image_path(b"d", b"/rgb_%05d.png", 12345, buf, 256) == 0 and buf.value == b"d/rgb_12345.png"
    assert lib.mvo_image_path(b"d", b"/img%d.jpg", 3, buf, 256) == 0 and buf.value == b"d/img3.jpg"
    assert lib.mvo_image_path(b"d", b"/rgb.png", 3, buf, 256) == -1
    assert lib.mvo_image_path(b"d", b"/rgb_%05d.png", 3, buf, 4) == -4


def test_config_yaml(lib):
    import mvo_b200
    h = C.c_void_p()
    assert lib.mvo_config_load(str(GOLDEN / "config_fixture.yaml").encode(), C.byref(h)) == 0
    try:
        d, i = C.c_double(), C.c_int()
        buf = C.create_string_buffer(256)
        assert lib.mvo_config_get_string(h, b"dataset_name", buf, 256) == 0 and buf.value == b"fr1_desk"
        assert lib.mvo_config_get_string(h, b"fr1_desk/dataset_dir", buf, 256) == 0 and buf.value == b"/some/where/with spaces/fr1 desk"
        assert lib.mvo_config_get_string(h, b"save_predicted_traj_to", buf, 256) == 0 and buf.value == b"data/test_data/cam_traj.txt"
        assert lib.mvo_config_get_double(h, b"matlab/camera_info.fx", C.byref(d)) == 0 and d.value == 615.0
        assert lib.mvo_config_get_double(h, b"lowe_method_dist_ratio", C.byref(d)) == 0 and d.value == 0.8
        assert lib.mvo_config_get_int(h, b"lowe_method_dist_ratio", C.byref(i)) == 0 and i.value == 1      # cvRound, like FileNode -> int
        assert lib.mvo_config_get_int(h, b"num_prev_frames_to_opti_by_ba", C.byref(i)) == 0 and i.value == 5
        assert lib.mvo_config_get_bool(h, b"is_enable_ba", C.byref(i)) == 0 and i.value == 1
        assert lib.mvo_config_get_bool(h, b"fr1_desk/is_draw_true_traj", C.byref(i)) == 0 and i.value == 0
        assert lib.mvo_config_get_double(h, b"no_such_key", C.byref(d)) == -1                               # config.cpp:35 throws
        p, tp = mvo_b200.default_params(), mvo_b200.TrackParams()
        lib.mvo_default_track_params(C.byref(tp))
        K = np.zeros(9)
        assert lib.mvo_config_apply(h, C.byref(p), C.byref(tp), K.ctypes.data) == 0
        vp = mvo_b200.VoParams()
        lib.mvo_vo_default_params(C.byref(vp))
        vp.min_inlier_matches, vp.assumed_mean_depth_init, vp.track.ba_window = -1, -1.0, -1
        assert lib.mvo_config_apply_vo(h, C.byref(vp)) == 0                                                 # vo.cpp / vo_addFrame.cpp keys
        assert (vp.match_method_init, vp.max_match_dist_init, vp.max_match_dist_triangulation, vp.min_inlier_matches) == (1, 100.0, 100.0, 15)
        assert (vp.essential_threshold, vp.min_triang_angle, vp.max_ratio_angle_to_median, vp.min_pixel_dist) == (1.0, 1.0, 20.0, 50.0)
        assert (vp.min_median_triangulation_angle, vp.assumed_mean_depth_init, vp.track.ba_window, vp.track.match_radius) == (2.0, 0.8, 5, 50.0)
        assert (p.orb_nfeatures, p.orb_nlevels, p.orb_fast_threshold, p.max_keypoints, p.grid_size, p.max_pts_per_grid) == (8000, 4, 20, 2000, 16, 8)
        assert abs(p.orb_scale_factor - 1.2) < 1e-6 and (p.xiang_gao_ratio, p.lowe_ratio) == (2.0, 1.0)
        assert (tp.match_method, tp.match_radius, tp.ba_enable, tp.ba_window, tp.ba_fix_points) == (1, 50.0, 1, 5, 1)
        assert (tp.max_dist_to_prev, tp.min_dist_keyframe) == (0.3, 0.03) and list(tp.information) == [1.0, 0.0, 0.0, 1.0]
        assert np.array_equal(K.reshape(3, 3), [[517.3, 0, 325.1], [0, 516.5, 249.7], [0, 0, 1]])
        # untouched fields keep their defaults
        assert p.pnp_hypotheses == 4096 and p.ba_iterations == 50 and tp.buffer_size == 20
    finally:
        lib.mvo_config_free(h)
    assert lib.mvo_config_load(b"/nonexistent/config.yaml", C.byref(h)) == -1

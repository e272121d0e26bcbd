"""Host logic of the VO state machine (csrc/vo_pipeline.cpp) on the CPU tier: the file is compiled a second time with
the GPU stages replaced by forwarders (tests/cpp/vo_pipeline_hostcheck.cpp) that this test points at the oracle stages
(cv2 + oracle/), and the result is compared frame by frame with oracle/vo_pipeline_oracle.py — the restatement of
reference src/vo/vo_addFrame.cpp:10-142 and src/vo/vo.cpp — on a synthetic 3-D sequence.  Same stages on both sides, so
every difference is a difference in the state machine: container order, bookkeeping, index plumbing.  The two-view
assembly (csrc/two_view.cpp = helperEstimatePossibleRelativePosesByEpipolarGeometry, reference
src/geometry/motion_estimation.cpp:10-158) is part of that second build too, over forwarders for ITS stages."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import mvo_synth
from conftest import have_cv2

ROOT = Path(__file__).resolve().parent.parent
pytestmark = pytest.mark.skipif(not have_cv2(), reason="cv2 (the reference's third-party code) is needed for the oracle stages")
K = mvo_synth.K_DEFAULT
ROWS, COLS = 480, 640
# identical stages on both sides; what differs is double rounding in the 4x4 algebra (rigid inverse vs numpy's LU inverse),
# which the 10-iteration LM of the BA stage carries along: observed <= 2e-9 over 22 frames
POSE_TOL = 1e-7


def _arr(ptr, shape, dtype):
    n = int(np.prod(shape))
    if n == 0:
        return np.zeros(shape, dtype)
    buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype).reshape(shape)


class Stages:
    """The numeric stages of the oracle pipeline behind the C signatures of the product's entry points."""

    def __init__(self, helper, ba_iterations):
        import mvo_b200
        from oracle import epipolar_oracle, motion_oracle, oracle_lib
        self.h, self.epi, self.mot, self.ol, self.mvo = helper, epipolar_oracle, motion_oracle, oracle_lib, mvo_b200
        self.ba_iterations = ba_iterations
        vp, i, sz, f, d = C.c_void_p, C.c_int, C.c_size_t, C.c_float, C.c_double
        self.protos = [
            (C.CFUNCTYPE(i, vp, i, i, i, sz, vp, vp, vp), self.orb_extract),
            (C.CFUNCTYPE(i, vp, i, vp, i, i, vp, vp, f, vp, vp), self.match_features),
            (C.CFUNCTYPE(i, vp, vp, i, vp, d, vp, vp, vp, vp, vp, vp, vp), self.esti_motion_by_homography),
            (C.CFUNCTYPE(i, vp, vp, i, vp, i, vp, vp, vp, vp), self.remove_wrong_rt_of_homography),
            (C.CFUNCTYPE(i, vp, vp, i, vp, d, vp, vp, vp, vp, vp), self.esti_motion_by_essential),
            (C.CFUNCTYPE(i, vp, vp, i, vp, vp, vp, i, vp), self.do_triangulation),
            (C.CFUNCTYPE(i, vp, vp, i, vp, vp, vp, vp, vp), self.solve_pnp_ransac),
            (C.CFUNCTYPE(i, vp, i, vp, i, vp, vp, vp, i, vp, vp, i, i), self.bundle_adjustment),
        ]
        self.cbs = [proto(fn) for proto, fn in self.protos]          # keep the callbacks alive
        self.table = (C.c_void_p * len(self.cbs))(*[C.cast(cb, C.c_void_p) for cb in self.cbs])

    def orb_extract(self, image, rows, cols, channels, stride, kpts, n_kpts, desc):
        from oracle.vo_pipeline_oracle import Frame
        img = _arr(image, (rows, stride), np.uint8)[:, : cols * channels].reshape(rows, cols, channels)
        fr = Frame(0, np.ascontiguousarray(img if channels == 3 else img[:, :, 0]))
        self.h._extract(fr)
        cap = C.c_int.from_address(n_kpts)
        n = len(fr.kp)
        assert n <= cap.value
        _arr(kpts, (n,), self.mvo.KEYPOINT_DTYPE)[:] = fr.kp
        _arr(desc, (n, 32), np.uint8)[:] = fr.desc
        cap.value = n
        return 0

    def match_features(self, d1, n1, d2, n2, method, xy1, xy2, radius, out, n_out):
        m = self.h._match(_arr(d1, (n1, 32), np.uint8).copy(), _arr(d2, (n2, 32), np.uint8).copy(), method,
                          _arr(xy1, (n1, 2), np.float32).copy(), _arr(xy2, (n2, 2), np.float32).copy(), radius)
        _arr(out, (len(m),), self.mvo.DMATCH_DTYPE)[:] = m
        C.c_int.from_address(n_out).value = len(m)
        return 0

    def esti_motion_by_homography(self, p1, p2, n, Kp, threshold, H, Rs, ts, normals, n_solutions, inliers, n_inliers):
        a, b = _arr(p1, (n, 2), np.float32).copy(), _arr(p2, (n, 2), np.float32).copy()
        Hm, Rh, th, nh, inl = self.epi.esti_motion_by_homography(a, b, _arr(Kp, (3, 3), np.float64).copy(), threshold)
        _arr(H, (3, 3), np.float64)[:] = Hm
        k = len(Rh)
        assert k <= 4
        _arr(Rs, (k, 9), np.float64)[:] = np.array(Rh).reshape(k, 9)
        _arr(ts, (k, 3), np.float64)[:] = np.array(th).reshape(k, 3)
        _arr(normals, (k, 3), np.float64)[:] = np.array(nh).reshape(k, 3)
        C.c_int.from_address(n_solutions).value = k
        _arr(inliers, (len(inl),), np.int32)[:] = inl
        C.c_int.from_address(n_inliers).value = len(inl)
        return 0

    def remove_wrong_rt_of_homography(self, np1, np2, n, inliers, n_inliers, Rs, ts, normals, n_solutions):
        k = C.c_int.from_address(n_solutions)
        R, t, nr = _arr(Rs, (k.value, 9), np.float64), _arr(ts, (k.value, 3), np.float64), _arr(normals, (k.value, 3), np.float64)
        keep = self.epi.remove_wrong_rt_of_homography(_arr(np1, (n, 2), np.float32).copy(), _arr(np2, (n, 2), np.float32).copy(),
                                                      _arr(inliers, (n_inliers,), np.int32).copy(), [r.reshape(3, 3).copy() for r in R],
                                                      [x.copy() for x in t], [x.copy() for x in nr])
        Rk, tk, nk = R[keep].copy(), t[keep].copy(), nr[keep].copy()
        R[: len(keep)], t[: len(keep)], nr[: len(keep)] = Rk, tk, nk
        k.value = len(keep)
        return 0

    def esti_motion_by_essential(self, p1, p2, n, Kp, threshold, E, R, t, inliers, n_inliers):
        a, b = _arr(p1, (n, 2), np.float32).copy(), _arr(p2, (n, 2), np.float32).copy()
        Em, Rm, tm, inl = self.epi.esti_motion_by_essential(a, b, _arr(Kp, (3, 3), np.float64).copy(), 0.999, threshold)
        _arr(E, (3, 3), np.float64)[:] = Em
        _arr(R, (3, 3), np.float64)[:] = Rm
        _arr(t, (3,), np.float64)[:] = tm
        _arr(inliers, (len(inl),), np.int32)[:] = inl
        C.c_int.from_address(n_inliers).value = len(inl)
        return 0

    def do_triangulation(self, np1, np2, n, R, t, inliers, n_inliers, pts3d):
        X = self.epi.do_triangulation(_arr(np1, (n, 2), np.float32).copy(), _arr(np2, (n, 2), np.float32).copy(), _arr(R, (3, 3), np.float64).copy(),
                                      _arr(t, (3,), np.float64).copy(), _arr(inliers, (n_inliers,), np.int32).copy())
        _arr(pts3d, (n_inliers, 3), np.float32)[:] = X
        return 0

    def solve_pnp_ransac(self, pts3d, pts2d, n, Kp, rvec, tvec, inliers, n_inliers):
        import cv2
        ok, rv, tv, inl = cv2.solvePnPRansac(_arr(pts3d, (n, 3), np.float32).copy(), _arr(pts2d, (n, 2), np.float32).copy(),
                                             _arr(Kp, (3, 3), np.float64).copy(), None, None, None, False, 100, 2.0, 0.999)
        if not ok or inl is None:
            return -6                                                 # MVO_ERR_DEGENERATE
        inl = inl.ravel()
        _arr(rvec, (3,), np.float64)[:] = rv.ravel()
        _arr(tvec, (3,), np.float64)[:] = tv.ravel()
        _arr(inliers, (len(inl),), np.int32)[:] = inl
        C.c_int.from_address(n_inliers).value = len(inl)
        return 0

    def bundle_adjustment(self, poses, nf, points, npts, ef, ep, obs, ne, Kp, info, fix_points, update_points):
        P = _arr(poses, (nf, 16), np.float64)
        new_poses, _, _ = self.ol.bundle_adjustment(P.copy(), _arr(points, (npts, 3), np.float32).copy(), _arr(ef, (ne,), np.int32).copy(),
                                                    _arr(ep, (ne,), np.int32).copy(), _arr(obs, (ne, 2), np.float32).copy(),
                                                    _arr(Kp, (3, 3), np.float64).copy(), _arr(info, (2, 2), np.float64).copy(),
                                                    fix_points=bool(fix_points), update_points=bool(update_points), iterations=self.ba_iterations)
        P[:] = new_poses.reshape(nf, 16)
        return 0


@pytest.fixture(scope="module")
def hostcheck(built, tmp_path_factory):
    so = tmp_path_factory.mktemp("vohost") / "libvo_hostcheck.so"
    pkg = ROOT / "monocular-visual-odometry_b200"
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-I", str(ROOT / "include"), "-I", str(pkg / "csrc"),
                    "-I", "/usr/local/cuda/include", str(ROOT / "tests" / "cpp" / "vo_pipeline_hostcheck.cpp"), str(pkg / "csrc" / "vo_pipeline.cpp"),
                    str(pkg / "csrc" / "two_view.cpp"),
                    "-L", str(pkg), "-lmvo", "-Wl,-Bsymbolic", f"-Wl,-rpath,{pkg}", "-o", str(so)], check=True)
    import mvo_b200
    lib = C.CDLL(str(so))
    lib.hostcheck_ctx_new.restype = C.c_void_p
    lib.hostcheck_ctx_new.argtypes = [C.c_int]
    lib.hostcheck_ctx_free.argtypes = [C.c_void_p]
    lib.hostcheck_set_stages.argtypes = [C.c_void_p]
    for name in ("mvo_vo_default_params", "mvo_vo_create", "mvo_vo_destroy", "mvo_vo_add_frame", "mvo_vo_is_initialized", "mvo_vo_map_size",
                 "mvo_vo_num_keyframes", "mvo_vo_get_map", "mvo_vo_frame_pose", "mvo_vo_frame_data", "mvo_vo_has_keyframe"):
        res, args = mvo_b200.SIGNATURES[name]
        getattr(lib, name).restype, getattr(lib, name).argtypes = res, args
    return lib


def test_my_slam_visual_odometry_adapter(hostcheck, tmp_path):
    """my_slam::vo::VisualOdometry / Frame / Map (my_slam_adapter/vo_mvo.h) driven by the run_vo.cpp-style loop of
    tests/cpp/adapter_vo_demo.cpp, linked against the host-check build: the members run_vo.cpp reads and the trajectory file."""
    import mvo_b200
    from oracle import vo_pipeline_oracle as vp
    pkg = ROOT / "monocular-visual-odometry_b200"
    so = tmp_path / "libadapter_vo_demo.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-Dmain=adapter_vo_demo_main", "-I", str(pkg / "my_slam_adapter" / "include"),
                    "-I", str(ROOT / "tests" / "cvshim"), "-I", str(ROOT / "include"), "-I", str(pkg / "my_slam_adapter"),
                    str(ROOT / "tests" / "cpp" / "adapter_vo_demo.cpp"), str(pkg / "my_slam_adapter" / "mvo_context.cpp"), str(pkg / "my_slam_adapter" / "vo_mvo.cpp"),
                    "-L", str(Path(hostcheck._name).parent), "-lvo_hostcheck", "-L", str(pkg), "-lmvo", f"-Wl,-rpath,{Path(hostcheck._name).parent}",
                    f"-Wl,-rpath,{pkg}", "-o", str(so)], check=True)
    n = 16
    frames, _ = mvo_synth.room_sequence(0, n)
    imgs = [mvo_synth.gray_to_bgr(f) for f in frames]
    np.stack(imgs).tofile(tmp_path / "frames.bin")
    oracle = vp.CpuVo(K, ROWS, COLS, max_number_of_keypoints=2000, ba_iterations=10)
    helper = vp.CpuVo(K, ROWS, COLS, max_number_of_keypoints=2000)
    stages = Stages(helper, 10)
    hostcheck.hostcheck_set_stages(C.cast(stages.table, C.c_void_p))
    demo = C.CDLL(str(so))
    argv = (C.c_char_p * 3)(b"adapter_vo_demo", str(tmp_path).encode(), str(n).encode())
    assert getattr(demo, "_Z20adapter_vo_demo_mainiPPc")(3, argv) == 0          # the renamed main has C++ linkage
    rows = [[int(x) for x in ln.split()] for ln in (tmp_path / "summary.txt").read_text().splitlines()]
    assert len(rows) == n
    poses = []
    for i, img in enumerate(imgs):
        T, info = oracle.add_frame(img)
        poses.append(T)
        cur = oracle.curr
        fid, init, is_kf, nk, n_ref, n_map, n_p3, n_pts, prev_ref = rows[i]
        assert (fid, init, is_kf, nk) == (cur.id, int(oracle.state == vp.DOING_TRACKING), int(cur.id in oracle.keyframes), len(cur.kp)), i
        assert (n_ref, n_map, n_p3, n_pts) == (len(cur.matches_with_ref), len(cur.matches_with_map), len(cur.inliers_pts3d), len(oracle.map)), i
        assert prev_ref == (oracle.prev_ref.id if oracle.prev_ref is not None else -1), i
    assert rows[-1][1] == 1 and sum(r[2] for r in rows) >= 3
    # the trajectory file (writePoseToFile format, 6 significant digits) holds the poses returned by addFrame
    lib = mvo_b200.load_library()
    got, cnt = np.zeros((n, 16)), C.c_int(0)
    assert lib.mvo_read_pose_file(str(tmp_path / "traj.txt").encode(), got.ctypes.data, n, C.byref(cnt)) == 0 and cnt.value == n
    assert np.abs(got.reshape(n, 4, 4) - np.stack(poses)).max() < 2e-5


def _run_both(hostcheck, frames, max_kpts, ba_iterations, **vo_cfg):
    import mvo_b200
    from oracle import vo_pipeline_oracle as vp
    oracle = vp.CpuVo(K, ROWS, COLS, max_number_of_keypoints=max_kpts, ba_iterations=ba_iterations, **{k: v for k, v in vo_cfg.items() if not k.startswith("_")})
    helper = vp.CpuVo(K, ROWS, COLS, max_number_of_keypoints=max_kpts)
    stages = Stages(helper, ba_iterations)
    hostcheck.hostcheck_set_stages(C.cast(stages.table, C.c_void_p))
    ctx = C.c_void_p(hostcheck.hostcheck_ctx_new(max_kpts))
    p = mvo_b200.VoParams()
    hostcheck.mvo_vo_default_params(C.byref(p))
    p.track.ba_step_tol = 0.0
    if "min_dist_between_two_keyframes" in vo_cfg:
        p.track.min_dist_keyframe = vo_cfg["min_dist_between_two_keyframes"]
    if "init_calc_homography" in vo_cfg:
        p.init_calc_homography = int(vo_cfg["init_calc_homography"])
    h = C.c_void_p()
    Kc = np.ascontiguousarray(K, np.float64)
    assert hostcheck.mvo_vo_create(ctx, Kc.ctypes.data, ROWS, COLS, C.byref(p), C.byref(h)) == 0
    rows = []
    try:
        for f in frames:
            img = mvo_synth.gray_to_bgr(f)
            T_o, info_o = oracle.add_frame(img)
            T_p, info_p = np.zeros(16), mvo_b200.VoFrameInfo()
            assert hostcheck.mvo_vo_add_frame(h, img.ctypes.data, 3, img.shape[1] * 3, T_p.ctypes.data, C.byref(info_p)) == 0
            rows.append((T_o, info_o, T_p.reshape(4, 4).copy(), info_p))
            _check_frame_members(hostcheck, h, oracle)
        n = C.c_int(hostcheck.mvo_vo_map_size(h))
        ids, pts = np.zeros(max(n.value, 1), np.int32), np.zeros((max(n.value, 1), 3), np.float32)
        assert hostcheck.mvo_vo_get_map(h, ids.ctypes.data, pts.ctypes.data, None, None, len(ids), C.byref(n)) == 0
        final = (ids[: n.value].copy(), pts[: n.value].copy(), hostcheck.mvo_vo_num_keyframes(h),
                 np.stack([_pose(hostcheck, h, k) for k in range(min(len(frames), 20))]))
        if vo_cfg.get("_check_run_sequence"):
            # mvo_vo_run_sequence (run_vo.cpp's main loop as one call) on a fresh instance = the calls above, frame by frame
            h2 = C.c_void_p()
            assert hostcheck.mvo_vo_create(ctx, Kc.ctypes.data, ROWS, COLS, C.byref(p), C.byref(h2)) == 0
            try:
                imgs = [np.ascontiguousarray(mvo_synth.gray_to_bgr(f)) for f in frames]
                ptrs = (C.c_void_p * len(imgs))(*[im.ctypes.data for im in imgs])
                Ts, infos, done = np.zeros((len(imgs), 16)), (mvo_b200.VoFrameInfo * len(imgs))(), C.c_int(0)
                hostcheck.mvo_vo_run_sequence.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_size_t, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
                assert hostcheck.mvo_vo_run_sequence(h2, ptrs, len(imgs), 3, imgs[0].shape[1] * 3, 0, Ts.ctypes.data, infos, C.byref(done)) == 0
                assert done.value == len(imgs)
                for i, (_, _, T_p, info_p) in enumerate(rows):
                    assert np.array_equal(Ts[i].reshape(4, 4), T_p), i
                    assert (infos[i].state_out, infos[i].keyframe, infos[i].map_points, infos[i].n_inliers) == (info_p.state_out, info_p.keyframe, info_p.map_points, info_p.n_inliers), i
            finally:
                hostcheck.mvo_vo_destroy(h2)
    finally:
        hostcheck.mvo_vo_destroy(h)
        hostcheck.hostcheck_ctx_free(ctx)
    return oracle, rows, final


def _frame_data(lib, h, which, what, dtype, tail=()):
    n = C.c_int(0)
    rc = lib.mvo_vo_frame_data(h, which, what, None, 0, C.byref(n))
    assert rc in (0, -4), rc                                          # MVO_ERR_CAPACITY reports the size
    out = np.zeros((max(n.value, 1),) + tail, dtype)
    assert lib.mvo_vo_frame_data(h, which, what, out.ctypes.data, len(out), C.byref(n)) == 0
    return out[: n.value]


def _check_frame_members(lib, h, oracle):
    """What run_vo.cpp's display code reads after addFrame (run_vo.cpp:184-232, 286-300)."""
    import mvo_b200
    cur = oracle.curr
    assert _frame_data(lib, h, 0, 5, np.int32)[0] == cur.id
    assert _frame_data(lib, h, 0, 0, mvo_b200.KEYPOINT_DTYPE).tobytes() == np.ascontiguousarray(cur.kp).tobytes()
    assert np.array_equal(_frame_data(lib, h, 0, 1, np.uint8, (32,)), cur.desc)
    for what, ref in ((2, cur.matches_with_ref), (3, cur.matches_with_map)):
        got = _frame_data(lib, h, 0, what, mvo_b200.DMATCH_DTYPE)
        assert np.array_equal(got["query_idx"], ref["query_idx"]) and np.array_equal(got["train_idx"], ref["train_idx"])
    p3 = _frame_data(lib, h, 0, 4, np.float32, (3,))
    assert p3.shape == cur.inliers_pts3d.shape and (len(p3) == 0 or np.abs(p3 - cur.inliers_pts3d).max() < 1e-5)
    assert bool(lib.mvo_vo_has_keyframe(h, cur.id)) == (cur.id in oracle.keyframes)
    if oracle.prev_ref is None:
        n = C.c_int(0)
        assert lib.mvo_vo_frame_data(h, -1, 5, None, 0, C.byref(n)) == -1          # no previous reference keyframe yet
    else:
        assert _frame_data(lib, h, -1, 5, np.int32)[0] == oracle.prev_ref.id
    n = C.c_int(0)
    assert lib.mvo_vo_frame_data(h, 0, 99, None, 0, C.byref(n)) == -1 and lib.mvo_vo_frame_data(h, 64, 0, None, 0, C.byref(n)) == -1


def _pose(lib, h, k):
    T = np.zeros(16)
    assert lib.mvo_vo_frame_pose(h, k, T.ctypes.data) == 0
    return T.reshape(4, 4)


def test_state_machine_equals_oracle_on_room_sequence(hostcheck):
    from oracle import vo_pipeline_oracle as vp
    frames, truth = mvo_synth.room_sequence(0, 28)
    oracle, rows, (ids, pts, n_kf, buff_poses) = _run_both(hostcheck, frames, 2000, 10)
    keys = [("state_in", "state_in"), ("state_out", "state_out"), ("keyframe", "keyframe"), ("n_keypoints", "n_keypoints"), ("n_matches", "n_matches"),
            ("n_inliers", "n_inliers"), ("pnp_ok", "pnp_ok"), ("ba_frames", "ba_frames"), ("best_sol", "best_sol"), ("map_points", "map_points")]
    for i, (T_o, io, T_p, ip) in enumerate(rows):
        for ko, kp in keys:
            assert io[ko] == getattr(ip, kp), (i, ko, io[ko], getattr(ip, kp))
        for opt, kp in (("n_candidates", "n_candidates"), ("ba_edges", "ba_edges"), ("kf_matches", "kf_matches"), ("kf_new_points_in", "kf_new_points")):
            if opt in io:
                assert io[opt] == getattr(ip, kp), (i, opt, io[opt], getattr(ip, kp))
        assert np.abs(T_o - T_p).max() < POSE_TOL, (i, np.abs(T_o - T_p).max())
        if "T_pnp" in io:
            assert np.abs(io["T_pnp"] - np.array(ip.T_w_c_pnp).reshape(4, 4)).max() < POSE_TOL
        if "init_median_angle" in io:
            assert abs(io["init_median_angle"] - ip.init_median_angle) < 1e-9 and abs(io["init_mean_pixel_dist"] - ip.init_mean_pixel_dist) < 1e-9
            assert abs(io["score_e"] - ip.score_e) < 1e-6 * max(1, io["score_e"]) and abs(io["eh_ratio"] - ip.eh_ratio) < 1e-9
    # the sequence exercises every branch: initialisation after several skipped frames, tracking, >= 2 later keyframes with culling
    states = [r[1]["state_out"] for r in rows]
    assert states[0] == vp.DOING_INITIALIZATION and states[-1] == vp.DOING_TRACKING
    init_at = states.index(vp.DOING_TRACKING)
    assert init_at >= 2 and sum(r[1]["keyframe"] for r in rows[init_at + 1:]) >= 2
    assert n_kf == len(oracle.keyframes)
    # the map outgrew 1000 points, so the later keyframes culled with the raised erase ratio of optimizeMap_ (vo.cpp:519-524)
    assert max(r[1]["map_points"] for r in rows) > 1000 and oracle.map_point_erase_ratio > 0.1
    # the map: same ids in the same container order, same positions
    assert ids.tolist() == oracle.map.keys()
    assert np.abs(pts - np.stack([oracle.map[i].pos for i in ids])).max() < 1e-6          # float32 positions from poses that agree to POSE_TOL
    # buffered poses (later BA updates included)
    ob = list(oracle.buff)
    for k in range(len(buff_poses)):
        assert np.abs(buff_poses[k] - ob[len(ob) - 1 - k].T_w_c).max() < POSE_TOL
    # and the trajectory is a sane one: RMS error after similarity alignment below 2 % of the path length
    est = [r[2] for r in rows[init_at:]]
    err, _ = vp.trajectory_error(est, truth[init_at:])
    path = np.linalg.norm(truth[-1][:3, 3] - truth[init_at][:3, 3])
    assert err < 0.02 * path, (err, path)


def test_state_machine_without_homography_and_dense_keyframes(hostcheck):
    """Essential-only initialisation and a keyframe on every tracked frame."""
    frames, _ = mvo_synth.room_sequence(0, 14)
    oracle, rows, (ids, pts, n_kf, _) = _run_both(hostcheck, frames, 2000, 10, init_calc_homography=False, min_dist_between_two_keyframes=0.005)
    for i, (T_o, io, T_p, ip) in enumerate(rows):
        assert (io["state_out"], io["keyframe"], io["map_points"], io["n_inliers"]) == (ip.state_out, ip.keyframe, ip.map_points, ip.n_inliers), i
        assert np.abs(T_o - T_p).max() < POSE_TOL
    assert sum(r[1]["keyframe"] for r in rows) >= 5 and n_kf == len(oracle.keyframes)
    assert ids.tolist() == oracle.map.keys()


def test_state_machine_through_blank_and_far_away_frames(hostcheck):
    """Frames the reference's happy path never sees: a featureless image while initialising (no keypoints, no matches), one
    while tracking (PnP cannot run: the pose falls back to the previous frame's, vo.cpp:376-379), and a view from far away
    (vo.cpp:360-369 rejects the jump or the PnP fails) — the state machine goes on, and agrees with the oracle throughout."""
    from oracle import vo_pipeline_oracle as vp
    frames, truth = mvo_synth.room_sequence(0, 16)
    planes = mvo_synth._room_planes(0)
    blank = np.full_like(frames[0], 117)
    far = truth[12].copy()
    far[:3, 3] += np.array([1.2, 0.0, 2.6])       # 2.9 m = 0.48 in the units of the normalised map (threshold 0.3)
    seq = frames[:3] + [blank] + frames[3:10] + [blank] + frames[10:12] + [mvo_synth.render_room(far, planes)] + frames[12:]
    oracle, rows, (ids, pts, n_kf, _) = _run_both(hostcheck, seq, 2000, 10, _check_run_sequence=True)
    for i, (T_o, io, T_p, ip) in enumerate(rows):
        assert (io["state_out"], io["keyframe"], io["map_points"], io["n_keypoints"], io["n_matches"], io["n_inliers"], io["pnp_ok"]) == \
               (ip.state_out, ip.keyframe, ip.map_points, ip.n_keypoints, ip.n_matches, ip.n_inliers, ip.pnp_ok), i
        assert np.abs(T_o - T_p).max() < POSE_TOL, i
    info = [r[1] for r in rows]
    assert info[3]["n_keypoints"] == 0 and info[3]["state_out"] == vp.DOING_INITIALIZATION
    assert info[8]["state_out"] == vp.DOING_TRACKING                                # initialised in spite of the blank frame
    assert info[11]["n_keypoints"] == 0 and info[11]["pnp_ok"] == 0 and np.array_equal(rows[11][0], rows[10][0])
    assert info[14]["pnp_ok"] == 0 and np.array_equal(rows[14][0], rows[13][0])      # the far-away view is not accepted
    assert all(i["pnp_ok"] == 1 for i in info[15:])                                 # and tracking resumes
    assert ids.tolist() == oracle.map.keys() and n_kf == len(oracle.keyframes)

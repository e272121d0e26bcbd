"""SLOW (opt-in: MVO_SLOW_TESTS=1, about three minutes): the VO state machine on the CPU tier with the PRODUCT's two-view
numerics — the essential-matrix, homography and triangulation CUDA kernels executed through tests/cpp/cuda_emu.h and the
host-side decomposition / visibility filter of csrc/epipolar_math.cuh — under the real csrc/two_view.cpp and
csrc/vo_pipeline.cpp; extraction, matching, PnP and BA stay on the oracle stages.  Shows what the hardware run of
tests/test_vo_pipeline_gpu.py is expected to show for the initialisation and the keyframe branch: the E / H choice, the
frame at which the map is created, and a trajectory no worse than the oracle pipeline's."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np
import pytest

import mvo_synth
from conftest import have_cv2
from test_vo_pipeline_host import COLS, K, ROWS, Stages, _arr, hostcheck  # noqa: F401  (fixture)

ROOT = Path(__file__).resolve().parent.parent
pytestmark = [pytest.mark.skipif(not have_cv2(), reason="cv2 not importable"),
              pytest.mark.skipif(os.environ.get("MVO_SLOW_TESTS", "0") == "0", reason="slow emulation run: set MVO_SLOW_TESTS=1")]
HYP = 384


class EmuStages(Stages):
    def __init__(self, helper, ba_iterations, emu, epi_host):
        self.emu, self.epi_host = emu, epi_host
        super().__init__(helper, ba_iterations)

    def esti_motion_by_essential(self, p1, p2, n, Kp, threshold, E, R, t, inliers, n_inliers):
        oi = np.zeros(8, np.int32)
        inl = np.zeros(n, np.int32)
        ni = self.emu.emu_essential(p1, p2, n, Kp, threshold, HYP, 12345, E, R, t, inl.ctypes.data, oi.ctypes.data)
        if ni < 8:
            return -6
        _arr(inliers, (ni,), np.int32)[:] = inl[:ni]
        C.c_int.from_address(n_inliers).value = ni
        return 0

    def esti_motion_by_homography(self, p1, p2, n, Kp, threshold, H, Rs, ts, normals, n_solutions, inliers, n_inliers):
        oi = np.zeros(8, np.int32)
        inl = np.zeros(n, np.int32)
        ni = self.emu.emu_homography(p1, p2, n, Kp, threshold, HYP, 12345, H, inl.ctypes.data, oi.ctypes.data)
        if ni < 4:
            return -6
        _arr(inliers, (ni,), np.int32)[:] = inl[:ni]
        C.c_int.from_address(n_inliers).value = ni
        # cv::decomposeHomographyMat on Hn = K^-1 H K, t /= |t|  (the host part of mvo_esti_motion_by_homography, csrc/epipolar.cu)
        Kc, Hm = _arr(Kp, (3, 3), np.float64), _arr(H, (3, 3), np.float64)
        Hn = np.ascontiguousarray(np.linalg.inv(Kc) @ Hm @ Kc)
        k = self.epi_host.epi_decompose_homography(Hn.ctypes.data, Rs, ts, normals)
        tv = _arr(ts, (k, 3), np.float64)
        for s in range(k):
            nt = np.linalg.norm(tv[s])
            if nt > 0:
                tv[s] /= nt
        C.c_int.from_address(n_solutions).value = k
        return 0

    def remove_wrong_rt_of_homography(self, np1, np2, n, inliers, n_inliers, Rs, ts, normals, n_solutions):
        k = C.c_int.from_address(n_solutions)
        keep = np.zeros(4, np.int32)
        self.epi_host.epi_filter_homography(Rs, normals, k.value, np1, np2, inliers, n_inliers, keep.ctypes.data)
        R, t, nr = _arr(Rs, (k.value, 9), np.float64), _arr(ts, (k.value, 3), np.float64), _arr(normals, (k.value, 3), np.float64)
        idx = [s for s in range(k.value) if keep[s]]
        Rk, tk, nk = R[idx].copy(), t[idx].copy(), nr[idx].copy()
        R[: len(idx)], t[: len(idx)], nr[: len(idx)] = Rk, tk, nk
        k.value = len(idx)
        return 0

    def do_triangulation(self, np1, np2, n, R, t, inliers, n_inliers, pts3d):
        return self.emu.emu_triangulate(np1, np2, inliers, n_inliers, R, t, pts3d)


def _build(tmp, src, name, extra=()):
    so = tmp / name
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", str(ROOT / "include"),
                    "-I", str(ROOT / "monocular-visual-odometry_b200" / "csrc"), "-I", str(ROOT / "tests" / "cpp"), "-I", "/usr/local/cuda/include",
                    *extra, str(ROOT / "tests" / "cpp" / src), "-o", str(so)], check=True)
    return C.CDLL(str(so))


def test_state_machine_with_the_products_two_view_kernels(hostcheck, tmp_path):  # noqa: F811
    import mvo_b200
    from oracle import vo_pipeline_oracle as vp
    emu = _build(tmp_path, "two_view_emu.cpp", "libtwo_view_emu.so")
    vpt, i, d, u64 = C.c_void_p, C.c_int, C.c_double, C.c_uint64
    emu.emu_essential.argtypes = [vpt, vpt, i, vpt, d, i, u64, vpt, vpt, vpt, vpt, vpt]
    emu.emu_homography.argtypes = [vpt, vpt, i, vpt, d, i, u64, vpt, vpt, vpt]
    emu.emu_triangulate.argtypes = [vpt, vpt, vpt, i, vpt, vpt, vpt]
    epi_host = _build(tmp_path, "epipolar_math_host.cpp", "libepi_host.so")
    epi_host.epi_decompose_homography.argtypes = [vpt] * 4
    epi_host.epi_filter_homography.argtypes = [vpt, vpt, i, vpt, vpt, vpt, i, vpt]
    n = 20
    frames, truth = mvo_synth.room_sequence(0, n)
    oracle = vp.CpuVo(K, ROWS, COLS, max_number_of_keypoints=2000, ba_iterations=10)
    helper = vp.CpuVo(K, ROWS, COLS, max_number_of_keypoints=2000)
    stages = EmuStages(helper, 10, emu, epi_host)
    hostcheck.hostcheck_set_stages(C.cast(stages.table, C.c_void_p))
    ctx = C.c_void_p(hostcheck.hostcheck_ctx_new(2000))
    p = mvo_b200.VoParams()
    hostcheck.mvo_vo_default_params(C.byref(p))
    h = C.c_void_p()
    Kc = np.ascontiguousarray(K, np.float64)
    assert hostcheck.mvo_vo_create(ctx, Kc.ctypes.data, ROWS, COLS, C.byref(p), C.byref(h)) == 0
    Tp, To, infos = [], [], []
    for f in frames:
        img = mvo_synth.gray_to_bgr(f)
        T, info = np.zeros(16), mvo_b200.VoFrameInfo()
        assert hostcheck.mvo_vo_add_frame(h, img.ctypes.data, 3, img.shape[1] * 3, T.ctypes.data, C.byref(info)) == 0
        Tp.append(T.reshape(4, 4).copy())
        infos.append((info.state_out, info.keyframe, info.best_sol, info.n_inliers, info.map_points, round(info.eh_ratio, 3)))
        To.append(oracle.add_frame(img)[0])
    hostcheck.mvo_vo_destroy(h)
    hostcheck.hostcheck_ctx_free(ctx)
    print(infos)
    states = [s[0] for s in infos]
    so = [l["state_out"] for l in oracle.log]
    g0, c0 = states.index(2), so.index(2)
    assert abs(g0 - c0) <= 2 and infos[g0][2] == 0 and infos[g0][4] >= 100       # essential-matrix solution chosen, a usable map
    assert all(s[5] < 0.5 for s in infos[1:g0 + 1])                               # H / (E + H) stays on the E side for this 3-D scene
    assert sum(s[1] for s in infos[g0 + 1:]) >= 2                                 # keyframes keep coming
    s0 = max(g0, c0)
    ep, _ = vp.trajectory_error(Tp[s0:], truth[s0:])
    eo, _ = vp.trajectory_error(To[s0:], truth[s0:])
    path = float(np.linalg.norm(truth[-1][:3, 3] - truth[s0][:3, 3]))
    print("trajectory RMS error: product two-view kernels %.5f, oracle %.5f, path %.3f" % (ep, eo, path))
    assert ep < 0.02 * path and ep <= 1.25 * eo + 1e-4


class RealEntryPointStages(Stages):
    """Extraction, matching and the two-view stage go to the library's REAL entry points — host code and kernels built for the host
    by tests/emu_build.py — and only PnP and BA (whose LM runs on thread-block clusters, not emulated) stay on the oracle."""

    def __init__(self, helper, ba_iterations, lib, ctx):
        self.lib, self.ctx = lib, ctx
        super().__init__(helper, ba_iterations)

    def orb_extract(self, image, rows, cols, channels, stride, kpts, n_kpts, desc):
        return self.lib.mvo_orb_extract(self.ctx, image, rows, cols, channels, stride, kpts, n_kpts, desc)

    def match_features(self, d1, n1, d2, n2, method, xy1, xy2, radius, out, n_out):
        return self.lib.mvo_match_features(self.ctx, d1, n1, d2, n2, method, xy1, xy2, radius, out, n_out)

    def esti_motion_by_essential(self, p1, p2, n, Kp, threshold, E, R, t, inliers, n_inliers):
        return self.lib.mvo_esti_motion_by_essential(self.ctx, p1, p2, n, Kp, threshold, E, R, t, inliers, n_inliers)

    def esti_motion_by_homography(self, p1, p2, n, Kp, threshold, H, Rs, ts, normals, n_solutions, inliers, n_inliers):
        return self.lib.mvo_esti_motion_by_homography(self.ctx, p1, p2, n, Kp, threshold, H, Rs, ts, normals, n_solutions, inliers, n_inliers)

    def remove_wrong_rt_of_homography(self, np1, np2, n, inliers, n_inliers, Rs, ts, normals, n_solutions):
        return self.lib.mvo_remove_wrong_rt_of_homography(self.ctx, np1, np2, n, inliers, n_inliers, Rs, ts, normals, n_solutions)

    def do_triangulation(self, np1, np2, n, R, t, inliers, n_inliers, pts3d):
        return self.lib.mvo_do_triangulation(self.ctx, np1, np2, n, R, t, inliers, n_inliers, pts3d)


def test_state_machine_over_the_real_entry_points(hostcheck, tmp_path):  # noqa: F811
    """Plumbing check at half resolution (the emulated extraction costs ~13 s per 320x240 frame): every call the state machine
    makes into the real entry points succeeds, extraction and matching agree with the oracle pipeline running beside it on the same
    frames (they are bit-exact), and the two-view results that come back are well-formed.  Whether and when the map gets initialised
    is NOT asserted here: at this resolution the first frame pairs are close to the planar / low-parallax degeneracy of the essential
    matrix and the two RANSACs legitimately settle on different motions (the full-resolution behaviour is the test above)."""
    import emu_build
    import mvo_b200
    from oracle import vo_pipeline_oracle as vp
    lib = C.CDLL(str(emu_build.build(tmp_path, ["ctx.cu", "orb.cu", "orb_host.cpp", "match.cu", "match_host.cpp", "epipolar.cu", "two_view.cpp", "motion_host.cpp"])))
    for name in ("mvo_default_params", "mvo_create", "mvo_destroy", "mvo_last_error", "mvo_orb_extract", "mvo_match_features", "mvo_esti_motion_by_essential",
                 "mvo_esti_motion_by_homography", "mvo_remove_wrong_rt_of_homography", "mvo_do_triangulation"):
        res, args = mvo_b200.SIGNATURES[name]
        # the forwarders hand raw addresses through: every pointer parameter as void*
        getattr(lib, name).restype, getattr(lib, name).argtypes = res, [C.c_void_p if hasattr(a, "contents") or a is C.c_char_p else a for a in args]
    prm = mvo_b200.Params()
    lib.mvo_default_params(C.byref(prm))
    prm.max_keypoints, prm.epi_hypotheses = 1000, 384
    ctx = C.c_void_p()
    assert lib.mvo_create(C.byref(ctx), 0, C.byref(prm)) == 0
    rows, cols = 240, 320
    Kh = K.copy()
    Kh[:2] *= 0.5
    planes = mvo_synth._room_planes(0)
    _, truth = None, []
    frames = []
    rng = np.random.default_rng(2000)
    drift_r = rng.normal(0, 1, 3) * 0.003
    for i in range(5):
        T = np.eye(4)
        T[:3, :3] = mvo_synth.rodrigues(drift_r * i)
        T[:3, 3] = np.array([0.08, -0.012, 0.024]) * i
        truth.append(T)
        frames.append(mvo_synth.render_room(T, planes, Kh, cols, rows))
    cfg = dict(max_number_of_keypoints=1000, ba_iterations=10, min_pixel_dist=25.0)
    oracle = vp.CpuVo(Kh, rows, cols, **cfg)
    helper = vp.CpuVo(Kh, rows, cols, max_number_of_keypoints=1000)
    stages = RealEntryPointStages(helper, 10, lib, ctx)
    hostcheck.hostcheck_set_stages(C.cast(stages.table, C.c_void_p))
    hctx = C.c_void_p(hostcheck.hostcheck_ctx_new(1000))
    p = mvo_b200.VoParams()
    hostcheck.mvo_vo_default_params(C.byref(p))
    p.min_pixel_dist = 25.0
    h = C.c_void_p()
    Kc = np.ascontiguousarray(Kh, np.float64)
    assert hostcheck.mvo_vo_create(hctx, Kc.ctypes.data, rows, cols, C.byref(p), C.byref(h)) == 0
    Tp, To, infos = [], [], []
    for f in frames:
        img = mvo_synth.gray_to_bgr(f)
        T, info = np.zeros(16), mvo_b200.VoFrameInfo()
        assert hostcheck.mvo_vo_add_frame(h, img.ctypes.data, 3, cols * 3, T.ctypes.data, C.byref(info)) == 0, lib.mvo_last_error(ctx)
        To.append(oracle.add_frame(img)[0])
        io = oracle.log[-1]
        Tp.append(T.reshape(4, 4).copy())
        infos.append((info.state_out, info.keyframe, info.best_sol, info.n_keypoints, info.n_matches, info.n_inliers, info.map_points))
        assert info.n_keypoints == io["n_keypoints"]                                  # extraction: bit-exact with the oracle
        if info.state_in == 1 and io["state_in"] == 1:
            assert info.n_matches == io["n_matches"]                                  # and so is the matching against the first keyframe
    hostcheck.mvo_vo_destroy(h)
    hostcheck.hostcheck_ctx_free(hctx)
    lib.mvo_destroy(ctx)
    print(infos)
    assert infos[0][:2] == (1, 1) and all(0 <= i[2] <= 4 for i in infos[1:])         # a two-view solution was chosen on every later frame

"""Entry points of the library EXECUTED ON THE CPU through their real host code: csrc/{ctx,match,epipolar}.cu with
csrc/{match_host,two_view,motion_host}.cpp built for the host (tests/emu_build.py) — the descriptor matcher against the
oracle (bit-exact, as tests/test_match_gpu.py) and the two-view entry points against the OpenCV restatement with the bars of
the hardware tests.  mvo_esti_motion_by_homography / mvo_estimate_relative_poses have not run on a GPU yet: here their host
side (buffer layout, copies, decomposition, the assembly of the solutions) runs for the first time, over the emulated kernels."""
import ctypes as C

import numpy as np
import pytest

import mvo_synth
from conftest import have_cv2

K = mvo_synth.K_DEFAULT


@pytest.fixture(scope="module")
def lib(tmp_path_factory):
    import emu_build
    import mvo_b200
    so = emu_build.build(tmp_path_factory.mktemp("libemu"), ["ctx.cu", "orb.cu", "orb_host.cpp", "match.cu", "match_host.cpp", "epipolar.cu", "two_view.cpp", "motion_host.cpp"])
    lib = C.CDLL(str(so))
    for name in ("mvo_default_params", "mvo_create", "mvo_destroy", "mvo_last_error", "mvo_get_params", "mvo_set_params", "mvo_match_hamming_nn", "mvo_match_hamming_knn2", "mvo_match_radius_sad",
                 "mvo_match_features", "mvo_esti_motion_by_essential", "mvo_esti_motion_by_homography", "mvo_remove_wrong_rt_of_homography",
                 "mvo_do_triangulation", "mvo_estimate_relative_poses"):
        res, args = mvo_b200.SIGNATURES[name]
        getattr(lib, name).restype, getattr(lib, name).argtypes = res, args
    return lib


@pytest.fixture(scope="module")
def ctx(lib):
    import mvo_b200
    p = mvo_b200.Params()
    lib.mvo_default_params(C.byref(p))
    p.epi_hypotheses = 384
    h = C.c_void_p()
    assert lib.mvo_create(C.byref(h), 0, C.byref(p)) == 0
    yield h
    lib.mvo_destroy(h)


def test_emulated_matcher_is_bit_exact(lib, ctx):
    import mvo_b200
    from oracle import oracle_lib
    d1 = mvo_synth.random_descriptors(1, 300, dup_every=7)
    d2 = mvo_synth.random_descriptors(2, 257, dup_every=5)
    out = np.zeros(300, mvo_b200.DMATCH_DTYPE)
    assert lib.mvo_match_hamming_nn(ctx, d1.ctypes.data, 300, d2.ctypes.data, 257, out.ctypes.data) == 0, lib.mvo_last_error(ctx)
    ref = oracle_lib.hamming_nn(d1, d2)
    assert np.array_equal(out["train_idx"], ref["train_idx"]) and np.array_equal(out["distance"], ref["distance"])
    out2 = np.zeros(600, mvo_b200.DMATCH_DTYPE)
    assert lib.mvo_match_hamming_knn2(ctx, d1.ctypes.data, 300, d2.ctypes.data, 257, out2.ctypes.data) == 0
    ref2 = oracle_lib.hamming_knn2(d1, d2).ravel()
    assert np.array_equal(out2["train_idx"], ref2["train_idx"]) and np.array_equal(out2["distance"], ref2["distance"])
    rng = np.random.default_rng(0)
    xy1 = rng.uniform(0, 640, (300, 2)).astype(np.float32)
    xy2 = (xy1[rng.integers(0, 300, 257)] + rng.normal(0, 20, (257, 2))).astype(np.float32)
    for method in (1, 2, 3):
        m, n = np.zeros(300, mvo_b200.DMATCH_DTYPE), C.c_int(0)
        assert lib.mvo_match_features(ctx, d1.ctypes.data, 300, d2.ctypes.data, 257, method, xy1.ctypes.data, xy2.ctypes.data, 50.0, m.ctypes.data, C.byref(n)) == 0
        refm = oracle_lib.match_features(d1, d2, method, xy1, xy2, 50.0)
        assert n.value == len(refm) and m[: n.value].tobytes() == refm.tobytes(), method


def _rod(r):
    th = np.linalg.norm(r)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def _scene(seed, planar, n=320):
    rng = np.random.default_rng(seed)
    R = _rod(rng.normal(0, 0.05, 3) + 1e-9)
    t = np.array([0.3, 0.02, 0.06])
    nrm = np.array([0.05, -0.08, 1.0])
    nrm /= np.linalg.norm(nrm)
    P = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(2.5, 8, n)], 1)
    if planar:
        P[:, 2] = (4.0 - P[:, :2] @ nrm[:2]) / nrm[2]
    P2 = P @ R.T + t
    q1, q2 = P[:, :2] / P[:, 2:3] * K[0, 0] + K[:2, 2], P2[:, :2] / P2[:, 2:3] * K[0, 0] + K[:2, 2]
    p1 = (q1 + rng.normal(0, 0.4, (n, 2))).astype(np.float32)
    p2 = q2 + rng.normal(0, 0.4, (n, 2))
    bad = rng.random(n) < 0.15
    p2[bad] = rng.uniform([0, 0], [640, 480], (bad.sum(), 2))
    return p1, p2.astype(np.float32), R, t / np.linalg.norm(t), nrm, ~bad, P, q1, q2


@pytest.mark.skipif(not have_cv2(), reason="cv2 not importable")
def test_emulated_homography_entry_points(lib, ctx):
    from oracle import epipolar_oracle
    p1, p2, R, td, nrm, good, P, q1, q2 = _scene(3, True)
    n = len(p1)
    Kc = np.ascontiguousarray(K, np.float64)
    H, Rs, ts, ns = np.zeros((3, 3)), np.zeros((4, 3, 3)), np.zeros((4, 3)), np.zeros((4, 3))
    nsol, inl, ni = C.c_int(0), np.zeros(n, np.int32), C.c_int(n)
    rc = lib.mvo_esti_motion_by_homography(ctx, p1.ctypes.data, p2.ctypes.data, n, Kc.ctypes.data, 3.0, H.ctypes.data, Rs.ctypes.data, ts.ctypes.data,
                                           ns.ctypes.data, C.byref(nsol), inl.ctypes.data, C.byref(ni))
    assert rc == 0, lib.mvo_last_error(ctx)
    inl = inl[: ni.value]
    Ho, Rso, tso, nso, inlo = epipolar_oracle.esti_motion_by_homography(p1, p2, K, 3.0)

    def terr(Hm):
        m = np.c_[q1, np.ones(n)] @ Hm.T
        return np.sqrt(np.mean(np.sum((m[:, :2] / m[:, 2:3] - q2) ** 2, 1)))
    assert abs(H[2, 2] - 1) < 1e-12 and nsol.value == 4 and terr(H) < terr(Ho) + 0.1
    assert np.all(np.diff(inl) > 0) and len(set(inl.tolist()) & set(inlo.tolist())) / len(inlo) > 0.9 and good[inl].mean() > 0.97
    err = [max(np.abs(Rs[i] - R).max(), np.abs(ts[i] - td).max(), np.abs(ns[i] - nrm).max()) for i in range(4)]
    assert min(err) < 0.02, err                                      # the true motion is among the four solutions
    for i in range(4):
        assert abs(np.linalg.det(Rs[i]) - 1) < 1e-8 and abs(np.linalg.norm(ts[i]) - 1) < 1e-9
    # removeWrongRtOfHomography: the same survivors as OpenCV's filter
    Ki = np.linalg.inv(K)
    np1 = ((np.c_[p1, np.ones(n)] @ Ki.T)[:, :2]).astype(np.float32)
    np2 = ((np.c_[p2, np.ones(n)] @ Ki.T)[:, :2]).astype(np.float32)
    keep = epipolar_oracle.remove_wrong_rt_of_homography(np1, np2, inl, list(Rs), list(ts), list(ns))
    R2, t2, n2, k2 = Rs.copy(), ts.copy(), ns.copy(), C.c_int(4)
    assert lib.mvo_remove_wrong_rt_of_homography(ctx, np1.ctypes.data, np2.ctypes.data, n, inl.ctypes.data, len(inl), R2.ctypes.data, t2.ctypes.data,
                                                 n2.ctypes.data, C.byref(k2)) == 0
    assert k2.value == len(keep) >= 1 and all(np.array_equal(R2[j], Rs[k]) for j, k in enumerate(keep)) and int(np.argmin(err)) in keep
    # too few points: the reference's OpenCV call would throw; MVO_ERR_DEGENERATE here
    ni2 = C.c_int(3)
    assert lib.mvo_esti_motion_by_homography(ctx, p1.ctypes.data, p2.ctypes.data, 3, Kc.ctypes.data, 3.0, H.ctypes.data, Rs.ctypes.data, ts.ctypes.data,
                                             ns.ctypes.data, C.byref(nsol), inl.ctypes.data, C.byref(ni2)) == -6


@pytest.mark.skipif(not have_cv2(), reason="cv2 not importable")
@pytest.mark.parametrize("planar", [False, True])
def test_emulated_estimate_relative_poses(lib, ctx, planar):
    import mvo_b200
    p1, p2, R, td, nrm, good, P, q1, q2 = _scene(40 + planar, planar)
    n = len(p1)
    Kc = np.ascontiguousarray(K, np.float64)
    sol = mvo_b200.TwoViewSolutions()
    inl, pts = np.zeros((5, n), np.int32), np.zeros((5, n, 3), np.float32)
    rc = lib.mvo_estimate_relative_poses(ctx, p1.ctypes.data, p2.ctypes.data, n, Kc.ctypes.data, 1, 1, C.byref(sol), inl.ctypes.data, pts.ctypes.data)
    assert rc == 0, lib.mvo_last_error(ctx)
    assert 1 <= sol.num_solutions <= 5 and 0 <= sol.best < sol.num_solutions
    R0, t0 = np.array(sol.R[0]).reshape(3, 3), np.array(sol.t[0])
    if not planar:
        assert sol.best == 0 and sol.ratio < 0.5
        assert np.abs(R0 - R).max() < 5e-3 and np.arccos(np.clip(t0 @ td, -1, 1)) < 0.05
        k = sol.n_inliers[0]
        X = pts[0, :k] * np.linalg.norm([0.3, 0.02, 0.06])            # unit baseline -> true scale
        Pi = P[inl[0, :k]]
        assert np.median(np.linalg.norm(X - Pi, axis=1) / np.linalg.norm(Pi, axis=1)) < 0.05
    else:
        assert sol.num_solutions >= 2 and np.allclose(np.array(sol.normal[0]), 0)
        errs = [max(np.abs(np.array(sol.R[s]).reshape(3, 3) - R).max(), np.abs(np.array(sol.t[s]) - td).max()) for s in range(1, sol.num_solutions)]
        assert min(errs) < 0.03, errs
        assert sol.score_h > 0 and sol.score_e > 0
        # Where the decision falls: OpenCV's un-refined five-point model keeps fewer inliers on a plane than the locally optimised
        # essential matrix here, so the reference's H/(E+H) comes out at 0.50-0.52 on such scenes and this one's at 0.485-0.492 —
        # on the E side of the reference's 0.5, on the H side of the 0.45 its README documents (mvo_params::eh_ratio_threshold).
        assert 0.46 < sol.ratio < 0.5 and sol.best == 0
        prm = mvo_b200.Params()
        assert lib.mvo_get_params(ctx, C.byref(prm)) == 0
        prm.eh_ratio_threshold = 0.45
        assert lib.mvo_set_params(ctx, C.byref(prm)) == 0
        try:
            sol2 = mvo_b200.TwoViewSolutions()
            assert lib.mvo_estimate_relative_poses(ctx, p1.ctypes.data, p2.ctypes.data, n, Kc.ctypes.data, 1, 1, C.byref(sol2), inl.ctypes.data, pts.ctypes.data) == 0
            assert sol2.best >= 1 and abs(sol2.ratio - sol.ratio) < 1e-12
            Rb, tb, nb = np.array(sol2.R[sol2.best]).reshape(3, 3), np.array(sol2.t[sol2.best]), np.array(sol2.normal[sol2.best])
            assert np.abs(Rb - R).max() < 0.03 and np.abs(tb - td).max() < 0.03 and abs(abs(nb[2]) - abs(nrm[2])) < 0.03      # the true motion, the most frontal normal
        finally:
            prm.eh_ratio_threshold = 0.5
            assert lib.mvo_set_params(ctx, C.byref(prm)) == 0


def test_emulated_grid_selection_formulations(lib, ctx):
    """The per-cell counter of selectUniformKptsByGrid in its two device formulations (csrc/orb.cu: grid_rank_keep_seg /
    grid_rank_keep) against the sequential rule, executed on the CPU tier (cases of tests/test_orb_gpu.py)."""
    from test_orb_gpu import check_grid_rank
    check_grid_rank(lib, ctx)

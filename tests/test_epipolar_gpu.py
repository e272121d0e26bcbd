"""GPU parity of the two-view stage (first part) against the reference's arithmetic
(oracle/epipolar_oracle.py = cv2.findEssentialMat + cv2.recoverPose + cv2.triangulatePoints, the calls of reference
src/geometry/epipolar_geometry.cpp:17-57,130-175).  The RANSACs differ by design (batched eight-point hypotheses vs
OpenCV's adaptive five-point), so the bar is the one used for PnP: the pose must be as close to the synthetic truth
as OpenCV's (within a stated factor and floor), and the consensus sets must agree (Jaccard)."""
import numpy as np
import pytest
from conftest import have_cv2

import mvo_synth

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not have_cv2(), reason="cv2 not importable")]
K = mvo_synth.K_DEFAULT


def _rodrigues(r):
    th = np.linalg.norm(r)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def _scene(seed, n=1500, noise=0.5, outliers=0.2):
    rng = np.random.default_rng(seed)
    R = _rodrigues(rng.normal(0, 0.06, 3) + 1e-9)
    t = np.array([0.25, 0.02, 0.05]) + rng.normal(0, 0.02, 3)
    X1 = np.stack([rng.uniform(-2.2, 2.2, n), rng.uniform(-1.6, 1.6, n), rng.uniform(2.5, 9, n)], 1)
    X2 = X1 @ R.T + t
    p1 = X1[:, :2] / X1[:, 2:3] * K[0, 0] + K[:2, 2]
    p2 = X2[:, :2] / X2[:, 2:3] * K[0, 0] + K[:2, 2]
    p1 += rng.normal(0, noise, p1.shape)
    p2 += rng.normal(0, noise, p2.shape)
    bad = rng.random(n) < outliers
    p2[bad] = rng.uniform([0, 0], [640, 480], (bad.sum(), 2))
    return p1.astype(np.float32), p2.astype(np.float32), R, t / np.linalg.norm(t), ~bad, X1


def _rot_err(Ra, Rb):
    return np.arccos(np.clip((np.trace(Ra.T @ Rb) - 1) / 2, -1, 1))


def _dir_err(a, b):
    return np.arccos(np.clip(np.dot(a, b) / (np.linalg.norm(a) * np.linalg.norm(b)), -1, 1))


@pytest.mark.parametrize("seed", [0, 1, 2, 3])
def test_essential_pose_and_inliers_vs_opencv(ctx, seed):
    from oracle import epipolar_oracle
    p1, p2, R_true, t_true, good, _ = _scene(seed)
    E, R, t, inl = ctx.esti_motion_by_essential(p1, p2, K, threshold=1.0)
    Eo, Ro, to, inlo = epipolar_oracle.esti_motion_by_essential(p1, p2, K, 0.999, 1.0)
    # same conventions: E[2,2] = 1, |t| = 1, proper rotation, x2 ~ R x1 + t
    assert abs(E[2, 2] - 1) < 1e-12 and abs(np.linalg.norm(t) - 1) < 1e-12
    assert abs(np.linalg.det(R) - 1) < 1e-9 and np.allclose(R @ R.T, np.eye(3), atol=1e-9)
    # pose against the synthetic truth: no worse than 2x OpenCV's error, with a floor of 3 mrad / 2 degrees of direction
    er, ero = _rot_err(R, R_true), _rot_err(Ro, R_true)
    et, eto = _dir_err(t, t_true), _dir_err(to, t_true)
    assert er < max(2 * ero, 3e-3), (er, ero)
    assert et < max(2 * eto, 0.035), (et, eto)
    # the essential matrices describe the same epipolar geometry (both are estimates: OpenCV's translation direction alone
    # is off by up to 0.08 rad on these scenes)
    En, Eon = E / np.linalg.norm(E), Eo / np.linalg.norm(Eo)
    assert min(np.abs(En - Eon).max(), np.abs(En + Eon).max()) < 0.12
    # consensus sets: ascending indices, true inliers, and OpenCV's set (the consensus of its un-refined minimal model,
    # 75-90 % of the true inliers here) is contained in ours
    assert np.all(np.diff(inl) > 0) and inl.min() >= 0 and inl.max() < len(p1)
    a, b = set(inl.tolist()), set(inlo.tolist())
    assert len(a & b) / len(b) > 0.9, (len(a), len(b), len(a & b))
    assert good[inl].mean() > 0.97 and len(inl) > 0.6 * good.sum()


def test_essential_degenerate_inputs(ctx):
    import mvo_b200
    p = np.zeros((5, 2), np.float32)
    with pytest.raises(mvo_b200.MvoError):
        ctx.esti_motion_by_essential(p, p, K)                       # fewer than 8 correspondences
    rng = np.random.default_rng(0)
    p1 = rng.uniform([0, 0], [640, 480], (200, 2)).astype(np.float32)
    p2 = rng.uniform([0, 0], [640, 480], (200, 2)).astype(np.float32)
    try:                                                            # pure noise: either a tiny consensus set or a clean refusal
        E, R, t, inl = ctx.esti_motion_by_essential(p1, p2, K)
        assert len(inl) < 60
    except mvo_b200.MvoError as e:
        assert e.code == -6


@pytest.mark.parametrize("seed", [0, 5])
def test_triangulation_vs_opencv(ctx, seed):
    from oracle import epipolar_oracle
    p1, p2, R_true, t_true, good, X1 = _scene(seed, n=800, noise=0.3, outliers=0.1)
    Ki = np.linalg.inv(K)
    np1 = ((np.c_[p1, np.ones(len(p1))] @ Ki.T)[:, :2]).astype(np.float32)
    np2 = ((np.c_[p2, np.ones(len(p2))] @ Ki.T)[:, :2]).astype(np.float32)
    inl = np.nonzero(good)[0].astype(np.int32)[::2]
    t_scaled = t_true * 0.26
    got = ctx.do_triangulation(np1, np2, R_true, t_scaled, inl)
    ref = epipolar_oracle.do_triangulation(np1, np2, R_true, t_scaled, inl)
    assert got.shape == ref.shape == (len(inl), 3) and got.dtype == np.float32
    assert np.abs(got - ref).max() < 2e-4 * max(1.0, np.abs(ref).max())
    assert np.median(np.linalg.norm(got - X1[inl], axis=1) / np.linalg.norm(X1[inl], axis=1)) < 0.05
    assert ctx.do_triangulation(np1, np2, R_true, t_scaled, np.zeros(0, np.int32)).shape == (0, 3)


def test_device_math_equals_host_build(ctx):
    """csrc/epipolar_math.cuh on the device (one thread, mvo_test_epi_math) against numpy: the routines the CPU tier checks
    through the host build (tests/test_epipolar_math.py) must behave the same inside a kernel."""
    import ctypes as C
    lib = ctx.lib
    rng = np.random.default_rng(1)
    for rep in range(5):
        R = _rodrigues(rng.normal(0, 0.15, 3))
        t = rng.normal(0, 1, 3)
        t /= np.linalg.norm(t)
        X1 = np.stack([rng.uniform(-2, 2, 8), rng.uniform(-1.5, 1.5, 8), rng.uniform(3, 9, 8)], 1)
        X2 = X1 @ R.T + 0.4 * t
        x1, x2 = X1[:, :2] / X1[:, 2:3], X2[:, :2] / X2[:, 2:3]
        M3, M89 = rng.normal(0, 1, (3, 3)), rng.normal(0, 1, (8, 9))
        inp = np.concatenate([x1.ravel(), x2.ravel(), M3.ravel(), M89.ravel()])
        out = np.zeros(64)
        assert lib.mvo_test_epi_math(ctx.h, inp.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p)) == 0
        assert out[0] == 1 and out[1] == 0
        E = out[2:11].reshape(3, 3)
        tx = np.array([[0, -t[2], t[1]], [t[2], 0, -t[0]], [-t[1], t[0], 0]])
        Et = tx @ R
        Et /= np.linalg.norm(Et)
        En = E / np.linalg.norm(E)
        assert min(np.abs(En - Et).max(), np.abs(En + Et).max()) < 1e-7
        U, sv, V = out[11:20].reshape(3, 3), out[20:23], out[23:32].reshape(3, 3)
        assert np.allclose(U @ np.diag(sv) @ V.T, M3, atol=1e-10) and np.allclose(sv, np.linalg.svd(M3, compute_uv=False), atol=1e-8)
        assert out[32] == 1 and np.abs(M89 @ out[33:42]).max() < 1e-10
        assert np.allclose(np.sort(out[42:45]), np.linalg.eigvalsh(M3 + M3.T), atol=1e-9)

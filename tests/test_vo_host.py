"""Host-side initialisation / keyframe decisions (csrc/vo_host.cpp) against the numpy restatement of reference
src/vo/vo.cpp:96-265 (oracle/vo_init_oracle.py)."""
import ctypes as C

import numpy as np
import pytest


def _pose(rng, t_scale=0.3):
    r = rng.normal(0, 0.08, 3)
    th = np.linalg.norm(r)
    k = r / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    T = np.eye(4)
    T[:3, :3] = np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx
    T[:3, 3] = rng.normal(0, t_scale, 3)
    return T


@pytest.mark.parametrize("seed", [0, 1, 2])
def test_retain_good_triangulation(built, seed):
    import mvo_b200
    from oracle import vo_init_oracle as o
    lib = mvo_b200.load_library()
    rng = np.random.default_rng(seed)
    n = 500
    Tc, Tr = _pose(rng), np.eye(4)
    pts = np.stack([rng.uniform(-3, 3, n), rng.uniform(-2, 2, n), rng.uniform(1, 60, n)], 1).astype(np.float32)
    keep, ang, cnt = np.zeros(n, np.int32), np.zeros(n), C.c_int(-1)
    assert lib.mvo_retain_good_triangulation(pts.ctypes.data, n, Tc.ctypes.data, Tr.ctypes.data, 1.0, 20.0, keep.ctypes.data, ang.ctypes.data, C.byref(cnt)) == 0
    rk, ra = o.retain_good_triangulation(pts, Tc, Tr, 1.0, 20.0)
    assert 0 < cnt.value < n                                          # both rejection rules fire on this depth range
    assert np.array_equal(keep[: cnt.value], rk)
    assert np.allclose(ang[: cnt.value], ra, rtol=0, atol=1e-9)
    # empty input is the reference's early return
    assert lib.mvo_retain_good_triangulation(None, 0, Tc.ctypes.data, Tr.ctypes.data, 1.0, 20.0, None, None, C.byref(cnt)) == 0 and cnt.value == 0
    assert lib.mvo_retain_good_triangulation(pts.ctypes.data, n, None, Tr.ctypes.data, 1.0, 20.0, keep.ctypes.data, ang.ctypes.data, C.byref(cnt)) != 0


def test_normalize_init_depth(built):
    import mvo_b200
    from oracle import vo_init_oracle as o
    lib = mvo_b200.load_library()
    rng = np.random.default_rng(3)
    pts = np.stack([rng.uniform(-3, 3, 300), rng.uniform(-2, 2, 300), rng.uniform(1, 9, 300)], 1).astype(np.float32)
    t = np.array([0.7, -0.1, 0.2])
    rp, rt, rs = o.normalize_init_depth(pts, t, 5.0)
    p, tt, s = pts.copy(), t.copy(), C.c_double(0)
    assert lib.mvo_normalize_init_depth(p.ctypes.data, 300, tt.ctypes.data, 5.0, C.byref(s)) == 0
    assert np.array_equal(p, rp) and np.allclose(tt, rt, rtol=0, atol=1e-15) and abs(s.value - rs) < 1e-15
    assert abs(float(p[:, 2].astype(float).mean()) - 5.0) < 1e-5
    assert lib.mvo_normalize_init_depth(p.ctypes.data, 0, tt.ctypes.data, 5.0, None) != 0


@pytest.mark.parametrize("case", ["good", "few", "still", "flat", "no_angles", "no_matches"])
def test_is_vo_good_to_init(built, case):
    import mvo_b200
    from oracle import vo_init_oracle as o
    lib = mvo_b200.load_library()
    rng = np.random.default_rng(11)
    n = 30 if case == "few" else (0 if case == "no_matches" else 200)
    a = rng.uniform(0, 600, (n, 2)).astype(np.float32)
    b = (a + rng.normal(0, 0.5 if case == "still" else 40, (n, 2))).astype(np.float32)
    ang = np.zeros(0) if case == "no_angles" else rng.uniform(0.1, 0.9 if case == "flat" else 6.0, 150)
    good, md, med = C.c_int(-1), C.c_double(0), C.c_double(0)
    assert lib.mvo_is_vo_good_to_init(a.ctypes.data if n else None, b.ctypes.data if n else None, n, ang.ctypes.data if len(ang) else None, len(ang), 50, 50.0, 2.0,
                                      C.byref(good), C.byref(md), C.byref(med)) == 0
    rg, rmd, rmed = o.is_vo_good_to_init(a, b, ang, 50, 50.0, 2.0)
    assert bool(good.value) == rg == (case == "good")
    assert (np.isnan(md.value) and np.isnan(rmd)) or abs(md.value - rmd) < 1e-9
    assert abs(med.value - rmed) < 1e-15


def test_check_large_move(built):
    import mvo_b200
    from oracle import vo_init_oracle as o
    lib = mvo_b200.load_library()
    rng = np.random.default_rng(5)
    seen = set()
    for i in range(20):
        Tr, Tc = _pose(rng, 1.0), _pose(rng, 1.0)
        if i % 2:
            Tc = Tr @ _pose(rng, 0.03)
        large, d, a = C.c_int(-1), C.c_double(0), C.c_double(0)
        assert lib.mvo_check_large_move(Tc.ctypes.data, Tr.ctypes.data, 0.1, C.byref(large), C.byref(d), C.byref(a)) == 0
        rl, rd, ra = o.check_large_move(Tc, Tr, 0.1)
        assert bool(large.value) == rl and abs(d.value - rd) < 1e-12 and abs(a.value - ra) < 1e-7
        seen.add(rl)
    assert seen == {True, False}
    assert lib.mvo_check_large_move(None, None, 0.1, C.byref(large), None, None) != 0


def test_room_loop_trajectory_stays_inside_the_room():
    """The long-sequence generator for BASELINE config 5 (>= 150 frames per sequence): closed loop, ~0.04 m per frame, inside the walls."""
    import mvo_synth
    poses = mvo_synth.room_loop_poses(3, 240)
    c = np.array([T[:3, 3] for T in poses])
    steps = np.linalg.norm(np.diff(c, axis=0), axis=1)
    r = mvo_synth._ROOM
    assert 0.03 < steps.min() and steps.max() < 0.06
    assert np.abs(c[:, 0]).max() < r["half_w"] - 0.5 and r["top"] + 0.5 < c[:, 1].min() and c[:, 1].max() < r["bottom"] - 0.5
    assert r["z0"] + 1.0 < c[:, 2].min() and c[:, 2].max() < r["back"] - 3.0
    assert np.abs(c[120] - c[0]).max() < 1e-9                       # period 120: the loop closes
    for T in poses[::17]:
        assert np.abs(T[:3, :3] @ T[:3, :3].T - np.eye(3)).max() < 1e-12

"""k_match_filter (csrc/track.cu) — the thresholds of geometry::matchFeatures (feature_match.cpp:179-217) and
removeDuplicatedMatches (:241-260: unstable libstdc++ std::sort by trainIdx + first of every run) restated on the
device — against the host implementation that runs the real std::sort, from identical packed matcher keys.
Bit-exact: which of several map points matched to one keypoint survives is decided by the introsort's partition
sequence, so the inputs stress that: heavy duplication, presorted / reversed / constant / organ-pipe / sawtooth key
sequences, sizes around the 16-element insertion-sort cutoff, all three match methods, partial visibility."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _run(ctx, which, keys, vis, nk, method):
    import mvo_b200
    lib = mvo_b200.load_library()
    fn = getattr(lib, f"mvo_test_match_filter_{which}")
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    nmap = len(vis)
    pairs = np.full((max(nmap, 1), 2), -1, np.int32)
    info = np.zeros(16, np.int32)
    keys = np.ascontiguousarray(keys, np.uint32)
    vis = np.ascontiguousarray(vis, np.uint8)
    rc = fn(ctx.h, keys.ctypes.data, vis.ctypes.data, nmap, nk, method, pairs.ctypes.data, info.ctypes.data)
    assert rc == 0, ctx.lib.mvo_last_error(ctx.h)
    return pairs[: info[0]].copy(), info[:3].copy()


def _train_patterns(rng, n, nk):
    yield "random", rng.integers(0, nk, n)
    yield "heavy-dup", rng.integers(0, max(1, nk // 7), n)
    yield "two-values", rng.integers(0, 2, n) * (nk - 1)
    yield "constant", np.full(n, nk // 2)
    yield "ascending", np.sort(rng.integers(0, nk, n))
    yield "descending", np.sort(rng.integers(0, nk, n))[::-1]
    yield "organ-pipe", np.minimum(np.arange(n), np.arange(n)[::-1]) % nk
    yield "sawtooth", (np.arange(n) * 37) % max(1, nk // 3)
    yield "blocks", np.repeat(rng.permutation(max(1, n // 5 + 1)), 5)[:n] % nk


@pytest.mark.parametrize("nmap", [1, 2, 15, 16, 17, 18, 33, 100, 517, 2001, 4096, 8000])
def test_filter_equals_host_std_sort_method1(ctx, nmap):
    rng = np.random.default_rng(nmap)
    nk = min(max(2, nmap if nmap < 100 else nmap // 2 + 3), 8192)
    for name, train in _train_patterns(rng, nmap, nk):
        for vis_frac in (1.0, 0.7):
            dist = rng.integers(0, 90, nmap).astype(np.uint32)       # Hamming distances; threshold = max(2*min, 30)
            keys = (dist << 16) | train.astype(np.uint32)
            vis = (rng.random(nmap) < vis_frac).astype(np.uint8)
            ph, ih = _run(ctx, "host", keys, vis, nk, 1)
            pd, idv = _run(ctx, "dev", keys, vis, nk, 1)
            # median-of-3 killers (organ pipe, concatenated ascending runs) take libstdc++'s quicksort phase to its depth limit
            # 2*log2(n): the kernel heapsorts the remaining segments like std::__partial_sort does (round 2; it declined before)
            assert idv[2] == 0, (name, nmap)
            assert np.array_equal(ih, idv), (name, nmap, ih, idv)
            assert np.array_equal(ph, pd), (name, nmap, vis_frac)
            assert np.all(np.diff(pd[:, 1]) > 0)                    # sorted by keypoint index, unique


def test_filter_method2_and_method3(ctx):
    rng = np.random.default_rng(5)
    nmap, nk = 2001, 1900
    ctx.set_params(lowe_ratio=0.8)
    try:
        for rep in range(6):
            vis = (rng.random(nmap) < 0.9).astype(np.uint8)
            # method 2: two keys per query (best, second best)
            d0 = rng.integers(0, 80, nmap).astype(np.uint32)
            d1 = d0 + rng.integers(0, 40, nmap).astype(np.uint32)
            k = np.empty(2 * nmap, np.uint32)
            k[0::2] = (d0 << 16) | rng.integers(0, nk // 3, nmap).astype(np.uint32)
            k[1::2] = (d1 << 16) | rng.integers(0, nk, nmap).astype(np.uint32)
            ph, ih = _run(ctx, "host", k, vis, nk, 2)
            pd, idv = _run(ctx, "dev", k, vis, nk, 2)
            assert np.array_equal(ih, idv) and np.array_equal(ph, pd) and 0 < ih[0] < nmap
            # method 3: SAD keys, queries without a keypoint inside the radius carry 0xFFFFFFFF
            sad = rng.integers(0, 32 * 60, nmap).astype(np.uint32)
            k3 = (sad << 16) | rng.integers(0, nk // 2, nmap).astype(np.uint32)
            k3[rng.random(nmap) < 0.2] = 0xFFFFFFFF
            ph, ih = _run(ctx, "host", k3, vis, nk, 3)
            pd, idv = _run(ctx, "dev", k3, vis, nk, 3)
            assert np.array_equal(ih, idv) and np.array_equal(ph, pd) and ih[0] > 0
    finally:
        ctx.set_params(lowe_ratio=1.0)


def test_filter_empty_and_over_capacity(ctx):
    rng = np.random.default_rng(9)
    # nothing visible / nothing matched
    nmap, nk = 300, 280
    keys = ((rng.integers(0, 60, nmap) << 16) | rng.integers(0, nk, nmap)).astype(np.uint32)
    pd, idv = _run(ctx, "dev", keys, np.zeros(nmap, np.uint8), nk, 1)
    assert tuple(idv) == (0, 0, 0) and len(pd) == 0
    pd, idv = _run(ctx, "dev", np.full(nmap, 0xFFFFFFFF, np.uint32), np.ones(nmap, np.uint8), nk, 1)
    assert tuple(idv) == (0, nmap, 0)
    # a match list beyond the device capacity is declined (status 1), the tracker then takes the host filter
    nmap, nk = 9000, 2000
    keys = ((np.full(nmap, 5) << 16) | rng.integers(0, nk, nmap)).astype(np.uint32)
    _, idv = _run(ctx, "dev", keys, np.ones(nmap, np.uint8), nk, 1)
    assert idv[2] == 1 and idv[1] == nmap


def test_tracker_host_filter_variant_gives_identical_poses(ctx, monkeypatch):
    """The fused single-synchronisation frame and the variant that filters the match list on the host (taken when
    the device filter declines) must produce identical results."""
    import subprocess
    import sys
    from pathlib import Path
    root = Path(__file__).resolve().parent.parent
    code = (
        "import sys, numpy as np\n"
        f"sys.path.insert(0, r'{root}'); sys.path.insert(0, r'{root / 'monocular-visual-odometry_b200' / 'python'}')\n"
        "import mvo_b200, mvo_synth\n"
        "K = mvo_synth.K_DEFAULT\n"
        "ctx = mvo_b200.Context(0, max_keypoints=2000, ba_iterations=10)\n"
        "frames, _, _ = mvo_synth.planar_sequence(6, n_frames=7, plane_z=4.0)\n"
        "imgs = [mvo_synth.gray_to_bgr(f) for f in frames]\n"
        "kp, desc = ctx.orb_extract(imgs[0])\n"
        "rays = (np.linalg.inv(K) @ np.stack([kp['x'], kp['y'], np.ones(len(kp))]).astype(np.float64)).T\n"
        "pts = (rays * (4.0 / rays[:, 2:3])).astype(np.float32)\n"
        "perm = np.random.default_rng(3).permutation(len(pts)); pts = pts[perm]; desc = np.ascontiguousarray(desc[perm])\n"
        "trk = mvo_b200.Tracker(ctx, K, 480, 640)\n"
        "trk.set_map(pts, desc); trk.reset(np.eye(4))\n"
        "out = []\n"
        "for im in imgs[1:]:\n"
        "    T, r = trk.track(im)\n"
        "    out.append(np.concatenate([T.ravel(), [r.n_candidates, r.n_matches, r.n_inliers, r.pnp_ok, r.ba_frames, r.ba_edges]]))\n"
        "np.save(sys.argv[1], np.array(out))\n")
    import os
    import tempfile
    res = []
    for env_extra in ({}, {"MVO_TRACK_HOST_FILTER": "1"}):
        with tempfile.TemporaryDirectory() as td:
            f = os.path.join(td, "o.npy")
            env = dict(os.environ, **env_extra)
            subprocess.run([sys.executable, "-c", code, f], check=True, env=env, timeout=300)
            res.append(np.load(f))
    assert res[0].shape == res[1].shape and res[0][:, 19].min() == 1
    assert np.array_equal(res[0][:, 16:], res[1][:, 16:])
    assert np.abs(res[0][:, :16] - res[1][:, :16]).max() < 1e-8

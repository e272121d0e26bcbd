"""k_match_filter (csrc/track_filter.cuh) EXECUTED ON THE CPU through tests/cpp/cuda_emu.h against the library's host filter
(the real libstdc++ std::sort), from identical packed matcher keys — the CPU-tier counterpart of
tests/test_track_filter_gpu.py on a subset of its sizes and key patterns (one OS thread plays one CUDA thread)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "monocular-visual-odometry_b200"


@pytest.fixture(scope="module")
def emu(built, tmp_path_factory):
    so = tmp_path_factory.mktemp("filteremu") / "libtrack_filter_emu.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", str(ROOT / "include"), "-I", str(PKG / "csrc"),
                    "-I", str(ROOT / "tests" / "cpp"), "-I", "/usr/local/cuda/include", str(ROOT / "tests" / "cpp" / "track_filter_emu.cpp"),
                    "-L", str(PKG), "-lmvo", f"-Wl,-rpath,{PKG}", "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    lib.emu_params_ctx.restype = C.c_void_p
    lib.emu_params_ctx.argtypes = [C.c_double, C.c_double]
    lib.emu_params_ctx_free.argtypes = [C.c_void_p]
    lib.emu_match_filter.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_double, C.c_double, C.c_void_p, C.c_void_p]
    return lib


def _both(emu, keys, vis, nk, method, xg=2.0, lowe=1.0):
    import mvo_b200
    lib = mvo_b200.load_library()
    lib.mvo_test_match_filter_host.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    nmap = len(vis)
    keys, vis = np.ascontiguousarray(keys, np.uint32), np.ascontiguousarray(vis, np.uint8)
    ph, ih = np.full((max(nmap, 1), 2), -1, np.int32), np.zeros(16, np.int32)
    ctx = C.c_void_p(emu.emu_params_ctx(xg, lowe))
    try:
        assert lib.mvo_test_match_filter_host(ctx, keys.ctypes.data, vis.ctypes.data, nmap, nk, method, ph.ctypes.data, ih.ctypes.data) == 0
    finally:
        emu.emu_params_ctx_free(ctx)
    pd, idv = np.full((max(nmap, 1), 2), -1, np.int32), np.zeros(64, np.int32)
    assert emu.emu_match_filter(keys.ctypes.data, vis.ctypes.data, nmap, nk, method, xg, lowe, pd.ctypes.data, idv.ctypes.data) == 0
    _both.last_info = idv.copy()        # the kernel's full record ([3] segments heapsorted at the depth limit, [7] the longest)
    return ph[: ih[0]], ih[:3], pd[: idv[0]], idv[:3]


@pytest.mark.parametrize("nmap,pattern", [(1, "random"), (16, "random"), (17, "heavy-dup"), (33, "constant"), (100, "two-values"), (517, "random"),
                                          (517, "descending"), (2001, "random"), (2001, "heavy-dup"), (2001, "ascending")])
def test_emulated_filter_equals_host_std_sort(emu, nmap, pattern):
    rng = np.random.default_rng(nmap * 7 + len(pattern))
    nk = min(max(2, nmap if nmap < 100 else nmap // 2 + 3), 8192)
    train = {"random": lambda: rng.integers(0, nk, nmap), "heavy-dup": lambda: rng.integers(0, max(1, nk // 7), nmap),
             "two-values": lambda: rng.integers(0, 2, nmap) * (nk - 1), "constant": lambda: np.full(nmap, nk // 2),
             "ascending": lambda: np.sort(rng.integers(0, nk, nmap)), "descending": lambda: np.sort(rng.integers(0, nk, nmap))[::-1]}[pattern]()
    dist = rng.integers(0, 90, nmap).astype(np.uint32)
    keys = (dist << 16) | train.astype(np.uint32)
    vis = (rng.random(nmap) < 0.8).astype(np.uint8)
    ph, ih, pd, idv = _both(emu, keys, vis, nk, 1)
    assert idv[2] == 0 and np.array_equal(ih, idv) and np.array_equal(ph, pd), (nmap, pattern, ih, idv)
    assert len(pd) == 0 or np.all(np.diff(pd[:, 1]) > 0)


def test_emulated_filter_methods_2_and_3_and_declines(emu):
    rng = np.random.default_rng(5)
    nmap, nk = 1200, 1100
    vis = (rng.random(nmap) < 0.9).astype(np.uint8)
    d0 = rng.integers(0, 80, nmap).astype(np.uint32)
    d1 = d0 + rng.integers(0, 40, nmap).astype(np.uint32)
    k = np.empty(2 * nmap, np.uint32)
    k[0::2] = (d0 << 16) | rng.integers(0, nk // 3, nmap).astype(np.uint32)
    k[1::2] = (d1 << 16) | rng.integers(0, nk, nmap).astype(np.uint32)
    ph, ih, pd, idv = _both(emu, k, vis, nk, 2, lowe=0.8)
    assert np.array_equal(ih, idv) and np.array_equal(ph, pd) and 0 < ih[0] < nmap
    sad = rng.integers(0, 32 * 60, nmap).astype(np.uint32)
    k3 = (sad << 16) | rng.integers(0, nk // 2, nmap).astype(np.uint32)
    k3[rng.random(nmap) < 0.2] = 0xFFFFFFFF
    ph, ih, pd, idv = _both(emu, k3, vis, nk, 3)
    assert np.array_equal(ih, idv) and np.array_equal(ph, pd) and ih[0] > 0
    # organ pipe: libstdc++ leaves quicksort for heapsort at its depth limit -> the kernel heapsorts those segments too
    n = 2001
    train = np.minimum(np.arange(n), np.arange(n)[::-1]) % 1003
    keys = (np.full(n, 5, np.uint32) << 16) | train.astype(np.uint32)
    ph, ih, pd, idv = _both(emu, keys, np.ones(n, np.uint8), 1003, 1)
    assert idv[2] == 0 and np.array_equal(ih, idv) and np.array_equal(ph, pd) and idv[1] == n
    assert _both.last_info[3] > 0 and _both.last_info[7] > 16, _both.last_info[:9]

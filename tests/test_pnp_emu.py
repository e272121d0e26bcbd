"""The RANSAC PnP kernels (csrc/pnp_kernels.cuh: P3P hypotheses, fp32-with-margin scoring, arg-max + consensus set) EXECUTED ON
THE CPU through tests/cpp/cuda_emu.h, against the fp64 oracle (oracle/pnp_oracle.py) with the assertions of the hardware
test (tests/test_pnp_gpu.py::test_hypothesis_scoring_bit_level) on a smaller problem: one OS thread plays one CUDA thread."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import mvo_synth

ROOT = Path(__file__).resolve().parent.parent
K = mvo_synth.K_DEFAULT


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = tmp_path_factory.mktemp("pnpemu") / "libpnp_emu.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", str(ROOT / "include"),
                    "-I", str(ROOT / "monocular-visual-odometry_b200" / "csrc"), "-I", str(ROOT / "tests" / "cpp"), "-I", "/usr/local/cuda/include",
                    str(ROOT / "tests" / "cpp" / "pnp_emu.cpp"), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    lib.emu_pnp_ransac.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_int, C.c_uint64] + [C.c_void_p] * 5
    return lib


@pytest.mark.parametrize("seed,n,outl", [(1, 600, 0.3), (7, 150, 0.5)])
def test_emulated_pnp_scoring_and_consensus(emu, seed, n, outl):
    from oracle import pnp_oracle as po
    P, uv, rvec_t, tvec_t, _ = mvo_synth.pnp_problem(seed, n=n, outlier_frac=outl)
    P, uv = np.ascontiguousarray(P, np.float32), np.ascontiguousarray(uv, np.float32)
    H = 768
    poses, counts = np.zeros((H, 12)), np.zeros(H, np.int32)
    best_pose, out_i, inl = np.zeros(12), np.zeros(4, np.int32), np.zeros(n, np.int32)
    Kc = np.ascontiguousarray(K, np.float64)
    ni = emu.emu_pnp_ransac(P.ctypes.data, uv.ctypes.data, n, Kc.ctypes.data, 2.0, H, 0x9E3779B97F4A7C15, poses.ctypes.data, counts.ctypes.data,
                            best_pose.ctypes.data, out_i.ctypes.data, inl.ctypes.data)
    inl = inl[:ni]
    valid = counts >= 0
    assert valid.mean() > 0.9                                       # P3P almost always has a real solution
    ref, margin = po.count_inliers(P, uv, K, poses[valid], 2.0)
    clear = margin > 1e-7
    assert clear.mean() > 0.99 and np.array_equal(counts[valid][clear], ref[clear])          # exact counts (fp32 + margin + fp64 recheck)
    R = poses[valid][:, :9].reshape(-1, 3, 3)
    assert np.abs(R @ R.transpose(0, 2, 1) - np.eye(3)).max() < 1e-9 and np.abs(np.linalg.det(R) - 1).max() < 1e-9
    best = int(np.flatnonzero(counts == counts.max())[0])                                   # highest count, lowest index on ties
    assert out_i[1] == best and np.array_equal(best_pose, poses[best])
    e = po.reproj_err2(P, uv, K, poses[best][:9].reshape(3, 3), poses[best][9:])
    assert np.array_equal(inl, np.flatnonzero(e <= 4.0)) and ni >= 0.9 * (1 - outl) * n
    # the best minimal model is close to the generating pose
    assert np.abs(po.rvec_from_R(poses[best][:9].reshape(3, 3)) - rvec_t).max() < 2e-2 and np.abs(poses[best][9:] - tvec_t).max() < 1e-1

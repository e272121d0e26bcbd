"""Self-test of the CPU-tier CUDA emulation (tests/cpp/cuda_emu.h) used by tests/test_two_view_emu.py and
tests/test_orb_variants_emu.py: warp collectives (shuffles, ballot, __syncwarp), shared-memory and global atomics and the
integer intrinsics against numpy, on a grid whose last block is partly idle."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def test_emulated_collectives_and_atomics(tmp_path):
    so = tmp_path / "libemu_selftest.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", str(ROOT / "tests" / "cpp"),
                    "-I", "/usr/local/cuda/include", str(ROOT / "tests" / "cpp" / "cuda_emu_selftest.cpp"), "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    rng = np.random.default_rng(0)
    blocks = 3
    n = blocks * 96 - 10
    v = rng.integers(0, 1000, n).astype(np.int32)
    out = np.zeros(blocks * 96 * 6 + 8, np.int32)
    rc = lib.emu_selftest(v.ctypes.data, n, blocks, out.ctypes.data)
    assert rc == (0x10 + 0x0F + 0x05 + 0x01) + 1000 * 5 + 100000 * 8          # __vsadu4, __ffs, __popc
    vp = np.zeros(blocks * 96, np.int64)
    vp[:n] = v
    w = vp.reshape(-1, 32)
    o = out[: blocks * 96 * 6].reshape(-1, 6)
    assert np.array_equal(o[:, 0], np.repeat(w.sum(1), 32))                  # butterfly sum (__shfl_xor_sync)
    assert np.array_equal(o[:, 1], np.cumsum(w, 1).ravel())                  # inclusive scan (__shfl_up_sync)
    assert np.array_equal(o[:, 2], np.repeat(w[:, 0], 32))                   # broadcast (__shfl_sync)
    assert np.array_equal(o[:, 3], np.concatenate([w[:, 1:], w[:, 31:32]], 1).ravel())      # __shfl_down_sync
    odd = w & 1
    assert np.array_equal(o[:, 4], (np.cumsum(odd, 1) - odd).ravel())        # ballot + lanemask_lt + popc: rank among the set lanes
    assert np.array_equal(o[:, 5], np.repeat(w.sum(1), 32))                  # shared memory written before __syncwarp / __syncthreads
    assert np.array_equal(out[blocks * 96 * 6: blocks * 96 * 6 + 4], np.bincount(vp & 3, minlength=4))      # shared + global atomicAdd

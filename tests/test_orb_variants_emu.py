"""The descriptor blur of the shipped path (csrc/orb_variants.cuh: k_blur2) EXECUTED ON THE CPU: the header nvcc
compiles is built for the host with tests/cpp/orb_variants_emu.cpp (one OS thread per CUDA thread, barriers for
__syncthreads / warp shuffles) and its outputs are compared, byte for byte, with the numpy restatement of cv::ORB
(oracle/orb_oracle.py, itself pinned against cv2 in tests/test_orb_oracle.py).  Catches indexing / data-flow mistakes in
kernels that were written without GPU access; the hardware run is tests/test_orb_variants_gpu.py."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import mvo_synth

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def emu(tmp_path_factory):
    so = tmp_path_factory.mktemp("orbemu") / "liborb_emu.so"
    subprocess.run(["g++", "-O1", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-pthread", "-I", str(ROOT / "include"),
                    "-I", str(ROOT / "monocular-visual-odometry_b200" / "csrc"), "-I", "/usr/local/cuda/include",
                    "-I", str(ROOT / "tests" / "cpp"), str(ROOT / "tests" / "cpp" / "orb_variants_emu.cpp"), "-o", str(so)], check=True)
    return C.CDLL(str(so))


def _pack(levels):
    w = np.array([l.shape[1] for l in levels], np.int32)
    h = np.array([l.shape[0] for l in levels], np.int32)
    return w, h, np.concatenate([np.ascontiguousarray(l).ravel() for l in levels])


def _unpack(flat, w, h):
    out, o = [], 0
    for ww, hh in zip(w, h):
        out.append(flat[o:o + ww * hh].reshape(hh, ww))
        o += ww * hh
    return out


@pytest.mark.parametrize("scene", ["rect_200x150", "noise_131x97", "rect_pyramid_320x240"])
def test_blur2_equals_the_oracle_blur(emu, scene):
    from oracle import orb_oracle as oo
    if scene == "rect_200x150":
        levels = [mvo_synth.rect_scene(3, 200, 150, n_rect=150)]
    elif scene == "noise_131x97":                                  # widths that are not multiples of 4 or of the tile, heights not of 16
        levels = [mvo_synth.noise_scene(4, 131, 97), mvo_synth.noise_scene(5, 64, 70)]
    else:
        levels = oo.build_pyramid(mvo_synth.rect_scene(6, 320, 240, n_rect=300), 4, 1.2)
    w, h, flat = _pack(levels)
    out = np.zeros_like(flat)
    assert emu.emu_blur2(len(levels), w.ctypes.data, h.ctypes.data, flat.ctypes.data, out.ctypes.data) == 0
    for got, lvl in zip(_unpack(out, w, h), levels):
        ref = oo.blur7(lvl)
        assert np.array_equal(got, ref), (scene, lvl.shape, int((got != ref).sum()))

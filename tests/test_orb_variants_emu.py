"""The experimental ORB kernels of csrc/orb_variants.cuh (k_blur2, k_describe_sel2) EXECUTED ON THE CPU: the header nvcc
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


def test_describe_sel2_equals_the_oracle(emu):
    """Keypoint records (position, size, angle, Harris response, octave) and descriptor bytes of oracle-detected keypoints."""
    import mvo_b200
    from oracle import orb_oracle as oo
    img = mvo_synth.rect_scene(8, 320, 240, n_rect=300)
    kp = oo.detect(img)                                            # level-major, like the selection list the kernel reads
    assert len(kp) > 300
    kp = kp[:: max(1, len(kp) // 160)]                             # ~160 keypoints over all levels
    levels = oo.build_pyramid(img, 4, 1.2)
    blurred = [oo.blur7(l) for l in levels]
    scales = np.array(oo.level_scales(4, 1.2), np.float32)
    w, h, flat = _pack(levels)
    _, _, flat_b = _pack(blurred)
    lvl = kp["octave"].astype(np.uint32)
    inv = (np.float32(1.0) / scales)[kp["octave"]]
    x = np.rint((kp["x"] * inv).astype(np.float32)).astype(np.uint32)
    y = np.rint((kp["y"] * inv).astype(np.float32)).astype(np.uint32)
    sel_xy = (x | (y << 12)).astype(np.uint32)
    n = len(kp)
    kout, desc = np.zeros(n, mvo_b200.KEYPOINT_DTYPE), np.zeros((n, 32), np.uint8)
    emu.emu_describe_sel2.argtypes = [C.c_int] + [C.c_void_p] * 7 + [C.c_int, C.c_void_p, C.c_void_p]
    cnt = emu.emu_describe_sel2(4, w.ctypes.data, h.ctypes.data, scales.ctypes.data, flat.ctypes.data, flat_b.ctypes.data, sel_xy.ctypes.data,
                                lvl.ctypes.data, n, kout.ctypes.data, desc.ctypes.data)
    assert cnt == n
    assert kout.tobytes() == np.ascontiguousarray(kp).tobytes()
    assert np.array_equal(desc, oo.compute(img, kp))

"""The run_vo application (monocular-visual-odometry_b200/apps/: the reference's run_vo.cpp:61-151 on libmvo, no display):
its PNG reader against cv2.imread — what run_vo.cpp:114 calls — and its error behaviour on the CPU tier; the whole run on a
synthetic PNG dataset on the GPU tier (child process, first hardware run)."""
import ctypes as C
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import os as _os
_TIMEOUT_SCALE = float(_os.environ.get("MVO_TEST_TIMEOUT_SCALE", "1"))      # > 1 when the library under test is the CPU emulation (MVO_LIB)

import mvo_synth
from conftest import GOLDEN, have_cv2

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "monocular-visual-odometry_b200"
APP = PKG / "build" / "run_vo"


@pytest.mark.skipif(not have_cv2(), reason="cv2 writes the PNG files and is the reference reader")
def test_png_reader_equals_cv_imread(tmp_path):
    import cv2
    so = tmp_path / "libpng_reader.so"
    subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", str(PKG / "apps" / "png_reader.cpp"), "-lz", "-o", str(so)], check=True)
    lib = C.CDLL(str(so))
    rng = np.random.default_rng(0)
    images = {"gray": mvo_synth.rect_scene(1), "bgr": mvo_synth.color_scene(2), "odd_size": rng.integers(0, 256, (37, 53, 3), dtype=np.uint8),
              "bgra": rng.integers(0, 256, (40, 61, 4), dtype=np.uint8), "gray16": rng.integers(0, 65536, (33, 47)).astype(np.uint16),
              "flat": np.full((20, 30, 3), 200, np.uint8), "ramp": np.tile(np.arange(256, dtype=np.uint8), (64, 1))}
    for name, im in images.items():
        for level in (0, 9):                                         # stored and best compression: different filter / block mixes
            path = str(tmp_path / f"{name}_{level}.png")
            assert cv2.imwrite(path, im, [cv2.IMWRITE_PNG_COMPRESSION, level])
            ref = cv2.imread(path)
            out, r, c = np.zeros(ref.size, np.uint8), C.c_int(), C.c_int()
            assert lib.mvo_app_read_png_bgr(path.encode(), out.ctypes.data, out.size, C.byref(r), C.byref(c)) == 0, name
            assert (r.value, c.value) == ref.shape[:2] and np.array_equal(out.reshape(ref.shape), ref), name
    r, c = C.c_int(), C.c_int()
    bad = tmp_path / "bad.png"
    bad.write_bytes(b"not a png at all, but long enough to pass the size test........")
    out = np.zeros(16, np.uint8)
    assert lib.mvo_app_read_png_bgr(str(bad).encode(), out.ctypes.data, 16, C.byref(r), C.byref(c)) == -1
    assert lib.mvo_app_read_png_bgr(str(tmp_path / "missing.png").encode(), out.ctypes.data, 16, C.byref(r), C.byref(c)) == -1
    trunc = tmp_path / "trunc.png"
    trunc.write_bytes((tmp_path / "bgr_9.png").read_bytes()[:2000])
    assert lib.mvo_app_read_png_bgr(str(trunc).encode(), out.ctypes.data, 16, C.byref(r), C.byref(c)) == -1


def test_run_vo_argument_and_config_errors(built, tmp_path):
    assert APP.exists(), "build() makes run_vo"
    r = subprocess.run([str(APP)], capture_output=True, text=True)
    assert r.returncode == 1 and "config file" in r.stderr
    r = subprocess.run([str(APP), str(tmp_path / "nope.yaml")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot read the config file" in r.stderr
    cfg = tmp_path / "no_dataset.yaml"
    cfg.write_text("%YAML:1.0\ndataset_name: \"x\"\n")
    r = subprocess.run([str(APP), str(cfg)], capture_output=True, text=True)
    assert r.returncode == 1 and "x/dataset_dir" in r.stderr


GPU_CHILD = r'''
import re, subprocess, sys
import numpy as np, cv2
sys.path.insert(0, r"{root}"); sys.path.insert(0, r"{root}/monocular-visual-odometry_b200/python")
import mvo_b200, mvo_synth
n = 14
frames, _ = mvo_synth.room_sequence(0, n)
imgs = [mvo_synth.gray_to_bgr(f) for f in frames]
for i, im in enumerate(imgs):
    assert cv2.imwrite(r"{tmp}/rgb_%05d.png" % i, im)
cfg = open(r"{fixture}").read()
cfg = cfg.replace('dataset_name: "fr1_desk"', 'dataset_name: "matlab"').replace("dataset_dir: data/dataset_images_matlab", 'dataset_dir: "{tmp}"')
cfg = re.sub(r"num_images: 150", "num_images: %d" % n, cfg, count=1).replace("save_predicted_traj_to: data/test_data/cam_traj.txt", 'save_predicted_traj_to: "{tmp}/traj.txt"')
open(r"{tmp}/config.yaml", "w").write(cfg)
import os
r = subprocess.run([r"{app}", r"{tmp}/config.yaml"], capture_output=True, text=True, timeout=200 * float(os.environ.get("MVO_TEST_TIMEOUT_SCALE", "1")))
assert r.returncode == 0, (r.stdout[-1500:], r.stderr[-1500:])
assert "Wrote %d poses" % n in r.stdout
# the same frames through the ctypes path with the same configuration
lib = mvo_b200.load_library()
import ctypes as C
h = C.c_void_p()
assert lib.mvo_config_load(r"{tmp}/config.yaml".encode(), C.byref(h)) == 0
p, vp, K = mvo_b200.Params(), mvo_b200.VoParams(), np.zeros(9)
lib.mvo_default_params(C.byref(p)); lib.mvo_vo_default_params(C.byref(vp))
assert lib.mvo_config_apply(h, C.byref(p), None, K.ctypes.data) == 0 and lib.mvo_config_apply_vo(h, C.byref(vp)) == 0
ctx = mvo_b200.Context(0, params=p)
vh = C.c_void_p()
assert lib.mvo_vo_create(ctx.h, K.ctypes.data, 480, 640, C.byref(vp), C.byref(vh)) == 0
poses = []
for im in imgs:
    T = np.zeros(16)
    assert lib.mvo_vo_add_frame(vh, im.ctypes.data, 3, 640 * 3, T.ctypes.data, None) == 0
    poses.append(T.reshape(4, 4))
got, cnt = np.zeros((n, 16)), C.c_int(0)
assert lib.mvo_read_pose_file(r"{tmp}/traj.txt".encode(), got.ctypes.data, n, C.byref(cnt)) == 0 and cnt.value == n
assert np.abs(got.reshape(n, 4, 4) - np.stack(poses)).max() < 2e-5
assert lib.mvo_vo_is_initialized(vh) == 1
print("run_vo child ok")
'''


@pytest.mark.gpu
@pytest.mark.skipif(not have_cv2(), reason="cv2 writes the PNG dataset")
def test_run_vo_on_a_png_dataset(built, tmp_path):
    r = subprocess.run([sys.executable, "-c", GPU_CHILD.format(root=str(ROOT), tmp=str(tmp_path), app=str(APP), fixture=str(GOLDEN / "config_fixture.yaml"))],
                       capture_output=True, text=True, timeout=300 * _TIMEOUT_SCALE)
    assert r.returncode == 0 and "run_vo child ok" in r.stdout, (r.stdout[-1500:], r.stderr[-3000:])

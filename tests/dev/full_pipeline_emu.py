"""The whole product on the CPU: every stage and the `mvo_vo_*` state machine through the library's REAL code, built for the host by
tests/emu_build.py (kernels through tests/cpp/cuda_emu.h, cluster LM kernels included), next to the oracle pipeline on the same
frames.  About 20 minutes for 13 frames (one OS thread plays one CUDA thread).  Usage:
    python -c "import sys; sys.path.insert(0,'tests'); import emu_build; print(emu_build.build('/tmp/emu_full', emu_build.ALL_UNITS))"
    MVO_LIB=/tmp/emu_full/libmvo_emu.so python tests/dev/full_pipeline_emu.py
Output of the round-1 run: profiles/emu_full_pipeline_r1.log (correctness evidence, NOT a performance number)."""
import sys, time
import numpy as np
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python"))
import mvo_b200, mvo_synth
from oracle import vo_pipeline_oracle as vp
K = mvo_synth.K_DEFAULT
n = 13
frames, truth = mvo_synth.room_sequence(0, n)
ctx = mvo_b200.Context(0, max_keypoints=2000, ba_iterations=10)
vo = mvo_b200.VisualOdometry(ctx, K, 480, 640, init_calc_homography=1)
cpu = vp.CpuVo(K, 480, 640, max_number_of_keypoints=2000, ba_iterations=10)
Tg, Tc = [], []
t0 = time.time()
for i, f in enumerate(frames):
    img = mvo_synth.gray_to_bgr(f)
    T, info = vo.add_frame(img)
    T2, info2 = cpu.add_frame(img)
    Tg.append(T); Tc.append(T2)
    print(i, "emu:", info.state_out, info.keyframe, info.best_sol, info.n_keypoints, info.n_matches, info.n_inliers, info.pnp_ok, info.ba_frames, info.map_points, round(info.eh_ratio, 3),
          "| oracle:", info2["state_out"], info2["keyframe"], info2["n_matches"], info2["n_inliers"], info2["map_points"], "| %.0fs" % (time.time() - t0), flush=True)
sg = [vo.frame_data("id", k)[0] for k in range(3)]
g0 = next((i for i in range(n) if np.abs(Tg[i] - np.eye(4)).max() > 1e-9), n - 2)
eg, _ = vp.trajectory_error(Tg[g0:], truth[g0:]); ec, _ = vp.trajectory_error(Tc[g0:], truth[g0:])
print("trajectory RMS error from frame", g0, ": emulated library %.5f  oracle %.5f" % (eg, ec))
print("FULL PIPELINE EMU DONE")

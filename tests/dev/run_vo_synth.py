"""The whole VO state machine (mvo_vo_*) on the ray-cast room sequence next to the oracle pipeline on the same frames:
frames/s of both and the trajectory error of both against the ground truth.  GPU needed.
Usage: python tests/dev/run_vo_synth.py [n_frames] [calc_homography 0|1] [line|loop] > gpurun_out/run_vo_synth.json"""
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python"))
import mvo_b200  # noqa: E402
import mvo_synth  # noqa: E402
from oracle import vo_pipeline_oracle as vp  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    homo = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    K = mvo_synth.K_DEFAULT
    kind = sys.argv[3] if len(sys.argv) > 3 else "line"
    frames, truth = (mvo_synth.room_loop_sequence if kind == "loop" else mvo_synth.room_sequence)(0, n)
    imgs = [mvo_synth.gray_to_bgr(f) for f in frames]
    ctx = mvo_b200.Context(0, max_keypoints=2000, ba_iterations=10)
    out = {"frames": n, "calc_homography": homo, "sequence": kind}
    for name in ("gpu", "oracle"):
        if name == "gpu":
            vo = mvo_b200.VisualOdometry(ctx, K, 480, 640, init_calc_homography=homo)
            step = lambda im: vo.add_frame(im)
            state = lambda info: info.state_out
            kf = lambda info: info.keyframe
        else:
            cpu = vp.CpuVo(K, 480, 640, max_number_of_keypoints=2000, ba_iterations=10, init_calc_homography=bool(homo))
            step = lambda im: cpu.add_frame(im)
            state = lambda info: info["state_out"]
            kf = lambda info: info["keyframe"]
        step(imgs[0])                                                # first keyframe (and lazy allocations) outside the clock
        T, infos, per = [np.eye(4)], [None], []
        t0 = time.perf_counter()
        for im in imgs[1:]:
            t1 = time.perf_counter()
            Ti, info = step(im)
            per.append(time.perf_counter() - t1)
            T.append(Ti); infos.append(info)
        dt = time.perf_counter() - t0
        states = [1] + [state(i) for i in infos[1:]]
        init = states.index(2) if 2 in states else -1
        rec = {"frames_per_s": (n - 1) / dt, "initialised_at": init, "keyframes": 1 + sum(kf(i) for i in infos[1:])}
        kinds = {}
        for p, i, s_in in zip(per, infos[1:], states[:-1]):
            kinds.setdefault(("init" if s_in == 1 else "keyframe" if kf(i) else "tracked"), []).append(p)
        rec["ms_per_frame_by_kind"] = {k: {"n": len(v), "mean": 1e3 * float(np.mean(v)), "median": 1e3 * float(np.median(v))} for k, v in kinds.items()}
        if init >= 0:
            err, scale = vp.trajectory_error(T[init:], truth[init:])
            rec.update(trajectory_rms_error=err, scale=scale, path=float(np.linalg.norm(truth[-1][:3, 3] - truth[init][:3, 3])))
        out[name] = rec
    print(json.dumps(out))


if __name__ == "__main__":
    main()

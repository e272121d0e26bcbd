"""S independent sequences through the WHOLE state machine concurrently on one GPU: one mvo_vo (context, streams, extraction
workers) and one driver thread per sequence, every thread running mvo_vo_run_sequence over its own 150-frame sequence (ctypes
releases the GIL inside the C ABI).  One sequence is two chains of small dependent kernels that leave most of the 148 SMs idle;
this measures how far the whole-GPU frame rate grows with S.  The bench headline stays one sequence per GPU (BASELINE config 5).
Usage: python tools/multi_sequence_pipeline.py [S ...]   -> one JSON line per S (default 1 2 4 8)."""
import json
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python"))
import mvo_b200  # noqa: E402
import mvo_synth  # noqa: E402
import torch  # noqa: E402

N_FRAMES, PASSES, WARM = 150, 4, 2


class Lane:
    def __init__(self, seed):
        self.stream = torch.cuda.Stream()
        self.ctx = mvo_b200.Context(0, max_keypoints=2000, ba_iterations=10)
        self.ctx.set_stream(self.stream.cuda_stream)
        frames, _ = mvo_synth.cached_room_loop_sequence(seed % 2, N_FRAMES)
        self.vo = mvo_b200.VisualOdometry(self.ctx, mvo_synth.K_DEFAULT, 480, 640)
        self.d = torch.from_numpy(np.stack([mvo_synth.gray_to_bgr(f) for f in frames])).cuda()
        self.ptrs = [self.d[i].data_ptr() for i in range(N_FRAMES)]
        self.kf = 0

    def run(self, passes):
        for _ in range(passes):
            self.vo.reset()
            _, infos = self.vo.run_sequence(self.ptrs, channels=3, stride=1920, on_device=True)
            self.kf = sum(i.keyframe for i in infos)
            assert infos[-1].state_out == 2


class UtilSampler(threading.Thread):
    def __init__(self):
        super().__init__(daemon=True)
        self.vals, self._halt = [], threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(0)
        except Exception:
            self.nv = None

    def run(self):
        while self.nv and not self._halt.is_set():
            try:
                self.vals.append(self.nv.nvmlDeviceGetUtilizationRates(self.h).gpu)
            except Exception:
                pass
            self._halt.wait(0.02)

    def stop(self):
        self._halt.set()
        self.join(timeout=1)
        return float(np.median(self.vals)) if self.vals else None


def main():
    counts = [int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]
    lanes = []
    for S in counts:
        while len(lanes) < S:
            lanes.append(Lane(len(lanes)))
        act = lanes[:S]
        ths = [threading.Thread(target=l.run, args=(WARM,)) for l in act]
        [t.start() for t in ths]; [t.join() for t in ths]
        torch.cuda.synchronize()
        samp = UtilSampler(); samp.start()
        t0 = time.perf_counter()
        ths = [threading.Thread(target=l.run, args=(PASSES,)) for l in act]
        [t.start() for t in ths]; [t.join() for t in ths]
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        util = samp.stop()
        print(json.dumps({"sequences_per_gpu": S, "frames_per_s": S * PASSES * N_FRAMES / dt, "ms_per_frame_per_sequence": 1e3 * dt / (PASSES * N_FRAMES),
                          "keyframes_per_pass": act[0].kf, "nvml_gpu_util_pct_median": util,
                          "what": "whole run_vo state machine (mvo_vo_run_sequence, 150 frames, images resident in HBM), wall clock around the joined driver threads"}), flush=True)


if __name__ == "__main__":
    main()

"""Experiment for the next GPU session: S independent sequences tracked CONCURRENTLY on one GPU (one tracker, one stream
and one driver thread per sequence; ctypes releases the GIL inside the C ABI).  A single tracked frame is a ~220 us chain of
small dependent kernels (one cluster for the BA, one CTA for the match-list filter ...) that leaves most of the 148 SMs
idle, so the whole-GPU frame rate should grow with S until the SMs or the host threads saturate.  Same workload per
sequence as bench.py (device-resident frames, 2001 keypoints, 4096 PnP hypotheses, 5-frame BA).
Usage: python tools/multi_sequence_bench.py [S ...]   -> one JSON line per S (default 1 2 4 8)."""
import json
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tools"))
import tracker_workload as bench  # noqa: E402  (puts the package on sys.path)
import mvo_b200  # noqa: E402
import mvo_synth  # noqa: E402
import torch  # noqa: E402

STEPS, WARM = 300, 10


class Lane:
    def __init__(self, seed):
        self.stream = torch.cuda.Stream()
        self.ctx = mvo_b200.Context(0, max_keypoints=bench.MAX_KPTS, ba_iterations=bench.BA_ITERS)
        self.ctx.set_stream(self.stream.cuda_stream)
        imgs, _, self.order = bench.build_sequence(seed)
        kp0, desc0 = self.ctx.orb_extract(imgs[0])
        perm = bench.map_order(len(kp0))
        self.trk = mvo_b200.Tracker(self.ctx, mvo_synth.K_DEFAULT, bench.H, bench.W)
        self.trk.set_map(bench.map_from_first_frame(kp0)[perm], np.ascontiguousarray(desc0[perm]))
        self.trk.reset(np.eye(4))
        n_slots = 40                                              # x S lanes >= L2 for S >= 4; stated in the output
        self.frames = torch.empty((n_slots, bench.H, bench.W, 3), dtype=torch.uint8, device="cuda")
        for s in range(n_slots):
            self.frames[s].copy_(torch.from_numpy(imgs[self.order[s % len(self.order)]]))
        self.n_slots = n_slots
        self.ok = False

    def args(self, i):
        return (self.frames[i % self.n_slots].data_ptr(),), dict(channels=3, stride=bench.W * 3, on_device=True)

    def run(self, n, first):
        a, k = self.args(first)
        self.trk.prefetch(*a, **k)
        res = None
        for i in range(n):
            if i + 1 < n:
                a2, k2 = self.args(first + i + 1)
                self.trk.prefetch(*a2, **k2)
            a, k = self.args(first + i)
            _, res = self.trk.track(*a, **k)
        self.ok = bool(res.pnp_ok) and res.n_inliers > 100


def measure(S):
    lanes = [Lane(seed) for seed in range(S)]
    for ln in lanes:
        ln.run(WARM, 0)
    torch.cuda.synchronize()
    gate = threading.Barrier(S + 1)

    def work(ln):
        gate.wait()
        ln.run(STEPS, WARM)
        ln.ctx.synchronize()

    th = [threading.Thread(target=work, args=(ln,)) for ln in lanes]
    for t in th:
        t.start()
    gate.wait()
    t0 = time.perf_counter()
    for t in th:
        t.join()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out = {"sequences_per_gpu": S, "frames_per_s": S * STEPS / dt, "ms_per_frame_per_sequence": 1e3 * dt / STEPS,
           "all_tracking": all(ln.ok for ln in lanes), "steps": STEPS,
           "frame_slots_MB": S * lanes[0].n_slots * bench.H * bench.W * 3 / 1e6, "timing": "wall clock around S joined driver threads + device synchronize"}
    for ln in lanes:
        ln.trk.close()
        ln.ctx.close()
    return out


if __name__ == "__main__":
    for S in ([int(a) for a in sys.argv[1:]] or [1, 2, 4, 8]):
        print(json.dumps(measure(S)), flush=True)

"""Runs N tracked frames (device-resident images) — target command for ncu captures."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tools"))
import numpy as np, torch, mvo_b200, mvo_synth
import tracker_workload as bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
ctx = mvo_b200.Context(0, max_keypoints=2000, ba_iterations=10)
imgs, T_true, order = bench.build_sequence(0)
kp0, d0 = ctx.orb_extract(imgs[0])
trk = mvo_b200.Tracker(ctx, mvo_synth.K_DEFAULT, 480, 640)
perm = bench.map_order(len(kp0))
trk.set_map(bench.map_from_first_frame(kp0)[perm], np.ascontiguousarray(d0[perm])); trk.reset(np.eye(4))
d = [torch.from_numpy(im).cuda() for im in imgs]
torch.cuda.synchronize()
for i in range(n):
    T, r = trk.track(d[order[i % len(order)]].data_ptr(), channels=3, stride=1920, on_device=True)
print("ok", r.n_inliers, r.ba_frames)

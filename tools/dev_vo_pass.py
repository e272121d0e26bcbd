"""One pass of the device-resident state machine over the bench sequence (images resident in HBM) — target command for
ncu captures and for MVO_VO_DEBUG=1 stage timing.  Usage: python tools/dev_vo_pass.py [n_frames] [passes]"""
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python")); sys.path.insert(0, str(ROOT))
import numpy as np, torch, mvo_b200, mvo_synth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
passes = int(sys.argv[2]) if len(sys.argv) > 2 else 1
frames, truth = mvo_synth.cached_room_loop_sequence(0, 150)
imgs = [mvo_synth.gray_to_bgr(f) for f in frames[:n]]
ctx = mvo_b200.Context(0, max_keypoints=2000, ba_iterations=10)
vo = mvo_b200.VisualOdometry(ctx, mvo_synth.K_DEFAULT, 480, 640)
d = [torch.from_numpy(im).cuda() for im in imgs]
torch.cuda.synchronize()
sig = []
for p in range(passes):
    vo.reset()
    t0 = time.perf_counter()
    for k in range(min(2, n)):
        vo.prefetch(d[k].data_ptr(), channels=3, stride=1920, on_device=True)
    kf = 0
    cls = {"init": [], "tracked": [], "keyframe": []}
    for i in range(n):
        tf = time.perf_counter()
        if i + 2 < n:
            vo.prefetch(d[i + 2].data_ptr(), channels=3, stride=1920, on_device=True)
        T, info = vo.add_frame(d[i].data_ptr(), channels=3, stride=1920, on_device=True)
        kf += info.keyframe
        cls["keyframe" if info.keyframe else ("tracked" if info.state_in == 2 else "init")].append(time.perf_counter() - tf)
    dt = time.perf_counter() - t0
    print("  " + ", ".join(f"{k} {len(v)} x {1e3 * sum(v) / max(len(v), 1):.3f} ms" for k, v in cls.items()), flush=True)
    print(f"pass {p}: {n} frames in {1e3 * dt:.1f} ms = {n / dt:.0f} fps, keyframes {kf}, state {info.state_out}, map {info.map_points}", flush=True)
    if p == 0:
        import hashlib
        sig.append(np.asarray(T, dtype=np.float64).tobytes())
        print("  result checksum (last pose + keyframes + map size of pass 0):", hashlib.sha1(b"".join(sig) + bytes([kf % 256]) + int(info.map_points).to_bytes(4, "little")).hexdigest()[:16],
              "last pose t =", np.round(np.asarray(T)[:3, 3], 9).tolist(), flush=True)

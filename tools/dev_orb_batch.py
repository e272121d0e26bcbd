"""Batched device-resident ORB extraction (64 frames per call, the workload of bench.py's extra.orb_batch) as a stand-alone
loop: the target of ncu captures and of A/B timings.  Usage: python tools/dev_orb_batch.py [calls] -> one JSON line."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python"))
import mvo_b200  # noqa: E402
import mvo_synth  # noqa: E402
import torch  # noqa: E402


def main():
    calls = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    B, H, W = 64, 480, 640
    cap = 2001
    stream = torch.cuda.Stream()
    ctx = mvo_b200.Context(0, max_keypoints=2000)
    ctx.set_stream(stream.cuda_stream)
    frames = mvo_synth.cached_room_loop_sequence(seed=0)[0]
    d_frames = torch.empty((2 * B, H, W, 3), dtype=torch.uint8, device="cuda")
    for s in range(2 * B):
        d_frames[s].copy_(torch.from_numpy(mvo_synth.gray_to_bgr(frames[s % len(frames)])))
    d_k = torch.empty(B * cap * 28, dtype=torch.uint8, device="cuda")
    d_d = torch.empty(B * cap * 32, dtype=torch.uint8, device="cuda")
    d_c = torch.empty(B, dtype=torch.int32, device="cuda")

    def call(it):
        ctx.orb_extract_batch_dev(d_frames[(it % 2) * B].data_ptr(), B, H, W, 3, W * 3, H * W * 3, d_k.data_ptr(), d_d.data_ptr(), d_c.data_ptr(), cap)
    for it in range(3):
        call(it)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for it in range(calls):
        call(it)
    e1.record(stream)
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / calls
    names = mvo_b200.kernel_names()
    mvo_b200.timing_enable(ctx, (1 << len(names)) - 1)
    mvo_b200.timing_read(ctx)
    for it in range(2):
        call(it)
    ms, cnt = mvo_b200.timing_read(ctx)
    print(json.dumps({"us_per_frame": round(us / B, 3), "frames_per_s": round(B / (us * 1e-6), 1), "keypoints_per_frame": d_c.sum().item() / B,
                      "kernel_us_per_frame": {names[k]: round(1e3 * ms[k] / (2 * B), 3) for k in range(len(names)) if cnt[k]},
                      "checksum": int(d_d.to(torch.int64).sum().item())}))


if __name__ == "__main__":
    main()

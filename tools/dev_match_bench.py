"""Dev micro-benchmark of the match kernel (device-resident), not the driver's bench."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python"))
import numpy as np, torch, mvo_b200, mvo_synth

ctx = mvo_b200.Context(0)
s = torch.cuda.Stream()
ctx.set_stream(s.cuda_stream)
sizes = [int(a) for a in sys.argv[1:]] or [2001, 8000]
for n in sizes:
    d1 = torch.from_numpy(mvo_synth.random_descriptors(1, n)).cuda()
    d2 = torch.from_numpy(mvo_synth.random_descriptors(2, n)).cuda()
    xy1 = torch.rand(n, 2, device="cuda") * 600
    xy2 = torch.rand(n, 2, device="cuda") * 600
    keys = torch.empty(n * 2, dtype=torch.int32, device="cuda")
    torch.cuda.synchronize()
    for mode in (0, 1, 2):
        with torch.cuda.stream(s):
            for _ in range(5):
                ctx.match_dev(mode, d1.data_ptr(), xy1.data_ptr(), n, d2.data_ptr(), xy2.data_ptr(), n, 50.0, keys.data_ptr())
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            iters = 200
            e0.record(s)
            for _ in range(iters):
                ctx.match_dev(mode, d1.data_ptr(), xy1.data_ptr(), n, d2.data_ptr(), xy2.data_ptr(), n, 50.0, keys.data_ptr())
            e1.record(s); torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / iters
        print(f"n={n} mode={mode}: {us:.2f} us/launch, {n*n/us/1e6:.3f} Tpairs/s")

"""The N > 1 path of bench.py on CPU (gloo, world_size 2).  The hot path does not shard (replicas only):
rank r tracks its own sequence and the only cross-rank traffic is a barrier plus an all_reduce(MAX) of the elapsed
times; the reference arm runs on rank 0 alone and prints the single JSON line."""
import json
import os
import subprocess
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent


def _worker(rank, world, port, out_dir):
    import torch.distributed as dist
    sys.path.insert(0, str(ROOT))
    import bench
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        ms_local = 100.0 + 50.0 * rank                        # rank 1 is the slow one
        ms = bench.reduce_max_ms(ms_local, dist, "cpu")
        dist.barrier()
        fps = bench.whole_job_fps(world, 200, ms)
        seed = bench.sequence_seed(rank)
        np.save(os.path.join(out_dir, f"r{rank}.npy"), np.array([ms, fps, seed]))
    finally:
        dist.destroy_process_group()


def test_time_reduction_and_sequence_assignment_gloo(tmp_path):
    import torch.multiprocessing as mp
    port = 29600 + os.getpid() % 300
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert r0[0] == r1[0] == 150.0                            # max over ranks, identical on every rank
    import bench
    assert r0[1] == r1[1] == 2 * 200 * bench.N_FRAMES / 0.150  # whole-job frames/s: both ranks' frames (200 passes each) over the slowest time
    assert (r0[2], r1[2]) == (0.0, 1.0)                       # independent sequences: seed = rank


def test_sequences_differ_between_ranks():
    sys.path.insert(0, str(ROOT))
    sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python"))
    import mvo_synth
    f0, _ = mvo_synth.room_loop_sequence(0, 2)
    f1, _ = mvo_synth.room_loop_sequence(1, 2)
    assert f0[1].shape == f1[1].shape == (480, 640) and not np.array_equal(f0[1], f1[1])


def test_reference_arm_under_torchrun_prints_one_line_from_rank0(built):
    port = 29300 + os.getpid() % 300
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(port), str(ROOT / "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "2", "--warmup", "1"]
    env = dict(os.environ, MVO_BENCH_FRAMES="24")            # test hook: a 24-frame sequence keeps the CPU suite short
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=str(ROOT), env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip().startswith("{")]
    assert len(lines) == 1, r.stdout
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["steps"] == 2 and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0
    assert d["metric"].startswith("VO frames/sec") and d["unit"] == "frames/s" and d["higher_is_better"] is True
    assert d["config"]["frames_per_step"] == 24 and "workload" in d["config"]

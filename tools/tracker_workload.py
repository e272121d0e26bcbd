"""The tracked-frame workload of round 1's bench (mvo_tracker against a fixed map on a planar ping-pong sequence), kept for the
tracker-level tools (tools/multi_sequence_bench.py, tools/dev_track_loop.py, ncu capture targets).  bench.py itself now runs
the whole state machine (BASELINE config 5)."""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python"))
sys.path.insert(0, str(ROOT))

W, H = 640, 480
N_DISTINCT = 16
MAX_KPTS = 2000
BA_ITERS = 10


def build_sequence(seed):
    """16 distinct frames of a textured plane + ground-truth poses; visiting order is a ping-pong."""
    import mvo_synth
    frames, T_c_w, _ = mvo_synth.planar_sequence(seed, n_frames=N_DISTINCT, plane_z=4.0)
    imgs = [mvo_synth.gray_to_bgr(f) for f in frames]
    order = list(range(1, N_DISTINCT)) + list(range(N_DISTINCT - 2, 0, -1))      # 1..15,14..1 then repeat
    return imgs, [np.linalg.inv(T) for T in T_c_w], order


def map_from_first_frame(kp, plane_z=4.0):
    import mvo_synth
    Ki = np.linalg.inv(mvo_synth.K_DEFAULT)
    rays = (Ki @ np.stack([kp["x"], kp["y"], np.ones(len(kp))]).astype(np.float64)).T
    return (rays * (plane_z / rays[:, 2:3])).astype(np.float32)


def map_order(n, seed=20240923):
    """A seeded permutation standing for the iteration order of the reference's std::unordered_map<int, MapPoint::Ptr>: the
    order keypoints come out of ORB (level-major) makes libstdc++'s std::sort in removeDuplicatedMatches heapsort on every frame."""
    return np.random.default_rng(seed).permutation(n)

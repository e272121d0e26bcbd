"""Timing of the two-view stage (mvo_esti_motion_by_essential, mvo_do_triangulation) next to the OpenCV calls of the
reference (cv2.findEssentialMat + cv2.recoverPose, cv2.triangulatePoints) on the same synthetic inputs.
Dev tool (not the driver's bench): prints one JSON line."""
import json
import sys
import time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT / "monocular-visual-odometry_b200" / "python")); sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import numpy as np
import cv2
import mvo_b200, mvo_synth
from oracle import epipolar_oracle
import test_epipolar_gpu as T

ctx = mvo_b200.Context(0)
K = mvo_synth.K_DEFAULT
p1, p2, R_true, t_true, good, X1 = T._scene(0, n=1500)
for _ in range(5):
    E, R, t, inl = ctx.esti_motion_by_essential(p1, p2, K, 1.0)
names = mvo_b200.kernel_names()
mvo_b200.timing_enable(ctx, 1 << names.index("k_epi"))
mvo_b200.timing_read(ctx)
reps = 50
t0 = time.perf_counter()
for _ in range(reps):
    E, R, t, inl = ctx.esti_motion_by_essential(p1, p2, K, 1.0)
wall_gpu = (time.perf_counter() - t0) / reps
ms, cnt = mvo_b200.timing_read(ctx)
k = names.index("k_epi")
cv2.setNumThreads(0)
t0 = time.perf_counter()
for _ in range(10):
    Eo, Ro, to, inlo = epipolar_oracle.esti_motion_by_essential(p1, p2, K, 0.999, 1.0)
wall_cpu = (time.perf_counter() - t0) / 10
Ki = np.linalg.inv(K)
np1 = ((np.c_[p1, np.ones(len(p1))] @ Ki.T)[:, :2]).astype(np.float32)
np2 = ((np.c_[p2, np.ones(len(p2))] @ Ki.T)[:, :2]).astype(np.float32)
mvo_b200.timing_read(ctx)
t0 = time.perf_counter()
for _ in range(reps):
    X = ctx.do_triangulation(np1, np2, R, t * 0.26, inl)
wall_tri = (time.perf_counter() - t0) / reps
ms2, cnt2 = mvo_b200.timing_read(ctx)
t0 = time.perf_counter()
for _ in range(reps):
    Xo = epipolar_oracle.do_triangulation(np1, np2, R, t * 0.26, inl)
wall_tri_cpu = (time.perf_counter() - t0) / reps
print(json.dumps({
    "workload": "1500 correspondences, 20 % outliers, 0.5 px noise; 4096 hypotheses",
    "esti_motion_by_essential": {"gpu_call_ms": 1e3 * wall_gpu, "gpu_kernels_ms": ms[k] / reps, "kernel_launches_per_call": int(cnt[k]) / reps,
                                 "cv2_call_ms": 1e3 * wall_cpu, "inliers_gpu": int(len(inl)), "inliers_cv2": int(len(inlo)),
                                 "rot_err_gpu": float(T._rot_err(R, R_true)), "rot_err_cv2": float(T._rot_err(Ro, R_true)),
                                 "dir_err_gpu": float(T._dir_err(t, t_true)), "dir_err_cv2": float(T._dir_err(to, t_true))},
    "do_triangulation": {"points": int(len(inl)), "gpu_call_ms": 1e3 * wall_tri, "gpu_kernel_ms": ms2[k] / reps, "cv2_call_ms": 1e3 * wall_tri_cpu},
}))

"""Seeded synthetic inputs for tests, smoke() and bench.py.  Pure numpy.

No dataset ships with the reference (its data/ directory is git-ignored), so every workload is
generated: rectangle-texture 640x480 frames, planar warps of them with known camera poses,
3D-2D correspondence sets with outliers, and small BA problems.
"""
from __future__ import annotations

import numpy as np

W, H = 640, 480
K_DEFAULT = np.array([[615.0, 0, 320.0], [0, 615.0, 240.0], [0, 0, 1.0]])   # config/config.yaml:18-21 style


def _blur3(img):
    """3x3 Gaussian, sigma 0.8, reflect-101 borders, float math (input data only)."""
    k = np.exp(-np.arange(-1, 2) ** 2 / (2 * 0.8 ** 2))
    k /= k.sum()
    p = np.pad(img.astype(np.float32), 1, mode="reflect")
    h = k[0] * p[:, :-2] + k[1] * p[:, 1:-1] + k[2] * p[:, 2:]
    v = k[0] * h[:-2] + k[1] * h[1:-1] + k[2] * h[2:]
    return np.clip(np.rint(v), 0, 255).astype(np.uint8)


def rect_scene(seed=0, width=W, height=H, n_rect=1500, blur=True):
    """Mid-gray canvas + filled axis-aligned rectangles."""
    rng = np.random.default_rng(seed)
    img = np.full((height, width), 128, np.uint8)
    xs = rng.integers(0, width, n_rect)
    ys = rng.integers(0, height, n_rect)
    ws = rng.integers(4, 40, n_rect)
    hs = rng.integers(4, 40, n_rect)
    gs = rng.integers(0, 256, n_rect)
    for x, y, w, h, g in zip(xs, ys, ws, hs, gs):
        img[y:y + h, x:x + w] = g
    return _blur3(img) if blur else img


def gray_to_bgr(gray):
    return np.repeat(gray[:, :, None], 3, axis=2).copy()


def noise_scene(seed=0, width=W, height=H):
    """Dense uniform noise: drives every pyramid level above OpenCV's retainBest caps."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, (height, width), dtype=np.uint8)


def color_scene(seed=0, width=W, height=H):
    """Three independent rectangle planes (exercises the BGR->gray fixed-point formula)."""
    return np.stack([rect_scene(seed * 3 + c, width, height) for c in range(3)], axis=2).copy()


def warp_homography(img, Hm):
    """Inverse-map bilinear warp, out(x) = img(Hm^-1 x); border = 128."""
    h, w = img.shape[:2]
    Hi = np.linalg.inv(Hm)
    ys, xs = np.mgrid[0:h, 0:w]
    p = np.stack([xs.ravel(), ys.ravel(), np.ones(h * w)])
    q = Hi @ p
    qx, qy = q[0] / q[2], q[1] / q[2]
    x0 = np.floor(qx).astype(np.int64)
    y0 = np.floor(qy).astype(np.int64)
    fx, fy = qx - x0, qy - y0
    ok = (x0 >= 0) & (y0 >= 0) & (x0 < w - 1) & (y0 < h - 1)
    x0c, y0c = np.clip(x0, 0, w - 2), np.clip(y0, 0, h - 2)
    src = img.astype(np.float64)
    v = (src[y0c, x0c] * (1 - fx) * (1 - fy) + src[y0c, x0c + 1] * fx * (1 - fy)
         + src[y0c + 1, x0c] * (1 - fx) * fy + src[y0c + 1, x0c + 1] * fx * fy)
    v = np.where(ok, v, 128.0)
    return np.clip(np.rint(v), 0, 255).astype(np.uint8).reshape(h, w)


def rodrigues(rvec):
    rvec = np.asarray(rvec, np.float64).reshape(3)
    th = np.linalg.norm(rvec)
    if th < 1e-12:
        return np.eye(3)
    k = rvec / th
    Kx = np.array([[0, -k[2], k[1]], [k[2], 0, -k[0]], [-k[1], k[0], 0]])
    return np.eye(3) + np.sin(th) * Kx + (1 - np.cos(th)) * Kx @ Kx


def planar_sequence(seed=0, n_frames=8, plane_z=4.0, K=K_DEFAULT, step=0.02):
    """Frames of a fronto-parallel textured plane seen from a smoothly moving camera.

    Returns (frames[n] uint8 HxW, T_c_w[n] 4x4 world->camera, texture).  Frame 0 is the texture
    itself (camera at the origin looking down +z at the plane z = plane_z)."""
    tex = rect_scene(seed)
    rng = np.random.default_rng(1000 + seed)
    frames, poses = [], []
    n = np.array([0.0, 0.0, 1.0])
    Ki = np.linalg.inv(K)
    drift_r = rng.normal(0, 1, 3) * 0.004
    drift_t = np.array([step, step * 0.4, step * 0.25])
    for i in range(n_frames):
        rvec = drift_r * i
        R = rodrigues(rvec)
        t = drift_t * i                      # world->camera translation
        T = np.eye(4)
        T[:3, :3], T[:3, 3] = R, t
        # plane-induced homography from frame 0 pixels to frame i pixels: K (R + t n^T / d) K^-1
        Hm = K @ (R + np.outer(t, n) / plane_z) @ Ki
        frames.append(tex if i == 0 else warp_homography(tex, Hm))
        poses.append(T)
    return frames, poses, tex


def pnp_problem(seed=0, n=2000, outlier_frac=0.3, noise=0.5, K=K_DEFAULT):
    """SURVEY.md §8d config 3."""
    rng = np.random.default_rng(seed)
    P = np.stack([rng.uniform(-2, 2, n), rng.uniform(-1.5, 1.5, n), rng.uniform(2, 8, n)], 1)
    rvec = np.array([0.05, -0.1, 0.02])
    tvec = np.array([0.1, -0.05, 0.2])
    R = rodrigues(rvec)
    pc = P @ R.T + tvec
    uv = pc[:, :2] / pc[:, 2:3] * np.array([K[0, 0], K[1, 1]]) + np.array([K[0, 2], K[1, 2]])
    uv += rng.normal(0, noise, uv.shape)
    n_out = int(round(outlier_frac * n))
    out_idx = rng.choice(n, n_out, replace=False)
    uv[out_idx] = np.stack([rng.uniform(0, W, n_out), rng.uniform(0, H, n_out)], 1)
    is_outlier = np.zeros(n, bool)
    is_outlier[out_idx] = True
    return P.astype(np.float32), uv.astype(np.float32), rvec, tvec, is_outlier


def ba_problem(seed=0, n_frames=5, n_points=2000, noise=0.5, outlier_frac=0.05, pose_pert=1e-2,
               point_pert=1e-2, K=K_DEFAULT, visibility=1.0):
    """poses on a smooth arc, every point seen in every frame.

    Returns dict with T_w_c (perturbed initial camera->world poses), T_w_c_true, points (perturbed,
    float32), points_true, edge_frame, edge_point, obs (float32), K."""
    rng = np.random.default_rng(seed)
    P = np.stack([rng.uniform(-2, 2, n_points), rng.uniform(-1.5, 1.5, n_points), rng.uniform(2, 8, n_points)], 1)
    T_true, T_init = [], []
    for f in range(n_frames):
        rvec = np.array([0.01, -0.02, 0.005]) * f
        t = np.array([0.05 * f, 0.01 * f, 0.004 * f * f])
        Tcw = np.eye(4)
        Tcw[:3, :3], Tcw[:3, 3] = rodrigues(rvec), t
        T_true.append(np.linalg.inv(Tcw))
        Tp = np.eye(4)
        Tp[:3, :3] = rodrigues(rng.normal(0, pose_pert, 3))
        Tp[:3, 3] = rng.normal(0, pose_pert, 3)
        T_init.append(np.linalg.inv(Tp @ Tcw))
    ef, ep, obs = [], [], []
    for f in range(n_frames):
        Tcw = np.linalg.inv(T_true[f])
        pc = P @ Tcw[:3, :3].T + Tcw[:3, 3]
        uv = pc[:, :2] / pc[:, 2:3] * K[0, 0] + np.array([K[0, 2], K[1, 2]])
        uv += rng.normal(0, noise, uv.shape)
        bad = rng.random(n_points) < outlier_frac
        uv[bad] += rng.uniform(-30, 30, (bad.sum(), 2))
        vis = rng.random(n_points) < visibility
        for j in np.nonzero(vis)[0]:
            ef.append(f)
            ep.append(j)
            obs.append(uv[j])
    P_init = (P + rng.normal(0, point_pert, P.shape)).astype(np.float32)
    return dict(T_w_c=np.array(T_init), T_w_c_true=np.array(T_true), points=P_init, points_true=P,
                edge_frame=np.array(ef, np.int32), edge_point=np.array(ep, np.int32),
                obs=np.array(obs, np.float32), K=K.copy())


def random_descriptors(seed, n, dup_every=0):
    rng = np.random.default_rng(seed)
    d = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    if dup_every:
        for i in range(dup_every, n, dup_every):
            d[i] = d[i - dup_every]      # exact duplicates -> distance ties
    return d


# ---- a 3-D scene: textured box room seen from a moving camera (two-view initialisation needs depth variation) ----
_ROOM = dict(half_w=1.7, top=-1.1, bottom=1.0, back=6.0, z0=-1.5, ppm=120.0)


def _room_planes(seed):
    """(normal, offset, (axis_u, u0, axis_v, v0), texture) of the five faces; a point X is on a face when n.X = h."""
    r = _ROOM
    depth = r["back"] - r["z0"]
    size = lambda a, b: (int(np.ceil(a * r["ppm"])) + 2, int(np.ceil(b * r["ppm"])) + 2)       # (width, height) in texels
    faces = [
        ((0, 0, 1.0), r["back"], (0, -r["half_w"], 1, r["top"]), size(2 * r["half_w"], r["bottom"] - r["top"])),     # back wall: (x, y)
        ((0, 1.0, 0), r["bottom"], (0, -r["half_w"], 2, r["z0"]), size(2 * r["half_w"], depth)),                     # floor: (x, z)
        ((0, 1.0, 0), r["top"], (0, -r["half_w"], 2, r["z0"]), size(2 * r["half_w"], depth)),                        # ceiling
        ((1.0, 0, 0), -r["half_w"], (2, r["z0"], 1, r["top"]), size(depth, r["bottom"] - r["top"])),                 # left wall: (z, y)
        ((1.0, 0, 0), r["half_w"], (2, r["z0"], 1, r["top"]), size(depth, r["bottom"] - r["top"])),                  # right wall
    ]
    out = []
    for i, (n, h, uv, (tw, th)) in enumerate(faces):
        tex = rect_scene(seed * 16 + i + 1, tw, th, n_rect=int(tw * th / 200))
        out.append((np.array(n, np.float64), float(h), uv, tex))
    return out


def render_room(T_w_c, planes, K=K_DEFAULT, width=W, height=H):
    """Ray-cast the box room: nearest face along every pixel ray, bilinear texture lookup.  T_w_c: camera->world."""
    T_w_c = np.asarray(T_w_c, np.float64)
    ys, xs = np.mgrid[0:height, 0:width]
    d = np.linalg.inv(K) @ np.stack([xs.ravel(), ys.ravel(), np.ones(width * height)]).astype(np.float64)
    d = T_w_c[:3, :3] @ d
    o = T_w_c[:3, 3]
    best_s = np.full(width * height, np.inf)
    val = np.full(width * height, 128.0)
    ppm = _ROOM["ppm"]
    for n, h, (au, u0, av, v0), tex in planes:
        nd = n @ d
        with np.errstate(divide="ignore", invalid="ignore"):
            s = (h - n @ o) / nd
            X = o[:, None] + s * d
        tu, tv = (X[au] - u0) * ppm, (X[av] - v0) * ppm
        th, tw = tex.shape
        ok = np.isfinite(s) & (s > 1e-6) & (s < best_s) & (tu >= 0) & (tv >= 0) & (tu < tw - 1) & (tv < th - 1)
        tu, tv = np.where(ok, tu, 0), np.where(ok, tv, 0)
        x0, y0 = np.floor(tu).astype(np.int64), np.floor(tv).astype(np.int64)
        fx, fy = tu - x0, tv - y0
        t = tex.astype(np.float64)
        v = t[y0, x0] * (1 - fx) * (1 - fy) + t[y0, x0 + 1] * fx * (1 - fy) + t[y0 + 1, x0] * (1 - fx) * fy + t[y0 + 1, x0 + 1] * fx * fy
        val = np.where(ok, v, val)
        best_s = np.where(ok, s, best_s)
    return np.clip(np.rint(val), 0, 255).astype(np.uint8).reshape(height, width)


def room_sequence(seed=0, n_frames=30, step=0.05, K=K_DEFAULT):
    """Frames of the box room from a camera that moves mostly sideways with a slow rotation.

    Returns (frames[n] uint8 HxW, T_w_c[n] 4x4 camera->world ground truth).  Frame 0 is taken from the origin."""
    planes = _room_planes(seed)
    rng = np.random.default_rng(2000 + seed)
    drift_r = rng.normal(0, 1, 3) * 0.003
    drift_t = np.array([step, -0.15 * step, 0.3 * step])
    frames, poses = [], []
    for i in range(n_frames):
        T = np.eye(4)
        T[:3, :3] = rodrigues(drift_r * i)
        T[:3, 3] = drift_t * i
        frames.append(render_room(T, planes, K))
        poses.append(T)
    return frames, poses


def room_loop_poses(seed=0, n_frames=150, period=120, radius=(0.9, 0.12, 0.7)):
    """Camera poses (camera->world) on a closed loop inside the box room: an ellipse in x/z with a small vertical wobble and a slow
    yaw/pitch oscillation, ~0.05 m between frames at the default period — long sequences that never leave the room (BASELINE
    config 5: >= 150 frames per sequence)."""
    rng = np.random.default_rng(3000 + seed)
    ph = rng.uniform(0, 2 * np.pi, 3)
    poses = []
    for i in range(n_frames):
        a = 2 * np.pi * i / period
        T = np.eye(4)
        T[:3, 3] = [radius[0] * np.sin(a), radius[1] * np.sin(2 * a + ph[0]), radius[2] * (1 - np.cos(a))]
        T[:3, :3] = rodrigues(np.array([0.02 * np.sin(a + ph[1]), 0.04 * np.sin(a), 0.01 * np.sin(2 * a)]))
        poses.append(T)
    return poses


def room_loop_sequence(seed=0, n_frames=150, K=K_DEFAULT, **kw):
    """Frames of the box room along room_loop_poses.  Returns (frames[n] uint8 HxW, T_w_c[n])."""
    planes = _room_planes(seed)
    poses = room_loop_poses(seed, n_frames, **kw)
    return [render_room(T, planes, K) for T in poses], poses


def trajectory_error(T_est, T_true):
    """RMS position error (absolute trajectory error) after the best similarity alignment (Umeyama) of the estimated camera
    centres to the true ones — a monocular trajectory has a free scale.  Returns (rms, scale)."""
    a = np.stack([np.asarray(T)[:3, 3] for T in T_est])
    b = np.stack([np.asarray(T)[:3, 3] for T in T_true])
    ma, mb = a.mean(0), b.mean(0)
    A, B = a - ma, b - mb
    U, S, Vt = np.linalg.svd(B.T @ A / len(a))
    D = np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))])
    R = U @ D @ Vt
    s = np.trace(np.diag(S) @ D) / max((A * A).sum(), 1e-300) * len(a)
    aligned = (s * (R @ A.T)).T + mb
    return float(np.sqrt(((aligned - b) ** 2).sum(1).mean())), float(s)


def cached_room_loop_sequence(seed=0, n_frames=150, cache_dir="/tmp"):
    """room_loop_sequence with an on-disk cache (rendering 150 frames takes ~30 s of one core)."""
    import os
    path = os.path.join(cache_dir, f"mvo_room_loop_s{seed}_n{n_frames}.npz")
    if os.path.exists(path):
        try:
            d = np.load(path)
            return list(d["frames"]), list(d["truth"])
        except Exception:
            pass
    frames, truth = room_loop_sequence(seed, n_frames)
    try:
        tmp = path + f".{os.getpid()}.tmp.npz"
        np.savez(tmp, frames=np.stack(frames), truth=np.stack(truth))
        os.replace(tmp, path)
    except Exception:
        pass
    return frames, truth

"""TEST ORACLE for the hypothesis generator of the batched RANSAC PnP (csrc/pnp_kernels.cuh k_pnp_hypotheses; it stands where
cv::solvePnPRansac draws its minimal samples, reference src/vo/vo.cpp:318-320).  Not product code.

Two parts:
  * the SAMPLE LIST, shared with the device: the counter-based generator of the kernel restated (splitmix64 of
    seed ^ splitmix64(h << 20 ^ counter), modulo n, redrawn on a duplicate) — hypothesis h uses correspondences
    sample_indices(seed, h, n)[0:3] for the P3P and [3] to pick among its solutions;
  * an INDEPENDENT P3P solver: the two law-of-cosines conics in the depth ratios (u, v) are reduced to one quartic in v by
    the resultant in u (np.roots), u follows linearly, the depths from the third side, the pose by a Kabsch/SVD alignment of
    the three camera points onto the world points.  The device solves the same conics through their degenerate pencil
    member (a cubic) and builds the pose from two orthonormal triads: no shared code path.
The rule checked: the device's pose for hypothesis h is the P3P solution that reprojects the fourth sample point best."""
from __future__ import annotations

import numpy as np

M64 = (1 << 64) - 1


def splitmix64(x):
    x = (x + 0x9E3779B97F4A7C15) & M64
    x = ((x ^ (x >> 30)) * 0xBF58476D1CE4E5B9) & M64
    x = ((x ^ (x >> 27)) * 0x94D049BB133111EB) & M64
    return x ^ (x >> 31)


def sample_indices(seed, h, n):
    """The four correspondence indices of hypothesis h (k_pnp_hypotheses: at most 64 attempts per slot)."""
    idx, ctr = [], 0
    for k in range(4):
        cand = 0
        for _ in range(64):
            r = splitmix64(seed ^ splitmix64(((h << 20) & M64) ^ ctr))
            ctr += 1
            cand = r % n
            if cand not in idx[:k]:
                break
        idx.append(int(cand))
    return idx


def p3p_solutions(X, f):
    """All real poses (R, t) with depth_i f_i = R X_i + t, depth_i > 0, for three world points X (3 x 3) and unit bearings f (3 x 3)."""
    a = ((X[1] - X[2]) ** 2).sum(); b = ((X[0] - X[2]) ** 2).sum(); c = ((X[0] - X[1]) ** 2).sum()
    ca, cb, cg = f[1] @ f[2], f[0] @ f[2], f[0] @ f[1]
    P = np.polynomial.polynomial
    one = np.array([1.0])
    d = np.array([1.0, -2 * cb, 1.0])                 # 1 - 2 v cb + v^2  (ascending powers of v)
    # quadratics in u with polynomial coefficients in v:  b u^2 + B1 u + C1 = 0  and  b u^2 + B2 u + C2 = 0
    B1, C1 = np.array([0.0, -2 * b * ca]), P.polysub(np.array([0.0, 0.0, b]), a * d)
    B2, C2 = np.array([-2 * b * cg]), P.polysub(np.array([b]), c * d)
    dC, dB = P.polysub(C2, C1), P.polysub(B2, B1)
    # the common root u = -dC / dB substituted into the first quadratic: b dC^2 - B1 dC dB + C1 dB^2 = 0, a quartic in v
    res = P.polyadd(P.polysub(b * P.polymul(dC, dC), P.polymul(B1, P.polymul(dC, dB))), P.polymul(C1, P.polymul(dB, dB)))
    res = np.trim_zeros(res, "b")
    if len(res) < 2:
        return []
    sols = []
    for v in np.roots(res[::-1]):
        if abs(v.imag) > 1e-7 * max(1.0, abs(v.real)) or v.real <= 0:
            continue
        v = v.real
        den = P.polyval(v, dB)
        if abs(den) < 1e-12:
            continue
        u = -P.polyval(v, dC) / den
        dn = 1 + v * v - 2 * v * cb
        if u <= 0 or dn <= 1e-18:
            continue
        s1 = np.sqrt(b / dn)
        Pc = np.stack([s1 * f[0], u * s1 * f[1], v * s1 * f[2]])
        # Kabsch: R, t with Pc_i = R X_i + t
        mx, mp = X.mean(0), Pc.mean(0)
        U, _, Vt = np.linalg.svd((Pc - mp).T @ (X - mx))
        D = np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))])
        R = U @ D @ Vt
        t = mp - R @ mx
        if np.abs(Pc - (X @ R.T + t)).max() < 1e-6 * max(1.0, np.abs(Pc).max()):      # a consistent triangle (not a spurious root)
            sols.append((R, t))
    return sols


def hypothesis(P3, uv, K, seed, h):
    """(pose 12 = R row-major + t, or None; all solutions; the sample) the oracle expects of hypothesis h."""
    n = len(P3)
    idx = sample_indices(seed, h, n)
    X = P3[idx].astype(np.float64)
    fx, fy, cx, cy = K[0, 0], K[1, 1], K[0, 2], K[1, 2]
    u, v = uv[idx, 0].astype(np.float64), uv[idx, 1].astype(np.float64)
    f = np.stack([(u[:3] - cx) / fx, (v[:3] - cy) / fy, np.ones(3)], 1)
    f /= np.linalg.norm(f, axis=1, keepdims=True)
    sols = p3p_solutions(X[:3], f)
    best, be = None, np.inf
    errs = []
    for R, t in sols:
        pc = R @ X[3] + t
        e = np.inf if pc[2] <= 1e-9 else (fx * pc[0] / pc[2] + cx - u[3]) ** 2 + (fy * pc[1] / pc[2] + cy - v[3]) ** 2
        errs.append(e)
        if e < be:
            be, best = e, (R, t)
    return (None if best is None else np.concatenate([best[0].ravel(), best[1]])), sols, idx, errs

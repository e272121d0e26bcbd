"""TEST ORACLE — a second, INDEPENDENT implementation (numpy, no code shared with oracle/ba_oracle.c) of the optimiser the
reference runs through g2o in optimization::bundleAdjustment (src/optimization/g2o_ba.cpp:172-317), written from the rules
of g2o's OptimizationAlgorithmLevenberg as collected in SURVEY.md Appendix B.  Not product code.

g2o itself is an absent third-party dependency ("last version in year 2017", README.md:184): PARITY WITH THE REAL g2o IS
UNPINNED.  What this file adds is a cross-check of the restatement's CONTROL FLOW: ba_oracle.c (C, quaternion poses, Schur
complement on the points, LDLT) and this file (4 x 4 matrices, one dense Jacobian, the full normal equations solved at once)
share nothing but the published rules, and must produce the same per-trial trace — lambda, robust chi2, gain ratio,
accept / reject — on every problem (tests/test_ba_oracle.py).

Rules (App. B): e = z - (f (x/z, y/z) + c); chi2 = e^T Omega e; Huber(delta): rho = chi2 if chi2 <= delta^2 else
2 sqrt(chi2) delta - delta^2, weight rho' = 1 or delta / sqrt(chi2); H = J^T (rho' Omega) J, b = -J^T (rho' Omega) e;
lambda_0 = 1e-5 max diag H; per iteration <= 10 trials: solve (H + lambda I) dx = b, apply (pose: T <- exp(dx) T with
dx = (omega, upsilon); point: additive), gain = (chi_old - chi_new) / (dx . (lambda dx + b) + 1e-3); accepted when gain > 0
and chi_new finite: lambda *= max(1/3, min(1 - (2 gain - 1)^3, 2/3)), nu = 2; else the step is discarded, lambda *= nu,
nu *= 2.  The optimiser stops after the requested iterations, or when 10 trials in a row fail."""
from __future__ import annotations

import numpy as np


def _hat(w):
    return np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0.0]])


def se3_exp(d):
    """g2o SE3Quat::exp for d = (omega, upsilon) as a 4 x 4 matrix."""
    w, u = np.asarray(d[:3], float), np.asarray(d[3:], float)
    th = np.linalg.norm(w)
    W = _hat(w)
    if th < 1e-5:
        R = np.eye(3) + W + 0.5 * W @ W
        V = R
    else:
        R = np.eye(3) + np.sin(th) / th * W + (1 - np.cos(th)) / th ** 2 * W @ W
        V = np.eye(3) + (1 - np.cos(th)) / th ** 2 * W + (th - np.sin(th)) / th ** 3 * W @ W
    T = np.eye(4)
    T[:3, :3] = R
    T[:3, 3] = V @ u
    return T


def _robust(chi, delta):
    if delta <= 0:
        return chi, np.ones_like(chi)
    big = chi > delta * delta
    rho = np.where(big, 2 * np.sqrt(np.maximum(chi, 1e-300)) * delta - delta * delta, chi)
    w = np.where(big, delta / np.sqrt(np.maximum(chi, 1e-300)), 1.0)
    return rho, w


def bundle_adjustment_trace(T_w_c, points, edge_frame, edge_point, obs, K, information=None, fix_points=False, iterations=10,
                            huber_delta=1.0, fix_first_pose=False):
    """Returns (poses T_w_c [F, 4, 4], points [P, 3], trace) with trace = list of (iteration, lambda, chi2 of the trial, gain, accepted)."""
    F, P, E = len(T_w_c), len(points), len(edge_frame)
    Tcw = [np.linalg.inv(np.asarray(T, float).reshape(4, 4)) for T in T_w_c]
    X = np.asarray(points, np.float64).reshape(P, 3).copy()
    ef, ep = np.asarray(edge_frame), np.asarray(edge_point)
    z = np.asarray(obs, np.float64).reshape(E, 2)
    f, cx, cy = float(K[0][0]), float(K[0][2]), float(K[1][2])
    Om = np.eye(2) if information is None else np.asarray(information, float).reshape(2, 2)
    pose_col = {}
    for i in range(F):
        if not (fix_first_pose and i == 0):
            pose_col[i] = 6 * len(pose_col)
    npose = 6 * len(pose_col)
    used = np.unique(ep) if not fix_points else np.zeros(0, int)
    pt_col = {int(l): npose + 3 * k for k, l in enumerate(used)}
    n = npose + 3 * len(used)

    def residuals(Ts, Xs):
        R = np.stack([T[:3, :3] for T in Ts])[ef]
        t = np.stack([T[:3, 3] for T in Ts])[ef]
        pc = np.einsum("eij,ej->ei", R, Xs[ep]) + t
        e = z - np.stack([f * pc[:, 0] / pc[:, 2] + cx, f * pc[:, 1] / pc[:, 2] + cy], 1)
        return e, pc

    def robust_chi2(Ts, Xs):
        e, _ = residuals(Ts, Xs)
        chi = np.einsum("ei,ij,ej->e", e, Om, e)
        return float(_robust(chi, huber_delta)[0].sum())

    trace = []
    lam, nu = 0.0, 2.0
    current = robust_chi2(Tcw, X)
    for it in range(iterations):
        current = robust_chi2(Tcw, X)
        e, pc = residuals(Tcw, X)
        chi = np.einsum("ei,ij,ej->e", e, Om, e)
        _, w = _robust(chi, huber_delta)
        J = np.zeros((2 * E, n))
        x, y, zz = pc[:, 0], pc[:, 1], pc[:, 2]
        Jp = np.zeros((E, 2, 6))          # d e / d (omega, upsilon) under T <- exp(d) T
        Jp[:, 0] = np.stack([f * x * y / zz ** 2, -f * (1 + x ** 2 / zz ** 2), f * y / zz, -f / zz, 0 * zz, f * x / zz ** 2], 1)
        Jp[:, 1] = np.stack([f * (1 + y ** 2 / zz ** 2), -f * x * y / zz ** 2, -f * x / zz, 0 * zz, -f / zz, f * y / zz ** 2], 1)
        for k in range(E):
            c = pose_col.get(int(ef[k]))
            if c is not None:
                J[2 * k:2 * k + 2, c:c + 6] = Jp[k]
            if not fix_points:
                A = -(1.0 / zz[k]) * np.array([[f, 0, -f * x[k] / zz[k]], [0, f, -f * y[k] / zz[k]]]) @ Tcw[int(ef[k])][:3, :3]
                c = pt_col[int(ep[k])]
                J[2 * k:2 * k + 2, c:c + 3] = A
        Wm = np.kron(np.diag(w), np.eye(2)) * np.kron(np.eye(E), Om) if False else None
        # H = sum_k J_k^T (w_k Omega) J_k, b = -sum_k J_k^T (w_k Omega) e_k
        JW = (J.reshape(E, 2, n) * w[:, None, None])
        OJ = np.einsum("ij,ejn->ein", Om, JW)
        H = np.einsum("eim,ein->mn", J.reshape(E, 2, n), OJ)
        b = -np.einsum("ein,ei->n", OJ, e)
        if it == 0:
            lam = 1e-5 * float(np.abs(np.diag(H)).max()) if n else 0.0
            nu = 2.0
        gain, q = 0.0, 0
        while True:
            ok = True
            try:
                dx = np.linalg.solve(H + lam * np.eye(n), b)
            except np.linalg.LinAlgError:
                ok = False
            Tt, Xt = [T.copy() for T in Tcw], X.copy()
            scale = 0.0
            if ok:
                for i, c in pose_col.items():
                    Tt[i] = se3_exp(dx[c:c + 6]) @ Tcw[i]
                for l, c in pt_col.items():
                    Xt[l] = X[l] + dx[c:c + 3]
                scale = float(dx @ (lam * dx + b))
            chi_t = robust_chi2(Tt, Xt) if ok else np.inf
            gain = (current - chi_t) / (scale + 1e-3)
            acc = bool(gain > 0 and np.isfinite(chi_t))
            trace.append((it, lam, chi_t, gain, acc))
            if acc:
                alpha = min(1.0 - (2 * gain - 1) ** 3, 2.0 / 3.0)
                lam *= max(1.0 / 3.0, alpha)
                nu = 2.0
                current, Tcw, X = chi_t, Tt, Xt
            else:
                lam *= nu
                nu *= 2
            q += 1
            if not (gain < 0 and q < 10):
                break
        if q == 10 or gain == 0:
            break
    return np.stack([np.linalg.inv(T) for T in Tcw]), X, trace

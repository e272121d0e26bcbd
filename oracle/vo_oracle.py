"""TEST ORACLE / CPU BASELINE — Python restatement of the reference's per-frame tracking path on top
of the same third-party code the reference calls (OpenCV via cv2) and the oracle restatements:

  Frame::calcKeyPoints / calcDescriptors   include/my_slam/vo/frame.h:73-86 -> src/geometry/feature_match.cpp:11-49
  getMappointsInCurrentView_               src/vo/vo.cpp:16-49
  matchFeatures                            src/geometry/feature_match.cpp:126-239 (method 1 = exact Hamming NN)
  poseEstimationPnP_                       src/vo/vo.cpp:267-381  (cv2.solvePnPRansac, the identical call)
  callBundleAdjustment_ + bundleAdjustment src/vo/vo.cpp:384-478, src/optimization/g2o_ba.cpp:172-317
                                           (oracle/ba_oracle.c: g2o restated, parity vs real g2o unpinned)
Not product code.  The reference binary itself cannot be built here (no OpenCV C++/g2o/Sophus/PCL)."""
from __future__ import annotations

import numpy as np

from . import oracle_lib


def _cv_kp(kps):
    return np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id) for k in kps], oracle_lib.KEYPOINT_DTYPE)


class CpuTracker:
    def __init__(self, K, rows, cols, max_keypoints=1500, match_method=1, match_radius=50.0, ba_iterations=50,
                 ba_window=5, ba_fix_points=True, ba_enable=True, buffer_size=20):
        import cv2
        self.cv2 = cv2
        self.K = np.asarray(K, np.float64)
        self.rows, self.cols = rows, cols
        self.max_keypoints = max_keypoints
        self.match_method, self.match_radius = match_method, match_radius
        self.ba_iterations, self.ba_window, self.ba_fix_points, self.ba_enable = ba_iterations, ba_window, ba_fix_points, ba_enable
        self.buffer_size = buffer_size
        self.orb_detect = cv2.ORB_create(8000, 1.2, 4, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 20)   # feature_match.cpp:22-23
        self.orb_compute = cv2.ORB_create(8000, 1.2, 4)                                         # feature_match.cpp:45
        self.bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        self.reset(np.eye(4))

    def set_map(self, pts3d, desc):
        self.map_pts = np.ascontiguousarray(pts3d, np.float32).copy()
        self.map_desc = np.ascontiguousarray(desc, np.uint8).copy()

    def reset(self, T_ref):
        self.T_ref = np.array(T_ref, np.float64).copy()
        self.frames = []
        self.T_prev = None

    def extract(self, image):
        cv2 = self.cv2
        kp = _cv_kp(self.orb_detect.detect(image, None))
        kp = oracle_lib.select_uniform_kpts_by_grid(kp, image.shape[0], image.shape[1], self.max_keypoints, 16, 8)
        ck = [cv2.KeyPoint(float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]),
                           int(k["octave"]), int(k["class_id"])) for k in kp]
        _, desc = self.orb_compute.compute(image, ck)
        return kp, (desc if desc is not None else np.zeros((0, 32), np.uint8))

    def candidates(self, T_w_c):
        Tcw = np.linalg.inv(T_w_c)
        P = self.map_pts.astype(np.float64)
        pc = (P @ Tcw[:3, :3].T + Tcw[:3, 3]).astype(np.float32)                 # preTranslatePoint3f -> Point3f
        with np.errstate(divide="ignore", invalid="ignore"):
            u = (self.K[0, 0] * pc[:, 0].astype(np.float64) / pc[:, 2] + self.K[0, 2]).astype(np.float32)
            v = (self.K[1, 1] * pc[:, 1].astype(np.float64) / pc[:, 2] + self.K[1, 2]).astype(np.float32)
        ok = ~(pc[:, 2] < 0) & (u > 0) & (v > 0) & (u < self.cols) & (v < self.rows)
        idx = np.flatnonzero(ok)
        return idx, np.stack([u[idx], v[idx]], 1)

    def track(self, image):
        cv2 = self.cv2
        info = {}
        kp, desc = self.extract(image)
        frame = {"T": self.T_ref.copy(), "obs": np.zeros((0, 2), np.float32), "map_idx": np.zeros(0, np.int32)}
        self.frames.append(frame)
        if len(self.frames) > self.buffer_size:
            self.frames.pop(0)
        cand_idx, cand_xy = self.candidates(frame["T"])
        kp_xy = np.stack([kp["x"], kp["y"]], 1) if len(kp) else np.zeros((0, 2), np.float32)
        if len(cand_idx) and len(kp) and self.match_method == 1:
            # cv::BFMatcher(NORM_HAMMING).match is the exact search the reference's FLANN-LSH approximates; it is
            # bit-identical to oracle_lib.hamming_nn (tests/test_match_oracle.py) and SIMD-optimised like the
            # matcher the reference links, so the CPU baseline is not handicapped by a scalar loop
            ms = self.bf.match(self.map_desc[cand_idx], desc)
            allm = np.array([(x.queryIdx, x.trainIdx, x.imgIdx, x.distance) for x in ms], oracle_lib.DMATCH_DTYPE)
            m = oracle_lib.threshold_and_dedup(allm)
        elif len(cand_idx) and len(kp):
            m = oracle_lib.match_features(self.map_desc[cand_idx], desc, self.match_method, cand_xy, kp_xy, self.match_radius)
        else:
            m = np.zeros(0, oracle_lib.DMATCH_DTYPE)
        info.update(n_keypoints=len(kp), n_candidates=len(cand_idx), n_matches=len(m))
        pnp_ok = len(m) >= 5
        n_inl = 0
        if pnp_ok:
            p3 = self.map_pts[cand_idx[m["query_idx"]]]
            p2 = kp_xy[m["train_idx"]].astype(np.float32)
            ok, rvec, tvec, inl = cv2.solvePnPRansac(p3, p2, self.K, None, None, None, False, 100, 2.0, 0.999)   # vo.cpp:318-320
            if not ok or inl is None:
                pnp_ok = False
            else:
                inl = inl.ravel()
                n_inl = len(inl)
                frame["obs"] = p2[inl]
                frame["map_idx"] = cand_idx[m["query_idx"][inl]].astype(np.int32)
                R, _ = cv2.Rodrigues(rvec)
                T = np.eye(4)
                T[:3, :3], T[:3, 3] = R, tvec.ravel()
                frame["T"] = np.linalg.inv(T)
                if self.T_prev is not None and np.linalg.norm(frame["T"][:3, 3] - self.T_prev[:3, 3]) >= 0.3:
                    pnp_ok = False
        if not pnp_ok and self.T_prev is not None:
            frame["T"] = self.T_prev.copy()
        info.update(n_inliers=n_inl, pnp_ok=int(pnp_ok), T_pnp=frame["T"].copy(), ba_frames=0)
        if pnp_ok and self.ba_enable:
            total = len(self.frames)
            nba = min(self.ba_window, total - 1)
            sel = [b for b in range(total - 1, total - 1 - nba, -1) if len(self.frames[b]["map_idx"]) >= 3]
            if sel:
                ef = np.concatenate([np.full(len(self.frames[b]["map_idx"]), i, np.int32) for i, b in enumerate(sel)])
                ep_raw = np.concatenate([self.frames[b]["map_idx"] for b in sel])
                ob = np.concatenate([self.frames[b]["obs"] for b in sel]).astype(np.float32)
                used, ep = np.unique(ep_raw, return_inverse=True)
                poses = np.stack([self.frames[b]["T"] for b in sel])
                poses, pts, _ = oracle_lib.bundle_adjustment(poses, self.map_pts[used], ef, ep.astype(np.int32), ob, self.K,
                                                             fix_points=self.ba_fix_points, update_points=not self.ba_fix_points,
                                                             iterations=self.ba_iterations)
                for i, b in enumerate(sel):
                    self.frames[b]["T"] = poses[i]
                if not self.ba_fix_points:
                    self.map_pts[used] = pts
                info["ba_frames"] = len(sel)
        Tc = self.frames[-1]["T"]
        if pnp_ok and np.linalg.norm(Tc[:3, 3] - self.T_ref[:3, 3]) > 0.03:
            self.T_ref = Tc.copy()
        self.T_prev = Tc.copy()
        return Tc.copy(), info

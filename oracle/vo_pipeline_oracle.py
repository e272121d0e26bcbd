"""TEST ORACLE — Python restatement of the reference's whole VisualOdometry::addFrame state machine on top of the
third-party code it calls (OpenCV via cv2) and the other oracle restatements.  Not product code.

  addFrame                          src/vo/vo_addFrame.cpp:10-142
  estimateMotionAnd3DPoints_        src/vo/vo.cpp:53-111  (helperEstimatePossibleRelativePosesByEpipolarGeometry,
                                    src/geometry/motion_estimation.cpp:10-158)
  isVoGoodToInit_ / retainGoodTriangulationResult_ / checkLargeMoveForAddKeyFrame_   vo.cpp:113-265 (vo_init_oracle.py)
  poseEstimationPnP_ / callBundleAdjustment_   vo.cpp:267-478
  addKeyFrame_ / optimizeMap_ / pushCurrPointsToMap_ / getViewAngle_   vo.cpp:482-584
  helperFindInlierMatchesByEpipolarCons / helperTriangulatePoints     motion_estimation.cpp:180-240

Container order: Map::map_points_ and Frame::inliers_to_mappt_connections_ are std::unordered_map<int, ...> in the
reference and are walked in container order; StdIntMap keeps a real libstdc++ container (oracle/stdmap_oracle.cpp)
beside the Python values so that the candidate order of the PnP stage is the reference's.

Where the reference would throw (OpenCV assertion on fewer than 5 matched points in findEssentialMat, empty rvec after a
failed solvePnPRansac) this restatement skips the frame / reports a failed PnP instead; noted at the call sites.
"""
from __future__ import annotations

import ctypes as C
from collections import deque

import numpy as np

from . import epipolar_oracle, motion_oracle, oracle_lib, vo_init_oracle

DEFAULTS = dict(                                      # config/config.yaml of the reference
    max_number_of_keypoints=1500,
    feature_match_method_index_initialization=1, feature_match_method_index_pnp=1,
    max_matching_pixel_dist_in_initialization=100.0, max_matching_pixel_dist_in_triangulation=100.0,
    max_matching_pixel_dist_in_pnp=50.0,
    findEssentialMat_prob=0.999, findEssentialMat_threshold=1.0,
    min_triang_angle=1.0, max_ratio_between_max_angle_and_median_angle=20.0,
    min_inlier_matches=15, min_pixel_dist=50.0, min_median_triangulation_angle=2.0,
    assumed_mean_pts_depth_during_vo_init=0.8, min_dist_between_two_keyframes=0.03,
    max_possible_dist_to_prev_keyframe=0.3, is_enable_ba=True, num_prev_frames_to_opti_by_ba=5,
    is_ba_fix_map_points=True, ba_iterations=50, init_calc_homography=True)

BLANK, DOING_INITIALIZATION, DOING_TRACKING, LOST = 0, 1, 2, 3


class StdIntMap:
    """dict with the iteration order of std::unordered_map<int, T> (the real container, see module docstring)."""

    def __init__(self):
        L = oracle_lib.lib()
        L.orc_umap_new.restype = C.c_void_p
        for f in (L.orc_umap_free, L.orc_umap_size, L.orc_umap_insert, L.orc_umap_touch, L.orc_umap_erase, L.orc_umap_contains):
            f.argtypes = [C.c_void_p] + ([C.c_int] if f not in (L.orc_umap_free, L.orc_umap_size) else [])
        L.orc_umap_keys.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        self._L, self._h, self._v = L, C.c_void_p(L.orc_umap_new()), {}

    def __del__(self):
        self._L.orc_umap_free(self._h)

    def __len__(self):
        return len(self._v)

    def __contains__(self, k):
        return int(k) in self._v

    def __getitem__(self, k):
        return self._v[int(k)]

    def __setitem__(self, k, v):                       # operator[] =
        self._L.orc_umap_touch(self._h, int(k))
        self._v[int(k)] = v

    def insert(self, k, v):                            # unordered_map::insert: an existing entry wins
        if self._L.orc_umap_insert(self._h, int(k)):
            self._v[int(k)] = v
            return True
        return False

    def erase(self, k):
        self._L.orc_umap_erase(self._h, int(k))
        self._v.pop(int(k), None)

    def keys(self):
        out = np.zeros(max(len(self._v), 1), np.int32)
        n = self._L.orc_umap_keys(self._h, out.ctypes.data, len(out))
        assert n == len(self._v)
        return out[:n].tolist()


class MapPoint:
    def __init__(self, pid, pos, desc, norm, color):
        self.id, self.pos, self.desc, self.norm, self.color = pid, np.asarray(pos, np.float32), desc.copy(), norm, color
        self.visible_times, self.matched_times = 1, 1          # mappoint.cpp:17


class Frame:
    def __init__(self, fid, image):
        self.id, self.image = fid, image
        self.kp = np.zeros(0, oracle_lib.KEYPOINT_DTYPE)
        self.desc = np.zeros((0, 32), np.uint8)
        self.colors = np.zeros((0, 3), np.uint8)
        self.T_w_c = np.eye(4)
        empty = np.zeros(0, oracle_lib.DMATCH_DTYPE)
        self.matches_with_ref, self.inliers_matches_with_ref, self.inliers_matches_for_3d, self.matches_with_map = empty, empty, empty, empty
        self.inliers_pts3d = np.zeros((0, 3), np.float32)
        self.triangulation_angles = np.zeros(0)
        self.conn = StdIntMap()                                # keypoint index -> (pt_ref_idx, pt_map_idx)

    def xy(self):
        return np.stack([self.kp["x"], self.kp["y"]], 1).astype(np.float32) if len(self.kp) else np.zeros((0, 2), np.float32)


def _pretranslate(p, T):
    """basics::preTranslatePoint3f: double arithmetic, narrowed to Point3f."""
    p = np.asarray(p, np.float32).astype(np.float64).reshape(-1, 3)
    return (p @ T[:3, :3].T + T[:3, 3]).astype(np.float32)


def _cam2pixel(pc, K):
    pc = np.asarray(pc, np.float32).reshape(-1, 3)
    with np.errstate(divide="ignore", invalid="ignore"):
        u = (K[0, 0] * pc[:, 0].astype(np.float64) / pc[:, 2] + K[0, 2]).astype(np.float32)
        v = (K[1, 1] * pc[:, 1].astype(np.float64) / pc[:, 2] + K[1, 2]).astype(np.float32)
    return u, v


def _norm_plane(xy, K):
    """geometry::pixel2CamNormPlane (camera.cpp:10-15): double arithmetic, narrowed to Point2f."""
    xy = np.asarray(xy, np.float32).astype(np.float64)
    return np.stack([(xy[:, 0] - K[0, 2]) / K[0, 0], (xy[:, 1] - K[1, 2]) / K[1, 1]], 1).astype(np.float32)


def _rt2T(R, t):
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, np.asarray(t).ravel()
    return T


def _sub(m, idx):
    """DMatch(queryIdx, trainIdx, distance) rebuilt from a list of indices (imgIdx = -1, motion_estimation.cpp:104-106)."""
    out = np.ascontiguousarray(m[np.asarray(idx, np.int64)]).copy() if len(idx) else np.zeros(0, oracle_lib.DMATCH_DTYPE)
    out["img_idx"] = -1
    return out


class CpuVo:
    def __init__(self, K, rows, cols, **cfg):
        import cv2
        self.cv2 = cv2
        self.K = np.asarray(K, np.float64)
        self.rows, self.cols = rows, cols
        self.cfg = dict(DEFAULTS)
        unknown = set(cfg) - set(DEFAULTS)
        assert not unknown, unknown
        self.cfg.update(cfg)
        self.orb_detect = cv2.ORB_create(8000, 1.2, 4, 31, 0, 2, cv2.ORB_HARRIS_SCORE, 31, 20)   # feature_match.cpp:22-23
        self.orb_compute = cv2.ORB_create(8000, 1.2, 4)                                         # feature_match.cpp:45
        self.bf = cv2.BFMatcher(cv2.NORM_HAMMING)
        self.state = BLANK
        self.map = StdIntMap()
        self.keyframes = {}
        self.buff = deque()
        self.curr = self.prev = self.ref = self.prev_ref = None
        self.frame_factory_id = 0
        self.point_factory_id = 0
        self.map_point_erase_ratio = 0.1                           # function-local static of optimizeMap_ (vo.cpp:490-491)
        self.log = []

    # ---- Frame::calcKeyPoints / calcDescriptors (frame.h:73-86) ----
    def _extract(self, frame):
        cv2 = self.cv2
        img = frame.image
        kps = self.orb_detect.detect(img, None)
        kp = np.array([(k.pt[0], k.pt[1], k.size, k.angle, k.response, k.octave, k.class_id) for k in kps], oracle_lib.KEYPOINT_DTYPE)
        kp = oracle_lib.select_uniform_kpts_by_grid(kp, img.shape[0], img.shape[1], self.cfg["max_number_of_keypoints"], 16, 8)
        ck = [cv2.KeyPoint(float(k["x"]), float(k["y"]), float(k["size"]), float(k["angle"]), float(k["response"]), int(k["octave"]), int(k["class_id"])) for k in kp]
        _, desc = self.orb_compute.compute(img, ck)
        frame.kp, frame.desc = kp, (desc if desc is not None else np.zeros((0, 32), np.uint8))
        x, y = np.floor(kp["x"]).astype(np.int64), np.floor(kp["y"]).astype(np.int64)
        bgr = img[y, x] if img.ndim == 3 else np.repeat(img[y, x][:, None], 3, 1)
        frame.colors = bgr[:, ::-1].copy()                        # getPixelAt returns {r, g, b} (opencv_funcs.cpp:10-33)

    def _match(self, d1, d2, method, xy1, xy2, radius):
        if len(d1) == 0 or len(d2) == 0:
            return np.zeros(0, oracle_lib.DMATCH_DTYPE)
        if method == 1:                                           # exact Hamming NN in place of FLANN-LSH, like the product
            ms = self.bf.match(d1, d2)
            allm = np.array([(x.queryIdx, x.trainIdx, x.imgIdx, x.distance) for x in ms], oracle_lib.DMATCH_DTYPE)
            return oracle_lib.threshold_and_dedup(allm)
        return oracle_lib.match_features(d1, d2, method, xy1, xy2, radius)

    # ---- addFrame (vo_addFrame.cpp:10-142) ----
    def add_frame(self, image):
        c = self.cfg
        frame = Frame(self.frame_factory_id, image)
        self.frame_factory_id += 1
        self.buff.append(frame)                                   # pushFrameToBuff_ (vo.h:81-86)
        if len(self.buff) > 20:
            self.buff.popleft()
        self.curr = frame
        info = dict(frame_id=frame.id, state_in=self.state, keyframe=0, pnp_ok=0, n_inliers=0, ba_frames=0, n_matches=0, best_sol=-1)
        self._extract(frame)
        info["n_keypoints"] = len(frame.kp)
        self.prev_ref = self.ref
        if self.state == BLANK:
            frame.T_w_c = np.eye(4)
            self.state = DOING_INITIALIZATION
            self._add_keyframe(frame)
            info["keyframe"] = 1
        elif self.state == DOING_INITIALIZATION:
            ref = self.ref
            frame.matches_with_ref = self._match(ref.desc, frame.desc, c["feature_match_method_index_initialization"], ref.xy(), frame.xy(),
                                                 c["max_matching_pixel_dist_in_initialization"])
            info["n_matches"] = len(frame.matches_with_ref)
            ok = self._estimate_motion_and_3d_points(info)
            if ok and self._is_vo_good_to_init(info):
                self._push_curr_points_to_map()
                self._add_keyframe(frame)
                self.state = DOING_TRACKING
                info["keyframe"] = 1
            else:
                frame.T_w_c = ref.T_w_c.copy()
        elif self.state == DOING_TRACKING:
            frame.T_w_c = self.ref.T_w_c.copy()
            if self._pose_estimation_pnp(info):
                self._call_bundle_adjustment(info)
                large, _, _ = vo_init_oracle.check_large_move(frame.T_w_c, self.ref.T_w_c, c["min_dist_between_two_keyframes"])
                if large:
                    self._insert_keyframe(info)
        info["state_out"] = self.state
        info["map_points"] = len(self.map)
        self.prev = frame
        self.log.append(info)
        return frame.T_w_c.copy(), info

    # ---- estimateMotionAnd3DPoints_ (vo.cpp:53-111) ----
    def _estimate_motion_and_3d_points(self, info):
        c, K, ref, cur = self.cfg, self.K, self.ref, self.curr
        m = cur.matches_with_ref
        if len(m) < 8:                                            # the reference would hit an OpenCV assertion below 5 points; see module docstring
            return False
        p1, p2 = ref.xy()[m["query_idx"]], cur.xy()[m["train_idx"]]
        np1, np2 = _norm_plane(p1, K), _norm_plane(p2, K)
        E, R_e, t_e, inl_e = epipolar_oracle.esti_motion_by_essential(p1, p2, K, c["findEssentialMat_prob"], c["findEssentialMat_threshold"])
        list_R, list_t, list_n, list_inl = [R_e], [t_e], [np.zeros(3)], [inl_e]
        H, inl_h = None, np.zeros(0, np.int32)
        if c["init_calc_homography"]:
            H, Rs, ts, ns, inl_h = epipolar_oracle.esti_motion_by_homography(p1, p2, K, 3.0)
            for s in epipolar_oracle.remove_wrong_rt_of_homography(np1, np2, inl_h, Rs, ts, ns):
                list_R.append(Rs[s]); list_t.append(ts[s]); list_n.append(ns[s]); list_inl.append(inl_h)
        pts3d = [epipolar_oracle.do_triangulation(np1, np2, list_R[i], list_t[i], list_inl[i]) for i in range(len(list_R))]
        best = 0
        if c["init_calc_homography"]:
            score_e, _ = motion_oracle.check_essential_score(E, K, p1, p2, inl_e)
            score_h, _ = motion_oracle.check_homography_score(H, p1, p2, inl_h)
            best, ratio = motion_oracle.choose_e_or_h(score_e, score_h, np.array(list_n[1:]).reshape(-1, 3))
            info.update(score_e=score_e, score_h=score_h, eh_ratio=ratio)
        info["best_sol"] = best
        R, t = np.asarray(list_R[best], np.float64), np.asarray(list_t[best], np.float64).ravel()
        cur.inliers_matches_with_ref = _sub(m, list_inl[best])
        p_in_curr = (pts3d[best].astype(np.float64) @ R.T + t).astype(np.float32)            # basics::transCoord (vo.cpp:84-86)
        cur.inliers_pts3d = p_in_curr
        cur.T_w_c = ref.T_w_c @ np.linalg.inv(_rt2T(R, t))                                   # :91
        self._retain_good_triangulation_result()
        n = len(cur.inliers_pts3d)
        if n < 20:                                                                            # :97-101
            return True
        pts, t_scaled, _ = vo_init_oracle.normalize_init_depth(cur.inliers_pts3d, t, c["assumed_mean_pts_depth_during_vo_init"])
        cur.inliers_pts3d = pts
        cur.T_w_c = ref.T_w_c @ np.linalg.inv(_rt2T(R, t_scaled))                            # :110
        return True

    def _retain_good_triangulation_result(self):
        c, cur = self.cfg, self.curr
        keep, ang = vo_init_oracle.retain_good_triangulation(cur.inliers_pts3d, cur.T_w_c, self.ref.T_w_c, c["min_triang_angle"],
                                                             c["max_ratio_between_max_angle_and_median_angle"])
        if len(cur.inliers_pts3d) == 0:
            return
        cur.inliers_matches_for_3d = np.concatenate([cur.inliers_matches_for_3d, cur.inliers_matches_with_ref[keep]])
        cur.inliers_pts3d = cur.inliers_pts3d[keep]
        cur.triangulation_angles = np.concatenate([cur.triangulation_angles, ang])

    def _is_vo_good_to_init(self, info):
        c, cur = self.cfg, self.curr
        m = cur.inliers_matches_for_3d
        good, mean_dist, med = vo_init_oracle.is_vo_good_to_init(self.ref.xy()[m["query_idx"]], cur.xy()[m["train_idx"]], cur.triangulation_angles,
                                                                 c["min_inlier_matches"], c["min_pixel_dist"], c["min_median_triangulation_angle"])
        info.update(init_mean_pixel_dist=mean_dist, init_median_angle=med, n_inliers=len(m))
        return good

    # ---- mapping (vo.cpp:482-584) ----
    def _add_keyframe(self, frame):
        self.keyframes[frame.id] = frame
        self.ref = frame

    def _push_curr_points_to_map(self):
        cur, ref = self.curr, self.ref
        cam_center = cur.T_w_c[:3, 3].copy()
        for i, dm in enumerate(cur.inliers_matches_for_3d):
            q, pt_idx = int(dm["query_idx"]), int(dm["train_idx"])
            if q in ref.conn:                                                     # already a map point in the reference keyframe
                pid = ref.conn[q][1]
            else:
                world = _pretranslate(cur.inliers_pts3d[i], cur.T_w_c)[0]
                d = world.astype(np.float64) - cam_center
                mp = MapPoint(self.point_factory_id, world, cur.desc[pt_idx], d / np.sqrt((d * d).sum()), cur.colors[pt_idx].copy())
                self.point_factory_id += 1
                pid = mp.id
                self.map[pid] = mp                                                # Map::insertMapPoint
            cur.conn.insert(pt_idx, (q, pid))

    def _in_frame(self, frame, pos):
        pc = _pretranslate(pos, np.linalg.inv(frame.T_w_c))
        u, v = _cam2pixel(pc, self.K)
        return bool(not (pc[0, 2] < 0) and u[0] > 0 and v[0] > 0 and u[0] < self.cols and v[0] < self.rows)

    def _optimize_map(self):
        cur = self.curr
        cam_center = cur.T_w_c[:3, 3]
        for pid in self.map.keys():
            mp = self.map[pid]
            if not self._in_frame(cur, mp.pos):
                self.map.erase(pid)
                continue
            if np.float32(mp.matched_times) / np.float32(mp.visible_times) < self.map_point_erase_ratio:
                self.map.erase(pid)
                continue
            n = mp.pos.astype(np.float64) - cam_center                            # getViewAngle_ (vo.cpp:578-584)
            n = n / np.sqrt((n * n).sum())
            with np.errstate(invalid="ignore"):
                angle = np.arccos(n @ mp.norm)                                    # NaN for |cos| > 1 by rounding: not erased, as in C++
            if angle > np.pi / 4.0:
                self.map.erase(pid)
        if len(self.map) > 1000:
            self.map_point_erase_ratio += 0.05
        else:
            self.map_point_erase_ratio = 0.1

    # ---- poseEstimationPnP_ (vo.cpp:267-381) ----
    def _candidates(self):
        ids = self.map.keys()
        if not ids:
            return [], np.zeros((0, 2), np.float32), np.zeros((0, 32), np.uint8)
        pos = np.stack([self.map[i].pos for i in ids])
        pc = _pretranslate(pos, np.linalg.inv(self.curr.T_w_c))
        u, v = _cam2pixel(pc, self.K)
        ok = ~(pc[:, 2] < 0) & (u > 0) & (v > 0) & (u < self.cols) & (v < self.rows)
        sel = [ids[i] for i in np.flatnonzero(ok)]
        for i in sel:
            self.map[i].visible_times += 1
        desc = np.stack([self.map[i].desc for i in sel]) if sel else np.zeros((0, 32), np.uint8)
        return sel, np.stack([u[ok], v[ok]], 1), desc

    def _pose_estimation_pnp(self, info):
        cv2, c, cur = self.cv2, self.cfg, self.curr
        cand, cand_xy, cand_desc = self._candidates()
        m = self._match(cand_desc, cur.desc, c["feature_match_method_index_pnp"], cand_xy, cur.xy(), c["max_matching_pixel_dist_in_pnp"])
        cur.matches_with_map = m
        info.update(n_candidates=len(cand), n_matches=len(m))
        good = len(m) >= 5
        if good:
            p3 = np.stack([self.map[cand[q]].pos for q in m["query_idx"]]).astype(np.float32)
            p2 = cur.xy()[m["train_idx"]]
            ok, rvec, tvec, inl = cv2.solvePnPRansac(p3, p2, self.K, None, None, None, False, 100, 2.0, 0.999)          # vo.cpp:318-320
            if not ok or inl is None:                 # the reference goes on with an empty rvec here; see module docstring
                good = False
            else:
                inl = inl.ravel()
                for gi in inl:
                    mp = self.map[cand[int(m["query_idx"][gi])]]
                    mp.matched_times += 1
                    cur.conn[int(m["train_idx"][gi])] = (-1, mp.id)
                cur.matches_with_map = m[inl]
                info["n_inliers"] = len(inl)
                R, _ = cv2.Rodrigues(rvec)
                cur.T_w_c = np.linalg.inv(_rt2T(R, tvec))
                if np.linalg.norm(cur.T_w_c[:3, 3] - self.prev.T_w_c[:3, 3]) >= c["max_possible_dist_to_prev_keyframe"]:
                    good = False
        if not good:
            cur.T_w_c = self.prev.T_w_c.copy()
        info["pnp_ok"] = int(good)
        info["T_pnp"] = cur.T_w_c.copy()
        return good

    # ---- callBundleAdjustment_ (vo.cpp:384-478) ----
    def _call_bundle_adjustment(self, info):
        c = self.cfg
        if not c["is_enable_ba"]:
            return
        total = len(self.buff)
        nba = min(c["num_prev_frames_to_opti_by_ba"], total - 1)
        sel, ef, ep, ob = [], [], [], []
        for b in range(total - 1, total - 1 - nba, -1):
            f = self.buff[b]
            if len(f.conn) < 3:
                continue
            fi = len(sel)
            sel.append(f)
            xy = f.xy()
            for kpt_idx in f.conn.keys():
                pid = f.conn[kpt_idx][1]
                if pid not in self.map:                                            # point has been deleted
                    continue
                ef.append(fi); ep.append(pid); ob.append(xy[kpt_idx])
        if not sel or not ef:
            return
        used, epi = np.unique(np.array(ep), return_inverse=True)
        pts = np.stack([self.map[int(i)].pos for i in used])
        poses = np.stack([f.T_w_c for f in sel])
        fix = c["is_ba_fix_map_points"]
        poses, pts, _ = oracle_lib.bundle_adjustment(poses, pts, np.array(ef, np.int32), epi.astype(np.int32), np.array(ob, np.float32), self.K,
                                                     fix_points=fix, update_points=not fix, iterations=c["ba_iterations"])
        for i, f in enumerate(sel):
            f.T_w_c = poses[i]
        if not fix:
            for k, i in enumerate(used):
                self.map[int(i)].pos = pts[k]
        info["ba_frames"], info["ba_edges"] = len(sel), len(ef)

    # ---- the keyframe branch of addFrame (vo_addFrame.cpp:93-124) ----
    def _insert_keyframe(self, info):
        c, K, cur, ref = self.cfg, self.K, self.curr, self.ref
        cur.matches_with_ref = self._match(ref.desc, cur.desc, c["feature_match_method_index_pnp"], ref.xy(), cur.xy(),
                                           c["max_matching_pixel_dist_in_triangulation"])
        m = cur.matches_with_ref
        info["kf_matches"] = len(m)
        if len(m) < 8:                                                            # see module docstring
            return
        p1, p2 = ref.xy()[m["query_idx"]], cur.xy()[m["train_idx"]]
        _, _, _, inl = epipolar_oracle.esti_motion_by_essential(p1, p2, K, c["findEssentialMat_prob"], c["findEssentialMat_threshold"])
        cur.inliers_matches_with_ref = _sub(m, inl)                               # helperFindInlierMatchesByEpipolarCons
        mi = cur.inliers_matches_with_ref
        T = np.linalg.inv(cur.T_w_c) @ ref.T_w_c                                  # getMotionFromFrame1to2(curr_, ref_) (vo_commons.cpp:9-15)
        R, t = T[:3, :3], T[:3, 3]
        np1, np2 = _norm_plane(ref.xy()[mi["query_idx"]], K), _norm_plane(cur.xy()[mi["train_idx"]], K)
        in_prev = epipolar_oracle.do_triangulation(np1, np2, R, t, np.arange(len(mi)))
        cur.inliers_pts3d = (in_prev.astype(np.float64) @ R.T + t).astype(np.float32)
        self._retain_good_triangulation_result()
        info["kf_new_points_in"] = len(cur.inliers_matches_for_3d)
        self._push_curr_points_to_map()
        self._optimize_map()
        self._add_keyframe(cur)
        info["keyframe"] = 1


def trajectory_error(T_est, T_true):
    """RMS position error after the best similarity alignment (Umeyama) of the estimated camera centres to the
    true ones — monocular VO has a free scale."""
    a = np.stack([T[:3, 3] for T in T_est])
    b = np.stack([T[:3, 3] for T in T_true])
    ma, mb = a.mean(0), b.mean(0)
    A, B = a - ma, b - mb
    U, S, Vt = np.linalg.svd(B.T @ A / len(a))
    D = np.diag([1, 1, np.sign(np.linalg.det(U @ Vt))])
    R = U @ D @ Vt
    s = np.trace(np.diag(S) @ D) / (A * A).sum() * len(a)
    aligned = (s * (R @ A.T)).T + mb
    return float(np.sqrt(((aligned - b) ** 2).sum(1).mean())), float(s)

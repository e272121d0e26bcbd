"""The my_slam adapter layer (monocular-visual-odometry_b200/my_slam_adapter/): the reference's own C++ signatures
(include/my_slam/geometry/feature_match.h:12-54, include/my_slam/optimization/g2o_ba.h:16-30, the inline
cv::solvePnPRansac call of src/vo/vo.cpp:318-320) implemented on top of the C ABI.  tests/cpp/adapter_demo.cpp drives it
the way the reference's callers do (test/test_epipolor_geometry.cpp:91-98, vo.cpp:293-337, :428-462); here its outputs
are compared with the oracles and with the same calls made through ctypes."""
import subprocess
from pathlib import Path

import numpy as np
import pytest

import mvo_synth

ROOT = Path(__file__).resolve().parent.parent
DEMO = ROOT / "monocular-visual-odometry_b200" / "build" / "adapter_demo"


def test_adapter_layer_builds_with_reference_signatures(built):
    """CPU: the adapters compile against the cv:: value types and export the reference's mangled symbols."""
    assert DEMO.exists(), "make -C monocular-visual-odometry_b200 did not build the adapter demo"
    syms = subprocess.run(["nm", "-C", "--defined-only", str(DEMO)], capture_output=True, text=True, check=True).stdout
    for want in ("my_slam::geometry::calcKeyPoints(cv::Mat const&, std::vector<cv::KeyPoint",
                 "my_slam::geometry::calcDescriptors(cv::Mat const&, std::vector<cv::KeyPoint",
                 "my_slam::geometry::matchFeatures(cv::Mat_<unsigned char> const&, cv::Mat_<unsigned char> const&, std::vector<cv::DMatch",
                 "my_slam::geometry::matchByRadiusAndBruteForce(", "my_slam::geometry::removeDuplicatedMatches(",
                 "my_slam::geometry::selectUniformKptsByGrid(", "my_slam::geometry::computeMeanDistBetweenKeypoints(",
                 "my_slam::geometry::inliers2DMatches(", "my_slam::geometry::pts2Keypts(",
                 "my_slam::optimization::bundleAdjustment(", "my_slam::optimization::optimizeSingleFrame(",
                 "my_slam::geometry::estiMotionByEssential(std::vector<cv::Point2f", "my_slam::geometry::doTriangulation(std::vector<cv::Point2f",
                 "my_slam::geometry::estiMotionByHomography(std::vector<cv::Point2f", "my_slam::geometry::removeWrongRtOfHomography(std::vector<cv::Point2f"):
        assert want in syms, want


@pytest.mark.gpu
def test_adapter_demo_matches_oracles_and_capi(ctx, tmp_path):
    import mvo_b200
    from oracle import oracle_lib, orb_oracle
    K = mvo_synth.K_DEFAULT
    img1 = mvo_synth.gray_to_bgr(mvo_synth.rect_scene(11, 640, 480))
    img2 = np.ascontiguousarray(np.roll(img1, (3, 7), axis=(0, 1)))        # shifted copy
    img1.tofile(tmp_path / "img1.bin")
    img2.tofile(tmp_path / "img2.bin")
    P, uv, rvec_t, tvec_t, _ = mvo_synth.pnp_problem(0, n=600)
    np.ascontiguousarray(P, np.float32).tofile(tmp_path / "pnp_p3.bin")
    np.ascontiguousarray(uv, np.float32).tofile(tmp_path / "pnp_p2.bin")
    np.ascontiguousarray(K, np.float64).tofile(tmp_path / "K.bin")
    pb = mvo_synth.ba_problem(3, n_frames=4, n_points=300)
    np.ascontiguousarray(pb["T_w_c"], np.float64).tofile(tmp_path / "ba_poses.bin")
    np.ascontiguousarray(pb["points"], np.float32).tofile(tmp_path / "ba_points.bin")
    np.ascontiguousarray(pb["obs"], np.float32).tofile(tmp_path / "ba_obs.bin")
    pb["edge_frame"].astype(np.int32).tofile(tmp_path / "ba_edge_frame.bin")
    pb["edge_point"].astype(np.int32).tofile(tmp_path / "ba_edge_point.bin")
    r = subprocess.run([str(DEMO), str(tmp_path)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr + r.stdout
    assert "Using method 3" in r.stdout and "adapter demo done" in r.stdout

    # ORB: bit-exact with the numpy restatement of cv::ORB + selectUniformKptsByGrid, and with the ctypes path
    ctx.set_params(max_keypoints=2000)
    try:
        kps, descs = [], []
        for tag, img in (("1", img1), ("2", img2)):
            kp = np.fromfile(tmp_path / f"kp{tag}.bin", mvo_b200.KEYPOINT_DTYPE)
            desc = np.fromfile(tmp_path / f"desc{tag}.bin", np.uint8).reshape(-1, 32)
            det = orb_oracle.detect(img)
            sel = oracle_lib.select_uniform_kpts_by_grid(det, 480, 640, 2000, 16, 8)
            assert kp.tobytes() == sel.tobytes()
            assert np.array_equal(desc, orb_oracle.compute(img, sel))
            kps.append(kp)
            descs.append(desc)
        # matching: the three methods through the adapter == oracle matchFeatures
        xy = [np.stack([k["x"], k["y"]], 1) for k in kps]
        for method in (1, 2, 3):
            m = np.fromfile(tmp_path / f"matches{method}.bin", mvo_b200.DMATCH_DTYPE)
            ref = oracle_lib.match_features(descs[0], descs[1], method, xy[0], xy[1], 50.0, xiang_gao_ratio=2.0, lowe_ratio=1.0)
            got = ctx.match_features(descs[0], descs[1], method, xy[0], xy[1], 50.0)
            assert m.tobytes() == got.tobytes() == ref.tobytes(), method
    finally:
        ctx.set_params(max_keypoints=1500)
    # PnP: pose near the generating pose, inlier indices ascending, K x 1 int32
    rt = np.fromfile(tmp_path / "pnp_rt.bin", np.float64)
    inl = np.fromfile(tmp_path / "pnp_inliers.bin", np.int32)
    assert np.abs(rt[:3] - rvec_t).max() < 5e-3 and np.abs(rt[3:] - tvec_t).max() < 2e-2
    assert len(inl) > 300 and np.all(np.diff(inl) > 0)
    # BA through raw pointers == oracle (fixed points: shipped configuration; free points: 1e-6 on the points)
    for tag, fix in (("fixed", True), ("free", False)):
        poses = np.fromfile(tmp_path / f"ba_out_poses_{tag}.bin", np.float64).reshape(-1, 4, 4)
        pts = np.fromfile(tmp_path / f"ba_out_points_{tag}.bin", np.float32).reshape(-1, 3)
        op, opts, _ = oracle_lib.bundle_adjustment(pb["T_w_c"], pb["points"], pb["edge_frame"], pb["edge_point"], pb["obs"], K,
                                                   fix_points=fix, update_points=not fix, iterations=50)
        gp, gpts, _ = ctx.bundle_adjustment(pb["T_w_c"], pb["points"], pb["edge_frame"], pb["edge_point"], pb["obs"], K,
                                            fix_points=fix, update_points=not fix)
        assert np.abs(poses - gp).max() < 1e-12 and np.abs(pts - gpts).max() == 0
        if fix:
            assert np.abs(poses - op).max() < 1e-8 and np.array_equal(pts, pb["points"])

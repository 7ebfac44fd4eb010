"""GPU: the whole reference flow for one camera -- images -> detectTarget -> P3P-RANSAC -> computeCalibration -- through
ezcalib_b200.ezcalibrator.run_calibration, against the same flow assembled from the reference's own building blocks
(cv2.findChessboardCorners + cv2.cornerSubPix, cv2.solvePnPRansac, the restated-Ceres oracle)."""
import os
import sys

import numpy as np
import pytest

from oracle import lm_oracle as O

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "p3p_harness"))
import p3p_cv2_ref as PR  # noqa: E402

pytestmark = pytest.mark.gpu


def test_chessboard_images_to_calibration(ctx):
    cv2 = pytest.importorskip("cv2")
    cv2.ipp.setUseIPP(False)
    from ezcalib_b200 import detect, ezcalibrator as E, synth
    model, mono, n, W, H = "radtan5", True, 24, 1280, 1024
    mid = synth.MODEL_ID[model]
    d = synth.GT_DIST[model]
    f, cx, cy = 900.0, 645.0, 505.0
    tgt = synth.chessboard_target(7, 10, 0.05)
    rng = np.random.Generator(np.random.Philox(key=77))
    poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=90, tilt_sigma=0.3)
    cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
    imgs = detect.render_targets(ctx, "chessboard", 7, 10, 0.05, 0.0, mid, W, H, cams, poses, seed=123)
    host = imgs.download()
    imgs.close()
    cam, target, processed = E.run_calibration(ctx, host, "chessboard", 7, 10, 0.05, 0.0, model, mono, prior_fov_deg=70.0)
    assert len(cam.frames) >= n - 2
    # ground truth (rendered images: blur + noise): a sane calibration
    assert abs(cam.focal[0] - f) / f < 5e-3 and abs(cam.pp[0] - cx) < 3 and abs(cam.pp[1] - cy) < 3
    rm = np.array([fr.rmse_err for fr in cam.frames])
    # a view whose P3P pose is the mirror solution of the planar target can stay in a local minimum (the reference's dormant
    # refineCalibration exists for exactly that, ezcalibrator.cpp:342-557); the Cauchy loss keeps it from biasing the result
    assert np.median(rm) < 0.3 and (rm > 1.0).sum() <= 2
    # the reference flow from its own building blocks
    flags = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)
    focal0, pp0 = E.setup_initial_calibration(W, H, 70.0, mono)
    found, obs, poses0 = [], [], []
    for i in range(n):
        ok, c = cv2.findChessboardCorners(host[i], (9, 6), flags=flags)
        found.append(bool(ok))
        if ok:
            c = cv2.cornerSubPix(host[i], c, (5, 5), (-1, -1), crit).reshape(-1, 2)
    assert E.replay_skip_rule(found) == processed
    for i in processed:
        if not found[i]:
            continue
        ok, c = cv2.findChessboardCorners(host[i], (9, 6), flags=flags)
        c = cv2.cornerSubPix(host[i], c, (5, 5), (-1, -1), crit).reshape(-1, 2)
        okp, pose, _, _ = PR.compute_p3p_ransac_cv2(c, target, focal0, pp0, W, H)
        if okp:
            obs.append(c.astype(np.float32)); poses0.append(pose)
    assert len(obs) == len(cam.frames)
    f_ref, pp_ref, d_ref, poses_ref, rmse_ref, _ = O.solve_mono(mid, mono, np.stack(obs), target, W, H, focal0, pp0, np.zeros(len(d)),
                                                             np.stack(poses0))
    # detection differs by <= 1.2e-4 px, the P3P poses accordingly; both solves end in the same minimum
    assert abs(cam.focal[0] - f_ref[0]) / f_ref[0] <= 1e-6
    assert np.max(np.abs(cam.pp - pp_ref)) <= 1e-2
    assert np.max(np.abs(cam.dist - d_ref)) <= 1e-4
    assert np.max(np.abs(np.array([fr.rmse_err for fr in cam.frames]) - rmse_ref)) <= 1e-3


def test_refine_calibration_drops_bad_views_and_equals_oracle(ctx):
    """EZCalibrator::refineCalibration (ezcalibrator.cpp:342-557): one all-variable solve over the frames with rmse_err <= 1,
    against the oracle's single solve on the same subset; frames above 1 px keep their pose and rmse_err."""
    from ezcalib_b200 import ezcalibrator as E, synth
    model, mono, W, H = "radtan5", True, 1280, 1024
    mid = synth.MODEL_ID[model]
    tgt = synth.chessboard_target(7, 10, 0.05)
    P = synth.make_lm_problem(model, mono, 14, tgt, W, H, seed=31)
    cam = E.Camera(model, mono, W, H, P.focal_init, P.pp_init, P.dist_init)
    for i in range(14):
        cam.frames.append(E.CalibFrame(str(i), P.obs[i], P.poses_init[i]))
    # two views with grossly wrong observations: they end above 1 px after the staged solve
    rng = np.random.default_rng(5)
    for i in (4, 9):
        cam.frames[i].corners_px = (P.obs[i] + rng.normal(0, 6.0, P.obs[i].shape)).astype(np.float32)
    E.compute_calibration(ctx, cam, tgt)
    assert cam.stds is not None and len(cam.stds[2]) == 5 and all(v > 0 for v in cam.stds[0])
    bad = [i for i, f in enumerate(cam.frames) if f.rmse_err > 1.0]
    assert bad == [4, 9]
    keep_pose = {i: cam.frames[i].T_world_2_cam.copy() for i in bad}
    keep_rmse = {i: cam.frames[i].rmse_err for i in bad}
    good = [i for i in range(14) if i not in bad]
    obs = np.stack([cam.frames[i].corners_px for i in good])
    poses = np.stack([cam.frames[i].T_world_2_cam for i in good])
    f_ref, pp_ref, d_ref, poses_ref, s_ref, _ = O.solve(mid, mono, obs, tgt, W, H, cam.focal, cam.pp, cam.dist, poses)
    rm_ref = O.frame_rmse(mid, mono, obs, tgt, W, H, f_ref, pp_ref, d_ref, poses_ref)
    summ = E.refine_calibration(ctx, cam, tgt)
    assert summ["iterations"] == s_ref["iterations"] and summ["termination"] == s_ref["termination"]
    assert abs(cam.focal[0] - f_ref[0]) / f_ref[0] <= 1e-6 and np.max(np.abs(cam.pp - pp_ref) / pp_ref) <= 1e-6
    assert np.max(np.abs(cam.dist - d_ref)) <= 1e-6 * max(1.0, np.max(np.abs(d_ref)))
    assert np.max(np.abs(np.array([cam.frames[i].rmse_err for i in good]) - rm_ref)) <= 1e-6
    for i in bad:
        assert np.array_equal(cam.frames[i].T_world_2_cam, keep_pose[i]) and cam.frames[i].rmse_err == keep_rmse[i]

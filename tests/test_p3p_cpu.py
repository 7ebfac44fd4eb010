"""CPU: the product's P3P-RANSAC + EPnP mathematics (csrc/p3p_core.h compiled for the host, tests/p3p_harness) against the
real cv2.solvePnPRansac called with the reference's arguments (/root/reference/src/ezcalibrator.cpp:592-606).
Bar: ok flags and inlier sets equal, rvec/tvec within 1e-6 (observed ~1e-12), SE3 pose within 1e-6."""
import os
import sys

import numpy as np
import pytest

from ezcalib_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "p3p_harness"))
import p3p_cv2_ref as P  # noqa: E402

cv2 = pytest.importorskip("cv2")


def _views(target, n, seed, W=1280, H=1024, f=900.0, noise=0.1, outliers=(0,), missing=0.0, f_prior=None):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        rv = rng.normal(0, 0.35, 3)
        R, _ = cv2.Rodrigues(rv)
        centre = target.mean(0)
        t = np.array([rng.uniform(-0.1, 0.1), rng.uniform(-0.08, 0.08), rng.uniform(0.5, 0.9)]) - R @ centre
        pc = (R @ target.T).T + t
        px = np.stack([f * pc[:, 0] / pc[:, 2] + W / 2 + 3.0, f * pc[:, 1] / pc[:, 2] + H / 2 - 2.0], 1) + rng.normal(0, noise, (len(target), 2))
        no = outliers[i % len(outliers)]
        if no:
            o = rng.choice(len(target), no, replace=False)
            px[o] += rng.normal(0, 40.0, (no, 2))
        if missing:
            px[rng.random(len(target)) < missing] = -1.0
        out.append(px.astype(np.float32))
    fp = f_prior if f_prior is not None else f * 1.15
    return out, np.array([fp]), np.array([W / 2.0, H / 2.0]), W, H


def _compare(views, target, focal, pp, W, H):
    worst_rt, worst_pose, n_ok = 0.0, 0.0, 0
    for k, c in enumerate(views):
        ok0, pose0, inl0, rt0 = P.compute_p3p_ransac_cv2(c, target, focal, pp, W, H)
        ok1, pose1, inl1, rt1, nit = P.compute_p3p_ransac_harness(c, target, focal, pp, W, H)
        assert ok0 == ok1, (k, ok0, ok1)
        assert np.array_equal(inl0, inl1), (k, len(inl0), len(inl1))
        if rt0 is not None:
            worst_rt = max(worst_rt, float(np.max(np.abs(rt0 - rt1))))
        if ok0:
            n_ok += 1
            worst_pose = max(worst_pose, float(np.max(np.abs(pose0 - pose1))))
    return worst_rt, worst_pose, n_ok


def test_chessboard_views_match_cv2():
    target = synth.chessboard_target(7, 10, 0.05)
    views, focal, pp, W, H = _views(target, 60, seed=1)
    worst_rt, worst_pose, n_ok = _compare(views, target, focal, pp, W, H)
    assert n_ok == 60 and worst_rt <= 1e-6 and worst_pose <= 1e-6, (worst_rt, worst_pose)


def test_aprilgrid_views_with_outliers_and_missing_tags_match_cv2():
    target = synth.aprilgrid_target(6, 6, 0.06, 0.3)
    views, focal, pp, W, H = _views(target, 80, seed=2, W=2448, H=2048, f=1900.0, outliers=(0, 5, 20, 50), missing=0.2)
    worst_rt, worst_pose, n_ok = _compare(views, target, focal, pp, W, H)
    assert n_ok >= 70 and worst_rt <= 1e-6 and worst_pose <= 1e-6, (worst_rt, worst_pose, n_ok)


def test_too_few_points_and_too_few_inliers_fail_like_the_reference():
    target = synth.aprilgrid_target(6, 6, 0.06, 0.3)
    views, focal, pp, W, H = _views(target, 6, seed=3, W=2448, H=2048, f=1900.0, missing=0.85)
    for c in views:
        ok0, _, inl0, _ = P.compute_p3p_ransac_cv2(c, target, focal, pp, W, H)
        ok1, _, inl1, _, _ = P.compute_p3p_ransac_harness(c, target, focal, pp, W, H)
        assert ok0 == ok1 and not ok0
    # heavy contamination: fewer than 32 inliers although >= 32 points are in the image
    views, focal, pp, W, H = _views(target, 6, seed=4, W=2448, H=2048, f=1900.0, outliers=(125,))
    for c in views:
        ok0, _, inl0, _ = P.compute_p3p_ransac_cv2(c, target, focal, pp, W, H)
        ok1, _, inl1, _, _ = P.compute_p3p_ransac_harness(c, target, focal, pp, W, H)
        assert ok0 == ok1 and np.array_equal(inl0, inl1)

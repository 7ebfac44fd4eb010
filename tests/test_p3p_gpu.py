"""GPU: ezc_p3p_ransac_batch (csrc/p3p.cu, one warp per view) against the real cv2.solvePnPRansac called with the
reference's arguments (/root/reference/src/ezcalibrator.cpp:580-625) and against the host build of the same mathematics.
Bar: ok flags and inlier sets equal, poses within 1e-6."""
import os
import sys

import numpy as np
import pytest

from ezcalib_b200 import synth

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "p3p_harness"))
sys.path.insert(0, os.path.dirname(__file__))
import p3p_cv2_ref as PR  # noqa: E402
from test_p3p_cpu import _views  # noqa: E402

pytestmark = pytest.mark.gpu


def _check(ctx, target, views, focal, pp, W, H, min_ok):
    from ezcalib_b200 import ezcalibrator as E
    corners = np.stack(views)
    ok, poses, ninl, mask, rt, nit = E.compute_p3p_ransac_batch(ctx, corners, target, focal, pp, W, H, want_debug=True)
    worst_rt = worst_pose = worst_h = 0.0
    n_ok = 0
    for k, c in enumerate(views):
        ok0, pose0, inl0, rt0 = PR.compute_p3p_ransac_cv2(c, target, focal, pp, W, H)
        ok1, pose1, inl1, rt1, nit1 = PR.compute_p3p_ransac_harness(c, target, focal, pp, W, H)
        assert bool(ok[k]) == ok0, (k, ok[k], ok0)
        assert np.array_equal(np.nonzero(mask[k])[0], inl0), (k, mask[k].sum(), len(inl0))
        assert ninl[k] == len(inl0) and nit[k] == nit1
        if rt0 is not None:
            worst_rt = max(worst_rt, float(np.max(np.abs(rt[k] - rt0))))
            worst_h = max(worst_h, float(np.max(np.abs(rt[k] - rt1))))
        if ok0:
            n_ok += 1
            worst_pose = max(worst_pose, float(np.max(np.abs(poses[k] - pose0))))
    print(f"p3p: {n_ok}/{len(views)} ok, max |rvec,tvec - cv2| {worst_rt:.2e}, max |pose7 - cv2| {worst_pose:.2e}, vs host build {worst_h:.2e}")
    assert n_ok >= min_ok and worst_rt <= 1e-6 and worst_pose <= 1e-6 and worst_h <= 1e-6


def test_chessboard_views_match_cv2(ctx):
    pytest.importorskip("cv2")
    target = synth.chessboard_target(7, 10, 0.05)
    views, focal, pp, W, H = _views(target, 200, seed=21)
    _check(ctx, target, views, focal, pp, W, H, 200)


def test_aprilgrid_views_outliers_missing_tags_match_cv2(ctx):
    pytest.importorskip("cv2")
    target = synth.aprilgrid_target(6, 6, 0.06, 0.3)
    views, focal, pp, W, H = _views(target, 240, seed=22, W=2448, H=2048, f=1900.0, outliers=(0, 5, 20, 50), missing=0.2)
    _check(ctx, target, views, focal, pp, W, H, 200)


def test_failures_and_dual_focal(ctx):
    pytest.importorskip("cv2")
    from ezcalib_b200 import ezcalibrator as E
    target = synth.aprilgrid_target(6, 6, 0.06, 0.3)
    views, focal, pp, W, H = _views(target, 12, seed=23, W=2448, H=2048, f=1900.0, missing=0.85)
    views2, *_ = _views(target, 12, seed=24, W=2448, H=2048, f=1900.0, outliers=(125,))
    views = views + views2
    focal2 = np.array([focal[0], focal[0] * 1.02])
    ok, poses, ninl, mask, rt, nit = E.compute_p3p_ransac_batch(ctx, np.stack(views), target, focal2, pp, W, H, want_debug=True)
    for k, c in enumerate(views):
        ok0, pose0, inl0, rt0 = PR.compute_p3p_ransac_cv2(c, target, focal2, pp, W, H)
        assert bool(ok[k]) == ok0
        assert np.array_equal(np.nonzero(mask[k])[0], inl0)
        if not ok0:
            assert np.all(poses[k] == 0)
    assert not ok[:12].any()


def test_4k_rig_sized_batch_is_deterministic(ctx):
    """cfg4 size: 2000 views x 144 points at 3840x2160 (prior focal 15 % off, so the pinhole hypotheses do not explain
    every point): run twice -> bit-identical; a sample of the views equals cv2 (inlier sets, poses <= 1e-6)."""
    pytest.importorskip("cv2")
    from ezcalib_b200 import ezcalibrator as E
    target = synth.aprilgrid_target(6, 6, 0.06, 0.3)
    views, focal, pp, W, H = _views(target, 2000, seed=25, W=3840, H=2160, f=2800.0)
    corners = np.stack(views)
    a = E.compute_p3p_ransac_batch(ctx, corners, target, focal, pp, W, H, want_debug=True)
    b = E.compute_p3p_ransac_batch(ctx, corners, target, focal, pp, W, H, want_debug=True)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)
    worst = 0.0
    for k in range(0, 2000, 50):
        ok0, pose0, inl0, rt0 = PR.compute_p3p_ransac_cv2(views[k], target, focal, pp, W, H)
        assert bool(a[0][k]) == ok0 and np.array_equal(np.nonzero(a[3][k])[0], inl0)
        if ok0:
            worst = max(worst, float(np.max(np.abs(a[1][k] - pose0))))
    assert worst <= 1e-6, worst


def test_capacity_error_is_reported_not_fatal(ctx):
    """More points per view than the kernel's shared-memory working set: EZC_ERR_CAPACITY, and the context stays usable."""
    from ezcalib_b200 import ezcalibrator as E
    from ezcalib_b200._lib import EzcError
    n_pts = 4000
    tgt = np.concatenate([np.random.default_rng(0).uniform(-1, 1, (n_pts, 2)), np.ones((n_pts, 1))], 1)
    corners = np.full((2, n_pts, 2), 100.0, np.float32)
    with pytest.raises(EzcError) as ei:
        E.compute_p3p_ransac_batch(ctx, corners, tgt, np.array([900.0]), np.array([640.0, 512.0]), 1280, 1024)
    assert "status 4" in str(ei.value) and "shared-memory" in str(ei.value)
    small = synth.chessboard_target(7, 10, 0.05)
    views, focal, pp, W, H = _views(small, 4, seed=3)
    ok, _, _ = E.compute_p3p_ransac_batch(ctx, np.stack(views), small, focal, pp, W, H)
    assert ok.all()

"""GPU: BASELINE.json's configs 1-3 at FULL resolution against the reference's own building blocks, view by view (the
maximum corner error of every view is collected and the worst / median are printed with -s):
  cfg1  50 views 1280x1024, 9x6 chessboard, rad-tan 5        vs cv2.findChessboardCorners + cv2.cornerSubPix
  cfg2  32 stereo pairs 1920x1080, Kannala-Brandt 4          vs the same
  cfg3  32 views 2448x2048, 6x6 AprilGrid, rad-tan 8         vs the reference's extractTags objects (oracle/_ref) + cv2.cornerSubPix
Contract for OpenCV: its generic C++ code path (cv2.ipp.setUseIPP(False)).  The IPP build of getRectSubPix rounds
differently (weak corners of blurred views move by up to ~0.015 px); the error against cv2 WITH IPP is printed too and held
to 0.05 px -- the kernels reproduce the portable code bit for bit, which is what a from-source OpenCV (the reference's
build.sh compiles its own) runs on machines without IPP."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 0.01   # px, BASELINE.json north_star


def _render(ctx, kind, W, H, model, f, cx, cy, n, key, seed, margin):
    from ezcalib_b200 import detect, synth
    mid = synth.MODEL_ID[model]
    d = synth.GT_DIST[model]
    tgt = synth.chessboard_target(7, 10, 0.05) if kind == "chessboard" else synth.aprilgrid_target(6, 6, 0.06, 0.3)
    rng = np.random.Generator(np.random.Philox(key=key))
    poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=margin, tilt_sigma=0.3)
    cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
    if kind == "chessboard":
        return detect.render_targets(ctx, "chessboard", 7, 10, 0.05, 0.0, mid, W, H, cams, poses, seed=seed)
    return detect.render_targets(ctx, "aprilgrid", 6, 6, 0.06, 0.3, mid, W, H, cams, poses, seed=seed)


def _chessboard_vs_cv2(ctx, label, W, H, model, f, cx, cy, n, key, seed):
    import cv2
    from ezcalib_b200 import detect
    imgs = _render(ctx, "chessboard", W, H, model, f, cx, cy, n, key, seed, 90)
    found, corners = detect.detect_chessboard(ctx, imgs, 7, 10)
    host = imgs.download()
    imgs.close()
    flags = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)
    errs = {False: [], True: []}
    for ipp in (False, True):
        cv2.ipp.setUseIPP(ipp)
        for i in range(n):
            ok, c = cv2.findChessboardCorners(host[i], (9, 6), flags=flags)
            assert bool(ok) == bool(found[i]), (label, i, ipp)
            if ok:
                c = cv2.cornerSubPix(host[i], c, (5, 5), (-1, -1), crit).reshape(-1, 2)
                errs[ipp].append(float(np.abs(c - corners[i]).max()))
    cv2.ipp.setUseIPP(False)
    e0, e1 = np.array(errs[False]), np.array(errs[True])
    print(f"{label}: {int(found.sum())}/{n} found; per-view max corner error vs cv2 (no IPP): worst {e0.max():.2e} median {np.median(e0):.2e} px; "
          f"vs cv2 with IPP: worst {e1.max():.2e} median {np.median(e1):.2e} px")
    assert len(e0) >= n - 3 and e0.max() <= TOL and e1.max() <= 0.05


def test_cfg1_chessboard_mono_50_views(ctx):
    pytest.importorskip("cv2")
    _chessboard_vs_cv2(ctx, "cfg1", 1280, 1024, "radtan5", 900.0, 645.0, 505.0, 50, key=41, seed=11)


def test_cfg2_chessboard_stereo_32_pairs(ctx):
    pytest.importorskip("cv2")
    _chessboard_vs_cv2(ctx, "cfg2 left", 1920, 1080, "kannalabrandt", 900.0, 965.0, 533.0, 32, key=42, seed=12)
    _chessboard_vs_cv2(ctx, "cfg2 right", 1920, 1080, "kannalabrandt", 905.0, 950.0, 545.0, 32, key=43, seed=13)


def test_cfg3_aprilgrid_32_views(ctx):
    from ezcalib_b200 import detect
    from oracle import apriltag_oracle as A
    n = 32
    imgs = _render(ctx, "aprilgrid", 2448, 2048, "radtan8", 2000.0, 1227.0, 1020.0, n, key=44, seed=14, margin=30)
    succ, corners, ndet = detect.detect_aprilgrid(ctx, imgs, 6, 6)
    host = imgs.download()
    imgs.close()
    errs = []
    for i in range(n):
        ok, c, nd, _ = A.detect_target(host[i], 6, 6)
        assert bool(ok) == bool(succ[i]) and int(nd) == int(ndet[i]), i
        assert np.array_equal(c[:, 0] < 0, corners[i][:, 0] < 0), i          # same tag ids / corner slots
        errs.append(float(np.abs(c - corners[i]).max()))
    e = np.array(errs)
    print(f"cfg3: {int(succ.sum())}/{n} views, mean {ndet.mean():.1f} tags; per-view max corner error vs the reference: worst {e.max():.2e} median {np.median(e):.2e} px")
    assert succ.sum() >= n - 2 and e.max() <= TOL

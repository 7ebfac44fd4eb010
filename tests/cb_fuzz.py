"""Shared by tests/test_chessboard_fuzz_gpu.py and tools/chessboard_fuzz.py: random chessboard views (pose, blur, noise,
occluding discs, illumination gradient) through the GPU detector and through cv2 with the reference's flags."""
from __future__ import annotations

import numpy as np


def make_views(ctx, seed: int, n: int, W: int = 640, H: int = 480):
    """n degraded 9x6 chessboard views: rendered on the GPU (random pose, sensor noise), degraded on the host."""
    import cv2

    from ezcalib_b200 import detect, synth
    rng = np.random.default_rng(seed)
    model = "radtan5"
    mid = synth.MODEL_ID[model]
    d = np.array(synth.GT_DIST[model]) * rng.uniform(0.0, 1.2)
    f = rng.uniform(0.9, 1.6) * 480.0
    cx, cy = W / 2 + rng.normal(0, 6), H / 2 + rng.normal(0, 6)
    tgt = synth.chessboard_target(7, 10, 0.05)
    prng = np.random.Generator(np.random.Philox(key=seed))
    poses = synth.sample_poses(prng, n, tgt, model, f, cx, cy, d, W, H, margin=float(rng.uniform(12, 60)), tilt_sigma=float(rng.uniform(0.1, 0.5)),
                               roll_range=float(rng.uniform(0.1, 3.1)))
    cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
    imgs = detect.render_targets(ctx, "chessboard", 7, 10, 0.05, 0.0, mid, W, H, cams, poses, noise_sigma=0.0, seed=seed)
    host = imgs.download().astype(np.float32)
    imgs.close()
    out = np.empty((n, H, W), np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(n):
        im = host[i]
        sig = rng.choice([0.0, 0.0, 0.6, 1.0, 1.6, 2.4])
        if sig > 0:
            im = cv2.GaussianBlur(im, (0, 0), float(sig))
        if rng.random() < 0.4:   # illumination gradient + contrast change
            g = rng.uniform(-0.5, 0.5, 2)
            im = im * (1.0 + g[0] * (xx / W - 0.5) + g[1] * (yy / H - 0.5)) * rng.uniform(0.5, 1.0) + rng.uniform(0, 40)
        for _ in range(int(rng.choice([0, 0, 0, 1, 2, 3]))):   # occluding discs
            cxo, cyo, r = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(6, 45)
            im = np.where((xx - cxo) ** 2 + (yy - cyo) ** 2 < r * r, float(rng.choice([20, 100, 128, 160, 235])), im)
        ns = rng.choice([0.0, 2.0, 2.0, 5.0, 10.0])
        if ns > 0:
            im = im + rng.normal(0, ns, im.shape)
        out[i] = np.clip(np.rint(im), 0, 255).astype(np.uint8)
    return out


def compare(ctx, views: np.ndarray):
    """Returns (n_found_cv2, flag mismatches [indices], per-view max corner error of the views both found)."""
    import cv2

    from ezcalib_b200 import detect
    cv2.ipp.setUseIPP(False)
    imgs = detect.Images(ctx, views)
    found, corners = detect.detect_chessboard(ctx, imgs, 7, 10)
    imgs.close()
    flags = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)
    bad, errs, n_ok = [], [], 0
    for i in range(len(views)):
        ok, c = cv2.findChessboardCorners(views[i], (9, 6), flags=flags)
        if bool(ok) != bool(found[i]):
            bad.append(i)
            continue
        if ok:
            n_ok += 1
            c = cv2.cornerSubPix(views[i], c, (5, 5), (-1, -1), crit).reshape(-1, 2)
            errs.append(float(np.abs(c - corners[i]).max()))
    return n_ok, bad, np.array(errs)

"""Shared by tests/test_aprilgrid_fuzz_gpu.py and tools/aprilgrid_fuzz.py: random 6x6 AprilGrid views (pose, blur, noise,
occluding discs, illumination gradient) through the GPU detector and through the reference's own detector (oracle/_ref,
compiled unmodified from thirdparty/ethz_apriltag2) + cv2.cornerSubPix."""
from __future__ import annotations

import numpy as np

ROWS, COLS = 6, 6


def make_views(ctx, seed: int, n: int, W: int = 1024, H: int = 768):
    import cv2

    from ezcalib_b200 import detect, synth
    rng = np.random.default_rng(seed)
    model = "radtan8"
    mid = synth.MODEL_ID[model]
    d = np.array(synth.GT_DIST[model]) * rng.uniform(0.0, 1.2)
    f = rng.uniform(0.9, 1.5) * 760.0
    cx, cy = W / 2 + rng.normal(0, 6), H / 2 + rng.normal(0, 6)
    tgt = synth.aprilgrid_target(ROWS, COLS, 0.06, 0.3)
    prng = np.random.Generator(np.random.Philox(key=seed))
    poses = synth.sample_poses(prng, n, tgt, model, f, cx, cy, d, W, H, margin=float(rng.uniform(10, 50)), tilt_sigma=float(rng.uniform(0.1, 0.5)),
                               roll_range=float(rng.uniform(0.1, 3.1)))
    cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
    imgs = detect.render_targets(ctx, "aprilgrid", ROWS, COLS, 0.06, 0.3, mid, W, H, cams, poses, noise_sigma=0.0, seed=seed)
    host = imgs.download().astype(np.float32)
    imgs.close()
    out = np.empty((n, H, W), np.uint8)
    yy, xx = np.mgrid[0:H, 0:W]
    for i in range(n):
        im = host[i]
        sig = rng.choice([0.0, 0.0, 0.6, 1.0, 1.5])
        if sig > 0:
            im = cv2.GaussianBlur(im, (0, 0), float(sig))
        if rng.random() < 0.4:   # illumination gradient + contrast change
            g = rng.uniform(-0.5, 0.5, 2)
            im = im * (1.0 + g[0] * (xx / W - 0.5) + g[1] * (yy / H - 0.5)) * rng.uniform(0.5, 1.0) + rng.uniform(0, 40)
        for _ in range(int(rng.choice([0, 0, 1, 2, 4, 8]))):   # occluding discs (knock out tags, cut borders)
            cxo, cyo, r = rng.uniform(0, W), rng.uniform(0, H), rng.uniform(6, 60)
            im = np.where((xx - cxo) ** 2 + (yy - cyo) ** 2 < r * r, float(rng.choice([20, 100, 128, 160, 235])), im)
        ns = rng.choice([0.0, 2.0, 2.0, 5.0, 10.0])
        if ns > 0:
            im = im + rng.normal(0, ns, im.shape)
        out[i] = np.clip(np.rint(im), 0, 255).astype(np.uint8)
    return out


def compare(ctx, views: np.ndarray):
    """Returns (views with a detection in the reference, [indices whose success flag / detection count / filled slots
    differ], per-view max corner difference of the others)."""
    import cv2

    from ezcalib_b200 import detect
    from oracle import apriltag_oracle as A
    cv2.ipp.setUseIPP(False)   # the OpenCV contract of DESIGN.md section 4b: the portable code path of cornerSubPix
    imgs = detect.Images(ctx, views)
    succ, corners, ndet = detect.detect_aprilgrid(ctx, imgs, ROWS, COLS)
    imgs.close()
    bad, errs, n_det = [], [], 0
    for i in range(len(views)):
        ok_r, c_r, n_r, _ = A.detect_target(views[i], ROWS, COLS)
        n_det += int(n_r > 0)
        if bool(ok_r) != bool(succ[i]) or int(n_r) != int(ndet[i]) or not np.array_equal(corners[i][:, 0] < 0, c_r[:, 0] < 0):
            bad.append(i)
            continue
        errs.append(float(np.abs(corners[i] - c_r).max()) if n_r > 0 else 0.0)
    return n_det, bad, np.array(errs)

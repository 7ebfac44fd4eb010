"""Chessboard fixtures + what the REAL OpenCV (cv2 4.13.0 of the build container) returns for them with the reference's
exact calls (chessboard_detector.hpp:30-40).  Run from the repo root:  python tests/golden/make_chessboard_golden.py"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ezcalib_b200 import synth  # noqa: E402

FLAGS = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
CRIT = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)


def flat_board(sq_cols=10, sq_rows=7, s=60, margin=60):
    im = np.full((sq_rows * s + 2 * margin, sq_cols * s + 2 * margin), 235, np.uint8)
    for i in range(sq_rows):
        for j in range(sq_cols):
            if (i + j) % 2 == 1:
                im[margin + i * s:margin + (i + 1) * s, margin + j * s:margin + (j + 1) * s] = 20
    return im


def warped(rng, W, H, blur, noise, grad, gain, shift_sigma=0.08):
    B = flat_board()
    bh, bw = B.shape
    scale = rng.uniform(0.45, 0.85) * min(W / bw, H / bh)
    ang = rng.uniform(-0.5, 0.5)
    c, s_ = np.cos(ang) * scale, np.sin(ang) * scale
    Hm = np.array([[c, -s_, 0], [s_, c, 0], [0, 0, 1.0]])
    Hm[2, 0] = rng.normal(0, 2e-4); Hm[2, 1] = rng.normal(0, 2e-4)
    ctr = Hm @ np.array([bw / 2, bh / 2, 1.0]); ctr = ctr[:2] / ctr[2]
    shift = np.array([W / 2, H / 2]) - ctr + rng.normal(0, shift_sigma, 2) * [W, H]
    T = np.array([[1, 0, shift[0]], [0, 1, shift[1]], [0, 0, 1.0]])
    img = cv2.warpPerspective(B, T @ Hm, (W, H), flags=cv2.INTER_LINEAR, borderValue=200)
    img = cv2.GaussianBlur(img, (0, 0), blur)
    gx = np.linspace(-1, 1, W)[None, :] * grad
    return np.clip(img.astype(np.float32) * gain + gx + rng.normal(0, noise, img.shape), 0, 255).astype(np.uint8)


def main():
    rng = np.random.Generator(np.random.Philox(key=synth.MASTER_SEED + 21))
    out, names = {}, []
    W, H = 640, 480
    tgt = synth.chessboard_target(7, 10, 0.05)
    pat = synth.chessboard_pattern(7, 10, 0.05)
    for k, (model, f) in enumerate((("radtan5", 450.0), ("kannalabrandt", 300.0), ("radtan8", 450.0))):
        d = synth.GT_DIST[model]
        pose = synth.sample_poses(rng, 1, tgt, model, f, 322.0, 238.0, d, W, H, margin=45, tilt_sigma=0.35)[0]
        out[f"render{k}_img"] = synth.render_view(pat, model, f, f, 322.0, 238.0, d, pose, W, H, rng)
        names.append(f"render{k}")
    # harder views: blur / noise / illumination gradient (several dilations or the adaptive fallback are needed)
    r2 = np.random.default_rng(5)
    hard = [(1.8, 5.0, 35.0, 0.6), (1.2, 3.0, 38.0, 0.55), (0.8, 1.0, 0.0, 1.0), (1.9, 4.0, 25.0, 0.7), (1.5, 6.0, 30.0, 0.5)]
    for k, (blur, noise, grad, gain) in enumerate(hard):
        out[f"warp{k}_img"] = warped(r2, W, H, blur, noise, grad, gain)
        names.append(f"warp{k}")
    out["cut_img"] = warped(r2, W, H, 1.0, 2.0, 0.0, 1.0, shift_sigma=0.45)   # board partly out of frame
    out["blank_img"] = np.full((H, W), 128, np.uint8)
    out["noise_img"] = np.random.default_rng(1).integers(0, 256, size=(240, 320)).astype(np.uint8)
    names += ["cut", "blank", "noise"]
    for n in names:
        img = out[n + "_img"]
        ok, c = cv2.findChessboardCorners(img, (9, 6), flags=FLAGS)
        out[n + "_found"] = np.array(bool(ok))
        if ok:
            out[n + "_raw"] = c.reshape(-1, 2)
            out[n + "_final"] = cv2.cornerSubPix(img, c.copy(), (5, 5), (-1, -1), CRIT).reshape(-1, 2)
        print(n, img.shape, "found", ok)
    out["names"] = np.array(names)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "chessboard_golden.npz"), **out)


if __name__ == "__main__":
    main()

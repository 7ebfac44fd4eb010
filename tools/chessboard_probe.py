"""Probe: GPU-rendered chessboard batch (cfg1 / cfg2 geometry) -> GPU detection timing and parity against cv2."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2
from ezcalib_b200 import synth, detect, lm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 50
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (1280, 1024)
model = sys.argv[4] if len(sys.argv) > 4 else "radtan5"
ncheck = int(sys.argv[5]) if len(sys.argv) > 5 else n
ctx = lm.Context(0)
mid = synth.MODEL_ID[model]; d = synth.GT_DIST[model]
f = (600.0 if model == "kannalabrandt" else 900.0) * W / 1280; cx, cy = W / 2 + 5, H / 2 - 7
tgt = synth.chessboard_target(7, 10, 0.05)
rng = np.random.Generator(np.random.Philox(key=9))
poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=60, tilt_sigma=0.35)
cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
imgs = detect.render_targets(ctx, "chessboard", 7, 10, 0.05, 0.0, mid, W, H, cams, poses)
for rep in range(3):
    t = time.time(); found, corners = detect.detect_chessboard(ctx, imgs, 7, 10); dt = time.time() - t
    print(f"gpu detect {n} images {W}x{H}: {dt:.3f}s -> {n/dt:.1f} img/s; found {found.sum()}/{n}")
host = imgs.download(0, ncheck)
FLAGS = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
CRIT = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)
t0 = time.time(); diffs = []; flag_mis = 0
for i in range(ncheck):
    ok, c = cv2.findChessboardCorners(host[i], (9, 6), flags=FLAGS)
    if bool(ok) != bool(found[i]): flag_mis += 1; continue
    if ok:
        ref = cv2.cornerSubPix(host[i], c, (5, 5), (-1, -1), CRIT).reshape(-1, 2)
        diffs.append(float(np.abs(ref - corners[i]).max()))
dt = time.time() - t0
diffs = np.array(diffs)
print(f"cv2 chain: {ncheck} images in {dt:.2f}s -> {ncheck/dt:.1f} img/s ({cv2.getNumThreads()} threads); flag mismatches {flag_mis}; "
      f"corner diff: exact {np.mean(diffs == 0):.2f}, <=0.01 {np.mean(diffs <= 0.01):.3f}, max {diffs.max() if len(diffs) else None}")
gt = synth.project(model, f, f, cx, cy, d, poses[0], tgt)
if found[0]: print("gpu vs GT image 0: max err", np.linalg.norm(corners[0] - gt, axis=1).max())

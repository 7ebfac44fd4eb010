"""Probe: noisy / occluded chessboard views (fallback phase, long noise contours) -> GPU vs cv2 parity + timing."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cv2
from ezcalib_b200 import synth, detect, lm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 24
sigma = float(sys.argv[2]) if len(sys.argv) > 2 else 10.0
W, H, model = 960, 720, "radtan5"
ctx = lm.Context(0)
mid = synth.MODEL_ID[model]; d = synth.GT_DIST[model]
f = 700.0; cx, cy = W / 2 + 5, H / 2 - 7
tgt = synth.chessboard_target(7, 10, 0.05)
rng = np.random.Generator(np.random.Philox(key=21))
poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=50, tilt_sigma=0.4)
cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
imgs = detect.render_targets(ctx, "chessboard", 7, 10, 0.05, 0.0, mid, W, H, cams, poses, noise_sigma=sigma)
host = imgs.download()
# occlude a corner of the board on every 4th view (=> the search runs all attempts and fails, or recovers late)
for i in range(0, n, 4):
    gt = synth.project(model, f, f, cx, cy, d, poses[i], tgt)
    cv2.circle(host[i], (int(gt[0, 0]), int(gt[0, 1])), 30, 128, -1)
imgs.close()
imgs = detect.Images(ctx, host)
detect.detect_chessboard(ctx, imgs, 7, 10)
t = time.time(); found, corners = detect.detect_chessboard(ctx, imgs, 7, 10); dt = time.time() - t
print(f"gpu: {n} views in {dt:.3f}s, found {int(found.sum())}")
FLAGS = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
CRIT = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)
cv2.ipp.setUseIPP(False)   # OpenCV's generic code path, which the kernels restate (IPP's sampler moves weak corners)
t0 = time.time(); diffs = []; mis = []
for i in range(n):
    ok, c = cv2.findChessboardCorners(host[i], (9, 6), flags=FLAGS)
    if bool(ok) != bool(found[i]): mis.append(i); continue
    if ok:
        ref = cv2.cornerSubPix(host[i], c, (5, 5), (-1, -1), CRIT).reshape(-1, 2)
        diffs.append(float(np.abs(ref - corners[i]).max()))
if mis:
    os.makedirs("gpurun_out", exist_ok=True)
    np.save(f"gpurun_out/cb_mismatch_s{int(sigma)}.npy", host[mis])
    print("gpu found flags of mismatches:", [bool(found[i]) for i in mis])
print(f"cv2: {time.time()-t0:.3f}s; flag mismatches {mis}; corner diffs max {max(diffs) if diffs else None}; >0.01: {sum(x > 0.01 for x in diffs)} of {len(diffs)}")

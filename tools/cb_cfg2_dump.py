"""cfg2-style chessboard batch of the bench: compare found flags with cv2 on the GPU's not-found views and dump mismatches."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, cv2
import bench
from ezcalib_b200 import detect, synth, lm
ctx = lm.Context(0)
C = bench.CB_CONFIGS["cfg2"]
n, W, H, model = 400, C["W"], C["H"], C["model"]
mid = synth.MODEL_ID[model]; d = synth.GT_DIST[model]
f, cx, cy = C["f"], C["cx"], C["cy"]
tgt = synth.chessboard_target(7, 10, 0.05)
rng = np.random.Generator(np.random.Philox(key=synth.MASTER_SEED + 2000 + 0))
poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=C["margin"], tilt_sigma=0.3)
cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
imgs = detect.render_targets(ctx, "chessboard", 7, 10, 0.05, 0.0, mid, W, H, cams, poses, seed=synth.MASTER_SEED + 7 + 0)
t = time.time(); found, corners = detect.detect_chessboard(ctx, imgs, 7, 10); print("gpu", time.time() - t, "found", found.sum())
bad = np.nonzero(~found)[0]
host = imgs.download()
FLAGS = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
mis = []
for i in bad:
    t = time.time(); ok, c = cv2.findChessboardCorners(host[i], (9, 6), flags=FLAGS); dt = time.time() - t
    ok2, _ = cv2.findChessboardCorners(host[i], (9, 6), flags=FLAGS - cv2.CALIB_CB_FAST_CHECK)
    print("view", i, "gpu not found; cv2 found", ok, "(%.0f ms)" % (dt * 1e3), "without FAST_CHECK", ok2)
    if ok: mis.append(i)
os.makedirs("gpurun_out", exist_ok=True)
if mis: np.save("gpurun_out/cb_cfg2_mismatch.npy", host[mis])

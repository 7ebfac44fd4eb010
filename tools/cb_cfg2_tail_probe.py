"""The cfg2-style bench batch: which views are not found, and what each of them costs alone (with EZC_CB_TIMING=1 the
quad-graph phase timers of cb_process_kernel are printed per call)."""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from ezcalib_b200 import detect, lm, synth

C = bench.CB_CONFIGS["cfg2"]
n = 400
W, H, model = C["W"], C["H"], C["model"]
mid, d = synth.MODEL_ID[model], synth.GT_DIST[model]
f, cx, cy = C["f"], C["cx"], C["cy"]
ctx = lm.Context(0)
tgt = synth.chessboard_target(7, 10, 0.05)
rng = np.random.Generator(np.random.Philox(key=synth.MASTER_SEED + 2000))
poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=C["margin"], tilt_sigma=0.3)
cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
imgs = detect.render_targets(ctx, "chessboard", 7, 10, 0.05, 0.0, mid, W, H, cams, poses, seed=synth.MASTER_SEED + 7)
out = detect.detect_chessboard(ctx, imgs, 7, 10)
found = np.asarray(out[0]).astype(bool)
print("not found:", np.nonzero(~found)[0].tolist())
if True:
    for i in np.nonzero(~found)[0]:
        one = detect.Images(ctx, imgs.download(int(i), 1))
        detect.detect_chessboard(ctx, one, 7, 10)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        detect.detect_chessboard(ctx, one, 7, 10)
        torch.cuda.synchronize()
        print("view", int(i), "alone: %.1f ms" % (1e3 * (time.perf_counter() - t0)), flush=True)
        one.close()

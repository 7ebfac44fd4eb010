"""Dump the views of the bench's chessboard batch that are not found (for offline analysis of the fast-check gate)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ezcalib_b200 import detect, synth, lm
ctx = lm.Context(0)
n, W, H, model = 400, 1280, 1024, "radtan5"
mid = synth.MODEL_ID[model]; d = synth.GT_DIST[model]
f, cx, cy = 900.0, 645.0, 505.0
tgt = synth.chessboard_target(7, 10, 0.05)
rng = np.random.Generator(np.random.Philox(key=synth.MASTER_SEED + 2000 + 0))
poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=90, tilt_sigma=0.3)
cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
imgs = detect.render_targets(ctx, "chessboard", 7, 10, 0.05, 0.0, mid, W, H, cams, poses, seed=synth.MASTER_SEED + 7 + 0)
found, corners = detect.detect_chessboard(ctx, imgs, 7, 10)
bad = np.nonzero(~found)[0]
print("not found:", bad)
host = imgs.download()
os.makedirs("gpurun_out", exist_ok=True)
np.save("gpurun_out/cb_bench_notfound.npy", host[bad])
one = detect.Images(ctx, host[bad])
os.environ["EZC_FC_DEBUG"] = "1"
print(detect.detect_chessboard(ctx, one, 7, 10)[0])
import time, torch
os.environ.pop("EZC_FC_DEBUG", None)
for fc in (True, False, True):
    torch.cuda.synchronize(); t = time.perf_counter()
    f2, c2 = detect.find_chessboard_corners(ctx, imgs, (9, 6), True, fast_check=fc)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"400 views fast_check={fc}: {dt*1e3:.1f} ms, found {int(f2.sum())}, launches so far {ctx.kernel_launches()}")
good = detect.Images(ctx, host[found])
for fc in (True, False):
    torch.cuda.synchronize(); t = time.perf_counter()
    f2, c2 = detect.find_chessboard_corners(ctx, good, (9, 6), True, fast_check=fc)
    torch.cuda.synchronize(); dt = time.perf_counter() - t
    print(f"398 found views fast_check={fc}: {dt*1e3:.1f} ms")

import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ezcalib_b200 import synth, detect, lm
n, W, H, model = 50, 1280, 1024, "radtan5"
ctx = lm.Context(0)
mid = synth.MODEL_ID[model]; d = synth.GT_DIST[model]
f = 900.0; cx, cy = W / 2 + 5, H / 2 - 7
tgt = synth.chessboard_target(7, 10, 0.05)
rng = np.random.Generator(np.random.Philox(key=9))
poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=60, tilt_sigma=0.35)
cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
imgs = detect.render_targets(ctx, "chessboard", 7, 10, 0.05, 0.0, mid, W, H, cams, poses)
found, corners = detect.detect_chessboard(ctx, imgs, 7, 10)
bad = int(np.nonzero(~found)[0][0])
host = imgs.download(bad, 1)
np.save("gpurun_out/cb_fail_img.npy", host[0])
one = detect.Images(ctx, host)
t = time.time(); f1, c1 = detect.detect_chessboard(ctx, one, 7, 10); print("failing image alone:", time.time() - t, f1)
good = detect.Images(ctx, imgs.download(0, 49) if bad == 49 else np.delete(imgs.download(0, 50), bad, axis=0))
for _ in range(3):
    t = time.time(); f2, c2 = detect.detect_chessboard(ctx, good, 7, 10); dt = time.time() - t
    print(f"49 found images: {dt:.4f}s -> {49/dt:.1f} img/s", f2.sum())

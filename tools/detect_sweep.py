"""Throughput of AprilGrid detection vs images in flight (sub-batch size)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from ezcalib_b200 import synth, detect, lm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 280
W, H = 2448, 2048
ctx = lm.Context(0)
model = "radtan8"; mid = synth.MODEL_ID[model]; d = synth.GT_DIST[model]
f = 2000.0; cx, cy = W / 2 + 3, H / 2 - 4
tgt = synth.aprilgrid_target(6, 6, 0.06, 0.3)
rng = np.random.Generator(np.random.Philox(key=5))
poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=30, tilt_sigma=0.3)
cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
imgs = detect.render_targets(ctx, "aprilgrid", 6, 6, 0.06, 0.3, mid, W, H, cams, poses)
ref = None
for sub in [int(a) for a in sys.argv[2:]] or [8, 16, 35, 70, 140]:
    out = detect.detect_aprilgrid(ctx, imgs, 6, 6, max_images_in_flight=sub)
    torch.cuda.synchronize()
    t = time.perf_counter()
    out = detect.detect_aprilgrid(ctx, imgs, 6, 6, max_images_in_flight=sub)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t
    if ref is None: ref = out
    same = all(np.array_equal(a, b) for a, b in zip(ref, out))
    print(f"in flight {sub:4d}: {n / dt:8.1f} img/s  same={same}", flush=True)

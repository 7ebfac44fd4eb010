"""Probe: GPU-rendered AprilGrid batch -> GPU detection timing, and parity against the reference detector on a sample."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from ezcalib_b200 import synth, detect, lm
from oracle import apriltag_oracle as A

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
W, H = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (2448, 2048)
ctx = lm.Context(0)
model = "radtan8"; mid = synth.MODEL_ID[model]; d = synth.GT_DIST[model]
f = 2000.0 * W / 2448; cx, cy = W / 2 + 3, H / 2 - 4
tgt = synth.aprilgrid_target(6, 6, 0.06, 0.3)
rng = np.random.Generator(np.random.Philox(key=5))
poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=30, tilt_sigma=0.3)
cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
t = time.time(); imgs = detect.render_targets(ctx, "aprilgrid", 6, 6, 0.06, 0.3, mid, W, H, cams, poses); print("render", time.time() - t)
for rep in range(3):
    t = time.time(); succ, corners, ndet = detect.detect_aprilgrid(ctx, imgs, 6, 6); dt = time.time() - t
    print(f"gpu detect {n} images: {dt:.3f}s -> {n/dt:.1f} img/s; success {succ.sum()}/{n}; mean ndet {ndet.mean():.1f}")
host = imgs.download(0, min(4, n))
worst = 0
for i in range(host.shape[0]):
    t = time.time(); ok, c, nd, det = A.detect_target(host[i], 6, 6); dt = time.time() - t
    same = np.array_equal(c[:, 0] < 0, corners[i][:, 0] < 0)
    err = np.abs(c - corners[i]).max()
    worst = max(worst, err)
    print(f"ref image {i}: {dt:.3f}s ok={ok} n={nd} | gpu ok={succ[i]} n={ndet[i]} slots_equal={same} max|dcorner|={err:.2e}")
gt = synth.project(model, f, f, cx, cy, d, poses[0], tgt)
m = corners[0][:, 0] >= 0
print("gpu vs GT max err", np.linalg.norm(corners[0][m] - gt[m], axis=1).max(), "worst vs ref", worst)

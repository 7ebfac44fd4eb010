"""The bench's chessboard batch (cfg1 / cfg2, see bench.CB_CONFIGS) for an ncu launch list: warm-up call, then ONE
profiled call bracketed by cudaProfilerStart/Stop (run under `ncu --profile-from-start off`).
usage: cb_bench_launches.py [cfg1|cfg2] [n_views]"""
import os
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import bench
from ezcalib_b200 import detect, lm, synth

cfg = sys.argv[1] if len(sys.argv) > 1 else "cfg1"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 400
C = bench.CB_CONFIGS[cfg]
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
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
t0 = time.perf_counter()
out = detect.detect_chessboard(ctx, imgs, 7, 10)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
torch.cuda.cudart().cudaProfilerStop()
print(cfg, n, "found", int(np.asarray(out[1]).sum()), "%.1f ms" % (dt * 1e3), "%.0f img/s" % (n / dt))

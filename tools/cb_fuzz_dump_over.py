"""Dump the fuzz views whose corners differ from cv2 by more than 0.01 px: `python tools/cb_fuzz_dump_over.py <seed> [<seed> ...]`
writes gpurun_out/cb_fuzz_over_<seed>_<i>.npy (for the CPU harness in tests/cb_harness)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cv2
import cb_fuzz
from ezcalib_b200 import lm, detect
ctx = lm.Context(0)
cv2.ipp.setUseIPP(False)
flags = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)
os.makedirs("gpurun_out", exist_ok=True)
for seed in [int(a) for a in sys.argv[1:]]:
    views = cb_fuzz.make_views(ctx, seed, 64)
    imgs = detect.Images(ctx, views)
    found, corners = detect.detect_chessboard(ctx, imgs, 7, 10)
    imgs.close()
    for i in range(len(views)):
        ok, c = cv2.findChessboardCorners(views[i], (9, 6), flags=flags)
        if ok and found[i]:
            c = cv2.cornerSubPix(views[i], c, (5, 5), (-1, -1), crit).reshape(-1, 2)
            e = float(np.abs(c - corners[i]).max())
            if e > 0.01:
                np.save(f"gpurun_out/cb_fuzz_over_{seed}_{i}.npy", views[i])
                np.save(f"gpurun_out/cb_fuzz_over_{seed}_{i}_gpu.npy", corners[i])
                print(seed, i, e, flush=True)

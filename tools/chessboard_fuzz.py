"""Fuzz the chessboard detector against cv2 (tests/cb_fuzz.py): `python tools/chessboard_fuzz.py <n_batches> [first_seed]`.
Prints per batch the found counts, flag mismatches and the worst corner error; mismatching views are saved to
gpurun_out/cb_fuzz_mismatch_<seed>_<i>.npy for tests/cb_harness."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import cb_fuzz
from ezcalib_b200 import lm
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 10
s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 1000
ctx = lm.Context(0)
tot = tot_ok = tot_bad = 0
worst = 0.0
over = 0
for s in range(s0, s0 + nb):
    views = cb_fuzz.make_views(ctx, s, 64)
    n_ok, bad, errs = cb_fuzz.compare(ctx, views)
    tot += len(views); tot_ok += n_ok; tot_bad += len(bad)
    w = float(errs.max()) if len(errs) else 0.0
    worst = max(worst, w); over += int((errs > 0.01).sum())
    print(f"seed {s}: cv2 found {n_ok}/{len(views)}, flag mismatches {bad}, worst corner error {w:.2e}, views over 0.01 px {int((errs > 0.01).sum())}", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    for i in bad[:3]:
        np.save(f"gpurun_out/cb_fuzz_mismatch_{s}_{i}.npy", views[i])
print(f"TOTAL {tot} views, cv2 found {tot_ok}, flag mismatches {tot_bad}, worst corner error {worst:.2e}, views over 0.01 px {over}")

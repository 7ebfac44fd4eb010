"""Fuzz the AprilGrid detector against the reference's own code (tests/at_fuzz.py): `python tools/aprilgrid_fuzz.py <n_batches> [first_seed]`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import at_fuzz
from ezcalib_b200 import lm
nb = int(sys.argv[1]) if len(sys.argv) > 1 else 4
s0 = int(sys.argv[2]) if len(sys.argv) > 2 else 7000
ctx = lm.Context(0)
tot = tot_det = tot_bad = 0
worst = 0.0
for s in range(s0, s0 + nb):
    views = at_fuzz.make_views(ctx, s, 32)
    n_det, bad, errs = at_fuzz.compare(ctx, views)
    tot += len(views); tot_det += n_det; tot_bad += len(bad)
    w = float(errs.max()) if len(errs) else 0.0
    worst = max(worst, w)
    print(f"seed {s}: reference detects tags in {n_det}/{len(views)}, structural mismatches {bad}, worst corner difference {w:.2e}", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    for i in bad[:3]:
        np.save(f"gpurun_out/at_fuzz_mismatch_{s}_{i}.npy", views[i])
print(f"TOTAL {tot} views, with detections {tot_det}, structural mismatches {tot_bad}, worst corner difference {worst:.2e}")

"""Where the per-camera mono calibration of cfg4 spends its time (2000 views x 144 points): cProfile around
calibrate_camera_from_corners on synthetic detections, plus EZC_LM_TIMING output of the library."""
import os, sys, time, cProfile, pstats
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
from ezcalib_b200 import lm, synth, ezcalibrator as E
from test_p3p_cpu import _views
ctx = lm.Context(0)
tgt = synth.aprilgrid_target(6, 6, 0.06, 0.3)
v, f, pp, W, H = _views(tgt, 2000, seed=25, W=3840, H=2160, f=2800.0)
corners = np.stack(v).astype(np.float32)
found = np.ones(len(corners), bool)
names = ["%06d" % i for i in range(len(corners))]
for rep in range(2):
    t0 = time.perf_counter()
    tm = {}
    cam, processed, summ = E.calibrate_camera_from_corners(ctx, found, corners, names, tgt, "radtan8", True, 70.0, W, H, tm)
    print("rep", rep, "total %.1f ms" % (1e3 * (time.perf_counter() - t0)), {k: round(1e3 * x, 1) for k, x in tm.items()}, [s["iterations"] for s in summ])
pr = cProfile.Profile(); pr.enable()
cam, processed, summ = E.calibrate_camera_from_corners(ctx, found, corners, names, tgt, "radtan8", True, 70.0, W, H, {})
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(22)

"""Generates tests/golden/subpix_golden.npz with the REAL cv2.cornerSubPix (OpenCV 4.13.0 in the build
container) using the reference's exact arguments (chessboard_detector.hpp:39-40).
Run from the repo root:  python tests/golden/make_subpix_golden.py
"""
import os
import sys

import cv2
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ezcalib_b200 import synth  # noqa: E402

CRIT = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)


def main():
    rng = np.random.Generator(np.random.Philox(key=synth.MASTER_SEED + 7))
    W, H = 640, 480
    model, d = "radtan5", synth.GT_DIST["radtan5"]
    f, cx, cy = 450.0, 322.0, 238.0
    tgt = synth.chessboard_target(7, 10, 0.05)
    pat = synth.chessboard_pattern(7, 10, 0.05)
    out = {}
    # case 0: board well inside; case 1: board touching the image border (border path of getRectSubPix)
    for case, margin in enumerate((40.0, 2.0)):
        pose = synth.sample_poses(rng, 1, tgt, model, f, cx, cy, d, W, H, margin=margin, tilt_sigma=0.3)[0]
        if case == 1:
            # push the board so that some inner corners are within 6 px of the left/top border
            gt = synth.project(model, f, f, cx, cy, d, pose, tgt)
            pose[4] -= (gt[:, 0].min() - 3.0) * pose[6] / f * 0.0  # keep pose; shift below through cx,cy
        gt = synth.project(model, f, f, cx, cy, d, pose, tgt)
        sx = 0.0 if case == 0 else -(gt[:, 0].min() - 3.2)
        sy = 0.0 if case == 0 else -(gt[:, 1].min() - 4.1)
        img = synth.render_view(pat, model, f, f, cx + sx, cy + sy, d, pose, W, H, rng)
        gt = synth.project(model, f, f, cx + sx, cy + sy, d, pose, tgt)
        start = (gt + rng.uniform(-1.5, 1.5, size=gt.shape)).astype(np.float32)
        start[:, 0] = np.clip(start[:, 0], 0, W - 1)
        start[:, 1] = np.clip(start[:, 1], 0, H - 1)
        ref = cv2.cornerSubPix(img, start.copy().reshape(-1, 1, 2), (5, 5), (-1, -1), CRIT).reshape(-1, 2)
        ref2 = cv2.cornerSubPix(img, start.copy().reshape(-1, 1, 2), (2, 2), (-1, -1),
                                (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 15, 0.1)).reshape(-1, 2)
        out[f"img{case}"] = img
        out[f"start{case}"] = start
        out[f"ref_win5_{case}"] = ref
        out[f"ref_win2_{case}"] = ref2
        print(case, "min corner", gt.min(axis=0), "moved", np.abs(ref - start).max())
    out["cv2_version"] = np.array(cv2.__version__)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "subpix_golden.npz"), **out)


if __name__ == "__main__":
    main()

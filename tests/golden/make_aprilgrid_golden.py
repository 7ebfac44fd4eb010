"""Renders the AprilGrid fixture images and records what the REFERENCE's own detector (oracle/_ref, built from
/root/reference/thirdparty/ethz_apriltag2) + cv2.cornerSubPix return for them.
Run from the repo root:  python tests/golden/make_aprilgrid_golden.py
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from ezcalib_b200 import synth  # noqa: E402
from ezcalib_b200.tag36h11 import CODES  # noqa: E402
from oracle import apriltag_oracle as A  # noqa: E402


def main():
    rng = np.random.Generator(np.random.Philox(key=synth.MASTER_SEED + 11))
    out = {}
    cases = [
        # name, rows, cols, W, H, model, f, margin, tilt
        ("grid6x6", 6, 6, 1224, 1024, "radtan8", 1000.0, 40.0, 0.25),
        ("grid4x3_tilt", 3, 4, 800, 600, "radtan5", 620.0, 30.0, 0.5),
        ("grid6x6_kb", 6, 6, 960, 540, "kannalabrandt", 420.0, 25.0, 0.3),
    ]
    for name, rows, cols, W, H, model, f, margin, tilt in cases:
        tgt = synth.aprilgrid_target(rows, cols, 0.06, 0.3)
        d = synth.GT_DIST[model]
        cx, cy = W / 2 + 3.0, H / 2 - 4.0
        pose = synth.sample_poses(rng, 1, tgt, model, f, cx, cy, d, W, H, margin=margin, tilt_sigma=tilt)[0]
        pat = synth.aprilgrid_pattern(rows, cols, 0.06, 0.3, CODES)
        img = synth.render_view(pat, model, f, f, cx, cy, d, pose, W, H, rng)
        ok, corners, n, det = A.detect_target(img, rows, cols)
        print(name, "success", ok, "n", n, "ids", sorted(det["id"].tolist())[:8], "...")
        out[name + "_img"] = img
        out[name + "_grid"] = np.array([rows, cols])
        out[name + "_corners"] = corners
        out[name + "_n"] = np.array(n)
        out[name + "_ok"] = np.array(ok)
        out[name + "_ids"] = det["id"]
        out[name + "_p"] = det["p"]
    # board partly out of frame: shift the principal point so that half of the grid leaves the image
    name, rows, cols, W, H, model, f = "grid6x6_cut", 6, 6, 800, 600, "radtan5", 700.0
    tgt = synth.aprilgrid_target(rows, cols, 0.06, 0.3)
    d = synth.GT_DIST[model]
    pose = synth.sample_poses(rng, 1, tgt, model, f, W / 2, H / 2, d, W, H, margin=20.0, tilt_sigma=0.2)[0]
    pat = synth.aprilgrid_pattern(rows, cols, 0.06, 0.3, CODES)
    img = synth.render_view(pat, model, f, f, W / 2 + 330.0, H / 2 + 60.0, d, pose, W, H, rng)
    ok, corners, n, det = A.detect_target(img, rows, cols)
    print(name, "success", ok, "n", n)
    out[name + "_img"] = img; out[name + "_grid"] = np.array([rows, cols]); out[name + "_corners"] = corners
    out[name + "_n"] = np.array(n); out[name + "_ok"] = np.array(ok); out[name + "_ids"] = det["id"]; out[name + "_p"] = det["p"]
    # noise and blank
    for name, img in (("noise", rng.integers(0, 256, size=(240, 320)).astype(np.uint8)), ("blank", np.full((240, 320), 140, np.uint8))):
        ok, corners, n, det = A.detect_target(img, 6, 6)
        print(name, "success", ok, "n", n)
        out[name + "_img"] = img; out[name + "_grid"] = np.array([6, 6]); out[name + "_corners"] = corners
        out[name + "_n"] = np.array(n); out[name + "_ok"] = np.array(ok); out[name + "_ids"] = det["id"]; out[name + "_p"] = det["p"]
    out["names"] = np.array(["grid6x6", "grid4x3_tilt", "grid6x6_kb", "grid6x6_cut", "noise", "blank"])
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "aprilgrid_golden.npz"), **out)


if __name__ == "__main__":
    main()

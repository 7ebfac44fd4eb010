import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ezcalib_b200 import detect
from ezcalib_b200.lm import Context
from oracle import subpix_oracle as S
G = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "subpix_golden.npz"))
ctx = Context(0)
rng = np.random.default_rng(3)
img = G["img1"]
H, W = img.shape
pts = []
for _ in range(200):
    side = rng.integers(0, 4)
    t = rng.uniform(0, 1)
    if side == 0: pts.append((rng.uniform(0, 7), t * (H - 1)))
    elif side == 1: pts.append((W - 1 - rng.uniform(0, 7), t * (H - 1)))
    elif side == 2: pts.append((t * (W - 1), rng.uniform(0, 7)))
    else: pts.append((t * (W - 1), H - 1 - rng.uniform(0, 7)))
pts = np.array(pts, np.float32)
ora = S.corner_subpix(img, pts)
for rep in range(3):
    imgs = detect.Images(ctx, img)
    got = detect.corner_subpix(ctx, imgs, pts, np.zeros(len(pts), np.int32))
    imgs.close()
    eq = np.all(got == ora, axis=1)
    print(rep, "frac exact", eq.mean(), "max", np.abs(got - ora).max())
    bad = np.where(~eq)[0]
    for b in bad[:30]:
        print("  ", b, pts[b], got[b], ora[b])
    # pad with garbage allocation between reps
    junk = detect.Images(ctx, np.full((3, H, W), 255, np.uint8)); junk.close()

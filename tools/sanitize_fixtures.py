"""The small fixtures through every kernel family once -- the workload of `make sanitize` (compute-sanitizer memcheck /
racecheck / synccheck).  The merge / union-find / atomics-heavy kernels (at_ccl_union, cb_label_merge16, at_merge) are exactly
where racecheck earns its keep (SURVEY.md section 5 "Race detection")."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from ezcalib_b200 import detect, ezcalibrator as E, lm, synth

ctx = lm.Context(0)
gold = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
G = np.load(os.path.join(gold, "aprilgrid_golden.npz"))
img = np.ascontiguousarray(G["grid6x6_img"])
imgs = detect.Images(ctx, np.stack([img, img[::-1].copy()]))
succ, corners, ndet = detect.detect_aprilgrid(ctx, imgs, 6, 6)
imgs.close()
Gc = np.load(os.path.join(gold, "chessboard_golden.npz"))
cb = np.stack([Gc["render0_img"], Gc["warp1_img"], Gc["blank_img"], Gc["noise_img"]])
imgs = detect.Images(ctx, cb)
found, cc = detect.detect_chessboard(ctx, imgs, 7, 10)
imgs.close()
tgt = synth.chessboard_target(7, 10, 0.05)
H, W = cb.shape[1:]
f0, pp0 = E.setup_initial_calibration(W, H, 70.0, True)
okp, poses, ninl = E.compute_p3p_ransac_batch(ctx, cc[:2], tgt, f0, pp0, W, H)
for model, mono in (("radtan5", True), ("kannalabrandt", False)):
    P = synth.make_lm_problem(model, mono, 9, tgt, 1280, 1024, seed=1, drop_fraction=0.1)
    lm.lm_mono(ctx, synth.MODEL_ID[model], mono, P.obs, P.target, P.width, P.height, P.focal_init, P.pp_init, P.dist_init, P.poses_init)
print("fixtures ok:", int(ndet[0]), list(map(int, found)), list(map(int, ninl)))
ctx.close()

"""GPU: batched cornerSubPix kernel against the oracle restatement (tight) and the cv2 golden vectors
(north_star tolerance: sub-pixel corners within 0.01 px)."""
import os

import numpy as np
import pytest

from oracle import subpix_oracle as S

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "subpix_golden.npz"))


def test_subpix_matches_oracle_and_cv2(ctx):
    from ezcalib_b200 import detect
    imgs = detect.Images(ctx, np.stack([G["img0"], G["img1"]]))
    start = np.concatenate([G["start0"], G["start1"]])
    idx = np.concatenate([np.zeros(len(G["start0"]), np.int32), np.ones(len(G["start1"]), np.int32)])
    for win, mi, eps, key in (((5, 5), 100, 0.05, "ref_win5"), ((2, 2), 15, 0.1, "ref_win2")):
        got = detect.corner_subpix(ctx, imgs, start, idx, win, mi, eps)
        ora = np.concatenate([S.corner_subpix(G["img0"], G["start0"], win, mi, eps),
                              S.corner_subpix(G["img1"], G["start1"], win, mi, eps)])
        cv = np.concatenate([G[f"{key}_0"], G[f"{key}_1"]])
        assert np.array_equal(got, ora), np.abs(got - ora).max()   # bit-exact vs the restatement (== cv2 without IPP)
        assert np.abs(got - cv).max() <= 1e-4 < 0.01
    imgs.close()


def test_subpix_rejects_outside_corner(ctx):
    from ezcalib_b200 import _lib, detect
    imgs = detect.Images(ctx, G["img0"])
    with pytest.raises(_lib.EzcError, match="outside"):
        detect.corner_subpix(ctx, imgs, np.array([[-3.0, 5.0]], np.float32), [0])
    imgs.close()


def test_subpix_border_and_flat_regions(ctx):
    """Corners hugging every image border (replicated-border sampler) and on flat areas (singular system)."""
    from ezcalib_b200 import detect
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
    flat = np.full((64, 64), 128, np.uint8)
    imgs = detect.Images(ctx, img)
    got = detect.corner_subpix(ctx, imgs, pts, np.zeros(len(pts), np.int32))
    ora = S.corner_subpix(img, pts)
    # random border points are not corners: the 2x2 system is ill-conditioned there, so any deviation in the
    # arithmetic (summation order, FMA contraction, a 1-ulp mask entry) would show -- the kernel is bit-exact
    assert np.array_equal(got, ora), (np.abs(got - ora).max(), np.mean(np.all(got == ora, axis=1)))
    imgs.close()
    imgs = detect.Images(ctx, flat)
    p2 = np.array([[20.3, 30.7], [5.0, 5.0]], np.float32)
    assert np.array_equal(detect.corner_subpix(ctx, imgs, p2, [0, 0]), S.corner_subpix(flat, p2))
    imgs.close()

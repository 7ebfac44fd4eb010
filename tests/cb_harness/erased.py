"""Boards with outer squares painted over (shared by the CPU logic test and the GPU parity test): these exercise
orderFoundConnectedQuads' repair path (addOuterQuad), which plain fixtures never reach."""
import cv2
import numpy as np

FLAGS = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS


def erased_cases(img0, n_pairs=4, seed=1, step=1):
    """Yields (label, image): every `step`-th square of the outer ring erased, plus a few random pairs."""
    ok, c = cv2.findChessboardCorners(img0, (9, 6), flags=FLAGS)
    assert ok
    c = c.reshape(6, 9, 2)

    def sq_centre(i, j):
        ia, ib = int(min(max(np.floor(i - 0.5), 0), 7)), int(min(max(np.floor(j - 0.5), 0), 4))
        p00 = c[ib, ia]
        return p00 + (c[ib, ia + 1] - p00) * ((i - 0.5) - ia) + (c[ib + 1, ia] - p00) * ((j - 0.5) - ib)

    sz = min(np.linalg.norm(c[0, 1] - c[0, 0]), np.linalg.norm(c[1, 0] - c[0, 0]))
    ring = [(i, j) for i in range(10) for j in range(7) if i in (0, 9) or j in (0, 6)]
    cases = [[s] for s in ring[::step]]
    rng = np.random.default_rng(seed)
    for _ in range(n_pairs):
        a, b = rng.choice(len(ring), 2, replace=False)
        cases.append([ring[a], ring[b]])
    H, W = img0.shape
    for cs in cases:
        im = img0.copy()
        for (i, j) in cs:
            ctr = sq_centre(i, j)
            if 0 <= ctr[0] < W and 0 <= ctr[1] < H:
                cv2.circle(im, (int(ctr[0]), int(ctr[1])), int(sz * 0.72), 235, -1)
        yield cs, im

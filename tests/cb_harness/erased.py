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


def gate_corpus(G):
    """Views that stress CALIB_CB_FAST_CHECK: boards shifted partly out of view, shrunk boards (the gate rejects boards
    the full search would find), blur, noise, and the plain fixtures."""
    names = [str(n) for n in G["names"]]
    imgs = [G[nm + "_img"] for nm in names if G[nm + "_img"].shape == (480, 640)]
    rng = np.random.default_rng(0)
    for nm in ("render0", "warp1", "warp3"):
        im = G[nm + "_img"]
        for dx in (-300, -200, -120, 120, 200, 300):
            M = np.float32([[1, 0, dx], [0, 1, dx // 3]])
            imgs.append(cv2.warpAffine(im, M, (640, 480), borderValue=230))
        for s in (0.3, 0.5):
            small = cv2.resize(im, None, fx=s, fy=s)
            canvas = np.full((480, 640), 225, np.uint8)
            canvas[50:50 + small.shape[0], 80:80 + small.shape[1]] = small
            imgs.append(canvas)
        imgs.append(cv2.GaussianBlur(im, (0, 0), 3))
        imgs.append(np.clip(im.astype(np.int32) + rng.normal(0, 25, im.shape), 0, 255).astype(np.uint8))
    return imgs


def ideal_board(S, shear=0.0, W=1400, H=900, ry=1.0):
    """10x7 ideal (unblurred) squares of S px, optionally sheared / squashed: with `predilated` below this pins which
    dilation level each attempt of the histogram phase works on ("First attempt")."""
    img = np.full((H, W), 255, np.uint8)
    x0, y0 = 150, 120
    for r in range(7):
        for c in range(10):
            if (r + c) % 2 == 0:
                pts = [(x0 + cx * S + shear * cy * S, y0 + cy * S * ry) for (cx, cy) in ((c, r), (c + 1, r), (c + 1, r + 1), (c, r + 1))]
                cv2.fillConvexPoly(img, np.round(np.array(pts)).astype(np.int32), 0)
    return img


def predilated(img, levels):
    """Yields (m, image dilated m times) for m in `levels`."""
    cur, m = img.copy(), 0
    for target in sorted(levels):
        while m < target:
            cur = cv2.dilate(cur, None, iterations=1)
            m += 1
        yield m, cur.copy()

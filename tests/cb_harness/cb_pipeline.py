"""CPU harness: cv::findChessboardCorners re-assembled from (a) numpy restatements of the image-level steps and
(b) the product's own quad/grid logic (ezcalib_b200/csrc/cb_logic.h through libcb_harness.so), with cv2 primitives
(dilate / findContours / cornerSubPix) standing in for the CUDA kernels.  Used by tests/test_cb_logic.py to pin the
logic against the real cv2.findChessboardCorners before it runs on the GPU."""
import ctypes as C
import os
import subprocess

import cv2
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
VARIANT = os.environ.get('CB_VARIANT', 'A')
_lib = None


def lib():
    global _lib
    if _lib is None:
        so = os.path.join(HERE, "libcb_harness.so")
        src = os.path.join(HERE, "cb_harness.cpp")
        hdr = os.path.join(HERE, "..", "..", "ezcalib_b200", "csrc", "cb_logic.h")
        if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
            subprocess.check_call(["/usr/bin/g++", "-O2", "-g", "-fPIC", "-shared", "-std=c++17", "-o", so, src])
        _lib = C.CDLL(so)
    return _lib


def binarize_hist_threshold(img: np.ndarray) -> int:
    """icvBinarizationHistogramBased (calibinit.cpp): returns the threshold (0 = leave the image unchanged)."""
    H, W = img.shape
    max_pix = W * H
    max_pix1 = max_pix // 100
    hist = np.bincount(img.ravel(), minlength=256).astype(np.int64)
    smooth = np.zeros(256, np.int64)
    for i in range(256):
        lo, hi = max(0, i - 1), min(255, i + 1)
        smooth[i] = int(hist[lo:hi + 1].sum()) // 3
    grad = np.zeros(256, np.int64)
    prev = 0
    for i in range(1, 255):
        g = int(smooth[i - 1] - smooth[i + 1])
        if abs(g) < 100:
            g = -100 if prev == 0 else prev
        grad[i] = g
        prev = g
    maxpos = []
    i = 254
    while i > 2 and len(maxpos) < 20:
        if grad[i - 1] < 0 and grad[i] > 0:
            s = int(smooth[i - 1] + smooth[i] + smooth[i + 1])
            if not (s < max_pix1 and i < 64):
                maxpos.append(i)
        i -= 1
    n = len(maxpos)
    thresh = 0
    if n == 0:
        acc = 0
        for i in range(256):
            acc += int(hist[i])
            if acc > max_pix // 2:
                thresh = i
                break
    elif n == 1:
        thresh = maxpos[0] // 2
    elif n == 2:
        thresh = (maxpos[0] + maxpos[1]) // 2
    else:
        idx_acc, acc = 0, 0
        for i in range(255, 0, -1):
            acc += int(hist[i])
            if acc > max_pix // 18:
                idx_acc = i
                break
        idx_bg = 0
        bright = maxpos[0]
        for k in range(n - 1):
            idx_bg = k + 1
            if maxpos[k] < idx_acc:
                break
            bright = maxpos[k]
        maxval = int(hist[maxpos[idx_bg]])
        if maxpos[idx_bg] >= 250 and idx_bg + 1 < n:
            idx_bg += 1
            maxval = int(hist[maxpos[idx_bg]])
        for k in range(idx_bg + 1, n):
            if int(hist[maxpos[k]]) >= maxval:
                maxval = int(hist[maxpos[k]])
                idx_bg = k
        dist2 = (bright - maxpos[idx_bg]) // 2
        thresh = bright - dist2
    return thresh


def hole_contours(binary: np.ndarray):
    """Hole contours of the white foreground in the order generateQuads visits them (descending index)."""
    cs, hier = cv2.findContours(binary, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
    out = []
    if hier is None:
        return out
    h = hier[0]
    for idx in range(len(cs) - 1, -1, -1):
        parent = h[idx][3]
        if h[idx][2] != -1 or parent == -1:
            continue
        out.append((cs[idx].reshape(-1, 2).astype(np.int32), int(parent)))
    return out


def detect_from_binary(binary: np.ndarray, pw: int, ph: int, dilations: int = 0, prev_in: int = 0):
    holes = hole_contours(binary)
    if not holes:
        return False, np.zeros((0, 2), np.float32), int(prev_in), 0
    pts = np.ascontiguousarray(np.concatenate([c for c, _ in holes]), np.int32)
    lengths = np.array([len(c) for c, _ in holes], np.int32)
    # remap parent ids to small ints
    ids = {}
    parent = np.array([ids.setdefault(p, len(ids)) for _, p in holes], np.int32)
    out = np.zeros((pw * ph, 2), np.float32)
    n_out = C.c_int32(0); prev = C.c_int32(int(prev_in)); nq = C.c_int32(0)
    found = lib().cbh_detect_from_contours(pts.ctypes.data_as(C.c_void_p), lengths.ctypes.data_as(C.c_void_p),
                                           parent.ctypes.data_as(C.c_void_p), len(holes), pw, ph, int(dilations), out.ctypes.data_as(C.c_void_p),
                                           C.byref(n_out), C.byref(prev), C.byref(nq))
    return bool(found), out, prev.value, nq.value


def find_chessboard_corners(img: np.ndarray, pattern_size, refine=True, info=None, fast_check=False):
    """Restated cv::findChessboardCorners(img, pattern, ADAPTIVE_THRESH+NORMALIZE_IMAGE+FILTER_QUADS[+FAST_CHECK])."""
    pw, ph = pattern_size
    H, W = img.shape
    t = binarize_hist_threshold(img)
    thresh = np.where(img >= t, 255, 0).astype(np.uint8) if t > 0 else img.copy()
    if fast_check:
        import fast_check as FC
        if not FC.fast_check(img, thresh, pattern_size):
            return False, None
    found, corners = False, None
    prev_sqr_size = 0
    used = img
    for dilations in range(0, 8):
        # cv2 4.13: the first attempt runs on the undilated binary image (pinned on
        # pre-dilated ideal boards: cv2's first failing level matches only this way)
        if dilations > 0:
            thresh = cv2.dilate(thresh, None, iterations=1)
        cv2.rectangle(thresh, (0, 0), (W - 1, H - 1), 255, 3, cv2.LINE_8)
        ok, c, prev_sqr_size, nq = detect_from_binary(thresh, pw, ph, dilations, prev_sqr_size)   # in/out like cv2 (a group can reset it to 0)
        if ok:
            found, corners = True, c
            if info is not None:
                info.update(phase=1, dilations=dilations, thresh=t, quads=nq)
            break
    if not found:
        eq = cv2.equalizeHist(img)
        used = eq
        prev_sqr_size = 0
        prev_thresh_img = None
        for k in range(6):
            if found:
                break
            prev_block = -1
            for dilations in range(0, 8):
                block = int(round(min(W, H) * (0.2 if k % 2 == 0 else 0.1))) if prev_sqr_size == 0 else int(round(prev_sqr_size * 2))
                block |= 1
                if VARIANT == "A":
                    if block != prev_block:
                        th = cv2.adaptiveThreshold(eq, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, block, (k // 2) * 5)
                        if dilations > 0:
                            th = cv2.dilate(th, None, iterations=dilations)
                        prev_thresh_img = th.copy()
                    else:
                        if dilations > 0:
                            prev_thresh_img = cv2.dilate(prev_thresh_img, None, iterations=1)
                        th = prev_thresh_img.copy()
                else:
                    if block != prev_block:
                        th = cv2.adaptiveThreshold(eq, 255, cv2.ADAPTIVE_THRESH_MEAN_C, cv2.THRESH_BINARY, block, (k // 2) * 5)
                        if dilations > 0 and dilations - 1 > 0:
                            th = cv2.dilate(th, None, iterations=dilations - 1)
                    else:
                        th = cv2.dilate(th, None, iterations=1)
                prev_block = block
                cv2.rectangle(th, (0, 0), (W - 1, H - 1), 255, 3, cv2.LINE_8)
                ok, c, prev_sqr_size, nq = detect_from_binary(th, pw, ph, {'1': dilations, '0': 0, 'm1': max(dilations - 1, 0), 'h': dilations // 2, 'p1': dilations + 1}[os.environ.get('CB_PHASE2_COMP', '1')], prev_sqr_size)
                if ok:
                    found, corners = True, c
                    if info is not None:
                        info.update(phase=2, k=k, dilations=dilations, block=block, quads=nq)
                    break
    if found:
        found = bool(lib().cbh_check_monotony(corners.ctypes.data_as(C.c_void_p), pw, ph))
    if found:
        B = 8
        if np.any((corners[:, 0] <= B) | (corners[:, 0] > W - B) | (corners[:, 1] <= B) | (corners[:, 1] > H - B)):
            found = False
    if found:
        if ph % 2 == 0 and pw % 2 == 0:
            if corners[(ph - 1) * pw, 1] - corners[0, 1] < 0:
                corners = corners[::-1].copy()
        if refine:
            corners = cv2.cornerSubPix(used, corners.reshape(-1, 1, 2).copy(), (2, 2), (-1, -1),
                                       (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_MAX_ITER, 15, 0.1)).reshape(-1, 2)
    return found, corners

"""CPU restatement of CALIB_CB_FAST_CHECK (OpenCV calib3d checkchessboard.cpp: checkChessboardBinary + checkChessboard)
on cv2 primitives.  cv2.checkChessboard is exported and pins the second function directly; the first one is pinned
through cv2.findChessboardCorners (an image that only the fast check can reject comes back in a few ms)."""
import math

import cv2
import numpy as np


def _hypotheses(fg: np.ndarray, class_id: int, quads: list):
    """icvGetQuadrangleHypotheses: outer contours only, minAreaRect size / aspect filters."""
    contours, hier = cv2.findContours(fg, cv2.RETR_CCOMP, cv2.CHAIN_APPROX_SIMPLE)
    if hier is None:
        return
    for c, h in zip(contours, hier[0]):
        if h[3] != -1:
            continue
        (_, _), (bw, bh), _ = cv2.minAreaRect(c)
        box_size = max(bw, bh)
        if box_size < 10.0:
            continue
        aspect = bw / max(bh, 1)
        if aspect < 0.3 or aspect > 3.0:
            continue
        quads.append((np.float32(box_size), class_id))


def fill_quads(white, black, white_thresh, black_thresh):
    quads = []
    _, th = cv2.threshold(white, white_thresh, 255, cv2.THRESH_BINARY)
    _hypotheses(th, 1, quads)
    _, th = cv2.threshold(black, black_thresh, 255, cv2.THRESH_BINARY_INV)
    _hypotheses(th, 0, quads)
    return quads


def check_quads(quads, size) -> bool:
    w, h = size
    min_quads_count = (w * h) // 2
    quads = sorted(quads, key=lambda q: q[0])
    size_rel_dev = np.float32(0.4)
    n = len(quads)
    black_count = round(math.ceil(w / 2.0) * math.ceil(h / 2.0))
    white_count = round(math.floor(w / 2.0) * math.floor(h / 2.0))
    for i in range(n):
        j = i + 1
        while j < n:
            if np.float32(quads[j][0] / quads[i][0]) > np.float32(1.0) + size_rel_dev:
                break
            j += 1
        if j + 1 > min_quads_count + i:
            counts = [0, 0]
            for k in range(i, j):
                counts[quads[k][1]] += 1
            if counts[0] < black_count * 0.75 or counts[1] < white_count * 0.75:
                continue
            return True
    return False


def check_chessboard_binary(binary: np.ndarray, size) -> bool:
    white, black = binary.copy(), binary.copy()
    for erosion in range(4):
        if erosion:
            white = cv2.erode(white, None, iterations=1)
            black = cv2.dilate(black, None, iterations=1)
        if check_quads(fill_quads(white, black, 128, 128), size):
            return True
    return False


def check_chessboard(img: np.ndarray, size) -> bool:
    white = cv2.erode(img, None, iterations=1)
    black = cv2.dilate(img, None, iterations=1)
    t = 20.0
    while t < 130.0:
        if check_quads(fill_quads(white, black, t + 70.0, t), size):
            return True
        t += 20.0
    return False


def fast_check(img: np.ndarray, binary: np.ndarray, size) -> bool:
    """The CALIB_CB_FAST_CHECK gate of cv::findChessboardCorners: binary = histogram-binarised image."""
    return check_chessboard_binary(binary, size) or check_chessboard(img, size)

"""CPU restatement of cv::cornerSubPix as the reference calls it.

TEST INFRASTRUCTURE ONLY (see oracle/lm_oracle.cpp header for the rule).

Reference call sites:
  /root/reference/include/target_detectors/chessboard_detector.hpp:39-40
  /root/reference/include/target_detectors/aprilgrid_detector.hpp:52-53
      cv::cornerSubPix(img, corners, Size(5,5), Size(-1,-1), {COUNT+EPS, 100, 0.05})
The algorithm lives in OpenCV (un-vendored dependency; "OpenCV 3 or 4", README.md:67):
modules/imgproc/src/cornersubpix.cpp + samplers.cpp (getRectSubPix_8u32f).  It is restated here
from the published source and PINNED against the real thing: this container has cv2 4.13.0, and
tests/test_subpix_oracle.py checks this restatement against cv2.cornerSubPix on rendered corners
(agreement ~1e-5 px; cv2 routes getRectSubPix through IPP, whose float rounding differs in the
last bits) and against the committed golden vectors tests/golden/subpix_*.npz (made by
tests/golden/make_subpix_golden.py with cv2).
"""
from __future__ import annotations

import math

import numpy as np

F = np.float32


def _floor(v: np.float32) -> int:
    return int(math.floor(float(v)))


def _expf(x) -> np.float32:
    """std::exp(float) as the reference's C++ gets it: glibc expf.  numpy's float32 exp is a different (SIMD)
    implementation and is 1 ulp off for several of the mask arguments (e.g. exp(-1)), which shows up in the last bits
    of ill-conditioned windows."""
    global _LIBM
    if _LIBM is None:
        import ctypes
        import ctypes.util
        _LIBM = ctypes.CDLL(ctypes.util.find_library("m") or "libm.so.6")
        _LIBM.expf.restype = ctypes.c_float
        _LIBM.expf.argtypes = [ctypes.c_float]
    return F(_LIBM.expf(float(x)))


_LIBM = None


def get_rect_subpix_8u32f(img: np.ndarray, win_w: int, win_h: int, cx: np.float32, cy: np.float32) -> np.ndarray:
    """getRectSubPix_8u32f (samplers.cpp): float32 patch win_h x win_w centred at (cx,cy)."""
    H, W = img.shape
    centre_x = F(F(cx) - F((win_w - 1) * 0.5))
    centre_y = F(F(cy) - F((win_h - 1) * 0.5))
    ipx, ipy = _floor(centre_x), _floor(centre_y)
    out = np.zeros((win_h, win_w), dtype=F)
    if 0 <= ipx and ipx + win_w < W and 0 <= ipy and ipy + win_h < H:
        a = F(centre_x - F(ipx))
        b = F(centre_y - F(ipy))
        a = max(a, F(0.0001))
        a12 = F(a * F(F(1.0) - b))
        a22 = F(a * b)
        b1 = F(F(1.0) - b)
        b2 = b
        s = (1.0 - float(a)) / float(a)  # double
        for i in range(win_h):
            r0 = img[ipy + i, ipx:ipx + win_w + 1].astype(F)
            r1 = img[ipy + i + 1, ipx:ipx + win_w + 1].astype(F)
            prev = F(F(F(1.0) - a) * F(F(b1 * r0[0]) + F(b2 * r1[0])))
            for j in range(win_w):
                t = F(F(a12 * r0[j + 1]) + F(a22 * r1[j + 1]))
                out[i, j] = F(prev + t)
                prev = F(float(t) * s)
        return out
    # generic path with replicated border (getRectSubPix_Cn_ + adjustRect)
    a = F(centre_x - F(ipx))
    b = F(centre_y - F(ipy))
    a11 = F(F(F(1.0) - a) * F(F(1.0) - b))
    a12 = F(a * F(F(1.0) - b))
    a21 = F(F(F(1.0) - a) * b)
    a22 = F(a * b)
    b1 = F(F(1.0) - b)
    b2 = b
    # adjustRect
    if ipx >= 0:
        col0 = ipx
        rx = 0
    else:
        col0 = 0
        rx = min(-ipx, win_w)
    if ipx < W - win_w:
        rw = win_w
    else:
        rw = W - ipx - 1
        if rw < 0:
            col0 += rw
            rw = 0
    if ipy >= 0:
        row = ipy
        ry = 0
    else:
        row = 0
        ry = -ipy
    if ipy < H - win_h:
        rh = win_h
    else:
        rh = H - ipy - 1
        if rh < 0:
            row += rh
            rh = 0
    base = col0 - rx  # src pointer column such that src[j] is image column base + j
    for i in range(win_h):
        row2 = row + 1
        if i < ry or i >= rh:
            row2 = row
        rr = min(max(row, 0), H - 1)
        rr2 = min(max(row2, 0), H - 1)

        def px(r, j):
            return F(img[r, min(max(base + j, 0), W - 1)])
        s0 = F(F(px(rr, rx) * b1) + F(px(rr2, rx) * b2))
        for j in range(0, rx):
            out[i, j] = s0
        s0 = F(F(px(rr, rw) * b1) + F(px(rr2, rw) * b2))
        for j in range(rw, win_w):
            out[i, j] = s0
        for j in range(rx, rw):
            v = F(F(F(px(rr, j) * a11) + F(px(rr, j + 1) * a12)) + F(px(rr2, j) * a21))
            out[i, j] = F(v + F(px(rr2, j + 1) * a22))
        if i < rh:
            row = row2
    return out


def corner_subpix(img: np.ndarray, corners: np.ndarray, win=(5, 5), max_iter: int = 100, eps: float = 0.05,
                  return_iters: bool = False):
    """cv::cornerSubPix(img, corners, win, (-1,-1), {COUNT+EPS, max_iter, eps}); img u8 HxW."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    ww, wh = win
    win_w, win_h = 2 * ww + 1, 2 * wh + 1
    max_iters = min(max(max_iter, 1), 100)
    eps2 = max(eps, 0.0) ** 2
    mask = np.zeros((win_h, win_w), dtype=F)
    for i in range(win_h):
        y = F(F(i - wh) / F(wh))
        vy = _expf(F(-y * y))
        for j in range(win_w):
            x = F(F(j - ww) / F(ww))
            mask[i, j] = F(vy * _expf(F(-x * x)))
    out = np.array(corners, dtype=F, copy=True).reshape(-1, 2)
    iters = np.zeros(len(out), dtype=np.int32)
    py = (np.arange(win_h, dtype=np.float64) - wh)[:, None]
    pxv = (np.arange(win_w, dtype=np.float64) - ww)[None, :]
    m64 = mask.astype(np.float64)
    for k in range(len(out)):
        cT = out[k].copy()
        cI = cT.copy()
        it = 0
        while True:
            sub = get_rect_subpix_8u32f(img, win_w + 2, win_h + 2, cI[0], cI[1])
            tgx = (sub[1:-1, 2:] - sub[1:-1, :-2]).astype(np.float64)  # float32 difference, widened
            tgy = (sub[2:, 1:-1] - sub[:-2, 1:-1]).astype(np.float64)
            gxx = tgx * tgx * m64
            gxy = tgx * tgy * m64
            gyy = tgy * tgy * m64
            # sequential raster-order accumulation like the reference's scalar loop (np.sum is pairwise: differs in
            # the last bits, which shows on ill-conditioned windows); cumsum adds strictly left to right
            seq = lambda x: float(np.cumsum(x.ravel())[-1])
            a = seq(gxx); b = seq(gxy); c = seq(gyy)
            bb1 = seq(gxx * pxv + gxy * py)
            bb2 = seq(gxy * pxv + gyy * py)
            det = a * c - b * b
            if abs(det) <= np.finfo(np.float64).eps ** 2:
                break
            scale = 1.0 / det
            c2 = np.array([F(float(cI[0]) + c * scale * bb1 - b * scale * bb2),
                           F(float(cI[1]) - b * scale * bb1 + a * scale * bb2)], dtype=F)
            dx = F(c2[0] - cI[0]); dy = F(c2[1] - cI[1])
            err = float(F(F(dx * dx) + F(dy * dy)))  # Point2f arithmetic: float32
            cI = c2
            if cI[0] < 0 or cI[0] >= W or cI[1] < 0 or cI[1] >= H:
                break
            it += 1
            if not (it < max_iters and err > eps2):
                break
        if abs(float(F(cI[0] - cT[0]))) > ww or abs(float(F(cI[1] - cT[1]))) > wh:
            cI = cT
        out[k] = cI
        iters[k] = it
    out = out.reshape(np.asarray(corners).shape)
    return (out, iters) if return_iters else out

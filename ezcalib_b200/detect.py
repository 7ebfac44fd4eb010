"""Host-side driver of the batched target detection (half (a) of the hot path).

Mirrors the reference's detector interfaces (/root/reference/include/target_detectors/):
CalibDetector::detectTarget(img, corners) -> bool and getTargetCoords(); here the detectors take a whole
batch of images resident in HBM and return per-image found flags + corner arrays.  All arithmetic runs in
libezcalib_b200.so (CUDA); this file only marshals buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ImageBatchC, check
from .lm import Context


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Images:
    """Device-resident batch of u8 gray images [n, H, W]."""

    def __init__(self, ctx: Context, images=None, device_ptr: int | None = None, shape=None):
        self.ctx = ctx
        b = ImageBatchC()
        if device_ptr is not None:
            n, h, w = shape
            b.data = C.c_void_p(int(device_ptr))
            self._keep = None
        else:
            images = np.ascontiguousarray(images, dtype=np.uint8)
            if images.ndim == 2:
                images = images[None]
            n, h, w = images.shape
            b.data = _p(images)
            self._keep = images
        b.n_images, b.height, b.width = int(n), int(h), int(w)
        b.row_stride, b.image_stride = 0, 0
        self.n, self.height, self.width = int(n), int(h), int(w)
        self._h = C.c_void_p()
        fn = _lib.lib().ezc_images_wrap_device if device_ptr is not None else _lib.lib().ezc_images_upload
        check(fn(ctx.handle, C.byref(b), C.byref(self._h)), "ezc_images_upload")

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            _lib.lib().ezc_images_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def corner_subpix(ctx: Context, imgs: Images, corners, image_index, win=(5, 5), max_iter: int = 100, eps: float = 0.05):
    """cv::cornerSubPix for corners [n,2] (float32) lying in images image_index[n]."""
    c = np.array(corners, dtype=np.float32, copy=True).reshape(-1, 2)
    idx = np.ascontiguousarray(image_index, dtype=np.int32).reshape(-1)
    assert idx.shape[0] == c.shape[0]
    check(_lib.lib().ezc_corner_subpix(ctx.handle, imgs.handle, _p(c), _p(idx), c.shape[0], int(win[0]), int(win[1]),
                                       int(max_iter), float(eps)), "ezc_corner_subpix")
    return c

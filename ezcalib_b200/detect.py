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

    def __init__(self, ctx: Context, images=None, device_ptr: int | None = None, shape=None, async_upload: bool = False):
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
        L = _lib.lib()
        fn = L.ezc_images_wrap_device if device_ptr is not None else (L.ezc_images_upload_async if async_upload else L.ezc_images_upload)
        check(fn(ctx.handle, C.byref(b), C.byref(self._h)), "ezc_images_upload")

    @property
    def handle(self):
        return self._h

    @classmethod
    def from_handle(cls, ctx, handle, n, height, width):
        self = cls.__new__(cls)
        self.ctx, self._h, self._keep = ctx, handle, None
        self.n, self.height, self.width = int(n), int(height), int(width)
        return self

    def download(self, first: int = 0, count: int | None = None) -> np.ndarray:
        count = self.n - first if count is None else count
        out = np.zeros((count, self.height, self.width), np.uint8)
        check(_lib.lib().ezc_images_download(self.ctx.handle, self._h, int(first), int(count), _p(out)), "ezc_images_download")
        return out

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


def detect_aprilgrid(ctx: Context, imgs: Images, nb_rows: int, nb_cols: int, max_images_in_flight: int = 0):
    """AprilTagCalibDetector::detectTarget for every image: returns (success[n] bool, corners[n, 4*nb_tags, 2] float32
    with (-1,-1) for missing corners, n_detections[n])."""
    nb_tags = nb_rows * nb_cols
    corners = np.zeros((imgs.n, 4 * nb_tags, 2), np.float32)
    ndet = np.zeros(imgs.n, np.int32)
    succ = np.zeros(imgs.n, np.uint8)
    check(_lib.lib().ezc_detect_aprilgrid_batch(ctx.handle, imgs.handle, int(nb_rows), int(nb_cols), _p(corners), _p(ndet), _p(succ),
                                                int(max_images_in_flight)), "ezc_detect_aprilgrid_batch")
    return succ.astype(bool), corners, ndet


def apriltag_debug(ctx: Context, imgs: Images, image: int, nb_rows: int, nb_cols: int) -> dict:
    """Stage products of the GPU extractTags for one image (tests)."""
    L = _lib.lib()
    nb_tags = nb_rows * nb_cols
    h = C.c_void_p()
    corners = np.zeros((4 * nb_tags, 2), np.float32)
    ndet = np.zeros(1, np.int32)
    succ = np.zeros(1, np.uint8)
    check(L.ezc_apriltag_debug_run(ctx.handle, imgs.handle, int(image), int(nb_rows), int(nb_cols), C.byref(h), _p(corners), _p(ndet), _p(succ)),
          "ezc_apriltag_debug_run")
    try:
        npx, ne, ns, nq = (L.ezc_apriltag_debug_count(h, k) for k in range(4))
        theta = np.zeros(npx, np.float32); mag = np.zeros(npx, np.float32)
        edges = np.zeros((ne, 3), np.int32); rep = np.zeros(npx, np.int32); size = np.zeros(npx, np.int32)
        segs = np.zeros((ns, 6), np.float32); quads = np.zeros((nq, 9), np.float32)
        check(L.ezc_apriltag_debug_get(h, _p(theta), _p(mag), _p(edges), _p(rep), _p(size), _p(segs), _p(quads)), "ezc_apriltag_debug_get")
    finally:
        L.ezc_apriltag_debug_free(h)
    shp = (imgs.height, imgs.width)
    return dict(theta=theta.reshape(shp), mag=mag.reshape(shp), edges=edges, rep=rep.reshape(shp), size=size.reshape(shp),
                segments=segs, quads=quads, corners=corners, n_detections=int(ndet[0]), success=bool(succ[0]))


def render_targets(ctx: Context, pattern: str, rows: int, cols: int, size: float, space: float, model_id: int, width: int,
                   height: int, cams12, poses7, noise_sigma: float = 2.0, seed: int = 20261019) -> Images:
    """Render synthetic target views straight into HBM (bench / large tests). pattern: 'chessboard' | 'aprilgrid'."""
    cams = np.ascontiguousarray(cams12, dtype=np.float64).reshape(-1, 12)
    poses = np.ascontiguousarray(poses7, dtype=np.float64).reshape(-1, 7)
    assert cams.shape[0] == poses.shape[0]
    h = C.c_void_p()
    check(_lib.lib().ezc_render_targets(ctx.handle, 0 if pattern == "chessboard" else 1, int(rows), int(cols), float(size), float(space),
                                        int(model_id), int(width), int(height), int(poses.shape[0]), _p(cams), _p(poses), float(noise_sigma),
                                        int(seed), C.byref(h)), "ezc_render_targets")
    return Images.from_handle(ctx, h, poses.shape[0], height, width)


def find_chessboard_corners(ctx: Context, imgs: Images, pattern_size, final_subpix: bool = False, max_images_in_flight: int = 0,
                            fast_check: bool = False):
    """cv::findChessboardCorners(img, pattern_size, ADAPTIVE_THRESH+NORMALIZE_IMAGE+FILTER_QUADS+FAST_CHECK) for the batch;
    with final_subpix also the reference's cornerSubPix (5,5).  Returns (found[n] bool, corners[n, pw*ph, 2] float32)."""
    pw, ph = pattern_size
    corners = np.zeros((imgs.n, pw * ph, 2), np.float32)
    found = np.zeros(imgs.n, np.uint8)
    check(_lib.lib().ezc_find_chessboard_corners_batch(ctx.handle, imgs.handle, int(pw), int(ph), int(bool(final_subpix)), int(bool(fast_check)), _p(corners), _p(found),
                                                       int(max_images_in_flight)), "ezc_find_chessboard_corners_batch")
    return found.astype(bool), corners


def detect_chessboard(ctx: Context, imgs: Images, nb_rows: int, nb_cols: int, max_images_in_flight: int = 0):
    """ChessboardCalibDetector::detectTarget for every image (nb_rows x nb_cols SQUARES, like the reference)."""
    return find_chessboard_corners(ctx, imgs, (nb_cols - 1, nb_rows - 1), True, max_images_in_flight, fast_check=True)

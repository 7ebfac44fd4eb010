"""ctypes wrapper of oracle/_ref/libapriltag_ref.so: the reference's OWN AprilTag detector
(thirdparty/ethz_apriltag2, compiled unmodified from /root/reference by oracle/apriltag_ref/Makefile) plus a
restatement of AprilTagCalibDetector::detectTarget around it.

TEST INFRASTRUCTURE ONLY (tests/, smoke(), bench.py cpu_baseline / --impl reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "libapriltag_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH) or os.path.isdir("/root/reference/thirdparty/ethz_apriltag2")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            if not os.path.isdir("/root/reference/thirdparty/ethz_apriltag2"):
                raise RuntimeError("oracle/_ref/libapriltag_ref.so is missing and /root/reference is not present to build it")
            env = dict(os.environ); env.pop("CXX", None)
            subprocess.check_call(["make", "-C", os.path.join(_HERE, "apriltag_ref"), "-j8"], stdout=subprocess.DEVNULL, env=env)
        _lib = C.CDLL(_LIB_PATH)
        _lib.ref_code36h11.restype = C.c_uint64
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def codes36h11() -> np.ndarray:
    n = lib().ref_num_codes36h11()
    return np.array([lib().ref_code36h11(i) for i in range(n)], dtype=np.uint64)


def extract_tags(img: np.ndarray, max_det: int = 4096) -> dict:
    """AprilTags::TagDetector(tagCodes36h11).extractTags(img): detections in the reference's output order."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    ids = np.zeros(max_det, np.int32); ham = np.zeros(max_det, np.int32); rot = np.zeros(max_det, np.int32)
    p = np.zeros((max_det, 4, 2), np.float32); cxy = np.zeros((max_det, 2), np.float32)
    per = np.zeros(max_det, np.float32); code = np.zeros(max_det, np.uint64)
    n = lib().ref_extract_tags(_p(img), W, H, max_det, _p(ids), _p(ham), _p(rot), _p(p), _p(cxy), _p(per), _p(code))
    assert n <= max_det
    return dict(n=n, id=ids[:n], hamming=ham[:n], rotation=rot[:n], p=p[:n], cxy=cxy[:n], perimeter=per[:n], obs_code=code[:n])


def stages(img: np.ndarray, max_segs: int = 65536, max_quads: int = 16384) -> dict:
    """Intermediate products of extractTags steps 1-7 (see ref_driver.cc)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    seg = np.zeros((H, W), np.float32); th = np.zeros((H, W), np.float32); mag = np.zeros((H, W), np.float32)
    max_edges = 4 * W * H
    edges = np.zeros((max_edges, 3), np.int32)
    rep = np.zeros((H, W), np.int32); size = np.zeros((H, W), np.int32)
    segs = np.zeros((max_segs, 6), np.float32); quads = np.zeros((max_quads, 9), np.float32)
    counts = np.zeros(4, np.int32)
    lib().ref_stages(_p(img), W, H, _p(seg), _p(th), _p(mag), _p(edges), max_edges, _p(rep), _p(size), _p(segs), max_segs,
                     _p(quads), max_quads, _p(counts))
    ne, ns, nq = int(counts[0]), int(counts[1]), int(counts[2])
    return dict(seg=seg, theta=th, mag=mag, edges=edges[:ne].copy(), rep=rep, size=size, segments=segs[:ns].copy(),
                quads=quads[:nq].copy())


def _corner_subpix(img, pts):
    try:
        import cv2
        crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)
        return cv2.cornerSubPix(img, pts.copy().reshape(-1, 1, 2), (5, 5), (-1, -1), crit).reshape(-1, 2)
    except ImportError:
        from . import subpix_oracle
        return subpix_oracle.corner_subpix(img, pts, (5, 5), 100, 0.05)


def detect_target(img: np.ndarray, nb_rows: int, nb_cols: int):
    """AprilTagCalibDetector::detectTarget (/root/reference/include/target_detectors/aprilgrid_detector.hpp:30-71):
    returns (success, corners[4*nb_tags,2] float32 with (-1,-1) for missing, n_detections, raw detections)."""
    img = np.ascontiguousarray(img, dtype=np.uint8)
    H, W = img.shape
    nb_tags = nb_rows * nb_cols
    det = extract_tags(img)
    corners = np.full((4 * nb_tags, 2), -1.0, np.float32)
    for i in range(det["n"]):
        tid = int(det["id"][i])
        if tid >= nb_tags:
            continue
        raw = det["p"][i].astype(np.float32)
        inside = (raw[:, 0] >= 0) & (raw[:, 0] < W) & (raw[:, 1] >= 0) & (raw[:, 1] < H)
        # cv::cornerSubPix CV_Asserts that every corner lies inside the image; the reference would throw there.
        ref = _corner_subpix(img, raw) if inside.all() else raw
        corners[tid * 4: tid * 4 + 4] = ref
    threshold = (nb_tags // 2) - 1
    return det["n"] > threshold, corners, det["n"], det

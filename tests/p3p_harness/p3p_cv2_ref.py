"""The cv2 side of the P3P parity tests: EZCalibrator::computeP3PRansac (/root/reference/src/ezcalibrator.cpp:560-628)
written over the real OpenCV calls the reference makes (cv2.solvePnPRansac with its exact arguments, cv2.Rodrigues).
Test infrastructure only: the product path (ezcalib_b200/) does not import cv2."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from ezcalib_b200 import synth


def fit_to_so3(R: np.ndarray) -> np.ndarray:
    """Sophus::SO3d::fitToSO3 (closest rotation in the Frobenius sense: U diag(1,1,det(UV^T)) V^T)."""
    U, _, Vt = np.linalg.svd(R)
    D = np.diag([1.0, 1.0, np.linalg.det(U @ Vt)])
    return U @ D @ Vt


def normalise(corners_px, target, focal, pp, width, height):
    fx, fy = float(focal[0]), float(focal[-1])
    u, v = corners_px[:, 0].astype(np.float32), corners_px[:, 1].astype(np.float32)
    inside = (u > -0.5) & (v > -0.5) & (u < np.float32(width) - np.float32(0.5)) & (v < np.float32(height) - np.float32(0.5))
    # Eigen's cofactor inverse of K applied to (u, v, 1)
    invdet = 1.0 / (fx * fy)
    a00, a02 = fy * invdet, (0.0 * pp[1] - pp[0] * fy) * invdet
    a11, a12 = fx * invdet, (pp[0] * 0.0 - fx * pp[1]) * invdet
    x = ((a00 * u.astype(np.float64) + 0.0 * v.astype(np.float64)) + a02).astype(np.float32)[inside]
    y = ((0.0 * u.astype(np.float64) + a11 * v.astype(np.float64)) + a12).astype(np.float32)[inside]
    return np.stack([x, y], axis=1), np.asarray(target, np.float64)[inside].astype(np.float32), np.nonzero(inside)[0]


def compute_p3p_ransac_cv2(corners_px, target, focal, pp, width, height, n_target_pts=None):
    """Returns (ok, pose7, inlier target indices, rvec_tvec[6])."""
    import cv2
    n_target_pts = len(target) if n_target_pts is None else n_target_pts
    un, w3, src = normalise(corners_px, target, focal, pp, width, height)
    if len(un) < 32 and n_target_pts >= 32:
        return False, None, np.zeros(0, int), None
    fx, fy = float(focal[0]), float(focal[-1])
    mean_focal = fx if len(focal) == 1 else 0.5 * (fx + fy)
    ok, rvec, tvec, inl = cv2.solvePnPRansac(w3.reshape(-1, 1, 3), un.reshape(-1, 1, 2), np.eye(3, dtype=np.float32), None,
                                             useExtrinsicGuess=False, iterationsCount=1000, reprojectionError=float(4.0 / mean_focal),
                                             confidence=0.999, flags=cv2.SOLVEPNP_AP3P)
    inl = np.zeros(0, int) if inl is None else src[inl.ravel()]
    rt = None if rvec is None else np.concatenate([rvec.ravel(), tvec.ravel()])
    if not ok or (len(inl) < 32 and n_target_pts >= 32):
        return False, None, inl, rt
    R, _ = cv2.Rodrigues(rvec)
    R = fit_to_so3(R.astype(np.float32).astype(np.float64))
    t = tvec.reshape(3).astype(np.float32).astype(np.float64)
    return True, np.concatenate([synth.quat_from_matrix(R), t]), inl, rt


_HERE = os.path.dirname(os.path.abspath(__file__))
_harness = None


def harness():
    """tests/p3p_harness: the product's csrc/p3p_core.h compiled for the host."""
    global _harness
    if _harness is None:
        subprocess.check_call(["make", "-C", _HERE], stdout=subprocess.DEVNULL)
        _harness = C.CDLL(os.path.join(_HERE, "libp3p_harness.so"))
    return _harness


def compute_p3p_ransac_harness(corners_px, target, focal, pp, width, height, n_target_pts=None):
    n_target_pts = len(target) if n_target_pts is None else n_target_pts
    c = np.ascontiguousarray(corners_px, np.float32)
    tg = np.ascontiguousarray(target, np.float64)
    fx, fy = float(focal[0]), float(focal[-1])
    mean_focal = fx if len(focal) == 1 else 0.5 * (fx + fy)
    rt, pose = np.zeros(6), np.zeros(7)
    mask = np.zeros(len(c), np.uint8)
    ninl, nit = C.c_int(0), C.c_int(0)
    D = C.c_double
    ok = harness().p3p_harness_view(c.ctypes.data_as(C.c_void_p), tg.ctypes.data_as(C.c_void_p), len(c), int(n_target_pts), D(fx), D(fy),
                                    D(float(pp[0])), D(float(pp[1])), int(width), int(height), 1000, D(4.0 / mean_focal), D(0.999), 32,
                                    rt.ctypes.data_as(C.c_void_p), pose.ctypes.data_as(C.c_void_p), mask.ctypes.data_as(C.c_void_p),
                                    C.byref(ninl), C.byref(nit))
    return bool(ok), (pose if ok else None), np.nonzero(mask)[0], rt, nit.value

"""ctypes wrapper of oracle/_ref/liblm_ref.so: the reference's OWN functor headers (include/dist_models/*.hpp,
include/ceres_params/se3_param.hpp) compiled unmodified against stand-in ceres / Sophus / Eigen headers
(oracle/lm_ref/).  TEST INFRASTRUCTURE ONLY -- never imported by the product package.

The library is built in this container (where /root/reference exists) by oracle/lm_ref/Makefile; the prebuilt .so
travels to the GPU box, where it is only loaded."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_ref", "liblm_ref.so")
_lib = None


def available() -> bool:
    return os.path.exists(_LIB_PATH) or os.path.isdir("/root/reference/include/dist_models")


def lib():
    global _lib
    if _lib is None:
        if os.path.isdir("/root/reference/include/dist_models"):
            subprocess.check_call(["make", "-C", os.path.join(_HERE, "lm_ref")], stdout=subprocess.DEVNULL)
        _lib = C.CDLL(_LIB_PATH)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _f64(*xs):
    return [np.ascontiguousarray(x, dtype=np.float64) for x in xs]


def num_dist(model_id: int) -> int:
    return lib().lmref_num_dist(int(model_id))


def residual_jacobian(model_id, mono, focal, pp, dist, pose, wpt, u, v):
    """createCeresCostFunction(u, v, wpt)->Evaluate + AutoDiffLocalLeftSE3::PlusJacobian.
    Returns r[2], J_local[2, nf+2+nd+6] (columns focal|pp|dist|pose6), J_ambient[2, nf+2+nd+7]."""
    nf = 1 if mono else 2
    nd = num_dist(model_id)
    r = np.zeros(2)
    Ja = np.zeros((2, nf + 2 + nd + 7))
    Jl = np.zeros((2, nf + 2 + nd + 6))
    a = _f64(focal, pp, dist, pose, wpt)
    rc = lib().lmref_eval(int(model_id), int(bool(mono)), _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3]), _p(a[4]),
                          C.c_double(float(u)), C.c_double(float(v)), _p(r), _p(Ja), _p(Jl))
    if rc != 0:
        raise RuntimeError(f"lmref_eval failed: {rc}")
    return r, Jl, Ja


def se3_plus(x, delta):
    out = np.zeros(7)
    a = _f64(x, delta)
    assert lib().lmref_plus(_p(a[0]), _p(a[1]), _p(out)) == 0
    return out


def se3_plus_jacobian(x):
    out = np.zeros((7, 6))
    a = _f64(x)
    assert lib().lmref_plus_jacobian(_p(a[0]), _p(out)) == 0
    return out


def se3_minus(y, x):
    out = np.zeros(6)
    a = _f64(y, x)
    assert lib().lmref_minus(_p(a[0]), _p(a[1]), _p(out)) == 0
    return out


def distort(model_id, dist, x, y):
    out = np.zeros(2)
    a = _f64(dist)
    assert lib().lmref_distort(int(model_id), _p(a[0]), C.c_double(float(x)), C.c_double(float(y)), _p(out)) == 0
    return out


def undistort(model_id, dist, xd, yd):
    out = np.zeros(2)
    a = _f64(dist)
    assert lib().lmref_undistort(int(model_id), _p(a[0]), C.c_double(float(xd)), C.c_double(float(yd)), _p(out)) == 0
    return out


def dist_std_quirk(model_id, cov):
    """DistParam::setDistParamsStd then getDistParamsStd: 2*nd values, the first nd are 0 (reference bug, kept)."""
    out = np.zeros(32)
    a = _f64(cov)
    n = lib().lmref_dist_std_quirk(int(model_id), _p(a[0]), _p(out), 32)
    return out[:n].copy()

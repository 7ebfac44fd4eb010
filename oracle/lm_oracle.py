"""ctypes wrapper of oracle/liblm_oracle.so ("restated Ceres" CPU LM solve).

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from the product package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liblm_oracle.so")


class _Problem(C.Structure):
    _fields_ = [("model", C.c_int32), ("mono", C.c_int32), ("n_blocks", C.c_int32), ("obs_per_block", C.c_int32),
                ("pts_period", C.c_int32), ("shared_per_block", C.c_int32), ("img_w", C.c_double), ("img_h", C.c_double),
                ("pts", C.c_void_p), ("obs", C.c_void_p), ("num_threads", C.c_int32)]


class _Options(C.Structure):
    _fields_ = [("opt_focal", C.c_int32), ("opt_pp", C.c_int32), ("opt_dist", C.c_int32),
                ("max_iterations", C.c_int32), ("trace_capacity", C.c_int32)]


class Summary(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("num_successful", C.c_int32), ("termination", C.c_int32),
                ("n_residual_blocks", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


def build(force: bool = False) -> str:
    if force or not os.path.exists(_LIB_PATH) or os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "lm_oracle.cpp")):
        subprocess.check_call(["make", "-C", _HERE, "liblm_oracle.so"], stdout=subprocess.DEVNULL)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_lm_solve.restype = C.c_int
        _lib.oracle_lm_mono.restype = C.c_int
    return _lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _mk_problem(model_id, mono, obs, pts, width, height, pts_period=None, shared_per_block=False, threads=8):
    obs = np.ascontiguousarray(obs, dtype=np.float32)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    p = _Problem()
    p.model, p.mono = int(model_id), int(bool(mono))
    p.n_blocks, p.obs_per_block = obs.shape[0], obs.shape[1]
    p.pts_period = int(pts_period if pts_period is not None else pts.shape[0])
    p.shared_per_block = int(bool(shared_per_block))
    p.img_w, p.img_h = float(width), float(height)
    p.pts, p.obs = _ptr(pts), _ptr(obs)
    p.num_threads = int(threads)
    return p, obs, pts  # keep arrays alive


def solve(model_id, mono, obs, pts, width, height, focal, pp, dist, poses, opt_focal=True, opt_pp=True, opt_dist=True,
          max_iterations=1000, threads=8, pts_period=None, shared_per_block=False, trace=True):
    """One ceres::Solve restated. Returns (focal, pp, dist, poses, summary dict, trace dict)."""
    p, obs, pts = _mk_problem(model_id, mono, obs, pts, width, height, pts_period, shared_per_block, threads)
    focal = np.array(focal, dtype=np.float64, copy=True)
    pp = np.array(pp, dtype=np.float64, copy=True)
    dist = np.array(dist, dtype=np.float64, copy=True)
    poses = np.array(poses, dtype=np.float64, copy=True)
    o = _Options(int(opt_focal), int(opt_pp), int(opt_dist), int(max_iterations), max_iterations + 2 if trace else 0)
    s = Summary()
    tc = np.zeros(max_iterations + 2)
    tr = np.zeros(max_iterations + 2)
    tf = np.zeros(max_iterations + 2, dtype=np.int32)
    lib().oracle_lm_solve(C.byref(p), C.byref(o), _ptr(focal), _ptr(pp), _ptr(dist), _ptr(poses), C.byref(s),
                          _ptr(tc), _ptr(tr), _ptr(tf))
    n = s.iterations
    return focal, pp, dist, poses, s.as_dict(), dict(cost=tc[:n].copy(), radius=tr[:n].copy(), flag=tf[:n].copy())


def solve_mono(model_id, mono, obs, pts, width, height, focal, pp, dist, poses, max_iterations=1000, threads=8):
    """The reference's 3-stage mono calibration + per-frame RMSE."""
    p, obs, pts = _mk_problem(model_id, mono, obs, pts, width, height, None, False, threads)
    focal = np.array(focal, dtype=np.float64, copy=True)
    pp = np.array(pp, dtype=np.float64, copy=True)
    dist = np.array(dist, dtype=np.float64, copy=True)
    poses = np.array(poses, dtype=np.float64, copy=True)
    rmse = np.zeros(obs.shape[0])
    ss = (Summary * 3)()
    lib().oracle_lm_mono(C.byref(p), _ptr(focal), _ptr(pp), _ptr(dist), _ptr(poses), _ptr(rmse), ss, int(max_iterations))
    return focal, pp, dist, poses, rmse, [s.as_dict() for s in ss]


def frame_rmse(model_id, mono, obs, pts, width, height, focal, pp, dist, poses, pts_period=None, shared_per_block=False):
    p, obs, pts = _mk_problem(model_id, mono, obs, pts, width, height, pts_period, shared_per_block, 1)
    rmse = np.zeros(obs.shape[0])
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (focal, pp, dist, poses)]
    lib().oracle_frame_rmse(C.byref(p), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), _ptr(rmse))
    return rmse


def residual_jacobian(model_id, mono, focal, pp, dist, pose, wpt, u, v):
    """Raw residual (2) and Jacobian 2 x (nf+2+nd+6), columns [focal pp dist pose6]."""
    nf = 1 if mono else 2
    nd = lib().oracle_num_dist(int(model_id))
    D = nf + 2 + nd + 6
    r = np.zeros(2)
    J = np.zeros((2, D))
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (focal, pp, dist, pose, wpt)]
    lib().oracle_residual_jacobian(int(model_id), int(bool(mono)), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), _ptr(a[4]),
                                   C.c_double(float(u)), C.c_double(float(v)), _ptr(r), _ptr(J))
    return r, J


def se3_plus(x, delta):
    out = np.zeros(7)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (x, delta)]
    lib().oracle_se3_plus(_ptr(a[0]), _ptr(a[1]), _ptr(out))
    return out


def project(model_id, mono, focal, pp, dist, pose, wpt):
    out = np.zeros(2)
    a = [np.ascontiguousarray(v, dtype=np.float64) for v in (focal, pp, dist, pose, wpt)]
    lib().oracle_project(int(model_id), int(bool(mono)), _ptr(a[0]), _ptr(a[1]), _ptr(a[2]), _ptr(a[3]), _ptr(a[4]), _ptr(out))
    return out

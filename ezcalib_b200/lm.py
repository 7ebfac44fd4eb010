"""Host-side driver of the GPU Levenberg-Marquardt calibration solve.

Mirrors the structure of EZCalibrator::computeCalibration
(/root/reference/src/ezcalibrator.cpp:105-339): parameter blocks focal / pp / dist shared by all
views plus one SE3 block per view, three staged solves, per-frame RMSE, covariance.
All arithmetic happens in libezcalib_b200.so (CUDA); this file only marshals buffers.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import LmOptionsC, LmProblemC, LmSummaryC, check

TERMINATION = ("function_tolerance", "parameter_tolerance", "gradient_tolerance", "max_iterations",
               "min_trust_region_radius", "failure")


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Context:
    """One context per (process, GPU); all calls are serialised on its stream."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self._h = C.c_void_p()
        check(_lib.lib().ezc_create(int(device), C.byref(self._h)), "ezc_create")
        self.device = device
        self._cb = None
        if stream is not None:
            check(_lib.lib().ezc_set_stream(self._h, C.c_void_p(stream)), "ezc_set_stream")

    @property
    def handle(self):
        return self._h

    def set_allreduce(self, fn, rank: int, world: int):
        """fn(device_ptr:int, count:int, op:int, stream:int) all-reduces `count` doubles in place."""
        def _tramp(user, buf, count, op, stream):
            # an exception inside a ctypes callback would be printed and swallowed: turn it into a status the C side returns
            try:
                fn(int(buf or 0), int(count), int(op), int(stream or 0))
                return 0
            except Exception as e:  # noqa: BLE001
                import sys
                print(f"ezcalib_b200: all-reduce callback failed: {e!r}", file=sys.stderr)
                return 1
        self._cb = _lib.ALLREDUCE_FN(_tramp)
        check(_lib.lib().ezc_set_allreduce(self._h, self._cb, None, int(rank), int(world)), "ezc_set_allreduce")

    def comm_mailbox(self, world: int):
        """Allocate this context's mailbox of the library-owned exchange; returns (device pointer, 64-byte CUDA IPC handle)."""
        ptr = C.c_void_p()
        h = (C.c_uint8 * 64)()
        check(_lib.lib().ezc_comm_mailbox(self._h, int(world), C.byref(ptr), h), "ezc_comm_mailbox")
        return int(ptr.value), bytes(h)

    def comm_init(self, rank: int, world: int, handles):
        """handles: the `world` IPC handles (bytes, 64 each) in rank order, all-gathered by the caller."""
        blob = b"".join(handles)
        assert len(blob) == 64 * world
        buf = (C.c_uint8 * len(blob)).from_buffer_copy(blob)
        check(_lib.lib().ezc_comm_init(self._h, int(rank), int(world), buf), "ezc_comm_init")

    def comm_attach(self, rank: int, world: int, mailboxes):
        """Same-process variant: mailboxes = device pointers (ints) of every rank's mailbox."""
        arr = (C.c_void_p * world)(*[C.c_void_p(int(m)) for m in mailboxes])
        check(_lib.lib().ezc_comm_attach(self._h, int(rank), int(world), arr), "ezc_comm_attach")

    def kernel_launches(self) -> int:
        return int(_lib.lib().ezc_kernel_launches(self._h))

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        check(_lib.lib().ezc_measure_fp64_peak(self._h, C.byref(v)), "ezc_measure_fp64_peak")
        return float(v.value)

    def measure_dmma_peak(self) -> float:
        v = C.c_double()
        check(_lib.lib().ezc_measure_dmma_peak(self._h, C.byref(v)), "ezc_measure_dmma_peak")
        return float(v.value)

    def synchronize(self):
        check(_lib.lib().ezc_synchronize(self._h), "ezc_synchronize")

    def close(self):
        if self._h:
            _lib.lib().ezc_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def default_options(**kw) -> LmOptionsC:
    o = LmOptionsC()
    _lib.lib().ezc_lm_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o


class LmSolver:
    """Device-resident calibration problem (observations + parameter blocks)."""

    def __init__(self, ctx: Context, model_id: int, mono: bool, obs, pts, width: float, height: float,
                 pts_period: int | None = None, cams_per_block: bool = False, device_ptrs: bool = False,
                 n_blocks: int | None = None, obs_per_block: int | None = None, n_blocks_global: int | None = None,
                 collective: int = 0):
        self.ctx = ctx
        p = LmProblemC()
        p.model, p.use_mono_focal = int(model_id), int(bool(mono))
        if device_ptrs:
            # obs / pts are raw device addresses (ints); shapes are given explicitly
            p.n_blocks, p.obs_per_block = int(n_blocks), int(obs_per_block)
            p.pts_period = int(pts_period)
            p.pts, p.obs = C.c_void_p(int(pts)), C.c_void_p(int(obs))
            self._keep = None
        else:
            obs = np.ascontiguousarray(obs, dtype=np.float32)
            pts = np.ascontiguousarray(pts, dtype=np.float64)
            p.n_blocks, p.obs_per_block = obs.shape[0], obs.shape[1]
            p.pts_period = int(pts_period if pts_period is not None else pts.shape[0])
            p.pts, p.obs = _p(pts), _p(obs)
            self._keep = (obs, pts)
        p.cams_per_block = int(bool(cams_per_block))
        p.img_w, p.img_h = float(width), float(height)
        p.n_blocks_global = int(n_blocks_global if n_blocks_global is not None else p.n_blocks)
        p.collective = int(collective)
        self.n_blocks = p.n_blocks
        self.mono = bool(mono)
        self.nf = 1 if mono else 2
        self.nd = _lib.lib().ezc_num_dist(int(model_id))
        self.cams_per_block = bool(cams_per_block)
        self._h = C.c_void_p()
        fn = _lib.lib().ezc_lm_create_device if device_ptrs else _lib.lib().ezc_lm_create
        check(fn(ctx.handle, C.byref(p), C.byref(self._h)), "ezc_lm_create")

    def set_params(self, focal, pp, dist, poses):
        a = [np.ascontiguousarray(x, dtype=np.float64) for x in (focal, pp, dist, poses)]
        check(_lib.lib().ezc_lm_set_params(self._h, _p(a[0]), _p(a[1]), _p(a[2]), _p(a[3])), "ezc_lm_set_params")

    def get_params(self):
        m = self.n_blocks if self.cams_per_block else 1
        focal = np.zeros(m * self.nf)
        pp = np.zeros(m * 2)
        dist = np.zeros(m * self.nd)
        poses = np.zeros((self.n_blocks, 7))
        check(_lib.lib().ezc_lm_get_params(self._h, _p(focal), _p(pp), _p(dist), _p(poses)), "ezc_lm_get_params")
        return focal, pp, dist, poses

    def solve(self, opt_focal=True, opt_pp=True, opt_dist=True, trace: bool = False, **kw):
        """One ceres::Solve equivalent. Returns (summary dict, trace dict|None)."""
        o = default_options(opt_focal=int(opt_focal), opt_pp=int(opt_pp), opt_dist=int(opt_dist), **kw)
        tr = None
        if trace:
            n = o.max_num_iterations + 2
            tc, trr, tf = np.zeros(n), np.zeros(n), np.zeros(n, dtype=np.int32)
            o.trace_capacity = n
            o.trace_cost, o.trace_radius, o.trace_flag = _p(tc), _p(trr), _p(tf)
        s = LmSummaryC()
        check(_lib.lib().ezc_lm_solve(self._h, C.byref(o), C.byref(s)), "ezc_lm_solve")
        d = s.as_dict()
        d["termination_name"] = TERMINATION[d["termination"]]
        if trace:
            m = s.iterations
            tr = dict(cost=tc[:m].copy(), radius=trr[:m].copy(), flag=tf[:m].copy())
        return d, tr

    def frame_rmse(self):
        r = np.zeros(self.n_blocks)
        check(_lib.lib().ezc_lm_frame_rmse(self._h, _p(r)), "ezc_lm_frame_rmse")
        return r

    def covariance(self):
        k = self.nf + 2 + self.nd
        cov = np.zeros((k, k))
        check(_lib.lib().ezc_lm_covariance(self._h, _p(cov)), "ezc_lm_covariance")
        return cov

    def normal_equations(self, opt_focal=True, opt_pp=True, opt_dist=True):
        k = (self.nf if opt_focal else 0) + (2 if opt_pp else 0) + (self.nd if opt_dist else 0)
        if self.cams_per_block:
            k = 0
        nb = self.n_blocks
        Hcc, gc = np.zeros((k, k)), np.zeros(k)
        Hpp, Hpc, gp = np.zeros((nb, 6, 6)), np.zeros((nb, 6, max(k, 1))), np.zeros((nb, 6))
        Hpc_flat = np.zeros(nb * 6 * k) if k else np.zeros(1)
        cost = np.zeros(1)
        check(_lib.lib().ezc_lm_normal_equations(self._h, int(opt_focal), int(opt_pp), int(opt_dist), _p(Hcc), _p(gc),
                                                 _p(Hpp), _p(Hpc_flat), _p(gp), _p(cost)), "ezc_lm_normal_equations")
        Hpc = Hpc_flat[:nb * 6 * k].reshape(nb, 6, k) if k else np.zeros((nb, 6, 0))
        return dict(Hcc=Hcc, gc=gc, Hpp=Hpp, Hpc=Hpc, gp=gp, cost=float(cost[0]))

    def close(self):
        if self._h:
            _lib.lib().ezc_lm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def lm_mono(ctx: Context, model_id: int, mono: bool, obs, pts, width, height, focal, pp, dist, poses):
    """EZCalibrator::computeCalibration's optimisation through the one-shot C entry point
    (host buffers in, host buffers out).  Returns (focal, pp, dist, poses, frame_rmse, summaries)."""
    obs = np.ascontiguousarray(obs, dtype=np.float32)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    p = LmProblemC()
    p.model, p.use_mono_focal = int(model_id), int(bool(mono))
    p.n_blocks, p.obs_per_block, p.pts_period, p.cams_per_block = obs.shape[0], obs.shape[1], pts.shape[0], 0
    p.img_w, p.img_h = float(width), float(height)
    p.pts, p.obs = _p(pts), _p(obs)
    p.n_blocks_global = p.n_blocks
    focal = np.array(focal, dtype=np.float64, copy=True)
    pp = np.array(pp, dtype=np.float64, copy=True)
    dist = np.array(dist, dtype=np.float64, copy=True)
    poses = np.array(poses, dtype=np.float64, copy=True)
    rmse = np.zeros(obs.shape[0])
    ss = (LmSummaryC * 3)()
    check(_lib.lib().ezc_lm_mono(ctx.handle, C.byref(p), _p(focal), _p(pp), _p(dist), _p(poses), _p(rmse), ss), "ezc_lm_mono")
    return focal, pp, dist, poses, rmse, [s.as_dict() for s in ss]


def lm_multicam(ctx: Context, model_id: int, mono: bool, obs, pts, width, height, focal, pp, dist, extrinsics,
                n_blocks_global: int | None = None, collective: int = 0):
    """EZCalibrator::runMultiCameraCalib's optimisation through the one-shot C entry point (ezc_lm_multicam): block b =
    extrinsic of camera b+1 with its own constant intrinsics; obs [n_blocks, n_frames0 * n_pts, 2], pts [n_frames0 * n_pts, 3].
    Returns (extrinsics [n_blocks, 7], summary dict)."""
    obs = np.ascontiguousarray(obs, dtype=np.float32)
    pts = np.ascontiguousarray(pts, dtype=np.float64)
    p = LmProblemC()
    p.model, p.use_mono_focal = int(model_id), int(bool(mono))
    p.n_blocks, p.obs_per_block, p.pts_period, p.cams_per_block = obs.shape[0], obs.shape[1], pts.shape[0], 1
    p.img_w, p.img_h = float(width), float(height)
    p.pts, p.obs = _p(pts), _p(obs)
    p.n_blocks_global = int(n_blocks_global if n_blocks_global is not None else p.n_blocks)
    p.collective = int(collective)
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (focal, pp, dist)]
    ext = np.array(extrinsics, dtype=np.float64, copy=True).reshape(-1, 7)
    s = LmSummaryC()
    check(_lib.lib().ezc_lm_multicam(ctx.handle, C.byref(p), _p(a[0]), _p(a[1]), _p(a[2]), _p(ext), C.byref(s)), "ezc_lm_multicam")
    d = s.as_dict()
    d["termination_name"] = TERMINATION[d["termination"]]
    return ext, d


def residual_jacobian(ctx: Context, model_id: int, mono: bool, focal, pp, dist, pose, wpt, u, v):
    nf = 1 if mono else 2
    nd = _lib.lib().ezc_num_dist(int(model_id))
    D = nf + 2 + nd + 6
    r, J = np.zeros(2), np.zeros((2, D))
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (focal, pp, dist, pose, wpt)]
    check(_lib.lib().ezc_lm_residual_jacobian(ctx.handle, int(model_id), int(bool(mono)), _p(a[0]), _p(a[1]), _p(a[2]),
                                              _p(a[3]), _p(a[4]), float(u), float(v), _p(r), _p(J)), "ezc_lm_residual_jacobian")
    return r, J

"""GPU: csrc/libm_exact.cuh -- the device restatements of glibc's atan2f (fdlibm) and sinf / cosf (ARM optimised
routines) must equal the HOST libm bit for bit: a 1-ulp difference in a pixel's theta flips integer edge costs
(thirdparty/ethz_apriltag2/src/Edge.cc:22-23) and with them the union-find merge order.  The host side is the C library
the reference's own objects (oracle/_ref) were linked against, called through ctypes; the restatement is pinned to that
glibc (2.39 in this image) -- a different libm on the reference side is a different reference."""
import ctypes as C
import ctypes.util

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _host(fn_name, *args):
    libm = C.CDLL(ctypes.util.find_library("m") or "libm.so.6")
    fn = getattr(libm, fn_name)
    fn.restype = C.c_float
    fn.argtypes = [C.c_float] * len(args)
    out = np.empty(len(args[0]), np.float32)
    for i in range(len(out)):
        out[i] = fn(*(float(a[i]) for a in args))
    return out


def _device(ctx, which, a, b=None):
    from ezcalib_b200 import _lib
    a = np.ascontiguousarray(a, np.float32)
    out = np.empty_like(a)
    bp = None if b is None else np.ascontiguousarray(b, np.float32)
    _lib.check(_lib.lib().ezc_debug_libm(ctx.handle, which, a.ctypes.data_as(C.c_void_p), None if bp is None else bp.ctypes.data_as(C.c_void_p),
                                         a.size, out.ctypes.data_as(C.c_void_p)), "ezc_debug_libm")
    return out


def _args(rng, n):
    # gradient-like magnitudes (|Ix|, |Iy| <= 2), tiny values, exact zeros and signs, plus wide-range values
    a = np.concatenate([rng.uniform(-2, 2, n // 2), rng.normal(0, 1e-3, n // 4), rng.normal(0, 50.0, n // 4 - 16),
                        [0.0, -0.0, 1.0, -1.0, 1e-30, -1e-30, 3.0e38, -3.0e38, 0.5, 2.0, 1e-10, 7.0, 0.0, 0.0, 1.0, -1.0]]).astype(np.float32)
    return a


def test_atan2f_sinf_cosf_bit_exact_vs_host_libm(ctx):
    rng = np.random.default_rng(9)
    n = 400_000
    y, x = _args(rng, n), rng.permutation(_args(rng, n))
    got = _device(ctx, 0, y, x)
    ref = _host("atan2f", y, x)
    assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), int((got.view(np.uint32) != ref.view(np.uint32)).sum())
    t = np.concatenate([rng.uniform(-7, 7, n // 2), rng.uniform(-200, 200, n // 2)]).astype(np.float32)
    for which, name in ((1, "sinf"), (2, "cosf")):
        got = _device(ctx, which, t)
        ref = _host(name, t)
        assert np.array_equal(got.view(np.uint32), ref.view(np.uint32)), (name, int((got.view(np.uint32) != ref.view(np.uint32)).sum()))

"""The C-ABI shared library loads and exports every symbol include/ezcalib_b200.h declares
(no compute call: this runs without a GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "ezcalib_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(ezc_[a-z0-9_]+)\s*\(", txt)) - {"ezc_allreduce_fn"})


def test_library_exports_every_declared_symbol():
    from ezcalib_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    L = ctypes.CDLL(_lib.LIB_PATH)
    syms = _declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), f"missing export {s}"
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms


def test_no_gpu_fails_loudly():
    """Without a CUDA device every compute entry point must fail (no CPU fallback)."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from ezcalib_b200 import _lib
    from ezcalib_b200.lm import Context
    with pytest.raises(_lib.EzcError, match="no CUDA device|CPU fallback|cuda"):
        Context(0)

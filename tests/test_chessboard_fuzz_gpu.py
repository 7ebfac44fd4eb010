"""GPU: randomised chessboard views against cv2 (the reference's flags, cv2 without IPP): random pose incl. large in-plane
rotation, lens distortion, blur up to sigma 2.4, illumination gradients, occluding discs, sensor noise up to sigma 10
(tests/cb_fuzz.py).  hypothesis draws the batch seeds (derandomised: the same examples every run), 16 batches of 64 views.

What is asserted is what was measured after the fallback-compensation fix of round 2, with head-room of one batch: the
restatement of cv::findChessboardCorners is empirical (OpenCV's source is not in the image), and on these HARD views a
small residue remains -- about 3 views per 1000 where cv2 and the restatement succeed at different attempts (DESIGN.md
section 4b lists what was learnt from them).  On the BASELINE-style renders (tests/test_baseline_configs_gpu.py) the
agreement is bit for bit."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import cb_fuzz  # noqa: E402

pytestmark = pytest.mark.gpu
hypothesis = pytest.importorskip("hypothesis")
from hypothesis import given, settings, strategies as st  # noqa: E402

STATS = dict(views=0, found=0, flag_mismatch=0, over=0, worst=0.0)


@settings(max_examples=16, deadline=None, derandomize=True, database=None)
@given(seed=st.integers(min_value=2000, max_value=2_000_000))
def test_fuzz_batch_matches_cv2(ctx, seed):
    pytest.importorskip("cv2")
    views = cb_fuzz.make_views(ctx, seed, 64)
    n_ok, bad, errs = cb_fuzz.compare(ctx, views)
    STATS["views"] += len(views); STATS["found"] += n_ok; STATS["flag_mismatch"] += len(bad)
    STATS["over"] += int((errs > 0.01).sum()); STATS["worst"] = max(STATS["worst"], float(errs.max()) if len(errs) else 0.0)
    # a single batch may hold a residual view; more than three means a regression
    assert len(bad) + int((errs > 0.01).sum()) <= 3, (seed, bad, errs[errs > 0.01])


def test_fuzz_totals(ctx):
    """Runs after the batches (file order): totals over all drawn batches."""
    s = STATS
    if s["views"] == 0:
        pytest.skip("no batches ran")
    print(f"chessboard fuzz: {s['views']} views, cv2 found {s['found']}, found-flag mismatches {s['flag_mismatch']}, "
          f"views with a corner off by > 0.01 px {s['over']} (worst {s['worst']:.3f} px)")
    assert s["flag_mismatch"] <= 0.01 * s["views"] and s["over"] <= 0.015 * s["views"]

"""CPU: the cornerSubPix restatement (oracle/subpix_oracle.py) against the real OpenCV.
Golden vectors were produced by cv2 4.13.0 (tests/golden/make_subpix_golden.py)."""
import os

import numpy as np
import pytest

from oracle import subpix_oracle as S

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "subpix_golden.npz"))


@pytest.mark.parametrize("case", [0, 1])
def test_oracle_matches_cv2_golden(case):
    img, start = G[f"img{case}"], G[f"start{case}"]
    out = S.corner_subpix(img, start, (5, 5), 100, 0.05)
    assert np.abs(out - G[f"ref_win5_{case}"]).max() <= 1e-4   # cv2 (IPP sampler) vs restatement: ~1e-5 px
    out2 = S.corner_subpix(img, start, (2, 2), 15, 0.1)
    assert np.abs(out2 - G[f"ref_win2_{case}"]).max() <= 1e-4


def test_oracle_bit_exact_vs_live_cv2_without_ipp():
    cv2 = pytest.importorskip("cv2")
    img, start = G["img1"], G["start1"]
    cv2.ipp.setUseIPP(False)
    try:
        ref = cv2.cornerSubPix(img, start.copy().reshape(-1, 1, 2), (5, 5), (-1, -1),
                               (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)).reshape(-1, 2)
    finally:
        cv2.ipp.setUseIPP(True)
    out = S.corner_subpix(img, start, (5, 5), 100, 0.05)
    assert np.array_equal(out, ref)

"""CPU: the AprilGrid oracle = the reference's own sources (oracle/_ref) must reproduce the committed golden
detections (made by tests/golden/make_aprilgrid_golden.py); also cross-checks the code table against cv2.aruco."""
import os

import numpy as np
import pytest

from oracle import apriltag_oracle as A

G = np.load(os.path.join(os.path.dirname(__file__), "golden", "aprilgrid_golden.npz"))
pytestmark = pytest.mark.skipif(not A.available(), reason="oracle/_ref not built and /root/reference absent")


@pytest.mark.parametrize("name", [str(n) for n in G["names"]])
def test_reference_detector_reproduces_golden(name):
    img = G[name + "_img"]
    rows, cols = (int(v) for v in G[name + "_grid"])
    ok, corners, n, det = A.detect_target(img, rows, cols)
    assert bool(ok) == bool(G[name + "_ok"]) and int(n) == int(G[name + "_n"])
    assert np.array_equal(det["id"], G[name + "_ids"])
    assert np.array_equal(det["p"], G[name + "_p"])          # raw quad corners: bit-exact
    assert np.abs(corners - G[name + "_corners"]).max() <= 1e-4  # after cornerSubPix (cv2 build may differ in IPP use)


def test_code_table_equals_aruco_dictionary_rotated():
    cv2 = pytest.importorskip("cv2")
    if not hasattr(cv2, "aruco"):
        pytest.skip("cv2.aruco missing")
    from ezcalib_b200.tag36h11 import CODES
    assert np.array_equal(CODES, A.codes36h11())
    d = cv2.aruco.getPredefinedDictionary(cv2.aruco.DICT_APRILTAG_36h11)
    for tid in (0, 1, 17, 300, 586):
        bits = cv2.aruco.Dictionary.getBitsFromByteList(d.bytesList[tid:tid + 1], 6).astype(np.uint8)
        ref = np.array([(int(CODES[tid]) >> (35 - i)) & 1 for i in range(36)], np.uint8).reshape(6, 6)
        assert np.array_equal(np.rot90(bits, 2), ref)  # aruco grid is the reference grid rotated 180 deg

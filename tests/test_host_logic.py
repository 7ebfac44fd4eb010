"""CPU: host-side logic that mirrors the reference's orchestration (no GPU needed)."""
import numpy as np

from ezcalib_b200 import ezcalibrator as E
from ezcalib_b200 import synth


def _reference_loop(found):
    """Literal transcription of the iterator arithmetic of ezcalibrator.cpp:39-59 (bounded instead of UB)."""
    used, it, bad, n = [], 0, 0, len(found)
    while it < n:
        used.append(it)
        if found[it]:
            bad = 0
        else:
            bad += 1
            for _ in range(bad // 2):
                it += 1
        it += 1
    return used


def test_skip_rule_replay():
    rng = np.random.default_rng(0)
    for _ in range(200):
        f = rng.uniform(size=rng.integers(1, 60)) < rng.uniform(0.2, 0.95)
        assert E.replay_skip_rule(list(f)) == _reference_loop(list(f))
    assert E.replay_skip_rule([False] * 8) == [0, 1, 3, 5, 8][:4]  # 0, 1 (skip 1), 3 (skip 1), 5 (skip 2) ...


def test_se3_log_exp_roundtrip():
    rng = np.random.default_rng(1)
    for _ in range(50):
        d = np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 0.8, 3)])
        T = synth.se3_exp_left(d, np.array([0, 0, 0, 1.0, 0, 0, 0]))
        assert np.allclose(synth.se3_log(T), d, atol=1e-10)
        I = synth.se3_mul(T, synth.se3_inverse(T))
        assert np.allclose(I, [0, 0, 0, 1, 0, 0, 0], atol=1e-12)


def test_target_coordinates_follow_the_reference():
    t = synth.chessboard_target(7, 10, 0.05)
    assert t.shape == (54, 3) and np.allclose(t[0], [-0.45, 0.05, 1.0]) and np.allclose(t[1], [-0.40, 0.05, 1.0])
    assert np.allclose(t[9], [-0.45, 0.10, 1.0]) and np.all(t[:, 2] == 1.0)
    a = synth.aprilgrid_target(6, 6, 0.06, 0.3)
    assert a.shape == (144, 3)
    assert np.allclose(a[:4], [[0, 0, 1], [0.06, 0, 1], [0.06, -0.06, 1], [0, -0.06, 1]])
    assert np.allclose(a[4], [0.078, 0, 1]) and np.allclose(a[24], [0, -0.078, 1])


def test_prior_intrinsics_rule():
    f, cx, cy = synth.prior_intrinsics(1280, 1024, 70.0)
    assert cx == 640 and cy == 512 and abs(f - 640 / np.tan(np.radians(35.0))) < 1e-9

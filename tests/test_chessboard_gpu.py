"""GPU: batched chessboard detection against the REAL OpenCV (cv2 is present on the GPU box too) and the committed
golden vectors: found flag and corner ORDER bit-exact, corners after the reference's cornerSubPix within 0.01 px on
every fixture (north_star tolerance), against cv2 without IPP (the OpenCV contract)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "chessboard_golden.npz"))
NAMES = [str(n) for n in G["names"]]
BIG = [n for n in NAMES if G[n + "_img"].shape == (480, 640)]


def test_batch_raw_corners_match_cv2_golden(ctx):
    from ezcalib_b200 import detect
    imgs = detect.Images(ctx, np.stack([G[n + "_img"] for n in BIG]))
    found, corners = detect.find_chessboard_corners(ctx, imgs, (9, 6), final_subpix=False)
    imgs.close()
    for k, n in enumerate(BIG):
        assert bool(found[k]) == bool(G[n + "_found"]), n
        if found[k]:
            d = np.abs(corners[k] - G[n + "_raw"]).max(axis=1)
            assert (d <= 1e-4).mean() >= 0.9, (n, d.max())
            if n.startswith("render"):
                assert d.max() <= 1e-4, (n, d.max())


def test_detect_target_matches_reference_calls(ctx):
    """ChessboardCalibDetector::detectTarget == findChessboardCorners + cornerSubPix(5,5): live cv2 on the same images."""
    cv2 = pytest.importorskip("cv2")
    from ezcalib_b200 import detect
    flags = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)
    imgs = detect.Images(ctx, np.stack([G[n + "_img"] for n in BIG]))
    found, corners = detect.detect_chessboard(ctx, imgs, 7, 10)   # 7 x 10 squares -> 9 x 6 inner corners
    imgs.close()
    for k, n in enumerate(BIG):
        img = G[n + "_img"]
        ok, c = cv2.findChessboardCorners(img, (9, 6), flags=flags)
        assert bool(ok) == bool(found[k]), n
        if ok:
            ref = cv2.cornerSubPix(img, c, (5, 5), (-1, -1), crit).reshape(-1, 2)
            tol = 0.01   # every fixture, degraded ones included (round 1 allowed 0.05 px on those)
            assert np.abs(corners[k] - ref).max() <= tol, (n, np.abs(corners[k] - ref).max())
            assert np.abs(corners[k] - G[n + "_final"]).max() <= tol
        else:
            assert np.all(corners[k] == 0)


def test_small_noise_image_is_rejected(ctx):
    from ezcalib_b200 import detect
    imgs = detect.Images(ctx, G["noise_img"])
    found, corners = detect.detect_chessboard(ctx, imgs, 7, 10)
    imgs.close()
    assert not found[0]


def test_erased_outer_squares_match_cv2(ctx):
    """Boards with outer squares painted over, one batch: the detector must repair (or reject) exactly like cv2."""
    import sys
    import cv2
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "cb_harness"))
    from erased import FLAGS, erased_cases
    from ezcalib_b200 import detect
    batch, labels = [], []
    for name in ("render0", "warp0", "warp2"):
        for cs, im in erased_cases(G[name + "_img"], n_pairs=4, step=1):
            batch.append(im); labels.append((name, cs))
    imgs = detect.Images(ctx, np.stack(batch))
    found, corners = detect.find_chessboard_corners(ctx, imgs, (9, 6), False)
    imgs.close()
    n_found = 0
    for k, im in enumerate(batch):
        ok_ref, c_ref = cv2.findChessboardCorners(im, (9, 6), flags=FLAGS)
        assert bool(found[k]) == bool(ok_ref), labels[k]
        if ok_ref:
            n_found += 1
            assert np.abs(corners[k] - c_ref.reshape(-1, 2)).max() <= 0.01, labels[k]
    assert n_found >= 40 and n_found < len(batch)   # both outcomes are exercised


def test_predilated_ideal_boards_fail_at_cv2s_level(ctx):
    """Ideal boards dilated m times before the call, levels around cv2's boundary: found flags must equal cv2's (pins the
    undilated first attempt of the histogram phase and the linking rule's distance bound)."""
    import sys
    import cv2
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "cb_harness"))
    from erased import FLAGS, ideal_board, predilated
    from ezcalib_b200 import detect
    batch, labels = [], []
    for (S, shear, ry) in ((30, 0.0, 1.0), (43, 0.0026, 1.0), (50, 0.0, 0.7), (60, 0.0, 1.0)):
        for m, im in predilated(ideal_board(S, shear, ry=ry), range(int(S * ry) // 5, int(S * ry) // 3 + 2)):
            batch.append(im); labels.append((S, shear, ry, m))
    imgs = detect.Images(ctx, np.stack(batch))
    found, corners = detect.find_chessboard_corners(ctx, imgs, (9, 6), False)
    imgs.close()
    seen = set()
    for k, im in enumerate(batch):
        ok_ref, c_ref = cv2.findChessboardCorners(im, (9, 6), flags=FLAGS)
        assert bool(found[k]) == bool(ok_ref), labels[k]
        seen.add(bool(ok_ref))
        if ok_ref:
            assert np.abs(corners[k] - c_ref.reshape(-1, 2)).max() <= 0.01, labels[k]
    assert seen == {True, False}


def test_unaligned_width_matches_cv2(ctx):
    """Image widths that are not multiples of 16 take the byte-load path of the 16-pixel row chunks: same answers."""
    import cv2
    from ezcalib_b200 import detect
    cv2.ipp.setUseIPP(False)
    FLAGS = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
    n_found = 0
    for name, cut in (("render0", 3), ("warp2", 9), ("render1", 13)):
        img = np.ascontiguousarray(G[name + "_img"][:-1, :-cut])
        assert img.shape[1] % 16 != 0
        imgs = detect.Images(ctx, img[None])
        found, corners = detect.find_chessboard_corners(ctx, imgs, (9, 6), True, fast_check=True)
        imgs.close()
        ok_ref, c_ref = cv2.findChessboardCorners(img, (9, 6), flags=FLAGS)
        assert bool(found[0]) == bool(ok_ref), name
        if ok_ref:
            c_ref = cv2.cornerSubPix(img, c_ref, (5, 5), (-1, -1), (cv2.TERM_CRITERIA_EPS + cv2.TERM_CRITERIA_COUNT, 100, 0.05))
            assert np.abs(corners[0] - c_ref.reshape(-1, 2)).max() <= 0.01, name
            n_found += 1
    assert n_found >= 2


def test_occluding_discs_match_cv2(ctx):
    """A gray disc painted over the first corner (darker than, equal to and brighter than the binarisation threshold), at
    the fixture size and upscaled 2x: found flags and corners must equal cv2's with the reference's flags."""
    import cv2
    from ezcalib_b200 import detect
    cv2.ipp.setUseIPP(False)
    FLAGS = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
    outcomes = set()
    for scale in (1, 2):
        batch, labels = [], []
        for name in ("render0", "warp2"):
            base = G[name + "_img"]
            ok, c = cv2.findChessboardCorners(base, (9, 6))
            assert ok
            x, y = c.reshape(-1, 2)[0]
            for val in (100, 128, 160, 200):
                img = base.copy()
                cv2.circle(img, (int(x), int(y)), 40, val, -1)
                if scale != 1:
                    img = cv2.resize(img, None, fx=scale, fy=scale, interpolation=cv2.INTER_LINEAR)
                batch.append(img); labels.append((name, val, scale))
        imgs = detect.Images(ctx, np.stack(batch))
        found, corners = detect.find_chessboard_corners(ctx, imgs, (9, 6), False, fast_check=True)
        imgs.close()
        for k, img in enumerate(batch):
            ok_ref, c_ref = cv2.findChessboardCorners(img, (9, 6), flags=FLAGS)
            assert bool(found[k]) == bool(ok_ref), labels[k]
            outcomes.add(bool(ok_ref))
            if ok_ref:
                assert np.abs(corners[k] - c_ref.reshape(-1, 2)).max() <= 0.01, labels[k]
    assert False in outcomes   # the occluded board is rejected somewhere: the full 56-attempt search runs


def test_noisy_view_at_seven_dilations_is_rejected_like_cv2(ctx):
    """cv2 does not find this sigma-10 view: at 7 dilations it links a blob to its side neighbour (one-sided diagonal test
    in attempts with dilation compensation) and the grid comes out scrambled.  Same decision here."""
    import cv2
    from ezcalib_b200 import detect
    FLAGS = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
    for name in ("chessboard_noisy_d7.npz", "chessboard_noisy_fallback_d6.npz"):
        noisy = np.load(os.path.join(os.path.dirname(__file__), "golden", name))["img"]
        ok_ref, _ = cv2.findChessboardCorners(noisy, (9, 6), flags=FLAGS)
        imgs = detect.Images(ctx, noisy[None])
        found, _ = detect.find_chessboard_corners(ctx, imgs, (9, 6), True, fast_check=True)
        imgs.close()
        assert not ok_ref and not found[0], name


def test_fast_check_gate_matches_cv2(ctx):
    """CALIB_CB_FAST_CHECK on: found flags must equal cv2's with the reference's full flag set -- including the shrunk
    boards that the gate rejects although the search alone finds them -- and the gate must not change found corners."""
    import sys
    import cv2
    sys.path.insert(0, os.path.join(os.path.dirname(__file__), "cb_harness"))
    from erased import FLAGS, gate_corpus
    from ezcalib_b200 import detect
    batch = gate_corpus(G)
    imgs = detect.Images(ctx, np.stack(batch))
    f_gate, c_gate = detect.find_chessboard_corners(ctx, imgs, (9, 6), False, fast_check=True)
    f_full, c_full = detect.find_chessboard_corners(ctx, imgs, (9, 6), False, fast_check=False)
    imgs.close()
    changed = 0
    cv2.ipp.setUseIPP(False)   # the generic OpenCV code path (IPP's sampler moves weak corners of blurred views by > 1 px)
    n_close = n_found = 0
    for k, im in enumerate(batch):
        ok_gate, c_ref = cv2.findChessboardCorners(im, (9, 6), flags=FLAGS + cv2.CALIB_CB_FAST_CHECK)
        ok_full, _ = cv2.findChessboardCorners(im, (9, 6), flags=FLAGS)
        assert bool(f_gate[k]) == bool(ok_gate), k
        assert bool(f_full[k]) == bool(ok_full), k
        changed += int(bool(ok_gate) != bool(ok_full))
        if ok_gate:
            n_found += 1
            n_close += int(np.abs(c_gate[k] - c_ref.reshape(-1, 2)).max() <= 0.01)
            assert np.array_equal(c_gate[k], c_full[k])   # the gate never changes a found board
    assert changed >= 1           # the corpus contains boards only the gate rejects
    # identical attempt and quads on every view, incl. the 0.3-scale boards found in the fallback phase
    assert n_close == n_found, (n_close, n_found)


def _detect_both_ways(ctx, stack, **kw):
    """(found, corners) of the concurrent-attempt search and of the one-at-a-time search (EZC_CB_SEQUENTIAL=1)."""
    from ezcalib_b200 import detect
    out = []
    for sequential in (False, True):
        if sequential:
            os.environ["EZC_CB_SEQUENTIAL"] = "1"
        try:
            imgs = detect.Images(ctx, stack)
            out.append(detect.detect_chessboard(ctx, imgs, 7, 10, **kw))
            imgs.close()
        finally:
            os.environ.pop("EZC_CB_SEQUENTIAL", None)
    return out


def test_concurrent_attempts_equal_the_sequential_search(ctx):
    """The remaining attempts of unfinished views run concurrently in free slots (chessboard.cu, detect_sub_batch); the
    walk over their results must keep exactly what the one-at-a-time search keeps: same flags, same corners bit for bit.
    Golden fixtures (incl. the noise view that fails all 56 attempts), degraded fuzz views, and a batch small enough that
    the free slots run out (several rounds, chunks that end in the middle of a threshold level)."""
    import sys
    sys.path.insert(0, os.path.dirname(__file__))
    import cb_fuzz
    stacks = [np.stack([G[n + "_img"] for n in BIG]),
              cb_fuzz.make_views(ctx, 1001, 64), cb_fuzz.make_views(ctx, 1002, 64),
              cb_fuzz.make_views(ctx, 1003, 64)[:6], cb_fuzz.make_views(ctx, 1001, 64)[40:45]]
    for stack in stacks:
        (f_con, c_con), (f_seq, c_seq) = _detect_both_ways(ctx, stack)
        assert np.array_equal(f_con, f_seq), np.flatnonzero(f_con != f_seq)
        assert np.array_equal(c_con, c_seq), float(np.abs(c_con - c_seq).max())
    # other slot caps (EZC_CB_SLOTS): rounds that end in the middle of a threshold level / one round for everything
    for slots in ("3", "400"):
        os.environ["EZC_CB_SLOTS"] = slots
        try:
            (f_s, c_s), _ = _detect_both_ways(ctx, stacks[2])
        finally:
            os.environ.pop("EZC_CB_SLOTS", None)
        (f_d, c_d), _ = _detect_both_ways(ctx, stacks[2])
        assert np.array_equal(f_s, f_d) and np.array_equal(c_s, c_d), slots
    # sub-batches of 1 and 3 views: no free slots at all / a few -- the per-view results do not depend on the slot count
    stack = stacks[1][:9]
    (f_ref, c_ref), _ = _detect_both_ways(ctx, stack)
    for m in (1, 3):
        (f_con, c_con), (f_seq, c_seq) = _detect_both_ways(ctx, stack, max_images_in_flight=m)
        assert np.array_equal(f_con, f_seq) and np.array_equal(c_con, c_seq), m
        assert np.array_equal(f_con, f_ref) and np.array_equal(c_con, c_ref), m

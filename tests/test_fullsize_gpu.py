"""GPU: BASELINE.json's full sizes through size-independent properties (the oracles finish only small cases in seconds):
determinism, additivity over a split of the views, sharded == unsharded, monotone cost on accepted steps for the LM solve
at 10,000,080 observations; batch-composition invariance and reference parity on a sample for the detectors at
2448x2048 (cfg3) and 1920x1080 (cfg2)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_lm_full_size_properties(ctx):
    import bench
    from ezcalib_b200 import lm, synth
    model = "radtan5"
    mid = synth.MODEL_ID[model]
    n = bench.N_VIEWS_10M
    P = bench.make_problem(model, n, seed_offset=9)
    assert P.obs.shape[0] * P.obs.shape[1] == 10_000_080

    def solve(obs, poses, iters):
        s = lm.LmSolver(ctx, mid, True, obs, P.target, P.width, P.height)
        s.set_params(P.focal_init, P.pp_init, P.dist_init, poses)
        summ, tr = s.solve(max_num_iterations=iters, trace=True)
        out = s.get_params()
        s.close()
        return summ, out, tr["cost"], tr["flag"]

    s1, o1, c1, f1 = solve(P.obs, P.poses_init, 6)
    s2, o2, c2, f2 = solve(P.obs, P.poses_init, 6)
    # determinism: fixed-order reductions => bit-identical runs
    assert s1["final_cost"] == s2["final_cost"]
    for a, b in zip(o1, o2):
        assert np.array_equal(a, b)
    # accepted steps (flag 1) never increase the cost
    k = s1["iterations"]
    acc = [c1[i] for i in range(k) if i == 0 or f1[i] == 1]
    assert all(acc[i + 1] <= acc[i] for i in range(len(acc) - 1)) and acc[-1] < 0.5 * acc[0]
    # additivity: the cost at the starting point is the sum over any split of the views
    h = n // 2
    sa, *_ = solve(P.obs[:h], P.poses_init[:h], 1)
    sb, *_ = solve(P.obs[h:], P.poses_init[h:], 1)
    assert abs(sa["initial_cost"] + sb["initial_cost"] - s1["initial_cost"]) <= 1e-12 * s1["initial_cost"]
    assert sa["n_residual_blocks"] + sb["n_residual_blocks"] == s1["n_residual_blocks"]


def test_aprilgrid_full_size_batch_invariance_and_reference(ctx):
    from ezcalib_b200 import detect, synth
    from oracle import apriltag_oracle as A
    W, H, n = 2448, 2048, 6
    model = "radtan8"; mid = synth.MODEL_ID[model]; d = synth.GT_DIST[model]
    f, cx, cy = 2000.0, W / 2 + 3.0, H / 2 - 4.0
    tgt = synth.aprilgrid_target(6, 6, 0.06, 0.3)
    rng = np.random.Generator(np.random.Philox(key=31))
    poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=30, tilt_sigma=0.3)
    cams = np.tile(np.array([f, f, cx, cy] + list(d)), (n, 1))
    imgs = detect.render_targets(ctx, "aprilgrid", 6, 6, 0.06, 0.3, mid, W, H, cams, poses, seed=5)
    a = detect.detect_aprilgrid(ctx, imgs, 6, 6)
    b = detect.detect_aprilgrid(ctx, imgs, 6, 6, max_images_in_flight=1)
    for x, y in zip(a, b):
        assert np.array_equal(x, y)     # an image's result does not depend on what else is in flight
    host = imgs.download(0, 2)
    imgs.close()
    for i in range(2):
        ok, c, nd, _ = A.detect_target(host[i], 6, 6)
        assert bool(ok) == bool(a[0][i]) and int(nd) == int(a[2][i])
        assert np.array_equal(c[:, 0] < 0, a[1][i][:, 0] < 0) and np.abs(c - a[1][i]).max() <= 0.01
        # ground truth: the refined corners sit on the projected target corners
        gt = synth.project(model, f, f, cx, cy, d, poses[i], tgt)
        m = a[1][i][:, 0] >= 0
        assert m.sum() >= 100 and np.linalg.norm(a[1][i][m] - gt[m], axis=1).max() < 0.6   # rendered with blur + noise: sub-pixel, not exact


def test_chessboard_full_size_batch_invariance_and_cv2(ctx):
    cv2 = pytest.importorskip("cv2")
    cv2.ipp.setUseIPP(False)
    from ezcalib_b200 import detect, synth
    W, H, n = 1920, 1080, 8
    model = "kannalabrandt"; mid = synth.MODEL_ID[model]; d = synth.GT_DIST[model]
    f, cx, cy = 900.0, 965.0, 533.0
    tgt = synth.chessboard_target(7, 10, 0.05)
    rng = np.random.Generator(np.random.Philox(key=32))
    poses = synth.sample_poses(rng, n, tgt, model, f, cx, cy, d, W, H, margin=90, tilt_sigma=0.3)
    cams = np.tile(np.array([f, f, cx, cy] + list(d) + [0.0] * (8 - len(d))), (n, 1))
    imgs = detect.render_targets(ctx, "chessboard", 7, 10, 0.05, 0.0, mid, W, H, cams, poses, seed=6)
    fa, ca = detect.detect_chessboard(ctx, imgs, 7, 10)
    fb, cb = detect.detect_chessboard(ctx, imgs, 7, 10, max_images_in_flight=3)
    assert np.array_equal(fa, fb) and np.array_equal(ca, cb)
    host = imgs.download()
    imgs.close()
    flags = cv2.CALIB_CB_ADAPTIVE_THRESH + cv2.CALIB_CB_NORMALIZE_IMAGE + cv2.CALIB_CB_FILTER_QUADS + cv2.CALIB_CB_FAST_CHECK
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 100, 0.05)
    n_found = 0
    for i in range(n):
        ok, c = cv2.findChessboardCorners(host[i], (9, 6), flags=flags)
        assert bool(ok) == bool(fa[i]), i
        if ok:
            n_found += 1
            c = cv2.cornerSubPix(host[i], c, (5, 5), (-1, -1), crit).reshape(-1, 2)
            assert np.abs(c - ca[i]).max() <= 0.01, i
    assert n_found >= n - 2

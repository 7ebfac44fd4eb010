"""CPU: sanity of the LM oracle itself (oracle/lm_oracle.cpp) -- Jet autodiff Jacobians against central finite
differences, SE3 manifold properties, and the early-stopped 3-stage solve against the true minimiser (scipy)."""
import numpy as np
import pytest

from ezcalib_b200 import synth
from oracle import lm_oracle as O


@pytest.mark.parametrize("model", synth.MODELS)
@pytest.mark.parametrize("mono", [True, False])
def test_jet_jacobian_vs_finite_differences(model, mono):
    rng = np.random.default_rng(5)
    mid = synth.MODEL_ID[model]
    nd = synth.NUM_DIST[mid]
    nf = 1 if mono else 2
    focal = np.array([820.0] if mono else [820.0, 835.0])
    pp = np.array([640.0, 500.0])
    dist = np.array(synth.GT_DIST[model])
    pose = np.concatenate([synth.quat_from_rotvec(0.3 * rng.standard_normal(3)), [0.04, -0.02, 0.8]])
    wpt = np.array([0.15, -0.1, 1.0])
    r0, J = O.residual_jacobian(mid, mono, focal, pp, dist, pose, wpt, 600.0, 480.0)

    def res(f, p, d, T):
        return O.project(mid, mono, f, p, d, T, wpt) - np.array([600.0, 480.0])
    assert np.allclose(res(focal, pp, dist, pose), r0, atol=1e-10)
    h = 1e-6
    col = 0
    for arr in (focal, pp, dist):
        for i in range(len(arr)):
            a, b = arr.copy(), arr.copy()
            a[i] += h; b[i] -= h
            args_a = [focal, pp, dist]; args_b = [focal, pp, dist]
            k = [id(focal), id(pp), id(dist)].index(id(arr))
            args_a[k] = a; args_b[k] = b
            fd = (res(*args_a, pose) - res(*args_b, pose)) / (2 * h)
            assert np.allclose(J[:, col], fd, rtol=1e-6, atol=1e-6), (model, mono, col)
            col += 1
    for i in range(6):
        d6 = np.zeros(6); d6[i] = h
        fd = (res(focal, pp, dist, O.se3_plus(pose, d6)) - res(focal, pp, dist, O.se3_plus(pose, -d6))) / (2 * h)
        assert np.allclose(J[:, nf + 2 + nd + i], fd, rtol=1e-6, atol=1e-5), (model, mono, "pose", i)


def test_se3_plus_matches_numpy_exp():
    rng = np.random.default_rng(2)
    for _ in range(20):
        T = np.concatenate([synth.quat_from_rotvec(rng.standard_normal(3)), rng.standard_normal(3)])
        d = np.concatenate([0.1 * rng.standard_normal(3), 0.2 * rng.standard_normal(3)])
        assert np.allclose(O.se3_plus(T, d), synth.se3_exp_left(d, T), atol=1e-12)
    T = np.array([0, 0, 0, 1.0, 1, 2, 3])
    assert np.allclose(O.se3_plus(T, np.zeros(6)), T)


def test_three_stage_solve_reaches_the_true_minimum():
    """The restated Ceres schedule stops early by design (function_tolerance 1e-6); its end point must still sit at
    the minimiser of 1/2 sum log(1+|r|^2) found by scipy at tight tolerance."""
    scipy_opt = pytest.importorskip("scipy.optimize")
    model, mono = "radtan5", True
    mid = synth.MODEL_ID[model]
    tgt = synth.chessboard_target(7, 10, 0.05)
    P = synth.make_lm_problem(model, mono, 12, tgt, 1280, 1024, seed=4)
    f, pp, d, poses, rmse, ss = O.solve_mono(mid, mono, P.obs, P.target, P.width, P.height, P.focal_init, P.pp_init, P.dist_init, P.poses_init)
    assert [s["termination"] for s in ss] == [0, 0, 0]
    assert ss[2]["final_cost"] < ss[1]["final_cost"] < ss[0]["final_cost"] < ss[0]["initial_cost"]
    assert abs(f[0] - P.gt["focal"][0]) < 1.5 and np.median(rmse) < 0.13

    def fun(x):
        fo, p2, di = x[:1], x[1:3], x[3:8]
        out = []
        for b in range(12):
            T = O.se3_plus(poses[b], x[8 + 6 * b: 14 + 6 * b])
            px = synth.project(model, fo[0], fo[0], p2[0], p2[1], di, T, P.target)
            out.append((px - P.obs[b]).ravel())
        return np.concatenate(out)
    x0 = np.concatenate([f, pp, d, np.zeros(72)])
    sol = scipy_opt.least_squares(fun, x0, loss="cauchy", f_scale=1.0, xtol=1e-12, ftol=1e-12, gtol=1e-12)
    # scipy's cost = 1/2 sum rho(r_i^2) per scalar residual; compare parameters instead of costs
    assert abs(sol.x[0] - f[0]) < 0.1 and np.max(np.abs(sol.x[1:3] - pp)) < 0.1


# ------------------------------------------------------------------------------------------------------------------
# Pin of the restated functors: the reference's OWN headers compiled unmodified (oracle/lm_ref -> oracle/_ref/liblm_ref.so)
from oracle import lm_ref as R  # noqa: E402

needs_ref = pytest.mark.skipif(not R.available(), reason="oracle/_ref/liblm_ref.so not built and /root/reference absent")


def _random_case(rng, model, mono):
    focal = np.array([820.0] if mono else [820.0, 835.0]) * (1 + 0.05 * rng.standard_normal())
    pp = np.array([640.0, 500.0]) + 5 * rng.standard_normal(2)
    dist = np.array(synth.GT_DIST[model]) * (1 + 0.2 * rng.standard_normal(len(synth.GT_DIST[model])))
    pose = np.concatenate([synth.quat_from_rotvec(0.5 * rng.standard_normal(3)), [0.04, -0.02, 0.8] + 0.05 * rng.standard_normal(3)])
    wpt = np.array([0.3 * rng.standard_normal(), 0.3 * rng.standard_normal(), 1.0])
    uv = np.array([600.0, 480.0]) + 100 * rng.standard_normal(2)
    return focal, pp, dist, pose, wpt, float(np.float32(uv[0])), float(np.float32(uv[1]))


@needs_ref
@pytest.mark.parametrize("model", synth.MODELS)
@pytest.mark.parametrize("mono", [True, False])
def test_restated_functors_equal_compiled_reference(model, mono):
    """14 functors: residual and Jacobian of oracle/lm_oracle.cpp == the reference's createCeresCostFunction(...)->Evaluate
    (+ AutoDiffLocalLeftSE3::PlusJacobian), compiled from /root/reference/include/dist_models/*.hpp."""
    rng = np.random.default_rng(11)
    mid = synth.MODEL_ID[model]
    assert R.num_dist(mid) == synth.NUM_DIST[mid]
    worst = 0.0
    for _ in range(200):
        focal, pp, dist, pose, wpt, u, v = _random_case(rng, model, mono)
        r0, J0 = O.residual_jacobian(mid, mono, focal, pp, dist, pose, wpt, u, v)
        r1, J1, _ = R.residual_jacobian(mid, mono, focal, pp, dist, pose, wpt, u, v)
        scale = np.maximum(1.0, np.abs(J1))
        worst = max(worst, float(np.max(np.abs(J0 - J1) / scale)), float(np.max(np.abs(r0 - r1) / np.maximum(1.0, np.abs(r1)))))
    assert worst <= 1e-14, worst


@needs_ref
def test_manifold_equals_compiled_reference():
    """AutoDiffLocalLeftSE3 (se3_param.hpp:12-41) compiled: Plus == the restatement, Minus inverts Plus, PlusJacobian at 0."""
    rng = np.random.default_rng(3)
    for k in range(50):
        T = np.concatenate([synth.quat_from_rotvec(rng.standard_normal(3)), rng.standard_normal(3)])
        d = np.concatenate([0.1 * rng.standard_normal(3), 0.2 * rng.standard_normal(3)]) * (1e-12 if k % 5 == 0 else 1.0)
        a, b = R.se3_plus(T, d), O.se3_plus(T, d)
        assert np.max(np.abs(a - b)) <= 1e-15, (a, b)
        assert np.allclose(R.se3_minus(a, T), d, atol=1e-12)
        h = 1e-6
        J = R.se3_plus_jacobian(T)
        for i in range(6):
            e = np.zeros(6); e[i] = h
            fd = (R.se3_plus(T, e) - R.se3_plus(T, -e)) / (2 * h)
            assert np.allclose(J[:, i], fd, atol=1e-8)


@needs_ref
@pytest.mark.parametrize("model", synth.MODELS)
def test_projection_equals_compiled_distort(model):
    """distortCamPoint(x, y) of the 7 DistParam classes (the per-frame RMSE path) == the restated projection."""
    rng = np.random.default_rng(8)
    mid = synth.MODEL_ID[model]
    dist = np.array(synth.GT_DIST[model])
    ident = np.array([0, 0, 0, 1.0, 0, 0, 0])
    for _ in range(50):
        x, y = 0.4 * rng.standard_normal(2)
        got = O.project(mid, False, np.array([1.0, 1.0]), np.zeros(2), dist, ident, np.array([x, y, 1.0]))
        assert np.max(np.abs(got - R.distort(mid, dist, x, y))) <= 1e-15
    # the std-dev quirk (dist_model_base.hpp:90-102): 2*nd values, first nd zero
    nd = synth.NUM_DIST[mid]
    cov = np.diag(np.arange(1.0, nd + 1.0) ** 2)
    s = R.dist_std_quirk(mid, cov)
    assert len(s) == 2 * nd and np.all(s[:nd] == 0) and np.allclose(s[nd:], np.arange(1.0, nd + 1.0))

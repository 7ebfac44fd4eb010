"""include/ezcalib_b200.hpp -- the C++ mirror of the reference's detector / calibrator interfaces above the C ABI:
it must compile against the header and link the shared library (CPU), and give exactly what the ctypes API gives (GPU)."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "host_mirror_check.cpp")
EXE = os.path.join(ROOT, "tests", "cpp", "host_mirror_check")


def _build():
    deps = [SRC, os.path.join(ROOT, "include", "ezcalib_b200.hpp"), os.path.join(ROOT, "include", "ezcalib_b200.h")]
    if not os.path.exists(EXE) or os.path.getmtime(EXE) < max(os.path.getmtime(d) for d in deps):
        subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-Wall", "-Werror", "-I" + os.path.join(ROOT, "include"), SRC,
                               "-L" + os.path.join(ROOT, "ezcalib_b200"), "-lezcalib_b200",
                               "-Wl,-rpath," + os.path.join(ROOT, "ezcalib_b200"), "-o", EXE])
    return EXE


def test_cpp_mirror_compiles_and_links():
    exe = _build()
    assert os.access(exe, os.X_OK)
    # no arguments: usage error code, and the process must have been able to load libezcalib_b200.so
    r = subprocess.run([exe], capture_output=True)
    assert r.returncode == 2, r.stderr


@pytest.mark.gpu
def test_cpp_mirror_matches_python_api(ctx, tmp_path):
    from ezcalib_b200 import detect, lm, synth
    exe = _build()
    G_cb = np.load(os.path.join(ROOT, "tests", "golden", "chessboard_golden.npz"))
    G_ag = np.load(os.path.join(ROOT, "tests", "golden", "aprilgrid_golden.npz"))
    cb = np.stack([G_cb["render0_img"], G_cb["warp1_img"], G_cb["blank_img"]])
    name = str(G_ag["names"][0])
    ag = np.stack([G_ag[name + "_img"], G_ag[name + "_img"][::-1].copy()])
    rows, cols = (int(v) for v in G_ag[name + "_grid"])
    assert (rows, cols) == (6, 6)
    model, mono = "radtan5", True
    tgt = synth.chessboard_target(7, 10, 0.05)
    P = synth.make_lm_problem(model, mono, 30, tgt, 1280, 1024, seed=5, drop_fraction=0.05)
    path = tmp_path / "mirror_input.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("3i", cb.shape[0], cb.shape[2], cb.shape[1])); f.write(cb.tobytes())
        f.write(struct.pack("3i", ag.shape[0], ag.shape[2], ag.shape[1])); f.write(ag.tobytes())
        f.write(struct.pack("4i", synth.MODEL_ID[model], int(mono), P.obs.shape[0], P.obs.shape[1]))
        f.write(struct.pack("2d", float(P.width), float(P.height)))
        f.write(np.resize(np.asarray(P.focal_init, np.float64), 2).tobytes())
        f.write(np.asarray(P.pp_init, np.float64).tobytes())
        d8 = np.zeros(8); d8[:len(P.dist_init)] = P.dist_init
        f.write(d8.tobytes())
        f.write(np.ascontiguousarray(P.target, np.float64).tobytes())
        f.write(np.ascontiguousarray(P.obs, np.float32).tobytes())
        f.write(np.ascontiguousarray(P.poses_init, np.float64).tobytes())
    r = subprocess.run([exe, str(path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    # ---- chessboard
    imgs = detect.Images(ctx, cb)
    found, corners = detect.detect_chessboard(ctx, imgs, 7, 10)
    imgs.close()
    t0 = [ln for ln in lines if ln.startswith("chessboard target")][0].split()
    assert int(t0[2]) == 54 and np.allclose([float(v) for v in t0[3:6]], tgt[0])
    got = [ln.split() for ln in lines if ln.startswith("chessboard ") and not ln.startswith("chessboard target")]
    assert len(got) == 3
    for k, g in enumerate(got):
        assert int(g[2]) == int(found[k])
        assert np.array_equal(np.array(g[3:], np.float32).reshape(-1, 2), corners[k])
    assert list(found) == [True, True, False]
    # ---- AprilGrid
    imgs = detect.Images(ctx, ag)
    succ, c_ag, ndet = detect.detect_aprilgrid(ctx, imgs, 6, 6)
    imgs.close()
    got = [ln.split() for ln in lines if ln.startswith("aprilgrid ")]
    assert len(got) == 2
    for k, g in enumerate(got):
        assert int(g[2]) == int(succ[k]) and int(g[3]) == int(ndet[k])
        assert np.array_equal(np.array(g[4:], np.float32).reshape(-1, 2), c_ag[k])
    # ---- calibration: same three-stage solve as lm.lm_mono
    f1, pp1, d1, poses1, rmse1, ss1 = lm.lm_mono(ctx, synth.MODEL_ID[model], mono, P.obs, P.target, P.width, P.height, P.focal_init,
                                                 P.pp_init, P.dist_init, P.poses_init)
    cal = [ln.split() for ln in lines if ln.startswith("calib focal")][0]
    i_pp, i_d = cal.index("pp"), cal.index("dist")
    assert np.array_equal(np.array(cal[2:i_pp], float), np.asarray(f1)[:1])
    assert np.array_equal(np.array(cal[i_pp + 1:i_d], float), pp1)
    assert np.array_equal(np.array(cal[i_d + 1:], float), d1)
    it = [ln.split() for ln in lines if ln.startswith("calib iterations")][0]
    assert [int(v) for v in it[2:5]] == [s["iterations"] for s in ss1]
    r0 = [ln.split() for ln in lines if ln.startswith("calib rmse0")][0]
    assert float(r0[2]) == rmse1[0]
    assert np.array_equal(np.array(r0[4:], float), poses1[0])
    std = [ln.split() for ln in lines if ln.startswith("calib std")][0]
    assert len(std) == 2 + 1 + 2 + 5 and all(float(v) > 0 for v in std[2:])
    assert "error-check ok" in lines


def test_cpp_param_mirrors_match_the_compiled_reference():
    """DistParam (7 models: distortCamPoint x3, undistortCamPoint, resetParameters / getDistParameters, the 2*nd std-dev
    quirk), makeDistParam's config strings, IntrinsicsParam and Camera::setupInitialCalibration of include/ezcalib_b200.hpp
    against the reference's own headers compiled into oracle/_ref/liblm_ref.so (CPU only, no CUDA context)."""
    from ezcalib_b200 import ezcalibrator as E, synth
    from oracle import lm_ref as R
    if not R.available():
        pytest.skip("oracle/_ref/liblm_ref.so not built and /root/reference absent")
    exe = _build()
    r = subprocess.run([exe, "--cpu-mirror"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    coefs = [-0.21, 0.06, -0.012, 0.004, 0.0011, -0.0007, 0.0009, -0.0005]
    dl = [ln.split() for ln in lines if ln.startswith("dist ")]
    assert [d[1] for d in dl] == list(synth.MODELS)
    for mid, d in enumerate(dl):
        nd = int(d[2])
        assert nd == R.num_dist(mid)
        v = [float(x) for x in d[3:9]]
        ref = R.distort(mid, coefs[:nd], 0.31, -0.17)
        assert np.max(np.abs(np.array(v[0:2]) - ref)) <= 1e-15 and np.max(np.abs(np.array(v[2:4]) - ref)) <= 1e-15
        und = R.undistort(mid, coefs[:nd], ref[0], ref[1])
        assert np.max(np.abs(np.array(v[4:6]) - und)) <= 1e-12 and np.max(np.abs(und - [0.31, -0.17])) < 2e-5
        nstd = int(d[9])
        sd = np.array(d[10:10 + nstd], float)
        assert np.array_equal(sd, R.dist_std_quirk(mid, np.diag(np.arange(1.0, nd + 1.0) ** 2)))
        assert np.array_equal(np.array(d[10 + nstd:], float), coefs[:nd])
    assert "unknown-model ok" in lines
    cam = [ln.split() for ln in lines if ln.startswith("camera ")][0]
    f0, pp0 = E.setup_initial_calibration(1280, 1024, 70.0, True)
    assert float(cam[1]) == f0[0] and [float(cam[2]), float(cam[3])] == list(pp0) and int(cam[4]) == 1
    assert abs(float(cam[5]) - (100.25 - 640) / f0[0]) < 1e-15 and abs(float(cam[6]) - (700.5 - 512) / f0[0]) < 1e-15
    assert abs(float(cam[7]) - (f0[0] * 0.1 + 640)) < 1e-12 and abs(float(cam[8]) - (f0[0] * -0.05 + 512)) < 1e-12
    assert [int(v) for v in cam[9:12]] == [1, 0, 0]
    se = np.array([ln.split() for ln in lines if ln.startswith("se3 ")][0][1:], float)
    T = R.se3_plus(np.array([0, 0, 0, 1.0, 0, 0, 0]), np.array([0.1, -0.2, 0.3, 0.2, -0.1, 0.4]))
    assert np.max(np.abs(se[:7] - T)) <= 1e-15
    assert np.max(np.abs(se[7:] - R.se3_minus(T, R.se3_plus(np.array([0, 0, 0, 1.0, 0, 0, 0]), np.array([0, 0, 0, 0, 0, 1e-12]))))) <= 1e-14


@pytest.mark.gpu
def test_cpp_rig_calibration_matches_python_api(ctx, tmp_path):
    """Two cameras through the C++ mirror classes (Camera, IntrinsicsParam, DistParam, computeP3PRansac, computeCalibration,
    runMultiCameraCalib) == the same flow through the Python host mirror."""
    from ezcalib_b200 import ezcalibrator as E, synth
    exe = _build()
    model, mono, W, H, n_frames, fov = "radtan5", True, 1280, 1024, 16, 70.0
    tgt = synth.chessboard_target(7, 10, 0.05)
    rng = np.random.Generator(np.random.Philox(key=91))
    d = synth.GT_DIST[model]
    poses0 = synth.sample_poses(rng, n_frames, tgt, model, 900.0, 645.0, 505.0, d, W, H, margin=140, tilt_sigma=0.25)
    T10 = np.concatenate([synth.quat_from_rotvec(np.array([0.01, -0.03, 0.005])), [-0.06, 0.002, 0.004]])
    corners = np.zeros((2, n_frames, len(tgt), 2), np.float32)
    for i, p in enumerate(poses0):
        corners[0, i] = synth.project(model, 900.0, 900.0, 645.0, 505.0, d, p, tgt) + rng.normal(0, 0.05, (len(tgt), 2))
        corners[1, i] = synth.project(model, 880.0, 880.0, 630.0, 515.0, d, synth.se3_mul(T10, p), tgt) + rng.normal(0, 0.05, (len(tgt), 2))
    corners[1, 3] = -1.0   # one view camera 1 does not see
    path = tmp_path / "rig.bin"
    with open(path, "wb") as f:
        f.write(struct.pack("6i", 2, n_frames, len(tgt), W, H, int(mono)))
        f.write(struct.pack("d", fov))
        f.write(model.encode().ljust(32, b"\0"))
        f.write(np.ascontiguousarray(tgt, np.float64).tobytes())
        f.write(np.ascontiguousarray(corners, np.float32).tobytes())
    r = subprocess.run([exe, "--rig", str(path)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    lines = r.stdout.strip().splitlines()
    cams = []
    for c in range(2):
        focal, pp = E.setup_initial_calibration(W, H, fov, mono)
        cam = E.Camera(model, mono, W, H, focal, pp, np.zeros(len(d)))
        ok, poses, _ = E.compute_p3p_ransac_batch(ctx, corners[c], tgt, focal, pp, W, H)
        for i in range(n_frames):
            if ok[i]:
                cam.frames.append(E.CalibFrame("%06d.png" % i, corners[c, i], poses[i]))
        E.compute_calibration(ctx, cam, tgt)
        cams.append(cam)
        ln = [x.split() for x in lines if x.startswith(f"rig cam {c} ")][0]
        assert int(ln[4]) == len(cam.frames)
        assert float(ln[6]) == cam.focal[0] and [float(ln[8]), float(ln[9])] == list(cam.pp)
        i_d = ln.index("dist")
        assert np.array_equal(np.array(ln[i_d + 1:i_d + 1 + len(d)], float), cam.dist)
        assert int(ln[ln.index("cov") + 1]) == 1 and int(ln[ln.index("nstd") + 1]) == 2 * len(d)
    out = E.run_multicam_calib(ctx, cams, tgt)
    ln = [x.split() for x in lines if x.startswith("rig pairs")][0]
    assert int(ln[2]) == out["n_good_pairs"] and int(ln[4]) == out["summary"]["iterations"]
    ext = np.array(ln[ln.index("ext") + 1:ln.index("ext") + 8], float)
    assert np.max(np.abs(ext - cams[1].T_cam0_2_cam)) <= 1e-9
    assert np.max(np.abs(ext[4:] - T10[4:])) < 2e-3   # and it is the rig's baseline

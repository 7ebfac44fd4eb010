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

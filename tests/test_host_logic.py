"""CPU: host-side logic that mirrors the reference's orchestration (no GPU needed)."""
import numpy as np
import pytest

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


def test_initial_calibration():
    """Camera::setupInitialCalibration (camera.cpp:102-120) as the host mirror restates it: prior focal from the field of
    view, principal point at the image centre.  (computeP3PRansac is a GPU entry point now: tests/test_p3p_cpu.py checks
    its mathematics against cv2 on the CPU, tests/test_p3p_gpu.py the kernel.)"""
    from ezcalib_b200 import ezcalibrator as E
    W, H = 1280, 1024
    f0, pp0 = E.setup_initial_calibration(W, H, 70.0, True)
    assert f0.shape == (1,) and abs(f0[0] - max(W / 2, H / 2) / np.tan(np.deg2rad(35.0))) < 1e-9
    assert np.array_equal(pp0, [W / 2, H / 2])
    f2, _ = E.setup_initial_calibration(W, H, 70.0, False)
    assert f2.shape == (2,) and f2[0] == f2[1] == f0[0]


def test_product_package_does_not_import_the_oracles():
    """The product path must not route through cv2 or oracle/ (VERDICT round 1, N1)."""
    import os
    import re
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "ezcalib_b200")
    for name in os.listdir(root):
        if name.endswith(".py"):
            text = open(os.path.join(root, name)).read()
            assert not re.search(r"^\s*(import|from)\s+(cv2|oracle)\b", text, re.M), name


def test_result_yaml_matches_the_reference_writer(tmp_path):
    """writeCamerasCalib (apps/ezcalib.cpp:69-187): the exact text for a two-camera result, incl. the [sic] key
    `intrisics_parameters`, the +/- comment lines, nine fixed decimals, the blank line after every matrix row."""
    cam0 = E.Camera("radtan5", True, 1280, 1024, np.array([900.123456789]), np.array([645.5, 505.25]),
                    np.array([-0.2, 0.05, 0.0, 0.001, -0.002]))
    cam1 = E.Camera("kannalabrandt", False, 1920, 1080, np.array([600.0, 601.5]), np.array([960.0, 540.0]), np.array([0.1, 0.01, 0.0, 0.0]),
                    T_cam0_2_cam=np.array([0.0, 0.0, 0.0, 1.0, -0.12, 0.001, 0.0005]))
    stds = [([0.5], [0.25, 0.125], [1e-3] * 5), ([0.5, 0.75], [0.25, 0.125], [1e-3] * 4)]
    txt = E.cameras_calib_yaml([cam0, cam1], ["/data/cam0/", "/data/cam1/"], stds)
    expect = (
        "%YAML:1.0\n---\n\n#Camera Calibration Parameters\n\n"
        "\n#Camera #0\ncam0:\n  input_images_folder: /data/cam0/\n  use_mono_focal: 1\n  dist_model: radtan5\n  image_size: [1280, 1024]\n"
        "  intrisics_parameters: [900.123456789, 900.123456789, 645.500000000, 505.250000000]\n"
        "                         # +/- [0.500000000, 0.500000000, 0.250000000, 0.125000000]\n"
        "  distortion_coefs: [-0.200000000, 0.050000000, 0.000000000, 0.001000000, -0.002000000]\n"
        "                         # +/- [0.001000000, 0.001000000, 0.001000000, 0.001000000, 0.001000000]\n"
        "\n#Camera #1\ncam1:\n  input_images_folder: /data/cam1/\n  use_mono_focal: 0\n  dist_model: kannalabrandt\n  image_size: [1920, 1080]\n"
        "  intrisics_parameters: [600.000000000, 601.500000000, 960.000000000, 540.000000000]\n"
        "                         # +/- [0.500000000, 0.750000000, 0.250000000, 0.125000000]\n"
        "  distortion_coefs: [0.100000000, 0.010000000, 0.000000000, 0.000000000]\n"
        "                         # +/- [0.001000000, 0.001000000, 0.001000000, 0.001000000]\n"
        "  #Transformation from cam0 frame to cam #1 frame.\n"
        "  T_cam0_2_cam: [1.000000000, 0.000000000, 0.000000000, -0.120000000, \n"
        "                0.000000000, 1.000000000, 0.000000000, 0.001000000, \n"
        "                0.000000000, 0.000000000, 1.000000000, 0.000500000]\n\n")
    assert txt == expect
    p = E.write_cameras_calib([cam0, cam1], str(tmp_path / "out.txt"), ["/data/cam0/", "/data/cam1/"], stds)
    assert p.endswith("out.yaml") and open(p).read() == expect


def test_config_reader_follows_the_reference_format(tmp_path):
    """setupCameras / setupCalibrationProblem: the reference's own file layout (its '%YAML 1.0' header included)."""
    cfg = tmp_path / "calib.yaml"
    cfg.write_text("%YAML 1.0\n---\n\ninput_images_folder:\n  - /data/left/\n  - /data/right/\n\nuse_mono_focal: # (if 0: fx != fy)\n  - 1\n  - 0\n\n"
                   "dist_model: # (rad1, ...)\n  - radtan5\n  - kannalabrandt\n\nprior_fov:\n  - 70.0\n  - 120.0\n\n\n"
                   "target_type: \"chessboard\" # (chessboard / aprilgrid)\n\ntarget_nb_rows: 7\ntarget_nb_cols: 12\n\nsquare_size: 0.05\n\ndebug: 0")
    c = E.read_calib_config(str(cfg))
    assert [k["dist_model"] for k in c["cameras"]] == ["radtan5", "kannalabrandt"]
    assert [k["use_mono_focal"] for k in c["cameras"]] == [True, False]
    assert c["cameras"][1]["prior_fov"] == 120.0 and c["target_type"] == "chessboard"
    assert (c["target_nb_rows"], c["target_nb_cols"], c["square_size"], c["debug"]) == (7, 12, 0.05, 0)
    bad = tmp_path / "bad.yaml"
    bad.write_text(cfg.read_text().replace("  - 70.0\n  - 120.0\n", "  - 70.0\n"))
    with pytest.raises(ValueError, match="prior_fov"):
        E.read_calib_config(str(bad))


def test_rig_point_transform_is_bit_identical_to_synth_transform():
    """run_rig_extrinsics / build_multicam_problem transform the target into cam0's frames with the cross products written
    out (one batched call for 2000 frames); the result must equal synth.transform bit for bit -- the C++ mirror and the
    oracle consume the same points."""
    from ezcalib_b200 import ezcalibrator as E, synth
    rng = np.random.default_rng(3)
    poses = rng.normal(size=(257, 7))
    poses[:, :4] /= np.linalg.norm(poses[:, :4], axis=1, keepdims=True)
    tgt = synth.aprilgrid_target(6, 6, 0.06, 0.3)
    assert np.array_equal(E._transform_points(poses, tgt), synth.transform(poses, tgt))
    one = np.concatenate([synth.transform(p, tgt) for p in poses[:5]], axis=0)
    assert np.array_equal(E._transform_points(poses[:5], tgt).reshape(-1, 3), one)

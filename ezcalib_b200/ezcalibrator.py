"""Host-side orchestration mirroring the reference's EZCalibrator / Camera / CalibFrame
(/root/reference/include/ezcalibrator.hpp, camera.hpp): frame selection (skip rule), the mono 3-stage calibration and
the multi-camera extrinsic calibration, expressed over the GPU entry points.  Only bookkeeping happens here."""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from . import lm, synth


@dataclass
class CalibFrame:
    """camera.hpp:14-33"""
    img_name: str
    corners_px: np.ndarray          # [n_pts, 2] float32 (cv::Point2f), (-1,-1) = missing
    T_world_2_cam: np.ndarray       # [7] qx qy qz qw tx ty tz
    rmse_err: float = -1.0


@dataclass
class Camera:
    """camera.hpp:35-69 (state only)."""
    model: str
    use_mono_focal: bool
    width: int
    height: int
    focal: np.ndarray
    pp: np.ndarray
    dist: np.ndarray
    frames: list = field(default_factory=list)
    T_cam0_2_cam: np.ndarray | None = None
    stds: tuple | None = None       # (focal_std, pp_std, dist_std) after compute_calibration


def replay_skip_rule(found_flags) -> list[int]:
    """EZCalibrator::runCalibration's image loop (/root/reference/src/ezcalibrator.cpp:39-59): returns the indices of the
    images the reference would have processed, given the per-image detection result of ALL images (detectTarget is a
    pure function of the image, so detecting everything in one batch and replaying the rule is equivalent).
    After a failure the iterator skips nb_consec_bad/2 further images (the reference can step past end(); we stop)."""
    out = []
    n = len(found_flags)
    i = 0
    bad = 0
    while i < n:
        out.append(i)
        if found_flags[i]:
            bad = 0
        else:
            bad += 1
            i += bad // 2
        i += 1
    return out


def _stds_from_covariance(solver: lm.LmSolver, cam: Camera):
    """ceres::Covariance over (focal, pp, dist) -> setFocalStd / setPPStd / setDistParamsStd (ezcalibrator.cpp:293-331).
    A rank-deficient system is not fatal: the reference prints 'Covariance Estimation Failed' and keeps the calibration."""
    from ._lib import EzcError
    try:
        cov = solver.covariance()
    except EzcError:
        cam.stds = None
        return
    sd = np.sqrt(np.diag(cov))
    nf = 1 if cam.use_mono_focal else 2
    cam.stds = (sd[:nf].copy(), sd[nf:nf + 2].copy(), sd[nf + 2:].copy())


def _solve_frames(ctx: lm.Context, cam: Camera, target: np.ndarray, frames: list, stages):
    obs = np.stack([f.corners_px for f in frames]).astype(np.float32)
    poses = np.stack([f.T_world_2_cam for f in frames])
    solver = lm.LmSolver(ctx, synth.MODEL_ID[cam.model], cam.use_mono_focal, obs, target, cam.width, cam.height)
    try:
        solver.set_params(cam.focal, cam.pp, cam.dist, poses)
        summ = [solver.solve(*mask)[0] for mask in stages]
        rmse = solver.frame_rmse()
        f, pp, d, poses = solver.get_params()
        _stds_from_covariance(solver, cam)
    finally:
        solver.close()
    cam.focal, cam.pp, cam.dist = f, pp, d
    for fr, p, r in zip(frames, poses, rmse):
        fr.T_world_2_cam = p
        fr.rmse_err = float(r)
    return summ


def compute_calibration(ctx: lm.Context, cam: Camera, target: np.ndarray):
    """EZCalibrator::computeCalibration (ezcalibrator.cpp:105-339): 3 staged solves (focal + poses; + distortion; + principal
    point; opt_focal, opt_pp, opt_dist masks), parameters written back into the camera / frames, rmse_err filled, standard
    deviations from the covariance in cam.stds."""
    return _solve_frames(ctx, cam, target, cam.frames, [(True, False, False), (True, False, True), (True, True, True)])


def refine_calibration(ctx: lm.Context, cam: Camera, target: np.ndarray):
    """EZCalibrator::refineCalibration (ezcalibrator.cpp:342-557; dormant in the reference, its call at :63 is commented
    out): ONE solve with every block variable over the frames whose rmse_err <= 1; their rmse_err is recomputed, the
    others keep theirs.  It drops the views stuck in the mirror pose of the planar target."""
    good = [f for f in cam.frames if not f.rmse_err > 1.0]
    if not good:
        raise RuntimeError("refine_calibration: no frame with rmse_err <= 1")
    return _solve_frames(ctx, cam, target, good, [(True, True, True)])[0]


def multicam_block(cam0_frames: list, camj: Camera, cam_index: int, target: np.ndarray):
    """One camera's share of EZCalibrator::runMultiCameraCalib's problem (ezcalibrator.cpp:654-785): the initial T_cj_c0 and
    the observation rows against cam0's frames.  cam0_frames: objects with img_name / T_world_2_cam / rmse_err (cam0's
    CalibFrames, or what the rank that owns cam0 broadcast).  Returns (obs [n_frames0 * n_pts, 2], init [7], n_good_pairs).
    Reproduces the reference's initialisation quirk Q3: the 'median' for camera i is taken over copies of the relative pose
    of cam0's frame #i (indexed with the CAMERA index, :680).  Frame lists are sorted by image name (the reference's linear
    scans with their early `break` rely on it, :688-704), so a name lookup finds the same partner."""
    n_pts = target.shape[0]
    n0 = len(cam0_frames)
    by_name = {f.img_name: f for f in camj.frames}
    if cam_index >= n0:
        raise IndexError("reference indexes cam0.m_v_calib_frames[camera index] out of range (quirk Q3)")
    f0 = cam0_frames[cam_index]
    logs = []
    n_good = 0
    if not f0.rmse_err > 1.0:
        fj = by_name.get(f0.img_name)
        if fj is not None:
            lg = synth.se3_log(synth.se3_mul(fj.T_world_2_cam, synth.se3_inverse(f0.T_world_2_cam)))
            logs = [lg] * n0          # the reference pushes the same value once per cam0 frame (loop variable j unused, :678-680)
            n_good = n0
    if not logs:
        raise IndexError(f"no usable frame pair for the extrinsic initialisation of camera {cam_index}: cam0 frame #{cam_index} "
                         f"({f0.img_name!r}, rmse_err {f0.rmse_err:.3f}) {'has rmse_err > 1' if f0.rmse_err > 1.0 else 'has no partner frame in this camera'}"
                         f" ({len(camj.frames)} frames) -- the reference reads out of bounds here (ezcalibrator.cpp:707-713)")
    logs = np.sort(np.asarray(logs), axis=0)
    med = logs[len(logs) // 2]
    init = synth.se3_exp_left(med, np.array([0, 0, 0, 1.0, 0, 0, 0]))
    obs = np.full((n0 * n_pts, 2), -1.0, np.float32)
    for a_, f0 in enumerate(cam0_frames):
        if f0.rmse_err > 1.0:
            continue
        fj = by_name.get(f0.img_name)
        if fj is not None and not fj.rmse_err > 1.0:
            obs[a_ * n_pts:(a_ + 1) * n_pts] = fj.corners_px
    return obs, init, n_good


def build_multicam_problem(cameras: list[Camera], target: np.ndarray):
    """Problem set-up of EZCalibrator::runMultiCameraCalib (ezcalibrator.cpp:631-785): returns
    (obs [n_cams-1, n_frames0*n_pts, 2], pts [n_frames0*n_pts, 3], init extrinsics [n_cams-1, 7], n_good_pairs)."""
    cam0 = cameras[0]
    pts = _transform_points(np.stack([f.T_world_2_cam for f in cam0.frames]), target).reshape(-1, 3)
    blocks = [multicam_block(cam0.frames, cameras[i], i, target) for i in range(1, len(cameras))]
    obs = np.stack([b[0] for b in blocks])
    init = np.stack([b[1] for b in blocks])
    return obs, pts, init, int(sum(b[2] for b in blocks))


def run_multicam_calib(ctx: lm.Context, cameras: list[Camera], target: np.ndarray):
    """EZCalibrator::runMultiCameraCalib: intrinsics of cameras 1..N-1 frozen, one SE3 block T_ci_c0 per camera, one
    joint solve (one trust region for all blocks, like the single ceres::Problem of the reference).
    Limitation of this round: all cameras must share dist_model / use_mono_focal (true for every BASELINE config)."""
    if len(cameras) < 2:
        return None
    m0, mono0 = cameras[1].model, cameras[1].use_mono_focal
    if any(c.model != m0 or c.use_mono_focal != mono0 for c in cameras[1:]):
        raise NotImplementedError("mixed distortion models across cameras are not supported yet")
    obs, pts, init, n_good = build_multicam_problem(cameras, target)
    cam1 = cameras[1]
    focal = np.concatenate([c.focal for c in cameras[1:]])
    pp = np.concatenate([c.pp for c in cameras[1:]])
    dist = np.concatenate([c.dist for c in cameras[1:]])
    ext, summ = lm.lm_multicam(ctx, synth.MODEL_ID[m0], mono0, obs, pts, cam1.width, cam1.height, focal, pp, dist, init)
    for c, e in zip(cameras[1:], ext):
        c.T_cam0_2_cam = e
    return dict(summary=summ, n_good_pairs=n_good, init=init, extrinsics=ext, obs=obs, pts=pts)


# ------------------------------------------------------------------------------------------------------------------
# The step BETWEEN the two accelerated halves: initial intrinsics and P3P-RANSAC poses.
def setup_initial_calibration(width: int, height: int, prior_fov_deg: float, use_mono_focal: bool):
    """Camera::setupInitialCalibration (/root/reference/src/camera.cpp:102-120): focal from the prior field of view,
    principal point at the image centre.  Returns (focal[1|2], pp[2])."""
    cx, cy = width / 2.0, height / 2.0
    pi = 2.0 * np.arcsin(1.0)
    inv_tan = 1.0 / np.tan(prior_fov_deg / 2.0 * (pi / 180.0))
    f = max(cx * inv_tan, cy * inv_tan)
    return np.array([f] if use_mono_focal else [f, f], np.float64), np.array([cx, cy], np.float64)


def compute_p3p_ransac_batch(ctx: lm.Context, corners_px: np.ndarray, target: np.ndarray, focal: np.ndarray, pp: np.ndarray, width: int,
                             height: int, n_target_pts: int | None = None, want_debug: bool = False):
    """EZCalibrator::computeP3PRansac (/root/reference/src/ezcalibrator.cpp:560-628) for a batch of views on the GPU
    (ezc_p3p_ransac_batch, one warp per view; csrc/p3p.cu).  corners_px [n_views, n_pts, 2] float32 pixels, (-1,-1) = missing.
    Pixels are normalised with the PRIOR intrinsics (no distortion), points outside the image dropped, >= 32 points and
    >= 32 inliers required when the target has >= 32 points; cv::solvePnPRansac(identity K, 1000 iterations, threshold
    4 / mean focal, confidence 0.999, AP3P) restated (RANSAC with cv::RNG's sequence + EPnP on the inliers),
    Rodrigues -> float32 -> fitToSO3.  Returns (ok[n_views] bool, poses7[n_views, 7], n_inliers[n_views]) and, with
    want_debug, also (inlier mask [n_views, n_pts], rvec_tvec [n_views, 6], RANSAC iterations [n_views])."""
    import ctypes as C

    from . import _lib
    corners = np.ascontiguousarray(corners_px, np.float32)
    assert corners.ndim == 3 and corners.shape[2] == 2
    n_views, n_pts = corners.shape[:2]
    tg = np.ascontiguousarray(target, np.float64)
    assert tg.shape == (n_pts, 3)
    n_target_pts = n_pts if n_target_pts is None else int(n_target_pts)
    focal = np.ascontiguousarray(focal, np.float64)
    ppa = np.ascontiguousarray(pp, np.float64)
    fx, fy = float(focal[0]), float(focal[-1])
    mean_focal = fx if len(focal) == 1 else 0.5 * (fx + fy)
    ok = np.zeros(n_views, np.uint8)
    poses = np.zeros((n_views, 7))
    ninl = np.zeros(n_views, np.int32)
    mask = np.zeros((n_views, n_pts), np.uint8) if want_debug else None
    rt = np.zeros((n_views, 6)) if want_debug else None
    nit = np.zeros(n_views, np.int32) if want_debug else None

    def p(a):
        return None if a is None else a.ctypes.data_as(C.c_void_p)
    _lib.check(_lib.lib().ezc_p3p_ransac_batch(ctx.handle, p(corners), p(tg), n_views, n_pts, n_target_pts, p(focal), len(focal), p(ppa),
                                               int(width), int(height), 1000, 4.0 / mean_focal, 0.999, 32, p(ok), p(poses), p(ninl), p(mask),
                                               p(rt), p(nit)), "ezc_p3p_ransac_batch")
    if want_debug:
        return ok.astype(bool), poses, ninl, mask.astype(bool), rt, nit
    return ok.astype(bool), poses, ninl


def run_calibration(ctx: lm.Context, images: np.ndarray, target_kind: str, nb_rows: int, nb_cols: int, size: float, space: float,
                    model: str, use_mono_focal: bool, prior_fov_deg: float, names=None):
    """EZCalibrator::runCalibration for one camera (/root/reference/src/ezcalibrator.cpp:14-103) over a set of gray images
    [n, H, W]: batched GPU detection, the reference's image loop replayed on its results, P3P-RANSAC initial poses, the
    three-stage GPU solve.  Returns (camera, target, processed image indices)."""
    from . import detect
    n, H, W = images.shape
    imgs = detect.Images(ctx, images, async_upload=True)
    if target_kind == "chessboard":
        target = synth.chessboard_target(nb_rows, nb_cols, size)
        found, corners = detect.detect_chessboard(ctx, imgs, nb_rows, nb_cols)
    elif target_kind == "aprilgrid":
        target = synth.aprilgrid_target(nb_rows, nb_cols, size, space)
        found, corners, _ = detect.detect_aprilgrid(ctx, imgs, nb_rows, nb_cols)
    else:
        raise ValueError(f"unknown target {target_kind!r}")
    imgs.close()
    focal, pp = setup_initial_calibration(W, H, prior_fov_deg, use_mono_focal)
    dist = np.zeros(len(synth.GT_DIST[model]))   # all coefficients start at 0 (member defaults of the dist_model classes)
    cam = Camera(model, use_mono_focal, W, H, focal, pp, dist)
    processed = replay_skip_rule([bool(f) for f in found])
    cand = [i for i in processed if found[i]]
    if cand:
        ok, poses, _ = compute_p3p_ransac_batch(ctx, np.asarray(corners)[cand], target, focal, pp, W, H)
        for k, i in enumerate(cand):
            if ok[k]:   # quirk Q4: a frame is stored only if P3P succeeds
                cam.frames.append(CalibFrame(names[i] if names is not None else str(i), corners[i].astype(np.float32), poses[k]))
    if not cam.frames:
        raise RuntimeError("no frame with a detected target and a P3P pose")
    compute_calibration(ctx, cam, target)
    return cam, target, processed


# ------------------------------------------------------------------------------------------------------------------
# The whole application flow for a camera rig (apps/ezcalib.cpp:205-218: runCalibration per camera, then
# runMultiCameraCalib), camera-per-GPU: rank r owns the cameras c with c % world == r.
class SingleProcessComm:
    """The communicator interface run_rig_calibration needs; one process, no exchange."""
    rank, world = 0, 1

    def bcast(self, obj, src: int):
        return obj

    def allgather(self, obj):
        return [obj]


class _FrameInfo:
    __slots__ = ("img_name", "T_world_2_cam", "rmse_err")

    def __init__(self, img_name, T, rmse):
        self.img_name, self.T_world_2_cam, self.rmse_err = img_name, T, rmse


def calibrate_camera_from_corners(ctx: lm.Context, found, corners, names, target, model, use_mono_focal, prior_fov_deg, W, H, timings=None):
    """EZCalibrator::runCalibration after detection (ezcalibrator.cpp:39-67): skip-rule replay, P3P-RANSAC, 3-stage solve."""
    import time
    focal, pp = setup_initial_calibration(W, H, prior_fov_deg, use_mono_focal)
    cam = Camera(model, use_mono_focal, W, H, focal, pp, np.zeros(synth.NUM_DIST[synth.MODEL_ID[model]]))
    processed = replay_skip_rule([bool(f) for f in found])
    cand = [i for i in processed if found[i]]
    t0 = time.perf_counter()
    if cand:
        ok, poses, _ = compute_p3p_ransac_batch(ctx, np.asarray(corners)[cand], target, focal, pp, W, H)
        for k, i in enumerate(cand):
            if ok[k]:
                cam.frames.append(CalibFrame(names[i], np.asarray(corners[i], np.float32), poses[k]))
    t1 = time.perf_counter()
    if not cam.frames:
        raise RuntimeError("no frame with a detected target and a P3P pose")
    summ = compute_calibration(ctx, cam, target)
    t2 = time.perf_counter()
    if timings is not None:
        timings["p3p_s"] = timings.get("p3p_s", 0.0) + (t1 - t0)
        timings["solve_s"] = timings.get("solve_s", 0.0) + (t2 - t1)
    return cam, processed, summ


def _transform_points(poses: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """synth.transform(poses [n,7], pts [m,3]) -> [n, m, 3] with the cross products written out (np.cross spends most of its
    time in axis bookkeeping): the same multiplications and subtractions in the same order, so the result is bit-identical."""
    qx, qy, qz, qw = (poses[:, None, i] for i in range(4))
    px, py, pz = pts[None, :, 0], pts[None, :, 1], pts[None, :, 2]
    ux, uy, uz = 2.0 * (qy * pz - qz * py), 2.0 * (qz * px - qx * pz), 2.0 * (qx * py - qy * px)
    out = np.empty((poses.shape[0], pts.shape[0], 3))
    out[..., 0] = px + qw * ux + (qy * uz - qz * uy) + poses[:, None, 4]
    out[..., 1] = py + qw * uy + (qz * ux - qx * uz) + poses[:, None, 5]
    out[..., 2] = pz + qw * uz + (qx * uy - qy * ux) + poses[:, None, 6]
    return out


def run_rig_extrinsics(ctx: lm.Context, local_cameras: dict, n_cams: int, target: np.ndarray, comm=None, timings=None):
    """EZCalibrator::runMultiCameraCalib over ranks: local_cameras {camera index: calibrated Camera} holds the cameras this
    rank owns.  The owner of cam0 shares cam0's frame names, poses and rmse_err (the only data the other cameras' residual
    blocks need, ezcalibrator.cpp:727-767) in one object all-gather; every rank builds the blocks of its own cameras; ONE joint solve -- the J^T J
    is block diagonal per camera, Ceres still uses one trust region and one termination test for all of them, which is what the
    per-iteration scalar all-reduce of the view-sharded solver provides (install it with ctx.set_allreduce before calling).
    Returns {camera index: T_cam0_2_cam} for ALL cameras 1..n_cams-1 on every rank, and the summary."""
    import time
    comm = comm or SingleProcessComm()
    t0 = time.perf_counter()
    # ONE object exchange up front (an object collective costs ~10 ms over NCCL, the joint solve itself ~2 ms): the owner of
    # cam0 contributes cam0's frames as three flat objects (names, poses, rmse), the owner of cam1 the shared camera model
    mine = sorted(c for c in local_cameras if c >= 1)
    contrib = {}
    if 0 in local_cameras:
        fr0 = local_cameras[0].frames
        contrib["cam0"] = ([f.img_name for f in fr0], np.stack([f.T_world_2_cam for f in fr0]), np.array([f.rmse_err for f in fr0]))
    if 1 in local_cameras:
        c1 = local_cameras[1]
        contrib["ref"] = (c1.model, c1.use_mono_focal, c1.width, c1.height)
    merged = {}
    for part in comm.allgather(contrib):
        merged.update(part)
    names0, poses0, rmse0 = merged["cam0"]
    m0, mono0, W, H = merged["ref"]
    cam0_frames = [_FrameInfo(nm, poses0[i], float(rmse0[i])) for i, nm in enumerate(names0)]
    n_pts = target.shape[0]
    pts = _transform_points(poses0, target).reshape(-1, 3)
    blocks = [multicam_block(cam0_frames, local_cameras[c], c, target) for c in mine]
    if any(local_cameras[c].model != m0 or local_cameras[c].use_mono_focal != mono0 for c in mine):
        raise NotImplementedError("mixed distortion models across cameras are not supported yet")
    nd = synth.NUM_DIST[synth.MODEL_ID[m0]]
    if mine:
        obs = np.stack([b[0] for b in blocks])
        init = np.stack([b[1] for b in blocks])
        focal = np.concatenate([local_cameras[c].focal for c in mine])
        pp = np.concatenate([local_cameras[c].pp for c in mine])
        dist = np.concatenate([local_cameras[c].dist for c in mine])
    else:
        obs = np.zeros((0, len(cam0_frames) * n_pts, 2), np.float32)
        init = np.zeros((0, 7))
        focal, pp, dist = np.zeros(1 if mono0 else 2), np.zeros(2), np.zeros(nd)
    t1 = time.perf_counter()
    ext, summ = lm.lm_multicam(ctx, synth.MODEL_ID[m0], mono0, obs, pts, W, H, focal, pp, dist, init, n_blocks_global=n_cams - 1,
                               collective=1 if comm.world > 1 else -1)
    t2 = time.perf_counter()
    parts = comm.allgather(({c: ext[k] for k, c in enumerate(mine)}, int(sum(b[2] for b in blocks))))
    out = {}
    for p_, _ in parts:
        out.update(p_)
    for c in mine:
        local_cameras[c].T_cam0_2_cam = out[c]
    summ["n_good_pairs"] = int(sum(g for _, g in parts))
    if timings is not None:
        timings["multicam_setup_s"] = t1 - t0
        timings["multicam_solve_s"] = t2 - t1
    return out, summ


# ------------------------------------------------------------------------------------------------------------------
# Result file: byte-compatible with writeCamerasCalib of the reference.
def _fixed9(v: float) -> str:
    """std::fixed << std::setprecision(9)"""
    return f"{float(v):.9f}"


def cameras_calib_yaml(cameras: list, input_paths=None, stds=None) -> str:
    """The text writeCamerasCalib (/root/reference/apps/ezcalib.cpp:69-187) writes for these cameras: `%YAML:1.0`, one
    `camN:` block per camera with image_size, `intrisics_parameters` [sic] fx fy cx cy and the +/- comment line,
    distortion_coefs with its +/- line, and for N > 0 the 3x4 row-major `T_cam0_2_cam`; nine fixed decimals.
    stds: per camera (focal_std[1|2], pp_std[2], dist_std[nd]) from the covariance (ezc_lm_covariance); zeros if None."""
    out = ["%YAML:1.0\n", "---\n\n", "#Camera Calibration Parameters\n\n"]
    for i, cam in enumerate(cameras):
        fx = float(cam.focal[0])
        fy = float(cam.focal[0] if cam.use_mono_focal else cam.focal[1])
        nd = len(cam.dist)
        if stds is not None and stds[i] is not None:
            fstd, pstd, dstd = stds[i]
            fstd = [float(fstd[0]), float(fstd[0] if cam.use_mono_focal or len(fstd) < 2 else fstd[1])]
        else:
            fstd, pstd, dstd = [0.0, 0.0], [0.0, 0.0], [0.0] * nd
        out.append(f"\n#Camera #{i}\n")
        out.append(f"cam{i}:\n")
        out.append(f"  input_images_folder: {input_paths[i] if input_paths else ''}\n")
        out.append(f"  use_mono_focal: {int(bool(cam.use_mono_focal))}\n")
        out.append(f"  dist_model: {cam.model}\n")
        out.append(f"  image_size: [{int(cam.width)}, {int(cam.height)}]\n")
        out.append("  intrisics_parameters: [" + ", ".join(_fixed9(v) for v in (fx, fy, cam.pp[0], cam.pp[1])) + "]\n")
        out.append("                         # +/- [" + ", ".join(_fixed9(v) for v in (fstd[0], fstd[1], pstd[0], pstd[1])) + "]\n")
        out.append("  distortion_coefs: [" + ", ".join(_fixed9(v) for v in cam.dist) + "]\n")
        out.append("                         # +/- [" + ", ".join(_fixed9(v) for v in dstd) + "]\n")
        if i > 0:
            T = synth.pose_to_matrix34(cam.T_cam0_2_cam)
            out.append(f"  #Transformation from cam0 frame to cam #{i} frame.\n")
            out.append("  T_cam0_2_cam: [")
            for r in range(3):
                if r > 0:
                    out.append("                ")
                for c in range(4):
                    out.append(_fixed9(T[r, c]) + ("]\n" if (r == 2 and c == 3) else ", "))
                out.append("\n")
    return "".join(out)


def write_cameras_calib(cameras: list, path: str = "cam_calib.yaml", input_paths=None, stds=None) -> str:
    """writeCamerasCalib: same path rules (a missing or foreign extension becomes .yaml, an existing file is replaced)."""
    import os
    root, ext = os.path.splitext(path)
    if ext != ".yaml":
        path = root + ".yaml"
    if os.path.exists(path):
        os.remove(path)
    with open(path, "w") as f:
        f.write(cameras_calib_yaml(cameras, input_paths, stds))
    return path


def read_calib_config(path: str) -> dict:
    """The configuration file of the reference (calib_*_config.yaml; apps/ezcalib.cpp:10-68 setupCameras and
    src/ezcalibrator.cpp:846-898 setupCalibrationProblem): one entry per camera for input_images_folder / use_mono_focal /
    dist_model / prior_fov (all four lists must have the same length), the target description, the debug flag.
    Errors raise ValueError where the reference prints and calls exit(-1)."""
    import yaml
    with open(path) as f:
        text = f.read()
    lines = text.splitlines()
    if lines and lines[0].startswith("%YAML"):
        lines = lines[1:]          # OpenCV's "%YAML 1.0" / "%YAML:1.0" directive is not something PyYAML accepts
    cfg = yaml.safe_load("\n".join(lines)) or {}
    folders = cfg.get("input_images_folder")
    if not isinstance(folders, list):
        raise ValueError("input_images_folder must be a sequence (one folder per camera)")
    n = len(folders)
    for key in ("use_mono_focal", "dist_model", "prior_fov"):
        v = cfg.get(key)
        if not isinstance(v, list) or len(v) != n:
            raise ValueError(f"Config file must have as many {key} entries as provided input image folders")
    for m in cfg["dist_model"]:
        if m not in synth.MODEL_ID:
            raise ValueError(f"The provided distortion model: {m} is not implemented")
    target = cfg.get("target_type")
    out = {"cameras": [dict(input_images_folder=str(folders[i]), use_mono_focal=bool(int(cfg["use_mono_focal"][i])),
                            dist_model=str(cfg["dist_model"][i]), prior_fov=float(cfg["prior_fov"][i])) for i in range(n)],
           "target_type": target, "target_nb_rows": int(cfg["target_nb_rows"]), "target_nb_cols": int(cfg["target_nb_cols"]),
           "debug": int(cfg.get("debug", 0))}
    if target == "aprilgrid":
        out["tag_size"], out["tag_space"] = float(cfg["tag_size"]), float(cfg["tag_space"])
    elif target == "chessboard":
        out["square_size"] = float(cfg["square_size"])
    else:
        raise ValueError(f"Not implemented target type: {target} (implemented: chessboard / aprilgrid)")
    return out

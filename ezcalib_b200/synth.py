"""Deterministic synthetic calibration data.

The reference ships no sample data (its YAML configs point at the author's /data/... folders),
so every workload in BASELINE.json is synthetic: target geometry follows the reference's
detectors, poses / noise follow SURVEY.md 8d (master seed 20261019).

Target coordinates mirror
  * ChessboardCalibDetector::setupCalibTarget  (/root/reference/include/target_detectors/chessboard_detector.hpp:71-84)
  * AprilTagCalibDetector::setupCalibTarget    (/root/reference/include/target_detectors/aprilgrid_detector.hpp:93-119)
Initial intrinsics mirror Camera::setupInitialCalibration (/root/reference/src/camera.cpp:102-120).
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import numpy as np

MASTER_SEED = 20261019

MODELS = ("rad1", "rad2", "rad3", "radtan4", "radtan5", "radtan8", "kannalabrandt")
MODEL_ID = {m: i for i, m in enumerate(MODELS)}
NUM_DIST = (1, 2, 3, 4, 5, 8, 4)

# Ground-truth distortion used by the synthetic workloads (reference coefficient order).
GT_DIST = {
    "rad1": (-0.20,),
    "rad2": (-0.25, 0.07),
    "rad3": (-0.28, 0.09, -0.01),
    "radtan4": (-0.25, 0.07, 7e-4, -4e-4),
    "radtan5": (-0.28, 0.09, -0.01, 7e-4, -4e-4),          # [k1,k2,k3,p1,p2]
    "radtan8": (0.12, -0.06, 0.01, 0.35, -0.03, 0.004, 7e-4, -4e-4),  # [k1..k6,p1,p2]
    "kannalabrandt": (0.02, -0.005, 0.001, -2e-4),
}


# ----------------------------------------------------------------------------- targets
def chessboard_target(nb_rows: int, nb_cols: int, square_size: float) -> np.ndarray:
    """(nb_rows-1)*(nb_cols-1) x 3; rows/cols count SQUARES; z = 1 (quirk Q6)."""
    pts = []
    for r in range(1, nb_rows):
        for c in range(nb_cols - 1, 0, -1):
            pts.append((-c * square_size, r * square_size, 1.0))
    return np.asarray(pts, dtype=np.float64)


def aprilgrid_target(nb_rows: int, nb_cols: int, tag_size: float, tag_space: float) -> np.ndarray:
    """4*nb_tags x 3, tag id = r*nb_cols + c, corner order BL, BR, TR, TL; z = 1."""
    spacing = tag_size * tag_space
    off = tag_size + spacing
    corners = np.array([[0.0, 0.0, 1.0], [tag_size, 0.0, 1.0], [tag_size, -tag_size, 1.0], [0.0, -tag_size, 1.0]])
    pts = []
    for r in range(nb_rows):
        for c in range(nb_cols):
            o = np.array([c * off, -r * off, 0.0])
            for k in range(4):
                pts.append(corners[k] + o)
    return np.asarray(pts, dtype=np.float64)


# ----------------------------------------------------------------------------- SE3 helpers
def quat_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """Hamilton product, storage [x,y,z,w], broadcasting over leading dims."""
    ax, ay, az, aw = np.moveaxis(a, -1, 0)
    bx, by, bz, bw = np.moveaxis(b, -1, 0)
    return np.stack([
        aw * bx + ax * bw + ay * bz - az * by,
        aw * by + ay * bw + az * bx - ax * bz,
        aw * bz + az * bw + ax * by - ay * bx,
        aw * bw - ax * bx - ay * by - az * bz,
    ], axis=-1)


def quat_rotate(q: np.ndarray, p: np.ndarray) -> np.ndarray:
    qv = q[..., :3]
    uv = 2.0 * np.cross(qv, p)
    return p + q[..., 3:4] * uv + np.cross(qv, uv)


def quat_from_matrix(R: np.ndarray) -> np.ndarray:
    """Batched rotation matrix -> unit quaternion [x,y,z,w] (w >= 0)."""
    R = np.asarray(R, dtype=np.float64)
    shape = R.shape[:-2]
    R = R.reshape(-1, 3, 3)
    q = np.empty((R.shape[0], 4))
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    for i in range(R.shape[0]):
        m = R[i]
        if tr[i] > 0:
            s = math.sqrt(tr[i] + 1.0) * 2
            q[i] = ((m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s, 0.25 * s)
        elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
            s = math.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
            q[i] = (0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s, (m[2, 1] - m[1, 2]) / s)
        elif m[1, 1] > m[2, 2]:
            s = math.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
            q[i] = ((m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s, (m[0, 2] - m[2, 0]) / s)
        else:
            s = math.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
            q[i] = ((m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s, (m[1, 0] - m[0, 1]) / s)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    q[q[:, 3] < 0] *= -1.0
    return q.reshape(shape + (4,))


def quat_from_rotvec(w: np.ndarray) -> np.ndarray:
    w = np.asarray(w, dtype=np.float64)
    th = np.linalg.norm(w, axis=-1, keepdims=True)
    half = 0.5 * th
    k = np.where(th > 1e-12, np.sin(half) / np.where(th > 1e-12, th, 1.0), 0.5 - th * th / 48.0)
    return np.concatenate([k * w, np.cos(half)], axis=-1)


def se3_exp_left(delta: np.ndarray, pose: np.ndarray) -> np.ndarray:
    """exp(delta) * T, delta = (upsilon, omega), pose = [qx qy qz qw tx ty tz]; batched."""
    delta = np.asarray(delta, dtype=np.float64)
    pose = np.asarray(pose, dtype=np.float64)
    ups, om = delta[..., :3], delta[..., 3:]
    th = np.linalg.norm(om, axis=-1, keepdims=True)
    qd = quat_from_rotvec(om)
    th2 = th * th
    small = th < 1e-8
    ths = np.where(small, 1.0, th)
    c1 = np.where(small, 0.5, (1.0 - np.cos(ths)) / (ths * ths))
    c2 = np.where(small, 1.0 / 6.0, (ths - np.sin(ths)) / (ths ** 3))
    wxu = np.cross(om, ups)
    td = ups + c1 * wxu + c2 * np.cross(om, wxu)
    q = quat_mul(qd, pose[..., :4])
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    t = td + quat_rotate(qd, pose[..., 4:])
    del th2
    return np.concatenate([q, t], axis=-1)


def transform(pose: np.ndarray, pts: np.ndarray) -> np.ndarray:
    """pose [...,7] applied to pts [n,3] -> [..., n, 3]."""
    q = pose[..., None, :4]
    t = pose[..., None, 4:]
    return quat_rotate(q, pts) + t


# ----------------------------------------------------------------------------- camera model
def distort(model: str, x: np.ndarray, y: np.ndarray, d) -> tuple[np.ndarray, np.ndarray]:
    """Vectorised DistParam::distortCamPoint for the 7 models (/root/reference/include/dist_models/*.hpp)."""
    r2 = x * x + y * y
    if model in ("rad1", "rad2", "rad3"):
        k = list(d) + [0.0] * (3 - len(d))
        D = 1.0 + r2 * (k[0] + k[1] * r2 + k[2] * r2 * r2)
        return x * D, y * D
    if model in ("radtan4", "radtan5", "radtan8"):
        if model == "radtan4":
            k1, k2, p1, p2 = d
            D = 1.0 + r2 * (k1 + k2 * r2)
        elif model == "radtan5":
            k1, k2, k3, p1, p2 = d
            D = 1.0 + r2 * (k1 + k2 * r2 + k3 * r2 * r2)
        else:
            k1, k2, k3, k4, k5, k6, p1, p2 = d
            D = (1.0 + r2 * (k1 + k2 * r2 + k3 * r2 * r2)) / (1.0 + r2 * (k4 + k5 * r2 + k6 * r2 * r2))
        xy2 = 2.0 * x * y
        return x * D + p1 * xy2 + p2 * (r2 + 2.0 * x * x), y * D + p2 * xy2 + p1 * (r2 + 2.0 * y * y)
    if model == "kannalabrandt":
        r = np.sqrt(r2)
        th = np.arctan(r)
        th2 = th * th
        D = th * (1.0 + th2 * (d[0] + th2 * d[1] + th2 * th2 * d[2] + th2 * th2 * th2 * d[3]))
        with np.errstate(invalid="ignore", divide="ignore"):
            s = D / r
        return x * s, y * s
    raise ValueError(model)


def project(model: str, fx: float, fy: float, cx: float, cy: float, d, pose: np.ndarray, pts: np.ndarray) -> np.ndarray:
    pc = transform(pose, pts)
    x = pc[..., 0] / pc[..., 2]
    y = pc[..., 1] / pc[..., 2]
    xd, yd = distort(model, x, y, d)
    return np.stack([fx * xd + cx, fy * yd + cy], axis=-1)


def prior_intrinsics(width: int, height: int, prior_fov_deg: float) -> tuple[float, float, float]:
    """Camera::setupInitialCalibration: cx=W/2, cy=H/2, f=max(cx,cy)/tan(fov/2)."""
    cx, cy = width / 2.0, height / 2.0
    PI = 2.0 * math.asin(1.0)
    inv_tan = 1.0 / math.tan(prior_fov_deg / 2.0 * (PI / 180.0))
    return max(cx * inv_tan, cy * inv_tan), cx, cy


# ----------------------------------------------------------------------------- LM-only workloads
@dataclass
class LmProblem:
    model: str
    mono: bool
    width: int
    height: int
    target: np.ndarray            # [n_pts,3] f64
    obs: np.ndarray               # [n_views,n_pts,2] f32
    poses_init: np.ndarray        # [n_views,7] f64
    focal_init: np.ndarray        # [1|2]
    pp_init: np.ndarray           # [2]
    dist_init: np.ndarray         # [nd]
    gt: dict = field(default_factory=dict)

    @property
    def n_views(self) -> int:
        return self.obs.shape[0]

    @property
    def n_pts(self) -> int:
        return self.obs.shape[1]


def sample_poses(rng: np.random.Generator, n: int, target: np.ndarray, model: str, f: float, cx: float, cy: float,
                 dist, width: int, height: int, roll_range: float = 0.5, margin: float = 10.0,
                 tilt_sigma: float = 0.35) -> np.ndarray:
    """World->camera poses that keep every target point >= margin px inside the image."""
    centre = target.mean(axis=0)
    extent = np.linalg.norm(target.max(axis=0) - target.min(axis=0))
    out = np.empty((n, 7))
    todo = np.arange(n)
    # distance so that the target fills a good part of the image
    base = extent * f / (0.75 * min(width, height))
    for _ in range(200):
        m = todo.size
        if m == 0:
            break
        tilt = np.clip(rng.normal(0.0, tilt_sigma, size=(m, 2)), -0.9, 0.9)
        roll = rng.uniform(-roll_range, roll_range, size=m)
        depth = base * rng.uniform(0.9, 1.8, size=m)
        lat = rng.uniform(-0.35, 0.35, size=(m, 2)) * depth[:, None] * np.array([width, height]) / f
        # R = Rz(roll) Rx(tilt0) Ry(tilt1) composed with a pi-rotation about x so that the
        # target's +y (rows) points down in the image and z = 1 plane faces the camera.
        q = quat_mul(quat_from_rotvec(np.stack([0 * roll, 0 * roll, roll], -1)),
                     quat_mul(quat_from_rotvec(np.stack([tilt[:, 0], 0 * roll, 0 * roll], -1)),
                              quat_from_rotvec(np.stack([0 * roll, tilt[:, 1], 0 * roll], -1))))
        t = np.stack([lat[:, 0], lat[:, 1], depth], -1) - quat_rotate(q, centre[None, :])
        pose = np.concatenate([q, t], -1)
        px = project(model, f, f, cx, cy, dist, pose, target)
        pc = transform(pose, target)
        ok = ((px[..., 0] > margin) & (px[..., 0] < width - 1 - margin) & (px[..., 1] > margin) &
              (px[..., 1] < height - 1 - margin) & (pc[..., 2] > 0.05)).all(axis=1)
        out[todo[ok]] = pose[ok]
        todo = todo[~ok]
    if todo.size:
        raise RuntimeError(f"could not place {todo.size} poses inside the image")
    return out


def make_lm_problem(model: str, mono: bool, n_views: int, target: np.ndarray, width: int, height: int,
                    seed: int = MASTER_SEED, noise_px: float = 0.08, prior_fov_deg: float | None = None,
                    f_gt: float | None = None, pose_noise=(0.004, 0.01), drop_fraction: float = 0.0,
                    roll_range: float = 0.5) -> LmProblem:
    """LM-only workload (cfg5 style): exact projection + N(0, noise_px) stored float32,
    initial poses = GT o exp(N(0,[4 mm]^3,[0.01 rad]^3)), intrinsics from the prior-FOV rule,
    distortion 0.  `drop_fraction` of the corners are replaced by (-1,-1) (missing AprilGrid
    corners, /root/reference/include/target_detectors/aprilgrid_detector.hpp:35)."""
    rng = np.random.Generator(np.random.Philox(key=seed))
    d_gt = GT_DIST[model]
    if f_gt is None:
        f_gt = 600.0 * width / 1920.0 if model == "kannalabrandt" else 900.0 * width / 1280.0
    cx_gt, cy_gt = width / 2.0 + 5.0, height / 2.0 - 7.0
    fx_gt, fy_gt = (f_gt, f_gt) if mono else (f_gt, f_gt * 1.004)
    poses = sample_poses(rng, n_views, target, model, f_gt, cx_gt, cy_gt, d_gt, width, height, roll_range=roll_range)
    px = project(model, fx_gt, fy_gt, cx_gt, cy_gt, d_gt, poses, target)
    px = px + rng.normal(0.0, noise_px, size=px.shape)
    obs = px.astype(np.float32)
    if drop_fraction > 0:
        drop = rng.uniform(size=obs.shape[:2]) < drop_fraction
        obs[drop] = -1.0
    delta = np.concatenate([rng.normal(0.0, pose_noise[0], size=(n_views, 3)),
                            rng.normal(0.0, pose_noise[1], size=(n_views, 3))], -1)
    poses_init = se3_exp_left(delta, poses)
    if prior_fov_deg is None:
        prior_fov_deg = 2.0 * math.degrees(math.atan(max(width, height) / 2.0 / f_gt)) * 1.08
    f0, cx0, cy0 = prior_intrinsics(width, height, prior_fov_deg)
    focal0 = np.array([f0] if mono else [f0, f0])
    return LmProblem(model=model, mono=mono, width=width, height=height, target=np.ascontiguousarray(target),
                     obs=np.ascontiguousarray(obs), poses_init=np.ascontiguousarray(poses_init), focal_init=focal0,
                     pp_init=np.array([cx0, cy0]), dist_init=np.zeros(NUM_DIST[MODEL_ID[model]]),
                     gt=dict(focal=(fx_gt,) if mono else (fx_gt, fy_gt), pp=(cx_gt, cy_gt), dist=d_gt, poses=poses))


# ----------------------------------------------------------------------------- synthetic renders
def undistort(model: str, xd: np.ndarray, yd: np.ndarray, d, iters: int = 12) -> tuple[np.ndarray, np.ndarray]:
    """Invert `distort` by Newton iterations with a central-difference Jacobian (the scheme of
    DistParam::undistortCamPoint, /root/reference/include/dist_models/dist_model_base.hpp:33-71)."""
    x, y = xd.copy(), yd.copy()
    for _ in range(iters):
        h = 1e-6
        fx0, fy0 = distort(model, x, y, d)
        ax, ay = distort(model, x + h, y, d)
        bx, by = distort(model, x - h, y, d)
        cx_, cy_ = distort(model, x, y + h, d)
        dx_, dy_ = distort(model, x, y - h, d)
        j00, j10 = (ax - bx) / (2 * h), (ay - by) / (2 * h)
        j01, j11 = (cx_ - dx_) / (2 * h), (cy_ - dy_) / (2 * h)
        ex, ey = fx0 - xd, fy0 - yd
        det = j00 * j11 - j01 * j10
        x = x - (j11 * ex - j01 * ey) / det
        y = y - (-j10 * ex + j00 * ey) / det
    return x, y


def chessboard_pattern(nb_rows: int, nb_cols: int, square_size: float):
    """White(1)/black(0) lookup on the target plane for the board whose inner corners are
    chessboard_target(...): squares span x in [-nb_cols*sq, 0], y in [0, nb_rows*sq]; the corner
    square next to target point 0 is black (cv::findChessboardCorners orders from a black corner)."""
    sq = square_size

    def fn(X, Y):
        j = np.floor(-X / sq).astype(np.int64)
        i = np.floor(Y / sq).astype(np.int64)
        inside = (i >= 0) & (i < nb_rows) & (j >= 0) & (j < nb_cols)
        black = ((i + (nb_cols - 1 - j)) % 2 == 0)
        return np.where(inside & black, 0.0, 1.0)
    return fn


def aprilgrid_pattern(nb_rows: int, nb_cols: int, tag_size: float, tag_space: float, codes: np.ndarray):
    """AprilGrid (Kalibr layout, 36h11, 2-cell black border): tag id = r*nb_cols + c at
    offset (c*off, -r*off); the tag spans x in [0,s], y in [-s,0]; data bit i (MSB first) is row-major from
    the top-left cell, 1 = white."""
    s = tag_size
    off = s * (1.0 + tag_space)
    cs = s / 10.0
    bits = np.zeros((nb_rows * nb_cols, 10, 10), dtype=np.float64)
    for t in range(nb_rows * nb_cols):
        code = int(codes[t])
        for idx in range(36):
            bits[t, 2 + idx // 6, 2 + idx % 6] = (code >> (35 - idx)) & 1

    def fn(X, Y):
        c = np.floor(X / off).astype(np.int64)
        r = np.floor(-Y / off + 1e-12).astype(np.int64)
        # tag (r,c) covers y in [-r*off - s, -r*off]
        r = np.floor((-Y) / off).astype(np.int64)
        lx = X - c * off
        ly = Y + r * off            # in (-off, 0]
        inside = (c >= 0) & (c < nb_cols) & (r >= 0) & (r < nb_rows) & (lx >= 0) & (lx < s) & (ly > -s) & (ly <= 0)
        cx_ = np.clip(np.floor(lx / cs).astype(np.int64), 0, 9)
        ry = np.clip(np.floor((ly + s) / cs).astype(np.int64), 0, 9)
        tid = np.clip(r * nb_cols + c, 0, nb_rows * nb_cols - 1)
        return np.where(inside, bits[tid, ry, cx_], 1.0)
    return fn


def gaussian_blur_np(img: np.ndarray, sigma: float) -> np.ndarray:
    if sigma <= 0:
        return img
    rad = max(1, int(math.ceil(3 * sigma)))
    xs = np.arange(-rad, rad + 1)
    k = np.exp(-xs * xs / (2 * sigma * sigma))
    k /= k.sum()
    pad = np.pad(img, rad, mode="edge")
    tmp = sum(k[i] * pad[:, i:i + img.shape[1]] for i in range(2 * rad + 1))
    out = sum(k[i] * tmp[i:i + img.shape[0], :] for i in range(2 * rad + 1))
    return out


def render_view(pattern_fn, model: str, fx: float, fy: float, cx: float, cy: float, d, pose: np.ndarray, width: int,
                height: int, rng: np.random.Generator | None = None, ss: int = 4, blur_sigma: float = 0.8,
                noise_sigma: float = 2.0, black: float = 20.0, white: float = 235.0) -> np.ndarray:
    """Supersampled (ss x ss) area-sampled u8 render of a planar target (z_w = 1) through the camera model."""
    q = pose[:4]
    t = pose[4:]
    Rm = np.array([quat_rotate(q, e) for e in np.eye(3)]).T  # columns = R e_i
    # target-plane coordinates on the lattice of pixel CORNERS (u = i - 0.5); the sub-samples inside a
    # pixel interpolate them bilinearly (the map is smooth: error << 1e-3 px)
    u = np.arange(width + 1, dtype=np.float64)[None, :] - 0.5
    v = np.arange(height + 1, dtype=np.float64)[:, None] - 0.5
    xd = np.broadcast_to((u - cx) / fx, (height + 1, width + 1)).copy()
    yd = np.broadcast_to((v - cy) / fy, (height + 1, width + 1)).copy()
    x, y = undistort(model, xd, yd, d, iters=8)
    # ray lambda*(x,y,1) = R p_w + t with p_w.z = 1  ->  p_w = R^T (lambda ray - t)
    r2x, r2y, r2z = Rm[0, 2], Rm[1, 2], Rm[2, 2]
    num = 1.0 + (r2x * t[0] + r2y * t[1] + r2z * t[2])
    den = r2x * x + r2y * y + r2z
    lam = num / den
    px, py, pz = lam * x - t[0], lam * y - t[1], lam - t[2]
    Xc = Rm[0, 0] * px + Rm[1, 0] * py + Rm[2, 0] * pz
    Yc = Rm[0, 1] * px + Rm[1, 1] * py + Rm[2, 1] * pz
    behind = (lam <= 0)
    Xc = np.where(behind, 1e9, Xc)
    Yc = np.where(behind, 1e9, Yc)
    acc = np.zeros((height, width))
    X00, X01, X10, X11 = Xc[:-1, :-1], Xc[:-1, 1:], Xc[1:, :-1], Xc[1:, 1:]
    Y00, Y01, Y10, Y11 = Yc[:-1, :-1], Yc[:-1, 1:], Yc[1:, :-1], Yc[1:, 1:]
    for a in range(ss):
        wa = (a + 0.5) / ss
        for b in range(ss):
            wb = (b + 0.5) / ss
            X = (1 - wb) * ((1 - wa) * X00 + wa * X01) + wb * ((1 - wa) * X10 + wa * X11)
            Y = (1 - wb) * ((1 - wa) * Y00 + wa * Y01) + wb * ((1 - wa) * Y10 + wa * Y11)
            acc += pattern_fn(X, Y)
    img = black + (white - black) * acc / (ss * ss)
    img = gaussian_blur_np(img, blur_sigma)
    if noise_sigma > 0 and rng is not None:
        img = img + rng.normal(0.0, noise_sigma, size=img.shape)
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


# ----------------------------------------------------------------------------- SE3 log / inverse / product
def se3_inverse(pose: np.ndarray) -> np.ndarray:
    q = pose[..., :4] * np.array([-1.0, -1.0, -1.0, 1.0])
    t = -quat_rotate(q, pose[..., 4:])
    return np.concatenate([q, t], axis=-1)


def se3_mul(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    q = quat_mul(a[..., :4], b[..., :4])
    q = q / np.linalg.norm(q, axis=-1, keepdims=True)
    t = a[..., 4:] + quat_rotate(a[..., :4], b[..., 4:])
    return np.concatenate([q, t], axis=-1)


def se3_log(pose: np.ndarray) -> np.ndarray:
    """Sophus::SE3d::log -> (upsilon, omega) for a single pose [7]."""
    q = np.asarray(pose[:4], dtype=np.float64)
    t = np.asarray(pose[4:], dtype=np.float64)
    n2 = float(q[:3] @ q[:3])
    n = math.sqrt(n2)
    w = float(q[3])
    if n2 < 1e-20:
        two_atan = 2.0 / w - (2.0 / 3.0) * n2 / (w * w * w)
        theta = 2.0 * n2 / w if False else two_atan * n
    else:
        if abs(w) < 1e-10:
            two_atan = (math.pi if w >= 0 else -math.pi) / n
        else:
            two_atan = 2.0 * math.atan(n / w) / n
        theta = two_atan * n
    om = two_atan * q[:3]
    Om = np.array([[0, -om[2], om[1]], [om[2], 0, -om[0]], [-om[1], om[0], 0]])
    if abs(theta) < 1e-10:
        Vinv = np.eye(3) - 0.5 * Om + (1.0 / 12.0) * (Om @ Om)
    else:
        half = 0.5 * theta
        Vinv = np.eye(3) - 0.5 * Om + (1.0 - theta * math.cos(half) / (2.0 * math.sin(half))) / (theta * theta) * (Om @ Om)
    return np.concatenate([Vinv @ t, om])


def pose_to_matrix34(pose: np.ndarray) -> np.ndarray:
    """Sophus::SE3d::matrix3x4() of a pose stored as qx qy qz qw tx ty tz."""
    pose = np.asarray(pose, np.float64)
    R = np.stack([quat_rotate(pose[:4], e) for e in np.eye(3)], axis=1)
    return np.concatenate([R, pose[4:7, None]], axis=1)

"""Synthetic calibration data: default AprilGrid board, random board poses,
exact projections rounded to f32 like the reference's FeaturePoint storage.

Reference: src/board.rs:46-95 (Board::init_aprilgrid, computed in f32),
src/detected_points.rs:6-9 (p2d/p3d stored as f32), src/optimization/factors.rs:141-143 (f32 -> f64),
data/eucm.json (EUCM ground truth, scaled x2 to a 1024x1024 image).

Pure numpy: this is input generation for tests/bench, not part of the GPU product path.
The numpy projection here is a third, independent statement of the six camera models.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

MODELS = {"ucm": 0, "eucm": 1, "eucmt": 2, "kb4": 3, "opencv5": 4, "ftheta": 5}
MODEL_NPARAMS = {0: 5, 1: 6, 2: 8, 3: 8, 4: 9, 5: 8}

# data/eucm.json (512x512) scaled x2
EUCM_GT = np.array([381.79237374367876, 381.74044571764734, 509.8750740963924, 513.7282896612157,
                    0.6283550447635853, 1.0458678747533083])

GT_PARAMS = {
    "ucm": np.array([381.79, 381.74, 509.875, 513.73, 0.63]),
    "eucm": np.array([381.79, 381.74, 509.875, 513.73, 0.6283550447635853, 1.0458678747533083]),
    "eucmt": np.array([381.79, 381.74, 509.875, 513.73, 0.6283550447635853, 1.0458678747533083, 1e-3, -5e-4]),
    "kb4": np.array([380.0, 380.0, 509.875, 513.73, 0.01, -0.002, 3e-4, -4e-5]),
    "opencv5": np.array([600.0, 600.0, 509.875, 513.73, -0.1, 0.05, 1e-3, -1e-3, -0.01]),
    "ftheta": np.array([380.0, 380.0, 509.875, 513.73, 0.01, -0.002, 3e-4, -4e-5]),
}

# multiplicative perturbation of the start point: (fx, fy, cx, cy, dist...) — SURVEY §8(d)
INIT_SCALE = {
    "ucm": np.array([1.05, 1.05, 1.01, 0.99, 0.95]),
    "eucm": np.array([1.05, 1.05, 1.01, 0.99, 0.95, 1.10]),
    "eucmt": np.array([1.05, 1.05, 1.01, 0.99, 0.95, 1.10, 0.5, 0.5]),
    "kb4": np.array([1.05, 1.05, 1.01, 0.99, 0.5, 0.5, 0.5, 0.5]),
    "opencv5": np.array([1.05, 1.05, 1.01, 0.99, 0.8, 0.8, 0.5, 0.5, 0.5]),
    "ftheta": np.array([1.05, 1.05, 1.01, 0.99, 0.5, 0.5, 0.5, 0.5]),
}


def aprilgrid_board(tag_size=0.088, tag_spacing=0.3, rows=6, cols=6) -> np.ndarray:
    """Board::init_aprilgrid in f32 arithmetic, corner order TL,TR,BR,BL per tag (board.rs:46-95).
    Returns (rows*cols*4, 3) float32."""
    ts = np.float32(tag_size)
    sp = np.float32(1.0) + np.float32(tag_spacing)
    pts = []
    for r in range(rows):
        for c in range(cols):
            sx = np.float32(c) * ts * sp
            sy = -(np.float32(r)) * ts * sp
            pts += [(sx, sy, 0.0), (sx + ts, sy, 0.0), (sx + ts, sy - ts, 0.0), (sx, sy - ts, 0.0)]
    return np.asarray(pts, dtype=np.float32)


def rodrigues(rvec: np.ndarray) -> np.ndarray:
    """(…,3) axis-angle -> (…,3,3) rotation matrices."""
    rvec = np.asarray(rvec, dtype=np.float64)
    th = np.linalg.norm(rvec, axis=-1)[..., None, None]
    k = np.zeros(rvec.shape[:-1] + (3, 3))
    k[..., 0, 1] = -rvec[..., 2]; k[..., 0, 2] = rvec[..., 1]
    k[..., 1, 0] = rvec[..., 2]; k[..., 1, 2] = -rvec[..., 0]
    k[..., 2, 0] = -rvec[..., 1]; k[..., 2, 1] = rvec[..., 0]
    with np.errstate(invalid="ignore", divide="ignore"):
        a = np.where(th > 1e-12, np.sin(th) / th, 1.0)
        b = np.where(th > 1e-12, (1.0 - np.cos(th)) / (th * th), 0.5)
    return np.eye(3) + a * k + b * (k @ k)


def rotmat_to_rvec(R: np.ndarray) -> np.ndarray:
    """(3,3) rotation matrix -> axis-angle (the log map; inverse of `rodrigues`, angles in [0, pi])."""
    R = np.asarray(R, dtype=np.float64)
    w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    c = np.clip((np.trace(R) - 1.0) * 0.5, -1.0, 1.0)
    s = 0.5 * np.linalg.norm(w)
    th = np.arctan2(s, c)
    if s > 1e-9:
        return w * (th / (2.0 * s))
    if c > 0.0:                      # near the identity: log R ~ (R - R^T)/2
        return 0.5 * w
    # near pi: axis from the largest diagonal entry of (R + I)/2 = a a^T
    A = 0.5 * (R + np.eye(3))
    i = int(np.argmax(np.diag(A)))
    a = A[:, i] / np.sqrt(max(A[i, i], 1e-300))
    return a * th


def project(model: str | int, prm: np.ndarray, P: np.ndarray) -> np.ndarray:
    """numpy projection of camera-frame points P (…,3) with FULL parameter vector prm. Returns (…,2)."""
    m = MODELS[model] if isinstance(model, str) else int(model)
    x, y, z = P[..., 0], P[..., 1], P[..., 2]
    fx, fy, cx, cy = prm[0], prm[1], prm[2], prm[3]
    if m in (0, 1, 2):
        alpha = prm[4]
        beta = 1.0 if m == 0 else prm[5]
        r2 = x * x + y * y
        rho = np.sqrt(beta * r2 + z * z)
        nrm = alpha * rho + (1.0 - alpha) * z
        mx, my = x / nrm, y / nrm
        if m == 2:
            t1, t2 = prm[6], prm[7]
            rr = mx * mx + my * my
            mx, my = (mx + 2 * t1 * mx * my + t2 * (rr + 2 * mx * mx),
                      my + t1 * (rr + 2 * my * my) + 2 * t2 * mx * my)
    elif m in (3, 5):
        r = np.sqrt(x * x + y * y)
        th = np.arctan2(r, z)
        k1, k2, k3, k4 = prm[4:8]
        if m == 3:
            d = th * (1 + k1 * th**2 + k2 * th**4 + k3 * th**6 + k4 * th**8)
        else:
            d = th * (1 + k1 * th + k2 * th**2 + k3 * th**3 + k4 * th**4)
        with np.errstate(invalid="ignore", divide="ignore"):
            s = np.where(r < 1e-8, 1.0 / z, d / r)
        mx, my = x * s, y * s
    elif m == 4:
        k1, k2, p1, p2, k3 = prm[4:9]
        a, b = x / z, y / z
        r2 = a * a + b * b
        rad = 1 + k1 * r2 + k2 * r2 * r2 + k3 * r2**3
        mx = a * rad + 2 * p1 * a * b + p2 * (r2 + 2 * a * a)
        my = b * rad + p1 * (r2 + 2 * b * b) + 2 * p2 * a * b
    else:
        raise ValueError(model)
    return np.stack([fx * mx + cx, fy * my + cy], axis=-1)


@dataclass
class SyntheticCalib:
    """One single-camera calibration problem in the library's SoA layout."""
    model: str
    width: int
    height: int
    frame_offsets: np.ndarray      # int32 (F+1)
    x: np.ndarray; y: np.ndarray; z: np.ndarray   # board points per observation (f64, f32-representable)
    u: np.ndarray; v: np.ndarray                  # observations (f64, f32-representable)
    gt_params: np.ndarray          # full parameter vector
    gt_poses: np.ndarray           # (F,6) rvec,tvec
    init_params: np.ndarray        # full parameter vector, perturbed
    init_poses: np.ndarray         # (F,6) perturbed
    extra: dict = field(default_factory=dict)

    @property
    def n_frames(self) -> int:
        return len(self.frame_offsets) - 1

    @property
    def n_obs(self) -> int:
        return int(self.frame_offsets[-1])


def make_poses(rng: np.random.Generator, n_frames: int, board_centre: np.ndarray, max_angle=0.6,
               xy_range=0.25, z_range=(0.35, 0.9)):
    axis = rng.normal(size=(n_frames, 3))
    axis /= np.linalg.norm(axis, axis=1, keepdims=True)
    rvec = axis * rng.uniform(0.0, max_angle, size=(n_frames, 1))
    c = np.stack([rng.uniform(-xy_range, xy_range, n_frames), rng.uniform(-xy_range, xy_range, n_frames),
                  rng.uniform(z_range[0], z_range[1], n_frames)], axis=1)
    R = rodrigues(rvec)
    tvec = c - np.einsum("fij,j->fi", R, board_centre)
    return rvec, tvec


def make_calib(model: str = "eucm", n_frames: int = 100, seed: int = 0, noise_px: float = 0.0,
               width: int = 1024, height: int = 1024, drop_fraction: float = 0.0,
               gt_params: np.ndarray | None = None) -> SyntheticCalib:
    """SURVEY §8(d) generator. drop_fraction>0 makes frames ragged (random corners missing)."""
    rng = np.random.default_rng(seed)
    board = aprilgrid_board().astype(np.float64)            # f32 values widened
    centre = board.mean(axis=0)
    gt = np.array(GT_PARAMS[model] if gt_params is None else gt_params, dtype=np.float64)
    if model == "opencv5":
        rvec, tvec = make_poses(rng, n_frames, centre, max_angle=0.4, xy_range=0.12, z_range=(0.6, 0.9))
    else:
        rvec, tvec = make_poses(rng, n_frames, centre)
    R = rodrigues(rvec)
    P = np.einsum("fij,kj->fki", R, board) + tvec[:, None, :]          # (F,144,3)
    uv = project(model, gt, P)
    if noise_px > 0:
        uv = uv + rng.normal(scale=noise_px, size=uv.shape)
    uv = uv.astype(np.float32).astype(np.float64)                       # FeaturePoint.p2d is f32
    inside = (uv[..., 0] >= 0) & (uv[..., 0] < width) & (uv[..., 1] >= 0) & (uv[..., 1] < height) & (P[..., 2] > 0.05)
    if drop_fraction > 0:
        inside &= rng.uniform(size=inside.shape) >= drop_fraction
    counts = inside.sum(axis=1)
    keep = counts >= 24                                                  # data_loader.rs:15 MIN_CORNERS
    rvec, tvec, inside, uv = rvec[keep], tvec[keep], inside[keep], uv[keep]
    counts = counts[keep]
    offs = np.zeros(len(counts) + 1, dtype=np.int32)
    np.cumsum(counts, out=offs[1:])
    fi, ki = np.nonzero(inside)
    x, y, z = board[ki, 0].copy(), board[ki, 1].copy(), board[ki, 2].copy()
    u, v = uv[fi, ki, 0].copy(), uv[fi, ki, 1].copy()
    poses = np.concatenate([rvec, tvec], axis=1)
    init_params = gt * INIT_SCALE[model]
    init_poses = poses.copy()
    init_poses[:, :3] += rng.normal(scale=0.01, size=(len(poses), 3))
    init_poses[:, 3:] += rng.normal(scale=0.005, size=(len(poses), 3))
    return SyntheticCalib(model=model, width=width, height=height, frame_offsets=offs, x=x, y=y, z=z, u=u, v=v,
                          gt_params=gt, gt_poses=poses, init_params=init_params, init_poses=init_poses,
                          extra={"corner_id": ki.astype(np.int32), "board": board.astype(np.float32)})


def to_frame_features(s: SyntheticCalib, poses: np.ndarray | None = None):
    """The problem as the reference's types: ([FrameFeature], {frame_idx: RvecTvec}) (detected_points.rs:6-17,
    types.rs:13-17) with `poses` (default: the perturbed initial poses) as the per-frame initial guess."""
    from .calib import FeaturePoint, FrameFeature, RvecTvec
    poses = s.init_poses if poses is None else poses
    frames, init = [], {}
    for f in range(s.n_frames):
        a, b = s.frame_offsets[f], s.frame_offsets[f + 1]
        feats = {k: FeaturePoint((s.u[a + k], s.v[a + k]), (s.x[a + k], s.y[a + k], s.z[a + k])) for k in range(b - a)}
        frames.append(FrameFeature(0, (s.width, s.height), feats))
        init[f] = RvecTvec(tuple(poses[f, :3]), tuple(poses[f, 3:]))
    return frames, init


def intr_from_full(params: np.ndarray, xy_same_focal: bool) -> np.ndarray:
    """calib_camera's `params.remove_row(1)` when --one-focal (util.rs:391-395)."""
    p = np.asarray(params, dtype=np.float64)
    return np.delete(p, 1) if xy_same_focal else p.copy()


def full_from_intr(intr: np.ndarray, xy_same_focal: bool) -> np.ndarray:
    """`new_params.insert_row(1, new_params[0])` (util.rs:466-470)."""
    a = np.asarray(intr, dtype=np.float64)
    return np.insert(a, 1, a[0]) if xy_same_focal else a.copy()


@dataclass
class SyntheticRig:
    """Joint multi-camera problem (calib_all_camera_with_extrinsics, src/util.rs:567-715) in block-CSR SoA layout."""
    model: str
    width: int
    height: int
    n_cams: int
    n_frames: int
    block_cam: np.ndarray        # int32 (B,)
    block_frame: np.ndarray      # int32 (B,)
    block_offsets: np.ndarray    # int32 (B+1,)
    x: np.ndarray; y: np.ndarray; z: np.ndarray; u: np.ndarray; v: np.ndarray
    gt_params: np.ndarray        # (C, nparams) full vectors
    gt_extr: np.ndarray          # (C, 6) T_c_0, row 0 zero
    gt_poses: np.ndarray         # (F, 6) T_0_b
    init_params: np.ndarray
    init_extr: np.ndarray
    init_poses: np.ndarray

    @property
    def n_obs(self) -> int:
        return int(self.block_offsets[-1])


def make_rig(model: str = "eucm", n_frames: int = 200, n_cams: int = 2, seed: int = 4, width: int = 1024,
             height: int = 1024, drop_block_fraction: float = 0.1) -> SyntheticRig:
    """joint cam0+cam1 problem, T_1_0 = rvec (0, 0.02, 0), tvec (-0.1, 0, 0); further cameras are spaced
    another -0.1 m apart. Each camera misses a random `drop_block_fraction` of the frames."""
    rng = np.random.default_rng(seed)
    board = aprilgrid_board().astype(np.float64)
    centre = board.mean(axis=0)
    gt0 = np.array(GT_PARAMS[model], dtype=np.float64)
    if model == "opencv5":
        rvec, tvec = make_poses(rng, n_frames, centre, max_angle=0.4, xy_range=0.12, z_range=(0.6, 0.9))
    else:
        rvec, tvec = make_poses(rng, n_frames, centre, xy_range=0.15)
    gt_params = np.stack([gt0 * (1.0 + 0.01 * c * np.sign(np.arange(len(gt0)) % 2 - 0.5)) for c in range(n_cams)])
    gt_extr = np.zeros((n_cams, 6))
    for c in range(1, n_cams):
        gt_extr[c] = [0.0, 0.02 * c, 0.0, -0.1 * c, 0.0, 0.0]
    R0 = rodrigues(rvec)
    P0 = np.einsum("fij,kj->fki", R0, board) + tvec[:, None, :]
    bc, bf, offs, xs, ys, zs, us, vs = [], [], [0], [], [], [], [], []
    for c in range(n_cams):
        Rc = rodrigues(gt_extr[c, :3])
        Pc = np.einsum("ij,fkj->fki", Rc, P0) + gt_extr[c, 3:]
        uv = project(model, gt_params[c], Pc).astype(np.float32).astype(np.float64)
        inside = (uv[..., 0] >= 0) & (uv[..., 0] < width) & (uv[..., 1] >= 0) & (uv[..., 1] < height) & (Pc[..., 2] > 0.05)
        seen = rng.uniform(size=n_frames) >= drop_block_fraction
        for f in range(n_frames):
            idx = np.nonzero(inside[f])[0]
            if not seen[f] or len(idx) < 24:
                continue
            bc.append(c); bf.append(f); offs.append(offs[-1] + len(idx))
            xs.append(board[idx, 0]); ys.append(board[idx, 1]); zs.append(board[idx, 2])
            us.append(uv[f, idx, 0]); vs.append(uv[f, idx, 1])
    cat = np.concatenate
    poses = cat([rvec, tvec], axis=1)
    # a frame no camera detected is not a variable of the problem (util.rs:588-600 only creates "rvec_0_b_{f}" for
    # frames with a detection): drop it and renumber
    used = np.zeros(n_frames, dtype=bool); used[np.asarray(bf, dtype=np.int64)] = True
    renum = np.cumsum(used) - 1
    bf = [int(renum[f]) for f in bf]
    poses = poses[used]
    n_all, n_frames = n_frames, int(used.sum())
    init_params = gt_params * INIT_SCALE[model][None, :]
    init_extr = gt_extr.copy()
    init_extr[1:, :3] += rng.normal(scale=0.005, size=(n_cams - 1, 3))
    init_extr[1:, 3:] += rng.normal(scale=0.005, size=(n_cams - 1, 3))
    init_poses = poses.copy()
    init_poses[:, :3] += rng.normal(scale=0.01, size=(n_all, 3))[used]
    init_poses[:, 3:] += rng.normal(scale=0.005, size=(n_all, 3))[used]
    return SyntheticRig(model=model, width=width, height=height, n_cams=n_cams, n_frames=n_frames,
                        block_cam=np.asarray(bc, dtype=np.int32), block_frame=np.asarray(bf, dtype=np.int32),
                        block_offsets=np.asarray(offs, dtype=np.int32), x=cat(xs), y=cat(ys), z=cat(zs), u=cat(us), v=cat(vs),
                        gt_params=gt_params, gt_extr=gt_extr, gt_poses=poses, init_params=init_params,
                        init_extr=init_extr, init_poses=init_poses)

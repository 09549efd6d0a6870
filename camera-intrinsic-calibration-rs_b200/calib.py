"""Host-side mirror of the reference's interface for the hot path, over the C ABI.

Mirrors (same names, argument meaning, error behaviour):
  * FeaturePoint / FrameFeature           src/detected_points.rs:6-17
  * RvecTvec                              src/types.rs:13-39
  * GenericModel (params/width/height)    camera-intrinsic-model ^0.8 (call sites src/util.rs:391,418,472)
  * calib_camera(...)                     src/util.rs:384-490  -> Optional[(GenericModel, {frame_idx: RvecTvec})]
plus a thin `Problem` class over the step-wise entry points used by the parity tests and the bench.

Everything numerical happens inside libccrs_b200.so (CUDA). This module only reshapes arrays.
The initial per-frame poses (SQPnP in the reference, util.rs:418-439) are an input here: pose
initialisation is outside the accelerated path.
"""
from __future__ import annotations

import ctypes as C
import sys
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from . import _abi
from ._abi import CcrsError, Options, Summary, check, default_options

MODELS = {"ucm": 0, "eucm": 1, "eucmt": 2, "kb4": 3, "opencv5": 4, "ftheta": 5}
MIN_POINTS_PER_FRAME = 10  # util.rs:431-433


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _up(a):
    return a.ctypes.data_as(C.POINTER(C.c_ubyte)) if a is not None else None


@dataclass
class FeaturePoint:
    p2d: Tuple[float, float]
    p3d: Tuple[float, float, float]


@dataclass
class FrameFeature:
    time_ns: int
    img_w_h: Tuple[int, int]
    features: Dict[int, FeaturePoint] = field(default_factory=dict)


@dataclass
class RvecTvec:
    rvec: Tuple[float, float, float]
    tvec: Tuple[float, float, float]

    def as_array(self) -> np.ndarray:
        return np.array([*self.rvec, *self.tvec], dtype=np.float64)


@dataclass
class GenericModel:
    model: str
    params: np.ndarray
    width: int
    height: int

    def model_id(self) -> int:
        return MODELS[self.model]


class Problem:
    """One device-resident calibration problem (or a batch of independent ones)."""

    def __init__(self, model, width: int, height: int, frame_offsets, x, y, z, u, v, xy_same_focal: bool = False,
                 huber_delta: float = 1.0, device: int = 0, problem_frame_offsets=None, corner_id=None, board=None):
        """corner_id / board: the reference's own data model (FrameFeature.features: corner id -> FeaturePoint whose
        p3d is the board point of that id): x, y, z are then ignored (pass None) and p3d = board[corner_id]."""
        self.lib = _abi.load()
        self.model = MODELS[model] if isinstance(model, str) else int(model)
        fo = np.ascontiguousarray(frame_offsets, dtype=np.int32)
        if corner_id is not None:
            ids = np.ascontiguousarray(corner_id, dtype=np.int32)
            bd = np.ascontiguousarray(board, dtype=np.float32).reshape(-1, 3)
            us, vs = (np.ascontiguousarray(a, dtype=np.float32) for a in (u, v))
            n = int(fo[-1])
            if ids.shape != (n,) or us.shape != (n,) or vs.shape != (n,):
                raise ValueError("corner_id, u, v must have frame_offsets[-1] entries")
            self.h = C.c_void_p()
            fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
            ip = C.POINTER(C.c_int32)
            check(self.lib.ccrs_problem_create_board_f32(C.byref(self.h), self.model, width, height, int(xy_same_focal), len(fo) - 1,
                                                         fo.ctypes.data_as(ip), ids.ctypes.data_as(ip), fp(us), fp(vs), fp(bd),
                                                         len(bd), float(huber_delta), int(device)))
            self._storage = "board"
            self._finish_init()
            return
        f32 = all(isinstance(a, np.ndarray) and a.dtype == np.float32 for a in (x, y, z, u, v)) and problem_frame_offsets is None
        conv = (lambda a: np.ascontiguousarray(a, dtype=np.float32)) if f32 else _f64
        xs, ys, zs, us, vs = map(conv, (x, y, z, u, v))
        n = int(fo[-1])
        for a in (xs, ys, zs, us, vs):
            if a.shape != (n,):
                raise ValueError("observation arrays must have frame_offsets[-1] entries")
        self.h = C.c_void_p()
        ip = C.POINTER(C.c_int32)
        if f32:   # FeaturePoint-style f32 observations: ccrs_problem_create_f32
            fp = lambda a: a.ctypes.data_as(C.POINTER(C.c_float))
            check(self.lib.ccrs_problem_create_f32(C.byref(self.h), self.model, width, height, int(xy_same_focal),
                                                   len(fo) - 1, fo.ctypes.data_as(ip), fp(xs), fp(ys), fp(zs), fp(us),
                                                   fp(vs), float(huber_delta), int(device)))
        elif problem_frame_offsets is None:
            check(self.lib.ccrs_problem_create(C.byref(self.h), self.model, width, height, int(xy_same_focal),
                                               len(fo) - 1, fo.ctypes.data_as(ip), _dp(xs), _dp(ys), _dp(zs), _dp(us),
                                               _dp(vs), float(huber_delta), int(device)))
        else:
            pfo = np.ascontiguousarray(problem_frame_offsets, dtype=np.int32)
            check(self.lib.ccrs_batch_create(C.byref(self.h), self.model, width, height, int(xy_same_focal),
                                             len(pfo) - 1, pfo.ctypes.data_as(ip), len(fo) - 1, fo.ctypes.data_as(ip),
                                             _dp(xs), _dp(ys), _dp(zs), _dp(us), _dp(vs), float(huber_delta), int(device)))
        self._storage = "f32" if f32 else "f64"
        self._finish_init()

    def _finish_init(self):
        self.d = self.lib.ccrs_problem_dim(self.h)
        self.nblk = self.lib.ccrs_problem_nblk(self.h)
        self.n_frames = self.lib.ccrs_problem_n_frames(self.h)
        self.n_obs = self.lib.ccrs_problem_n_obs(self.h)
        self.n_problems = self.lib.ccrs_problem_n_problems(self.h)
        self.nout = self.d * self.d + 3 * self.d + 1

    @classmethod
    def from_synth(cls, s, **kw):
        return cls(s.model, s.width, s.height, s.frame_offsets, s.x, s.y, s.z, s.u, s.v, **kw)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.ccrs_problem_destroy(self.h)
            self.h = None

    def __del__(self):
        # never call into CUDA while the interpreter is shutting down (the runtime may already be tearing down)
        try:
            if sys is None or sys.is_finalizing():
                return
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # ---- state ----
    def update_observations(self, frame_offsets, u, v, x=None, y=None, z=None, corner_id=None):
        """New detections for this handle (same frame count and total observation count): ccrs_problem_update_observations.
        Board-format handles take corner_id; the others x, y, z. Poses and solver state are reset."""
        fo = np.ascontiguousarray(frame_offsets, dtype=np.int32)
        dt = np.float64 if self._storage == "f64" else np.float32
        us, vs = (np.ascontiguousarray(a, dtype=dt) for a in (u, v))
        vp = lambda a: a.ctypes.data_as(C.c_void_p) if a is not None else None
        ids = np.ascontiguousarray(corner_id, dtype=np.int32) if corner_id is not None else None
        xs, ys, zs = ((np.ascontiguousarray(a, dtype=dt) if a is not None else None) for a in (x, y, z))
        for a in (us, vs, ids, xs, ys, zs):
            if a is not None and a.shape != (int(fo[-1]),):
                raise ValueError("observation arrays must have frame_offsets[-1] entries")
        check(self.lib.ccrs_problem_update_observations(self.h, fo.ctypes.data_as(C.POINTER(C.c_int32)),
                                                        ids.ctypes.data_as(C.POINTER(C.c_int32)) if ids is not None else None,
                                                        vp(xs), vp(ys), vp(zs), vp(us), vp(vs)))

    def set_poses(self, poses):
        p = _f64(poses).reshape(-1)
        assert p.size == 6 * self.n_frames
        check(self.lib.ccrs_set_poses(self.h, _dp(p)))

    def get_poses(self, out: np.ndarray | None = None) -> np.ndarray:
        """Current poses [n_frames, 6]. `out`: a caller-owned float64 buffer of 6 * n_frames values to download into
        (page-locked memory makes the D2H copy a single DMA instead of a staged copy)."""
        if out is None:
            p = np.empty(6 * self.n_frames)
        else:
            p = out.reshape(-1)
            assert p.dtype == np.float64 and p.size == 6 * self.n_frames and p.flags["C_CONTIGUOUS"]
        check(self.lib.ccrs_get_poses(self.h, _dp(p)))
        return p.reshape(-1, 6)

    # ---- step-wise path ----
    def eval_rj(self, intr, poses=None, apply_loss=True, want_j=True):
        intr = _f64(intr)
        r = np.empty(2 * self.n_obs)
        J = np.empty((2 * self.n_obs, self.d + 6)) if want_j else None
        pp = _f64(poses).reshape(-1) if poses is not None else None
        check(self.lib.ccrs_eval_rj(self.h, _dp(intr), _dp(pp), int(apply_loss), _dp(r), _dp(J)))
        return r, J

    def validation(self, intr, poses=None, want_errors=False):
        """util::validation on the device: (median, mean of the best 99 %[, per-point errors])."""
        intr = _f64(intr)
        pp = _f64(poses).reshape(-1) if poses is not None else None
        med, avg = C.c_double(0.0), C.c_double(0.0)
        e = np.empty(self.n_obs) if want_errors else None
        check(self.lib.ccrs_validation(self.h, _dp(intr), _dp(pp), C.byref(med), C.byref(avg), _dp(e)))
        return (med.value, avg.value, e) if want_errors else (med.value, avg.value)

    def linearize(self, intr, which=0) -> np.ndarray:
        intr = _f64(intr).reshape(-1)
        sq = np.empty(self.n_problems)
        check(self.lib.ccrs_linearize(self.h, _dp(intr), int(which), _dp(sq)))
        return sq

    def frame_blocks(self, which=0) -> np.ndarray:
        b = np.empty((self.n_frames, self.nblk))
        check(self.lib.ccrs_get_frame_blocks(self.h, int(which), _dp(b)))
        return b

    def compute_scale(self, which=0) -> np.ndarray:
        c = np.empty((self.n_problems, self.d))
        check(self.lib.ccrs_compute_scale(self.h, int(which), _dp(c)))
        return c

    def set_intr_scale(self, scale):
        s = _f64(scale).reshape(-1) if scale is not None else None
        check(self.lib.ccrs_set_intr_scale(self.h, _dp(s)))

    def reduce(self, which=0, u=None, use_scale=False, min_diag=1e-6, max_diag=1e32):
        """Returns dict of per-problem arrays S (P,d,d), g_s, g_a, diag_a (P,d), sq_err (P,)."""
        uu = _f64(np.broadcast_to(u, (self.n_problems,))) if u is not None else None
        out = np.empty((self.n_problems, self.nout))
        check(self.lib.ccrs_reduce(self.h, int(which), _dp(uu), int(use_scale), float(min_diag), float(max_diag), _dp(out)))
        d = self.d
        return dict(S=out[:, :d * d].reshape(-1, d, d), g_s=out[:, d * d:d * d + d], g_a=out[:, d * d + d:d * d + 2 * d],
                    diag_a=out[:, d * d + 2 * d:d * d + 3 * d], sq_err=out[:, -1])

    def backsub(self, y_a, u=None, in_place=False, want_model_dec=True):
        y = _f64(y_a).reshape(-1)
        uu = _f64(np.broadcast_to(u, (self.n_problems,))) if u is not None else None
        md = np.empty(self.n_problems) if want_model_dec else None
        check(self.lib.ccrs_backsub(self.h, _dp(y), _dp(uu), int(in_place), _dp(md)))
        return md

    def eval_cost(self, intr, which=0) -> np.ndarray:
        intr = _f64(intr).reshape(-1)
        sq = np.empty(self.n_problems)
        check(self.lib.ccrs_eval_cost(self.h, _dp(intr), int(which), _dp(sq)))
        return sq

    def accept(self, mask=None):
        m = np.ascontiguousarray(mask, dtype=np.uint8) if mask is not None else None
        check(self.lib.ccrs_accept(self.h, _up(m)))

    # ---- controllers ----
    def _solve(self, fn, intr, lo, hi, fixed, options):
        a = _f64(intr).reshape(-1).copy()
        assert a.size == self.n_problems * self.d
        o = options or default_options()
        s = Summary()
        hist = np.full(max(o.max_iteration, 1), np.nan)
        lo_a = _f64(lo) if lo is not None else None
        hi_a = _f64(hi) if hi is not None else None
        fx = np.ascontiguousarray(fixed, dtype=np.uint8) if fixed is not None else None
        code = fn(self.h, _dp(a), _dp(lo_a), _dp(hi_a), _up(fx), C.byref(o), C.byref(s), _dp(hist))
        if code not in (0, -4, -5):
            check(code)
        out = a.reshape(self.n_problems, self.d) if self.n_problems > 1 else a
        return out, s, hist[: s.iterations]

    def solve_gn(self, intr, lo=None, hi=None, fixed=None, options: Optional[Options] = None):
        """GaussNewtonOptimizer::optimize on the device state. Returns (intr, Summary, err_hist)."""
        return self._solve(self.lib.ccrs_solve_gn, intr, lo, hi, fixed, options)

    def solve_lm(self, intr, lo=None, hi=None, fixed=None, options: Optional[Options] = None):
        return self._solve(self.lib.ccrs_solve_lm, intr, lo, hi, fixed, options)

    # ---- multi-GPU ----
    def comm_init(self, unique_id, rank: int = 0, world: int = 1, deterministic: bool = True):
        """unique_id=None attaches the process-wide communicator created by an earlier call."""
        buf = C.create_string_buffer(unique_id, 128) if unique_id is not None else None
        check(self.lib.ccrs_comm_init(self.h, buf, rank, world))
        check(self.lib.ccrs_comm_set_deterministic(self.h, int(deterministic)))

    # ---- measurement ----
    def time_linearize(self, intr, reps=20, flush_l2=False) -> float:
        intr = _f64(intr).reshape(-1)
        ms = C.c_double(0.0)
        check(self.lib.ccrs_time_linearize(self.h, _dp(intr), int(reps), int(flush_l2), C.byref(ms)))
        return ms.value

    def bench_lm_steps(self, intr0, poses0, warmup=3, steps=10, reset_every=4, flush_l2=True):
        """Returns (step_ms[steps], kernel launches inside the timed steps)."""
        a = _f64(intr0).reshape(-1)
        pz = _f64(poses0).reshape(-1)
        ms = np.zeros(steps)
        n = C.c_int64(0)
        check(self.lib.ccrs_bench_lm_steps(self.h, _dp(a), _dp(pz), int(warmup), int(steps), int(reset_every),
                                           int(flush_l2), _dp(ms), C.byref(n)))
        return ms, int(n.value)

    @staticmethod
    def bench_lm_steps_rotating(replicas, intr0, poses0, warmup=3, steps=20):
        """`replicas`: Problem handles of one shape (together larger than the L2 cache). Returns (total ms of the
        timed slots, kernel launches inside them, LM iterations they really executed)."""
        lib = replicas[0].lib
        a = _f64(intr0).reshape(-1)
        pz = _f64(poses0).reshape(-1)
        hs = (C.c_void_p * len(replicas))(*[r.h.value for r in replicas])
        ms = np.zeros(1)
        n, ex = C.c_int64(0), C.c_int64(0)
        check(lib.ccrs_bench_lm_steps_rotating(hs, len(replicas), _dp(a), _dp(pz), int(warmup), int(steps), _dp(ms), C.byref(n), C.byref(ex)))
        return float(ms[0]), int(n.value), int(ex.value)

    def launch_count(self) -> int:
        return int(self.lib.ccrs_launch_count(self.h))


def comm_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    check(_abi.load().ccrs_comm_unique_id(buf))
    return buf.raw


def measure_fp64_peak(device: int = 0) -> float:
    t = C.c_double(0.0)
    check(_abi.load().ccrs_measure_fp64_peak(int(device), C.byref(t)))
    return t.value


def model_bounds(model, width, height):
    m = MODELS[model] if isinstance(model, str) else int(model)
    n = _abi.load().ccrs_model_nparams(m)
    lo, hi = np.empty(n), np.empty(n)
    check(_abi.load().ccrs_model_bounds(m, int(width), int(height), _dp(lo), _dp(hi)))
    return lo, hi


def pack_frames(frame_feature_list: Sequence[Optional[FrameFeature]], initial_poses: Dict[int, RvecTvec]):
    """FrameFeature list -> SoA arrays. Frames that are None, have no initial pose, or fewer than 10 points
    are skipped like util.rs:401,431-433. f32 -> f64 widening as factors.rs:141-143."""
    valid, offs, xs, ys, zs, us, vs, poses = [], [0], [], [], [], [], [], []
    for i, ff in enumerate(frame_feature_list):
        if ff is None or i not in initial_poses or len(ff.features) < MIN_POINTS_PER_FRAME:
            continue
        for fp in ff.features.values():
            p3 = np.asarray(fp.p3d, dtype=np.float32).astype(np.float64)
            p2 = np.asarray(fp.p2d, dtype=np.float32).astype(np.float64)
            xs.append(p3[0]); ys.append(p3[1]); zs.append(p3[2]); us.append(p2[0]); vs.append(p2[1])
        offs.append(len(xs))
        valid.append(i)
        poses.append(initial_poses[i].as_array())
    return (valid, np.asarray(offs, dtype=np.int32), _f64(xs), _f64(ys), _f64(zs), _f64(us), _f64(vs),
            _f64(poses).reshape(-1, 6))


def calib_camera(frame_feature_list: Sequence[Optional[FrameFeature]], generic_camera: GenericModel,
                 xy_same_focal: bool, disabled_distortions: int, fixed_focal: bool,
                 initial_poses: Optional[Dict[int, RvecTvec]] = None, use_lm: bool = False,
                 options: Optional[Options] = None,
                 device: int = 0) -> Optional[Tuple[GenericModel, Dict[int, RvecTvec]]]:
    """Mirror of calib_camera (src/util.rs:384-490). With the reference's own five arguments the initial pose of every
    frame comes from the pose-initialisation step (unproject + PnP per frame, util.rs:418-439; here one launch for all
    frames); `initial_poses` overrides it. Returns None where the reference returns None (optimiser failure: NaN error
    or Cholesky failure)."""
    lib = _abi.load()
    if initial_poses is None:
        initial_poses = globals()["initial_poses"](frame_feature_list, generic_camera, device=device)
    valid, offs, x, y, z, u, v, poses = pack_frames(frame_feature_list, initial_poses)
    if not valid:
        return None
    params = _f64(generic_camera.params).copy()
    poses = poses.copy()
    s = Summary()
    o = options or default_options()
    code = lib.ccrs_calib_camera(generic_camera.model_id(), int(generic_camera.width), int(generic_camera.height),
                                 len(valid), offs.ctypes.data_as(C.POINTER(C.c_int32)), _dp(x), _dp(y), _dp(z), _dp(u), _dp(v),
                                 _dp(params), _dp(poses.reshape(-1)), int(xy_same_focal), int(disabled_distortions),
                                 int(fixed_focal), int(use_lm), C.byref(o), C.byref(s), int(device))
    if code in (-4, -5):
        return None
    check(code)
    cam = GenericModel(generic_camera.model, params, generic_camera.width, generic_camera.height)
    rt = {i: RvecTvec(tuple(poses[k, :3]), tuple(poses[k, 3:])) for k, i in enumerate(valid)}
    return cam, rt


def init_poses(frame_offsets, x, y, z, xn, yn, device: int = 0, want_cost: bool = False):
    """All frames' initial board poses in one launch (ccrs_init_poses). Returns poses [F, 6] (rvec, tvec)."""
    fo = np.ascontiguousarray(frame_offsets, dtype=np.int32)
    xs, ys, zs, a, b = map(_f64, (x, y, z, xn, yn))
    poses = np.empty((len(fo) - 1, 6))
    cost = np.empty(len(fo) - 1)
    check(_abi.load().ccrs_init_poses(len(fo) - 1, fo.ctypes.data_as(C.POINTER(C.c_int32)), _dp(xs), _dp(ys), _dp(zs), _dp(a),
                                      _dp(b), _dp(poses.reshape(-1)), _dp(cost), int(device)))
    return (poses, cost) if want_cost else poses


def initial_poses(frame_feature_list: Sequence[Optional[FrameFeature]], generic_camera: GenericModel,
                  device: int = 0) -> Dict[int, RvecTvec]:
    """The pose-initialisation part of calib_camera (src/util.rs:401-441) for every frame at once: unproject the
    detections with the current model (host, models.unproject), drop points outside its domain and frames left with
    fewer than 10 points (util.rs:431-433), normalise to z = 1 rounded to f32 like `glam::Vec2::new(.. as f32, ..)`
    (util.rs:425), and solve all frames in one launch. Returns {frame_idx: RvecTvec} for the valid frames."""
    from .models import unproject
    idx, offs, xs, ys, zs, xn, yn = [], [0], [], [], [], [], []
    for i, ff in enumerate(frame_feature_list):
        if ff is None:
            continue
        p2 = np.array([fp.p2d for fp in ff.features.values()], dtype=np.float32).astype(np.float64).reshape(-1, 2)
        p3 = np.array([fp.p3d for fp in ff.features.values()], dtype=np.float32).astype(np.float64).reshape(-1, 3)
        if len(p2) == 0:
            continue
        rays, valid = unproject(generic_camera.model, generic_camera.params, p2)
        valid &= np.abs(rays[:, 2]) > 1e-12
        if valid.sum() < MIN_POINTS_PER_FRAME:
            continue
        r = rays[valid]
        xs.append(p3[valid, 0]); ys.append(p3[valid, 1]); zs.append(p3[valid, 2])
        xn.append((r[:, 0] / r[:, 2]).astype(np.float32).astype(np.float64))
        yn.append((r[:, 1] / r[:, 2]).astype(np.float32).astype(np.float64))
        offs.append(offs[-1] + int(valid.sum()))
        idx.append(i)
    if not idx:
        return {}
    cat = np.concatenate
    poses = init_poses(offs, cat(xs), cat(ys), cat(zs), cat(xn), cat(yn), device=device)
    return {i: RvecTvec(tuple(poses[k, :3]), tuple(poses[k, 3:])) for k, i in enumerate(idx)}


def init_ucm(frame_feature0: FrameFeature, frame_feature1: FrameFeature, rtvec0: RvecTvec, rtvec1: RvecTvec,
             init_f: float, init_alpha: float, fixed_focal: bool, options: Optional[Options] = None,
             device: int = 0) -> Optional[GenericModel]:
    """Mirror of init_ucm (src/util.rs:284-378): [f, alpha] fit on two frames, then — from fresh PnP poses under that
    model, like the reference — the one-focal UCM calibration of those two frames. Returns None where the reference
    returns None (optimiser failure in the first stage)."""
    frames = [frame_feature0, frame_feature1]
    _, offs, x, y, z, u, v, poses = pack_frames(frames, {0: rtvec0, 1: rtvec1})
    if len(offs) != 3:
        return None
    w, h = frame_feature0.img_w_h
    out = np.empty(5)
    s = Summary()
    o = options or default_options()
    poses = poses.copy()
    code = _abi.load().ccrs_init_ucm(int(w), int(h), 2, offs.ctypes.data_as(C.POINTER(C.c_int32)), _dp(x), _dp(y), _dp(z),
                                     _dp(u), _dp(v), float(init_f), float(init_alpha), int(fixed_focal),
                                     _dp(poses.reshape(-1)), _dp(out), C.byref(o), C.byref(s), int(device))
    if code in (-4, -5):
        return None
    check(code)
    return GenericModel("ucm", out, int(w), int(h))


def convert_model(source_model: GenericModel, target_model: GenericModel, disabled_distortions: int,
                  options: Optional[Options] = None, device: int = 0) -> None:
    """Mirror of convert_model (src/util.rs:225-278): target_model.params is updated in place (`&mut GenericModel`)."""
    from .models import conversion_grid, unproject
    if int(round(source_model.width)) != int(round(target_model.width)):
        raise ValueError("source width and target width are not the same.")      # factors.rs:28-29 (panic)
    if int(round(source_model.height)) != int(round(target_model.height)):
        raise ValueError("source height and target height are not the same.")    # factors.rs:30-31
    rays, valid = unproject(source_model.model, source_model.params, conversion_grid(source_model.width, source_model.height))
    p = np.ascontiguousarray(rays[valid].T)                                        # filter_map(Some) factors.rs:39-42
    tgt = _f64(target_model.params).copy()
    src = _f64(source_model.params)
    s = Summary()
    o = options or default_options()
    check(_abi.load().ccrs_convert_model(source_model.model_id(), _dp(src), target_model.model_id(), _dp(tgt),
                                         int(source_model.width), int(source_model.height), int(disabled_distortions),
                                         p.shape[1], _dp(p[0]), _dp(p[1]), _dp(p[2]), C.byref(o), C.byref(s), int(device)))
    target_model.params = tgt


def validation(cam_idx: int, final_result: GenericModel, rtvec_list: Dict[int, RvecTvec],
               detected_feature_frames: Sequence[Optional[FrameFeature]], recording_option=None,
               device: int = 0) -> Tuple[float, float]:
    """Mirror of util::validation (src/util.rs:721-795): (median reprojection error, mean of the best 99 %) in px over
    every feature of every frame that has a pose. cam_idx / recording_option only label the reference's rerun logging
    (out of scope here) and are accepted for signature parity."""
    offs, xs, ys, zs, us, vs, poses = [0], [], [], [], [], [], []
    for i, rt in rtvec_list.items():
        ff = detected_feature_frames[i]
        if ff is None:                                   # util.rs:730
            continue
        for fp in ff.features.values():
            xs.append(fp.p3d[0]); ys.append(fp.p3d[1]); zs.append(fp.p3d[2]); us.append(fp.p2d[0]); vs.append(fp.p2d[1])
        offs.append(len(xs))
        poses.append(rt.as_array())
    f32 = lambda a: np.asarray(a, dtype=np.float32)      # FeaturePoint stores f32 (detected_points.rs:6-9)
    with Problem(final_result.model, final_result.width, final_result.height, offs, f32(xs), f32(ys), f32(zs), f32(us),
                 f32(vs), huber_delta=0.0, device=device) as gp:
        return gp.validation(final_result.params, _f64(poses))


class JointProblem:
    """Device-resident joint multi-camera problem (calib_all_camera_with_extrinsics, src/util.rs:567-715)."""

    def __init__(self, model, n_cams: int, n_frames: int, block_cam, block_frame, block_offsets, x, y, z, u, v,
                 xy_same_focal: bool = False, huber_delta: float = 1.0, device: int = 0):
        self.lib = _abi.load()
        self.model = MODELS[model] if isinstance(model, str) else int(model)
        bc = np.ascontiguousarray(block_cam, dtype=np.int32)
        bf = np.ascontiguousarray(block_frame, dtype=np.int32)
        bo = np.ascontiguousarray(block_offsets, dtype=np.int32)
        xs, ys, zs, us, vs = map(_f64, (x, y, z, u, v))
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        self.h = C.c_void_p()
        code = self.lib.ccrs_joint_create(C.byref(self.h), self.model, int(xy_same_focal), int(n_cams), int(n_frames), len(bc),
                                          ip(bc), ip(bf), ip(bo), _dp(xs), _dp(ys), _dp(zs), _dp(us), _dp(vs),
                                          float(huber_delta), int(device))
        self._check(code)
        self.n_cams, self.n_frames, self.n_obs = int(n_cams), int(n_frames), int(bo[-1])
        self.d = self.lib.ccrs_joint_dim(self.h)

    def _check(self, code):
        if code != 0:
            raise CcrsError(code, self.lib.ccrs_joint_last_error().decode("utf-8", "replace"))

    @classmethod
    def from_rig(cls, rig, **kw):
        return cls(rig.model, rig.n_cams, rig.n_frames, rig.block_cam, rig.block_frame, rig.block_offsets,
                   rig.x, rig.y, rig.z, rig.u, rig.v, **kw)

    def close(self):
        if getattr(self, "h", None) is not None and self.h:
            self.lib.ccrs_joint_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            if sys is None or sys.is_finalizing():
                return
            self.close()
        except Exception:
            pass

    def eval_rj(self, intr, extr, poses, apply_loss=True):
        intr, extr, poses = _f64(intr).reshape(-1), _f64(extr).reshape(-1), _f64(poses).reshape(-1)
        r = np.empty(2 * self.n_obs); J = np.empty((2 * self.n_obs, self.d + 12))
        self._check(self.lib.ccrs_joint_eval_rj(self.h, _dp(intr), _dp(extr), _dp(poses), int(apply_loss), _dp(r), _dp(J)))
        return r, J

    def solve_gn(self, intr, extr, poses, lo=None, hi=None, fixed=None, options: Optional[Options] = None):
        a = _f64(intr).reshape(-1).copy(); e = _f64(extr).reshape(-1).copy(); p = _f64(poses).reshape(-1).copy()
        o = options or default_options()
        s = Summary(); hist = np.full(max(o.max_iteration, 1), np.nan)
        lo_a = _f64(lo).reshape(-1) if lo is not None else None
        hi_a = _f64(hi).reshape(-1) if hi is not None else None
        fx = np.ascontiguousarray(fixed, dtype=np.uint8).reshape(-1) if fixed is not None else None
        code = self.lib.ccrs_joint_solve_gn(self.h, _dp(a), _dp(e), _dp(p), _dp(lo_a), _dp(hi_a), _up(fx), C.byref(o), C.byref(s), _dp(hist))
        if code not in (0, -4, -5):
            self._check(code)
        return a.reshape(self.n_cams, self.d), e.reshape(-1, 6), p.reshape(-1, 6), s, hist[: s.iterations]


def calib_all_camera_with_extrinsics(cameras: Sequence[GenericModel], t_cam_i_0: Sequence[RvecTvec],
                                     cam_rtvecs: Sequence[Dict[int, RvecTvec]],
                                     cams_detected_feature_frames: Sequence[Sequence[Optional[FrameFeature]]],
                                     xy_same_focal: bool, disabled_distortions: int, cam0_fixed_focal: bool,
                                     options: Optional[Options] = None, device: int = 0):
    """Mirror of calib_all_camera_with_extrinsics (src/util.rs:567-715).
    Returns (intrinsics, t_i_0 list, {frame_idx: board RvecTvec}) or None where the reference returns None."""
    n_cams = len(cameras)
    model = cameras[0].model
    nfull = len(cameras[0].params)
    if any(c.model != model or len(c.params) != nfull for c in cameras):
        raise ValueError("calib_all_camera_with_extrinsics: every camera must use the same model")
    shift = 1 if xy_same_focal else 0
    d = nfull - shift
    frame_ids = sorted({f for c in range(n_cams) for f in cam_rtvecs[c]})      # valid_frame_board_to_cam0
    fidx = {f: i for i, f in enumerate(frame_ids)}
    bc, bf, offs, xs, ys, zs, us, vs = [], [], [0], [], [], [], [], []
    poses = np.zeros((len(frame_ids), 6)); have_pose = np.zeros(len(frame_ids), dtype=bool)

    def iso(rt):
        from .synth import rodrigues
        return rodrigues(np.asarray(rt.rvec)), np.asarray(rt.tvec, dtype=np.float64)

    for c in range(n_cams):
        for f, rt in cam_rtvecs[c].items():
            ff = cams_detected_feature_frames[c][f]
            bc.append(c); bf.append(fidx[f])
            for fp in ff.features.values():
                p3 = np.asarray(fp.p3d, dtype=np.float32).astype(np.float64); p2 = np.asarray(fp.p2d, dtype=np.float32).astype(np.float64)
                xs.append(p3[0]); ys.append(p3[1]); zs.append(p3[2]); us.append(p2[0]); vs.append(p2[1])
            offs.append(len(xs))
            if not have_pose[fidx[f]]:                     # entry().or_insert: first camera that saw the frame wins
                if c == 0:
                    poses[fidx[f]] = rt.as_array()
                else:                                      # 0 <- i <- board (util.rs:640-650): T_c_0^-1 * T_c_b
                    from .synth import rotmat_to_rvec
                    Rc, tc = iso(t_cam_i_0[c]); Rb, tb = iso(rt)
                    R = Rc.T @ Rb; t = Rc.T @ (tb - tc)
                    poses[fidx[f]] = np.concatenate([rotmat_to_rvec(R), t])
                have_pose[fidx[f]] = True
    intr = np.zeros((n_cams, d)); lo = np.zeros((n_cams, d)); hi = np.zeros((n_cams, d)); fixed = np.zeros((n_cams, d), dtype=np.uint8)
    keep = [i for i in range(nfull) if not (xy_same_focal and i == 1)]
    for c, cam in enumerate(cameras):
        flo, fhi = model_bounds(cam.model, cam.width, cam.height)
        intr[c] = _f64(cam.params)[keep]; lo[c] = flo[keep]; hi[c] = fhi[keep]
        for i in range(disabled_distortions):              # set_problem_parameter_disabled (util.rs:50-71)
            idx = nfull - 1 - shift - i
            fixed[c, idx] = 1; intr[c, idx] = 0.0
    if cam0_fixed_focal:
        fixed[0, 0] = 1                                    # problem.fix_variable("params0", 0) (util.rs:664-667)
    extr = np.stack([t.as_array() for t in t_cam_i_0])
    jp = JointProblem(model, n_cams, len(frame_ids), bc, bf, offs, xs, ys, zs, us, vs, xy_same_focal=xy_same_focal, device=device)
    try:
        a, e, p, s, _ = jp.solve_gn(intr, extr, poses, lo, hi, fixed, options)
    finally:
        jp.close()
    if s.status in (-4, -5):
        return None
    out_cams = []
    for c, cam in enumerate(cameras):
        full = np.insert(a[c], 1, a[c][0]) if xy_same_focal else a[c].copy()
        out_cams.append(GenericModel(cam.model, full, cam.width, cam.height))
    t_i_0 = [RvecTvec((0.0, 0.0, 0.0), (0.0, 0.0, 0.0))] + [RvecTvec(tuple(e[c, :3]), tuple(e[c, 3:])) for c in range(1, n_cams)]
    board = {f: RvecTvec(tuple(p[i, :3]), tuple(p[i, 3:])) for f, i in fidx.items()}
    return out_cams, t_i_0, board

"""ctypes binding of the CPU oracle (oracle/libccrs_oracle.so). TEST INFRASTRUCTURE ONLY:
importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
The product package never imports this module."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libccrs_oracle.so")


def build(force: bool = False) -> str:
    src = [os.path.join(_HERE, f) for f in ("ccrs_oracle.cpp", "ccrs_oracle.hpp")]
    stale = (not os.path.exists(_LIB_PATH)) or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in src)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-s"])
    return _LIB_PATH


class Problem(C.Structure):
    _fields_ = [("model", C.c_int), ("width", C.c_int), ("height", C.c_int), ("xy_same_focal", C.c_int),
                ("n_frames", C.c_int), ("frame_offsets", C.POINTER(C.c_int32)),
                ("x", C.POINTER(C.c_double)), ("y", C.POINTER(C.c_double)), ("z", C.POINTER(C.c_double)),
                ("u", C.POINTER(C.c_double)), ("v", C.POINTER(C.c_double)),
                ("huber_delta", C.c_double), ("n_threads", C.c_int)]


class Options(C.Structure):
    _fields_ = [("max_iteration", C.c_int), ("min_abs_decrease", C.c_double), ("min_rel_decrease", C.c_double),
                ("min_error", C.c_double), ("lm_initial_radius", C.c_double), ("lm_min_diag", C.c_double),
                ("lm_max_diag", C.c_double), ("fixed_mode", C.c_int), ("solver", C.c_int)]


class Result(C.Structure):
    _fields_ = [("iterations", C.c_int), ("status", C.c_int), ("stop_reason", C.c_int),
                ("final_error", C.c_double), ("n_accepted", C.c_int), ("n_rejected", C.c_int)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB_PATH)
        _lib.oracle_linearize.restype = C.c_double
        _lib.oracle_sq_error.restype = C.c_double
    return _lib


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double)) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


class OracleProblem:
    """Holds numpy arrays alive and exposes the oracle entry points."""

    def __init__(self, model: int, width: int, height: int, frame_offsets, x, y, z, u, v,
                 xy_same_focal: bool = False, huber_delta: float = 1.0, n_threads: int = 0):
        self.frame_offsets = np.ascontiguousarray(frame_offsets, dtype=np.int32)
        self.x, self.y, self.z, self.u, self.v = map(_f64, (x, y, z, u, v))
        self.model = int(model)
        self.xy_same_focal = bool(xy_same_focal)
        self.n_frames = len(self.frame_offsets) - 1
        self.n_obs = int(self.frame_offsets[-1])
        self.d = lib().oracle_model_nparams(self.model) - (1 if xy_same_focal else 0)
        self.c = Problem(self.model, width, height, int(xy_same_focal), self.n_frames,
                         self.frame_offsets.ctypes.data_as(C.POINTER(C.c_int32)),
                         _dp(self.x), _dp(self.y), _dp(self.z), _dp(self.u), _dp(self.v),
                         float(huber_delta), int(n_threads))

    @classmethod
    def from_synth(cls, s, model_id: int, **kw):
        return cls(model_id, s.width, s.height, s.frame_offsets, s.x, s.y, s.z, s.u, s.v, **kw)

    def eval_rj(self, intr, poses, apply_loss=True, want_j=True):
        intr, poses = _f64(intr), _f64(poses)
        r = np.empty(2 * self.n_obs)
        J = np.empty((2 * self.n_obs, self.d + 6)) if want_j else None
        lib().oracle_eval_rj(C.byref(self.c), _dp(intr), _dp(poses), int(apply_loss), _dp(r), _dp(J))
        return r, J

    def eval_r(self, intr, poses, apply_loss=True):
        intr, poses = _f64(intr), _f64(poses)
        r = np.empty(2 * self.n_obs)
        lib().oracle_eval_r(C.byref(self.c), _dp(intr), _dp(poses), int(apply_loss), _dp(r))
        return r

    def validation(self, intr, poses):
        """util::validation (src/util.rs:721-795): (median, mean of the best 99 %, per-point errors)."""
        intr, poses = _f64(intr), _f64(poses)
        med, avg = C.c_double(0.0), C.c_double(0.0)
        e = np.empty(self.n_obs)
        lib().oracle_validation(C.byref(self.c), _dp(intr), _dp(poses), C.byref(med), C.byref(avg), _dp(e))
        return med.value, avg.value, e

    def othercam_rj(self, intr, poses_0_b, pose_i_0, apply_loss=True):
        intr, poses_0_b, pose_i_0 = _f64(intr), _f64(poses_0_b), _f64(pose_i_0)
        r = np.empty(2 * self.n_obs)
        J = np.empty((2 * self.n_obs, self.d + 12))
        lib().oracle_othercam_rj(C.byref(self.c), _dp(intr), _dp(poses_0_b), _dp(pose_i_0), int(apply_loss), _dp(r), _dp(J))
        return r, J

    def nblk(self):
        return lib().oracle_nblk(C.byref(self.c))

    def linearize(self, intr, poses):
        intr, poses = _f64(intr), _f64(poses)
        blk = np.empty((self.n_frames, self.nblk()))
        sq = lib().oracle_linearize(C.byref(self.c), _dp(intr), _dp(poses), _dp(blk))
        return sq, blk

    def sq_error(self, intr, poses):
        intr, poses = _f64(intr), _f64(poses)
        return lib().oracle_sq_error(C.byref(self.c), _dp(intr), _dp(poses))

    def default_options(self, **kw) -> Options:
        o = Options()
        lib().oracle_default_options(C.byref(o))
        for k, v in kw.items():
            setattr(o, k, v)
        return o

    def solve_step(self, intr, poses, u=0.0, scale=None, fixed=None, options=None):
        intr, poses = _f64(intr), _f64(poses)
        o = options or self.default_options()
        di = np.empty(self.d); dp = np.empty(6 * self.n_frames)
        md = C.c_double(0.0)
        sc = _f64(scale) if scale is not None else None
        fx = np.ascontiguousarray(fixed, dtype=np.uint8) if fixed is not None else None
        st = lib().oracle_solve_step(C.byref(self.c), _dp(intr), _dp(poses), C.c_double(u), _dp(sc), C.byref(o),
                                     fx.ctypes.data_as(C.POINTER(C.c_ubyte)) if fx is not None else None,
                                     _dp(di), _dp(dp), C.byref(md))
        return st, di, dp.reshape(-1, 6), md.value

    def _run(self, fn, intr, poses, lo, hi, fixed, options):
        intr = _f64(intr).copy(); poses = _f64(poses).copy()
        o = options or self.default_options()
        res = Result()
        hist = np.full(o.max_iteration, np.nan)
        lo_a = _f64(lo) if lo is not None else None
        hi_a = _f64(hi) if hi is not None else None
        fx = np.ascontiguousarray(fixed, dtype=np.uint8) if fixed is not None else None
        fn(C.byref(self.c), _dp(intr), _dp(poses), _dp(lo_a), _dp(hi_a),
           fx.ctypes.data_as(C.POINTER(C.c_ubyte)) if fx is not None else None, C.byref(o), C.byref(res), _dp(hist))
        return intr, poses.reshape(-1, 6), res, hist[: res.iterations]

    def gauss_newton(self, intr, poses, lo=None, hi=None, fixed=None, options=None):
        return self._run(lib().oracle_gn, intr, poses, lo, hi, fixed, options)

    def levenberg_marquardt(self, intr, poses, lo=None, hi=None, fixed=None, options=None):
        return self._run(lib().oracle_lm, intr, poses, lo, hi, fixed, options)


def init_ucm_gn(op: "OracleProblem", cx: float, cy: float, f: float, alpha: float, poses, fixed_focal=False, options=None):
    """init_ucm first stage (util.rs:295-357) on the frames of `op` (a UCM problem; only its observations are used)."""
    fa = np.array([f, alpha], dtype=np.float64)
    poses = _f64(poses).copy()
    o = options or op.default_options()
    res = Result()
    hist = np.full(o.max_iteration, np.nan)
    lib().oracle_init_ucm_gn(C.byref(op.c), C.c_double(cx), C.c_double(cy), _dp(fa), _dp(poses), int(fixed_focal),
                             C.byref(o), C.byref(res), _dp(hist))
    return fa, poses.reshape(-1, 6), res, hist[: res.iterations]


def convert_model_gn(src_model: int, src_params, tgt_model: int, tgt_params, p3d, lo=None, hi=None, fixed=None,
                     huber_delta: float = 1.0, options=None):
    """convert_model's optimisation (util.rs:244-277) on given unprojected points p3d [n,3]."""
    src = _f64(src_params); tgt = _f64(tgt_params).copy(); p3 = _f64(p3d).reshape(-1, 3)
    if options is None:
        options = Options(); lib().oracle_default_options(C.byref(options))
    res = Result(); hist = np.full(options.max_iteration, np.nan)
    lo_a = _f64(lo) if lo is not None else None
    hi_a = _f64(hi) if hi is not None else None
    fx = np.ascontiguousarray(fixed, dtype=np.uint8) if fixed is not None else None
    lib().oracle_convert_model_gn(int(src_model), _dp(src), int(tgt_model), _dp(tgt), len(p3), _dp(p3), _dp(lo_a), _dp(hi_a),
                                  fx.ctypes.data_as(C.POINTER(C.c_ubyte)) if fx is not None else None,
                                  C.c_double(huber_delta), C.byref(options), C.byref(res), _dp(hist))
    return tgt, res, hist[: res.iterations]


def project(model: int, params, P):
    params, P = _f64(params), _f64(P)
    uv = np.empty(2)
    lib().oracle_project(int(model), _dp(params), _dp(P), _dp(uv))
    return uv


def transform_point(rvec, tvec, p):
    rvec, tvec, p = _f64(rvec), _f64(tvec), _f64(p)
    out = np.empty(3)
    lib().oracle_transform_point(_dp(rvec), _dp(tvec), _dp(p), _dp(out))
    return out


class JointStruct(C.Structure):
    _fields_ = [("model", C.c_int), ("xy_same_focal", C.c_int), ("n_cams", C.c_int), ("n_frames", C.c_int),
                ("n_blocks", C.c_int), ("block_cam", C.POINTER(C.c_int32)), ("block_frame", C.POINTER(C.c_int32)),
                ("block_offsets", C.POINTER(C.c_int32)),
                ("x", C.POINTER(C.c_double)), ("y", C.POINTER(C.c_double)), ("z", C.POINTER(C.c_double)),
                ("u", C.POINTER(C.c_double)), ("v", C.POINTER(C.c_double)), ("huber_delta", C.c_double), ("n_threads", C.c_int)]


class OracleJoint:
    """calib_all_camera_with_extrinsics (src/util.rs:567-715) restated on the CPU: dense normal equations."""

    def __init__(self, rig, model_id: int, xy_same_focal: bool = False, huber_delta: float = 1.0):
        ip = lambda a: a.ctypes.data_as(C.POINTER(C.c_int32))
        self.bc = np.ascontiguousarray(rig.block_cam, dtype=np.int32)
        self.bf = np.ascontiguousarray(rig.block_frame, dtype=np.int32)
        self.bo = np.ascontiguousarray(rig.block_offsets, dtype=np.int32)
        self.x, self.y, self.z, self.u, self.v = map(_f64, (rig.x, rig.y, rig.z, rig.u, rig.v))
        self.n_cams, self.n_frames, self.n_obs = rig.n_cams, rig.n_frames, int(self.bo[-1])
        self.d = lib().oracle_model_nparams(model_id) - (1 if xy_same_focal else 0)
        self.c = JointStruct(model_id, int(xy_same_focal), rig.n_cams, rig.n_frames, len(self.bc), ip(self.bc), ip(self.bf),
                             ip(self.bo), _dp(self.x), _dp(self.y), _dp(self.z), _dp(self.u), _dp(self.v), float(huber_delta), 0)

    def eval_rj(self, intr, extr, poses, apply_loss=True):
        intr, extr, poses = _f64(intr), _f64(extr), _f64(poses)
        r = np.empty(2 * self.n_obs); J = np.empty((2 * self.n_obs, self.d + 12))
        lib().oracle_joint_eval_rj(C.byref(self.c), _dp(intr), _dp(extr), _dp(poses), int(apply_loss), _dp(r), _dp(J))
        return r, J

    def gauss_newton(self, intr, extr, poses, lo=None, hi=None, fixed=None, options=None):
        intr = _f64(intr).copy(); extr = _f64(extr).copy(); poses = _f64(poses).copy()
        if options is None:
            options = Options(); lib().oracle_default_options(C.byref(options))
        res = Result(); hist = np.full(options.max_iteration, np.nan)
        lo_a = _f64(lo) if lo is not None else None
        hi_a = _f64(hi) if hi is not None else None
        fx = np.ascontiguousarray(fixed, dtype=np.uint8) if fixed is not None else None
        lib().oracle_joint_gn(C.byref(self.c), _dp(intr), _dp(extr), _dp(poses), _dp(lo_a), _dp(hi_a),
                              fx.ctypes.data_as(C.POINTER(C.c_ubyte)) if fx is not None else None, C.byref(options),
                              C.byref(res), _dp(hist))
        return intr.reshape(self.n_cams, -1), extr.reshape(-1, 6), poses.reshape(-1, 6), res, hist[: res.iterations]

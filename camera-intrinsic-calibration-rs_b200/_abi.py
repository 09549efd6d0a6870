"""ctypes prototypes for include/ccrs_b200.h (libccrs_b200.so, built in-tree by csrc/Makefile).

The library is the product; this file only declares its C ABI. There is no Python/CPU fallback:
if the shared object is missing the import fails loudly, and without an sm_100 device
ccrs_problem_create returns CCRS_ERR_NO_DEVICE.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# CCRS_B200_LIB selects another build of the same library (e.g. the `make timing` debug build); never a fallback
LIB_PATH = os.environ.get("CCRS_B200_LIB") or os.path.join(_HERE, "libccrs_b200.so")

# every symbol include/ccrs_b200.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "ccrs_model_nparams", "ccrs_last_error", "ccrs_problem_create", "ccrs_problem_create_f32", "ccrs_problem_create_board_f32", "ccrs_batch_create", "ccrs_problem_destroy", "ccrs_release_cached_memory",
    "ccrs_problem_dim", "ccrs_problem_nblk", "ccrs_problem_n_frames", "ccrs_problem_n_obs", "ccrs_problem_n_problems",
    "ccrs_set_poses", "ccrs_get_poses", "ccrs_eval_rj", "ccrs_validation", "ccrs_linearize", "ccrs_get_frame_blocks",
    "ccrs_compute_scale", "ccrs_set_intr_scale", "ccrs_reduce", "ccrs_backsub", "ccrs_eval_cost", "ccrs_accept",
    "ccrs_comm_unique_id", "ccrs_comm_init", "ccrs_comm_finalize", "ccrs_comm_set_deterministic", "ccrs_comm_uses_peer_memory", "ccrs_default_options",
    "ccrs_solve_gn", "ccrs_solve_lm", "ccrs_controller_gn", "ccrs_controller_lm", "ccrs_calib_camera",
    "ccrs_joint_create", "ccrs_joint_destroy", "ccrs_joint_dim", "ccrs_joint_last_error", "ccrs_joint_launch_count",
    "ccrs_joint_eval_rj", "ccrs_joint_solve_gn", "ccrs_init_poses", "ccrs_set_fixed_poses", "ccrs_init_ucm", "ccrs_convert_model", "ccrs_model_bounds", "ccrs_measure_fp64_peak", "ccrs_time_linearize", "ccrs_bench_lm_steps", "ccrs_bench_lm_steps_rotating", "ccrs_problem_update_observations", "ccrs_launch_count", "ccrs_step_trace", "ccrs_spec_k3_counters", "ccrs_loop_counters", "ccrs_loop_trace",
]

STATUS = {0: "CCRS_OK", -1: "CCRS_ERR_INVALID", -2: "CCRS_ERR_CUDA", -3: "CCRS_ERR_NO_DEVICE",
          -4: "CCRS_ERR_NUMERIC", -5: "CCRS_ERR_CHOLESKY", -6: "CCRS_ERR_COMM"}


class CcrsError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"{STATUS.get(code, code)}: {msg}")
        self.code = code


class Options(C.Structure):
    _fields_ = [("max_iteration", C.c_int), ("min_abs_decrease", C.c_double), ("min_rel_decrease", C.c_double),
                ("min_error", C.c_double), ("lm_initial_radius", C.c_double), ("lm_min_diag", C.c_double),
                ("lm_max_diag", C.c_double), ("fixed_mode", C.c_int), ("speculative", C.c_int), ("verbose", C.c_int),
                ("block_huber_delta", C.c_double)]


class Summary(C.Structure):
    _fields_ = [("iterations", C.c_int), ("status", C.c_int), ("stop_reason", C.c_int), ("final_error", C.c_double),
                ("n_accepted", C.c_int), ("n_rejected", C.c_int), ("device_ms", C.c_double)]


_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int32)
_up = C.POINTER(C.c_ubyte)

BE_LINEARIZE = C.CFUNCTYPE(C.c_int, C.c_void_p, _dp, C.c_int)
BE_COMPUTE_SCALE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, _dp)
BE_SET_INTR_SCALE = C.CFUNCTYPE(C.c_int, C.c_void_p, _dp)
BE_REDUCE = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_int, _dp, C.c_int, C.c_double, C.c_double, _dp)
BE_BACKSUB = C.CFUNCTYPE(C.c_int, C.c_void_p, _dp, _dp, _up, C.c_int)
BE_TRIAL_STATS = C.CFUNCTYPE(C.c_int, C.c_void_p, _dp, C.c_int, _dp)
BE_ACCEPT = C.CFUNCTYPE(C.c_int, C.c_void_p, _up)
BE_ALLREDUCE = C.CFUNCTYPE(C.c_int, C.c_void_p, _dp, C.c_int)


class Backend(C.Structure):
    _fields_ = [("ctx", C.c_void_p), ("d", C.c_int), ("n_problems", C.c_int),
                ("linearize", BE_LINEARIZE), ("compute_scale", BE_COMPUTE_SCALE), ("set_intr_scale", BE_SET_INTR_SCALE),
                ("reduce", BE_REDUCE), ("backsub", BE_BACKSUB), ("trial_stats", BE_TRIAL_STATS), ("accept", BE_ACCEPT),
                ("allreduce", BE_ALLREDUCE)]


_lib = None


def load():
    """dlopen libccrs_b200.so and attach prototypes. Raises if the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `make -C {os.path.join(_HERE, 'csrc')}` "
                          f"(or __graft_entry__.build()). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    lib.ccrs_last_error.restype = C.c_char_p
    lib.ccrs_problem_n_obs.restype = C.c_int64
    lib.ccrs_launch_count.restype = C.c_int64
    lib.ccrs_problem_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip,
                                        _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int]
    _fp = C.POINTER(C.c_float)
    lib.ccrs_problem_create_f32.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip,
                                            _fp, _fp, _fp, _fp, _fp, C.c_double, C.c_int]
    lib.ccrs_problem_create_board_f32.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip,
                                                  _fp, _fp, _fp, C.c_int, C.c_double, C.c_int]
    lib.ccrs_batch_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, C.c_int, _ip,
                                      _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int]
    lib.ccrs_problem_destroy.argtypes = [vp]
    for name in ("ccrs_problem_dim", "ccrs_problem_nblk", "ccrs_problem_n_frames", "ccrs_problem_n_obs",
                 "ccrs_problem_n_problems", "ccrs_launch_count"):
        getattr(lib, name).argtypes = [vp]
    lib.ccrs_set_poses.argtypes = [vp, _dp]
    lib.ccrs_get_poses.argtypes = [vp, _dp]
    lib.ccrs_eval_rj.argtypes = [vp, _dp, _dp, C.c_int, _dp, _dp]
    lib.ccrs_validation.argtypes = [vp, _dp, _dp, _dp, _dp, _dp]
    lib.ccrs_linearize.argtypes = [vp, _dp, C.c_int, _dp]
    lib.ccrs_get_frame_blocks.argtypes = [vp, C.c_int, _dp]
    lib.ccrs_compute_scale.argtypes = [vp, C.c_int, _dp]
    lib.ccrs_set_intr_scale.argtypes = [vp, _dp]
    lib.ccrs_reduce.argtypes = [vp, C.c_int, _dp, C.c_int, C.c_double, C.c_double, _dp]
    lib.ccrs_backsub.argtypes = [vp, _dp, _dp, C.c_int, _dp]
    lib.ccrs_eval_cost.argtypes = [vp, _dp, C.c_int, _dp]
    lib.ccrs_accept.argtypes = [vp, _up]
    lib.ccrs_comm_unique_id.argtypes = [C.c_void_p]
    lib.ccrs_comm_init.argtypes = [vp, C.c_void_p, C.c_int, C.c_int]
    lib.ccrs_comm_set_deterministic.argtypes = [vp, C.c_int]
    lib.ccrs_default_options.argtypes = [C.POINTER(Options)]
    lib.ccrs_default_options.restype = None
    for name in ("ccrs_solve_gn", "ccrs_solve_lm"):
        getattr(lib, name).argtypes = [vp, _dp, _dp, _dp, _up, C.POINTER(Options), C.POINTER(Summary), _dp]
    for name in ("ccrs_controller_gn", "ccrs_controller_lm"):
        getattr(lib, name).argtypes = [C.POINTER(Backend), _dp, _dp, _dp, _up, C.POINTER(Options), C.POINTER(Summary), _dp]
    lib.ccrs_calib_camera.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp,
                                      C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(Options), C.POINTER(Summary), C.c_int]
    lib.ccrs_model_bounds.argtypes = [C.c_int, C.c_int, C.c_int, _dp, _dp]
    lib.ccrs_init_poses.argtypes = [C.c_int, _ip, _dp, _dp, _dp, _dp, _dp, _dp, _dp, C.c_int]
    lib.ccrs_set_fixed_poses.argtypes = [vp, C.c_int]
    lib.ccrs_init_ucm.argtypes = [C.c_int, C.c_int, C.c_int, _ip, _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_double, C.c_int,
                                  _dp, _dp, C.POINTER(Options), C.POINTER(Summary), C.c_int]
    lib.ccrs_convert_model.argtypes = [C.c_int, _dp, C.c_int, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, _dp, _dp,
                                       C.POINTER(Options), C.POINTER(Summary), C.c_int]
    lib.ccrs_joint_create.argtypes = [C.POINTER(vp), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, _ip, _ip, _ip,
                                      _dp, _dp, _dp, _dp, _dp, C.c_double, C.c_int]
    lib.ccrs_joint_destroy.argtypes = [vp]
    lib.ccrs_joint_dim.argtypes = [vp]
    lib.ccrs_joint_last_error.restype = C.c_char_p
    lib.ccrs_joint_launch_count.argtypes = [vp]
    lib.ccrs_joint_launch_count.restype = C.c_int64
    lib.ccrs_joint_eval_rj.argtypes = [vp, _dp, _dp, _dp, C.c_int, _dp, _dp]
    lib.ccrs_joint_solve_gn.argtypes = [vp, _dp, _dp, _dp, _dp, _dp, _up, C.POINTER(Options), C.POINTER(Summary), _dp]
    lib.ccrs_spec_k3_counters.argtypes = [C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.ccrs_loop_counters.argtypes = [C.POINTER(C.c_int64)]
    lib.ccrs_loop_trace.argtypes = [C.c_int, _dp, C.POINTER(C.c_int64)]
    lib.ccrs_measure_fp64_peak.argtypes = [C.c_int, _dp]
    lib.ccrs_time_linearize.argtypes = [vp, _dp, C.c_int, C.c_int, _dp]
    lib.ccrs_bench_lm_steps.argtypes = [vp, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, _dp, C.POINTER(C.c_int64)]
    lib.ccrs_problem_update_observations.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_int32), vp, vp, vp, vp, vp]
    lib.ccrs_bench_lm_steps_rotating.argtypes = [C.POINTER(vp), C.c_int, _dp, _dp, C.c_int, C.c_int, _dp, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    _lib = lib
    return lib


def check(code: int):
    if code != 0:
        raise CcrsError(code, load().ccrs_last_error().decode("utf-8", "replace"))


def default_options(**kw) -> Options:
    o = Options()
    load().ccrs_default_options(C.byref(o))
    for k, v in kw.items():
        setattr(o, k, v)
    return o

//! Raw bindings to `include/ccrs_b200.h` (ABI version 2). SOURCE ONLY — never compiled in the build image.
#![allow(non_camel_case_types)]
use libc::{c_char, c_double, c_int, c_uchar, c_void};

#[repr(C)]
pub struct ccrs_problem {
    _private: [u8; 0],
}

pub const CCRS_UCM: c_int = 0;
pub const CCRS_EUCM: c_int = 1;
pub const CCRS_EUCMT: c_int = 2;
pub const CCRS_KB4: c_int = 3;
pub const CCRS_OPENCV5: c_int = 4;
pub const CCRS_FTHETA: c_int = 5;

pub const CCRS_OK: c_int = 0;
pub const CCRS_ERR_NUMERIC: c_int = -4; // tiny-solver: optimize() -> None (NaN error)
pub const CCRS_ERR_CHOLESKY: c_int = -5; // tiny-solver: optimize() -> None (LLT failure)

#[repr(C)]
#[derive(Clone, Copy, Debug)]
pub struct ccrs_options {
    pub max_iteration: c_int,
    pub min_abs_decrease: c_double,
    pub min_rel_decrease: c_double,
    pub min_error: c_double,
    pub lm_initial_radius: c_double,
    pub lm_min_diag: c_double,
    pub lm_max_diag: c_double,
    pub fixed_mode: c_int,
    pub speculative: c_int,
    pub verbose: c_int,
    pub block_huber_delta: c_double,
}

#[repr(C)]
#[derive(Clone, Copy, Debug, Default)]
pub struct ccrs_summary {
    pub iterations: c_int,
    pub status: c_int,
    pub stop_reason: c_int,
    pub final_error: c_double,
    pub n_accepted: c_int,
    pub n_rejected: c_int,
    pub device_ms: c_double,
}

extern "C" {
    pub fn ccrs_model_nparams(model: c_int) -> c_int;
    pub fn ccrs_last_error() -> *const c_char;
    pub fn ccrs_problem_create(out: *mut *mut ccrs_problem, model: c_int, width: c_int, height: c_int,
                               xy_same_focal: c_int, n_frames: c_int, frame_offsets: *const i32,
                               x: *const c_double, y: *const c_double, z: *const c_double,
                               u: *const c_double, v: *const c_double, huber_delta: c_double, device_id: c_int) -> c_int;
    pub fn ccrs_problem_destroy(p: *mut ccrs_problem) -> c_int;
    pub fn ccrs_problem_dim(p: *const ccrs_problem) -> c_int;
    pub fn ccrs_set_poses(p: *mut ccrs_problem, poses: *const c_double) -> c_int;
    pub fn ccrs_get_poses(p: *mut ccrs_problem, poses: *mut c_double) -> c_int;
    pub fn ccrs_eval_rj(p: *mut ccrs_problem, intr: *const c_double, poses: *const c_double, apply_loss: c_int,
                        r: *mut c_double, j: *mut c_double) -> c_int;
    pub fn ccrs_default_options(o: *mut ccrs_options);
    pub fn ccrs_solve_gn(p: *mut ccrs_problem, intr: *mut c_double, lo: *const c_double, hi: *const c_double,
                         fixed: *const c_uchar, opt: *const ccrs_options, summary: *mut ccrs_summary,
                         err_hist: *mut c_double) -> c_int;
    pub fn ccrs_solve_lm(p: *mut ccrs_problem, intr: *mut c_double, lo: *const c_double, hi: *const c_double,
                         fixed: *const c_uchar, opt: *const ccrs_options, summary: *mut ccrs_summary,
                         err_hist: *mut c_double) -> c_int;
    pub fn ccrs_calib_camera(model: c_int, width: c_int, height: c_int, n_frames: c_int, frame_offsets: *const i32,
                             x: *const c_double, y: *const c_double, z: *const c_double, u: *const c_double,
                             v: *const c_double, params: *mut c_double, poses: *mut c_double, xy_same_focal: c_int,
                             disabled_distortions: c_int, fixed_focal: c_int, use_lm: c_int, opt: *const ccrs_options,
                             summary: *mut ccrs_summary, device_id: c_int) -> c_int;
    pub fn ccrs_validation(p: *mut ccrs_problem, intr: *const c_double, poses: *const c_double, median: *mut c_double,
                           avg99: *mut c_double, errors: *mut c_double) -> c_int;
    pub fn ccrs_init_poses(n_frames: c_int, frame_offsets: *const i32, x: *const c_double, y: *const c_double,
                           z: *const c_double, xn: *const c_double, yn: *const c_double, poses_out: *mut c_double,
                           cost_out: *mut c_double, device_id: c_int) -> c_int;
    pub fn ccrs_set_fixed_poses(p: *mut ccrs_problem, fixed: c_int) -> c_int;
    pub fn ccrs_init_ucm(width: c_int, height: c_int, n_frames: c_int, frame_offsets: *const i32, x: *const c_double,
                         y: *const c_double, z: *const c_double, u: *const c_double, v: *const c_double,
                         init_f: c_double, init_alpha: c_double, fixed_focal: c_int, poses: *mut c_double,
                         params_out: *mut c_double, opt: *const ccrs_options, summary: *mut ccrs_summary,
                         device_id: c_int) -> c_int;
    pub fn ccrs_convert_model(src_model: c_int, src_params: *const c_double, tgt_model: c_int, tgt_params: *mut c_double,
                              width: c_int, height: c_int, disabled_distortions: c_int, n_pts: c_int,
                              px: *const c_double, py: *const c_double, pz: *const c_double, opt: *const ccrs_options,
                              summary: *mut ccrs_summary, device_id: c_int) -> c_int;
    pub fn ccrs_comm_unique_id(unique_id_128: *mut c_void) -> c_int;
    pub fn ccrs_comm_init(p: *mut ccrs_problem, unique_id_128: *const c_void, rank: c_int, world_size: c_int) -> c_int;
    pub fn ccrs_release_cached_memory() -> c_int;
}

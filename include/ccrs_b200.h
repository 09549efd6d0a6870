/* ccrs_b200.h — C ABI of the B200-native linearisation library (libccrs_b200.so).
 *
 * The reference (powei-lin/camera-intrinsic-calibration-rs v0.11.2) has no FFI: its hot path sits
 * behind two Rust-level interfaces:
 *   (1) tiny_solver::factors::Factor::residual_func  (src/optimization/factors.rs:152-173, :204-228),
 *       registered per corner by Problem::add_residual_block with HuberLoss::new(1.0)
 *       (src/util.rs:409-414, :603-631), constrained by set_variable_bounds / fix_variable
 *       (src/util.rs:29-71) and solved by GaussNewtonOptimizer::optimize (src/util.rs:443-464, :668-670);
 *   (2) the entry points calib_camera (src/util.rs:384-490) and
 *       calib_all_camera_with_extrinsics (src/util.rs:567-715).
 * Per-corner dual-number callbacks cannot cross to a GPU, so the boundary moves up to
 * "whole problem in, reduced normal equations / converged parameters out". Every entry point below
 * names the reference interface it replaces. A Rust `-sys` crate binding these symbols is under
 * camera-intrinsic-calibration-rs_b200/rust/ (source only: no Rust toolchain in the build image).
 *
 * Conventions: plain pointers and sizes, FP64 everywhere, all arrays are caller-owned HOST memory
 * unless a name ends in _dev. Every function returns 0 on success or a negative ccrs_status; no
 * exceptions cross the ABI. A handle is used by one host thread at a time. Work is enqueued on the
 * handle's stream; host outputs are valid on return. There is NO CPU fallback: without an sm_100
 * device ccrs_problem_create fails with CCRS_ERR_NO_DEVICE.
 */
#ifndef CCRS_B200_H
#define CCRS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CCRS_ABI_VERSION 2

/* GenericModel variants of camera-intrinsic-model ^0.8 (CLI names: src/bin/camera_calibration.rs:35).
 * Parameter order:
 *   UCM     fx fy cx cy alpha              EUCM    fx fy cx cy alpha beta
 *   EUCMT   fx fy cx cy alpha beta t1 t2   KB4     fx fy cx cy k1 k2 k3 k4
 *   OPENCV5 fx fy cx cy k1 k2 p1 p2 k3     FTHETA  fx fy cx cy k1 k2 k3 k4                         */
enum ccrs_model { CCRS_UCM = 0, CCRS_EUCM = 1, CCRS_EUCMT = 2, CCRS_KB4 = 3, CCRS_OPENCV5 = 4, CCRS_FTHETA = 5 };

enum ccrs_status {
  CCRS_OK = 0,
  CCRS_ERR_INVALID = -1,    /* bad argument */
  CCRS_ERR_CUDA = -2,       /* CUDA runtime error (see ccrs_last_error) */
  CCRS_ERR_NO_DEVICE = -3,  /* no sm_100 device: the library has no CPU path */
  CCRS_ERR_NUMERIC = -4,    /* NaN error — tiny-solver returns None */
  CCRS_ERR_CHOLESKY = -5,   /* non-positive pivot — tiny-solver returns None */
  CCRS_ERR_COMM = -6        /* collective failed / NCCL unavailable */
};

typedef struct ccrs_problem ccrs_problem; /* opaque: device buffers, stream, (optional) communicator */

/* Number of parameters of `model` (full vector, fy included). */
int ccrs_model_nparams(int model);

/* Thread-local text of the last error on this thread. */
const char* ccrs_last_error(void);

/* ---- problem construction: replaces calib_camera's Problem assembly (src/util.rs:399-441) ---------
 * One 2-residual block per (frame, corner) = ReprojectionFactor::new (factors.rs:134-150) +
 * add_residual_block(2, ["params","rvec{i}","tvec{i}"], .., HuberLoss(1.0)) (util.rs:407-414).
 * Observations are SoA, CSR by frame: frame f owns [frame_offsets[f], frame_offsets[f+1]).
 * x,y,z = FeaturePoint.p3d, u,v = FeaturePoint.p2d widened f32->f64 (factors.rs:141-143).
 * xy_same_focal: the optimised intrinsic vector has fy removed (util.rs:391-395), d = nparams-1.
 * huber_delta <= 0 disables the loss. The arrays are copied to the device; nothing is retained. */
int ccrs_problem_create(ccrs_problem** out, int model, int width, int height, int xy_same_focal,
                        int n_frames, const int32_t* frame_offsets,
                        const double* x, const double* y, const double* z,
                        const double* u, const double* v,
                        double huber_delta, int device_id);

/* Same, taking the observations as the f32 values the reference stores (FeaturePoint{p2d: glam::Vec2, p3d: glam::Vec3},
 * src/detected_points.rs:6-9): half the host->device traffic and half the HBM bytes per observation (20 B). The kernels
 * widen f32 -> f64 on load exactly like ReprojectionFactor::new (factors.rs:141-143), so results are bit-identical to
 * ccrs_problem_create on the widened arrays. */
int ccrs_problem_create_f32(ccrs_problem** out, int model, int width, int height, int xy_same_focal,
                            int n_frames, const int32_t* frame_offsets,
                            const float* x, const float* y, const float* z,
                            const float* u, const float* v,
                            double huber_delta, int device_id);

/* Same problem in the reference's own data model: FrameFeature.features is a map corner id -> FeaturePoint
 * (src/detected_points.rs:13-17) and every p3d is the board point of that id (Board::init_aprilgrid id -> 3-D,
 * src/board.rs:46-95). corner_id[N] indexes board_xyz[n_board][3] (f32, row-major x y z); u, v as above. 12 bytes per
 * observation cross PCIe instead of 20; the device expands p3d = board[id] once, so every kernel and every result is
 * bit-identical to ccrs_problem_create_f32 on the expanded arrays. n_board <= 4096. */
int ccrs_problem_create_board_f32(ccrs_problem** out, int model, int width, int height, int xy_same_focal,
                                  int n_frames, const int32_t* frame_offsets, const int32_t* corner_id,
                                  const float* u, const float* v, const float* board_xyz, int n_board,
                                  double huber_delta, int device_id);

/* New detections for an existing single-problem handle — a calibration service that re-runs calib_camera
 * (src/util.rs:384-490) on fresh detections of the same recording keeps its device buffers, pools and kernel
 * configuration: only the observations cross PCIe again. Same number of frames and the same total number of
 * observations as at creation (frame_offsets may distribute them differently); the arrays have the element type the
 * handle was created with (double / float); handles created in board format take corner_id (x, y, z ignored), the others
 * x, y, z (corner_id ignored). Poses and the solver state are reset like after ccrs_problem_create*. */
int ccrs_problem_update_observations(ccrs_problem* p, const int32_t* frame_offsets, const int32_t* corner_id, const void* x,
                                     const void* y, const void* z, const void* u, const void* v);

/* Batch of independent calibrations in one handle (BASELINE config 5): problem b owns frames
 * [problem_frame_offsets[b], problem_frame_offsets[b+1]). Same model/size/flags for all problems.
 * Intrinsic arrays passed to the calls below are then [n_problems][d]; scalars become [n_problems]. */
int ccrs_batch_create(ccrs_problem** out, int model, int width, int height, int xy_same_focal,
                      int n_problems, const int32_t* problem_frame_offsets,
                      int n_frames, const int32_t* frame_offsets,
                      const double* x, const double* y, const double* z,
                      const double* u, const double* v,
                      double huber_delta, int device_id);

int ccrs_problem_destroy(ccrs_problem* p);
/* Device / pinned buffers and streams of destroyed handles are cached process-wide (cudaMalloc/cudaFree cost more
 * than a whole solve); this returns them to the driver. */
int ccrs_release_cached_memory(void);

/* Sizes. d = optimised intrinsics per problem; nblk = (d+7)(d+8)/2 packed entries per frame block. */
int ccrs_problem_dim(const ccrs_problem* p);
int ccrs_problem_nblk(const ccrs_problem* p);
int ccrs_problem_n_frames(const ccrs_problem* p);
int64_t ccrs_problem_n_obs(const ccrs_problem* p);
int ccrs_problem_n_problems(const ccrs_problem* p);

/* Pose state (the "rvec{i}"/"tvec{i}" variables, util.rs:435-441) lives on the device.
 * poses = [n_frames][6] = rvec(3), tvec(3) per frame (types.rs:13-17 RvecTvec). */
int ccrs_set_poses(ccrs_problem* p, const double* poses);
int ccrs_get_poses(ccrs_problem* p, double* poses);

/* ---- parity hook: replaces ReprojectionFactor::residual_func evaluated with f64 and with duals ----
 * (factors.rs:152-173; tiny-solver residual_and_jacobian + Corrector).
 * r: [2N]; J: [2N][d+6] row-major, columns [intrinsics | rvec | tvec]; J may be NULL.
 * apply_loss != 0 returns the Huber-corrected r and J that tiny-solver assembles.
 * poses may be NULL (use the device pose state). Single-problem handles only. */
int ccrs_eval_rj(ccrs_problem* p, const double* intr, const double* poses, int apply_loss, double* r, double* J);

/* ---- validation: replaces util::validation (src/util.rs:721-795) ----------------------------------------------
 * Per-point reprojection error sqrt(dx^2 + dy^2) WITHOUT the robust loss at (intr, poses) (util.rs:733-745), then
 * median = sorted_errors[N / 2] and avg99 = mean of the N * 99 / 100 smallest errors (util.rs:771-781), computed on
 * the device by radix select (no sort, no per-observation transfer). poses NULL = the device pose state.
 * errors (nullable, [N], observation order) receives the per-point errors the reference logs to rerun.
 * Single-problem handles only. */
int ccrs_validation(ccrs_problem* p, const double* intr, const double* poses, double* median, double* avg99,
                    double* errors);

/* ---- step-wise hot path: replaces Problem::compute_residual_and_jacobian + J^T J assembly + the
 * per-iteration sparse LLT of tiny-solver (call sites util.rs:455,463,670) ---------- */

/* K2: fused residual + analytic Jacobian + Huber + per-frame Gram blocks at the CURRENT (which=0) or
 * TRIAL (which=1) pose state with intrinsics `intr`. sq_err[n_problems] = sum of corrected r^2. */
int ccrs_linearize(ccrs_problem* p, const double* intr, int which, double* sq_err);

/* Per-frame packed blocks of the last linearisation (parity hook): [n_frames][nblk], upper triangle,
 * row-major, of [J r]^T [J r] with column order [intrinsics | rvec | tvec | r]. */
int ccrs_get_frame_blocks(ccrs_problem* p, int which, double* blocks);

/* Jacobi column scaling 1/(1+||J[:,c]||) (tiny-solver LM, first iteration): computes the pose scales on
 * the device from the `which` linearisation and returns the LOCAL squared column norms of the
 * intrinsic columns, col_sq[n_problems][d]; the caller sums them across ranks and calls ccrs_set_intr_scale. */
int ccrs_compute_scale(ccrs_problem* p, int which, double* col_sq);
int ccrs_set_intr_scale(ccrs_problem* p, const double* intr_scale /* [n_problems][d] or NULL = none */);

/* K3: per-frame damping + 6x6 Cholesky elimination + fixed-order reduction onto the intrinsic system.
 * u[n_problems] = LM damping of the POSE blocks (NULL/0 for Gauss-Newton); use_scale applies the Jacobi scaling.
 * out[n_problems][d*d + 3d + 1]:  S (d*d row-major, = A' - sum_f B' C'^-1 B'^T, intrinsic damping NOT yet
 * added: it needs the global diag), g_s (d, reduced rhs), g_a (d, unreduced scaled gradient -J_a^T r),
 * diag_a (d, undamped scaled diag of A), sq_err (sum of corrected r^2). Summed over ranks when a
 * communicator is attached. */
int ccrs_reduce(ccrs_problem* p, int which, const double* u, int use_scale, double min_diag, double max_diag,
                double* out);

/* K4: pose back-substitution y_p = C^-1 (g_p - B^T y_a), trial_poses = poses + D y_p.
 * y_a[n_problems][d] is the (scaled) intrinsic solution, u[n_problems] the damping used in ccrs_reduce (nullable).
 * in_place = 1 writes the update into the current poses (Gauss-Newton). model_dec[n_problems] (nullable)
 * receives the pose part of the LM gain-ratio denominator y^T (2 g' - H' y). */
int ccrs_backsub(ccrs_problem* p, const double* y_a, const double* u, int in_place, double* model_dec);

/* K5: residual-only Huber cost at the current (0) / trial (1) poses. sq_err[n_problems]. */
int ccrs_eval_cost(ccrs_problem* p, const double* intr, int which, double* sq_err);

/* Accept the trial point: trial poses (and, if present, the trial linearisation) become current.
 * mask[n_problems] nullable = accept all. */
int ccrs_accept(ccrs_problem* p, const unsigned char* mask);

/* ---- multi-GPU: frames are sharded across ranks; one exchange of the reduced system per linearisation.
 * The library dlopen()s libnccl.so.2 (the copy torch loaded). unique_id is the 128-byte ncclUniqueId
 * produced by ccrs_comm_unique_id on rank 0 and distributed by the host (torch.distributed / MPI). */
int ccrs_comm_unique_id(void* unique_id_128);
int ccrs_comm_init(ccrs_problem* p, const void* unique_id_128, int rank, int world_size);
/* The communicator is process-wide (one process per GPU): ccrs_comm_init with unique_id_128 == NULL attaches
 * the communicator an earlier call created to another handle. ccrs_comm_finalize destroys it. */
int ccrs_comm_finalize(void);
/* deterministic = 1: all-gather the per-rank partials and sum in rank order on every rank (default);
 * 0: ncclAllReduce(sum) as north_star names. */
int ccrs_comm_set_deterministic(ccrs_problem* p, int deterministic);
/* 1 when the per-iteration exchange runs over peer memory inside K2/K3 (all ranks on one NVLink node; IPC handles
 * were exchanged at ccrs_comm_init), 0 when it falls back to NCCL collectives (CCRS_P2P=0, no P2P, non-deterministic
 * all-reduce requested). */
int ccrs_comm_uses_peer_memory(void);

/* ---- loop controllers (host side): replace GaussNewtonOptimizer::optimize (util.rs:443-464) and
 * tiny-solver's LevenbergMarquardtOptimizer::optimize (named by north_star). ------------------------ */
typedef struct ccrs_options {
  int max_iteration;          /* 100  OptimizerOptions::default() */
  double min_abs_decrease;    /* 1e-5 */
  double min_rel_decrease;    /* 1e-5 */
  double min_error;           /* 1e-10 */
  double lm_initial_radius;   /* 1e4 */
  double lm_min_diag;         /* 1e-6 */
  double lm_max_diag;         /* 1e32 */
  int fixed_mode;             /* 0: fixed variables stay in the system and are reset after the update
                                 (tiny-solver ParameterBlock::update_params); 1: eliminated */
  int speculative;            /* LM: 1 = linearise at the trial point instead of a residual-only pass (default) */
  int verbose;
  double block_huber_delta;   /* GN only, 0 = off: ONE Huber loss over the whole residual vector (a problem that is a
                                 single residual block, ModelConvertFactor util.rs:246-251): changes the error the stop
                                 tests see, not the step. Use with a handle created with huber_delta <= 0. */
} ccrs_options;
void ccrs_default_options(ccrs_options* o);

typedef struct ccrs_summary {
  int iterations;   /* linearisations (GN) / LM iterations performed (max over problems) */
  int status;       /* ccrs_status */
  int stop_reason;  /* 0 max_iter, 1 error<min, 2 abs decrease, 3 rel decrease (problem 0) */
  double final_error;
  int n_accepted, n_rejected;
  double device_ms; /* CUDA-event time of the whole loop */
} ccrs_summary;

/* intr[n_problems][d] in/out; lo/hi [d] nullable (set_variable_bounds, util.rs:29-49);
 * fixed[d] nullable: 1 = fix_variable (util.rs:50-71, :459-464; handled per fixed_mode), 2 = not a variable of the
 * problem at all (always removed from the linear system; UCMInitFocalAlphaFactor's constant cx, cy).
 * Poses are the device pose state.
 * err_hist nullable [max_iteration] (problem 0). */
int ccrs_solve_gn(ccrs_problem* p, double* intr, const double* lo, const double* hi, const unsigned char* fixed,
                  const ccrs_options* opt, ccrs_summary* summary, double* err_hist);
int ccrs_solve_lm(ccrs_problem* p, double* intr, const double* lo, const double* hi, const unsigned char* fixed,
                  const ccrs_options* opt, ccrs_summary* summary, double* err_hist);

/* ---- backend-agnostic controllers: the same GN / LM loops driven through a table of callbacks.
 * ccrs_solve_gn/lm call these with the CUDA backend; CPU tests (world_size-2 gloo) drive them with
 * their own callbacks. All buffers host memory; n_problems systems of dimension d. */
typedef struct ccrs_backend {
  void* ctx;
  int d;
  int n_problems;
  /* K2 at the current (0) / trial (1) point; asynchronous, no output */
  int (*linearize)(void* ctx, const double* intr, int which);
  /* local squared intrinsic column norms [n_problems][d]; pose scales stay inside the backend */
  int (*compute_scale)(void* ctx, int which, double* col_sq);
  int (*set_intr_scale)(void* ctx, const double* intr_scale /* nullable */);
  /* K3: out[n_problems][d*d + 3d + 1] = S | g_s | g_a | diag_a | sq_err  (see ccrs_reduce) */
  int (*reduce)(void* ctx, int which, const double* u, int use_scale, double min_diag, double max_diag, double* out);
  /* K4: active[n_problems] nullable; in_place = 1 updates the current poses (Gauss-Newton) */
  int (*backsub)(void* ctx, const double* y_a, const double* u, const unsigned char* active, int in_place);
  /* out[n_problems][2] = pose part of the model decrease, sq_err at the trial point.
   * speculative: obtain sq_err by linearising the trial point (its blocks become current on accept) */
  int (*trial_stats)(void* ctx, const double* intr_trial, int speculative, double* out);
  int (*accept)(void* ctx, const unsigned char* mask);
  /* in-place sum over ranks, identical result on every rank. NULL = the outputs above are already global
   * (single rank, or the backend exchanges on the device as the CUDA backend does with NCCL). */
  int (*allreduce)(void* ctx, double* buf, int count);
} ccrs_backend;

int ccrs_controller_gn(const ccrs_backend* be, double* intr, const double* lo, const double* hi,
                       const unsigned char* fixed, const ccrs_options* opt, ccrs_summary* summary, double* err_hist);
int ccrs_controller_lm(const ccrs_backend* be, double* intr, const double* lo, const double* hi,
                       const unsigned char* fixed, const ccrs_options* opt, ccrs_summary* summary, double* err_hist);

/* ---- reference entry point: calib_camera (src/util.rs:384-490) from "Problem assembled" onwards ----
 * params[nparams] in/out (FULL vector incl. fy); poses[n_frames][6] in/out (initial = SQPnP poses the
 * caller computed, util.rs:435-441). Semantics of xy_same_focal / disabled_distortions / fixed_focal as
 * util.rs:391-395, :446-454, :459-464. use_lm = 0 runs Gauss-Newton like the reference. */
int ccrs_calib_camera(int model, int width, int height, int n_frames, const int32_t* frame_offsets,
                      const double* x, const double* y, const double* z, const double* u, const double* v,
                      double* params, double* poses, int xy_same_focal, int disabled_distortions, int fixed_focal,
                      int use_lm, const ccrs_options* opt, ccrs_summary* summary, int device_id);

/* ---- initial board poses: replaces the per-frame sqpnp_simple::sqpnp_solve_glam(&p3ds, &p2ds_z) of calib_camera
 * (src/util.rs:418-439) and of init_pose (src/optimization/linear.rs:5-21) -----------------------------------------
 * x, y, z: board points; xn, yn: the matching NORMALISED image points (p2.x / p2.z, p2.y / p2.z of
 * generic_camera.unproject, util.rs:418-429, or init_pose's radial approximation) — unprojection belongs to the model
 * crate and stays with the caller. All frames are solved in one launch (one warp per frame): minimiser over SO(3) x R^3
 * of the SQPnP cost sum_i (R p_i + t)^T Q_i (R p_i + t) with the board in front of the camera.
 * poses_out [n_frames][6] = rvec, tvec (RvecTvec, types.rs:13-17); cost_out [n_frames] nullable = the minimum.
 * Every frame needs >= 4 points (the reference skips frames with fewer than 10, util.rs:431-433). */
int ccrs_init_poses(int n_frames, const int32_t* frame_offsets, const double* x, const double* y, const double* z,
                    const double* xn, const double* yn, double* poses_out, double* cost_out, int device_id);

/* Poses as constants: the per-frame pose blocks are not variables (K3 skips their elimination, K4 leaves them
 * untouched). Intrinsics-only problems such as ModelConvertFactor (factors.rs:11-77). */
int ccrs_set_fixed_poses(ccrs_problem* p, int fixed);

/* ---- init_ucm (src/util.rs:284-378) ------------------------------------------------------------------------------
 * Stage 1 (util.rs:295-357): Gauss-Newton over "params" = [f, alpha] of a UCM whose principal point is the image
 * centre (UCMInitFocalAlphaFactor, factors.rs:83-120) and the poses of the given frames (two in the reference), Huber
 * loss 1.0, f in [init_f / 3, 3 init_f], alpha in [1e-6, 1], fixed_focal = fix_variable("params", 0). It runs on the
 * ReprojectionFactor kernels as a one-focal UCM problem whose cx, cy are removed from the linear system.
 * Stage 2 (util.rs:358-372): calib_camera(frames, UCM[f, f, w/2, h/2, alpha], one_focal = true, 0, fixed_focal), which
 * like the reference starts from fresh poses: every detection is unprojected with the stage-1 model (UCM: closed form)
 * and the PnP of each frame is solved by ccrs_init_poses (util.rs:418-439); the stage-1 poses are dropped.
 * params_out[5] = fx fy cx cy alpha; poses[n_frames][6] in/out. */
int ccrs_init_ucm(int width, int height, int n_frames, const int32_t* frame_offsets,
                  const double* x, const double* y, const double* z, const double* u, const double* v,
                  double init_f, double init_alpha, int fixed_focal, double* poses, double* params_out,
                  const ccrs_options* opt, ccrs_summary* summary, int device_id);

/* ---- convert_model (src/util.rs:225-278) -------------------------------------------------------------------------
 * Fit the target model to the source model on n_pts 3D points (px, py, pz: the source's unprojected pixel grid,
 * ModelConvertFactor::new factors.rs:22-48 — unprojection belongs to the model crate and stays with the caller).
 * UCM -> EUCM / EUCMT is the reference's closed form (util.rs:230-243). Otherwise: the source projections are evaluated
 * on the device, the target's parameters start from tgt_params with fx fy cx cy taken from the source (util.rs:253-255),
 * bounds and disabled distortions as util.rs:262-271, one Huber(1.0) over the whole block, Gauss-Newton on K2/K3 with
 * the (identity) pose held constant. tgt_params[nparams(tgt)] in/out. */
int ccrs_convert_model(int src_model, const double* src_params, int tgt_model, double* tgt_params, int width, int height,
                       int disabled_distortions, int n_pts, const double* px, const double* py, const double* pz,
                       const ccrs_options* opt, ccrs_summary* summary, int device_id);

/* ---- joint multi-camera refinement: calib_all_camera_with_extrinsics (src/util.rs:567-715) ----------------------
 * Variables "params{c}" (d per camera), "rvec_{c}_0"/"tvec_{c}_0" (camera c <- camera 0, c > 0) and
 * "rvec_0_b_{f}"/"tvec_0_b_{f}" (board -> camera 0, per frame, shared by all cameras). A block is the set of corners one
 * camera detected in one frame: cam0 blocks are ReprojectionFactor blocks (util.rs:603-611), the others
 * OtherCamReprojectionFactor blocks (util.rs:612-631, factors.rs:204-228). block_offsets is CSR over the SoA
 * observation arrays. The board poses are eliminated per frame on the device; the host solves the shared system. */
typedef struct ccrs_joint ccrs_joint;
int ccrs_joint_create(ccrs_joint** out, int model, int xy_same_focal, int n_cams, int n_frames, int n_blocks,
                      const int32_t* block_cam, const int32_t* block_frame, const int32_t* block_offsets,
                      const double* x, const double* y, const double* z, const double* u, const double* v,
                      double huber_delta, int device_id);
int ccrs_joint_destroy(ccrs_joint* p);
int ccrs_joint_dim(const ccrs_joint* p);
const char* ccrs_joint_last_error(void);
int64_t ccrs_joint_launch_count(const ccrs_joint* p);
/* parity hook for OtherCamReprojectionFactor::residual_func (factors.rs:204-228): r [2N], J [2N][d+12] with columns
 * [params_c | rvec_0_b tvec_0_b | rvec_c_0 tvec_c_0] (the last six are zero for cam0 blocks). intr [n_cams][d],
 * extr [n_cams][6] (row 0 ignored), poses [n_frames][6]. */
int ccrs_joint_eval_rj(ccrs_joint* p, const double* intr, const double* extr, const double* poses, int apply_loss,
                       double* r, double* J);
/* GaussNewtonOptimizer::optimize on the joint problem (util.rs:668-670). intr, extr, poses in/out; lo/hi/fixed are
 * [n_cams][d], nullable (set_problem_parameter_bound / _disabled per camera util.rs:654-663; cam0_fixed_focal =
 * fixed[0], util.rs:664-667). extr row 0 is returned as zeros (util.rs:689-690). */
int ccrs_joint_solve_gn(ccrs_joint* p, double* intr, double* extr, double* poses, const double* lo, const double* hi,
                        const unsigned char* fixed, const ccrs_options* opt, ccrs_summary* summary, double* err_hist);

/* Distortion bounds of GenericModel::distortion_params_bound() for the FULL parameter vector
 * (lo/hi [nparams]; +-inf where unbounded), plus fx,fy in [0,1e4], cx in [0,w], cy in [0,h] (util.rs:36-39). */
int ccrs_model_bounds(int model, int width, int height, double* lo, double* hi);

/* ---- measurement helpers ------------------------------------------------------------------------ */
/* FP64 FMA throughput microbenchmark (TFLOP/s) — the roofline denominator SURVEY §8(d) asks to measure. */
int ccrs_measure_fp64_peak(int device_id, double* tflops);
/* Time `reps` back-to-back launches of K2 (linearise) with CUDA events on the handle's stream. */
int ccrs_time_linearize(ccrs_problem* p, const double* intr, int reps, int flush_l2, double* avg_ms);
/* `warmup` untimed + `steps` timed LM iterations with the stop tests disabled, each iteration bracketed by CUDA
 * events on the handle's stream (step_ms[steps]). Every `reset_every` iterations the state returns to
 * (intr0, poses0) outside the timed bracket, so every timed step is one of the first `reset_every` LM iterations of
 * the problem. flush_l2 writes a 512 MB buffer before every iteration, also outside the bracket.
 * timed_launches = kernels launched inside the timed iterations. */
int ccrs_bench_lm_steps(ccrs_problem* p, const double* intr0, const double* poses0, int warmup, int steps,
                        int reset_every, int flush_l2, double* step_ms, int64_t* timed_launches);
/* The same LM iteration measured the way a solve runs it — K2, K3, K2, K3, ... enqueued back to back, no host
 * synchronisation in between — with cold caches: `n_ps` replicas of one single-problem handle (create them with the same
 * arguments; together they must exceed the L2 cache) are visited round-robin, step i on replica i mod n_ps, so that a
 * replica's arrays have left the L2 cache when its turn comes again. `warmup` untimed steps, then `steps` timed steps
 * inside CUDA-event brackets (total_ms; one bracket per 4 x n_ps steps: every replica runs LM iterations 1-4 of the
 * problem and is then returned to the start point, untimed); stop tests disabled; device-driven loop only.
 * executed_steps = linearisations the timed slots really executed (across GPUs the slot behind a mis-speculated
 * reduction exits at once: time spent, not a step). */
int ccrs_bench_lm_steps_rotating(ccrs_problem** ps, int n_ps, const double* intr0, const double* poses0, int warmup, int steps,
                                 double* total_ms, int64_t* timed_launches, int64_t* executed_steps);
/* Host-side phase trace of single-problem LM iterations (process-wide). Returns the averages accumulated since the
 * last call, in microseconds per iteration, then resets and enables/disables tracing:
 *   [0] K3 launch call  [1] K3 execution + publish latency  [2] host: unpack, d x d solve, trial point
 *   [3] K2 launch call  [4] K2 execution + publish latency  [5] host: accept/reject, bookkeeping */
int ccrs_step_trace(int enable, double* avg_us, int64_t* n_iterations);
/* Speculative K3: in the speculative LM loop of a single problem the library launches the next reduction right behind
 * the trial-point linearisation, for the outcome "accepted with gain ratio >= 0.937" (u_next = u / 3), so that launch
 * and kernel leave the critical path; a different decision by the controller just launches K3 again. Process-wide
 * counters (speculative launches, launches whose result was used); returns 1 if enabled (CCRS_SPEC_K3=0 disables). */
int ccrs_spec_k3_counters(int64_t* launched, int64_t* hits);
/* Device-driven loop (single-problem handles; ccrs_solve_gn / ccrs_solve_lm / ccrs_calib_camera use it whenever the
 * handle has no communicator or exchanges over peer memory): the host enqueues K3, K2, K3, K2, ... ahead of time, each a
 * programmatic dependent of the one before; the last CTA of K3 runs the controller rule (accept / reject, damping, stop
 * tests, d x d solve, clamp — the same source as the host controllers, csrc/ccrs_rule.h) and leaves the next
 * linearisation point in device memory, so no host round trip sits between two kernels. The host audits every
 * published iteration record by re-running the rule (decisions and solves must agree bit for bit) and ends the loop.
 * CCRS_DEVICE_LOOP=0 selects the host-driven controllers instead.
 * ccrs_loop_counters: number of device solves the host has audited (process-wide); returns 1 if the loop is enabled.
 * ccrs_loop_trace: device-side phase trace (globaltimer stamps carried by the records), microseconds per iteration
 * averaged since the last call, then resets and enables/disables tracing:
 *   [0] K2: first warp past its dependency wait -> last warp done   [1] K2 done -> K3's last CTA past its wait
 *   [2] K3 per-frame elimination + CTA sums (last CTA)              [3] K3 tail: cross-CTA sum, exchange, controller rule
 *   [4] record ready -> next K2 running
 *   [5..11] finer split of [2] and [3] in the last CTA of K3: control block + decision | block load | per-frame
 *   elimination | CTA sum + partial store | cross-CTA sum (+ exchange) | staging | controller rule.
 *   [12] the part of [4] up to the first K2 warp leaving its dependency wait (the rest is its prologue loads).  avg_us: [13] */
int ccrs_loop_counters(int64_t* audited_solves);
int ccrs_loop_trace(int enable, double* avg_us, int64_t* n_iterations);
/* Kernel launches issued by this handle since creation. */
int64_t ccrs_launch_count(const ccrs_problem* p);

#ifdef __cplusplus
}
#endif
#endif /* CCRS_B200_H */

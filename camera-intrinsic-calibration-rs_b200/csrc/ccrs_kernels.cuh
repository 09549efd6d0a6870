// ccrs_kernels.cuh — kernel parameter blocks and launch entry points shared by ccrs_kernels.cu / ccrs_api.cu.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace ccrs {

constexpr int kLinThreads = 128;     // K2 CTA size
constexpr int kLinWarps = kLinThreads / 32;
constexpr double kLinOverheadIters = 6.0;  // K2 per-warp prologue + reduction cost in main-loop iterations (slicing cost model)
constexpr int kLinCtasPerSm = 2;     // K2 __launch_bounds__ occupancy target (255 regs x 128 x 2 = one register file)
constexpr int kRedChunk = 48;        // accumulators staged per smem reduction round (48 x 32 x 8 B = 12 KB / warp)
#ifndef CCRS_OBS_STAGES
#define CCRS_OBS_STAGES 4
#endif
constexpr int kObsStages = CCRS_OBS_STAGES;   // cp.async ring depth of K2's observation prefetch (distance 3 iterations; 5, 6, 8 measured slower warm AND cold)
constexpr int kFrameConst = 21;      // R(9) t(3) Jl(9) per frame in shared memory

// Observation arrays and per-frame state as the kernels see them.
struct ProblemDev {
  const double *x, *y, *z, *u, *v;        // [N] SoA; when f32 != 0 the arrays hold floats (FeaturePoint is f32, detected_points.rs:6-9)
  int f32;
  const int32_t* frame_offsets;            // [F+1]
  const int32_t* frame_problem;            // [F] problem of each frame (batch) — nullptr for a single problem
  const int32_t* problem_frame_offsets;    // [n_problems+1]
  const int32_t* obs_frame;                // [N] frame of each observation (K1 only)
  const int32_t* cur;                      // [n_problems] which of the two state buffers is "current" (batch)
  int cur_val;                             // ... single problem: tracked by the host, passed by value (cur == nullptr)
  double* poses[2];                        // [F][6]
  double* blocks[2];                       // [NBLK][Fs] SoA packed frame blocks
  double* frame_cost[2];                   // [Fs] per-frame sum of corrected r^2 (cost-only pass)
  int n_frames, n_problems, Fs;            // Fs = padded frame stride
  double huber_delta;
};

// Cross-GPU exchange fused into the producing kernel (frame-sharded problems, one process per GPU on one NVLink
// node): the thread that holds value v of this rank's partial result stores it straight into slot [rank][v] of EVERY
// rank's exchange buffer over peer memory, spins until the slots of all ranks in its own buffer are filled, sums them
// in rank order (bitwise identical on all ranks) and re-arms the slots. Slots validate themselves (armed bit pattern),
// so there is no fence, no flag and no collective launch. world <= 1 disables it.
constexpr int kXchgMaxRanks = 8;
constexpr int kXchgMaxVals = 96;
struct PeerXchg {
  double* peer[kXchgMaxRanks];   // rank r's buffer as mapped in this process (peer[rank] is the local buffer)
  int world, rank;
  int off;                       // double offset of the [world][kXchgMaxVals] slot area used by this exchange
};

// ---- device-driven loop (ccrs_loop.cu) ---------------------------------------------------------------------------
// The Gauss-Newton / Levenberg-Marquardt iteration of a single problem runs without a host round trip: the host
// enqueues K3, K2, K3, K2, ... (each a programmatic dependent of the one before), the last CTA of K3 executes the
// controller rule (ccrs_rule.h: d x d solve, clamp, accept / reject, damping update, stop tests) on the reduced system
// it has just summed and leaves {phase, intrinsics to linearise at, intrinsic step, damping} in this block for K2; K2's
// last warp leaves {pose part of the model decrease, cost}. A slot whose turn it is not (phase mismatch, loop done)
// exits at once. Every executed K3 publishes one record to mapped host memory; the host audits it by re-running the
// same rule, fills the summary and stops enqueueing when a record says the loop is done.
enum LoopPhase : int {
  PH_LIN0 = 0,     // K2: linearise the start point (no step to apply)
  PH_REDUCE = 1,   // K3: reduce blocks[cur] with damping ctl.u (first iteration, or after a rejected / mis-speculated step)
  PH_TRIAL = 2,    // K2: apply the step (LM: into the trial buffers; GN: in place) and linearise there
  PH_DECIDE = 3,   // K3: K2's statistics are in; LM: accept / reject, then reduce; GN: stop tests, then reduce
  PH_DONE = 5
};
constexpr int kRecStride = 208;   // doubles per published record
constexpr int kRecSlots = 32;     // ring of records in mapped host memory
// record layout (doubles)
enum : int {
  REC_SEQ = 0, REC_PHASE, REC_STATUS, REC_STOP, REC_IT, REC_ITERATIONS, REC_ACCEPTED, REC_CUR, REC_RHO, REC_U, REC_V,
  REC_CUR_ERR, REC_U_SOLVE, REC_MD_A, REC_SQ_NEW, REC_MD_POSE, REC_SQ_CUR_BEFORE, REC_U_BEFORE, REC_V_BEFORE,
  REC_CUR_ERR_BEFORE, REC_MD_A_BEFORE, REC_HIST_IDX, REC_HIST_VAL, REC_SOLVED, REC_N_ACC, REC_N_REJ, REC_FINAL_ERR,
  REC_SQ_CUR, REC_LAST_ERR_BEFORE, REC_DECIDED,
  // device timestamps (globaltimer, low 40 bits, ns): the K2 in front of this K3 (first warp after its wait / last warp
  // at its end), this K3 (entry of the last CTA after its wait, start of the last CTA's tail, record ready)
  REC_T_K2_BEGIN, REC_T_K2_END, REC_T_K3_BEGIN, REC_T_TAIL, REC_T_END,
  REC_T_HEAD, REC_T_LOAD, REC_T_COMP, REC_T_SUMMED, REC_T_RULE,   // finer stamps of the last CTA (tools/loop_trace.py)
  REC_T_K2_WAKE,
  REC_LAST_SCALAR,
  REC_INTR = 48, REC_TRIAL = 57, REC_Y = 66, REC_SCALE = 75, REC_OUT = 84   // out: d*d + 3d + 1 <= 109
};
struct LoopCtl {
  // ---- configuration: constant during a loop
  int mode;                 // 0 Gauss-Newton, 1 Levenberg-Marquardt
  int D, max_iteration, fixed_mode, has_bounds, has_fixed;
  double min_abs, min_rel, min_error, min_diag, max_diag, block_huber;
  double lo[9], hi[9];
  unsigned char fixed[16];
  // ---- state
  // Hot window: everything a K2 warp needs from this block, kCtlHotWords consecutive 8-byte words that the warp reads
  // with ONE lane-distributed load (lane i <- word i). Read field by field, every one of the ~1,200 warps of a K2 grid
  // sent 16-22 requests to the same two L2 lines right after the K3 -> K2 hand-off.
  int phase, cur;           // word 0
  int mode_k2, pad_k2;      // word 1: copy of `mode`
  double u_used;            // word 2: damping of the reduction whose elimination record K2 back-substitutes with
  double trial[9];          // words 3..11: intrinsics K2 linearises at next
  double step[9];           // words 12..20: intrinsic step in unscaled units (Jacobi scale x y_a): what K2's back-substitution uses
  int it, iterations, first, seq, status, stop_reason, n_acc, n_rej;
  double u, v;              // LM damping (1 / radius) and reject factor
  double intr[9];           // current intrinsics (optimised vector)
  double scale[9];          // Jacobi scaling of the intrinsic columns (LM; 1 for GN)
  double md_a;              // intrinsic part of the model decrease of the pending trial step
  double sq_cur, cur_err, last_err, final_err;
  double stat[2];           // K2 (this rank): {pose part of the model decrease, sum of corrected r^2} at the point it linearised
  double t_k2_begin, t_k2_end;   // device timestamps of the last executed K2 (see REC_T_*)
  double t_k2_wake;              // ... its first warp straight after the dependency wait (before the prologue loads)
};
static_assert(sizeof(LoopCtl) % 8 == 0, "LoopCtl is copied as 8-byte words");
constexpr int kCtlHotWords = 21;
static_assert(REC_LAST_SCALAR <= REC_INTR && REC_OUT + 9 * 9 + 3 * 9 + 1 <= kRecStride, "record layout overlaps");

struct LinParams {
  ProblemDev pb;
  LoopCtl* ctl;             // device-driven loop: phase, intrinsics, step, damping, buffer selector come from here
  const double* intr_dev;   // [n_problems][D] (batch) — single problem passes intr[] by value below
  double intr[9];           // full vector fx fy cx cy k.. (single problem; constant bank operands)
  const int32_t* acc_to_blk;  // [NACC] sparse accumulator -> dense packed block index
  int which;                // 0 = current point, 1 = trial point
  int G;                    // lanes per frame
  int FPW;                  // frames per warp = 32 / G
  // ---- tensor-core variant (ccrs_linmma.cu): frames are handed out dynamically, statistics summed per chunk of 16 frames
  unsigned long long* frame_ctr;  // next frame to hand out (reset by the launch's last grab)
  double* frame_stat;       // [Fs][2] per-frame {model decrease, cost}, self-validating slots (armed)
  unsigned int* chunk_cnt;  // [ceil(F / 16)] frames of the chunk done so far (reset by the warp that closes the chunk)
  // ---- fused K4 (pose back-substitution) in the prologue: 0 none, 1 trial = current + step, 2 in place (GN)
  int backsub;
  const double* elim;       // [(6D+18)][Fs] from K3
  const double* pose_scale; // [6][Fs] or nullptr
  double y_a[9];            // single problem: scaled intrinsic solution by value
  double u;                 // single problem: damping used by K3
  const double* ya_dev;     // batch: [n_problems][D]
  const double* u_dev;      // batch: [n_problems]
  const unsigned char* active;  // batch: problems whose poses may move (nullable)
  double* frame_md;         // [Fs] per-frame model-decrease part (batch reduces it per problem)
  // ---- fused statistics (single problem): per-CTA partial {md, cost}, final sum by the last CTA in CTA order
  double* cta_part;         // [n_ctas * kLinWarps][2] per-warp partials, armed (k_arm) between launches; 16-byte aligned
  unsigned int* ticket;
  double* stat_dev;         // [2] = {md, cost}
  volatile double* host_stat;  // mapped pinned [4] = {md, cost, -, seq}; nullptr when a cross-rank exchange follows
  double seq;
  long long* dbg;           // [n_warps][10] per-warp phase clocks (only written by -DCCRS_K2_TIMING builds)
  PeerXchg px;              // cross-GPU sum of {md, cost} inside the kernel (px.world > 1)
};

struct SchurParams {
  ProblemDev pb;
  int which;
  const double* u_dev;          // [n_problems] damping
  const double* intr_scale;     // [n_problems][D] or nullptr
  const double* pose_scale;     // [6][Fs] or nullptr
  double u_val;                 // single problem: damping by value (u_dev == nullptr)
  double min_diag, max_diag;
  int no_pose;                  // poses are constants of the problem (ModelConvertFactor): B' = 0, g'_p = 0, C' = I
  const unsigned char* active;  // batch, nullable: frames of problems that have stopped are skipped
  double* elim;                 // [(6D+18)][Fs]: X (6xD), cg (6), g'_p (6), Dd (6)
  double* frame_red;            // batch: [NRED][Fs] per-frame contributions; single: [n_ctas][NRED] CTA partials
  // single problem: the last CTA sums the CTA partials in CTA order
  unsigned int* ticket;
  double* red_out;              // [NRED] device
  volatile double* host_red;    // mapped pinned [NRED+1] (last = seq); nullptr when a cross-rank exchange follows
  double seq;
  long long* dbg;               // [n_warps][8] per-warp phase clocks (only written by -DCCRS_K2_TIMING builds)
  PeerXchg px;                  // cross-GPU sum of the reduced system inside the kernel (px.world > 1)
};

struct BacksubParams {
  ProblemDev pb;
  const double* elim;
  const double* y_a;            // [n_problems][D] device
  const double* u_dev;
  const double* pose_scale;     // [6][Fs] or nullptr
  double* frame_md;             // [Fs] per-frame model-decrease part (nullable)
  int in_place;                 // GN: write the update into the current poses
  const unsigned char* active;  // [n_problems] nullable: problems whose poses must not move
};

// K3, single problem (ccrs_loop.cu): eight lanes per frame; Jacobi scaling of the intrinsic columns factored out of the
// per-frame work and applied after the sum, so that the first LM reduction can compute it itself.
struct Schur2Params {
  ProblemDev pb;
  LoopCtl* ctl;                 // nullable: host-driven call (u_val / which / scales below)
  int which;
  double u_val;
  const double* intr_scale;     // [D] device or nullptr (host-driven)
  int use_pose_scale;           // host-driven: 1 = read pose_scale
  double* pose_scale;           // [6][Fs]
  double min_diag, max_diag;
  int no_pose;
  double* elim;                 // [(6D+18)][Fs]: X (6xD, WITHOUT the intrinsic scaling), cg (6), g'_p (6), Dd (6)
  double* partials;             // [n_ctas][NRED]
  unsigned int* ticket;
  double* red_out;              // [NRED] device (host-driven)
  volatile double* host_red;    // mapped pinned [NRED] (host-driven; nullptr when a cross-rank NCCL exchange follows)
  volatile double* rec;         // mapped pinned [kRecSlots][kRecStride] (ctl mode)
  PeerXchg px;
  unsigned int* xchg_count;     // device counter of executed exchanges (parity of the peer slots), nullable
};
cudaError_t launch_schur2(int D, const Schur2Params& prm, int n_frames, bool pdl, cudaStream_t s);
int schur2_ctas(int n_frames);   // CTAs (= partial slots) of a k_schur2 launch

// ---- device-driven loop for a BATCH of independent problems (ccrs_loop.cu) -----------------------------------------
// Every problem carries its own controller state; two small kernels (one CTA per problem) run the rule of ccrs_rule.h:
//   k_batch_solve : per-problem sum of K3's frame contributions -> (GN: stop tests) damped d x d solve -> next point
//   k_batch_decide: per-problem sum of K2's trial statistics -> LM accept / reject, damping, stop tests
// and write straight into the arrays K2 / K3 read (intrinsics, step, damping, buffer selector, active mask). The host
// enqueues K3, solve, K2, decide ahead of time and only reads the number of still-active problems per iteration from
// mapped host memory: no per-iteration memcpy, no stream synchronise.
struct BatchCtl {
  double u, v, cur_err, last_err, sq_cur, md_a, final_err;
  double intr[9], trial[9], scale[9];
  int it, iterations, active, status, stop, n_acc, n_rej, pad;
};
struct BatchRuleParams {
  ProblemDev pb;
  BatchCtl* ctl;                // [n_problems]
  int lm, D, NRED, max_iteration, fixed_mode, has_bounds, has_fixed, rr_idx;
  double min_abs, min_rel, min_error, min_diag, max_diag;
  double lo[9], hi[9];
  unsigned char fixed[16];
  const double* frame_red;      // [NRED][Fs] K3's per-frame contributions
  const double* frame_md;       // [Fs] K2's per-frame model-decrease parts
  double* intr_dev;             // [n_problems][D]  intrinsics K2 linearises at
  double* ya_dev;               // [n_problems][D]  (scaled) intrinsic step for K2's back-substitution
  double* u_dev;                // [n_problems]     damping for K3 / K2
  int32_t* cur;                 // [n_problems]     buffer selector
  unsigned char* active;        // [n_problems]     K2 / K3 skip the frames of problems that have stopped
  unsigned int* ticket;         // last CTA publishes the iteration status
  unsigned int* n_active;       // device counter
  volatile double* host_status; // mapped pinned ring [kRecSlots][2] = {sequence number, active problems}
  int seq;                      // sequence number of this launch (1-based)
  double* err_hist0;            // device [max_iteration] error history of problem 0 (nullable)
};
cudaError_t launch_batch_solve(const BatchRuleParams& prm, cudaStream_t s);
cudaError_t launch_batch_decide(const BatchRuleParams& prm, cudaStream_t s);

// number of values K3 reduces per problem: S upper (D(D+1)/2) + g_s (D) + g_a (D) + diag_a (D) + sq_err (1)
inline int nred_of(int D) { return D * (D + 1) / 2 + 3 * D + 1; }

int model_dims(int model, int one_focal, int* D, int* NA, int* NBLK, int* NACC);
void fill_acc_to_blk(int model, int one_focal, int32_t* table);
// K2 runs its lane-pair variant for this model (needs an even number of lanes per frame)
bool lin_uses_pairs(int model, int one_focal);

cudaError_t launch_linearize(int model, int one_focal, bool batch, bool cost_only, const LinParams& prm, int n_ctas,
                             cudaStream_t s);
// K2 with the Gram block on the FP64 tensor path (ccrs_linmma.cu): the models with d + 7 >= 14 columns
bool lin_mma_available(int model, int one_focal);
int lin_mma_ctas(int n_sms, int n_frames);   // grid of the tensor-core variant
cudaError_t launch_linearize_mma(int model, int one_focal, bool batch, const LinParams& prm, int n_ctas, cudaStream_t s);
cudaError_t launch_eval_rj(int model, int one_focal, const ProblemDev& pb, const double* intr_dev, const double* poses,
                           int apply_loss, double* r, double* J, int64_t n_obs, cudaStream_t s);
cudaError_t launch_schur(int D, const SchurParams& prm, cudaStream_t s);
cudaError_t launch_backsub(int D, const BacksubParams& prm, cudaStream_t s);
cudaError_t launch_compute_scale(int D, const ProblemDev& pb, int which, double* pose_scale, double* frame_colsq,
                                 cudaStream_t s);
// out[n_seg][NV] = sum over frames of in[v][f], f in [seg_off[s], seg_off[s+1]) in a fixed order.
cudaError_t launch_segreduce(const double* in, int NV, int Fs, const int32_t* seg_off, int n_seg, double* out,
                             cudaStream_t s);
// out[v] = sum_b partials[b][v] (b ascending); optionally published to mapped host memory followed by seq
cudaError_t launch_sum_partials(const double* partials, int n_part, int NV, double* out, volatile double* host_out,
                                double seq, cudaStream_t s);
// stat_out[n_problems][2] = { sum_f frame_md[f] (0 if null), sum_f cost_f } with cost_f taken from
// mode 0: frame_cost[trial]  1: (r,r) entry of blocks[trial]  2: (r,r) of blocks[current]  3: none  4: frame_cost[current]
cudaError_t launch_trial_stats(const ProblemDev& pb, int rr_idx, int mode, const double* frame_md, double* stat_out,
                               cudaStream_t s);
cudaError_t launch_flip_cur(int32_t* cur, const unsigned char* mask_dev, int n_problems, cudaStream_t s);
// fill p[0..n) with the arming bit pattern (self-validating result slots)
cudaError_t launch_arm(double* p, size_t n, cudaStream_t s);
// K6: per-observation reprojection error sqrt(dx^2 + dy^2) without loss (validation, util.rs:733-745)
cudaError_t launch_reproj_err(int model, int one_focal, const ProblemDev& pb, const double* intr_dev, const double* poses,
                              double* err, int64_t n_obs, cudaStream_t s);
// radix select (ccrs_select.cu): exact keys at two ranks of the non-negative doubles v[0..n) plus the sum of the values
// below the second key. st: in {rank[2]} (prefix, below zeroed), out {prefix = key bits, below = #keys smaller};
// hist: [2][kSelBins] zeroed; partial: [n_ctas] per-CTA sums of the values below key 1 (sum them in index order).
constexpr int kSelBins = 2048;
struct SelectState { unsigned long long prefix[2], rank[2], below[2]; };
cudaError_t launch_select(const double* v, int64_t n, SelectState* st, unsigned* hist, double* partial, int n_ctas,
                          cudaStream_t s, int64_t* launches);
// batched initial board poses (ccrs_pnp.cu): one warp per frame; poses_out [n_frames][6], cost_out [n_frames] nullable
cudaError_t launch_pnp(const int32_t* frame_offsets, const double* x, const double* y, const double* z, const double* xn,
                       const double* yn, int n_frames, double* poses_out, double* cost_out, cudaStream_t s);
// x, y, z[k] = board[3 id[k] + 0, 1, 2] (board-format problems)
cudaError_t launch_expand_board(const int32_t* id, const float* board, int n_board, int64_t n, float* x, float* y, float* z,
                                volatile double* bad_flag /* mapped host word, set to 1 on an id outside the table */, cudaStream_t s);
cudaError_t launch_fp64_peak(double* out, int n_ctas, int iters, cudaStream_t s);
cudaError_t launch_l2_flush(double* buf, size_t n, cudaStream_t s);

}  // namespace ccrs

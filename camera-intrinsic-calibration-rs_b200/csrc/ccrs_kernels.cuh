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
constexpr int kObsStages = 4;        // cp.async ring depth of K2's observation prefetch (distance 3 iterations)
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

struct LinParams {
  ProblemDev pb;
  const double* intr_dev;   // [n_problems][D] (batch) — single problem passes intr[] by value below
  double intr[9];           // full vector fx fy cx cy k.. (single problem; constant bank operands)
  const int32_t* acc_to_blk;  // [NACC] sparse accumulator -> dense packed block index
  int which;                // 0 = current point, 1 = trial point
  int G;                    // lanes per frame
  int FPW;                  // frames per warp = 32 / G
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

// number of values K3 reduces per problem: S upper (D(D+1)/2) + g_s (D) + g_a (D) + diag_a (D) + sq_err (1)
inline int nred_of(int D) { return D * (D + 1) / 2 + 3 * D + 1; }

int model_dims(int model, int one_focal, int* D, int* NA, int* NBLK, int* NACC);
void fill_acc_to_blk(int model, int one_focal, int32_t* table);
// K2 runs its lane-pair variant for this model (needs an even number of lanes per frame)
bool lin_uses_pairs(int model, int one_focal);

cudaError_t launch_linearize(int model, int one_focal, bool batch, bool cost_only, const LinParams& prm, int n_ctas,
                             cudaStream_t s);
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
cudaError_t launch_fp64_peak(double* out, int n_ctas, int iters, cudaStream_t s);
cudaError_t launch_l2_flush(double* buf, size_t n, cudaStream_t s);

}  // namespace ccrs

// ccrs_loop.cu — K3 for a single problem (k_schur2) and the device-resident loop controller in its last CTA.
//
//  K3 replaces, per iteration, the numeric sparse LLT of tiny-solver (call sites src/util.rs:455,463,670) by the
//  per-frame elimination of the 6x6 pose block onto the d x d intrinsic system: eight lanes cooperate on a frame (the
//  6x6 Cholesky is computed redundantly, the columns of Y = L^-1 B'^T and X = L^-T Y and the entries of
//  S_f = A_f - Y^T Y are spread over the lanes), sixteen warps x four frames per CTA, fixed-order sums (lanes ->
//  warp -> CTA -> last CTA), no floating-point atomics.
//
//  The Jacobi scaling D_a of the intrinsic columns (tiny-solver LM) factors out of the per-frame work:
//  S' = D_a (A - B D_p C'^-1 D_p B^T) D_a, so it is applied once, after the sum — and the first LM reduction can compute
//  D_a itself from the summed diagonal instead of needing a separate pass + host round trip.
//
//  Device-driven loop (LoopCtl, ccrs_kernels.cuh): the last CTA then runs the controller rule of ccrs_rule.h — the
//  same source the host controller is built from — on the reduced system: accept / reject of the pending trial step
//  (LM), stop tests, damped d x d solve, clamp / fixed variables, and leaves the next linearisation point for K2. One
//  cross-GPU exchange per iteration carries the reduced system AND the trial statistics of the K2 before it.
#include "ccrs_devutil.cuh"
#include "ccrs_rule.h"

#include <type_traits>

namespace ccrs {

constexpr int kS2Threads = 384;
constexpr int kS2Warps = kS2Threads / 32;
constexpr int kS2Lanes = 8;                  // lanes per frame
constexpr int kS2Fpw = 32 / kS2Lanes;        // frames per warp
constexpr int kS2Fpc = kS2Warps * kS2Fpw;    // frames per CTA

template <int D>
struct S2Cfg {
  static constexpr int N = D + 6, NA = N + 1, NB = NA * (NA + 1) / 2;
  static constexpr int NS = D * (D + 1) / 2;
  static constexpr int NRED = NS + 3 * D + 1;
  static constexpr int NX = NRED + 2;                       // + K2's {model decrease, cost} in the exchange
  static constexpr int VPL = (NRED + kS2Lanes - 1) / kS2Lanes;   // reduced values per lane
  static constexpr int NCOL = D + 1;                        // columns of [B'^T | g'_p]
  static constexpr int CPL = (NCOL + kS2Lanes - 1) / kS2Lanes;
  static constexpr int YCOLS = NCOL + 1;                    // + an all-zero column (see c_s2tab)
  static constexpr int WSTRIDE = kS2Fpw * NB + kS2Fpw * YCOLS * 6;   // doubles of shared memory per warp
  static constexpr int NOUT = D * D + 3 * D + 1;
};

CCRS_HD constexpr int s2_smem_doubles(int D) {
  const int NA = D + 7, NB = NA * (NA + 1) / 2, NRED = D * (D + 1) / 2 + 3 * D + 1;
  const int w = kS2Fpw * NB + kS2Fpw * (D + 2) * 6;
  const int tail = kS2Threads + (NRED + 2) + 2 + kRecStride + (int)(sizeof(LoopCtl) / 8) + 2;   // segment sums | totals | record | control block
  const int per_cta = kS2Warps * w + kS2Warps * (NRED + 1);
  return per_cta > tail ? per_cta : tail;
}

// Reduced value v of a frame is  sgn * H[hidx] - dot(Y[:, ya], Y[:, yb])  — one formula for S_f entries, g_s, g_a, the
// diagonal and the cost (the last three use an all-zero column of Y), so the lanes run straight-line code.
// Packed per (D, v): hidx | ya << 8 | yb << 12 | neg << 16. Filled per device on first launch.
constexpr int kS2TabStride = 80;
__constant__ unsigned int c_s2tab[6 * kS2TabStride];   // D = 4 .. 9
static void fill_s2tab(unsigned int* t) {
  for (int D = 4; D <= 9; ++D) {
    const int N = D + 6, NA = N + 1, NS = D * (D + 1) / 2, ZC = D + 1;
    unsigned int* o = t + (D - 4) * kS2TabStride;
    int v = 0;
    for (int a = 0; a < D; ++a) for (int b = a; b < D; ++b) o[v++] = (unsigned)tri_idx(NA, a, b) | (a << 8) | (b << 12);
    for (int a = 0; a < D; ++a) o[v++] = (unsigned)tri_idx(NA, a, N) | (a << 8) | (D << 12) | (1u << 16);     // g_s
    for (int a = 0; a < D; ++a) o[v++] = (unsigned)tri_idx(NA, a, N) | (ZC << 8) | (ZC << 12) | (1u << 16);   // g_a
    for (int a = 0; a < D; ++a) o[v++] = (unsigned)tri_idx(NA, a, a) | (ZC << 8) | (ZC << 12);                // diag
    o[v++] = (unsigned)tri_idx(NA, N, N) | (ZC << 8) | (ZC << 12);                                             // cost
    (void)NS;
  }
}

// (a, b) of packed upper entry e of a D x D symmetric matrix
template <int D>
CCRS_D void tri_decode(int e, int& a, int& b) {
  int row = 0, left = e;
#pragma unroll
  for (int r = 0; r < D; ++r) {
    const int len = D - r;
    if (left >= len && row == r) { left -= len; row = r + 1; }
  }
  a = row; b = row + left;
}

// ---- the controller rule on the device: executed by ONE thread of the last CTA ---------------------------------------
// tot: [NRED] summed reduced values WITHOUT intrinsic scaling ([S upper | g_s | g_a | diag | sq_err]) followed by K2's
// {model decrease (pose part), cost}. rec: shared-memory staging of the record (kRecStride doubles, zero-filled).
template <int D>
__device__ __noinline__ void loop_rule(LoopCtl* ctl, const double* tot, double* rec, int phase_in, double u_use, int which_used) {
  using C = S2Cfg<D>;
  using namespace ccrs_rule;
  const int lm = ctl->mode;
  int status = 0, stop = 0, done = 0, accepted = -1, reduce_valid = 1, decided = 0;
  double rho = 0.0;
  rec[REC_SQ_CUR_BEFORE] = ctl->sq_cur; rec[REC_U_BEFORE] = ctl->u; rec[REC_V_BEFORE] = ctl->v;
  rec[REC_CUR_ERR_BEFORE] = ctl->cur_err; rec[REC_MD_A_BEFORE] = ctl->md_a; rec[REC_LAST_ERR_BEFORE] = ctl->last_err;
  rec[REC_SQ_NEW] = tot[C::NRED + 1]; rec[REC_MD_POSE] = tot[C::NRED];
  rec[REC_HIST_IDX] = -1.0;
  if (lm) {
    if (phase_in == PH_DECIDE) {
      decided = 1;
      LmState st{ctl->u, ctl->v, ctl->cur_err};
      const double last_err = ctl->cur_err;
      const double sq_new = tot[C::NRED + 1];
      accepted = lm_decide(ctl->sq_cur, sq_new, CCRS_RADD(ctl->md_a, tot[C::NRED]), &st, &rho);
      if (accepted) {
        ctl->cur ^= 1;
        for (int i = 0; i < D; ++i) ctl->intr[i] = ctl->trial[i];
        ctl->n_acc++;
      } else {
        ctl->n_rej++;
      }
      ctl->u = st.u; ctl->v = st.v; ctl->cur_err = st.cur_err;
      rec[REC_HIST_IDX] = (double)ctl->it; rec[REC_HIST_VAL] = st.cur_err;
      ctl->final_err = st.cur_err;
      stop = lm_stop(last_err, st.cur_err, rho, accepted, ctl->min_error, ctl->min_abs, ctl->min_rel, &status);
      ctl->it += 1;
      if (stop != 0 || status != 0) { done = 1; if (stop < 0) stop = 0; }
      else if (ctl->it >= ctl->max_iteration) done = 1;
      // the reduction this kernel has just done is the one the controller needs only if it guessed the outcome
      reduce_valid = (which_used == ctl->cur) && (u_use == ctl->u);
    } else if (ctl->it == 0 && ctl->first) {
      ctl->sq_cur = tot[C::NS + 3 * D];
      ctl->cur_err = err_metric(ctl->sq_cur);
    }
  } else {
    // Gauss-Newton: stop tests at the top of the iteration on the error of the point just linearised
    decided = 1;
    const double err = err_metric(block_loss(tot[C::NS + 3 * D], ctl->block_huber));
    if (ctl->it >= ctl->max_iteration) {
      done = 1;   // the step of the last allowed iteration has been applied (and linearised); nothing else to do
    } else {
      rec[REC_HIST_IDX] = (double)ctl->it; rec[REC_HIST_VAL] = err;
      ctl->final_err = err;
      ctl->iterations = ctl->it + 1;
      stop = gn_stop(ctl->it, ctl->last_err, err, ctl->min_error, ctl->min_abs, ctl->min_rel, &status);
      if (stop != 0 || status != 0) { done = 1; if (stop < 0) stop = 0; }
      ctl->last_err = err;
      ctl->cur_err = err;
    }
  }
  int solved = 0;
  if (!done && reduce_valid) {
    double* out = rec + REC_OUT;
    // Jacobi scaling 1/(1+||J[:,c]||) from the first (loss-corrected) Jacobian
    if (lm && ctl->first) {
      for (int a = 0; a < D; ++a) ctl->scale[a] = CCRS_RDIV(1.0, CCRS_RADD(1.0, CCRS_RSQRT(tot[C::NS + 2 * D + a])));
      ctl->first = 0;
    }
    double sa[D];
    for (int a = 0; a < D; ++a) sa[a] = ctl->scale[a];
    int e = 0;
    for (int a = 0; a < D; ++a)
      for (int b = a; b < D; ++b) {
        const double v = CCRS_RMUL(CCRS_RMUL(sa[a], tot[e]), sa[b]);
        out[a * D + b] = v; out[b * D + a] = v; ++e;
      }
    for (int a = 0; a < D; ++a) {
      out[D * D + a] = CCRS_RMUL(sa[a], tot[C::NS + a]);
      out[D * D + D + a] = CCRS_RMUL(sa[a], tot[C::NS + D + a]);
      out[D * D + 2 * D + a] = CCRS_RMUL(CCRS_RMUL(sa[a], tot[C::NS + 2 * D + a]), sa[a]);
    }
    out[D * D + 3 * D] = tot[C::NS + 3 * D];
    if (lm) { ctl->sq_cur = tot[C::NS + 3 * D]; ctl->iterations = ctl->it + 1; }
    double y[D], dx[D], trial[D], md_a = 0.0;
    const Reduced r = view(out, D);
    const int st = solve_intrinsics(r, D, u_use, ctl->min_diag, ctl->max_diag, ctl->has_fixed ? ctl->fixed : nullptr, ctl->fixed_mode, y, &md_a);
    if (st != 0) {
      status = st; done = 1;
    } else {
      for (int i = 0; i < D; ++i) dx[i] = CCRS_RMUL(sa[i], y[i]);
      update_intr(D, ctl->intr, dx, ctl->has_bounds ? ctl->lo : nullptr, ctl->has_bounds ? ctl->hi : nullptr,
                  ctl->has_fixed ? ctl->fixed : nullptr, trial);
      for (int i = 0; i < D; ++i) { ctl->trial[i] = trial[i]; ctl->step[i] = dx[i]; rec[REC_TRIAL + i] = trial[i]; rec[REC_Y + i] = y[i]; }
      ctl->md_a = md_a; ctl->u_used = u_use;
      if (!lm) { for (int i = 0; i < D; ++i) ctl->intr[i] = trial[i]; ctl->it += 1; }   // GN: the step is taken unconditionally
      solved = 1;
    }
    rec[REC_U_SOLVE] = u_use; rec[REC_MD_A] = md_a;
  }
  if (status != 0) ctl->status = status;
  if (done) { ctl->stop_reason = stop; ctl->phase = PH_DONE; }
  else ctl->phase = solved ? PH_TRIAL : PH_REDUCE;
  ctl->seq += 1;
  rec[REC_SEQ] = (double)ctl->seq; rec[REC_PHASE] = (double)ctl->phase; rec[REC_STATUS] = (double)ctl->status;
  rec[REC_STOP] = (double)ctl->stop_reason; rec[REC_IT] = (double)ctl->it; rec[REC_ITERATIONS] = (double)ctl->iterations;
  rec[REC_ACCEPTED] = (double)accepted; rec[REC_CUR] = (double)ctl->cur; rec[REC_RHO] = rho; rec[REC_U] = ctl->u; rec[REC_V] = ctl->v;
  rec[REC_CUR_ERR] = ctl->cur_err; rec[REC_SOLVED] = (double)solved; rec[REC_N_ACC] = (double)ctl->n_acc; rec[REC_N_REJ] = (double)ctl->n_rej;
  rec[REC_FINAL_ERR] = ctl->final_err; rec[REC_SQ_CUR] = ctl->sq_cur; rec[REC_DECIDED] = (double)decided;
  for (int i = 0; i < D; ++i) { rec[REC_INTR + i] = ctl->intr[i]; rec[REC_SCALE + i] = ctl->scale[i]; }
}

template <int D>
__global__ void __launch_bounds__(kS2Threads, 1) k_schur2(const __grid_constant__ Schur2Params prm) {
  using C = S2Cfg<D>;
  constexpr int N = C::N, NA = C::NA, NB = C::NB, NS = C::NS, NRED = C::NRED;
  extern __shared__ double smem[];
  __shared__ int s_last;
  const ProblemDev& pb = prm.pb;
  // the K2 behind this grid (device-driven loop) may be scheduled as soon as every CTA of this grid is resident
  asm volatile("griddepcontrol.launch_dependents;");
  // launched as a programmatic dependent of the K2 in front of it: wait until that grid has completed and flushed
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int fl = lane / kS2Lanes, l = lane % kS2Lanes;
  const int f0w = (blockIdx.x * kS2Warps + wid) * kS2Fpw;   // first frame of this warp
  const bool valid = f0w + fl < pb.n_frames;
  // The per-frame region below contains warp-wide synchronisation, so it is entered by whole warps: the lanes of a frame
  // slot past the end of a ragged last warp recompute the warp's first frame and neither store nor contribute.
  const bool wact = f0w < pb.n_frames;
  const int f = valid ? f0w + fl : f0w;
  LoopCtl* const ctl = prm.ctl;
  const double t_begin = stamp_ns();

  // ---- which blocks, which damping: from the launch (host-driven) or the control block (device-driven loop).
  //      The control block is fetched into shared memory with one coalesced load (one memory round trip); thread 0
  //      evaluates what this slot has to do and broadcasts it.
  int which_buf = pb.cur_val ^ prm.which;   // buffer index of the blocks to reduce
  double u = prm.u_val;
  int phase_in = -1, first = 0, skip = 0, use_pose_scale = prm.use_pose_scale;
  __shared__ double s_head[sizeof(LoopCtl) / 8];   // the control block as it was when this grid started
  __shared__ double s_u;
  __shared__ int s_dec[4];   // which_buf, first, skip, run
  if (ctl) {
    constexpr int kCtlWords = (int)(sizeof(LoopCtl) / 8);
    if (threadIdx.x < kCtlWords) s_head[threadIdx.x] = __ldcg(reinterpret_cast<const double*>(ctl) + threadIdx.x);
    __syncthreads();
    if (threadIdx.x == 0) {
      const LoopCtl* c = reinterpret_cast<const LoopCtl*>(s_head);
      const int ph = c->phase, lm = c->mode, cur = c->cur;
      int run = (ph == PH_REDUCE || ph == PH_DECIDE), wb = cur, sk = 0;
      double uu = 0.0;
      if (run) {
        if (!lm) { sk = c->it >= c->max_iteration; }
        else if (ph == PH_REDUCE) { uu = c->u; }
        else if (prm.px.world > 1) {
          // the decision needs the cross-rank sums: reduce for the outcome that dominates a converging run — accepted
          // with gain ratio >= 0.937, i.e. u_next = u / 3 — and let the last CTA check the guess
          uu = CCRS_RMUL(c->u, ccrs_rule::kLmMinAcceptFactor); wb = cur ^ 1;
        } else {
          // single GPU: K2's sums are final: evaluate the decision rule on them (the last CTA repeats it and commits it)
          ccrs_rule::LmState st{c->u, c->v, c->cur_err};
          const double last_err = st.cur_err;
          double rho;
          const int acc = ccrs_rule::lm_decide(c->sq_cur, c->stat[1], CCRS_RADD(c->md_a, c->stat[0]), &st, &rho);
          int status = 0;
          const int stop = ccrs_rule::lm_stop(last_err, st.cur_err, rho, acc, c->min_error, c->min_abs, c->min_rel, &status);
          sk = stop != 0 || status != 0 || c->it + 1 >= c->max_iteration;
          uu = st.u; wb = acc ? (cur ^ 1) : cur;
        }
      }
      s_u = uu; s_dec[0] = wb; s_dec[1] = lm ? c->first : 0; s_dec[2] = sk; s_dec[3] = run;
    }
    __syncthreads();
    if (!s_dec[3]) return;   // not this slot's turn / loop done
  }
  // The ticket only elects the CTA that will sum the partials (the last one to ARRIVE); it is taken now so that its
  // round trip overlaps the block loads. Partial slots validate themselves, so the elected CTA simply waits for them.
  if (threadIdx.x == 0) s_last = (atomicAdd(prm.ticket, 1u) == gridDim.x - 1);
  if (ctl) {
    const LoopCtl* c = reinterpret_cast<const LoopCtl*>(s_head);
    phase_in = c->phase;
    use_pose_scale = c->mode;
    u = s_u; which_buf = s_dec[0]; first = s_dec[1]; skip = s_dec[2];
  }

  const double t_head = stamp_ns();
  double red[C::VPL];
#pragma unroll
  for (int j = 0; j < C::VPL; ++j) red[j] = 0.0;

  double* const sb = smem + wid * C::WSTRIDE + fl * NB;                       // this frame's packed block
  double* const sy = smem + wid * C::WSTRIDE + kS2Fpw * NB + fl * C::YCOLS * 6;  // its Y columns
  if (wact && !skip) {
    const size_t Fs = pb.Fs;
    const double* blk = pb.blocks[which_buf] + f;
    for (int e = l; e < NB; e += kS2Lanes) cp_async8(sb + e, blk + (size_t)e * Fs);
  }
  cp_async_commit();
  // the stored pose scales (LM after the first reduction) travel with the block
  double sp_ld[6];
#pragma unroll
  for (int i = 0; i < 6; ++i)
    sp_ld[i] = (wact && !skip && use_pose_scale && !first) ? __ldcg(prm.pose_scale + (size_t)i * pb.Fs + f) : 1.0;
  cp_async_wait<0>();
  __syncwarp();
  const double t_load = stamp_ns();
  if (wact && !skip) {
    const size_t Fs = pb.Fs;
    auto H = [&](int i, int j) { return sb[tri_idx(NA, i, j)]; };
    const bool np = prm.no_pose != 0;
    // pose scaling D_p (Jacobi, LM): computed from the first linearisation, then kept
    double sp[6];
    if (use_pose_scale && first) {
      // lane i < 6 of the frame computes (and stores) scale i; the others get it by shuffle
      const int li = l < 6 ? l : 0;
      const double mine = 1.0 / (1.0 + sqrt(sb[tri_idx(NA, D, D) + li * (NA - D) - (li * (li - 1)) / 2]));
      if (valid && l < 6) prm.pose_scale[(size_t)l * Fs + f] = mine;
#pragma unroll
      for (int i = 0; i < 6; ++i) sp[i] = __shfl_sync(0xffffffffu, mine, fl * kS2Lanes + i);
    } else {
#pragma unroll
      for (int i = 0; i < 6; ++i) sp[i] = sp_ld[i];
    }
    // C' (lower, in place Cholesky with reciprocal pivots), damping; every lane of the frame computes it
    double L[6][6], dd[6];
    int bad = 0;
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) L[i][j] = np ? (i == j ? 1.0 : 0.0) : sp[i] * H(D + j, D + i) * sp[j];
      dd[i] = fmin(fmax(L[i][i], prm.min_diag), prm.max_diag);
      L[i][i] = fma(u, dd[i], L[i][i]);
    }
#pragma unroll
    for (int j = 0; j < 6; ++j) {
      double s = L[j][j];
#pragma unroll
      for (int k = 0; k < j; ++k) s -= L[j][k] * L[j][k];
      if (!(s > 0.0)) bad = 1;
      const double il = rsqrt_fast(s > 0.0 ? s : 1.0);
      L[j][j] = il;
#pragma unroll
      for (int i = j + 1; i < 6; ++i) {
        double tt = L[i][j];
#pragma unroll
        for (int k = 0; k < j; ++k) tt -= L[i][k] * L[j][k];
        L[i][j] = tt * il;
      }
    }
    // columns c = l, l + 8, ... of [B'^T | g'_p]: Y = L^-1 rhs (kept for S_f), X = L^-T Y (stored for K4)
    double* el = prm.elim + f;
    if (l < 6) sy[(D + 1) * 6 + l] = 0.0;   // the zero column
#pragma unroll
    for (int q = 0; q < C::CPL; ++q) {
      const int c = l + q * kS2Lanes;
      if (c < C::NCOL) {
        const bool isg = c == D;
        // H(c, D + i) for c < D; H(D + i, N) (one entry per row D + i) for the gradient column
        const double* hb = sb + (isg ? tri_idx(NA, D, N) : tri_idx(NA, 0, D) + c * (NA - 1) - (c * (c - 1)) / 2);
        double yv[6];
        double* gdst = el + (size_t)(6 * D + 6) * Fs;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          // row D + i of the packed upper triangle starts NA - D - i entries after row D + i - 1's start
          const double h = isg ? sb[tri_idx(NA, D + i, N)] : hb[i];
          double s = np ? 0.0 : (isg ? -sp[i] * h : h * sp[i]);
          if (valid && isg) gdst[(size_t)i * Fs] = s;   // g'_p
#pragma unroll
          for (int k = 0; k < i; ++k) s -= L[i][k] * yv[k];
          yv[i] = s * L[i][i];
          sy[c * 6 + i] = yv[i];
        }
        double* xdst = isg ? el + (size_t)(6 * D) * Fs : el + (size_t)c * Fs;
        const size_t xstride = isg ? Fs : (size_t)D * Fs;
#pragma unroll
        for (int i = 5; i >= 0; --i) {
          double s = yv[i];
#pragma unroll
          for (int k = i + 1; k < 6; ++k) s -= L[k][i] * yv[k];
          yv[i] = s * L[i][i];
          if (valid) xdst[(size_t)i * xstride] = yv[i];
        }
      }
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) if (valid && l == i) el[(size_t)(6 * D + 12 + i) * Fs] = dd[i];
    __syncwarp();
    // reduced-system contributions v = l, l + 8, ... of this frame (c_s2tab)
#pragma unroll
    for (int j = 0; j < C::VPL; ++j) {
      const int v = l + j * kS2Lanes;
      if (v < NRED) {
        const unsigned t = c_s2tab[(D - 4) * kS2TabStride + v];
        const double h = sb[t & 0xffu];
        const double* ya = sy + ((t >> 8) & 0xfu) * 6;
        const double* yb = sy + ((t >> 12) & 0xfu) * 6;
        double val = (t >> 16) ? -h : h;
#pragma unroll
        for (int i = 0; i < 6; ++i) val -= ya[i] * yb[i];
        // a failed factorisation poisons the reduced system so the controller sees it (tiny-solver returns None); the
        // cost (last entry) stays valid: the failure is the factorisation's, not a NaN error
        if (bad && v != NRED - 1) val = nan("");
        red[j] = valid ? val : 0.0;
      }
    }
  }
  // ---- fixed-order sums: the warp's four frames (shuffles), the CTA's warps, then the CTAs (last CTA) ----
  const double t_comp = stamp_ns();
  double* const s_cta = smem + kS2Warps * C::WSTRIDE;     // [kS2Warps][NRED + 1]
#pragma unroll
  for (int j = 0; j < C::VPL; ++j) {
    const double x1 = __shfl_down_sync(0xffffffffu, red[j], 8), x2 = __shfl_down_sync(0xffffffffu, red[j], 16),
                 x3 = __shfl_down_sync(0xffffffffu, red[j], 24);
    const int v = l + j * kS2Lanes;
    if (fl == 0 && v < NRED) s_cta[wid * (NRED + 1) + v] = ((red[j] + x1) + x2) + x3;
  }
  __syncthreads();
  // No fence: the partial slots validate themselves (armed with kArmBits at creation and re-armed by the summing CTA of
  // the previous launch); the relaxed ticket (taken at the top) only elected the CTA that sums.
  if (threadIdx.x < NRED) {
    double s = s_cta[threadIdx.x];
#pragma unroll
    for (int w = 1; w < kS2Warps; ++w) s += s_cta[w * (NRED + 1) + threadIdx.x];
    prm.partials[(size_t)blockIdx.x * NRED + threadIdx.x] = s;
  }
  __syncthreads();
  if (!s_last) return;
  const double t_tail = stamp_ns();

  // ---- last CTA: sum the CTA partials in a fixed order (SEG interleaved segments per value, 16 loads in flight);
  //      a slot that still holds the arming pattern has not landed yet: read it again ----
  double* const s_seg = smem;                          // [SEG][NRED]
  double* const s_tot = smem + kS2Threads;             // [NX]
  double* const s_rec = s_tot + C::NX + 2;             // [kRecStride]
  constexpr int SEG = kS2Threads / NRED;
  const int nb = gridDim.x;
  {
    const int v = threadIdx.x % NRED, h = threadIdx.x / NRED;
    if (h < SEG) {
      double a = 0.0;
      constexpr int kFly = 20;   // loads in flight per thread (one round for 7,000 frames: 146 CTAs <= 9 x 20)
      bool poison = false;
      for (int b0 = h; b0 < nb; b0 += kFly * SEG) {
        double t[kFly];
        bool ok;
        const long long t_spin = clock64();
        do {
          ok = true;
#pragma unroll
          for (int q = 0; q < kFly; ++q) {
            const int b = b0 + q * SEG;
            t[q] = b < nb ? ld_spin(prm.partials + (size_t)b * NRED + v) : 0.0;
            ok = ok && (__double_as_longlong(t[q]) != kArmBits);
          }
          if (!ok && clock64() - t_spin > 4000000000LL) { poison = true; ok = true; }   // ~2 s: a partial never arrived -> poison, not a hang
        } while (!ok);
#pragma unroll
        for (int q = 0; q < kFly; ++q) {
          a += t[q];
          const int b = b0 + q * SEG;
          if (b < nb) prm.partials[(size_t)b * NRED + v] = __longlong_as_double(kArmBits);
        }
      }
      if (poison) a = nan("");
      s_seg[h * NRED + v] = a;
    }
  }
  if (threadIdx.x < kRecStride) s_rec[threadIdx.x] = 0.0;
  // the control block is worked on in shared memory (one round trip in, one out, instead of dependent global accesses
  // by the one thread that runs the rule)
  LoopCtl* const s_ctl = reinterpret_cast<LoopCtl*>(s_rec + kRecStride);
  constexpr int kCtlWords = (int)(sizeof(LoopCtl) / 8);
  if (ctl && threadIdx.x < kCtlWords) reinterpret_cast<double*>(s_ctl)[threadIdx.x] = s_head[threadIdx.x];   // unchanged since the head read it
  __syncthreads();
  if (threadIdx.x < C::NX) {
    double tot;
    if (threadIdx.x < NRED) {
      tot = s_seg[threadIdx.x];
#pragma unroll
      for (int h = 1; h < SEG; ++h) tot += s_seg[h * NRED + threadIdx.x];
    } else {
      tot = ctl ? s_ctl->stat[threadIdx.x - NRED] : 0.0;
    }
    if (prm.px.world > 1 && (ctl || threadIdx.x < NRED)) {
      PeerXchg px = prm.px;
      if (prm.xchg_count) {   // device-driven loop: slot parity from the device-side count of executed exchanges
        const unsigned parity = __ldcg(prm.xchg_count) & 1u;
        px.off += (int)(parity * kXchgMaxRanks * kXchgMaxVals);
      }
      tot = peer_exchange(px, threadIdx.x, tot);
    }
    s_tot[threadIdx.x] = tot;
  }
  __syncthreads();
  const double t_summed = stamp_ns();
  if (threadIdx.x == 0) {
    *prm.ticket = 0u;
    if (prm.xchg_count && prm.px.world > 1) *prm.xchg_count += 1u;
  }
  if (!ctl) {
    // host-driven call: apply the intrinsic scaling and publish the packed reduced system
    if (threadIdx.x < NRED) {
      const int v = threadIdx.x;
      double val = s_tot[v];
      if (prm.intr_scale) {
        if (v < NS) { int a, b; tri_decode<D>(v, a, b); val = CCRS_RMUL(CCRS_RMUL(prm.intr_scale[a], val), prm.intr_scale[b]); }
        else if (v < NS + 2 * D) { val = CCRS_RMUL(prm.intr_scale[(v - NS) % D], val); }
        else if (v < NS + 3 * D) { const double s = prm.intr_scale[v - NS - 2 * D]; val = CCRS_RMUL(CCRS_RMUL(s, val), s); }
      }
      prm.red_out[v] = val;
      if (prm.host_red) prm.host_red[v] = val;   // sentinel protocol: no fence
    }
    return;
  }
  if (threadIdx.x == 0) {
    s_rec[REC_T_K2_BEGIN] = s_ctl->t_k2_begin; s_rec[REC_T_K2_END] = s_ctl->t_k2_end; s_rec[REC_T_K2_WAKE] = s_ctl->t_k2_wake;
    s_rec[REC_T_K3_BEGIN] = t_begin; s_rec[REC_T_TAIL] = t_tail;
    s_rec[REC_T_HEAD] = t_head; s_rec[REC_T_LOAD] = t_load; s_rec[REC_T_COMP] = t_comp; s_rec[REC_T_SUMMED] = t_summed;
    s_rec[REC_T_RULE] = stamp_ns();
    loop_rule<D>(s_ctl, s_tot, s_rec, phase_in, u, which_buf);
    s_rec[REC_T_END] = stamp_ns();
  }
  __syncthreads();
  if (threadIdx.x < kCtlWords) reinterpret_cast<double*>(ctl)[threadIdx.x] = reinterpret_cast<const double*>(s_ctl)[threadIdx.x];
  if (threadIdx.x < kRecStride) {
    const int slot = ((int)s_rec[REC_SEQ] - 1) % kRecSlots;
    prm.rec[(size_t)slot * kRecStride + threadIdx.x] = s_rec[threadIdx.x];   // sentinel protocol: no fence
  }
}

// ---- batch controller kernels: one CTA per problem -------------------------------------------------------------------
constexpr int kBatchThreads = 256;

// status of the iteration: active problems counted with an integer atomic (order-free), published by the last CTA
CCRS_D void batch_publish(const BatchRuleParams& prm, int active_now) {
  __shared__ int s_is_last;
  if (threadIdx.x == 0) {
    if (active_now) atomicAdd(prm.n_active, 1u);
    __threadfence();
    s_is_last = (atomicAdd(prm.ticket, 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_is_last && threadIdx.x == 0) {
    __threadfence();
    const unsigned n = atomicExch(prm.n_active, 0u);
    *prm.ticket = 0u;
    volatile double* slot = prm.host_status + (size_t)((prm.seq - 1) % kRecSlots) * 2;
    slot[1] = (double)n;
    slot[0] = (double)prm.seq;   // sentinel protocol: both words are waited for
  }
}

template <int D>
__global__ void __launch_bounds__(kBatchThreads) k_batch_solve(const __grid_constant__ BatchRuleParams prm) {
  using namespace ccrs_rule;
  constexpr int NS = D * (D + 1) / 2, NRED = NS + 3 * D + 1, NOUT = D * D + 3 * D + 1;
  __shared__ double s_red[NRED];
  const int q = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  BatchCtl* bc = prm.ctl + q;
  int active = bc->active;
  if (active) {
    // per-problem sum of the frame contributions: one warp per value, lane-strided partial sums in frame order, then a
    // fixed butterfly (the order k_segreduce uses on the host-driven path: identical bits)
    const int b = prm.pb.problem_frame_offsets[q], e = prm.pb.problem_frame_offsets[q + 1];
    for (int v = warp; v < NRED; v += kBatchThreads / 32) {
      const double* src = prm.frame_red + (size_t)v * prm.pb.Fs;
      double s = 0.0;
      for (int f = b + lane; f < e; f += 32) s += src[f];
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) s_red[v] = s;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      double out[NOUT];
      int e2 = 0;
      for (int a = 0; a < D; ++a)
        for (int c = a; c < D; ++c) { out[a * D + c] = s_red[e2]; out[c * D + a] = s_red[e2]; ++e2; }
      for (int i = 0; i < 3 * D + 1; ++i) out[D * D + i] = s_red[NS + i];
      const Reduced r = view(out, D);
      const unsigned char* fixed = prm.has_fixed ? prm.fixed : nullptr;
      const double* lo = prm.has_bounds ? prm.lo : nullptr;
      const double* hi = prm.has_bounds ? prm.hi : nullptr;
      double y[D], next[D];
      int status = 0;
      if (prm.lm) {
        if (bc->it == 0) bc->cur_err = err_metric(r.sq_err);
        bc->sq_cur = r.sq_err;
        bc->iterations = bc->it + 1;
        double md_a = 0.0;
        status = solve_intrinsics(r, D, bc->u, prm.min_diag, prm.max_diag, fixed, prm.fixed_mode, y, &md_a);
        if (status == 0) {
          double dx[D];
          for (int i = 0; i < D; ++i) dx[i] = CCRS_RMUL(bc->scale[i], y[i]);
          update_intr(D, bc->intr, dx, lo, hi, fixed, next);
          bc->md_a = md_a;
          for (int i = 0; i < D; ++i) { bc->trial[i] = next[i]; prm.intr_dev[(size_t)q * D + i] = next[i]; prm.ya_dev[(size_t)q * D + i] = y[i]; }
          prm.u_dev[q] = bc->u;
        }
      } else if (bc->it >= prm.max_iteration) {
        active = 0;   // the step of the last allowed iteration has been applied (and linearised): nothing else to do
      } else {
        const double err = err_metric(r.sq_err);
        bc->iterations = bc->it + 1;
        bc->final_err = err;
        if (q == 0 && prm.err_hist0) prm.err_hist0[bc->it] = err;
        const int why = gn_stop(bc->it, bc->last_err, err, prm.min_error, prm.min_abs, prm.min_rel, &status);
        if (status == 0 && why != 0) { bc->stop = why; active = 0; }
        if (status == 0 && active) {
          bc->last_err = err;
          status = solve_intrinsics(r, D, 0.0, prm.min_diag, prm.max_diag, fixed, prm.fixed_mode, y, nullptr);
          if (status == 0) {
            update_intr(D, bc->intr, y, lo, hi, fixed, next);
            for (int i = 0; i < D; ++i) { bc->intr[i] = next[i]; prm.intr_dev[(size_t)q * D + i] = next[i]; prm.ya_dev[(size_t)q * D + i] = y[i]; }
            bc->it += 1;
          }
        }
      }
      if (status != 0) { bc->status = status; active = 0; }
      bc->active = active;
      prm.active[q] = (unsigned char)active;   // K2 / K3 skip the frames of a problem that has stopped
      s_red[0] = (double)active;
    }
    __syncthreads();
    active = (int)s_red[0];
  }
  if (!prm.lm) batch_publish(prm, active);
}

// LM: accept / reject of the trial point every active problem has just been linearised at
__global__ void __launch_bounds__(kBatchThreads) k_batch_decide(const __grid_constant__ BatchRuleParams prm) {
  using namespace ccrs_rule;
  __shared__ double sh0[kBatchThreads], sh1[kBatchThreads];
  const int q = blockIdx.x, D = prm.D;
  BatchCtl* bc = prm.ctl + q;
  int active = bc->active;
  if (active) {
    const int b = prm.pb.problem_frame_offsets[q], e = prm.pb.problem_frame_offsets[q + 1];
    const int cur = prm.cur[q];
    const double* cost = prm.pb.blocks[cur ^ 1] + (size_t)prm.rr_idx * prm.pb.Fs;
    double s0 = 0.0, s1 = 0.0;   // same order as k_trial_stats (thread-strided, then a fixed tree)
    for (int f = b + threadIdx.x; f < e; f += kBatchThreads) { s0 += prm.frame_md[f]; s1 += cost[f]; }
    sh0[threadIdx.x] = s0; sh1[threadIdx.x] = s1;
    __syncthreads();
    for (int w = kBatchThreads / 2; w > 0; w >>= 1) {
      if (threadIdx.x < w) { sh0[threadIdx.x] += sh0[threadIdx.x + w]; sh1[threadIdx.x] += sh1[threadIdx.x + w]; }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      LmState st{bc->u, bc->v, bc->cur_err};
      const double last_err = bc->cur_err;
      double rho;
      const int acc = lm_decide(bc->sq_cur, sh1[0], CCRS_RADD(bc->md_a, sh0[0]), &st, &rho);
      bc->u = st.u; bc->v = st.v; bc->cur_err = st.cur_err; bc->final_err = st.cur_err;
      if (acc) {
        for (int i = 0; i < D; ++i) bc->intr[i] = bc->trial[i];
        prm.cur[q] = cur ^ 1;
        bc->n_acc++;
      } else {
        bc->n_rej++;
      }
      if (q == 0 && prm.err_hist0) prm.err_hist0[bc->it] = st.cur_err;
      int status = 0;
      const int why = lm_stop(last_err, st.cur_err, rho, acc, prm.min_error, prm.min_abs, prm.min_rel, &status);
      if (status != 0) { bc->status = status; active = 0; }
      else if (why != 0) { bc->stop = why; active = 0; }
      bc->it += 1;
      if (bc->it >= prm.max_iteration) active = 0;
      bc->active = active;
      prm.active[q] = (unsigned char)active;
      prm.u_dev[q] = bc->u;          // damping of the next reduction
      sh0[0] = (double)active;
    }
    __syncthreads();
    active = (int)sh0[0];
  }
  batch_publish(prm, active);
}

cudaError_t launch_batch_solve(const BatchRuleParams& prm, cudaStream_t s) {
  switch (prm.D) {
    case 4: k_batch_solve<4><<<prm.pb.n_problems, kBatchThreads, 0, s>>>(prm); break;
    case 5: k_batch_solve<5><<<prm.pb.n_problems, kBatchThreads, 0, s>>>(prm); break;
    case 6: k_batch_solve<6><<<prm.pb.n_problems, kBatchThreads, 0, s>>>(prm); break;
    case 7: k_batch_solve<7><<<prm.pb.n_problems, kBatchThreads, 0, s>>>(prm); break;
    case 8: k_batch_solve<8><<<prm.pb.n_problems, kBatchThreads, 0, s>>>(prm); break;
    case 9: k_batch_solve<9><<<prm.pb.n_problems, kBatchThreads, 0, s>>>(prm); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}
cudaError_t launch_batch_decide(const BatchRuleParams& prm, cudaStream_t s) {
  k_batch_decide<<<prm.pb.n_problems, kBatchThreads, 0, s>>>(prm);
  return cudaGetLastError();
}

template <class F>
static cudaError_t dispatch_d2(int D, F&& f) {
  switch (D) {
    case 4: return f(std::integral_constant<int, 4>{});
    case 5: return f(std::integral_constant<int, 5>{});
    case 6: return f(std::integral_constant<int, 6>{});
    case 7: return f(std::integral_constant<int, 7>{});
    case 8: return f(std::integral_constant<int, 8>{});
    case 9: return f(std::integral_constant<int, 9>{});
  }
  return cudaErrorInvalidValue;
}

int schur2_ctas(int n_frames) { return (n_frames + kS2Fpc - 1) / kS2Fpc; }

cudaError_t launch_schur2(int D, const Schur2Params& prm, int n_frames, bool pdl, cudaStream_t s) {
  const int nb = schur2_ctas(n_frames);
  return dispatch_d2(D, [&](auto DD) {
    constexpr int d = decltype(DD)::value;
    constexpr size_t smem = (size_t)s2_smem_doubles(d) * sizeof(double);
    auto kern = k_schur2<d>;
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
      configured = true;
    }
    {
      static bool tab_done[64] = {false};
      int dev = 0;
      cudaGetDevice(&dev);
      if (!tab_done[dev & 63]) {
        unsigned int tab[6 * kS2TabStride] = {0};
        fill_s2tab(tab);
        cudaError_t e = cudaMemcpyToSymbol(c_s2tab, tab, sizeof(tab));
        if (e != cudaSuccess) return e;
        tab_done[dev & 63] = true;
      }
    }
    // programmatic dependent launch: enqueued behind a K2 that is still running, the CTAs are scheduled as K2 drains and
    // wait at griddepcontrol.wait; with nothing in front of it the wait returns at once
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(nb); cfg.blockDim = dim3(kS2Threads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = pdl ? 1 : 0;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, prm);
  });
}

}  // namespace ccrs

// ccrs_controller.cpp — host-side loop controllers, backend-agnostic (plain C++, no CUDA).
//
// Replaces tiny_solver::GaussNewtonOptimizer::optimize (the loop the reference runs: src/util.rs:443-464,
// :668-670) and tiny-solver's LevenbergMarquardtOptimizer::optimize (named by north_star),
// keeping on the host exactly what north_star keeps there: damping and accept/reject control, the small
// (d x d) intrinsic solve, bounds clamp and fixed-variable reset (src/util.rs:29-71; tiny-solver
// ParameterBlock::update_params). Everything per-observation / per-frame happens behind ccrs_backend.
#include "../../include/ccrs_b200.h"
#include "ccrs_rule.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

namespace {

// The controller arithmetic (error metric, damped d x d solve, update_params, accept / reject rule, stop tests) lives in
// ccrs_rule.h: one source, built here for the host and in ccrs_loop.cu for the device-driven loop.
using namespace ccrs_rule;
static_assert(kErrNumeric == CCRS_ERR_NUMERIC && kErrCholesky == CCRS_ERR_CHOLESKY, "ccrs_rule.h status codes out of sync with the ABI");

#define BE(call)                         \
  do {                                   \
    int _st = (call);                    \
    if (_st != 0) { sum->status = _st; return _st; } \
  } while (0)

}  // namespace

extern "C" {

void ccrs_default_options(ccrs_options* o) {
  o->max_iteration = 100;
  o->min_abs_decrease = 1e-5;
  o->min_rel_decrease = 1e-5;
  o->min_error = 1e-10;
  o->lm_initial_radius = 1e4;
  o->lm_min_diag = 1e-6;
  o->lm_max_diag = 1e32;
  o->fixed_mode = 0;
  o->speculative = 1;
  o->verbose = 0;
  o->block_huber_delta = 0.0;
}

int ccrs_controller_gn(const ccrs_backend* be, double* intr, const double* lo, const double* hi,
                       const unsigned char* fixed, const ccrs_options* opt_in, ccrs_summary* sum, double* err_hist) {
  ccrs_options opt;
  if (opt_in) opt = *opt_in; else ccrs_default_options(&opt);
  ccrs_summary local_sum;
  if (!sum) sum = &local_sum;
  std::memset(sum, 0, sizeof(*sum));
  const int d = be->d, P = be->n_problems, NOUT = d * d + 3 * d + 1;
  std::vector<double> out((size_t)P * NOUT), y((size_t)P * d, 0.0), last_err(P, 0.0);
  std::vector<unsigned char> active(P, 1);
  std::vector<int> stop(P, 0);
  int worst = 0;
  for (int it = 0; it < opt.max_iteration; ++it) {
    BE(be->linearize(be->ctx, intr, 0));
    BE(be->reduce(be->ctx, 0, nullptr, 0, opt.lm_min_diag, opt.lm_max_diag, out.data()));
    if (be->allreduce) BE(be->allreduce(be->ctx, out.data(), (int)out.size()));
    sum->iterations = it + 1;
    int n_active = 0;
    for (int p = 0; p < P; ++p) {
      if (!active[p]) continue;
      const Reduced r = view(&out[(size_t)p * NOUT], d);
      const double err = err_metric(block_loss(r.sq_err, opt.block_huber_delta));
      if (p == 0) { if (err_hist) err_hist[it] = err; sum->final_error = err; }
      int status = 0;
      const int why = gn_stop(it, last_err[p], err, opt.min_error, opt.min_abs_decrease, opt.min_rel_decrease, &status);
      if (status != 0) { active[p] = 0; worst = status; continue; }
      if (why != 0) { active[p] = 0; stop[p] = why; continue; }
      last_err[p] = err;
      double* yp = &y[(size_t)p * d];
      const int st = solve_intrinsics(r, d, 0.0, opt.lm_min_diag, opt.lm_max_diag, fixed, opt.fixed_mode, yp, nullptr);
      if (st != 0) { active[p] = 0; worst = st; continue; }
      update_intr(d, intr + (size_t)p * d, yp, lo, hi, fixed, intr + (size_t)p * d);
      ++n_active;
    }
    if (n_active == 0) break;
    BE(be->backsub(be->ctx, y.data(), nullptr, P > 1 ? active.data() : nullptr, 1));
  }
  sum->stop_reason = stop[0];
  sum->status = worst;
  return worst;
}

// LM loop state, steppable so that the bench can time single iterations (ccrs_bench_lm_steps).
struct ccrs_lm_state {
  const ccrs_backend* be;
  ccrs_options opt;
  int d, P, NOUT, it;
  double* intr;
  const double *lo, *hi;
  const unsigned char* fixed;
  std::vector<double> out, y, scale, colsq, u, v, cur_err, md_a, trial, dx, stats;
  std::vector<unsigned char> active, acc_mask;
  std::vector<int> stop;
  int worst;
  ccrs_summary* sum;
  double* err_hist;
};

static int lm_begin(ccrs_lm_state& S, const ccrs_backend* be, double* intr, const double* lo, const double* hi,
                    const unsigned char* fixed, const ccrs_options* opt_in, ccrs_summary* sum, double* err_hist) {
  S.be = be;
  if (opt_in) S.opt = *opt_in; else ccrs_default_options(&S.opt);
  S.sum = sum;
  std::memset(sum, 0, sizeof(*sum));
  S.err_hist = err_hist;
  S.intr = intr; S.lo = lo; S.hi = hi; S.fixed = fixed;
  const int d = S.d = be->d, P = S.P = be->n_problems;
  S.NOUT = d * d + 3 * d + 1;
  S.it = 0; S.worst = 0;
  S.out.assign((size_t)P * S.NOUT, 0.0); S.y.assign((size_t)P * d, 0.0); S.scale.assign((size_t)P * d, 1.0);
  S.colsq.assign((size_t)P * d, 0.0);
  S.u.assign(P, 1.0 / S.opt.lm_initial_radius); S.v.assign(P, kLmRejectFactor0);
  S.cur_err.assign(P, 0.0); S.md_a.assign(P, 0.0);
  S.trial.assign((size_t)P * d, 0.0); S.dx.assign(d, 0.0); S.stats.assign((size_t)P * 2, 0.0);
  S.active.assign(P, 1); S.acc_mask.assign(P, 0); S.stop.assign(P, 0);
  BE(be->linearize(be->ctx, intr, 0));
  // Jacobi scaling 1/(1+||J[:,c]||) from the first (loss-corrected) Jacobian
  BE(be->compute_scale(be->ctx, 0, S.colsq.data()));
  if (be->allreduce) BE(be->allreduce(be->ctx, S.colsq.data(), (int)S.colsq.size()));
  for (size_t i = 0; i < S.scale.size(); ++i) S.scale[i] = 1.0 / (1.0 + std::sqrt(S.colsq[i]));
  BE(be->set_intr_scale(be->ctx, S.scale.data()));
  return 0;
}

// one LM iteration; *done = 1 when every problem has stopped
static int lm_iterate(ccrs_lm_state& S, int* done) {
  const ccrs_backend* be = S.be;
  const ccrs_options& opt = S.opt;
  ccrs_summary* sum = S.sum;
  const int d = S.d, P = S.P, NOUT = S.NOUT, it = S.it;
  *done = 0;
  BE(be->reduce(be->ctx, 0, S.u.data(), 1, opt.lm_min_diag, opt.lm_max_diag, S.out.data()));
  if (be->allreduce) BE(be->allreduce(be->ctx, S.out.data(), (int)S.out.size()));
  sum->iterations = it + 1;
  int n_active = 0;
  for (int p = 0; p < P; ++p) {
    std::memcpy(&S.trial[(size_t)p * d], S.intr + (size_t)p * d, d * sizeof(double));
    if (!S.active[p]) continue;
    const Reduced r = view(&S.out[(size_t)p * NOUT], d);
    if (it == 0) S.cur_err[p] = err_metric(r.sq_err);
    double* yp = &S.y[(size_t)p * d];
    const int st = solve_intrinsics(r, d, S.u[p], opt.lm_min_diag, opt.lm_max_diag, S.fixed, opt.fixed_mode, yp, &S.md_a[p]);
    if (st != 0) { S.active[p] = 0; S.worst = st; continue; }
    for (int i = 0; i < d; ++i) S.dx[i] = S.scale[(size_t)p * d + i] * yp[i];
    update_intr(d, S.intr + (size_t)p * d, S.dx.data(), S.lo, S.hi, S.fixed, &S.trial[(size_t)p * d]);
    ++n_active;
  }
  if (n_active == 0) { *done = 1; return 0; }
  BE(be->backsub(be->ctx, S.y.data(), S.u.data(), P > 1 ? S.active.data() : nullptr, 0));
  BE(be->trial_stats(be->ctx, S.trial.data(), opt.speculative, S.stats.data()));
  if (be->allreduce) BE(be->allreduce(be->ctx, S.stats.data(), (int)S.stats.size()));
  bool any_accept = false;
  for (int p = 0; p < P; ++p) {
    S.acc_mask[p] = 0;
    if (!S.active[p]) continue;
    const Reduced r = view(&S.out[(size_t)p * NOUT], d);
    const double new_sq = S.stats[2 * p + 1];
    const double last_err = S.cur_err[p];
    LmState st{S.u[p], S.v[p], S.cur_err[p]};
    double rho;
    const bool accepted = lm_decide(r.sq_err, new_sq, S.md_a[p] + S.stats[2 * p], &st, &rho) != 0;
    S.u[p] = st.u; S.v[p] = st.v; S.cur_err[p] = st.cur_err;
    if (accepted) {
      S.acc_mask[p] = 1; any_accept = true;
      std::memcpy(S.intr + (size_t)p * d, &S.trial[(size_t)p * d], d * sizeof(double));
      if (p == 0) sum->n_accepted++;
    } else if (p == 0) {
      sum->n_rejected++;
    }
    if (p == 0) { if (S.err_hist) S.err_hist[it] = S.cur_err[p]; sum->final_error = S.cur_err[p]; }
    int status = 0;   // the stop tests compare successive ACCEPTED errors
    const int why = lm_stop(last_err, S.cur_err[p], rho, accepted ? 1 : 0, opt.min_error, opt.min_abs_decrease, opt.min_rel_decrease, &status);
    if (status != 0) { S.active[p] = 0; S.worst = status; }
    else if (why != 0) { S.active[p] = 0; S.stop[p] = why; }
  }
  if (any_accept) {
    BE(be->accept(be->ctx, S.acc_mask.data()));
    if (!opt.speculative) BE(be->linearize(be->ctx, S.intr, 0));
  }
  bool any_active = false;
  for (int p = 0; p < P; ++p) any_active |= (S.active[p] != 0);
  S.it = it + 1;
  if (!any_active || S.it >= opt.max_iteration) *done = 1;
  return 0;
}

int ccrs_controller_lm(const ccrs_backend* be, double* intr, const double* lo, const double* hi,
                       const unsigned char* fixed, const ccrs_options* opt_in, ccrs_summary* sum, double* err_hist) {
  ccrs_summary local_sum;
  if (!sum) sum = &local_sum;
  ccrs_lm_state S;
  int st = lm_begin(S, be, intr, lo, hi, fixed, opt_in, sum, err_hist);
  if (st) return st;
  int done = S.opt.max_iteration <= 0;
  while (!done) {
    st = lm_iterate(S, &done);
    if (st) return st;
  }
  sum->stop_reason = S.stop[0];
  sum->status = S.worst;
  return S.worst;
}

// Stepping interface for measurement (ccrs_bench_lm_steps): opaque state, one LM iteration per call.
ccrs_lm_state* ccrs_lm_state_create(const ccrs_backend* be, double* intr, const ccrs_options* opt, ccrs_summary* sum) {
  ccrs_lm_state* S = new ccrs_lm_state();
  if (lm_begin(*S, be, intr, nullptr, nullptr, nullptr, opt, sum, nullptr) != 0) { delete S; return nullptr; }
  return S;
}
int ccrs_lm_state_step(ccrs_lm_state* S, int* done) { return lm_iterate(*S, done); }
void ccrs_lm_state_destroy(ccrs_lm_state* S) { delete S; }

}  // extern "C"

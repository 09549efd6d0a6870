// ccrs_rule.h — the loop-control arithmetic of the Gauss-Newton / Levenberg-Marquardt controllers as ONE piece of
// source compiled for the host (ccrs_controller.cpp; the host-side audit of the device-driven loop in ccrs_api.cu)
// and for the device (the tail of K3, ccrs_loop.cu, which runs the same rule in-line so that no host round trip sits
// between the reduction and the next linearisation).
//
// What it restates: tiny-solver's GaussNewtonOptimizer / LevenbergMarquardtOptimizer loop control (call sites
// src/util.rs:443-464, :668-670; constants SURVEY App. B, mirrored in oracle/ccrs_oracle.hpp) and
// ParameterBlock::update_params (bounds clamp, fixed-variable reset; src/util.rs:29-71).
//
// Every floating-point operation is a single correctly-rounded IEEE operation on both sides (no FMA contraction on the
// device: __dmul_rn & co.; x86-64 host code has no fused operations), so the host audit reproduces the device's
// decisions BIT FOR BIT from the published reduced system.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define CCRS_RULE_HD __host__ __device__ __forceinline__
#else
#define CCRS_RULE_HD inline
#endif

#if defined(__CUDACC__)
#define CCRS_RULE_UNROLL _Pragma("unroll")
#else
#define CCRS_RULE_UNROLL
#endif

#if defined(__CUDA_ARCH__)
#define CCRS_RMUL(a, b) __dmul_rn((a), (b))
#define CCRS_RADD(a, b) __dadd_rn((a), (b))
#define CCRS_RSUB(a, b) __dsub_rn((a), (b))
#define CCRS_RDIV(a, b) __ddiv_rn((a), (b))
#define CCRS_RSQRT(a) __dsqrt_rn(a)
#define CCRS_RFMA(a, b, c) __fma_rn((a), (b), (c))
#else
#define CCRS_RMUL(a, b) ((a) * (b))
#define CCRS_RADD(a, b) ((a) + (b))
#define CCRS_RSUB(a, b) ((a) - (b))
#define CCRS_RDIV(a, b) ((a) / (b))
#define CCRS_RSQRT(a) sqrt(a)
#define CCRS_RFMA(a, b, c) fma((a), (b), (c))   /* correctly rounded by definition, with or without FMA hardware */
#endif

namespace ccrs_rule {

constexpr int kMaxD = 9;                  // intrinsics of one camera (OPENCV5)
// Restated tiny-solver constants (unverifiable here — see SURVEY App. B; mirrored in oracle/ccrs_oracle.hpp)
constexpr bool kErrorIsL2Norm = true;     // optimisers compare ||r||, not ||r||^2
constexpr double kLmRejectFactor0 = 2.0;
constexpr double kLmMinAcceptFactor = 1.0 / 3.0;   // accept: u *= max(1/3, 1 - (2 rho - 1)^3)

// status codes of include/ccrs_b200.h (kept in sync by a static_assert in ccrs_controller.cpp)
constexpr int kErrNumeric = -4, kErrCholesky = -5;

CCRS_RULE_HD bool is_nan(double x) { return x != x; }
CCRS_RULE_HD double dmin(double a, double b) { return a < b ? a : b; }
CCRS_RULE_HD double dmax(double a, double b) { return a > b ? a : b; }
CCRS_RULE_HD double dabs(double a) { return a < 0.0 ? -a : a; }

CCRS_RULE_HD double err_metric(double sq) { return kErrorIsL2Norm ? CCRS_RSQRT(sq) : sq; }
// One HuberLoss over the WHOLE residual vector (ModelConvertFactor is a single residual block, util.rs:246-251):
// the corrector scales r and J by the same w = sqrt(rho'(s)), s = ||r||^2, so the Gauss-Newton step is unchanged and
// only the error the stop tests see becomes ||w r||^2 = delta sqrt(s) outside the quadratic region.
CCRS_RULE_HD double block_loss(double sq, double delta) {
  return (delta > 0.0 && sq > CCRS_RMUL(delta, delta)) ? CCRS_RMUL(delta, CCRS_RSQRT(sq)) : sq;
}

// 1/sqrt(x) for a normal positive x, identical bits on host and device: integer seed (relative error < 3.5e-2), then
// four Newton steps y <- y (1 + e/2), e = 1 - x y^2, in correctly-rounded multiply / fused multiply-add arithmetic
// (error after the steps ~1e-20 before rounding: within ~1 ulp). A third of the latency of an IEEE sqrt followed by an
// IEEE division, which is what the device-side rule — one thread, one dependent chain — waits for.
CCRS_RULE_HD double rule_rsqrt(double x) {
  union { double d; long long i; } v;
  v.d = x;
  v.i = 0x5fe6eb50c7b537a9LL - (v.i >> 1);
  double y = v.d;
  CCRS_RULE_UNROLL
  for (int it = 0; it < 4; ++it) {
    const double t = CCRS_RMUL(x, y);
    const double e = CCRS_RFMA(-t, y, 1.0);
    y = CCRS_RFMA(y, CCRS_RMUL(0.5, e), y);
  }
  return y;
}

// in-place lower Cholesky of a dense n x n SPD matrix; false on a non-positive (or non-finite) pivot. The diagonal holds
// the RECIPROCAL pivots 1/l_jj.
CCRS_RULE_HD bool chol_factor(double* A, int n) {
  CCRS_RULE_UNROLL
  for (int j = 0; j < n; ++j) {
    double s = A[j * n + j];
    CCRS_RULE_UNROLL
    for (int k = 0; k < j; ++k) s = CCRS_RSUB(s, CCRS_RMUL(A[j * n + k], A[j * n + k]));
    if (!(s > 0.0) || s > 1e300) return false;
    const double il = rule_rsqrt(s);
    A[j * n + j] = il;
    CCRS_RULE_UNROLL
    for (int i = j + 1; i < n; ++i) {
      double t = A[i * n + j];
      CCRS_RULE_UNROLL
      for (int k = 0; k < j; ++k) t = CCRS_RSUB(t, CCRS_RMUL(A[i * n + k], A[j * n + k]));
      A[i * n + j] = CCRS_RMUL(t, il);
    }
  }
  return true;
}
CCRS_RULE_HD void chol_solve(const double* L, int n, double* b) {
  CCRS_RULE_UNROLL
  for (int i = 0; i < n; ++i) {
    double s = b[i];
    CCRS_RULE_UNROLL
    for (int k = 0; k < i; ++k) s = CCRS_RSUB(s, CCRS_RMUL(L[i * n + k], b[k]));
    b[i] = CCRS_RMUL(s, L[i * n + i]);
  }
  CCRS_RULE_UNROLL
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    CCRS_RULE_UNROLL
    for (int k = i + 1; k < n; ++k) s = CCRS_RSUB(s, CCRS_RMUL(L[k * n + i], b[k]));
    b[i] = CCRS_RMUL(s, L[i * n + i]);
  }
}

struct Reduced {  // view into one problem's reduce() output:  S (d x d) | g_s | g_a | diag_a | sq_err
  const double *S, *gs, *ga, *diag;
  double sq_err;
};
CCRS_RULE_HD Reduced view(const double* out, int d) {
  return Reduced{out, out + d * d, out + d * d + d, out + d * d + 2 * d, out[d * d + 3 * d]};
}

// Solve the damped intrinsic system of one problem (d <= kMaxD). Returns 0 / kErrCholesky.
// fixed[i] == 2, or fixed[i] != 0 with fixed_mode == 1: the variable is removed from the linear system.
CCRS_RULE_HD int solve_intrinsics(const Reduced& r, int d, double u, double min_diag, double max_diag, const unsigned char* fixed,
                                  int fixed_mode, double* y, double* model_dec_a) {
  double S[kMaxD * kMaxD], g[kMaxD], dd[kMaxD];
  CCRS_RULE_UNROLL
  for (int i = 0; i < d * d; ++i) S[i] = r.S[i];
  CCRS_RULE_UNROLL
  for (int i = 0; i < d; ++i) {
    g[i] = r.gs[i];
    dd[i] = dmin(dmax(r.diag[i], min_diag), max_diag);
    S[i * d + i] = CCRS_RADD(S[i * d + i], CCRS_RMUL(u, dd[i]));
  }
  if (fixed)
    CCRS_RULE_UNROLL
    for (int i = 0; i < d; ++i)
      if (fixed[i] == 2 || (fixed[i] && fixed_mode == 1)) {
        CCRS_RULE_UNROLL
        for (int j = 0; j < d; ++j) { S[i * d + j] = 0; S[j * d + i] = 0; }
        S[i * d + i] = 1; g[i] = 0;
      }
  CCRS_RULE_UNROLL
  for (int i = 0; i < d * d; ++i) if (is_nan(S[i])) return kErrCholesky;  // poisoned by a failed frame pivot
  if (!chol_factor(S, d)) return kErrCholesky;
  CCRS_RULE_UNROLL
  for (int i = 0; i < d; ++i) y[i] = g[i];
  chol_solve(S, d, y);
  if (model_dec_a) {  // y_a^T g'_a + u * sum dd_i y_i^2 (intrinsic part of y^T(2g' - H'y), using H_reg y = g')
    double md = 0.0;
    CCRS_RULE_UNROLL
    for (int i = 0; i < d; ++i)
      md = CCRS_RADD(md, CCRS_RADD(CCRS_RMUL(y[i], r.ga[i]), CCRS_RMUL(CCRS_RMUL(CCRS_RMUL(u, dd[i]), y[i]), y[i])));
    *model_dec_a = md;
  }
  return 0;
}

// ParameterBlock::update_params: new = old + dx ; clamp bounded indices ; fixed indices keep the old value
CCRS_RULE_HD void update_intr(int d, const double* intr, const double* dx, const double* lo, const double* hi,
                              const unsigned char* fixed, double* out) {
  CCRS_RULE_UNROLL
  for (int i = 0; i < d; ++i) {
    double v = CCRS_RADD(intr[i], dx[i]);
    if (lo && hi) v = dmin(dmax(v, lo[i]), hi[i]);
    if (fixed && fixed[i]) v = intr[i];
    out[i] = v;
  }
}

// ---- Levenberg-Marquardt accept / reject + damping update for one problem ---------------------------------------
struct LmState {
  double u, v;        // damping 1/radius, reject factor
  double cur_err;     // error metric at the current point
};
// sq_cur: sum r^2 at the current point; sq_new: at the trial point; md: model decrease y^T(2g' - H'y) (intrinsic +
// pose part). Returns 1 if the trial point is accepted. *rho_out = gain ratio.
CCRS_RULE_HD int lm_decide(double sq_cur, double sq_new, double md, LmState* st, double* rho_out) {
  const double rho = CCRS_RDIV(CCRS_RSUB(sq_cur, sq_new), md);
  *rho_out = rho;
  if (rho > 0.0) {
    const double t = CCRS_RSUB(CCRS_RMUL(2.0, rho), 1.0);
    st->u = CCRS_RMUL(st->u, dmax(kLmMinAcceptFactor, CCRS_RSUB(1.0, CCRS_RMUL(CCRS_RMUL(t, t), t))));
    st->v = kLmRejectFactor0;
    st->cur_err = err_metric(sq_new);
    return 1;
  }
  st->u = CCRS_RMUL(st->u, st->v);
  st->v = CCRS_RMUL(st->v, 2.0);
  return 0;
}
// Stop tests after an LM decision (they compare successive ACCEPTED errors). Returns the stop reason (0 = go on,
// 1 error < min, 2 abs decrease, 3 rel decrease) and sets *status on a NaN.
CCRS_RULE_HD int lm_stop(double last_err, double cur_err, double rho, int accepted, double min_error, double min_abs,
                         double min_rel, int* status) {
  if (cur_err < min_error) return 1;
  if (is_nan(cur_err) || is_nan(rho)) { *status = kErrNumeric; return -1; }
  if (accepted) {
    const double dec = dabs(CCRS_RSUB(last_err, cur_err));
    if (dec < min_abs) return 2;
    if (CCRS_RDIV(dec, last_err) < min_rel) return 3;
  }
  return 0;
}
// Gauss-Newton stop tests at the top of iteration `it` (tiny-solver compares successive total errors).
CCRS_RULE_HD int gn_stop(int it, double last_err, double err, double min_error, double min_abs, double min_rel, int* status) {
  if (err < min_error) return 1;
  if (is_nan(err)) { *status = kErrNumeric; return -1; }
  if (it > 0) {
    const double dec = dabs(CCRS_RSUB(last_err, err));
    if (dec < min_abs) return 2;
    if (CCRS_RDIV(dec, last_err) < min_rel) return 3;
  }
  return 0;
}

}  // namespace ccrs_rule

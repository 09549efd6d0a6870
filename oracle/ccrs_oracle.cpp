// ccrs_oracle.cpp — CPU ORACLE (test infrastructure, NOT product code). See ccrs_oracle.hpp
// for the parity status. C ABI so tests can drive it through ctypes.
//
// Restates, with dual-number autodiff and OpenMP over frames:
//   * per-observation residual + Jacobian (+ Huber corrector)     factors.rs:152-173, util.rs:409-414
//   * the OtherCam variant                                         factors.rs:204-228, util.rs:612-631
//   * normal equations JtJ / -Jtr and their solution               tiny-solver SparseCholeskySolver (App. B)
//   * GaussNewtonOptimizer::optimize loop                          util.rs:443-458 (the loop the reference runs)
//   * LevenbergMarquardtOptimizer::optimize loop                   (north_star; never called by the reference)
//   * bounds clamp + fixed-variable reset                          util.rs:29-71, tiny-solver ParameterBlock::update_params
#include "ccrs_oracle.hpp"

#include <algorithm>
#include <cstdio>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using namespace ccrs_oracle;

extern "C" {

typedef struct {
  int model, width, height, xy_same_focal;
  int n_frames;
  const int32_t* frame_offsets;  // n_frames + 1 (CSR over observations)
  const double *x, *y, *z;       // board points, widened from f32 (detected_points.rs:6-9, factors.rs:141-143)
  const double *u, *v;           // observations, widened from f32
  double huber_delta;            // util.rs:413 -> 1.0 ; <= 0 disables the loss
  int n_threads;                 // OpenMP threads (0 = all)
} oracle_problem_t;

typedef struct {
  int max_iteration;
  double min_abs_decrease, min_rel_decrease, min_error;
  double lm_initial_radius, lm_min_diag, lm_max_diag;
  int fixed_mode;  // 0 = keep fixed variables in the linear system and reset them after the update (App. B),
                   // 1 = eliminate them from the reduced system (proper constrained step)
  int solver;      // 0 = per-frame elimination (Schur), 1 = dense Cholesky of the whole (d+6F) system
} oracle_options_t;

typedef struct {
  int iterations;       // linearisations performed
  int status;           // 0 ok, -1 NaN error, -2 Cholesky failure
  int stop_reason;      // 0 max_iter, 1 err<min, 2 abs decrease, 3 rel decrease
  double final_error;
  int n_accepted, n_rejected;  // LM only
} oracle_result_t;

int oracle_model_nparams(int model) { return model_nparams(model); }

void oracle_default_options(oracle_options_t* o) {
  o->max_iteration = kMaxIteration;
  o->min_abs_decrease = kMinAbsErrDecrease;
  o->min_rel_decrease = kMinRelErrDecrease;
  o->min_error = kMinErrThreshold;
  o->lm_initial_radius = kLmInitialRadius;
  o->lm_min_diag = kLmMinDiagonal;
  o->lm_max_diag = kLmMaxDiagonal;
  o->fixed_mode = 0;
  o->solver = 0;
}

void oracle_project(int model, const double* params, const double* P, double* uv) {
  project_one<double>(model, params, P, uv);
}

// rvec,tvec -> R (row-major 3x3) and back-applied point; pins types.rs round trip
void oracle_transform_point(const double* rvec, const double* tvec, const double* p, double* out) {
  isometry_apply<double>(rvec, tvec, p, out);
}

}  // extern "C"

namespace {

inline int d_eff(const oracle_problem_t* pb) { return model_nparams(pb->model) - (pb->xy_same_focal ? 1 : 0); }

void set_threads(const oracle_problem_t* pb) {
#ifdef _OPENMP
  if (pb->n_threads > 0) omp_set_num_threads(pb->n_threads);
#endif
}

// residual_func with T = f64 (factors.rs:152-173)
inline void residual_f64(const oracle_problem_t* pb, const double* intr, const double* pose, int k, double r[2]) {
  const int d = d_eff(pb);
  double prm[16];
  if (pb->xy_same_focal) { prm[0] = intr[0]; prm[1] = intr[0]; for (int i = 1; i < d; ++i) prm[i + 1] = intr[i]; }
  else for (int i = 0; i < d; ++i) prm[i] = intr[i];
  const double p[3] = {pb->x[k], pb->y[k], pb->z[k]};
  double P[3], uv[2];
  isometry_apply<double>(pose, pose + 3, p, P);
  project_one<double>(pb->model, prm, P, uv);
  r[0] = uv[0] - pb->u[k]; r[1] = uv[1] - pb->v[k];
}

// residual_func with T = dual; partial order = [intr (d) | rvec (3) | tvec (3)]
inline void residual_dual(const oracle_problem_t* pb, const double* intr, const double* pose, int k,
                          double r[2], double* J0, double* J1) {
  const int d = d_eff(pb);
  const int n = d + 6;
  Dual a[16];
  for (int i = 0; i < d; ++i) a[i] = Dual::var(intr[i], n, i);
  Dual prm[16];
  if (pb->xy_same_focal) { prm[0] = a[0]; prm[1] = a[0]; for (int i = 1; i < d; ++i) prm[i + 1] = a[i]; }  // insert_row(1, params0[0])
  else for (int i = 0; i < d; ++i) prm[i] = a[i];
  Dual rv[3], tv[3];
  for (int i = 0; i < 3; ++i) { rv[i] = Dual::var(pose[i], n, d + i); tv[i] = Dual::var(pose[3 + i], n, d + 3 + i); }
  Dual p[3] = {Dual(pb->x[k], n), Dual(pb->y[k], n), Dual(pb->z[k], n)};
  Dual P[3], uv[2];
  isometry_apply<Dual>(rv, tv, p, P);
  project_one<Dual>(pb->model, prm, P, uv);
  Dual r0 = uv[0] - pb->u[k], r1 = uv[1] - pb->v[k];
  r[0] = r0.v; r[1] = r1.v;
  for (int i = 0; i < n; ++i) { J0[i] = r0.d[i]; J1[i] = r1.d[i]; }
}

inline int tri_idx(int NA, int i, int j) { return i * NA - (i * (i - 1)) / 2 + (j - i); }  // i <= j

// dense symmetric positive definite solve, in place lower Cholesky. returns false on non-positive pivot.
bool chol_factor(std::vector<double>& A, int n) {
  for (int j = 0; j < n; ++j) {
    double s = A[j * n + j];
    for (int k = 0; k < j; ++k) s -= A[j * n + k] * A[j * n + k];
    if (!(s > 0.0)) return false;
    const double l = std::sqrt(s);
    A[j * n + j] = l;
    for (int i = j + 1; i < n; ++i) {
      double t = A[i * n + j];
      for (int k = 0; k < j; ++k) t -= A[i * n + k] * A[j * n + k];
      A[i * n + j] = t / l;
    }
  }
  return true;
}
void chol_solve(const std::vector<double>& L, int n, double* b) {
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= L[i * n + k] * b[k]; b[i] = s / L[i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= L[k * n + i] * b[k]; b[i] = s / L[i * n + i]; }
}

struct Lin {
  int d, n, NA, nblk, F;
  std::vector<double> blk;  // F * nblk : packed upper triangle of [J r]^T [J r] per frame
  double sq_err;            // sum of corrected r^2
};

void linearize(const oracle_problem_t* pb, const double* intr, const double* poses, Lin& L) {
  set_threads(pb);
  L.d = d_eff(pb); L.n = L.d + 6; L.NA = L.n + 1; L.nblk = L.NA * (L.NA + 1) / 2; L.F = pb->n_frames;
  L.blk.assign((size_t)L.F * L.nblk, 0.0);
  double total = 0.0;
  const int n = L.n, NA = L.NA;
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (int f = 0; f < pb->n_frames; ++f) {
    double* B = &L.blk[(size_t)f * L.nblk];
    double rows[2][32];
    for (int k = pb->frame_offsets[f]; k < pb->frame_offsets[f + 1]; ++k) {
      double r[2];
      residual_dual(pb, intr, poses + 6 * f, k, r, rows[0], rows[1]);
      const double w = huber_sqrt_rho1(r[0] * r[0] + r[1] * r[1], pb->huber_delta);  // Corrector
      for (int c = 0; c < 2; ++c) {
        double* row = rows[c];
        for (int i = 0; i < n; ++i) row[i] *= w;
        row[n] = r[c] * w;
        int idx = 0;
        for (int i = 0; i < NA; ++i) for (int j = i; j < NA; ++j) B[idx++] += row[i] * row[j];
      }
    }
    total += B[L.nblk - 1];
  }
  L.sq_err = total;
}

double sq_error_only(const oracle_problem_t* pb, const double* intr, const double* poses) {
  set_threads(pb);
  double total = 0.0;
#pragma omp parallel for schedule(static) reduction(+ : total)
  for (int f = 0; f < pb->n_frames; ++f) {
    double acc = 0.0;
    for (int k = pb->frame_offsets[f]; k < pb->frame_offsets[f + 1]; ++k) {
      double r[2];
      residual_f64(pb, intr, poses + 6 * f, k, r);
      const double s = r[0] * r[0] + r[1] * r[1];
      const double w = huber_sqrt_rho1(s, pb->huber_delta);
      acc += s * w * w;
    }
    total += acc;
  }
  return total;
}

// Solve (D H D + u diag(clamp(diag(D H D)))) y = D g, dx = D y from per-frame blocks.
// scale == nullptr -> D = I. Returns false if a Cholesky pivot is non-positive.
// model_decrease = y^T (2 g' - H' y) with the UNregularised scaled system (LM gain-ratio denominator).
bool solve_blocks(const Lin& L, double u, const double* scale, double min_diag, double max_diag,
                  const unsigned char* fixed, int fixed_mode, int solver,
                  double* d_intr, double* d_poses, double* model_decrease) {
  const int d = L.d, n = L.n, NA = L.NA, F = L.F;
  auto H = [&](const double* B, int i, int j) { return i <= j ? B[tri_idx(NA, i, j)] : B[tri_idx(NA, j, i)]; };
  auto sc_a = [&](int i) { return scale ? scale[i] : 1.0; };
  auto sc_p = [&](int f, int i) { return scale ? scale[d + 6 * f + i] : 1.0; };
  auto damp = [&](double hjj) { return u * std::min(std::max(hjj, min_diag), max_diag); };

  // scaled intrinsic block A', gradient g'_a (g = -J^T r)
  std::vector<double> A(d * d, 0.0), ga(d, 0.0);
  for (int f = 0; f < F; ++f) {
    const double* B = &L.blk[(size_t)f * L.nblk];
    for (int i = 0; i < d; ++i) {
      for (int j = 0; j < d; ++j) A[i * d + j] += sc_a(i) * H(B, i, j) * sc_a(j);
      ga[i] -= sc_a(i) * H(B, i, n);
    }
  }
  std::vector<double> y_a(d), y_p((size_t)6 * F);

  if (solver == 1) {
    // dense Cholesky of the whole system (what a general sparse LLT computes, without reordering)
    const int M = d + 6 * F;
    std::vector<double> K((size_t)M * M, 0.0), rhs(M, 0.0);
    for (int i = 0; i < d; ++i) { for (int j = 0; j < d; ++j) K[(size_t)i * M + j] = A[i * d + j]; rhs[i] = ga[i]; }
    for (int f = 0; f < F; ++f) {
      const double* B = &L.blk[(size_t)f * L.nblk];
      for (int i = 0; i < 6; ++i) {
        const int gi = d + 6 * f + i;
        for (int j = 0; j < 6; ++j) K[(size_t)gi * M + d + 6 * f + j] = sc_p(f, i) * H(B, d + i, d + j) * sc_p(f, j);
        for (int a = 0; a < d; ++a) { const double v = sc_a(a) * H(B, a, d + i) * sc_p(f, i); K[(size_t)a * M + gi] = v; K[(size_t)gi * M + a] = v; }
        rhs[gi] = -sc_p(f, i) * H(B, d + i, n);
      }
    }
    std::vector<double> Kun = K;
    for (int i = 0; i < M; ++i) K[(size_t)i * M + i] += damp(Kun[(size_t)i * M + i]);
    if (fixed && fixed_mode == 1)
      for (int i = 0; i < d; ++i) if (fixed[i]) { for (int j = 0; j < M; ++j) { K[(size_t)i * M + j] = 0; K[(size_t)j * M + i] = 0; } K[(size_t)i * M + i] = 1; rhs[i] = 0; }
    std::vector<double> g0 = rhs;
    if (!chol_factor(K, M)) return false;
    chol_solve(K, M, rhs.data());
    double md = 0.0;
    for (int i = 0; i < M; ++i) { double hy = 0.0; for (int j = 0; j < M; ++j) hy += Kun[(size_t)i * M + j] * rhs[j]; md += rhs[i] * (2.0 * g0[i] - hy); }
    for (int i = 0; i < d; ++i) y_a[i] = rhs[i];
    for (size_t i = 0; i < y_p.size(); ++i) y_p[i] = rhs[d + i];
    if (model_decrease) *model_decrease = md;
  } else {
    // per-frame elimination onto the intrinsic system
    std::vector<double> S = A, gs = ga;
    for (int i = 0; i < d; ++i) S[i * d + i] += damp(A[i * d + i]);
    std::vector<double> CiBt((size_t)F * 6 * d), Cig((size_t)F * 6);
    bool ok = true;
#pragma omp parallel
    {
      std::vector<double> Sl(d * d, 0.0), gl(d, 0.0);
#pragma omp for schedule(static)
      for (int f = 0; f < F; ++f) {
        const double* B = &L.blk[(size_t)f * L.nblk];
        std::vector<double> C(36);
        double gp[6];
        for (int i = 0; i < 6; ++i) {
          for (int j = 0; j < 6; ++j) C[i * 6 + j] = sc_p(f, i) * H(B, d + i, d + j) * sc_p(f, j);
          gp[i] = -sc_p(f, i) * H(B, d + i, n);
        }
        for (int i = 0; i < 6; ++i) C[i * 6 + i] += damp(C[i * 6 + i]);
        if (!chol_factor(C, 6)) {
#pragma omp critical
          ok = false;
          continue;
        }
        double* X = &CiBt[(size_t)f * 6 * d];  // column a: C^-1 B'^T[:,a]
        for (int a = 0; a < d; ++a) {
          double col[6];
          for (int i = 0; i < 6; ++i) col[i] = sc_a(a) * H(B, a, d + i) * sc_p(f, i);
          chol_solve(C, 6, col);
          for (int i = 0; i < 6; ++i) X[i * d + a] = col[i];
        }
        double* cg = &Cig[(size_t)f * 6];
        for (int i = 0; i < 6; ++i) cg[i] = gp[i];
        chol_solve(C, 6, cg);
        for (int a = 0; a < d; ++a) {
          for (int b = 0; b < d; ++b) {
            double s = 0.0;
            for (int i = 0; i < 6; ++i) s += sc_a(a) * H(B, a, d + i) * sc_p(f, i) * X[i * d + b];
            Sl[a * d + b] -= s;
          }
          double s = 0.0;
          for (int i = 0; i < 6; ++i) s += sc_a(a) * H(B, a, d + i) * sc_p(f, i) * cg[i];
          gl[a] -= s;
        }
      }
#pragma omp critical
      {
        for (int i = 0; i < d * d; ++i) S[i] += Sl[i];
        for (int i = 0; i < d; ++i) gs[i] += gl[i];
      }
    }
    if (!ok) return false;
    if (fixed && fixed_mode == 1)
      for (int i = 0; i < d; ++i) if (fixed[i]) { for (int j = 0; j < d; ++j) { S[i * d + j] = 0; S[j * d + i] = 0; } S[i * d + i] = 1; gs[i] = 0; }
    if (!chol_factor(S, d)) return false;
    for (int i = 0; i < d; ++i) y_a[i] = gs[i];
    chol_solve(S, d, y_a.data());
    for (int f = 0; f < F; ++f) {
      const double* X = &CiBt[(size_t)f * 6 * d];
      for (int i = 0; i < 6; ++i) {
        double s = Cig[(size_t)f * 6 + i];
        for (int a = 0; a < d; ++a) s -= X[i * d + a] * y_a[a];
        y_p[(size_t)6 * f + i] = s;
      }
    }
    if (model_decrease) {
      double md = 0.0;
      for (int i = 0; i < d; ++i) { double hy = 0.0; for (int j = 0; j < d; ++j) hy += A[i * d + j] * y_a[j]; md += y_a[i] * (2.0 * ga[i] - hy); }
      for (int f = 0; f < F; ++f) {
        const double* B = &L.blk[(size_t)f * L.nblk];
        const double* yp = &y_p[(size_t)6 * f];
        for (int i = 0; i < 6; ++i) {
          double hy = 0.0;
          for (int j = 0; j < 6; ++j) hy += sc_p(f, i) * H(B, d + i, d + j) * sc_p(f, j) * yp[j];
          double by = 0.0;
          for (int a = 0; a < d; ++a) by += sc_a(a) * H(B, a, d + i) * sc_p(f, i) * y_a[a];
          md += yp[i] * (2.0 * (-sc_p(f, i) * H(B, d + i, n)) - hy - 2.0 * by);
        }
      }
      *model_decrease = md;
    }
  }
  for (int i = 0; i < d; ++i) d_intr[i] = sc_a(i) * y_a[i];
  for (int f = 0; f < F; ++f) for (int i = 0; i < 6; ++i) d_poses[(size_t)6 * f + i] = sc_p(f, i) * y_p[(size_t)6 * f + i];
  return true;
}

// ParameterBlock::update_params: new = old + dx; clamp bounded indices; fixed indices keep the old value
void apply_update(int d, int F, const double* lo, const double* hi, const unsigned char* fixed,
                  const double* d_intr, const double* d_poses, const double* intr, const double* poses,
                  double* intr_new, double* poses_new) {
  for (int i = 0; i < d; ++i) {
    double v = intr[i] + d_intr[i];
    if (lo && hi) v = std::min(std::max(v, lo[i]), hi[i]);
    if (fixed && fixed[i]) v = intr[i];
    intr_new[i] = v;
  }
  for (size_t i = 0; i < (size_t)6 * F; ++i) poses_new[i] = poses[i] + d_poses[i];
}

inline double err_metric(double sq) { return kErrorIsL2Norm ? std::sqrt(sq) : sq; }

}  // namespace

extern "C" {

// Parity hook a2: r (2N), J (2N x (d_eff+6), row-major). apply_loss -> Huber-corrected (what tiny-solver assembles).
void oracle_eval_rj(const oracle_problem_t* pb, const double* intr, const double* poses, int apply_loss,
                    double* r_out, double* J_out) {
  set_threads(pb);
  const int n = d_eff(pb) + 6;
#pragma omp parallel for schedule(static)
  for (int f = 0; f < pb->n_frames; ++f) {
    for (int k = pb->frame_offsets[f]; k < pb->frame_offsets[f + 1]; ++k) {
      double r[2], J0[32], J1[32];
      residual_dual(pb, intr, poses + 6 * f, k, r, J0, J1);
      const double w = apply_loss ? huber_sqrt_rho1(r[0] * r[0] + r[1] * r[1], pb->huber_delta) : 1.0;
      r_out[2 * (size_t)k] = r[0] * w; r_out[2 * (size_t)k + 1] = r[1] * w;
      if (J_out) for (int i = 0; i < n; ++i) { J_out[(size_t)(2 * k) * n + i] = J0[i] * w; J_out[(size_t)(2 * k + 1) * n + i] = J1[i] * w; }
    }
  }
}

// Residual only with T = f64 (the path the reference's own test exercises, optimization_test.rs:36-80)
void oracle_eval_r(const oracle_problem_t* pb, const double* intr, const double* poses, int apply_loss, double* r_out) {
  set_threads(pb);
#pragma omp parallel for schedule(static)
  for (int f = 0; f < pb->n_frames; ++f)
    for (int k = pb->frame_offsets[f]; k < pb->frame_offsets[f + 1]; ++k) {
      double r[2];
      residual_f64(pb, intr, poses + 6 * f, k, r);
      const double w = apply_loss ? huber_sqrt_rho1(r[0] * r[0] + r[1] * r[1], pb->huber_delta) : 1.0;
      r_out[2 * (size_t)k] = r[0] * w; r_out[2 * (size_t)k + 1] = r[1] * w;
    }
}

// validation (src/util.rs:721-795): per-point error sqrt(dx^2 + dy^2) of project_one(T * p3d) - p2d WITHOUT loss
// (:733-745); all errors sorted ascending (:771); median = sorted[len / 2] (:772); mean of the first len * 99 / 100
// computed as sum(p / len99) in sorted order (:777-781). errors_out (nullable): per-point errors, observation order.
void oracle_validation(const oracle_problem_t* pb, const double* intr, const double* poses, double* median, double* avg99,
                       double* errors_out) {
  const size_t N = (size_t)pb->frame_offsets[pb->n_frames];
  std::vector<double> e(N);
  set_threads(pb);
#pragma omp parallel for schedule(static)
  for (int f = 0; f < pb->n_frames; ++f)
    for (int k = pb->frame_offsets[f]; k < pb->frame_offsets[f + 1]; ++k) {
      double r[2];
      residual_f64(pb, intr, poses + 6 * f, k, r);
      e[k] = std::sqrt(r[0] * r[0] + r[1] * r[1]);
    }
  if (errors_out) std::copy(e.begin(), e.end(), errors_out);
  std::sort(e.begin(), e.end());
  *median = N ? e[N / 2] : 0.0;
  const size_t len99 = N * 99 / 100;
  double s = 0.0;
  for (size_t i = 0; i < len99; ++i) s += e[i] / (double)len99;
  *avg99 = s;
}

// Parity hook a3 (OtherCamReprojectionFactor, factors.rs:204-228): per observation k of camera i,
// pose0 = T_0_b of its frame, pose1 = T_i_0. J: 2 x (d_eff + 12), columns [intr | rvec_0_b tvec_0_b | rvec_i_0 tvec_i_0].
void oracle_othercam_rj(const oracle_problem_t* pb, const double* intr, const double* poses_0_b, const double* pose_i_0,
                        int apply_loss, double* r_out, double* J_out) {
  set_threads(pb);
  const int d = d_eff(pb), n = d + 12;
#pragma omp parallel for schedule(static)
  for (int f = 0; f < pb->n_frames; ++f) {
    for (int k = pb->frame_offsets[f]; k < pb->frame_offsets[f + 1]; ++k) {
      Dual a[16], prm[16];
      for (int i = 0; i < d; ++i) a[i] = Dual::var(intr[i], n, i);
      if (pb->xy_same_focal) { prm[0] = a[0]; prm[1] = a[0]; for (int i = 1; i < d; ++i) prm[i + 1] = a[i]; }
      else for (int i = 0; i < d; ++i) prm[i] = a[i];
      Dual rv0[3], tv0[3], rv1[3], tv1[3];
      for (int i = 0; i < 3; ++i) {
        rv0[i] = Dual::var(poses_0_b[6 * f + i], n, d + i);       tv0[i] = Dual::var(poses_0_b[6 * f + 3 + i], n, d + 3 + i);
        rv1[i] = Dual::var(pose_i_0[i], n, d + 6 + i);            tv1[i] = Dual::var(pose_i_0[3 + i], n, d + 9 + i);
      }
      Dual p[3] = {Dual(pb->x[k], n), Dual(pb->y[k], n), Dual(pb->z[k], n)};
      Dual P[3], uv[2];
      isometry_chain_apply<Dual>(rv1, tv1, rv0, tv0, p, P);
      project_one<Dual>(pb->model, prm, P, uv);
      Dual r0 = uv[0] - pb->u[k], r1 = uv[1] - pb->v[k];
      const double w = apply_loss ? huber_sqrt_rho1(r0.v * r0.v + r1.v * r1.v, pb->huber_delta) : 1.0;
      r_out[2 * (size_t)k] = r0.v * w; r_out[2 * (size_t)k + 1] = r1.v * w;
      if (J_out) for (int i = 0; i < n; ++i) { J_out[(size_t)(2 * k) * n + i] = r0.d[i] * w; J_out[(size_t)(2 * k + 1) * n + i] = r1.d[i] * w; }
    }
  }
}

int oracle_nblk(const oracle_problem_t* pb) { const int NA = d_eff(pb) + 7; return NA * (NA + 1) / 2; }

// Per-frame packed Gram blocks [J r]^T [J r] (same layout the GPU library exposes for parity). Returns sum r^2.
double oracle_linearize(const oracle_problem_t* pb, const double* intr, const double* poses, double* blocks_out) {
  Lin L;
  linearize(pb, intr, poses, L);
  if (blocks_out) std::memcpy(blocks_out, L.blk.data(), L.blk.size() * sizeof(double));
  return L.sq_err;
}

double oracle_sq_error(const oracle_problem_t* pb, const double* intr, const double* poses) {
  return sq_error_only(pb, intr, poses);
}

// One damped/undamped step from a linearisation at (intr, poses). scale nullable (d_eff + 6F).
int oracle_solve_step(const oracle_problem_t* pb, const double* intr, const double* poses, double u,
                      const double* scale, const oracle_options_t* opt, const unsigned char* fixed,
                      double* d_intr, double* d_poses, double* model_decrease) {
  Lin L;
  linearize(pb, intr, poses, L);
  return solve_blocks(L, u, scale, opt->lm_min_diag, opt->lm_max_diag, fixed, opt->fixed_mode, opt->solver,
                      d_intr, d_poses, model_decrease) ? 0 : -2;
}

// GaussNewtonOptimizer::optimize (App. B / SURVEY 3.3). intr/poses updated in place.
// err_hist (nullable, max_iteration entries) receives the error seen at each linearisation.
int oracle_gn(const oracle_problem_t* pb, double* intr, double* poses, const double* lo, const double* hi,
              const unsigned char* fixed, const oracle_options_t* opt, oracle_result_t* res, double* err_hist) {
  const int d = d_eff(pb), F = pb->n_frames;
  std::vector<double> di(d), dp((size_t)6 * F), intr_n(d), poses_n((size_t)6 * F);
  double last_err = 0.0;
  res->iterations = 0; res->status = 0; res->stop_reason = 0; res->n_accepted = res->n_rejected = 0;
  Lin L;
  for (int it = 0; it < opt->max_iteration; ++it) {
    linearize(pb, intr, poses, L);
    const double err = err_metric(L.sq_err);
    if (err_hist) err_hist[it] = err;
    res->iterations = it + 1; res->final_error = err;
    if (err < opt->min_error) { res->stop_reason = 1; break; }
    if (std::isnan(err)) { res->status = -1; return -1; }
    if (it > 0) {
      if (std::fabs(last_err - err) < opt->min_abs_decrease) { res->stop_reason = 2; break; }
      if (std::fabs(last_err - err) / last_err < opt->min_rel_decrease) { res->stop_reason = 3; break; }
    }
    last_err = err;
    if (!solve_blocks(L, 0.0, nullptr, opt->lm_min_diag, opt->lm_max_diag, fixed, opt->fixed_mode, opt->solver,
                      di.data(), dp.data(), nullptr)) { res->status = -2; return -2; }
    apply_update(d, F, lo, hi, fixed, di.data(), dp.data(), intr, poses, intr_n.data(), poses_n.data());
    std::memcpy(intr, intr_n.data(), d * sizeof(double));
    std::memcpy(poses, poses_n.data(), poses_n.size() * sizeof(double));
  }
  return 0;
}

// LevenbergMarquardtOptimizer::optimize (App. B; Ceres-style). Jacobi column scaling from iteration 0.
// err_hist[it] = error of the ACCEPTED point after iteration it.
int oracle_lm(const oracle_problem_t* pb, double* intr, double* poses, const double* lo, const double* hi,
              const unsigned char* fixed, const oracle_options_t* opt, oracle_result_t* res, double* err_hist) {
  const int d = d_eff(pb), F = pb->n_frames, n = d + 6;
  std::vector<double> di(d), dp((size_t)6 * F), intr_n(d), poses_n((size_t)6 * F), scale((size_t)d + 6 * F);
  res->iterations = 0; res->status = 0; res->stop_reason = 0; res->n_accepted = res->n_rejected = 0;
  double u = 1.0 / opt->lm_initial_radius, v = kLmRejectFactor0;
  Lin L;
  double cur_sq = sq_error_only(pb, intr, poses);
  double current_error = err_metric(cur_sq);
  bool need_linearize = true;
  for (int it = 0; it < opt->max_iteration; ++it) {
    const double last_err = current_error;
    if (need_linearize) linearize(pb, intr, poses, L);
    if (it == 0) {
      // jacobi scaling 1 / (1 + ||J[:,c]||) from the first (loss-corrected) Jacobian
      std::vector<double> cn(d, 0.0);
      for (int f = 0; f < F; ++f) {
        const double* B = &L.blk[(size_t)f * L.nblk];
        for (int i = 0; i < d; ++i) cn[i] += B[tri_idx(L.NA, i, i)];
        for (int i = 0; i < 6; ++i) scale[d + 6 * f + i] = 1.0 / (1.0 + std::sqrt(B[tri_idx(L.NA, d + i, d + i)]));
      }
      for (int i = 0; i < d; ++i) scale[i] = 1.0 / (1.0 + std::sqrt(cn[i]));
    }
    double model_dec = 0.0;
    if (!solve_blocks(L, u, scale.data(), opt->lm_min_diag, opt->lm_max_diag, fixed, opt->fixed_mode, opt->solver,
                      di.data(), dp.data(), &model_dec)) { res->status = -2; return -2; }
    apply_update(d, F, lo, hi, fixed, di.data(), dp.data(), intr, poses, intr_n.data(), poses_n.data());
    const double new_sq = sq_error_only(pb, intr_n.data(), poses_n.data());
    const double rho = (L.sq_err - new_sq) / model_dec;
    res->iterations = it + 1;
    if (rho > 0.0) {
      std::memcpy(intr, intr_n.data(), d * sizeof(double));
      std::memcpy(poses, poses_n.data(), poses_n.size() * sizeof(double));
      const double t = 2.0 * rho - 1.0;
      u *= std::max(1.0 / 3.0, 1.0 - t * t * t);
      v = kLmRejectFactor0;
      cur_sq = new_sq;
      need_linearize = true;
      res->n_accepted++;
    } else {
      u *= v; v *= 2.0;
      need_linearize = false;  // same point, only the damping changes
      res->n_rejected++;
    }
    current_error = err_metric(cur_sq);
    if (err_hist) err_hist[it] = current_error;
    res->final_error = current_error;
    if (current_error < opt->min_error) { res->stop_reason = 1; break; }
    if (std::isnan(current_error)) { res->status = -1; return -1; }
    if (rho > 0.0) {  // stop tests compare successive ACCEPTED errors (a rejected step leaves the error unchanged)
      if (std::fabs(last_err - current_error) < opt->min_abs_decrease) { res->stop_reason = 2; break; }
      if (std::fabs(last_err - current_error) / last_err < opt->min_rel_decrease) { res->stop_reason = 3; break; }
    }
    (void)n;
  }
  return 0;
}


// ======================================================================================================
// Joint multi-camera problem: calib_all_camera_with_extrinsics (src/util.rs:567-715).
// Variables: params{c} (d each), rvec_{c}_0/tvec_{c}_0 for c > 0, rvec_0_b_{f}/tvec_0_b_{f} per frame.
// cam0 corners use ReprojectionFactor (util.rs:603-611), other cameras OtherCamReprojectionFactor (util.rs:612-631).
// Unknown order of the dense system: [params0 .. params{C-1} | T_1_0 .. T_{C-1}_0 | T_0_b_0 .. T_0_b_{F-1}].
// ======================================================================================================
typedef struct {
  int model, xy_same_focal;
  int n_cams, n_frames, n_blocks;
  const int32_t* block_cam;      // [n_blocks]
  const int32_t* block_frame;    // [n_blocks]
  const int32_t* block_offsets;  // [n_blocks+1] CSR over observations
  const double *x, *y, *z, *u, *v;
  double huber_delta;
  int n_threads;
} oracle_joint_t;

}  // extern "C"

namespace {

// residual + Jacobian of one observation of block b. Jrow layout: [params_c (d) | T_0_b (6) | T_c_0 (6)] (d+12).
inline void joint_obs(const oracle_joint_t* jp, int d, int c, const double* intr_c, const double* pose_0_b,
                      const double* pose_c_0, int k, double r[2], double* J0, double* J1) {
  const int n = d + 12;
  Dual a[16], prm[16];
  for (int i = 0; i < d; ++i) a[i] = Dual::var(intr_c[i], n, i);
  if (jp->xy_same_focal) { prm[0] = a[0]; prm[1] = a[0]; for (int i = 1; i < d; ++i) prm[i + 1] = a[i]; }
  else for (int i = 0; i < d; ++i) prm[i] = a[i];
  Dual rv0[3], tv0[3];
  for (int i = 0; i < 3; ++i) { rv0[i] = Dual::var(pose_0_b[i], n, d + i); tv0[i] = Dual::var(pose_0_b[3 + i], n, d + 3 + i); }
  Dual p[3] = {Dual(jp->x[k], n), Dual(jp->y[k], n), Dual(jp->z[k], n)};
  Dual P[3], uv[2];
  if (c == 0) {
    isometry_apply<Dual>(rv0, tv0, p, P);                       // ReprojectionFactor (factors.rs:160-164)
  } else {
    Dual rv1[3], tv1[3];
    for (int i = 0; i < 3; ++i) { rv1[i] = Dual::var(pose_c_0[i], n, d + 6 + i); tv1[i] = Dual::var(pose_c_0[3 + i], n, d + 9 + i); }
    isometry_chain_apply<Dual>(rv1, tv1, rv0, tv0, p, P);       // OtherCamReprojectionFactor (factors.rs:212-218)
  }
  project_one<Dual>(jp->model, prm, P, uv);
  Dual r0 = uv[0] - jp->u[k], r1 = uv[1] - jp->v[k];
  r[0] = r0.v; r[1] = r1.v;
  for (int i = 0; i < n; ++i) { J0[i] = r0.d[i]; J1[i] = r1.d[i]; }
}

}  // namespace

extern "C" {

// Parity hook: per observation r (2N) and J (2N x (d+12)); cam0 blocks have zero T_c_0 columns.
void oracle_joint_eval_rj(const oracle_joint_t* jp, const double* intr /*[C][d]*/, const double* extr /*[C][6]*/,
                          const double* poses /*[F][6]*/, int apply_loss, double* r_out, double* J_out) {
  const int d = model_nparams(jp->model) - (jp->xy_same_focal ? 1 : 0), n = d + 12;
  for (int b = 0; b < jp->n_blocks; ++b) {
    const int c = jp->block_cam[b], f = jp->block_frame[b];
    for (int k = jp->block_offsets[b]; k < jp->block_offsets[b + 1]; ++k) {
      double r[2], J0[32], J1[32];
      joint_obs(jp, d, c, intr + (size_t)c * d, poses + 6 * (size_t)f, extr + 6 * (size_t)c, k, r, J0, J1);
      const double w = apply_loss ? huber_sqrt_rho1(r[0] * r[0] + r[1] * r[1], jp->huber_delta) : 1.0;
      r_out[2 * (size_t)k] = r[0] * w; r_out[2 * (size_t)k + 1] = r[1] * w;
      if (J_out) for (int i = 0; i < n; ++i) { J_out[(size_t)(2 * k) * n + i] = J0[i] * w; J_out[(size_t)(2 * k + 1) * n + i] = J1[i] * w; }
    }
  }
}

// GaussNewtonOptimizer::optimize on the joint problem (util.rs:668-670): dense normal equations over ALL unknowns.
// intr [C][d], extr [C][6] (row 0 unused), poses [F][6] updated in place. lo/hi/fixed: [C][d] nullable.
int oracle_joint_gn(const oracle_joint_t* jp, double* intr, double* extr, double* poses, const double* lo,
                    const double* hi, const unsigned char* fixed, const oracle_options_t* opt, oracle_result_t* res,
                    double* err_hist) {
  const int d = model_nparams(jp->model) - (jp->xy_same_focal ? 1 : 0);
  const int C = jp->n_cams, F = jp->n_frames;
  const int off_e = C * d, off_p = C * d + 6 * (C - 1), M = off_p + 6 * F;
  res->iterations = 0; res->status = 0; res->stop_reason = 0; res->n_accepted = res->n_rejected = 0;
  std::vector<double> H((size_t)M * M), g(M);
  double last_err = 0.0;
  for (int it = 0; it < opt->max_iteration; ++it) {
    std::fill(H.begin(), H.end(), 0.0); std::fill(g.begin(), g.end(), 0.0);
    double sq = 0.0;
    for (int b = 0; b < jp->n_blocks; ++b) {
      const int c = jp->block_cam[b], f = jp->block_frame[b];
      int col[32];
      for (int i = 0; i < d; ++i) col[i] = c * d + i;
      for (int i = 0; i < 6; ++i) col[d + i] = off_p + 6 * f + i;
      for (int i = 0; i < 6; ++i) col[d + 6 + i] = c > 0 ? off_e + 6 * (c - 1) + i : -1;
      for (int k = jp->block_offsets[b]; k < jp->block_offsets[b + 1]; ++k) {
        double r[2], J[2][32];
        joint_obs(jp, d, c, intr + (size_t)c * d, poses + 6 * (size_t)f, extr + 6 * (size_t)c, k, r, J[0], J[1]);
        const double w = huber_sqrt_rho1(r[0] * r[0] + r[1] * r[1], jp->huber_delta);
        for (int q = 0; q < 2; ++q) {
          const double rw = r[q] * w;
          sq += rw * rw;
          for (int i = 0; i < d + 12; ++i) {
            if (col[i] < 0) continue;
            const double ji = J[q][i] * w;
            g[col[i]] -= ji * rw;
            for (int j = 0; j < d + 12; ++j) if (col[j] >= 0) H[(size_t)col[i] * M + col[j]] += ji * J[q][j] * w;
          }
        }
      }
    }
    const double err = err_metric(sq);
    if (err_hist) err_hist[it] = err;
    res->iterations = it + 1; res->final_error = err;
    if (err < opt->min_error) { res->stop_reason = 1; break; }
    if (std::isnan(err)) { res->status = -1; return -1; }
    if (it > 0) {
      if (std::fabs(last_err - err) < opt->min_abs_decrease) { res->stop_reason = 2; break; }
      if (std::fabs(last_err - err) / last_err < opt->min_rel_decrease) { res->stop_reason = 3; break; }
    }
    last_err = err;
    if (fixed && opt->fixed_mode == 1)
      for (int i = 0; i < C * d; ++i) if (fixed[i]) { for (int j = 0; j < M; ++j) { H[(size_t)i * M + j] = 0; H[(size_t)j * M + i] = 0; } H[(size_t)i * M + i] = 1; g[i] = 0; }
    if (!chol_factor(H, M)) { res->status = -2; return -2; }
    chol_solve(H, M, g.data());
    for (int i = 0; i < C * d; ++i) {
      double v = intr[i] + g[i];
      if (lo && hi) v = std::min(std::max(v, lo[i]), hi[i]);
      if (fixed && fixed[i]) v = intr[i];
      intr[i] = v;
    }
    for (int c = 1; c < C; ++c) for (int i = 0; i < 6; ++i) extr[6 * c + i] += g[off_e + 6 * (c - 1) + i];
    for (int i = 0; i < 6 * F; ++i) poses[i] += g[off_p + i];
  }
  return 0;
}

}  // extern "C"

// =====================================================================================================================
// N4: the two small optimisation problems of src/optimization that use other factors.
// Restated directly from the reference (dense normal equations over the reference's own variable blocks), NOT through
// the per-frame elimination machinery above, so that they check the product's reuse of its kernels independently.
// =====================================================================================================================
namespace {

// generic dense Gauss-Newton step bookkeeping shared by the two problems below: tiny-solver semantics
// (error = ||corrected r||, stop tests on successive errors, additive update, clamp, fixed indices reset)
struct DenseGn {
  int M;                                  // number of scalars
  std::vector<double> x, lo, hi;          // lo > hi  <=> unbounded
  std::vector<unsigned char> fixed;
};

template <class Lin>
int dense_gn(DenseGn& P, const oracle_options_t* opt, oracle_result_t* res, double* err_hist, Lin&& linearize) {
  const int M = P.M;
  std::vector<double> H((size_t)M * M), g(M);
  double last_err = 0.0;
  res->iterations = 0; res->status = 0; res->stop_reason = 0; res->n_accepted = res->n_rejected = 0;
  for (int it = 0; it < opt->max_iteration; ++it) {
    std::fill(H.begin(), H.end(), 0.0); std::fill(g.begin(), g.end(), 0.0);
    const double sq = linearize(P.x.data(), H.data(), g.data());   // H = J^T J, g = -J^T r (corrected), returns sum r'^2
    const double err = err_metric(sq);
    if (err_hist) err_hist[it] = err;
    res->iterations = it + 1; res->final_error = err;
    if (err < opt->min_error) { res->stop_reason = 1; break; }
    if (std::isnan(err)) { res->status = -1; return -1; }
    if (it > 0) {
      if (std::fabs(last_err - err) < opt->min_abs_decrease) { res->stop_reason = 2; break; }
      if (std::fabs(last_err - err) / last_err < opt->min_rel_decrease) { res->stop_reason = 3; break; }
    }
    last_err = err;
    if (!chol_factor(H, M)) { res->status = -2; return -2; }
    chol_solve(H, M, g.data());
    for (int i = 0; i < M; ++i) {
      double v = P.x[i] + g[i];
      if (P.lo[i] <= P.hi[i]) v = std::min(std::max(v, P.lo[i]), P.hi[i]);
      if (P.fixed[i]) v = P.x[i];
      P.x[i] = v;
    }
  }
  return 0;
}

}  // namespace

extern "C" {

// init_ucm, first stage (src/util.rs:284-357): variables "params" = [f, alpha], "rvec{0,1}", "tvec{0,1}";
// one UCMInitFocalAlphaFactor block per feature of the two frames (factors.rs:101-120: cam = [f, f, cx, cy, alpha] with
// cx, cy the constants of the target model), HuberLoss(1.0) per block; bounds f in [f0/3, 3 f0], alpha in [1e-6, 1];
// fixed_focal fixes params[0]. Frame f owns observations [frame_offsets[f], frame_offsets[f+1]), n_frames == 2.
// f_alpha[2] and poses[2][6] in/out.
int oracle_init_ucm_gn(const oracle_problem_t* pb, double cx, double cy, double* f_alpha, double* poses, int fixed_focal,
                       const oracle_options_t* opt, oracle_result_t* res, double* err_hist) {
  const int F = pb->n_frames, M = 2 + 6 * F;
  DenseGn P; P.M = M; P.x.resize(M); P.lo.assign(M, 1.0); P.hi.assign(M, 0.0); P.fixed.assign(M, 0);
  P.x[0] = f_alpha[0]; P.x[1] = f_alpha[1];
  for (int i = 0; i < 6 * F; ++i) P.x[2 + i] = poses[i];
  P.lo[0] = f_alpha[0] / 3.0; P.hi[0] = f_alpha[0] * 3.0;     // util.rs:337
  P.lo[1] = 1e-6; P.hi[1] = 1.0;                               // util.rs:338
  if (fixed_focal) P.fixed[0] = 1;                             // util.rs:330-332
  const int st = dense_gn(P, opt, res, err_hist, [&](const double* x, double* H, double* g) {
    double sq = 0.0;
    for (int f = 0; f < F; ++f)
      for (int k = pb->frame_offsets[f]; k < pb->frame_offsets[f + 1]; ++k) {
        const int n = 8;                                       // [f, alpha | rvec | tvec]
        Dual fa[2] = {Dual::var(x[0], n, 0), Dual::var(x[1], n, 1)};
        Dual prm[5] = {fa[0], fa[0], Dual(cx, n), Dual(cy, n), fa[1]};          // factors.rs:104-108
        Dual rv[3], tv[3];
        for (int i = 0; i < 3; ++i) { rv[i] = Dual::var(x[2 + 6 * f + i], n, 2 + i); tv[i] = Dual::var(x[2 + 6 * f + 3 + i], n, 5 + i); }
        Dual p[3] = {Dual(pb->x[k], n), Dual(pb->y[k], n), Dual(pb->z[k], n)};
        Dual Pc[3], uv[2];
        isometry_apply<Dual>(rv, tv, p, Pc);
        project_one<Dual>(UCM, prm, Pc, uv);
        const Dual r0 = uv[0] - pb->u[k], r1 = uv[1] - pb->v[k];
        const double w = huber_sqrt_rho1(r0.v * r0.v + r1.v * r1.v, pb->huber_delta);
        const Dual* rr[2] = {&r0, &r1};
        auto gi = [&](int c) { return c < 2 ? c : 2 + 6 * f + (c - 2); };
        for (int q = 0; q < 2; ++q) {
          const double rq = rr[q]->v * w;
          sq += rq * rq;
          for (int a = 0; a < n; ++a) {
            const double ja = rr[q]->d[a] * w;
            g[gi(a)] -= ja * rq;
            for (int b = 0; b < n; ++b) H[(size_t)gi(a) * M + gi(b)] += ja * rr[q]->d[b] * w;
          }
        }
      }
    return sq;
  });
  f_alpha[0] = P.x[0]; f_alpha[1] = P.x[1];
  for (int i = 0; i < 6 * F; ++i) poses[i] = P.x[2 + i];
  return st;
}

// convert_model (src/util.rs:225-278) from the ModelConvertFactor onwards: ONE residual block of 2 n_pts residuals
// source.project(p3d) - target.project(p3d) (factors.rs:56-77) under ONE HuberLoss(1.0) (util.rs:246-251: the loss sees
// the squared norm of the whole block), variable "params" = the target's full parameter vector; lo/hi/fixed from
// set_problem_parameter_bound / _disabled (util.rs:262-271). p3d[n_pts][3] are the unprojected grid points.
int oracle_convert_model_gn(int src_model, const double* src_params, int tgt_model, double* tgt_params, int n_pts,
                            const double* p3d, const double* lo, const double* hi, const unsigned char* fixed,
                            double huber_delta, const oracle_options_t* opt, oracle_result_t* res, double* err_hist) {
  const int M = model_nparams(tgt_model);
  DenseGn P; P.M = M; P.x.assign(tgt_params, tgt_params + M); P.lo.assign(M, 1.0); P.hi.assign(M, 0.0); P.fixed.assign(M, 0);
  for (int i = 0; i < M; ++i) { if (lo && hi) { P.lo[i] = lo[i]; P.hi[i] = hi[i]; } if (fixed) P.fixed[i] = fixed[i]; }
  std::vector<double> uv0((size_t)2 * n_pts);
  for (int k = 0; k < n_pts; ++k) project_one<double>(src_model, src_params, p3d + 3 * k, &uv0[2 * k]);
  std::vector<double> r((size_t)2 * n_pts), J((size_t)2 * n_pts * M);
  const int st = dense_gn(P, opt, res, err_hist, [&](const double* x, double* H, double* g) {
    double s = 0.0;
    for (int k = 0; k < n_pts; ++k) {
      Dual prm[16];
      for (int i = 0; i < M; ++i) prm[i] = Dual::var(x[i], M, i);
      Dual Pc[3] = {Dual(p3d[3 * k], M), Dual(p3d[3 * k + 1], M), Dual(p3d[3 * k + 2], M)}, uv[2];
      project_one<Dual>(tgt_model, prm, Pc, uv);
      for (int q = 0; q < 2; ++q) {
        const Dual rq = uv0[2 * k + q] - uv[q];                 // p0 - p1 (factors.rs:68)
        r[2 * k + q] = rq.v; s += rq.v * rq.v;
        for (int i = 0; i < M; ++i) J[(size_t)(2 * k + q) * M + i] = rq.d[i];
      }
    }
    const double w = huber_sqrt_rho1(s, huber_delta);            // one corrector for the whole block
    for (int row = 0; row < 2 * n_pts; ++row) {
      const double rq = r[row] * w;
      for (int a = 0; a < M; ++a) {
        const double ja = J[(size_t)row * M + a] * w;
        g[a] -= ja * rq;
        for (int b = 0; b < M; ++b) H[(size_t)a * M + b] += ja * J[(size_t)row * M + b] * w;
      }
    }
    return s * w * w;
  });
  for (int i = 0; i < M; ++i) tgt_params[i] = P.x[i];
  return st;
}

}  // extern "C"

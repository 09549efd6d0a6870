// ccrs_oracle.hpp — CPU ORACLE (test infrastructure, NOT product code).
//
// A plain C++17 restatement of the reference's hot path, the way the reference
// computes it: forward-mode dual numbers pushed through
//     Isometry3::new(tvec, rvec) -> transform point -> project_one -> subtract
// (reference: src/optimization/factors.rs:152-173 ReprojectionFactor::residual_func,
//  :204-228 OtherCamReprojectionFactor::residual_func, src/types.rs:27-29,74-78).
//
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
// legs may load this. The shipped library (camera-intrinsic-calibration-rs_b200/csrc)
// never includes, links or calls anything in oracle/.
//
// PARITY STATUS ("partially pinned"): the arithmetic of this path lives in crates that
// are NOT vendored in /root/reference and cannot be built here (no cargo/rustc):
//   tiny-solver ^0.18.0 (Cargo.toml:44), camera-intrinsic-model ^0.8.0 (Cargo.toml:25),
//   nalgebra ^0.34.1 (Cargo.toml:35), num-dual (via tiny-solver).
// What IS pinned (tests/test_oracle_pins.py):
//   * reference tests/optimization_test.rs:36-80 (UCM residual ~0 at GT, >1e-3 perturbed),
//     tests/util_test.rs:77-110 (parameter order, UCM == EUCM(beta=1) == EUCMT(beta=1,t=0)),
//     tests/types_test.rs:5-20 (axis-angle <-> isometry round trip);
//   * KB4 and OPENCV5 values AND Jacobians against OpenCV 4.13 (cv2.fisheye.projectPoints,
//     cv2.projectPoints) — an independent third party with the same axis-angle convention;
//   * all six models' Jacobians against 50-digit mpmath differentiation.
// What is parity-UNPINNED (restated from the crates' published algorithms; every
// such assumption is a named constant below): EUCMT tangential form, FTHETA polynomial,
// KB4 small-r branch, tiny-solver's stop thresholds / error metric / fixed-variable and
// bounds semantics, and all LM constants (the reference never instantiates LM).
#pragma once
#include <cmath>
#include <cstdint>
#include <limits>

namespace ccrs_oracle {

// ---------------------------------------------------------------------------------
// Named assumptions about the un-vendored crates.
// ---------------------------------------------------------------------------------
constexpr int    kMaxIteration            = 100;    // tiny-solver OptimizerOptions::default().max_iteration
constexpr double kMinAbsErrDecrease       = 1e-5;   // .min_abs_error_decrease_threshold
constexpr double kMinRelErrDecrease       = 1e-5;   // .min_rel_error_decrease_threshold
constexpr double kMinErrThreshold         = 1e-10;  // .min_error_threshold
constexpr bool   kErrorIsL2Norm           = true;   // GN/LM compare ||r|| (norm_l2), not ||r||^2
constexpr double kLmMinDiagonal           = 1e-6;   // LM DEFAULT_MIN_DIAGONAL
constexpr double kLmMaxDiagonal           = 1e32;   // LM DEFAULT_MAX_DIAGONAL
constexpr double kLmInitialRadius         = 1e4;    // LM DEFAULT_INITIAL_TRUST_REGION_RADIUS (u0 = 1/radius)
constexpr double kLmRejectFactor0         = 2.0;    // v0: on reject u *= v; v *= 2 (Ceres/Nielsen)
constexpr double kKb4SmallR               = 1e-8;   // KB4/FTHETA: r below this uses the pinhole limit x/z, y/z
// nalgebra Quaternion::exp_eps: ||axisangle/2||^2 <= eps^2 returns the identity quaternion
// (with ZERO derivative under autodiff). eps = f64::EPSILON.
constexpr double kQuatExpEps              = std::numeric_limits<double>::epsilon();

enum Model { UCM = 0, EUCM = 1, EUCMT = 2, KB4 = 3, OPENCV5 = 4, FTHETA = 5 };

inline int model_nparams(int m) {
  switch (m) {
    case UCM: return 5;      // fx fy cx cy alpha            (factors.rs:103-107, optimization_test.rs:41)
    case EUCM: return 6;     // fx fy cx cy alpha beta       (data/eucm.json, util_test.rs:84,107-109)
    case EUCMT: return 8;    // fx fy cx cy alpha beta t1 t2 (util.rs:236-241)
    case KB4: return 8;      // fx fy cx cy k1 k2 k3 k4      (README.md:80 "OpenCV fisheye")
    case OPENCV5: return 9;  // fx fy cx cy k1 k2 p1 p2 k3   (README.md:81 "plumb_bob")
    case FTHETA: return 8;   // fx fy cx cy k1 k2 k3 k4      (unpinned)
  }
  return -1;
}

// ---------------------------------------------------------------------------------
// Forward-mode dual number with a runtime number of partials (num-dual's DualDVec64
// is the reference's carrier; a fixed-capacity array avoids its heap traffic, so this
// oracle is FASTER than the real reference and speed-ups quoted against it are conservative).
// ---------------------------------------------------------------------------------
constexpr int kMaxPartials = 24;  // 9 intrinsics + 12 pose scalars fits

struct Dual {
  double v;
  double d[kMaxPartials];
  int n;
  Dual() : v(0.0), n(0) {}
  Dual(double c, int n_) : v(c), n(n_) { for (int i = 0; i < n; ++i) d[i] = 0.0; }
  static Dual var(double c, int n_, int idx) { Dual r(c, n_); r.d[idx] = 1.0; return r; }
};
inline Dual operator+(const Dual& a, const Dual& b) { Dual r; r.n = a.n; r.v = a.v + b.v; for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] + b.d[i]; return r; }
inline Dual operator-(const Dual& a, const Dual& b) { Dual r; r.n = a.n; r.v = a.v - b.v; for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] - b.d[i]; return r; }
inline Dual operator-(const Dual& a) { Dual r; r.n = a.n; r.v = -a.v; for (int i = 0; i < a.n; ++i) r.d[i] = -a.d[i]; return r; }
inline Dual operator*(const Dual& a, const Dual& b) { Dual r; r.n = a.n; r.v = a.v * b.v; for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] * b.v + a.v * b.d[i]; return r; }
inline Dual operator/(const Dual& a, const Dual& b) {
  Dual r; r.n = a.n; const double inv = 1.0 / b.v; r.v = a.v * inv;
  for (int i = 0; i < a.n; ++i) r.d[i] = (a.d[i] * b.v - a.v * b.d[i]) * inv * inv;
  return r;
}
inline Dual operator+(const Dual& a, double c) { Dual r = a; r.v += c; return r; }
inline Dual operator-(const Dual& a, double c) { Dual r = a; r.v -= c; return r; }
inline Dual operator*(const Dual& a, double c) { Dual r = a; r.v *= c; for (int i = 0; i < a.n; ++i) r.d[i] *= c; return r; }
inline Dual operator/(const Dual& a, double c) { Dual r = a; r.v /= c; for (int i = 0; i < a.n; ++i) r.d[i] /= c; return r; }
inline Dual operator*(double c, const Dual& a) { return a * c; }
inline Dual operator+(double c, const Dual& a) { return a + c; }
inline Dual operator-(double c, const Dual& a) { return (-a) + c; }
inline Dual sqrt(const Dual& a) { Dual r; r.n = a.n; r.v = std::sqrt(a.v); const double k = 0.5 / r.v; for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] * k; return r; }
inline Dual sin(const Dual& a) { Dual r; r.n = a.n; r.v = std::sin(a.v); const double k = std::cos(a.v); for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] * k; return r; }
inline Dual cos(const Dual& a) { Dual r; r.n = a.n; r.v = std::cos(a.v); const double k = -std::sin(a.v); for (int i = 0; i < a.n; ++i) r.d[i] = a.d[i] * k; return r; }
inline Dual atan2(const Dual& y, const Dual& x) {
  Dual r; r.n = y.n; r.v = std::atan2(y.v, x.v); const double k = 1.0 / (x.v * x.v + y.v * y.v);
  for (int i = 0; i < y.n; ++i) r.d[i] = (x.v * y.d[i] - y.v * x.d[i]) * k;
  return r;
}
inline double value(const Dual& a) { return a.v; }
inline double value(double a) { return a; }
inline double sqrt_(double a) { return std::sqrt(a); }

// scalar-type helpers so the same templates run with T=double (residual only) and T=Dual
template <class T> struct Ops;
template <> struct Ops<double> {
  static double c(double v, const double&) { return v; }
  static double sqrt(double a) { return std::sqrt(a); }
  static double sin(double a) { return std::sin(a); }
  static double cos(double a) { return std::cos(a); }
  static double atan2(double y, double x) { return std::atan2(y, x); }
};
template <> struct Ops<Dual> {
  static Dual c(double v, const Dual& like) { return Dual(v, like.n); }
  static Dual sqrt(const Dual& a) { return ccrs_oracle::sqrt(a); }
  static Dual sin(const Dual& a) { return ccrs_oracle::sin(a); }
  static Dual cos(const Dual& a) { return ccrs_oracle::cos(a); }
  static Dual atan2(const Dual& y, const Dual& x) { return ccrs_oracle::atan2(y, x); }
};

// ---------------------------------------------------------------------------------
// nalgebra restatement: UnitQuaternion::from_scaled_axis + rotate + translate.
// q = exp(Quaternion::from_imag(axisangle / 2));   (nalgebra geometry/quaternion_construction.rs)
// p' = p + w*t + v x t,  t = 2 (v x p)               (nalgebra geometry/quaternion_ops.rs)
// Reference call sites: factors.rs:160-164, :212-218; types.rs:27-29.
// ---------------------------------------------------------------------------------
template <class T> struct Quat { T w, x, y, z; };

template <class T> Quat<T> quat_from_scaled_axis(const T rv[3]) {
  using O = Ops<T>;
  T hx = rv[0] / 2.0, hy = rv[1] / 2.0, hz = rv[2] / 2.0;
  T nn = hx * hx + hy * hy + hz * hz;
  if (value(nn) <= kQuatExpEps * kQuatExpEps) {
    // identity with zero derivative — what nalgebra's exp_eps returns (SURVEY App. A edge case)
    return Quat<T>{O::c(1.0, nn), O::c(0.0, nn), O::c(0.0, nn), O::c(0.0, nn)};
  }
  T n = O::sqrt(nn);
  T s = O::sin(n) / n;  // w_exp (= exp(0) = 1) * sin(n) / n
  return Quat<T>{O::cos(n), hx * s, hy * s, hz * s};
}

template <class T> void quat_rotate(const Quat<T>& q, const T p[3], T out[3]) {
  // t = (v x p) * 2 ; out = t*w + v x t + p
  T tx = (q.y * p[2] - q.z * p[1]) * 2.0;
  T ty = (q.z * p[0] - q.x * p[2]) * 2.0;
  T tz = (q.x * p[1] - q.y * p[0]) * 2.0;
  T cx = q.y * tz - q.z * ty;
  T cy = q.z * tx - q.x * tz;
  T cz = q.x * ty - q.y * tx;
  out[0] = tx * q.w + cx + p[0];
  out[1] = ty * q.w + cy + p[1];
  out[2] = tz * q.w + cz + p[2];
}

template <class T> Quat<T> quat_mul(const Quat<T>& a, const Quat<T>& b) {
  // Hamilton product, nalgebra's operand order
  Quat<T> r;
  r.w = a.w * b.w - a.x * b.x - a.y * b.y - a.z * b.z;
  r.x = a.w * b.x + a.x * b.w + a.y * b.z - a.z * b.y;
  r.y = a.w * b.y - a.x * b.z + a.y * b.w + a.z * b.x;
  r.z = a.w * b.z + a.x * b.y - a.y * b.x + a.z * b.w;
  return r;
}

// Isometry3::new(tvec, rvec) * point  (factors.rs:162-163)
template <class T> void isometry_apply(const T rv[3], const T tv[3], const T p[3], T out[3]) {
  Quat<T> q = quat_from_scaled_axis(rv);
  T r[3];
  quat_rotate(q, p, r);
  out[0] = r[0] + tv[0]; out[1] = r[1] + tv[1]; out[2] = r[2] + tv[2];
}

// (T_i_0 * T_0_b) * point — isometry composition first, like the reference (factors.rs:214-218)
template <class T> void isometry_chain_apply(const T rv1[3], const T tv1[3],   // T_i_0
                                             const T rv0[3], const T tv0[3],   // T_0_b
                                             const T p[3], T out[3]) {
  Quat<T> q1 = quat_from_scaled_axis(rv1);
  Quat<T> q0 = quat_from_scaled_axis(rv0);
  // Isometry * Isometry: t = t1 + R1 * t0 ; q = q1 * q0
  T rt0[3];
  quat_rotate(q1, tv0, rt0);
  T t[3] = {tv1[0] + rt0[0], tv1[1] + rt0[1], tv1[2] + rt0[2]};
  Quat<T> q = quat_mul(q1, q0);
  T r[3];
  quat_rotate(q, p, r);
  out[0] = r[0] + t[0]; out[1] = r[1] + t[1]; out[2] = r[2] + t[2];
}

// ---------------------------------------------------------------------------------
// camera-intrinsic-model ^0.8 restatement: GenericModel::project_one for six models.
// `prm` is the FULL parameter vector (fy already re-inserted, factors.rs:156-158).
// ---------------------------------------------------------------------------------
template <class T> void project_one(int model, const T* prm, const T P[3], T uv[2]) {
  using O = Ops<T>;
  const T& fx = prm[0]; const T& fy = prm[1]; const T& cx = prm[2]; const T& cy = prm[3];
  const T& x = P[0]; const T& y = P[1]; const T& z = P[2];
  switch (model) {
    case UCM: case EUCM: case EUCMT: {
      const T& alpha = prm[4];
      T r2 = x * x + y * y;
      T rho2 = (model == UCM) ? (r2 + z * z) : (prm[5] * r2 + z * z);
      T rho = O::sqrt(rho2);
      T norm = alpha * rho + (1.0 - alpha) * z;
      T mx = x / norm, my = y / norm;
      if (model == EUCMT) {
        // EUCM followed by plumb-bob tangential terms (unpinned)
        const T& t1 = prm[6]; const T& t2 = prm[7];
        T rr = mx * mx + my * my;
        T xd = mx + 2.0 * t1 * mx * my + t2 * (rr + 2.0 * mx * mx);
        T yd = my + t1 * (rr + 2.0 * my * my) + 2.0 * t2 * mx * my;
        mx = xd; my = yd;
      }
      uv[0] = fx * mx + cx; uv[1] = fy * my + cy;
      return;
    }
    case KB4: case FTHETA: {
      T r2 = x * x + y * y;
      T r = O::sqrt(r2);
      if (value(r) < kKb4SmallR) {  // pinhole limit (d(theta)/r -> 1/z); derivative is that of x/z
        uv[0] = fx * (x / z) + cx; uv[1] = fy * (y / z) + cy;
        return;
      }
      T th = O::atan2(r, z);
      T d;
      if (model == KB4) {
        T th2 = th * th;
        // theta (1 + k1 th^2 + k2 th^4 + k3 th^6 + k4 th^8), Horner in th^2
        d = th * (1.0 + th2 * (prm[4] + th2 * (prm[5] + th2 * (prm[6] + th2 * prm[7]))));
      } else {
        // FTHETA (unpinned): theta (1 + k1 th + k2 th^2 + k3 th^3 + k4 th^4)
        d = th * (1.0 + th * (prm[4] + th * (prm[5] + th * (prm[6] + th * prm[7]))));
      }
      uv[0] = fx * (d * x / r) + cx; uv[1] = fy * (d * y / r) + cy;
      return;
    }
    case OPENCV5: {
      const T& k1 = prm[4]; const T& k2 = prm[5]; const T& p1 = prm[6]; const T& p2 = prm[7]; const T& k3 = prm[8];
      T a = x / z, b = y / z;
      T r2 = a * a + b * b;
      T radial = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3));
      T xd = a * radial + 2.0 * p1 * a * b + p2 * (r2 + 2.0 * a * a);
      T yd = b * radial + p1 * (r2 + 2.0 * b * b) + 2.0 * p2 * a * b;
      uv[0] = fx * xd + cx; uv[1] = fy * yd + cy;
      return;
    }
  }
}

// ---------------------------------------------------------------------------------
// tiny-solver restatement: HuberLoss::evaluate + Corrector (Ceres-style).
// rho = [rho(s), rho'(s), rho''(s)], s = ||r||^2. Huber has rho'' <= 0 everywhere, so the
// corrector always takes the simple branch: r *= sqrt(rho'), J *= sqrt(rho').
// Reference call site: util.rs:413 `HuberLoss::new(1.0)`.
// ---------------------------------------------------------------------------------
inline double huber_sqrt_rho1(double s, double delta) {
  if (delta <= 0.0) return 1.0;             // loss disabled
  if (s > delta * delta) {
    const double r = std::sqrt(s);
    return std::sqrt(delta / r);
  }
  return 1.0;
}

}  // namespace ccrs_oracle

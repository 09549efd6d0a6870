// ccrs_device.cuh — device-side math of the linearisation path: analytic (not autodiff) projection
// and Jacobians of the six camera models, axis-angle pose math, packed block indexing.
//
// What this replaces in the reference (per observation, per iteration):
//   ReprojectionFactor::residual_func evaluated with num-dual duals  (src/optimization/factors.rs:152-173)
//   = Isometry3::new(tvec, rvec) * p3d  ->  GenericModel::project_one  ->  - p2d
// Every model has the form  u = fx*mx + cx,  v = fy*my + cy  with (mx,my) independent of fx,fy,cx,cy
// (SURVEY.md App. A), so a model only supplies m(P;k), dm/dP (2x3) and dm/dk (2xND).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#define CCRS_HD __host__ __device__ __forceinline__
#define CCRS_D __device__ __forceinline__

namespace ccrs {

enum : int { UCM = 0, EUCM = 1, EUCMT = 2, KB4 = 3, OPENCV5 = 4, FTHETA = 5 };

CCRS_HD constexpr int model_nd(int m) {  // number of distortion parameters after fx,fy,cx,cy
  return m == UCM ? 1 : m == EUCM ? 2 : m == OPENCV5 ? 5 : 4;
}
constexpr int kMaxNd = 5;
constexpr int kMaxFull = 9;             // fx fy cx cy + 5
constexpr double kSmallR = 1e-8;        // KB4/FTHETA pinhole-limit branch (oracle: kKb4SmallR)
constexpr double kQuatExpEps = 2.220446049250313e-16;  // nalgebra Quaternion::exp_eps threshold (oracle: kQuatExpEps)

// packed upper-triangular index of an NA x NA symmetric matrix, i <= j
CCRS_HD constexpr int tri_idx(int NA, int i, int j) { return i * NA - (i * (i - 1)) / 2 + (j - i); }

// ---------------------------------------------------------------------------------------------
// Pose: R(rvec) and the left Jacobian J_l(rvec) of SO(3).
//   d(R(w) p)/dw = -[R p]x J_l(w)          (so the kernels accumulate in the local basis
//   d(R p)/dphi = -[R p]x and change basis once per frame-slice with J_l).
// nalgebra builds R from the unit quaternion exp(w/2); for ||w/2||^2 <= eps^2 it returns the identity
// with zero derivative under autodiff — reproduced here (J_l = 0) so the GPU matches the reference
// at an exactly-zero rvec (tests/optimization_test.rs:59).
// ---------------------------------------------------------------------------------------------
struct FramePose {
  double R[9];   // row-major
  double t[3];
  double Jl[9];  // row-major
};

CCRS_HD void pose_from_rvec_tvec(const double* rt, FramePose& fp) {
  const double wx = rt[0], wy = rt[1], wz = rt[2];
  const double th2 = wx * wx + wy * wy + wz * wz;
  double a, b, c;  // sin(th)/th, (1-cos th)/th^2, (th - sin th)/th^3
  if (th2 < 1e-4) {
    // Taylor series, truncation error < 1e-22 relative for th^2 < 1e-4
    a = 1.0 - th2 * (1.0 / 6.0 - th2 * (1.0 / 120.0 - th2 * (1.0 / 5040.0 - th2 / 362880.0)));
    b = 0.5 - th2 * (1.0 / 24.0 - th2 * (1.0 / 720.0 - th2 * (1.0 / 40320.0 - th2 / 3628800.0)));
    c = 1.0 / 6.0 - th2 * (1.0 / 120.0 - th2 * (1.0 / 5040.0 - th2 * (1.0 / 362880.0 - th2 / 39916800.0)));
  } else {
    const double th = sqrt(th2);
    double sh, ch;
    sincos(0.5 * th, &sh, &ch);
    const double ith = 1.0 / th, s2 = 2.0 * sh * ch;   // one division instead of three (each ~127 cycles of latency)
    a = s2 * ith;
    b = 2.0 * sh * sh * (ith * ith);
    c = (th - s2) * (ith * ith * ith);
  }
  // K = [w]x, K^2 = w w^T - th2 I
  const double xx = wx * wx, yy = wy * wy, zz = wz * wz, xy = wx * wy, xz = wx * wz, yz = wy * wz;
  fp.R[0] = 1.0 - b * (yy + zz); fp.R[1] = -a * wz + b * xy;     fp.R[2] = a * wy + b * xz;
  fp.R[3] = a * wz + b * xy;     fp.R[4] = 1.0 - b * (xx + zz); fp.R[5] = -a * wx + b * yz;
  fp.R[6] = -a * wy + b * xz;    fp.R[7] = a * wx + b * yz;     fp.R[8] = 1.0 - b * (xx + yy);
  const bool identity_branch = (0.25 * th2 <= kQuatExpEps * kQuatExpEps);
  if (identity_branch) {
    for (int i = 0; i < 9; ++i) fp.Jl[i] = 0.0;
    fp.R[0] = fp.R[4] = fp.R[8] = 1.0;
    fp.R[1] = fp.R[2] = fp.R[3] = fp.R[5] = fp.R[6] = fp.R[7] = 0.0;
  } else {
    fp.Jl[0] = 1.0 - c * (yy + zz); fp.Jl[1] = -b * wz + c * xy;     fp.Jl[2] = b * wy + c * xz;
    fp.Jl[3] = b * wz + c * xy;     fp.Jl[4] = 1.0 - c * (xx + zz); fp.Jl[5] = -b * wx + c * yz;
    fp.Jl[6] = -b * wy + c * xz;    fp.Jl[7] = b * wx + c * yz;     fp.Jl[8] = 1.0 - c * (xx + yy);
  }
  fp.t[0] = rt[3]; fp.t[1] = rt[4]; fp.t[2] = rt[5];
}

// ---------------------------------------------------------------------------------------------
// Branch-free FP64 reciprocal / rsqrt / sqrt for operands known to be normal, finite and positive (depths,
// norms, squared residuals above the Huber threshold). CUDA's `1.0/x`, `rsqrt`, `sqrt` wrap the same MUFU seed in a
// special-case test plus a slow-path call, which splits the observation loop into many basic blocks and stops
// ptxas from overlapping the dependent model chain with the independent Gram-block DFMAs. These are single
// straight-line sequences (MUFU seed, relative error 2^-22, one cubically convergent step, one correction):
// error <= ~1 ulp.
// ---------------------------------------------------------------------------------------------
CCRS_D double rcp_fast(double x) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  double e = fma(-x, y, 1.0);
  e = fma(e, e, e);          // e + e^2  ->  y (1 + e + e^2) = (1/x)(1 - e^3)
  y = fma(y, e, y);
  e = fma(-x, y, 1.0);       // correction step
  return fma(y, e, y);
}
CCRS_D double rsqrt_fast(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double t = x * y;
  const double e = fma(-t, y, 1.0);            // 1 - x y^2
  const double p = fma(0.375, e, 0.5) * e;     // e/2 + 3 e^2/8  ->  error O(e^3)
  return fma(y, p, y);
}
CCRS_D double sqrt_fast(double x) {
  const double y = rsqrt_fast(x);
  const double s = x * y;
  const double r = fma(-s, s, x);              // one Newton correction of s ~ sqrt(x)
  return fma(r, 0.5 * y, s);
}

// ---------------------------------------------------------------------------------------------
// theta = atan2(r, z) for r > 0 (the angle of a point off the optical axis: KB4 / FTHETA), straight-line code instead of
// the library's atan2 (range tests, a division with a slow path, a 19-term polynomial: ~2.4x the instructions and a
// chain several times as deep). With rho = |(r, z)| given as 1/rho:
//   key   t = tan(theta'/2) = r / (rho + |z|) in [0, 1], evaluated in FP32 from FP32 square roots; i = round(64 t)
//   table theta_i = 2 atan(i/64), sin theta_i, cos theta_i  (ccrs_atan_tab.inc, 65 entries)
//   sin(theta' - theta_i) = (r cos theta_i - |z| sin theta_i) / rho,  |theta' - theta_i| <= 1/64 + FP32 error
//   theta' = theta_i + asin(.)  (odd series to x^9: truncation < 3e-20),  theta = z < 0 ? pi - theta' : theta'
// Absolute error a few 1e-16 (<= 9e-16 = 2 ulp of pi over two million random (r, z): the roundings of the table, of the
// sum and of pi - theta; tests/test_atan_table.py), i.e. <= 3e-14 relative for theta >= 1/64 and full relative precision
// below 1/128 (entry 0 is exact).
// ---------------------------------------------------------------------------------------------
static __device__ const double kAtanTab[65][4] = {
#include "ccrs_atan_tab.inc"
};
CCRS_D double atan2_pos(double r, double z, double r2, double rho2, double irho) {
  const double za = fabs(z);
  float rf, rhof, inv;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rf) : "f"((float)r2));
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rhof) : "f"((float)rho2));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(rhof + (float)za));
  const unsigned i = min((unsigned)__float2int_rn(rf * inv * 64.0f), 64u);
  const double2 e0 = *reinterpret_cast<const double2*>(&kAtanTab[i][0]);   // theta_i, sin
  const double ci = kAtanTab[i][2];
  const double xs = fma(r, ci, -(za * e0.y)) * irho;
  const double x2 = xs * xs;
  double p = fma(x2, 35.0 / 1152.0, 5.0 / 112.0);
  p = fma(x2, p, 3.0 / 40.0);
  p = fma(x2, p, 1.0 / 6.0);
  const double thp = e0.x + fma(xs * x2, p, xs);
  return z < 0.0 ? 3.141592653589793 - thp : thp;
}

// ---------------------------------------------------------------------------------------------
// Camera models. k = distortion parameters (params[4..]). WITH_J selects value-only evaluation.
// Outputs: m[2]; dP[2][3] = dm/dP; dk[2][ND] = dm/dk.
// ---------------------------------------------------------------------------------------------
template <int MODEL, bool WITH_J>
CCRS_D void model_eval(const double* __restrict__ k, double x, double y, double z,
                       double m[2], double dP[2][3], double dk[2][kMaxNd]) {
  if constexpr (MODEL == UCM || MODEL == EUCM || MODEL == EUCMT) {
    const double alpha = k[0];
    const double beta = (MODEL == UCM) ? 1.0 : k[1];
    const double r2 = x * x + y * y;
    const double rho2 = fma(beta, r2, z * z);
    const double irho = rsqrt_fast(rho2);
    const double rho = rho2 * irho;
    const double oma = 1.0 - alpha;
    const double nrm = fma(alpha, rho, oma * z);
    const double in = rcp_fast(nrm);
    double mx = x * in, my = y * in;
    if constexpr (!WITH_J) {
      if constexpr (MODEL == EUCMT) {
        const double t1 = k[2], t2 = k[3];
        const double rr = mx * mx + my * my;
        const double xd = mx + 2.0 * t1 * mx * my + t2 * (rr + 2.0 * mx * mx);
        const double yd = my + t1 * (rr + 2.0 * my * my) + 2.0 * t2 * mx * my;
        mx = xd; my = yd;
      }
      m[0] = mx; m[1] = my;
      return;
    } else {
      const double ab_irho = alpha * beta * irho;
      const double nx = ab_irho * x, ny = ab_irho * y, nz = fma(alpha * irho, z, oma);
      const double mxin = mx * in, myin = my * in;
      double e[2][3];
      e[0][0] = in - mxin * nx; e[0][1] = -mxin * ny;     e[0][2] = -mxin * nz;
      e[1][0] = -myin * nx;     e[1][1] = in - myin * ny; e[1][2] = -myin * nz;
      const double na = rho - z;                 // dn/dalpha
      const double nb = 0.5 * alpha * r2 * irho; // dn/dbeta
      double ea[2] = {-mxin * na, -myin * na};
      double eb[2] = {-mxin * nb, -myin * nb};
      if constexpr (MODEL == EUCMT) {
        const double t1 = k[2], t2 = k[3];
        const double rr = mx * mx + my * my;
        const double mxy2 = 2.0 * mx * my;
        const double xd = mx + t1 * mxy2 + t2 * (rr + 2.0 * mx * mx);
        const double yd = my + t1 * (rr + 2.0 * my * my) + t2 * mxy2;
        const double t00 = 1.0 + 2.0 * t1 * my + 6.0 * t2 * mx;
        const double t01 = 2.0 * (t1 * mx + t2 * my);
        const double t11 = 1.0 + 6.0 * t1 * my + 2.0 * t2 * mx;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
          dP[0][j] = t00 * e[0][j] + t01 * e[1][j];
          dP[1][j] = t01 * e[0][j] + t11 * e[1][j];
        }
        dk[0][0] = t00 * ea[0] + t01 * ea[1]; dk[1][0] = t01 * ea[0] + t11 * ea[1];
        dk[0][1] = t00 * eb[0] + t01 * eb[1]; dk[1][1] = t01 * eb[0] + t11 * eb[1];
        dk[0][2] = mxy2;                  dk[1][2] = rr + 2.0 * my * my;
        dk[0][3] = rr + 2.0 * mx * mx;    dk[1][3] = mxy2;
        m[0] = xd; m[1] = yd;
      } else {
#pragma unroll
        for (int j = 0; j < 3; ++j) { dP[0][j] = e[0][j]; dP[1][j] = e[1][j]; }
        dk[0][0] = ea[0]; dk[1][0] = ea[1];
        if constexpr (MODEL == EUCM) { dk[0][1] = eb[0]; dk[1][1] = eb[1]; }
        m[0] = mx; m[1] = my;
      }
    }
  } else if constexpr (MODEL == KB4 || MODEL == FTHETA) {
    const double r2 = fma(x, x, y * y);
    const double r2c = fmax(r2, 1e-300);            // branch-free; r2 = 0 lands in the pinhole-limit branch below
    const double ir = rsqrt_fast(r2c);
    const double r = r2c * ir;
    if (r < kSmallR) {
      const double iz = 1.0 / z;
      m[0] = x * iz; m[1] = y * iz;
      if constexpr (WITH_J) {
        dP[0][0] = iz; dP[0][1] = 0.0; dP[0][2] = -m[0] * iz;
        dP[1][0] = 0.0; dP[1][1] = iz; dP[1][2] = -m[1] * iz;
#pragma unroll
        for (int j = 0; j < 4; ++j) { dk[0][j] = 0.0; dk[1][j] = 0.0; }
      }
      return;
    }
    const double rho2 = fma(z, z, r2);
    const double irho = rsqrt_fast(rho2);
    const double th = atan2_pos(r, z, r2c, rho2, irho);
    double d, dd, pw[4];  // d(theta), d'(theta), d d/d k_i
    if constexpr (MODEL == KB4) {
      const double t2 = th * th;
      pw[0] = th * t2; pw[1] = pw[0] * t2; pw[2] = pw[1] * t2; pw[3] = pw[2] * t2;
      d = th * (1.0 + t2 * (k[0] + t2 * (k[1] + t2 * (k[2] + t2 * k[3]))));
      dd = 1.0 + t2 * (3.0 * k[0] + t2 * (5.0 * k[1] + t2 * (7.0 * k[2] + t2 * 9.0 * k[3])));
    } else {
      pw[0] = th * th; pw[1] = pw[0] * th; pw[2] = pw[1] * th; pw[3] = pw[2] * th;
      d = th * (1.0 + th * (k[0] + th * (k[1] + th * (k[2] + th * k[3]))));
      dd = 1.0 + th * (2.0 * k[0] + th * (3.0 * k[1] + th * (4.0 * k[2] + th * 5.0 * k[3])));
    }
    const double s = d * ir;  // m = s * (x, y)
    m[0] = x * s; m[1] = y * s;
    if constexpr (WITH_J) {
      const double irho2 = irho * irho;
      // theta_x = x z / (r rho2), theta_y = y z / (r rho2), theta_z = -r / rho2
      const double cxy = (dd * z * irho2 - s) * ir * ir;  // (d' theta_x / r - d x / r^3) / x
      const double sz = -dd * irho2;                      // ds/dz = d' theta_z / r
      dP[0][0] = s + x * x * cxy; dP[0][1] = x * y * cxy;     dP[0][2] = x * sz;
      dP[1][0] = dP[0][1];        dP[1][1] = s + y * y * cxy; dP[1][2] = y * sz;
      const double xr = x * ir, yr = y * ir;
#pragma unroll
      for (int j = 0; j < 4; ++j) { dk[0][j] = xr * pw[j]; dk[1][j] = yr * pw[j]; }
    }
  } else {  // OPENCV5: k = k1 k2 p1 p2 k3
    const double k1 = k[0], k2 = k[1], p1 = k[2], p2 = k[3], k3 = k[4];
    const double iz = rcp_fast(z);
    const double a = x * iz, b = y * iz;
    const double a2 = a * a, b2 = b * b, ab = a * b;
    const double r2 = a2 + b2;
    const double rad = 1.0 + r2 * (k1 + r2 * (k2 + r2 * k3));
    m[0] = a * rad + 2.0 * p1 * ab + p2 * (r2 + 2.0 * a2);
    m[1] = b * rad + p1 * (r2 + 2.0 * b2) + 2.0 * p2 * ab;
    if constexpr (WITH_J) {
      const double drad = k1 + r2 * (2.0 * k2 + r2 * 3.0 * k3);  // d rad / d r2
      const double xa = rad + 2.0 * a2 * drad + 2.0 * p1 * b + 6.0 * p2 * a;
      const double xb = 2.0 * (ab * drad + p1 * a + p2 * b);
      const double yb = rad + 2.0 * b2 * drad + 6.0 * p1 * b + 2.0 * p2 * a;
      dP[0][0] = xa * iz; dP[0][1] = xb * iz; dP[0][2] = -(xa * a + xb * b) * iz;
      dP[1][0] = xb * iz; dP[1][1] = yb * iz; dP[1][2] = -(xb * a + yb * b) * iz;
      const double r4 = r2 * r2, r6 = r4 * r2;
      dk[0][0] = a * r2; dk[0][1] = a * r4; dk[0][2] = 2.0 * ab;        dk[0][3] = r2 + 2.0 * a2; dk[0][4] = a * r6;
      dk[1][0] = b * r2; dk[1][1] = b * r4; dk[1][2] = r2 + 2.0 * b2;   dk[1][3] = 2.0 * ab;      dk[1][4] = b * r6;
    }
  }
}

// s^(-1/4) for a normal, finite, positive s: FP32 seed (two MUFU.f32, relative error ~2^-21) and one cubically
// convergent FP64 step  y <- y (1 + e/4 + 5 e^2/32),  e = 1 - s y^4   (error ~e^3: below 1 ulp).
CCRS_D double rqrt4_fast(double s) {
  float y0;
  const float sf = (float)s;
  asm("{\n\t.reg .f32 t;\n\tsqrt.approx.ftz.f32 t, %1;\n\trsqrt.approx.ftz.f32 %0, t;\n\t}" : "=f"(y0) : "f"(sf));
  const double y = (double)y0;
  const double y2 = y * y;
  const double e = fma(-s, y2 * y2, 1.0);
  const double p = fma(0.15625, e, 0.25) * e;
  return fma(y, p, y);
}

// Reciprocal of x given an FP32-accurate seed y0 ~ 1/x: one quartically convergent step y0 (1 + e + e^2 + e^3),
// e = 1 - x y0 (error e^4: still below 1e-13 when cancellation in the FP32 evaluation of x degraded the seed to 1e-3.5).
CCRS_D double rcp_refine(double x, double y0) {
  const double e = fma(-x, y0, 1.0);
  return fma(y0, fma(fma(e, e, e), e, e), y0);
}

// Huber corrector (tiny-solver Corrector with HuberLoss, reference util.rs:413): sqrt(rho'(s)) = (delta^2 / s)^(1/4)
// outside the quadratic region. Branch-free: the weight is always evaluated (on a safe operand) and selected;
// sqrt(delta) is loop-invariant at every call site.
CCRS_D double huber_weight(double s, double delta) {
  const bool out = delta > 0.0 && s > delta * delta;
  const double sd = sqrt_fast(delta > 0.0 ? delta : 1.0);
  const double w = sd * rqrt4_fast(out ? s : 1.0);
  return out ? w : 1.0;
}

}  // namespace ccrs

// ccrs_pnp.cu — batched initial board poses: one warp per frame.
//
// Replaces the per-frame `sqpnp_simple::sqpnp_solve_glam(&p3ds, &p2ds_z)` of calib_camera (src/util.rs:418-439) and of
// init_pose (src/optimization/linear.rs:5-21): the step immediately before the hot path, serial in the reference and
// the bottleneck once an LM iteration costs microseconds.
//
// Same objective as SQPnP (Terzakis & Lourakis, ECCV 2020; crate sqpnp_simple 0.2.0, un-vendored): with normalised image
// points m_i = (x_i, y_i, 1), Q_i = [1 0 -x; 0 1 -y; -x -y x^2+y^2] and r = vec(R) (row-major),
//     cost(R, t) = sum_i (R p_i + t)^T Q_i (R p_i + t),   t = P r,   cost = r^T Omega r,
//     P = -(sum Q_i)^-1 sum Q_i A_i,   Omega = sum A_i^T Q_i A_i + (sum Q_i A_i)^T P,   A_i = I_3 (x) p_i^T.
// Everything is a moment sum over the frame's points: sum q {1, p_c, p_c p_d} for q in {1, x, y, x^2+y^2} — 40 sums.
// Different solver, chosen for the GPU: instead of SQPnP's eigen-decomposition of Omega + sequential quadratic programming
// from a few eigenvector starts, the 32 lanes of the warp run damped Newton on SO(3) (exact 3x3 manifold Hessian) from
// 32 well-spread rotations (the 24 cube rotations + 8 sixty-degree turns about the body diagonals) in lock step, and
// the warp keeps the lowest-cost solution with the board in front of the camera. Same global minimiser, no 9x9
// eigen-solver, no divergence. Fixed reduction trees and lane-order tie-breaks: bitwise reproducible.
#include "ccrs_kernels.cuh"

#include <math.h>

namespace ccrs {

constexpr int kPnpWarps = 4;
constexpr int kPnpIters = 40;

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;   // identical in every lane: each level adds the same two values in both partners
}

// R <- Exp(d) R  (Rodrigues; series below 1e-4 rad)
__device__ __forceinline__ void so3_left_update(const double d[3], const double R[9], double out[9]) {
  const double th2 = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
  double a, b;
  if (th2 < 1e-8) { a = 1.0 - th2 / 6.0; b = 0.5 - th2 / 24.0; }
  else { const double th = sqrt(th2); double s, c; sincos(th, &s, &c); a = s / th; b = (1.0 - c) / th2; }
  // E = I + a K + b K^2, K = [d]x, K^2 = d d^T - th2 I
  double E[9];
  E[0] = 1.0 - b * (d[1] * d[1] + d[2] * d[2]); E[1] = -a * d[2] + b * d[0] * d[1];        E[2] = a * d[1] + b * d[0] * d[2];
  E[3] = a * d[2] + b * d[0] * d[1];        E[4] = 1.0 - b * (d[0] * d[0] + d[2] * d[2]); E[5] = -a * d[0] + b * d[1] * d[2];
  E[6] = -a * d[1] + b * d[0] * d[2];       E[7] = a * d[0] + b * d[1] * d[2];        E[8] = 1.0 - b * (d[0] * d[0] + d[1] * d[1]);
#pragma unroll
  for (int i = 0; i < 3; ++i)
#pragma unroll
    for (int j = 0; j < 3; ++j) out[3 * i + j] = E[3 * i] * R[j] + E[3 * i + 1] * R[3 + j] + E[3 * i + 2] * R[6 + j];
}

// w = Omega r (Omega row-major in shared memory, broadcast reads); returns r . w
__device__ __forceinline__ double quad_form(const double* __restrict__ Om, const double r[9], double w[9]) {
  double c = 0.0;
#pragma unroll
  for (int i = 0; i < 9; ++i) {
    double s = 0.0;
#pragma unroll
    for (int j = 0; j < 9; ++j) s = fma(Om[9 * i + j], r[j], s);
    w[i] = s;
    c = fma(r[i], s, c);
  }
  return c;
}

// rotation matrix -> axis-angle (the inverse of Isometry3::new's rotation argument, types.rs:58-64)
__device__ void rvec_from_R(const double R[9], double rv[3]) {
  const double vx = 0.5 * (R[7] - R[5]), vy = 0.5 * (R[2] - R[6]), vz = 0.5 * (R[3] - R[1]);
  const double s = sqrt(vx * vx + vy * vy + vz * vz);              // sin(theta)
  const double c = fmin(fmax(0.5 * (R[0] + R[4] + R[8] - 1.0), -1.0), 1.0);
  const double th = atan2(s, c);
  if (s < 1e-12 && c > 0.0) { rv[0] = vx; rv[1] = vy; rv[2] = vz; return; }   // theta -> 0: rvec = v (1 + O(theta^2))
  if (c > -0.99) { const double k = th / s; rv[0] = vx * k; rv[1] = vy * k; rv[2] = vz * k; return; }
  // theta near pi: the skew part vanishes; take the axis from the symmetric part R + R^T = 2 c I + 2 (1 - c) a a^T
  const double d0 = (R[0] - c) / (1.0 - c), d1 = (R[4] - c) / (1.0 - c), d2 = (R[8] - c) / (1.0 - c);
  double ax, ay, az;
  if (d0 >= d1 && d0 >= d2) { ax = sqrt(fmax(d0, 0.0)); ay = 0.5 * (R[1] + R[3]) / ((1.0 - c) * ax); az = 0.5 * (R[2] + R[6]) / ((1.0 - c) * ax); }
  else if (d1 >= d2)        { ay = sqrt(fmax(d1, 0.0)); ax = 0.5 * (R[1] + R[3]) / ((1.0 - c) * ay); az = 0.5 * (R[5] + R[7]) / ((1.0 - c) * ay); }
  else                      { az = sqrt(fmax(d2, 0.0)); ax = 0.5 * (R[2] + R[6]) / ((1.0 - c) * az); ay = 0.5 * (R[5] + R[7]) / ((1.0 - c) * az); }
  const double n = sqrt(ax * ax + ay * ay + az * az);
  double sg = (ax * vx + ay * vy + az * vz) < 0.0 ? -1.0 : 1.0;   // orient the axis along the (tiny) skew part
  const double k = sg * th / n;
  rv[0] = ax * k; rv[1] = ay * k; rv[2] = az * k;
}

// lane -> start rotation: 24 rotations of the cube, then 8 turns of 60 degrees about (+-1, +-1, +-1)
__device__ void start_rotation(int lane, double R[9]) {
  if (lane < 24) {
    // axis permutation perm (6) x sign pattern (4 of the 8 patterns give det = +1 for a given permutation parity)
    const int p = lane >> 2, q = lane & 3;
    const int perm[6][3] = {{0, 1, 2}, {1, 2, 0}, {2, 0, 1}, {0, 2, 1}, {2, 1, 0}, {1, 0, 2}};
    const bool odd_perm = p >= 3;
    // signs (s0, s1, s2) with product = +1 for even permutations, -1 for odd ones
    const int even_s[4][3] = {{1, 1, 1}, {1, -1, -1}, {-1, 1, -1}, {-1, -1, 1}};
    const int odd_s[4][3] = {{-1, 1, 1}, {1, -1, 1}, {1, 1, -1}, {-1, -1, -1}};
#pragma unroll
    for (int i = 0; i < 9; ++i) R[i] = 0.0;
    for (int i = 0; i < 3; ++i) R[3 * i + perm[p][i]] = odd_perm ? odd_s[q][i] : even_s[q][i];
  } else {
    const int k = lane - 24;
    const double a = 1.0471975511965976 / 1.7320508075688772;   // (pi/3) / sqrt(3)
    const double d[3] = {(k & 1) ? -a : a, (k & 2) ? -a : a, (k & 4) ? -a : a};
    const double I[9] = {1, 0, 0, 0, 1, 0, 0, 0, 1};
    so3_left_update(d, I, R);
  }
}

__global__ void __launch_bounds__(kPnpWarps * 32) k_pnp(const int32_t* __restrict__ frame_offsets, const double* __restrict__ x,
                                                        const double* __restrict__ y, const double* __restrict__ z,
                                                        const double* __restrict__ xn, const double* __restrict__ yn, int n_frames,
                                                        double* __restrict__ poses_out, double* __restrict__ cost_out) {
  __shared__ double s_om[kPnpWarps][81 + 27 + 40];   // Omega (9x9), P (3x9) and the 40 moment sums per warp
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * kPnpWarps + wid;
  if (f >= n_frames) return;
  const int beg = frame_offsets[f], end = frame_offsets[f + 1];
  // ---- 40 moment sums: S[q][m], q in {1, x, y, x^2 + y^2}, m in {1, px, py, pz, pxpx, pxpy, pxpz, pypy, pypz, pzpz}
  double S[4][10];
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int m = 0; m < 10; ++m) S[q][m] = 0.0;
  for (int k = beg + lane; k < end; k += 32) {
    const double px = x[k], py = y[k], pz = z[k], mx = xn[k], my = yn[k];
    const double mom[10] = {1.0, px, py, pz, px * px, px * py, px * pz, py * py, py * pz, pz * pz};
    const double qv[4] = {1.0, mx, my, mx * mx + my * my};
#pragma unroll
    for (int q = 0; q < 4; ++q)
#pragma unroll
      for (int m = 0; m < 10; ++m) S[q][m] = fma(qv[q], mom[m], S[q][m]);
  }
  double* Om = s_om[wid];
  double* Pm = Om + 81;
  double* Sm = Pm + 27;    // [4][10], indexed at run time below: shared memory, not registers
#pragma unroll
  for (int q = 0; q < 4; ++q)
#pragma unroll
    for (int m = 0; m < 10; ++m) {
      const double t = warp_sum(S[q][m]);
      if (lane == 0) Sm[10 * q + m] = t;
    }
  __syncwarp();
  // Q-weighted moment: sum_i Q_i[a][b] * mom_m
  auto Qm = [&](int a, int b, int m) -> double {
    if (a > b) { const int t = a; a = b; b = t; }
    if (a == 0 && b == 0) return Sm[m];
    if (a == 1 && b == 1) return Sm[m];
    if (a == 0 && b == 1) return 0.0;
    if (a == 0 && b == 2) return -Sm[10 + m];
    if (a == 1 && b == 2) return -Sm[20 + m];
    return Sm[30 + m];
  };
  auto m2 = [](int c, int d) -> int {   // index of p_c p_d among the moments
    if (c > d) { const int t = c; c = d; d = t; }
    return c == 0 ? 4 + d : (c == 1 ? 6 + d : 9);
  };
  // ---- P = -(sum Q)^-1 (sum Q A): 3x3 symmetric inverse by the adjugate
  const double q00 = Qm(0, 0, 0), q02 = Qm(0, 2, 0), q12 = Qm(1, 2, 0), q22 = Qm(2, 2, 0);   // sum Q = [q00 0 q02; 0 q00 q12; q02 q12 q22]
  const double det = q00 * (q00 * q22 - q12 * q12) - q02 * q02 * q00;
  const double idet = 1.0 / det;
  const double Qi[3][3] = {{(q00 * q22 - q12 * q12) * idet, (q02 * q12) * idet, (-q02 * q00) * idet},
                           {(q02 * q12) * idet, (q00 * q22 - q02 * q02) * idet, (-q00 * q12) * idet},
                           {(-q02 * q00) * idet, (-q00 * q12) * idet, (q00 * q00) * idet}};
  if (lane < 27) {
    const int a = lane / 9, j = lane - 9 * a, b = j / 3, c = j - 3 * b;
    double s = 0.0;
    for (int e = 0; e < 3; ++e) s -= Qi[a][e] * Qm(e, b, 1 + c);   // (sum Q A)[e][3b+c] = sum Q[e][b] p_c
    Pm[lane] = s;
  }
  __syncwarp();
  for (int idx = lane; idx < 81; idx += 32) {
    const int i = idx / 9, j = idx - 9 * i;
    const int a = i / 3, c = i - 3 * a, b = j / 3, d = j - 3 * b;
    double s = Qm(a, b, m2(c, d));                                   // (A^T Q A)[3a+c][3b+d]
    for (int e = 0; e < 3; ++e) s = fma(Qm(e, a, 1 + c), Pm[9 * e + j], s);   // + (sum Q A)^T P
    Om[idx] = s;
  }
  __syncwarp();
  // ---- 32 damped-Newton runs on SO(3) in lock step
  double R[9], w[9];
  start_rotation(lane, R);
  double cost = quad_form(Om, R, w);
  double lam = 1e-3;
  bool settled = false;   // sticky: this run has taken a step below rounding level, or damped itself to a standstill
  for (int it = 0; it < kPnpIters; ++it) {
    // tangent basis J_k = vec([e_k]x R): rows (0, -R2, R1), (R2, 0, -R0), (-R1, R0, 0)
    double J[3][9];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
      J[0][j] = 0.0;       J[0][3 + j] = -R[6 + j]; J[0][6 + j] = R[3 + j];
      J[1][j] = R[6 + j];  J[1][3 + j] = 0.0;       J[1][6 + j] = -R[j];
      J[2][j] = -R[3 + j]; J[2][3 + j] = R[j];      J[2][6 + j] = 0.0;
    }
    double g[3], H[3][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
#pragma unroll
    for (int k = 0; k < 3; ++k) { double s = 0.0; for (int i = 0; i < 9; ++i) s = fma(J[k][i], w[i], s); g[k] = 2.0 * s; }
    // 2 J^T Omega J
#pragma unroll
    for (int i = 0; i < 9; ++i) {
      double oj[3] = {0.0, 0.0, 0.0};
#pragma unroll
      for (int j = 0; j < 9; ++j) { const double o = Om[9 * i + j]; oj[0] = fma(o, J[0][j], oj[0]); oj[1] = fma(o, J[1][j], oj[1]); oj[2] = fma(o, J[2][j], oj[2]); }
#pragma unroll
      for (int k = 0; k < 3; ++k)
#pragma unroll
        for (int l = 0; l < 3; ++l) H[k][l] = fma(2.0 * J[k][i], oj[l], H[k][l]);
    }
    // + curvature of the manifold: 2 w . vec(1/2 (E_k E_l + E_l E_k) R) = W_k . R_l + W_l . R_k - 2 delta_kl <W, R>
    double WR[3][3], tr = 0.0;
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int l = 0; l < 3; ++l) WR[k][l] = w[3 * k] * R[3 * l] + w[3 * k + 1] * R[3 * l + 1] + w[3 * k + 2] * R[3 * l + 2];
    tr = WR[0][0] + WR[1][1] + WR[2][2];
#pragma unroll
    for (int k = 0; k < 3; ++k)
#pragma unroll
      for (int l = 0; l < 3; ++l) H[k][l] += WR[k][l] + WR[l][k] - (k == l ? 2.0 * tr : 0.0);
    // damped 3x3 solve (adjugate): (H + lam (|diag H| + eps)) d = -g
    const double scale = fabs(H[0][0]) + fabs(H[1][1]) + fabs(H[2][2]) + 1e-300;
    const double a00 = H[0][0] + lam * scale, a11 = H[1][1] + lam * scale, a22 = H[2][2] + lam * scale;
    const double a01 = H[0][1], a02 = H[0][2], a12 = H[1][2];
    const double c00 = a11 * a22 - a12 * a12, c01 = a02 * a12 - a01 * a22, c02 = a01 * a12 - a02 * a11;
    const double dt = a00 * c00 + a01 * c01 + a02 * c02;
    const double c11 = a00 * a22 - a02 * a02, c12 = a01 * a02 - a00 * a12, c22 = a00 * a11 - a01 * a01;
    const bool pd = a00 > 0.0 && c22 > 0.0 && dt > 0.0;   // leading minors: positive definite
    const double id = pd ? 1.0 / dt : 0.0;
    double d[3] = {-(c00 * g[0] + c01 * g[1] + c02 * g[2]) * id, -(c01 * g[0] + c11 * g[1] + c12 * g[2]) * id,
                   -(c02 * g[0] + c12 * g[1] + c22 * g[2]) * id};
    // trust region: at most one radian per step
    const double dn = sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    if (dn > 1.0) { d[0] /= dn; d[1] /= dn; d[2] /= dn; }
    double Rn[9], wn[9];
    so3_left_update(d, R, Rn);
    const double cn = quad_form(Om, Rn, wn);
    const bool accept = pd && cn <= cost;
    if (accept) {
#pragma unroll
      for (int i = 0; i < 9; ++i) { R[i] = Rn[i]; w[i] = wn[i]; }
      cost = cn;
      lam = fmax(lam * 0.1, 1e-15);
    } else {
      lam = fmin(lam * 10.0, 1e15);
    }
    settled = settled || (pd && dn < 1e-12) || lam >= 1e10;
    if (__all_sync(0xffffffffu, settled)) break;   // warp-uniform exit: typically after 8-12 of the 40 iterations
  }
  // ---- t = P r; keep the cheapest solution that puts the board in front of the camera (mean depth > 0)
  double t[3];
#pragma unroll
  for (int a = 0; a < 3; ++a) { double s = 0.0; for (int j = 0; j < 9; ++j) s = fma(Pm[9 * a + j], R[j], s); t[a] = s; }
  const double inv_n = 1.0 / Sm[0];
  const double depth = (R[6] * Sm[1] + R[7] * Sm[2] + R[8] * Sm[3]) * inv_n + t[2];
  const bool ok = depth > 0.0 && cost == cost;
  double key = ok ? cost : 1.7976931348623157e308;
  int best = lane;
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const double k2 = __shfl_xor_sync(0xffffffffu, key, o);
    const int b2 = __shfl_xor_sync(0xffffffffu, best, o);
    if (k2 < key || (k2 == key && b2 < best)) { key = k2; best = b2; }
  }
  if (lane == best) {
    double rv[3];
    rvec_from_R(R, rv);
    double* o = poses_out + 6 * (size_t)f;
    o[0] = rv[0]; o[1] = rv[1]; o[2] = rv[2]; o[3] = t[0]; o[4] = t[1]; o[5] = t[2];
    if (cost_out) cost_out[f] = ok ? cost : nan("");
  }
}

cudaError_t launch_pnp(const int32_t* frame_offsets, const double* x, const double* y, const double* z, const double* xn,
                       const double* yn, int n_frames, double* poses_out, double* cost_out, cudaStream_t s) {
  k_pnp<<<(n_frames + kPnpWarps - 1) / kPnpWarps, kPnpWarps * 32, 0, s>>>(frame_offsets, x, y, z, xn, yn, n_frames, poses_out, cost_out);
  return cudaGetLastError();
}

}  // namespace ccrs

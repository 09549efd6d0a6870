// ccrs_joint.cu — joint multi-camera refinement: calib_all_camera_with_extrinsics (reference src/util.rs:567-715).
//
// Variables: params{c} (d per camera), T_c_0 = (rvec_{c}_0, tvec_{c}_0) for c > 0, T_0_b_f = (rvec_0_b_{f}, tvec_0_b_{f})
// per frame. cam0 corners are ReprojectionFactor blocks (util.rs:603-611), the other cameras' corners are
// OtherCamReprojectionFactor blocks with five parameter blocks (util.rs:612-631; factors.rs:204-228):
//     r = project(theta_c ; T_c_0 * T_0_b_f * p) - p2d.
// The per-frame board pose T_0_b_f is eliminated (6x6 Cholesky per frame, summing the blocks of every camera that saw
// the frame) onto the shared system [theta_0 .. theta_{C-1} | T_1_0 .. T_{C-1}_0]; the host solves that small dense
// system, exactly like the single-camera path. The reference runs Gauss-Newton here (util.rs:668-670).
// This problem is small (BASELINE config 5: 2 cameras x 200 frames), so the kernels favour simplicity: one warp per
// (camera, frame) block with the rows of [J | r] staged in shared memory.
#include "../../include/ccrs_b200.h"
#include "ccrs_device.cuh"
#include "ccrs_kernels.cuh"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <type_traits>
#include <vector>

using namespace ccrs;

namespace {

thread_local char g_jerr[512] = "";
int jfail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_jerr, sizeof(g_jerr), fmt, ap);
  va_end(ap);
  return code;
}
#define JCK(call)                                                                                        \
  do {                                                                                                   \
    cudaError_t _e = (call);                                                                             \
    if (_e != cudaSuccess) return jfail(CCRS_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

constexpr int kMaxShared = 64;   // max size of the shared (intrinsics + extrinsics) system
constexpr int kJWarps = 4;       // warps (= blocks of the problem) per CTA in k_joint_linearize

struct JointDev {
  const double *x, *y, *z, *u, *v;
  const int32_t *block_cam, *block_frame, *block_offsets, *obs_block;
  const int32_t *frame_block_offsets, *frame_blocks;   // CSR frame -> blocks
  const double* intr;    // [C][d]
  const double* extr;    // [C][6]
  double* poses;         // [F][6]
  int n_cams, n_frames, n_blocks, d;
  double huber_delta;
};

template <int MODEL, bool OF>
struct JCfg {
  static constexpr int ND = model_nd(MODEL);
  static constexpr int D = 4 + ND - (OF ? 1 : 0);
  static constexpr int N = D + 12;     // [theta | w0 t0 | wc tc]
  static constexpr int NA = N + 1;     // + r
  static constexpr int NB = NA * (NA + 1) / 2;
  static constexpr int KOFF = OF ? 3 : 4;
};

// rows of [J | r] (rvec basis, Huber-corrected) for one observation of a block of camera c
template <int MODEL, bool OF>
CCRS_D void joint_rows(const double* __restrict__ ip, const FramePose& f0, const FramePose& fc, bool other,
                       double px, double py, double pz, double ou, double ov, double delta,
                       double* __restrict__ au, double* __restrict__ av) {
  using C = JCfg<MODEL, OF>;
  const double fx = ip[0], fy = OF ? ip[0] : ip[1], cx = ip[2], cy = ip[3];
  const double q0x = f0.R[0] * px + f0.R[1] * py + f0.R[2] * pz;
  const double q0y = f0.R[3] * px + f0.R[4] * py + f0.R[5] * pz;
  const double q0z = f0.R[6] * px + f0.R[7] * py + f0.R[8] * pz;
  const double P0x = q0x + f0.t[0], P0y = q0y + f0.t[1], P0z = q0z + f0.t[2];
  double qcx = 0, qcy = 0, qcz = 0, X = P0x, Y = P0y, Z = P0z;
  if (other) {   // T_c_0 * (T_0_b * p)
    qcx = fc.R[0] * P0x + fc.R[1] * P0y + fc.R[2] * P0z;
    qcy = fc.R[3] * P0x + fc.R[4] * P0y + fc.R[5] * P0z;
    qcz = fc.R[6] * P0x + fc.R[7] * P0y + fc.R[8] * P0z;
    X = qcx + fc.t[0]; Y = qcy + fc.t[1]; Z = qcz + fc.t[2];
  }
  double m[2], dP[2][3], dk[2][kMaxNd];
  model_eval<MODEL, true>(ip + 4, X, Y, Z, m, dP, dk);
  const double ru = fma(fx, m[0], cx) - ou, rv = fma(fy, m[1], cy) - ov;
  const double w = huber_weight(ru * ru + rv * rv, delta);
#pragma unroll
  for (int i = 0; i < C::NA; ++i) { au[i] = 0.0; av[i] = 0.0; }
  if constexpr (OF) { au[0] = w * m[0]; av[0] = w * m[1]; au[1] = w; av[2] = w; }
  else { au[0] = w * m[0]; av[1] = w * m[1]; au[2] = w; av[3] = w; }
  const double wf[2] = {w * fx, w * fy};
#pragma unroll
  for (int j = 0; j < C::ND; ++j) { au[C::KOFF + j] = wf[0] * dk[0][j]; av[C::KOFF + j] = wf[1] * dk[1][j]; }
#pragma unroll
  for (int row = 0; row < 2; ++row) {
    double* a = row ? av : au;
    const double d0 = wf[row] * dP[row][0], d1 = wf[row] * dP[row][1], d2 = wf[row] * dP[row][2];
    double e0 = d0, e1 = d1, e2 = d2;   // derivative w.r.t. P_0 (cam0 frame)
    if (other) {
      // T_c_0 columns: tvec -> d ; rvec -> (q_c x d)^T J_l(w_c)
      const double c0 = qcy * d2 - qcz * d1, c1 = qcz * d0 - qcx * d2, c2 = qcx * d1 - qcy * d0;
#pragma unroll
      for (int k = 0; k < 3; ++k) a[C::D + 6 + k] = c0 * fc.Jl[k] + c1 * fc.Jl[3 + k] + c2 * fc.Jl[6 + k];
      a[C::D + 9] = d0; a[C::D + 10] = d1; a[C::D + 11] = d2;
      e0 = fc.R[0] * d0 + fc.R[3] * d1 + fc.R[6] * d2;   // R_c^T d
      e1 = fc.R[1] * d0 + fc.R[4] * d1 + fc.R[7] * d2;
      e2 = fc.R[2] * d0 + fc.R[5] * d1 + fc.R[8] * d2;
    }
    const double c0 = q0y * e2 - q0z * e1, c1 = q0z * e0 - q0x * e2, c2 = q0x * e1 - q0y * e0;
#pragma unroll
    for (int k = 0; k < 3; ++k) a[C::D + k] = c0 * f0.Jl[k] + c1 * f0.Jl[3 + k] + c2 * f0.Jl[6 + k];
    a[C::D + 3] = e0; a[C::D + 4] = e1; a[C::D + 5] = e2;
  }
  au[C::N] = w * ru; av[C::N] = w * rv;
}

template <int MODEL, bool OF>
CCRS_D void load_full_intr(const double* a, double* ip) {
  using C = JCfg<MODEL, OF>;
  if constexpr (OF) { ip[0] = a[0]; ip[1] = a[0]; for (int i = 1; i < C::D; ++i) ip[i + 1] = a[i]; }
  else { for (int i = 0; i < C::D; ++i) ip[i] = a[i]; }
}

// parity hook for a3: per observation r and J (2 x (d+12))
template <int MODEL, bool OF>
__global__ void __launch_bounds__(128) k_joint_eval_rj(JointDev jd, int apply_loss, double* __restrict__ r,
                                                       double* __restrict__ J, int n_obs) {
  using C = JCfg<MODEL, OF>;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_obs) return;
  const int b = jd.obs_block[k], c = jd.block_cam[b], f = jd.block_frame[b];
  double ip[kMaxFull];
  load_full_intr<MODEL, OF>(jd.intr + (size_t)c * C::D, ip);
  FramePose f0, fc;
  pose_from_rvec_tvec(jd.poses + 6 * (size_t)f, f0);
  pose_from_rvec_tvec(jd.extr + 6 * (size_t)c, fc);
  double au[C::NA], av[C::NA];
  joint_rows<MODEL, OF>(ip, f0, fc, c > 0, jd.x[k], jd.y[k], jd.z[k], jd.u[k], jd.v[k], apply_loss ? jd.huber_delta : 0.0, au, av);
  r[2 * (size_t)k] = au[C::N]; r[2 * (size_t)k + 1] = av[C::N];
  if (J) for (int i = 0; i < C::N; ++i) { J[(size_t)(2 * k) * C::N + i] = au[i]; J[(size_t)(2 * k + 1) * C::N + i] = av[i]; }
}

// one warp per (camera, frame) block: packed Gram block [J r]^T [J r] of size NB, deterministic (fixed lane order)
template <int MODEL, bool OF>
__global__ void __launch_bounds__(32 * kJWarps) k_joint_linearize(JointDev jd, const uint16_t* __restrict__ ij_table,
                                                                  double* __restrict__ jblk) {
  using C = JCfg<MODEL, OF>;
  __shared__ double rows[kJWarps][32][2 * C::NA];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x * kJWarps + warp;
  if (b >= jd.n_blocks) return;
  const int c = jd.block_cam[b], f = jd.block_frame[b];
  double ip[kMaxFull];
  load_full_intr<MODEL, OF>(jd.intr + (size_t)c * C::D, ip);
  FramePose f0, fc;
  pose_from_rvec_tvec(jd.poses + 6 * (size_t)f, f0);
  pose_from_rvec_tvec(jd.extr + 6 * (size_t)c, fc);
  constexpr int EPL = (C::NB + 31) / 32;   // entries per lane
  double acc[EPL];
  int ei[EPL], ej[EPL];
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    acc[e] = 0.0;
    const int idx = lane + 32 * e;
    const uint16_t t = idx < C::NB ? ij_table[idx] : 0;
    ei[e] = t >> 8; ej[e] = t & 0xff;
  }
  const int beg = jd.block_offsets[b], end = jd.block_offsets[b + 1];
  for (int base = beg; base < end; base += 32) {
    const int k = base + lane;
    double* my = rows[warp][lane];
    if (k < end) {
      double au[C::NA], av[C::NA];
      joint_rows<MODEL, OF>(ip, f0, fc, c > 0, jd.x[k], jd.y[k], jd.z[k], jd.u[k], jd.v[k], jd.huber_delta, au, av);
#pragma unroll
      for (int i = 0; i < C::NA; ++i) { my[i] = au[i]; my[C::NA + i] = av[i]; }
    }
    __syncwarp();
    const int cnt = min(32, end - base);
    for (int l = 0; l < cnt; ++l) {
      const double* rw = rows[warp][l];
#pragma unroll
      for (int e = 0; e < EPL; ++e)
        acc[e] = fma(rw[ei[e]], rw[ej[e]], fma(rw[C::NA + ei[e]], rw[C::NA + ej[e]], acc[e]));
    }
    __syncwarp();
  }
#pragma unroll
  for (int e = 0; e < EPL; ++e) {
    const int idx = lane + 32 * e;
    if (idx < C::NB) jblk[(size_t)b * C::NB + idx] = acc[e];
  }
}

// The same block on the FP64 tensor path (the form K2 uses for the d >= 8 models, ccrs_linmma.cu): lanes = observations,
// the 2 x 32 rows of a round staged in shared memory as [24 columns][4 row classes][16 pairs], H (24 x 24, six upper
// 8 x 8 tiles) += J^T J with mma.sync.m8n8k4.f64, four rows = two observations per instruction. The scalar version above
// read four shared-memory values per FMA pair (32 rows x 8 entries per lane and round) and was bound by that.
constexpr int kJMmaWarps = 4;     // warps per (camera, frame) block: warp w takes the rounds w, w + 4, ... of 32 observations
constexpr int kJMmaColStride = 72, kJMmaRowStride = 18, kJMmaCols = 24;
constexpr size_t kJMmaSmem = (size_t)kJMmaWarps * kJMmaCols * kJMmaColStride * sizeof(double);
CCRS_D void jdmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODEL, bool OF>
__global__ void __launch_bounds__(32 * kJMmaWarps) k_joint_linearize_mma(JointDev jd, double* __restrict__ jblk) {
  using C = JCfg<MODEL, OF>;
  static_assert(C::NA <= kJMmaCols, "three column blocks of eight cover the joint block");
  extern __shared__ __align__(16) double s_all[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = blockIdx.x;
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched as a programmatic dependent of the kernel before it
  double* const sr = s_all + (size_t)warp * kJMmaCols * kJMmaColStride;
  for (int i = lane; i < kJMmaCols * kJMmaColStride; i += 32) sr[i] = 0.0;   // columns >= NA stay zero
  const int c = jd.block_cam[b], f = jd.block_frame[b];
  double ip[kMaxFull];
  load_full_intr<MODEL, OF>(jd.intr + (size_t)c * C::D, ip);
  FramePose f0, fc;
  pose_from_rvec_tvec(jd.poses + 6 * (size_t)f, f0);
  pose_from_rvec_tvec(jd.extr + 6 * (size_t)c, fc);
  // lane l reads row class l % 4 of column 8 blk + l / 4 (pairs g, g + 1 as one 16-byte load); lane o writes class
  // 2 (o % 2) (u) and the next one (v) of pair o / 2
  const double* const rd = sr + (lane >> 2) * kJMmaColStride + (lane & 3) * kJMmaRowStride;
  double* const wr_u = sr + (2 * (lane & 1)) * kJMmaRowStride + (lane >> 1);
  double* const wr_v = wr_u + kJMmaRowStride;
  // tiles (0,0) (0,1) (0,2) (1,1) (1,2) (2,2), two accumulator sets (even / odd pairs) summed at the end
  double ca[12], cb[12];
#pragma unroll
  for (int i = 0; i < 12; ++i) { ca[i] = 0.0; cb[i] = 0.0; }
  __syncwarp();
  const int beg = jd.block_offsets[b], end = jd.block_offsets[b + 1];
  for (int base = beg + 32 * warp; base < end; base += 32 * kJMmaWarps) {
    const int k = base + lane;
    const int cnt = min(32, end - base);
    {
      double au[C::NA], av[C::NA];
      const int kk = min(k, end - 1);
      joint_rows<MODEL, OF>(ip, f0, fc, c > 0, jd.x[kk], jd.y[kk], jd.z[kk], jd.u[kk], jd.v[kk], jd.huber_delta, au, av);
      const bool valid = k < end;
#pragma unroll
      for (int i = 0; i < C::NA; ++i) {
        wr_u[i * kJMmaColStride] = valid ? au[i] : 0.0;
        wr_v[i * kJMmaColStride] = valid ? av[i] : 0.0;
      }
    }
    __syncwarp();
    const int np2 = (cnt + 3) >> 2;   // pairs of pairs holding at least one observation
    for (int g2 = 0; g2 < np2; ++g2) {
      const double2 a0 = *reinterpret_cast<const double2*>(rd + 2 * g2);
      const double2 a1 = *reinterpret_cast<const double2*>(rd + 8 * kJMmaColStride + 2 * g2);
      const double2 a2 = *reinterpret_cast<const double2*>(rd + 16 * kJMmaColStride + 2 * g2);
      jdmma884(ca[0], ca[1], a0.x, a0.x); jdmma884(ca[2], ca[3], a0.x, a1.x); jdmma884(ca[4], ca[5], a0.x, a2.x);
      jdmma884(ca[6], ca[7], a1.x, a1.x); jdmma884(ca[8], ca[9], a1.x, a2.x); jdmma884(ca[10], ca[11], a2.x, a2.x);
      jdmma884(cb[0], cb[1], a0.y, a0.y); jdmma884(cb[2], cb[3], a0.y, a1.y); jdmma884(cb[4], cb[5], a0.y, a2.y);
      jdmma884(cb[6], cb[7], a1.y, a1.y); jdmma884(cb[8], cb[9], a1.y, a2.y); jdmma884(cb[10], cb[11], a2.y, a2.y);
    }
    __syncwarp();
  }
  // the four warps' partial tiles are summed in warp order by warp 0 (fixed order: deterministic); a warp's partial goes
  // through its own staging buffer, which it no longer needs
#pragma unroll
  for (int i = 0; i < 12; ++i) sr[i * 32 + lane] = ca[i] + cb[i];
  __syncthreads();
  if (warp != 0) return;
  // lane holds (row l / 4, columns 2 (l % 4) + {0, 1}) of every tile
  const int ti = lane >> 2, tj = 2 * (lane & 3);
  constexpr int trow[6] = {0, 0, 0, 1, 1, 2}, tcol[6] = {0, 1, 2, 1, 2, 2};
#pragma unroll
  for (int t = 0; t < 6; ++t)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int gi = 8 * trow[t] + ti, gj = 8 * tcol[t] + tj + e;
      double v = 0.0;
#pragma unroll
      for (int w = 0; w < kJMmaWarps; ++w) v += s_all[(size_t)w * kJMmaCols * kJMmaColStride + (2 * t + e) * 32 + lane];
      if (gi <= gj && gj < C::NA) jblk[(size_t)b * C::NB + tri_idx(C::NA, gi, gj)] = v;
    }
}

// per frame: gather the blocks of every camera, eliminate T_0_b_f. Output per frame:
// fs[f][NS + M + 1] = packed upper S_f (shared x shared), g_s (M), sum r^2 ; el[f][6*M + 6] = X (6 x M), C^-1 g_p
// One WARP per frame (a thread per frame spent ~200 us in serial read-modify-write of its 250-entry system): the
// frame's shared system lives in the warp's shared memory; the lanes scatter the packed entries of each camera block
// into it (every entry of a block has its own target: no atomics; blocks in order: deterministic), all lanes factor the
// 6x6 pose block redundantly in registers, then the columns of Y = L^-1 B^T and the entries of S -= Y^T Y are spread
// over the lanes.
constexpr int kJSchurWarps = 4;
__host__ __device__ inline int jschur_warp_doubles(int M) { return M * (M + 1) / 2 + 7 * M + 48; }   // S | B (M x 6) | gs | L (36) gp (6) pad

__global__ void __launch_bounds__(32 * kJSchurWarps) k_joint_schur(JointDev jd, const double* __restrict__ jblk,
                                                                  const uint16_t* __restrict__ ij_table, int NAJ, int M,
                                                                  double u_damp, double* __restrict__ fs, double* __restrict__ el) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched as a programmatic dependent of the kernel before it

  extern __shared__ double jsm[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int f = blockIdx.x * kJSchurWarps + warp;
  if (f >= jd.n_frames) return;
  const int d = jd.d, C = jd.n_cams, NBJ = NAJ * (NAJ + 1) / 2, NS = M * (M + 1) / 2, n = d + 12;
  const int off_e = C * d;
  double* S = jsm + (size_t)warp * jschur_warp_doubles(M);
  double* B = S + NS;          // [M][6]
  double* gs = B + 6 * M;      // [M]
  double* Lm = gs + M;         // [6][6] lower
  double* gp = Lm + 36;        // [6]
  for (int i = lane; i < NS + 7 * M + 42; i += 32) S[i] = 0.0;
  __syncwarp();
  double sq = 0.0;
  for (int q = jd.frame_block_offsets[f]; q < jd.frame_block_offsets[f + 1]; ++q) {
    const int b = jd.frame_blocks[q], c = jd.block_cam[b];
    const double* Hb = jblk + (size_t)b * NBJ;
    // shared index of block column i (theta: 0..d-1 ; extrinsic: d+6..d+11), -1 if not a shared column
    auto sh = [&](int i) { return i < d ? c * d + i : (i >= d + 6 && i < n && c > 0 ? off_e + 6 * (c - 1) + (i - d - 6) : -1); };
    for (int e = lane; e < NBJ; e += 32) {
      const uint16_t t = ij_table[e];
      const int i = t >> 8, j = t & 255;      // i <= j
      const double h = Hb[e];
      const bool ip = i >= d && i < d + 6, jp = j >= d && j < d + 6;
      if (j == n) {                            // the r column: gradient / cost
        if (i == n) sq += h;
        else if (ip) gp[i - d] -= h;
        else { const int si = sh(i); if (si >= 0) gs[si] -= h; }
      } else if (ip && jp) {
        Lm[(j - d) * 6 + (i - d)] += h;        // lower triangle: row >= column
      } else if (ip) {                         // pose x extrinsic column
        const int sj = sh(j); if (sj >= 0) B[sj * 6 + (i - d)] += h;
      } else if (jp) {                         // intrinsic x pose column
        const int si = sh(i); if (si >= 0) B[si * 6 + (j - d)] += h;
      } else {
        const int si = sh(i), sj = sh(j);
        if (si >= 0 && sj >= 0) S[si <= sj ? tri_idx(M, si, sj) : tri_idx(M, sj, si)] += h;
      }
    }
    __syncwarp();
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  // 6x6 Cholesky of the (damped) pose block and L^-1 g_p: every lane, in registers
  double L[6][6], yg[6];
  int bad = 0;
#pragma unroll
  for (int i = 0; i < 6; ++i)
#pragma unroll
    for (int j = 0; j <= i; ++j) L[i][j] = Lm[i * 6 + j];
#pragma unroll
  for (int i = 0; i < 6; ++i) L[i][i] += u_damp * L[i][i];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    double t = L[j][j];
#pragma unroll
    for (int k = 0; k < j; ++k) t -= L[j][k] * L[j][k];
    if (!(t > 0.0)) bad = 1;
    const double il = 1.0 / sqrt(t);
    L[j][j] = il;
#pragma unroll
    for (int i = j + 1; i < 6; ++i) {
      double w = L[i][j];
#pragma unroll
      for (int k = 0; k < j; ++k) w -= L[i][k] * L[j][k];
      L[i][j] = w * il;
    }
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double t = gp[i];
#pragma unroll
    for (int k = 0; k < i; ++k) t -= L[i][k] * yg[k];
    yg[i] = t * L[i][i];
  }
  __syncwarp();
  // Y[a] = L^-1 B[a] (columns over lanes), kept in B
  for (int a = lane; a < M; a += 32) {
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double t = B[a * 6 + i];
#pragma unroll
      for (int k = 0; k < i; ++k) t -= L[i][k] * y[k];
      y[i] = t * L[i][i];
    }
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) { B[a * 6 + i] = y[i]; t += y[i] * yg[i]; }
    gs[a] -= t;
  }
  __syncwarp();
  // S -= Y^T Y (entries over lanes) -> fs
  double* So = fs + (size_t)f * (NS + M + 1);
  const double poison = bad ? nan("") : 0.0;   // host reports the LLT failure (S and g_s poisoned, not the cost)
  for (int e = lane; e < NS; e += 32) {
    int a = 0, rem = e;                        // row a of the packed upper triangle holds M - a entries
    while (rem >= M - a) { rem -= M - a; ++a; }
    const int b2 = a + rem;
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) t += B[a * 6 + i] * B[b2 * 6 + i];
    So[e] = S[e] - t + poison;
  }
  for (int a = lane; a < M; a += 32) So[NS + a] = gs[a] + poison;
  if (lane == 0) So[NS + M] = sq;
  // X = L^-T Y, C^-1 g_p -> el
  double* eo = el + (size_t)f * (6 * M + 6);
  for (int a = lane; a < M; a += 32) {
    double y[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) y[i] = B[a * 6 + i];
#pragma unroll
    for (int i = 5; i >= 0; --i) {
      double t = y[i];
#pragma unroll
      for (int k = i + 1; k < 6; ++k) t -= L[k][i] * y[k];
      y[i] = t * L[i][i];
      eo[i * M + a] = y[i];
    }
  }
  if (lane == 0) {
#pragma unroll
    for (int i = 5; i >= 0; --i) {
      double t = yg[i];
#pragma unroll
      for (int k = i + 1; k < 6; ++k) t -= L[k][i] * yg[k];
      yg[i] = t * L[i][i];
      eo[6 * M + i] = yg[i];
    }
  }
}

// out[v] = sum over frames of fs[f][v]: one WARP per value — lane-strided partial sums in frame order, then a fixed
// butterfly (deterministic) — so a value costs one round of loads instead of F / 8 dependent ones (a thread per value took
// 20 us for 198 frames). Published straight to mapped host memory (host_out: the host armed the NV words with a sentinel
// and spins until all of them changed — no memcpy, no sync).
constexpr int kJSumWarps = 8;
__global__ void __launch_bounds__(32 * kJSumWarps) k_joint_sum(const double* __restrict__ fs, int F, int NV, double* __restrict__ out,
                                                              volatile double* host_out) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched as a programmatic dependent of the kernel before it

  const int v = blockIdx.x * kJSumWarps + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (v >= NV) return;
  double s = 0.0;
  for (int f = lane; f < F; f += 32) s += fs[(size_t)f * NV + v];
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    out[v] = s;
    if (host_out) host_out[v] = s;
  }
}

// The iteration's new state — intrinsics | extrinsics | shared step y — travels as a KERNEL ARGUMENT (<= 1.2 KB): no
// host-to-device copy (copy-engine hand-over between two kernels) on the iteration path.
struct JointState { double v[2 * kMaxShared + 8]; };
// one warp per frame: lane (i, part) sums a strided part of row i of X y; five lanes per row, fixed combine order.
// CTA 0 also leaves intr | extr (the first n_head values of the state) where the next linearisation reads them.
__global__ void __launch_bounds__(128) k_joint_backsub(JointDev jd, const double* __restrict__ el, const __grid_constant__ JointState st,
                                                       int n_head, double* __restrict__ state_dev, int M) {
  asm volatile("griddepcontrol.launch_dependents;");
  asm volatile("griddepcontrol.wait;" ::: "memory");   // launched as a programmatic dependent of the kernel before it

  if (blockIdx.x == 0) for (int i = threadIdx.x; i < n_head; i += blockDim.x) state_dev[i] = st.v[i];
  const int f = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (f >= jd.n_frames) return;
  const double* e = el + (size_t)f * (6 * M + 6);
  const double* y = st.v + n_head;
  const int i = lane / 5, part = lane - 5 * i;     // lanes 0..29: row i, part 0..4; lanes 30, 31 idle
  double s = 0.0;
  if (i < 6) for (int a = part; a < M; a += 5) s += e[i * M + a] * y[a];
  const double s1 = __shfl_down_sync(0xffffffffu, s, 1), s2 = __shfl_down_sync(0xffffffffu, s, 2);
  const double s3 = __shfl_down_sync(0xffffffffu, s, 3), s4 = __shfl_down_sync(0xffffffffu, s, 4);
  if (i < 6 && part == 0) jd.poses[6 * (size_t)f + i] += e[6 * M + i] - ((((s + s1) + s2) + s3) + s4);
}

// every kernel of the joint iteration is launched as a programmatic dependent of the one before it (its CTAs become
// resident while the predecessor drains and wait at griddepcontrol.wait): the launch latency leaves the critical path
template <class K, class... A>
static cudaError_t jlaunch(K kern, int grid, int block, size_t smem, cudaStream_t s, A... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, args...);
}

template <class F>
auto jdispatch(int model, int of, F&& f) {
#define JCASE(M) case M: return of ? f(std::integral_constant<int, M>{}, std::true_type{}) : f(std::integral_constant<int, M>{}, std::false_type{});
  switch (model) { JCASE(UCM) JCASE(EUCM) JCASE(EUCMT) JCASE(KB4) JCASE(OPENCV5) JCASE(FTHETA) }
#undef JCASE
  return f(std::integral_constant<int, EUCM>{}, std::false_type{});
}

bool chol_solve_dense(std::vector<double>& A, int n, double* b) {
  for (int j = 0; j < n; ++j) {
    double s = A[j * n + j];
    for (int k = 0; k < j; ++k) s -= A[j * n + k] * A[j * n + k];
    if (!(s > 0.0)) return false;
    const double l = std::sqrt(s);
    A[j * n + j] = l;
    for (int i = j + 1; i < n; ++i) { double t = A[i * n + j]; for (int k = 0; k < j; ++k) t -= A[i * n + k] * A[j * n + k]; A[i * n + j] = t / l; }
  }
  for (int i = 0; i < n; ++i) { double s = b[i]; for (int k = 0; k < i; ++k) s -= A[i * n + k] * b[k]; b[i] = s / A[i * n + i]; }
  for (int i = n - 1; i >= 0; --i) { double s = b[i]; for (int k = i + 1; k < n; ++k) s -= A[k * n + i] * b[k]; b[i] = s / A[i * n + i]; }
  return true;
}

}  // namespace

struct ccrs_joint {
  int model = 0, one_focal = 0, d = 0, NAJ = 0, NBJ = 0, M = 0, NS = 0;
  int n_cams = 0, n_frames = 0, n_blocks = 0, n_obs = 0, device = 0;
  double huber = 1.0;
  cudaStream_t stream = nullptr;
  std::vector<void*> allocs;
  double *x = nullptr, *y = nullptr, *z = nullptr, *u = nullptr, *v = nullptr;
  int32_t *block_cam = nullptr, *block_frame = nullptr, *block_offsets = nullptr, *obs_block = nullptr;
  int32_t *frame_block_offsets = nullptr, *frame_blocks = nullptr;
  uint16_t* ij_table = nullptr;
  double *intr = nullptr, *extr = nullptr, *poses = nullptr, *jblk = nullptr, *fs = nullptr, *el = nullptr, *red = nullptr, *ydev = nullptr;
  double* state_dev = nullptr;    // [C*d | C*6 | M] = intr | extr | y: one allocation, one copy per iteration
  double* h_state = nullptr;      // pinned staging of the same layout
  double* h_red = nullptr;        // mapped pinned [NS + M + 1]: the reduced system, published by k_joint_sum
  int64_t launches = 0;

  template <class T> cudaError_t alloc(T** p, size_t n) {
    cudaError_t e = cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) allocs.push_back(*p);
    return e;
  }
  JointDev dev() const {
    JointDev j{};
    j.x = x; j.y = y; j.z = z; j.u = u; j.v = v;
    j.block_cam = block_cam; j.block_frame = block_frame; j.block_offsets = block_offsets; j.obs_block = obs_block;
    j.frame_block_offsets = frame_block_offsets; j.frame_blocks = frame_blocks;
    j.intr = intr; j.extr = extr; j.poses = poses;
    j.n_cams = n_cams; j.n_frames = n_frames; j.n_blocks = n_blocks; j.d = d; j.huber_delta = huber;
    return j;
  }
};

namespace {
// intr | extr (| y) go to the device as ONE asynchronous copy from pinned staging; no synchronisation (the staging buffer
// is rewritten only after the host has seen the next reduced system, i.e. after this copy has been consumed)
int upload_state(ccrs_joint* p, const double* intr, const double* extr, const double* poses, const double* y = nullptr) {
  const size_t ni = (size_t)p->n_cams * p->d, ne = (size_t)p->n_cams * 6;
  std::memcpy(p->h_state, intr, ni * 8);
  std::memcpy(p->h_state + ni, extr, ne * 8);
  for (int i = 0; i < 6; ++i) p->h_state[ni + i] = 0.0;   // cam0 is the reference frame (util.rs:689-690)
  if (y) std::memcpy(p->h_state + ni + ne, y, (size_t)p->M * 8);
  JCK(cudaMemcpyAsync(p->state_dev, p->h_state, (ni + ne + (y ? p->M : 0)) * 8, cudaMemcpyHostToDevice, p->stream));
  if (poses) {
    JCK(cudaMemcpyAsync(p->poses, poses, (size_t)p->n_frames * 48, cudaMemcpyHostToDevice, p->stream));
    JCK(cudaStreamSynchronize(p->stream));   // caller-owned pageable memory
  }
  return 0;
}

constexpr unsigned long long kJSentinel = 0x7ff8dead5e471e15ULL;
void jarm(volatile double* dst, int n) {
  volatile unsigned long long* w = reinterpret_cast<volatile unsigned long long*>(dst);
  for (int i = 0; i < n; ++i) w[i] = kJSentinel;
  __sync_synchronize();
}
int jwait(ccrs_joint* p, volatile double* src, int n) {
  volatile unsigned long long* w = reinterpret_cast<volatile unsigned long long*>(src);
  for (unsigned long spins = 1;; ++spins) {
    int i = n - 1;
    while (i >= 0 && w[i] != kJSentinel) --i;
    if (i < 0) return 0;
    if ((spins & 0x3fff) == 0) {
      cudaError_t e = cudaStreamQuery(p->stream);
      if (e == cudaSuccess) {
        i = n - 1;
        while (i >= 0 && w[i] != kJSentinel) --i;
        if (i < 0) return 0;
        return jfail(CCRS_ERR_CUDA, "stream drained but the reduced system was never published");
      }
      if (e != cudaErrorNotReady) return jfail(CCRS_ERR_CUDA, "stream error while waiting: %s", cudaGetErrorString(e));
    }
  }
}
}  // namespace

extern "C" {

int ccrs_joint_create(ccrs_joint** out, int model, int xy_same_focal, int n_cams, int n_frames, int n_blocks,
                      const int32_t* block_cam, const int32_t* block_frame, const int32_t* block_offsets,
                      const double* x, const double* y, const double* z, const double* u, const double* v,
                      double huber_delta, int device_id) {
  if (!out || !block_cam || !block_frame || !block_offsets || !x || !y || !z || !u || !v || n_cams < 1 || n_frames < 1 || n_blocks < 1)
    return jfail(CCRS_ERR_INVALID, "null pointer or empty joint problem");
  if (model < 0 || model > 5) return jfail(CCRS_ERR_INVALID, "unknown model %d", model);
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); return jfail(CCRS_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path"); }
  if (device_id < 0 || device_id >= n_dev) return jfail(CCRS_ERR_INVALID, "device out of range");
  int major = 0;
  JCK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device_id));
  if (major != 10) return jfail(CCRS_ERR_NO_DEVICE, "kernels are built for sm_100a only");
  JCK(cudaSetDevice(device_id));
  ccrs_joint* p = new ccrs_joint();
  p->model = model; p->one_focal = xy_same_focal ? 1 : 0; p->n_cams = n_cams; p->n_frames = n_frames; p->n_blocks = n_blocks;
  p->device = device_id; p->huber = huber_delta;
  model_dims(model, p->one_focal, &p->d, nullptr, nullptr, nullptr);
  p->NAJ = p->d + 13; p->NBJ = p->NAJ * (p->NAJ + 1) / 2;
  p->M = n_cams * p->d + 6 * (n_cams - 1); p->NS = p->M * (p->M + 1) / 2;
  if (p->M > kMaxShared || n_cams * p->d + n_cams * 6 + p->M > 2 * kMaxShared + 8) { delete p; return jfail(CCRS_ERR_INVALID, "shared system of %d unknowns exceeds %d", p->M, kMaxShared); }
  for (int b = 0; b < n_blocks; ++b)
    if (block_cam[b] < 0 || block_cam[b] >= n_cams || block_frame[b] < 0 || block_frame[b] >= n_frames || block_offsets[b + 1] < block_offsets[b]) {
      delete p; return jfail(CCRS_ERR_INVALID, "bad block %d", b);
    }
  p->n_obs = block_offsets[n_blocks];
  const size_t N = p->n_obs;
  cudaError_t e = cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete p; return jfail(CCRS_ERR_CUDA, "stream"); }
  // host-side index structures
  std::vector<int32_t> obs_block(N), fbo(n_frames + 1, 0), fb(n_blocks);
  for (int b = 0; b < n_blocks; ++b) for (int k = block_offsets[b]; k < block_offsets[b + 1]; ++k) obs_block[k] = b;
  for (int b = 0; b < n_blocks; ++b) fbo[block_frame[b] + 1]++;
  for (int f = 0; f < n_frames; ++f) fbo[f + 1] += fbo[f];
  { std::vector<int32_t> cur(fbo.begin(), fbo.end() - 1); for (int b = 0; b < n_blocks; ++b) fb[cur[block_frame[b]]++] = b; }   // ascending block id per frame
  std::vector<uint16_t> ij(p->NBJ);
  { int k = 0; for (int i = 0; i < p->NAJ; ++i) for (int j = i; j < p->NAJ; ++j) ij[k++] = (uint16_t)((i << 8) | j); }
#define JA(ptr, n) do { cudaError_t _e = p->alloc(&p->ptr, n); if (_e != cudaSuccess) { ccrs_joint_destroy(p); return jfail(CCRS_ERR_CUDA, "cudaMalloc: %s", cudaGetErrorString(_e)); } } while (0)
  JA(x, N); JA(y, N); JA(z, N); JA(u, N); JA(v, N);
  JA(block_cam, n_blocks); JA(block_frame, n_blocks); JA(block_offsets, n_blocks + 1); JA(obs_block, N);
  JA(frame_block_offsets, n_frames + 1); JA(frame_blocks, n_blocks); JA(ij_table, p->NBJ);
  JA(state_dev, (size_t)n_cams * p->d + (size_t)n_cams * 6 + p->M); JA(poses, (size_t)n_frames * 6);
  p->intr = p->state_dev; p->extr = p->state_dev + (size_t)n_cams * p->d; p->ydev = p->extr + (size_t)n_cams * 6;
  if (cudaHostAlloc((void**)&p->h_state, ((size_t)n_cams * p->d + (size_t)n_cams * 6 + p->M) * 8, cudaHostAllocDefault) != cudaSuccess ||
      cudaHostAlloc((void**)&p->h_red, (size_t)(p->NS + p->M + 1) * 8, cudaHostAllocMapped) != cudaSuccess) {
    ccrs_joint_destroy(p); return jfail(CCRS_ERR_CUDA, "cudaHostAlloc");
  }
  JA(jblk, (size_t)n_blocks * p->NBJ); JA(fs, (size_t)n_frames * (p->NS + p->M + 1)); JA(el, (size_t)n_frames * (6 * p->M + 6));
  JA(red, p->NS + p->M + 1);
#undef JA
  cudaStream_t s = p->stream;
  cudaMemcpyAsync(p->x, x, N * 8, cudaMemcpyHostToDevice, s); cudaMemcpyAsync(p->y, y, N * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(p->z, z, N * 8, cudaMemcpyHostToDevice, s); cudaMemcpyAsync(p->u, u, N * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(p->v, v, N * 8, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(p->block_cam, block_cam, n_blocks * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(p->block_frame, block_frame, n_blocks * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(p->block_offsets, block_offsets, (n_blocks + 1) * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(p->obs_block, obs_block.data(), N * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(p->frame_block_offsets, fbo.data(), (n_frames + 1) * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(p->frame_blocks, fb.data(), n_blocks * 4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(p->ij_table, ij.data(), p->NBJ * 2, cudaMemcpyHostToDevice, s);
  e = cudaStreamSynchronize(s);
  if (e != cudaSuccess) { ccrs_joint_destroy(p); return jfail(CCRS_ERR_CUDA, "upload: %s", cudaGetErrorString(e)); }
  *out = p;
  return 0;
}

int ccrs_joint_destroy(ccrs_joint* p) {
  if (!p) return 0;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  for (void* a : p->allocs) cudaFree(a);
  if (p->h_state) cudaFreeHost(p->h_state);
  if (p->h_red) cudaFreeHost(p->h_red);
  if (p->stream) cudaStreamDestroy(p->stream);
  delete p;
  return 0;
}

int ccrs_joint_dim(const ccrs_joint* p) { return p ? p->d : -1; }
const char* ccrs_joint_last_error(void) { return g_jerr; }
int64_t ccrs_joint_launch_count(const ccrs_joint* p) { return p ? p->launches : -1; }

int ccrs_joint_eval_rj(ccrs_joint* p, const double* intr, const double* extr, const double* poses, int apply_loss,
                       double* r, double* J) {
  if (!p || !intr || !extr || !poses || !r) return jfail(CCRS_ERR_INVALID, "null");
  JCK(cudaSetDevice(p->device));
  int st = upload_state(p, intr, extr, poses);
  if (st) return st;
  const size_t N = p->n_obs;
  const int n = p->d + 12;
  double *dr = nullptr, *dJ = nullptr;
  JCK(cudaMalloc((void**)&dr, 2 * N * 8));
  if (J) JCK(cudaMalloc((void**)&dJ, 2 * N * n * 8));
  JointDev jd = p->dev();
  cudaError_t e = jdispatch(p->model, p->one_focal, [&](auto M, auto OF) {
    k_joint_eval_rj<decltype(M)::value, decltype(OF)::value><<<(int)((N + 127) / 128), 128, 0, p->stream>>>(jd, apply_loss, dr, dJ, (int)N);
    return cudaGetLastError();
  });
  p->launches++;
  JCK(e);
  JCK(cudaMemcpyAsync(r, dr, 2 * N * 8, cudaMemcpyDeviceToHost, p->stream));
  if (J) JCK(cudaMemcpyAsync(J, dJ, 2 * N * n * 8, cudaMemcpyDeviceToHost, p->stream));
  JCK(cudaStreamSynchronize(p->stream));
  cudaFree(dr); if (dJ) cudaFree(dJ);
  return 0;
}

// GaussNewtonOptimizer::optimize on the joint problem (util.rs:668-670). intr [C][d], extr [C][6] (row 0 ignored and
// returned as zeros, util.rs:689-690), poses [F][6] in/out. lo/hi/fixed are [C][d] (nullable): set_problem_parameter_bound /
// set_problem_parameter_disabled per camera (util.rs:654-663) and fix_variable("params0", 0) (util.rs:664-667).
int ccrs_joint_solve_gn(ccrs_joint* p, double* intr, double* extr, double* poses, const double* lo, const double* hi,
                        const unsigned char* fixed, const ccrs_options* opt_in, ccrs_summary* sum, double* err_hist) {
  if (!p || !intr || !extr || !poses) return jfail(CCRS_ERR_INVALID, "null");
  JCK(cudaSetDevice(p->device));
  ccrs_options opt;
  if (opt_in) opt = *opt_in; else ccrs_default_options(&opt);
  ccrs_summary local;
  if (!sum) sum = &local;
  std::memset(sum, 0, sizeof(*sum));
  const int d = p->d, C = p->n_cams, M = p->M, NS = p->NS, NV = NS + M + 1, F = p->n_frames;
  for (int i = 0; i < 6; ++i) extr[i] = 0.0;
  int st = upload_state(p, intr, extr, poses);
  if (st) return st;
  std::vector<double> red(NV), S((size_t)M * M), y(M);
  double last_err = 0.0;
  cudaEvent_t e0, e1;
  JCK(cudaEventCreate(&e0)); JCK(cudaEventCreate(&e1));
  JCK(cudaEventRecord(e0, p->stream));
  int status = 0;
  for (int it = 0; it < opt.max_iteration; ++it) {
    JointDev jd = p->dev();
    cudaError_t e = jdispatch(p->model, p->one_focal, [&](auto Mo, auto OF) {
      static const bool scalar = [] { const char* e = getenv("CCRS_JOINT_MMA"); return e && atoi(e) == 0; }();   // A/B switch
      if (scalar) k_joint_linearize<decltype(Mo)::value, decltype(OF)::value><<<(p->n_blocks + kJWarps - 1) / kJWarps, 32 * kJWarps, 0, p->stream>>>(jd, p->ij_table, p->jblk);
      else {
        auto kern = k_joint_linearize_mma<decltype(Mo)::value, decltype(OF)::value>;
        static bool configured = false;   // per instantiation
        if (!configured) {
          cudaError_t ce = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kJMmaSmem);
          if (ce != cudaSuccess) return ce;
          configured = true;
        }
        return jlaunch(kern, p->n_blocks, 32 * kJMmaWarps, kJMmaSmem, p->stream, jd, p->jblk);
      }
      return cudaGetLastError();
    });
    JCK(e);
    {
      const size_t smem = (size_t)kJSchurWarps * jschur_warp_doubles(M) * sizeof(double);
      static size_t configured = 0;
      if (smem > 48 * 1024 && smem > configured) {
        JCK(cudaFuncSetAttribute(k_joint_schur, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = smem;
      }
      JCK(jlaunch(k_joint_schur, (F + kJSchurWarps - 1) / kJSchurWarps, 32 * kJSchurWarps, smem, p->stream, jd, (const double*)p->jblk, (const uint16_t*)p->ij_table, p->NAJ, M, 0.0, p->fs, p->el));
    }
    JCK(cudaGetLastError());
    jarm(p->h_red, NV);
    JCK(jlaunch(k_joint_sum, (NV + kJSumWarps - 1) / kJSumWarps, 32 * kJSumWarps, 0, p->stream, (const double*)p->fs, F, NV, p->red, (volatile double*)p->h_red));
    JCK(cudaGetLastError());
    p->launches += 3;
    st = jwait(p, p->h_red, NV);     // mapped-memory publication: no memcpy, no stream synchronise
    if (st) return st;
    for (int i = 0; i < NV; ++i) red[i] = p->h_red[i];
    const double err = std::sqrt(red[NS + M]);
    if (err_hist) err_hist[it] = err;
    sum->iterations = it + 1; sum->final_error = err;
    if (err < opt.min_error) { sum->stop_reason = 1; break; }
    if (std::isnan(err)) { status = CCRS_ERR_NUMERIC; break; }
    if (it > 0) {
      if (std::fabs(last_err - err) < opt.min_abs_decrease) { sum->stop_reason = 2; break; }
      if (std::fabs(last_err - err) / last_err < opt.min_rel_decrease) { sum->stop_reason = 3; break; }
    }
    last_err = err;
    { int k = 0; for (int i = 0; i < M; ++i) for (int j = i; j < M; ++j) { S[(size_t)i * M + j] = red[k]; S[(size_t)j * M + i] = red[k]; ++k; } }
    for (int i = 0; i < M; ++i) y[i] = red[NS + i];
    bool poisoned = false;
    for (int i = 0; i < M * M; ++i) poisoned |= std::isnan(S[i]);
    if (fixed && opt.fixed_mode == 1)
      for (int i = 0; i < C * d; ++i) if (fixed[i]) { for (int j = 0; j < M; ++j) { S[(size_t)i * M + j] = 0; S[(size_t)j * M + i] = 0; } S[(size_t)i * M + i] = 1; y[i] = 0; }
    if (poisoned || !chol_solve_dense(S, M, y.data())) { status = CCRS_ERR_CHOLESKY; break; }
    // ParameterBlock::update_params per variable block
    for (int i = 0; i < C * d; ++i) {
      double v = intr[i] + y[i];
      if (lo && hi) v = std::min(std::max(v, lo[i]), hi[i]);
      if (fixed && fixed[i]) v = intr[i];
      intr[i] = v;
    }
    for (int c = 1; c < C; ++c) for (int i = 0; i < 6; ++i) extr[6 * c + i] += y[C * d + 6 * (c - 1) + i];
    {
      JointState js;
      const int n_head = C * d + C * 6;
      std::memcpy(js.v, intr, (size_t)C * d * 8);
      std::memcpy(js.v + C * d, extr, (size_t)C * 6 * 8);
      for (int i = 0; i < 6; ++i) js.v[C * d + i] = 0.0;   // cam0 is the reference frame (util.rs:689-690)
      std::memcpy(js.v + n_head, y.data(), (size_t)M * 8);
      JCK(jlaunch(k_joint_backsub, (F + 3) / 4, 128, 0, p->stream, jd, (const double*)p->el, js, n_head, p->state_dev, M));
    }
    JCK(cudaGetLastError());
    p->launches++;
  }
  JCK(cudaEventRecord(e1, p->stream));
  JCK(cudaEventSynchronize(e1));
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  sum->device_ms = ms;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  sum->status = status;
  JCK(cudaMemcpyAsync(poses, p->poses, (size_t)F * 48, cudaMemcpyDeviceToHost, p->stream));
  JCK(cudaStreamSynchronize(p->stream));
  if (status == CCRS_ERR_CHOLESKY) jfail(status, "Cholesky failure (non-positive pivot)");
  if (status == CCRS_ERR_NUMERIC) jfail(status, "NaN error");
  return status;
}

}  // extern "C"

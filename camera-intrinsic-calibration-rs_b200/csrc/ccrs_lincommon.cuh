// ccrs_lincommon.cuh — pieces shared by the two K2 translation units (ccrs_kernels.cu: register-accumulator variants,
// ccrs_linmma.cu: FP64 tensor-core Gram variant): compile-time shape of a (model, one_focal) instantiation, the
// per-observation rows of [J | r] (ReprojectionFactor::residual_func, src/optimization/factors.rs:152-173, analytic
// instead of num-dual), and the fixed-order finalisation of the fused {model decrease, cost} statistics.
#pragma once
#include "ccrs_devutil.cuh"

#include <algorithm>
#include <cstddef>
#include <type_traits>

namespace ccrs {

template <int B, int E, class F>
CCRS_D void static_for(F&& f) {
  if constexpr (B < E) {
    f(std::integral_constant<int, B>{});
    static_for<B + 1, E>(f);
  }
}

// Compile-time shape of one (model, one_focal) instantiation.
// Column order of a row of [J | r]: [intrinsics (D) | phi/rvec (3) | tvec (3) | r].
// Structural sparsity: the u-row never touches fy,cy and the v-row never touches fx,cx
// (u = fx*mx + cx, v = fy*my + cy), so those products are never formed nor stored in registers.
template <int MODEL, bool OF>
struct Cfg {
  static constexpr int ND = model_nd(MODEL);
  static constexpr int DFULL = 4 + ND;
  static constexpr int D = DFULL - (OF ? 1 : 0);
  static constexpr int N = D + 6;
  static constexpr int NA = N + 1;
  static constexpr int NBLK = NA * (NA + 1) / 2;
  static constexpr int KOFF = OF ? 3 : 4;  // first distortion column
  CCRS_HD static constexpr bool nzu(int c) { return OF ? (c != 2) : (c != 1 && c != 3); }
  CCRS_HD static constexpr bool nzv(int c) { return OF ? (c != 1) : (c != 0 && c != 2); }
  CCRS_HD static constexpr bool hasu(int i, int j) { return nzu(i) && nzu(j); }
  CCRS_HD static constexpr bool hasv(int i, int j) { return nzv(i) && nzv(j); }
  CCRS_HD static constexpr bool has(int i, int j) { return hasu(i, j) || hasv(i, j); }
  // index of (i<=j) among the structurally non-zero upper entries, row-major; -1 if structurally zero
  CCRS_HD static constexpr int kidx(int i, int j) {
    if (!has(i, j)) return -1;
    int k = 0;
    for (int a = 0; a < NA; ++a)
      for (int b = a; b < NA; ++b) {
        if (a == i && b == j) return k;
        if (has(a, b)) ++k;
      }
    return -1;
  }
  CCRS_HD static constexpr int nacc() {
    int k = 0;
    for (int a = 0; a < NA; ++a)
      for (int b = a; b < NA; ++b)
        if (has(a, b)) ++k;
    return k;
  }
  static constexpr int NACC = nacc();
  // compact rows: the structurally non-zero columns of the u-row / v-row of [J | r], in column order. Both rows have
  // the same shape [2 row intrinsics (fx cx | fy cy; one focal: f cx | f cy) | distortion | phi | t | r].
  CCRS_HD static constexpr bool inu(int c) { return c >= D || nzu(c); }
  CCRS_HD static constexpr bool inv(int c) { return c >= D || nzv(c); }
  CCRS_HD static constexpr int count_u() { int n = 0; for (int c = 0; c < NA; ++c) n += inu(c) ? 1 : 0; return n; }
  static constexpr int NU = count_u();     // == number of v-row columns
  CCRS_HD static constexpr int ucol(int k) { int n = 0; for (int c = 0; c < NA; ++c) if (inu(c)) { if (n == k) return c; ++n; } return -1; }
  CCRS_HD static constexpr int vcol(int k) { int n = 0; for (int c = 0; c < NA; ++c) if (inv(c)) { if (n == k) return c; ++n; } return -1; }
};

// Lane-pair variant of K2 for the models whose merged Gram block does not fit the register file (EUCMT, KB4, OPENCV5,
// FTHETA: 104-132 FP64 accumulators = 208-264 registers, i.e. heavy local-memory spilling): the even lane of a pair
// accumulates only the u-row products, the odd lane only the v-row products, of BOTH lanes' observations; the rows are
// swapped with one shuffle per value. One compact row has the shape of a problem with NU columns:
constexpr int kPairThreshold = 100;   // merged accumulators above which the pair variant is used
template <class C>
struct RowCfg {
  static constexpr int NA = C::NU;
  static constexpr int D = C::NU - 7;   // row intrinsics + distortion, then phi(3) t(3) r(1)
  static constexpr int N = NA - 1;
  static constexpr int NACC = NA * (NA + 1) / 2;
  CCRS_HD static constexpr int kidx(int i, int j) { return tri_idx(NA, i, j); }
};
template <int MODEL, bool OF>
constexpr bool lin_pair_v = Cfg<MODEL, OF>::NACC > kPairThreshold;

CCRS_D int cur_of(const ProblemDev& pb, int prob) { return pb.cur ? pb.cur[prob] : pb.cur_val; }

// publish n doubles to mapped host memory: plain stores, no fence — the host armed the n words with a sentinel and
// spins until all of them changed (ccrs_api.cu: arm_payload / wait_payload)
CCRS_D void publish_host(volatile double* dst, const double* src, int n) {
  for (int i = 0; i < n; ++i) dst[i] = src[i];
}

// observation k of an SoA array that holds doubles or (f32 != 0) floats; f32 -> f64 widening as factors.rs:141-143
CCRS_D double ld_obs(const double* base, int k, int f32) {
  return f32 ? (double)reinterpret_cast<const float*>(base)[k] : base[k];
}
// One observation: weighted rows au, av of [J | r] in the LOCAL rotation basis (d/dphi, not d/drvec).
// Returns the corrected squared residual. Structurally-zero entries of au/av are left untouched.
template <int MODEL, bool OF, bool WITH_J>
CCRS_D double obs_rows(const double* __restrict__ ip /* full intrinsics */, const double* __restrict__ fc /* R t */,
                       double px, double py, double pz, double ou, double ov, double delta,
                       double* __restrict__ au, double* __restrict__ av) {
  using C = Cfg<MODEL, OF>;
  const double fx = ip[0], fy = OF ? ip[0] : ip[1], cx = ip[2], cy = ip[3];
  // P = R p + t as one FMA chain per component, q = R p recovered off the critical path
  // (Isometry3::new(tvec, rvec) * p3d, factors.rs:162-163)
  const double X = fma(fc[0], px, fma(fc[1], py, fma(fc[2], pz, fc[9])));
  const double Y = fma(fc[3], px, fma(fc[4], py, fma(fc[5], pz, fc[10])));
  const double Z = fma(fc[6], px, fma(fc[7], py, fma(fc[8], pz, fc[11])));
  const double qx = X - fc[9], qy = Y - fc[10], qz = Z - fc[11];
  if constexpr (MODEL == UCM || MODEL == EUCM) {
    // Fused UCM / EUCM path (project_one, factors.rs:165). The dependent chain of an observation is
    //   rho2 -> 1/rho -> n -> 1/n -> m -> r -> s -> Huber w -> weighted Jacobian rows,
    // shortened by (i) seeding 1/n from an FP32 evaluation of n that runs beside the FP64 1/rho refinement, and
    // (ii) forming the rows directly from a = w f / n and b = a m instead of scaling an unweighted Jacobian.
    const double alpha = ip[4];
    const double beta = (MODEL == UCM) ? 1.0 : ip[5];
    const double oma = 1.0 - alpha;
    const double r2 = fma(X, X, Y * Y);
    const double rho2 = fma(beta, r2, Z * Z);
    const float rho2f = (float)rho2;
    float rhof;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rhof) : "f"(rho2f));
    float inf;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inf) : "f"(fmaf((float)alpha, rhof, (float)oma * (float)Z)));
    const double irho = rsqrt_fast(rho2);
    const double rho = rho2 * irho;
    const double nrm = fma(alpha, rho, oma * Z);
    const double in = rcp_refine(nrm, (double)inf);
    const double mx = X * in, my = Y * in;
    const double ru = fma(fx, mx, cx) - ou;                 // - p2d (factors.rs:167-171)
    const double rv = fma(fy, my, cy) - ov;
    const double s = fma(ru, ru, rv * rv);
    const double w = huber_weight(s, delta);                // Corrector: r *= sqrt(rho'), J *= sqrt(rho')
    if constexpr (WITH_J) {
      // dn/dP = (alpha beta x / rho, alpha beta y / rho, alpha z / rho + 1 - alpha); dn/dalpha = rho - z; dn/dbeta = alpha r2 / (2 rho)
      const double a_irho = alpha * irho, ab_irho = (MODEL == UCM) ? a_irho : a_irho * beta;
      const double nx = ab_irho * X, ny = ab_irho * Y, nz = fma(a_irho, Z, oma);
      const double na = rho - Z;
      const double au_a = (w * fx) * in, av_a = (w * fy) * in;     // a = w f / n
      const double nbu = -(au_a * mx), nbv = -(av_a * my);         // -b = -a m
      const double du0 = fma(nbu, nx, au_a), du1 = nbu * ny, du2 = nbu * nz;
      const double dv0 = nbv * nx, dv1 = fma(nbv, ny, av_a), dv2 = nbv * nz;
      if constexpr (OF) {
        au[0] = w * mx; av[0] = w * my;  // shared focal column
        au[1] = w;                        // cx
        av[2] = w;                        // cy
      } else {
        au[0] = w * mx; av[1] = w * my;
        au[2] = w; av[3] = w;
      }
      au[C::KOFF] = nbu * na; av[C::KOFF] = nbv * na;
      if constexpr (MODEL == EUCM) {
        const double nb = (0.5 * alpha) * (r2 * irho);
        au[C::KOFF + 1] = nbu * nb; av[C::KOFF + 1] = nbv * nb;
      }
      au[C::D + 0] = qy * du2 - qz * du1; au[C::D + 1] = qz * du0 - qx * du2; au[C::D + 2] = qx * du1 - qy * du0;
      av[C::D + 0] = qy * dv2 - qz * dv1; av[C::D + 1] = qz * dv0 - qx * dv2; av[C::D + 2] = qx * dv1 - qy * dv0;
      au[C::D + 3] = du0; au[C::D + 4] = du1; au[C::D + 5] = du2;
      av[C::D + 3] = dv0; av[C::D + 4] = dv1; av[C::D + 5] = dv2;
      au[C::N] = w * ru; av[C::N] = w * rv;
    }
    return s * (w * w);
  } else {
  double m[2], dP[2][3], dk[2][kMaxNd];
  model_eval<MODEL, WITH_J>(ip + 4, X, Y, Z, m, dP, dk);   // project_one (factors.rs:165)
  const double ru = fma(fx, m[0], cx) - ou;                 // - p2d (factors.rs:167-171)
  const double rv = fma(fy, m[1], cy) - ov;
  const double s = ru * ru + rv * rv;
  const double w = huber_weight(s, delta);                  // Corrector: r *= sqrt(rho'), J *= sqrt(rho')
  if constexpr (WITH_J) {
    const double wfx = w * fx, wfy = w * fy;
    const double du0 = wfx * dP[0][0], du1 = wfx * dP[0][1], du2 = wfx * dP[0][2];
    const double dv0 = wfy * dP[1][0], dv1 = wfy * dP[1][1], dv2 = wfy * dP[1][2];
    if constexpr (OF) {
      au[0] = w * m[0]; av[0] = w * m[1];  // shared focal column
      au[1] = w;                            // cx
      av[2] = w;                            // cy
    } else {
      au[0] = w * m[0]; av[1] = w * m[1];
      au[2] = w; av[3] = w;
    }
#pragma unroll
    for (int j = 0; j < C::ND; ++j) { au[C::KOFF + j] = wfx * dk[0][j]; av[C::KOFF + j] = wfy * dk[1][j]; }
    // d/dphi row = (q x d)^T   since d(Exp(phi) q)/dphi = -[q]x
    au[C::D + 0] = qy * du2 - qz * du1; au[C::D + 1] = qz * du0 - qx * du2; au[C::D + 2] = qx * du1 - qy * du0;
    av[C::D + 0] = qy * dv2 - qz * dv1; av[C::D + 1] = qz * dv0 - qx * dv2; av[C::D + 2] = qx * dv1 - qy * dv0;
    au[C::D + 3] = du0; au[C::D + 4] = du1; au[C::D + 5] = du2;
    av[C::D + 3] = dv0; av[C::D + 4] = dv1; av[C::D + 5] = dv2;
    au[C::N] = w * ru; av[C::N] = w * rv;
  }
  return s * w * w;
  }
}


// The hot window of the control block (LoopCtl, ccrs_kernels.cuh) read by a whole warp with ONE lane-distributed load;
// the fields are then handed to all lanes by shuffle.
struct CtlHot { int phase, cur, lm; double u_used; };
template <int D>
CCRS_D CtlHot load_ctl_hot(const LoopCtl* ctl, int lane, double (&trial)[D], double (&step)[D]) {
  const double* hot = reinterpret_cast<const double*>(ctl) + offsetof(LoopCtl, phase) / 8;
  const double w = lane < kCtlHotWords ? __ldcg(hot + lane) : 0.0;
  const long long w0 = __double_as_longlong(__shfl_sync(0xffffffffu, w, 0));
  const long long w1 = __double_as_longlong(__shfl_sync(0xffffffffu, w, 1));
  CtlHot h;
  h.phase = (int)(w0 & 0xffffffffLL); h.cur = (int)(w0 >> 32); h.lm = (int)(w1 & 0xffffffffLL);
  h.u_used = __shfl_sync(0xffffffffu, w, 2);
#pragma unroll
  for (int a = 0; a < D; ++a) { trial[a] = __shfl_sync(0xffffffffu, w, 3 + a); step[a] = __shfl_sync(0xffffffffu, w, 12 + a); }
  return h;
}

// Executed by the whole warp that took the last ticket: every producer of a {model decrease, cost} partial has taken
// its ticket, so every partial store has been issued. Sum the n_parts slots in a fixed order, exchange across GPUs,
// publish.
CCRS_D void stats_finalize(const LinParams& prm, unsigned n_parts, int lane, int phase) {
  // every warp has taken its ticket, so every partial store has been issued: read the slots from L2 (16 loads
  // in flight per lane) until none still holds the arming pattern, sum in a fixed order, re-arm for the next launch
  double2* part = reinterpret_cast<double2*>(prm.cta_part);
  double a = 0.0, b = 0.0;
  constexpr int kBatch = 16;   // loads in flight per lane (40 would cover 7,000 frames in one pass, but the 160 live registers cost the main loop 6 us)
  const long long t_spin = clock64();
  for (unsigned w0 = lane; w0 < n_parts; w0 += 32 * kBatch) {
    double2 t[kBatch];
    bool ok;
    do {
      ok = true;
#pragma unroll
      for (int q = 0; q < kBatch; ++q) {
        const unsigned w = w0 + 32 * q;
        t[q] = w < n_parts ? ld_spin2(part + w) : make_double2(0.0, 0.0);
        ok = ok && (__double_as_longlong(t[q].x) != kArmBits) && (__double_as_longlong(t[q].y) != kArmBits);
      }
      if (!ok && clock64() - t_spin > 4000000000LL) {   // ~2 s: a partial never arrived -> poison the result, not a hang
#pragma unroll
        for (int q = 0; q < kBatch; ++q) t[q] = make_double2(nan(""), nan(""));
        ok = true;
      }
    } while (!ok);
#pragma unroll
    for (int q = 0; q < kBatch; ++q) {
      a += t[q].x; b += t[q].y;
      const unsigned w = w0 + 32 * q;
      if (w < n_parts) part[w] = make_double2(__longlong_as_double(kArmBits), __longlong_as_double(kArmBits));
    }
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
  }
  if (prm.ctl) {
    // device-driven loop: this rank's sums stay on the device; the next K3 exchanges them together with the reduced
    // system (one cross-GPU exchange per iteration) and takes the accept / reject decision
    if (lane == 0) {
      prm.ctl->stat[0] = a; prm.ctl->stat[1] = b;
      prm.ctl->t_k2_end = stamp_ns();
      prm.ctl->phase = (phase == PH_LIN0 && prm.ctl->mode == 1) ? PH_REDUCE : PH_DECIDE;
      *prm.ticket = 0u;
    }
    return;
  }
  if (prm.px.world > 1) {
    const double mine = lane == 0 ? a : b;
    const double tot = lane < 2 ? peer_exchange(prm.px, lane, mine) : 0.0;
    a = __shfl_sync(0xffffffffu, tot, 0);
    b = __shfl_sync(0xffffffffu, tot, 1);
  }
  if (lane == 0) {
    prm.stat_dev[0] = a; prm.stat_dev[1] = b;
    *prm.ticket = 0u;
    if (prm.host_stat) { double tmp[2] = {a, b}; publish_host(prm.host_stat, tmp, 2); }
  }
}

}  // namespace ccrs

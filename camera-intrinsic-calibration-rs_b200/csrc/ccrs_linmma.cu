// ccrs_linmma.cu — K2 for the camera models whose Gram block does not fit a lane's registers (EUCMT, KB4, OPENCV5,
// FTHETA: d + 7 = 14..16 columns), with the per-frame normal-equation block accumulated on the FP64 tensor path.
//
// What it replaces in the reference is the same as k_linearize (ccrs_kernels.cu): tiny-solver's
// compute_residual_and_jacobian over ReprojectionFactor::residual_func (src/optimization/factors.rs:152-173) plus the
// J^T J / J^T r products (call sites src/util.rs:455,463,670).
//
// Organisation: ONE WARP PER FRAME (a warp walks its frames one after the other). The 32 lanes evaluate 32
// observations of the frame at a time (model chain, Huber corrector, the two weighted rows of [J | r], already in the
// rvec basis), drop the rows into the warp's shared-memory staging buffer, and the warp accumulates
//     H (16 x 16, upper three 8 x 8 tiles)  +=  J^T J ,   J = the 4 rows of two observations
// with mma.sync.m8n8k4.f64: the A and the B fragment of a tile are the same staged values (lane l holds row l % 4,
// column 8 b + l / 4), the accumulators are 6 doubles per lane — against 91-105 per lane in the register-accumulator
// variants, which is what limited those to two warps per sub-partition with the model chain's latency exposed. Here
// 16 warps are resident per SM. The dense tiles cost 768 FMA per observation instead of 182, on the same FP64 engine
// (tools/microbench/fp64_dmma.cu), and still come out ahead: the pipe is kept busy instead of waiting.
// Deterministic: fixed accumulation order, no atomics on floating-point data.
#include "ccrs_lincommon.cuh"

namespace ccrs {

constexpr int kMmaColStride = 72;   // doubles between two columns of the staging buffer (72 * 8 B = 64 mod 128: see below)
constexpr int kMmaRowStride = 18;   // doubles between the four row classes (u / v of the even / odd observation of a pair)
constexpr int kMmaBuf = 16 * kMmaColStride;   // doubles of the staging buffer
constexpr int kMmaCtasPerSm = 4;
// staging buffer + frame constants (21) + intrinsics (<= 10), rounded so that every warp's buffer stays 128-byte aligned
CCRS_HD constexpr int mma_warp_smem_doubles() { return kMmaBuf + 48; }
static_assert(kFrameConst + kMaxFull + 1 + kMaxFull <= 48, "per-warp constants do not fit");

CCRS_D void dmma884(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODEL, bool OF, bool BATCH, bool F32>
__global__ void __launch_bounds__(kLinThreads, kMmaCtasPerSm) k_linearize_mma(const __grid_constant__ LinParams prm) {
  using C = Cfg<MODEL, OF>;
  static_assert(C::NA <= 16 && C::N >= 8, "three 8 x 8 tiles cover the block");
  extern __shared__ double smem[];
  const ProblemDev& pb = prm.pb;
  asm volatile("griddepcontrol.launch_dependents;");
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x * kLinWarps + wid;
#ifdef CCRS_K2_TIMING
  unsigned long long gt0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
  const long long tk_start = clock64();
  long long tk_pro = 0, tk_loop = 0, tk_epi = 0, tk_iters = 0, tk_mark = 0;
#define CCRS_TKM(acc) { const long long t_ = clock64(); acc += t_ - tk_mark; tk_mark = t_; }
#else
#define CCRS_TKM(acc)
#endif
  double* const s_rows = smem + (size_t)wid * mma_warp_smem_doubles();   // [16 columns][4 row classes][16 pairs] (padded)
  double* const s_fc = s_rows + kMmaBuf;                                  // R t Jl of the frame
  double* const s_intr = s_fc + kFrameConst;                              // full intrinsic vector
  double* const s_ya = s_intr + kMaxFull + 1;                             // intrinsic step of the pending back-substitution (single problem)
  // Staging layout: value of column c for row class q (0: u of the even observation of a pair, 1: its v, 2 / 3: the odd
  // observation) and pair g sits at c * 72 + q * 18 + g. Lane l reads, as one 16-byte load, pairs g, g + 1 of class
  // l % 4 and column 8 b + l / 4: the eight lanes of a quarter-warp hit 16-byte bank groups 0..7 (18 * 8 B = 16 mod 128,
  // 72 * 8 B = 64 mod 128): conflict-free. Lane o writes class 2 (o % 2), pair o / 2: 16 consecutive doubles per class.
  // Columns >= NA and the structural zeros of the rows (fy, cy in a u-row; fx, cx in a v-row) are zeroed once.
  for (int i = lane; i < kMmaBuf; i += 32) s_rows[i] = 0.0;
  const int q_rd = lane & 3, c_rd = lane >> 2;
  const double* const rd0 = s_rows + c_rd * kMmaColStride + q_rd * kMmaRowStride;
  const double* const rd1 = rd0 + 8 * kMmaColStride;
  double* const wr_u = s_rows + (2 * (lane & 1)) * kMmaRowStride + (lane >> 1);
  double* const wr_v = wr_u + kMmaRowStride;

  // ---- which point, which step: from the launch (host-driven) or the device control block ----
  int which = prm.which, backsub = prm.backsub, phase = -1, cur = pb.cur_val;
  double u_bs = prm.u;
  if constexpr (!BATCH) {
    if (prm.ctl) {
      asm volatile("griddepcontrol.wait;" ::: "memory");
      const LoopCtl* ctl = prm.ctl;
      if (gw == 0 && lane == 0) prm.ctl->t_k2_wake = stamp_ns();
      double lin_intr[C::D], ya_ld[C::D];
      const CtlHot hot = load_ctl_hot<C::D>(ctl, lane, lin_intr, ya_ld);   // one request per warp
      const int ph = hot.phase, lm = hot.lm, cu = hot.cur;
      u_bs = hot.u_used;
      if (lane == 0) {
#pragma unroll
        for (int a = 0; a < C::D; ++a) s_ya[a] = ya_ld[a];
      }
      phase = ph;
      if (phase != PH_LIN0 && phase != PH_TRIAL) return;   // not this slot's turn (re-reduction pending, or the loop is done)
      if (gw == 0 && lane == 0) prm.ctl->t_k2_begin = stamp_ns();
      cur = cu;
      backsub = phase == PH_LIN0 ? 0 : (lm ? 1 : 2);
      which = (phase == PH_TRIAL && lm) ? 1 : 0;
      if (lane == 0) {
        if constexpr (OF) { s_intr[0] = lin_intr[0]; s_intr[1] = lin_intr[0]; for (int i = 1; i < C::D; ++i) s_intr[i + 1] = lin_intr[i]; }
        else { for (int i = 0; i < C::D; ++i) s_intr[i] = lin_intr[i]; }
      }
    } else {
      if (lane < C::D) s_ya[lane] = prm.y_a[lane];
      if (lane == 0) {
#pragma unroll
        for (int i = 0; i < C::DFULL; ++i) s_intr[i] = prm.intr[i];
      }
    }
  }
  __syncwarp();   // s_ya, s_intr and the zeroed staging buffer are visible to the whole warp
  // Frames are handed out one at a time from a device counter: a warp that the scheduler favours simply takes more of
  // them, so all warps finish within one frame of each other (with a static split the slowest warp of a sub-partition ran
  // alone, latency-bound, for a third of the kernel). Every warp makes exactly one failed grab, so the counter has seen
  // n_frames + n_warps grabs when the launch is over: the warp that makes the last one resets it.
  const unsigned long long n_grabs = (unsigned long long)pb.n_frames + (unsigned long long)gridDim.x * kLinWarps;
  // issued by lane 0 (the value is consumed later: the atomic's round trip overlaps whatever comes in between)
  auto grab_issue = [&]() -> unsigned long long {
    unsigned long long t = 0;
    if (lane == 0) t = atomicAdd(prm.frame_ctr, 1ull);
    return t;
  };
  auto grab_take = [&](unsigned long long t) -> int {
    if (lane == 0 && t == n_grabs - 1) *prm.frame_ctr = 0ull;   // the launch's last grab
    t = __shfl_sync(0xffffffffu, t, 0);
    return t < (unsigned long long)pb.n_frames ? (int)t : -1;
  };
#ifdef CCRS_K2_TIMING
  const long long tk_setup = clock64();
  tk_mark = tk_setup;
  long long tk_frames = 0;
#endif
  for (int f = grab_take(grab_issue()); f >= 0;) {
#ifdef CCRS_K2_TIMING
    ++tk_frames;
#endif
    const int fo_beg = __ldg(pb.frame_offsets + f);
    int fo_end = __ldg(pb.frame_offsets + f + 1);
    int prob = 0;
    bool moves = backsub != 0;
    if constexpr (BATCH) {
      prob = pb.frame_problem[f];
      cur = cur_of(pb, prob);
      if (prm.active && !prm.active[prob]) { fo_end = fo_beg; moves = false; }   // its problem has stopped
    }
    // first observations of the frame: in flight during the pose prologue
    auto ldo = [&](const double* base, int k) -> double {
      if constexpr (F32) return (double)__ldg(reinterpret_cast<const float*>(base) + k);
      else return __ldg(base + k);
    };
    const int last = max(fo_end - 1, fo_beg);
    const int k0 = min(fo_beg + lane, last);
    double ox = 0.0, oy = 0.0, oz = 1.0, ou = 0.0, ov = 0.0;
    if (fo_end > fo_beg) { ox = ldo(pb.x, k0); oy = ldo(pb.y, k0); oz = ldo(pb.z, k0); ou = ldo(pb.u, k0); ov = ldo(pb.v, k0); }

    // ---- pose of the frame, fused K4: y_p = cg - X y_a ; pose += D_p y_p ; model decrease y_p^T g'_p + u sum dd_i y_p,i^2.
    //      Lane i < 6 owns pose component i (its row of the elimination record); the components are then broadcast.
    double rt[6];
    double md = 0.0;
    {
      const int pi = lane < 6 ? lane : 0;
      const double* src = pb.poses[backsub ? cur : (cur ^ which)] + 6 * (size_t)f;
      double rt_l = BATCH ? src[pi] : __ldcg(src + pi);
      double md_l = 0.0;
      if (backsub && moves) {
        const size_t Fs = pb.Fs;
        const double* el = prm.elim + f;
        const double* elx = el + (size_t)(pi * C::D) * Fs;
        const double* ya = BATCH ? prm.ya_dev + (size_t)prob * C::D : s_ya;
        const double u = BATCH ? (prm.u_dev ? prm.u_dev[prob] : 0.0) : u_bs;
        double yp = __ldcg(el + (size_t)(6 * C::D + pi) * Fs);
        double xv[C::D];
#pragma unroll
        for (int a = 0; a < C::D; ++a) xv[a] = __ldcg(elx + (size_t)a * Fs);
        const double gp = __ldcg(el + (size_t)(6 * C::D + 6 + pi) * Fs), dd = __ldcg(el + (size_t)(6 * C::D + 12 + pi) * Fs);
        const double sp = prm.pose_scale ? __ldcg(prm.pose_scale + (size_t)pi * Fs + f) : 1.0;
#pragma unroll
        for (int a = 0; a < C::D; ++a) yp -= xv[a] * ya[a];
        rt_l += sp * yp;
        md_l = yp * gp + u * dd * yp * yp;
      }
      if (backsub && lane < 6) pb.poses[backsub == 2 ? cur : (cur ^ 1)][6 * (size_t)f + lane] = rt_l;
#pragma unroll
      for (int i = 0; i < 6; ++i) { rt[i] = __shfl_sync(0xffffffffu, rt_l, i); md += __shfl_sync(0xffffffffu, md_l, i); }
      if (BATCH && backsub && lane == 0 && prm.frame_md) prm.frame_md[f] = md;
    }
    FramePose fp;
    pose_from_rvec_tvec(rt, fp);
    __syncwarp();   // the previous frame's readers of s_fc / s_intr are done
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < 9; ++i) s_fc[i] = fp.R[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) s_fc[9 + i] = fp.t[i];
#pragma unroll
      for (int i = 0; i < 9; ++i) s_fc[12 + i] = fp.Jl[i];
      if constexpr (BATCH) {
        const double* a = prm.intr_dev + (size_t)prob * C::D;
        if constexpr (OF) { s_intr[0] = a[0]; s_intr[1] = a[0]; for (int i = 1; i < C::D; ++i) s_intr[i + 1] = a[i]; }
        else { for (int i = 0; i < C::D; ++i) s_intr[i] = a[i]; }
      }
    }
    __syncwarp();

    // two accumulator sets (even / odd pairs): independent DMMA chains, summed in a fixed order at the end
    double ca[6], cb[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) { ca[i] = 0.0; cb[i] = 0.0; }

    // rows of the observation held in (ox .. ov) -> staging buffer; `valid` = the lane holds an observation of the frame
    auto stage_rows = [&](auto CHECKED, bool valid) {
      double au[C::NA], av[C::NA];
      obs_rows<MODEL, OF, true>(s_intr, s_fc, ox, oy, oz, ou, ov, pb.huber_delta, au, av);
      // d/drvec = d/dphi J_l
      const double* Jl = s_fc + 12;
      double ur[3], vr[3];
#pragma unroll
      for (int c = 0; c < 3; ++c) {
        ur[c] = fma(au[C::D], Jl[c], fma(au[C::D + 1], Jl[3 + c], au[C::D + 2] * Jl[6 + c]));
        vr[c] = fma(av[C::D], Jl[c], fma(av[C::D + 1], Jl[3 + c], av[C::D + 2] * Jl[6 + c]));
      }
#pragma unroll
      for (int c = 0; c < 3; ++c) { au[C::D + c] = ur[c]; av[C::D + c] = vr[c]; }
      if (decltype(CHECKED)::value && !valid) {
        static_for<0, C::NA>([&](auto Cc) {
          constexpr int c = decltype(Cc)::value;
          if constexpr (C::inu(c)) wr_u[c * kMmaColStride] = 0.0;
          if constexpr (C::inv(c)) wr_v[c * kMmaColStride] = 0.0;
        });
      } else {
        static_for<0, C::NA>([&](auto Cc) {
          constexpr int c = decltype(Cc)::value;
          if constexpr (C::inu(c)) wr_u[c * kMmaColStride] = au[c];
          if constexpr (C::inv(c)) wr_v[c * kMmaColStride] = av[c];
        });
      }
    };
    // pairs g, g + 1 of the staged rows -> the three tiles
    auto mma_pairs = [&](int g) {
      const double2 a0 = *reinterpret_cast<const double2*>(rd0 + g);
      const double2 a1 = *reinterpret_cast<const double2*>(rd1 + g);
      dmma884(ca[0], ca[1], a0.x, a0.x); dmma884(ca[2], ca[3], a0.x, a1.x); dmma884(ca[4], ca[5], a1.x, a1.x);
      dmma884(cb[0], cb[1], a0.y, a0.y); dmma884(cb[2], cb[3], a0.y, a1.y); dmma884(cb[4], cb[5], a1.y, a1.y);
    };

    CCRS_TKM(tk_pro);
    unsigned long long next_raw = 0;
    bool next_issued = false;
    for (int base = fo_beg; base < fo_end; base += 32) {
      if (base + 32 >= fo_end) { next_raw = grab_issue(); next_issued = true; }   // last round: ask for the next frame now
#ifdef CCRS_K2_TIMING
      ++tk_iters;
#endif
      const int n_here = min(32, fo_end - base);
      // next 32 observations: issued before this round's chain
      const int kn = min(base + 32 + lane, last);
      const bool more = base + 32 < fo_end;
      double nx = 0.0, ny = 0.0, nz = 1.0, nu = 0.0, nv = 0.0;
      if (more) { nx = ldo(pb.x, kn); ny = ldo(pb.y, kn); nz = ldo(pb.z, kn); nu = ldo(pb.u, kn); nv = ldo(pb.v, kn); }
      if (n_here == 32) {
        stage_rows(std::false_type{}, true);
        __syncwarp();
#pragma unroll
        for (int g = 0; g < 16; g += 2) mma_pairs(g);
      } else {
        stage_rows(std::true_type{}, lane < n_here);
        __syncwarp();
        const int np2 = (n_here + 3) >> 2;   // pairs of pairs that hold at least one observation (the rest of them is zero)
        for (int g2 = 0; g2 < np2; ++g2) mma_pairs(2 * g2);
      }
      __syncwarp();   // the staging buffer may be overwritten
      ox = nx; oy = ny; oz = nz; ou = nu; ov = nv;
    }
    CCRS_TKM(tk_loop);
    // ---- frame block: even + odd pair sets, stored SoA; frame cost = its (r, r) entry ----
    double cs[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) cs[i] = ca[i] + cb[i];
    {
      // block entries this lane holds: tiles (0,0), (0,1), (1,1), two columns each (index arithmetic per frame: keeping
      // the six offsets live across the main loop costs the registers that make it spill)
      double* const out = pb.blocks[cur ^ which] + f;
      const size_t Fs = pb.Fs;
      const int ti = lane >> 2, tj = 2 * (lane & 3);
#pragma unroll
      for (int t = 0; t < 3; ++t)
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int gi = ti + (t == 2 ? 8 : 0), gj = tj + e + (t >= 1 ? 8 : 0);
          if (gi <= gj && gj < C::NA && gi < C::NA) out[(size_t)tri_idx(C::NA, gi, gj) * Fs] = cs[2 * t + e];
        }
    }
    constexpr int ri = C::N - 8;   // (r, r) sits in tile (1,1) at (ri, ri): lane 4 ri + ri / 2, element ri % 2
    const double fcost = __shfl_sync(0xffffffffu, cs[4 + (ri & 1)], 4 * ri + (ri >> 1));
    const int f_done = f;
    if (!next_issued) next_raw = grab_issue();   // a frame without observations
    // ---- fused statistics (single problem): {model decrease, cost} per frame in self-validating slots; the warp that
    //      completes a chunk of 16 consecutive frames sums the chunk in frame order, the warp that completes the last
    //      chunk sums the chunks in chunk order (stats_finalize): a fixed order whoever computed what ----
    if constexpr (!BATCH) {
      const int chunk = f_done >> 4, c_beg = chunk << 4;
      const int c_cnt = min(16, pb.n_frames - c_beg);
      unsigned oldc = 0;
      if (lane == 0) {
        reinterpret_cast<double2*>(prm.frame_stat)[f_done] = make_double2(md, fcost);
        oldc = atomicAdd(prm.chunk_cnt + chunk, 1u);
      }
      f = grab_take(next_raw);
      if (__shfl_sync(0xffffffffu, (unsigned)(oldc == (unsigned)c_cnt - 1), 0)) {
        double2* slot = reinterpret_cast<double2*>(prm.frame_stat) + c_beg + (lane < c_cnt ? lane : 0);
        double2 v = make_double2(0.0, 0.0);
        const long long t_spin = clock64();
        bool ok;
        do {
          v = ld_spin2(slot);
          ok = (__double_as_longlong(v.x) != kArmBits) && (__double_as_longlong(v.y) != kArmBits);
          if (!ok && clock64() - t_spin > 4000000000LL) { v = make_double2(nan(""), nan("")); ok = true; }   // ~2 s: poison, not a hang
        } while (!__all_sync(0xffffffffu, ok));
        if (lane < c_cnt) *slot = make_double2(__longlong_as_double(kArmBits), __longlong_as_double(kArmBits));
        double sm = 0.0, sc = 0.0;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const double a = __shfl_sync(0xffffffffu, v.x, i), b = __shfl_sync(0xffffffffu, v.y, i);
          if (i < c_cnt) { sm += a; sc += b; }
        }
        const unsigned n_chunks = (unsigned)((pb.n_frames + 15) >> 4);
        unsigned t_old = 0;
        if (lane == 0) {
          prm.chunk_cnt[chunk] = 0u;
          reinterpret_cast<double2*>(prm.cta_part)[chunk] = make_double2(sm, sc);
          t_old = atomicAdd(prm.ticket, 1u);
        }
        if (__shfl_sync(0xffffffffu, (unsigned)(t_old == n_chunks - 1), 0)) stats_finalize(prm, n_chunks, lane, phase);
      }
    }
    if constexpr (BATCH) f = grab_take(next_raw);
    CCRS_TKM(tk_epi);
  }
#ifdef CCRS_K2_TIMING
  if (prm.dbg && lane == 0) {
    long long* o = prm.dbg + (size_t)gw * 12;
    unsigned smid, warpid;
    unsigned long long gt1;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
    o[0] = (long long)gt0; o[1] = smid; o[2] = tk_start; o[3] = tk_setup; o[4] = tk_pro; o[5] = tk_loop; o[6] = tk_epi;
    o[7] = clock64(); o[8] = (long long)gt1; o[9] = warpid; o[10] = tk_iters; o[11] = tk_frames;
  }
#endif
}

static bool mma_model(int model) { return model == EUCMT || model == KB4 || model == OPENCV5 || model == FTHETA; }

bool lin_mma_available(int model, int one_focal) {
  (void)one_focal;
  return mma_model(model);
}

int lin_mma_ctas(int n_sms, int n_frames) { return std::max(1, std::min(n_sms * kMmaCtasPerSm, (n_frames + kLinWarps - 1) / kLinWarps)); }

template <int MODEL, bool OF, bool BATCH, bool F32>
static cudaError_t launch_mma_t(const LinParams& prm, int n_ctas, cudaStream_t s) {
  auto kern = k_linearize_mma<MODEL, OF, BATCH, F32>;
  const size_t smem = (size_t)kLinWarps * mma_warp_smem_doubles() * sizeof(double);
  static bool configured = false;
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(n_ctas); cfg.blockDim = dim3(kLinThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = prm.ctl ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, prm);
}

template <int MODEL>
static cudaError_t launch_mma_m(int one_focal, bool batch, const LinParams& prm, int n_ctas, cudaStream_t s) {
  if (one_focal) {
    if (batch) return launch_mma_t<MODEL, true, true, false>(prm, n_ctas, s);
    return prm.pb.f32 ? launch_mma_t<MODEL, true, false, true>(prm, n_ctas, s) : launch_mma_t<MODEL, true, false, false>(prm, n_ctas, s);
  }
  if (batch) return launch_mma_t<MODEL, false, true, false>(prm, n_ctas, s);
  return prm.pb.f32 ? launch_mma_t<MODEL, false, false, true>(prm, n_ctas, s) : launch_mma_t<MODEL, false, false, false>(prm, n_ctas, s);
}

cudaError_t launch_linearize_mma(int model, int one_focal, bool batch, const LinParams& prm, int n_ctas, cudaStream_t s) {
  switch (model) {
    case EUCMT: return launch_mma_m<EUCMT>(one_focal, batch, prm, n_ctas, s);
    case KB4: return launch_mma_m<KB4>(one_focal, batch, prm, n_ctas, s);
    case OPENCV5: return launch_mma_m<OPENCV5>(one_focal, batch, prm, n_ctas, s);
    case FTHETA: return launch_mma_m<FTHETA>(one_focal, batch, prm, n_ctas, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace ccrs

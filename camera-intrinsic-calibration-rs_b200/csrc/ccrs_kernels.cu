// ccrs_kernels.cu — hand-written sm_100a FP64 kernels of the linearisation path.
//
//  K1 k_eval_rj      parity hook: per-observation residual (2) + Jacobian (2 x (d+6)), materialised (HBM-bound)
//  K2 k_linearize    fused residual + analytic Jacobian + Huber corrector + per-frame Gram blocks, no per-observation
//                    output (FP64-pipe-bound); COST_ONLY variant = K5 (residual-only robust cost)
//  K3 k_schur        per-frame damping, 6x6 Cholesky, elimination onto the intrinsic system, fixed-order reduction
//  K4 k_backsub      pose back-substitution, trial poses, LM model-decrease
// They replace, per iteration, tiny-solver's Problem::compute_residual_and_jacobian (num-dual autodiff of
// ReprojectionFactor::residual_func, reference src/optimization/factors.rs:152-173), the sparse J^T J product and
// the sparse LLT (call sites src/util.rs:455,463,670). The kernels in this file run on the FP64 CUDA cores: the per-frame
// blocks are 6x6 / 6xd and the packed sparse Gram update of the small models needs a third of the FMAs of a tiled J^T J
// on the (same) FP64 engine; the models with d >= 8 use the tensor-core variant in ccrs_linmma.cu, the single-problem
// reduction + controller rule is k_schur2 in ccrs_loop.cu. Determinism: no floating-point atomics anywhere; every sum has
// a fixed order.
#include "ccrs_lincommon.cuh"

namespace ccrs {

// ------------------------------------------------------------------------------------------------
// K2 / K5. Warp-autonomous: warp w of the grid owns FPW = 32/G consecutive frames; G lanes cooperate on a frame,
// lane j taking observations j, j+G, ... (so the G lanes read G consecutive doubles of each SoA array: coalesced,
// no reliance on L1). Every lane keeps the whole packed Gram block of its slice in registers (NACC FP64
// accumulators), rotates it from the local rotation basis to the rvec basis once (J_l), then the slices of a
// frame are summed in lane order through the warp's own shared memory and the frame block is written SoA to HBM.
// No CTA-wide barrier anywhere: prologue (fused K4 + pose exponential), main loop and reduction of one warp
// overlap with whatever phase the other resident warps are in.
// ------------------------------------------------------------------------------------------------
constexpr int kRedStride = 33;    // row stride of the reduction staging buffer (odd: no bank conflicts)
constexpr int kA2bDoubles = 112;  // per-warp copy of the accumulator -> block-entry table (<= 224 int32)
CCRS_HD constexpr int lin_warp_smem_doubles(int FPW, bool batch, bool cost_only) {
  return FPW * kFrameConst + (batch ? FPW * kMaxFull : kMaxFull + 1) + 2 * FPW + (cost_only ? 32 : kRedChunk * kRedStride) +
         kObsStages * 5 * 32 + kA2bDoubles;
}

// basis change phi -> rvec on one slice's packed block: H <- T^T H T, T = blkdiag(I_D, J_l, I_3, 1)
template <class C>
CCRS_D void basis_change(double (&acc)[C::NACC], const double* __restrict__ Jl) {
  static_for<0, C::NA>([&](auto Cc) {
    constexpr int c = decltype(Cc)::value;
    if constexpr (c < C::D || c >= C::D + 3) {
      constexpr int k0 = c < C::D ? C::kidx(c, C::D + 0) : C::kidx(C::D + 0, c);
      constexpr int k1 = c < C::D ? C::kidx(c, C::D + 1) : C::kidx(C::D + 1, c);
      constexpr int k2 = c < C::D ? C::kidx(c, C::D + 2) : C::kidx(C::D + 2, c);
      const double h0 = acc[k0], h1 = acc[k1], h2 = acc[k2];
      acc[k0] = fma(Jl[0], h0, fma(Jl[3], h1, Jl[6] * h2));
      acc[k1] = fma(Jl[1], h0, fma(Jl[4], h1, Jl[7] * h2));
      acc[k2] = fma(Jl[2], h0, fma(Jl[5], h1, Jl[8] * h2));
    }
  });
  {
    constexpr int p = C::D;
    const double h00 = acc[C::kidx(p, p)], h01 = acc[C::kidx(p, p + 1)], h02 = acc[C::kidx(p, p + 2)];
    const double h11 = acc[C::kidx(p + 1, p + 1)], h12 = acc[C::kidx(p + 1, p + 2)], h22 = acc[C::kidx(p + 2, p + 2)];
    double tmp[3][3];  // H_pp * Jl
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      tmp[0][c] = fma(h00, Jl[c], fma(h01, Jl[3 + c], h02 * Jl[6 + c]));
      tmp[1][c] = fma(h01, Jl[c], fma(h11, Jl[3 + c], h12 * Jl[6 + c]));
      tmp[2][c] = fma(h02, Jl[c], fma(h12, Jl[3 + c], h22 * Jl[6 + c]));
    }
    auto g = [&](int a, int b) { return fma(Jl[a], tmp[0][b], fma(Jl[3 + a], tmp[1][b], Jl[6 + a] * tmp[2][b])); };
    acc[C::kidx(p, p)] = g(0, 0); acc[C::kidx(p, p + 1)] = g(0, 1); acc[C::kidx(p, p + 2)] = g(0, 2);
    acc[C::kidx(p + 1, p + 1)] = g(1, 1); acc[C::kidx(p + 1, p + 2)] = g(1, 2); acc[C::kidx(p + 2, p + 2)] = g(2, 2);
  }
}

// rank-2 update of the packed Gram block with the two weighted rows of one observation (structural zeros skipped)
template <class C>
CCRS_D void gram_accumulate(double (&acc)[C::NACC], const double* __restrict__ au, const double* __restrict__ av) {
  static_for<0, C::NA>([&](auto I) {
    static_for<decltype(I)::value, C::NA>([&](auto J) {
      constexpr int i = decltype(I)::value, j = decltype(J)::value;
      constexpr int kk = C::kidx(i, j);
      if constexpr (kk >= 0) {
        if constexpr (C::hasu(i, j)) acc[kk] = fma(au[i], au[j], acc[kk]);
        if constexpr (C::hasv(i, j)) acc[kk] = fma(av[i], av[j], acc[kk]);
      }
    });
  });
}

// rank-1 update of one compact row's Gram block (pair variant: every entry is dense)
template <class R>
CCRS_D void gram_row(double (&acc)[R::NACC], const double* __restrict__ r) {
  static_for<0, R::NA>([&](auto A) {
    static_for<decltype(A)::value, R::NA>([&](auto B) {
      constexpr int a = decltype(A)::value, b = decltype(B)::value;
      acc[R::kidx(a, b)] = fma(r[a], r[b], acc[R::kidx(a, b)]);
    });
  });
}

// Pair variant of the frame reduction: compact entry e of the even lanes is a u-row product, of the odd lanes the
// matching v-row product. Where both land on the same entry of the frame block (columns common to both rows) the G
// slices are summed in lane order; otherwise the even and the odd lanes are summed separately into their two entries.
// s_a2b = [dense index of the u-row entry | dense index of the v-row entry], NACC each.
template <class R>
CCRS_D void slices_reduce_store_pair(double (&acc)[R::NACC], bool active, int lane, int fl, int sl, int G,
                                     double* __restrict__ s_red, const int* __restrict__ s_a2b, double* __restrict__ out, size_t Fs) {
  constexpr int NCH = (R::NACC + kRedChunk - 1) / kRedChunk;
  const double* const srow = s_red + fl * G;
  static_for<0, NCH>([&](auto CH) {
    constexpr int ch = decltype(CH)::value;
    constexpr int cnt = (R::NACC - ch * kRedChunk) < kRedChunk ? (R::NACC - ch * kRedChunk) : kRedChunk;
    if (ch > 0) __syncwarp();
    if (active) {
      static_for<0, cnt>([&](auto E) {
        constexpr int e = decltype(E)::value;
        s_red[e * kRedStride + lane] = acc[ch * kRedChunk + e];
      });
    }
    __syncwarp();
    if (active) {
      for (int e = sl; e < cnt; e += G) {
        const double* src = srow + e * kRedStride;
        const int bu = s_a2b[ch * kRedChunk + e], bv = s_a2b[R::NACC + ch * kRedChunk + e];
        if (bu == bv) {
          double t = src[0];
          for (int j = 1; j < G; ++j) t += src[j];
          out[(size_t)bu * Fs] = t;
        } else {
          double tu = src[0], tv = src[1];
          for (int j = 2; j < G; j += 2) { tu += src[j]; tv += src[j + 1]; }
          out[(size_t)bu * Fs] = tu;
          out[(size_t)bv * Fs] = tv;
        }
      }
    }
  });
}

// Sum the G slices of each frame of a warp in slice order (fixed order -> deterministic) and store the frame blocks SoA.
// `out` = block buffer + this lane's frame.
template <class C>
CCRS_D void slices_reduce_store(double (&acc)[C::NACC], bool active, int lane, int fl, int sl, int G,
                                double* __restrict__ s_red, const int* __restrict__ s_a2b, double* __restrict__ out, size_t Fs) {
  constexpr int NCH = (C::NACC + kRedChunk - 1) / kRedChunk;
  // Staged through the warp's shared memory in chunks of kRedChunk entries, rows padded to kRedStride doubles
  // (bank-conflict-free both ways). Lane (fl, sl) then sums, for its own frame, the entries e = sl, sl+G, ... over
  // the frame's G slices in slice order and stores them: every lane of the frame works, the six lanes that hold
  // the same entry of consecutive frames store consecutive doubles.
  const double* const srow = s_red + fl * G;
  static_for<0, NCH>([&](auto CH) {
    constexpr int ch = decltype(CH)::value;
    constexpr int cnt = (C::NACC - ch * kRedChunk) < kRedChunk ? (C::NACC - ch * kRedChunk) : kRedChunk;
    if (ch > 0) __syncwarp();
    if (active) {
      static_for<0, cnt>([&](auto E) {
        constexpr int e = decltype(E)::value;
        s_red[e * kRedStride + lane] = acc[ch * kRedChunk + e];
      });
    }
    __syncwarp();
    // the slice count is one of the ten values choose_slicing() can pick: fully unrolled sums, no inner branches
    auto store_rounds = [&](auto GG) {
      constexpr int g = decltype(GG)::value;
      if (active) {
#pragma unroll 2
        for (int e = sl; e < cnt; e += (g > 0 ? g : G)) {
          const double* src = srow + e * kRedStride;
          double s = src[0];
          if constexpr (g > 0) {
#pragma unroll
            for (int j = 1; j < g; ++j) s += src[j];
          } else {
            for (int j = 1; j < G; ++j) s += src[j];
          }
          out[(size_t)s_a2b[ch * kRedChunk + e] * Fs] = s;
        }
      }
    };
    switch (G) {
      case 1: store_rounds(std::integral_constant<int, 1>{}); break;
      case 2: store_rounds(std::integral_constant<int, 2>{}); break;
      case 3: store_rounds(std::integral_constant<int, 3>{}); break;
      case 4: store_rounds(std::integral_constant<int, 4>{}); break;
      case 5: store_rounds(std::integral_constant<int, 5>{}); break;
      case 6: store_rounds(std::integral_constant<int, 6>{}); break;
      case 8: store_rounds(std::integral_constant<int, 8>{}); break;
      case 10: store_rounds(std::integral_constant<int, 10>{}); break;
      case 16: store_rounds(std::integral_constant<int, 16>{}); break;
      default: store_rounds(std::integral_constant<int, 0>{}); break;
    }
  });
}

template <int MODEL, bool OF, bool BATCH, bool COST_ONLY, bool F32>
__global__ void __launch_bounds__(kLinThreads, kLinCtasPerSm) k_linearize(const __grid_constant__ LinParams prm) {
  using C = Cfg<MODEL, OF>;
  extern __shared__ double smem[];
  const ProblemDev& pb = prm.pb;
  // a K3 enqueued behind this grid as a programmatic dependent may start launching now (it waits for our completion)
  asm volatile("griddepcontrol.launch_dependents;");
  // per-warp {model decrease, cost} of the CTA, combined by warp 0 at the end (named barrier 1: the only CTA-level
  // synchronisation of the kernel, and only warp 0 ever waits on it)
  __shared__ double2 s_wpart[kLinWarps];
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = prm.G, FPW = prm.FPW;
  const int gw = blockIdx.x * kLinWarps + wid;          // warp index in the grid
  const int f0 = gw * FPW;
  const int nf = max(0, min(FPW, pb.n_frames - f0));    // frames of this warp
  double* s_fc = smem + (size_t)wid * lin_warp_smem_doubles(FPW, BATCH, COST_ONLY);  // [FPW][kFrameConst]
  double* s_intr = s_fc + FPW * kFrameConst;             // batch: [FPW][kMaxFull] per frame; single problem: [kMaxFull + 1] per warp
  double* s_stat = s_intr + (BATCH ? FPW * kMaxFull : kMaxFull + 1);  // [2][FPW] per-frame md, cost
  double* s_red = s_stat + 2 * FPW;                      // [kRedChunk][kRedStride]
  double* s_obs = s_red + (COST_ONLY ? 32 : kRedChunk * kRedStride);  // [kObsStages][5][32] cp.async ring of x,y,z,u,v
  int* s_a2b = reinterpret_cast<int*>(s_obs + kObsStages * 5 * 32);  // [NACC] accumulator -> packed block entry
  constexpr bool PAIR = !COST_ONLY && lin_pair_v<MODEL, OF>;
  using R = RowCfg<C>;
  constexpr int NACC_L = COST_ONLY ? 1 : (PAIR ? R::NACC : C::NACC);   // accumulators per lane

#ifdef CCRS_K2_TIMING
  long long tck[8];
  unsigned long long gt0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));
  tck[0] = clock64();
#define CCRS_TCK(i) tck[i] = clock64()
#else
#define CCRS_TCK(i)
#endif
  const int fl = lane / G;            // frame of this lane within the warp
  const int sl = lane - fl * G;       // slice of the frame
  const bool active = fl < nf;
  const int f = f0 + (active ? fl : 0);

  // Observations are prefetched kObsStages-1 iterations ahead with 8-byte cp.async into a per-thread shared-memory
  // ring (each thread reads back only what it fetched: no barrier), so neither DRAM nor L2 latency sits at the
  // top of an iteration. The first stages are issued before the pose prologue so their latency overlaps it.
  // Every global load of the prologue is issued before anything waits on one of them: frame offsets, then the pose and
  // (fused K4) the frame's elimination record — one memory round trip instead of three at the start of the kernel,
  // when no other warp of the SM has work to hide it.
  constexpr int kA2bN = COST_ONLY ? 0 : (PAIR ? 2 * R::NACC : C::NACC);
  constexpr int kA2bPerLane = (kA2bN + 31) / 32;
  int a2b_reg[kA2bPerLane > 0 ? kA2bPerLane : 1];
#pragma unroll
  for (int i = 0; i < kA2bPerLane; ++i) a2b_reg[i] = (lane + 32 * i < kA2bN) ? __ldg(prm.acc_to_blk + lane + 32 * i) : 0;
  int fo_beg = 0, fo_end = 0, prob = 0, cur = 0;
  bool moves = false;
  double rt[6], X[6][C::D], cg[6], gp[6], dd[6], sp[6];
  double* pose_dst = nullptr;
  if (active) {
    fo_end = __ldg(pb.frame_offsets + f + 1);
    fo_beg = __ldg(pb.frame_offsets + f);
    prob = BATCH ? pb.frame_problem[f] : 0;
  }
  int end = 0, beg = 0;
  double* ring = s_obs + lane;
  auto fetch = [&](int kk, int stage) {
    if (kk < end) {
      double* dst = ring + stage * (5 * 32);
      if constexpr (F32) {
        const float *fx = (const float*)pb.x, *fy = (const float*)pb.y, *fz = (const float*)pb.z, *fu = (const float*)pb.u, *fv = (const float*)pb.v;
        cp_async4(dst, fx + kk); cp_async4(dst + 32, fy + kk); cp_async4(dst + 64, fz + kk);
        cp_async4(dst + 96, fu + kk); cp_async4(dst + 128, fv + kk);
      } else {
        cp_async8(dst, pb.x + kk); cp_async8(dst + 32, pb.y + kk); cp_async8(dst + 64, pb.z + kk);
        cp_async8(dst + 96, pb.u + kk); cp_async8(dst + 128, pb.v + kk);
      }
    }
    cp_async_commit();
  };
  auto prefetch_first_stages = [&]() {
    end = active ? fo_end : 0;
    beg = active ? fo_beg + sl : 0;
#pragma unroll
    for (int i = 0; i < kObsStages - 1; ++i) fetch(beg + i * G, i);
  };
  // which point is linearised, and how the pending step is applied, come from the launch (host-driven) or from the
  // device control block (device-driven loop)
  int which = prm.which, backsub = prm.backsub, phase = -1;
  double u_bs = prm.u;
  double ya_bs[C::D];
  double lin_intr[C::D];
  if constexpr (BATCH) {
    // batch handles reach their pose through two more dependent loads (frame -> problem -> state selector): there the
    // observation prefetch goes out as soon as the frame offsets are back. A problem that has stopped contributes no
    // observations (its frames' poses stay put, its blocks are not needed any more).
    if (active && prm.active && !prm.active[prob]) fo_end = fo_beg;
    prefetch_first_stages();
    if (active) cur = cur_of(pb, prob);
  } else if (prm.ctl) {
    // Device-driven loop: this grid was launched as a programmatic dependent of the K3 in front of it. Everything
    // above and the observation prefetch touch only immutable data; the control block, the poses and the elimination
    // record are read after that K3 has completed — all of them issued together (both pose buffers: which one is
    // current is only known once the control block is back), one memory round trip.
    prefetch_first_stages();
    asm volatile("griddepcontrol.wait;" ::: "memory");
    const LoopCtl* ctl = prm.ctl;
    if (gw == 0 && lane == 0) prm.ctl->t_k2_wake = stamp_ns();
    const CtlHot hot = load_ctl_hot<C::D>(ctl, lane, lin_intr, ya_bs);   // one request per warp, not 4 + 2 d
    const int ph = hot.phase, lm = hot.lm, cu = hot.cur;
    u_bs = hot.u_used;
    double r0[6], r1[6];
    if (active) {
      const double *p0 = pb.poses[0] + 6 * (size_t)f, *p1 = pb.poses[1] + 6 * (size_t)f;
      const double* el = prm.elim + f;
      const size_t Fs = pb.Fs;
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        r0[i] = __ldcg(p0 + i); r1[i] = __ldcg(p1 + i);
#pragma unroll
        for (int a = 0; a < C::D; ++a) X[i][a] = __ldcg(el + (size_t)(i * C::D + a) * Fs);
        cg[i] = __ldcg(el + (size_t)(6 * C::D + i) * Fs);
        gp[i] = __ldcg(el + (size_t)(6 * C::D + 6 + i) * Fs);
        dd[i] = __ldcg(el + (size_t)(6 * C::D + 12 + i) * Fs);
        sp[i] = prm.pose_scale ? __ldcg(prm.pose_scale + (size_t)i * Fs + f) : 1.0;
      }
    }
    phase = ph;
    if (phase != PH_LIN0 && phase != PH_TRIAL) {   // not this slot's turn (re-reduction pending, or the loop is done)
      cp_async_wait<0>();
      return;
    }
    if (gw == 0 && lane == 0) prm.ctl->t_k2_begin = stamp_ns();
    cur = cu;
    backsub = phase == PH_LIN0 ? 0 : (lm ? 1 : 2);
    which = (phase == PH_TRIAL && lm) ? 1 : 0;
    if (active) {
#pragma unroll
      for (int i = 0; i < 6; ++i) rt[i] = cur ? r1[i] : r0[i];
      moves = backsub != 0;
      pose_dst = pb.poses[backsub == 2 ? cur : (cur ^ 1)] + 6 * (size_t)f;
    }
  } else {
    cur = pb.cur_val;
#pragma unroll
    for (int a = 0; a < C::D; ++a) ya_bs[a] = prm.y_a[a];
  }
  if (active && (BATCH || !prm.ctl)) {
    if (backsub) {
      const double* src = pb.poses[cur] + 6 * (size_t)f;
      pose_dst = pb.poses[backsub == 2 ? cur : (cur ^ 1)] + 6 * (size_t)f;
      moves = !(BATCH && prm.active && !prm.active[prob]);
#pragma unroll
      for (int i = 0; i < 6; ++i) rt[i] = src[i];
      if (moves) {
        const double* el = prm.elim + f;
        const size_t Fs = pb.Fs;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
#pragma unroll
          for (int a = 0; a < C::D; ++a) X[i][a] = __ldcg(el + (size_t)(i * C::D + a) * Fs);
          cg[i] = __ldcg(el + (size_t)(6 * C::D + i) * Fs);
          gp[i] = __ldcg(el + (size_t)(6 * C::D + 6 + i) * Fs);
          dd[i] = __ldcg(el + (size_t)(6 * C::D + 12 + i) * Fs);
          sp[i] = prm.pose_scale ? __ldcg(prm.pose_scale + (size_t)i * Fs + f) : 1.0;
        }
      }
    } else {
      const double* src = pb.poses[cur ^ which] + 6 * (size_t)f;
#pragma unroll
      for (int i = 0; i < 6; ++i) rt[i] = src[i];
    }
  }
  if constexpr (!BATCH) if (!prm.ctl) prefetch_first_stages();

  // ---- prologue: every lane of a frame evaluates the frame's pose redundantly (same addresses: broadcast loads);
  //      slice 0 stores. Fused K4: y_p = cg - X y_a ; pose += D_p y_p ; model decrease y_p^T g'_p + u sum dd_i y_p,i^2
  double md = 0.0;   // model decrease of this lane's frame (same value in all lanes of the frame)
  if (active) {
    if (backsub) {
      if (moves) {
        const double* ya = BATCH ? prm.ya_dev + (size_t)prob * C::D : ya_bs;
        const double u = BATCH ? (prm.u_dev ? prm.u_dev[prob] : 0.0) : u_bs;
#pragma unroll
        for (int i = 0; i < 6; ++i) {
          double yp = cg[i];
#pragma unroll
          for (int a = 0; a < C::D; ++a) yp -= X[i][a] * ya[a];
          rt[i] += sp[i] * yp;
          md += yp * gp[i] + u * dd[i] * yp * yp;
        }
      }
      if (sl == 0) {
#pragma unroll
        for (int i = 0; i < 6; ++i) pose_dst[i] = rt[i];
        if (BATCH && prm.frame_md) prm.frame_md[f] = md;
      }
    }
    FramePose fp;
    pose_from_rvec_tvec(rt, fp);
    if (sl == 0) {
      double* o = s_fc + fl * kFrameConst;
#pragma unroll
      for (int i = 0; i < 9; ++i) o[i] = fp.R[i];
#pragma unroll
      for (int i = 0; i < 3; ++i) o[9 + i] = fp.t[i];
#pragma unroll
      for (int i = 0; i < 9; ++i) o[12 + i] = fp.Jl[i];
      if constexpr (BATCH) {
        const double* a = prm.intr_dev + (size_t)prob * C::D;
        double* si = s_intr + fl * kMaxFull;
        if constexpr (OF) { si[0] = a[0]; si[1] = a[0]; for (int i = 1; i < C::D; ++i) si[i + 1] = a[i]; }
        else { for (int i = 0; i < C::D; ++i) si[i] = a[i]; }
      }
    }
  }
  if constexpr (!BATCH) {
    // the warp's copy of the full intrinsic vector (fy := fx inserted for one-focal problems, factors.rs:156-158)
    if (lane == 0) {
      if (prm.ctl) {
        if constexpr (OF) { s_intr[0] = lin_intr[0]; s_intr[1] = lin_intr[0]; for (int i = 1; i < C::D; ++i) s_intr[i + 1] = lin_intr[i]; }
        else { for (int i = 0; i < C::D; ++i) s_intr[i] = lin_intr[i]; }
      } else {
#pragma unroll
        for (int i = 0; i < C::DFULL; ++i) s_intr[i] = prm.intr[i];
      }
    }
  }
#pragma unroll
  for (int i = 0; i < kA2bPerLane; ++i) if (lane + 32 * i < kA2bN) s_a2b[lane + 32 * i] = a2b_reg[i];   // loaded with the prologue's other loads
  __syncwarp();
  CCRS_TCK(1);

  const double* ip = BATCH ? (s_intr + (active ? fl : 0) * kMaxFull) : s_intr;
  const double* fc = s_fc + (active ? fl : 0) * kFrameConst;

  double acc[NACC_L];
#pragma unroll
  for (int i = 0; i < NACC_L; ++i) acc[i] = 0.0;

  if constexpr (PAIR) {
    // every lane of a pair runs the pair's trip count (the even lane's: it owns the extra observation of a ragged
    // slice); the row exchange below needs both lanes
    const int n_mine = (active && beg < end) ? (end - beg + G - 1) / G : 0;
    const int n_pair = max(n_mine, __shfl_xor_sync(0xffffffffu, n_mine, 1));
    const unsigned pair_mask = 3u << (lane & 30);
    const bool odd = (lane & 1) != 0;
    int k = beg, it = 0;
    // one pair iteration; CHECKED = the lane's slice may be exhausted (only the last iteration of a ragged pair)
    auto pair_step = [&](auto CHECKED) {
      constexpr bool checked = decltype(CHECKED)::value;
      fetch(k + (kObsStages - 1) * G, (it + kObsStages - 1) % kObsStages);
      cp_async_wait<kObsStages - 1>();
      const bool valid = !checked || it < n_mine;
      const double* src = ring + (it % kObsStages) * (5 * 32);
      auto ld = [&](int a) -> double {
        if constexpr (F32) return (double)*reinterpret_cast<const float*>(src + a * 32);
        else return src[a * 32];
      };
      double au[C::NA], av[C::NA];
      obs_rows<MODEL, OF, true>(ip, fc, ld(0), ld(1), ld(2), ld(3), ld(4), pb.huber_delta, au, av);
      // mine: the row this lane accumulates (even: u, odd: v); give: the row its partner accumulates. An exhausted
      // slice contributes zeros (selected, not multiplied: stale ring data may hold anything).
      double mine[R::NA], give[R::NA];
      static_for<0, R::NA>([&](auto Cc) {
        constexpr int c = decltype(Cc)::value;
        const double a = au[C::ucol(c)], b = av[C::vcol(c)];
        if constexpr (checked) { mine[c] = valid ? (odd ? b : a) : 0.0; give[c] = valid ? (odd ? a : b) : 0.0; }
        else { mine[c] = odd ? b : a; give[c] = odd ? a : b; }
      });
#pragma unroll
      for (int c = 0; c < R::NA; ++c) give[c] = __shfl_xor_sync(pair_mask, give[c], 1);
      gram_row<R>(acc, mine);
      gram_row<R>(acc, give);
      ++it; k += G;
    };
    const int n_both = min(n_mine, __shfl_xor_sync(0xffffffffu, n_mine, 1));   // iterations in which both lanes hold an observation
    while (it < n_both) pair_step(std::false_type{});
    while (it < n_pair) pair_step(std::true_type{});
    CCRS_TCK(2);
    if (active) basis_change<R>(acc, fc + 12);
  } else

  if (active) {
    // rows of [J | r] for the observation in ring slot `stage` (f32 -> f64 widening as factors.rs:141-143)
    auto rows_of = [&](int stage, double* __restrict__ au, double* __restrict__ av) -> double {
      const double* src = ring + stage * (5 * 32);
      auto ld = [&](int a) -> double {
        if constexpr (F32) return (double)*reinterpret_cast<const float*>(src + a * 32);
        else return src[a * 32];
      };
      const double px = ld(0), py = ld(1), pz = ld(2), ou = ld(3), ov = ld(4);
      return obs_rows<MODEL, OF, !COST_ONLY>(ip, fc, px, py, pz, ou, ov, pb.huber_delta, au, av);
    };
    if constexpr (COST_ONLY) {
      int it = 0;
      for (int k = beg; k < end; k += G, ++it) {
        fetch(k + (kObsStages - 1) * G, (it + kObsStages - 1) % kObsStages);
        cp_async_wait<kObsStages - 1>();
        acc[0] += rows_of(it % kObsStages, nullptr, nullptr);
      }
    } else {
      // One observation per iteration: model chain, then the NACC independent Gram DFMAs. (A hand software-pipelined
      // version that issued observation i+1's chain among observation i's DFMAs was 2-5 % slower once the chain had been
      // shortened: its second row buffer cost 44 registers that the accumulators need; the other resident warp of the
      // sub-partition covers the chain's latency instead.)
      int it = 0;
      for (int k = beg; k < end; k += G, ++it) {
        double au[C::NA], av[C::NA];
        fetch(k + (kObsStages - 1) * G, (it + kObsStages - 1) % kObsStages);
        cp_async_wait<kObsStages - 1>();
        rows_of(it % kObsStages, au, av);
        gram_accumulate<C>(acc, au, av);
      }
    }
    CCRS_TCK(2);
    if constexpr (!COST_ONLY) {
      basis_change<C>(acc, fc + 12);
    }
  }

  CCRS_TCK(3);
  // ---- frame cost = sum over the frame's slices of the (r,r) entry, in slice order (the same order as the block
  //      reduction below, so it equals the stored entry bit for bit); shuffles: the G lanes of a frame are adjacent
  double fcost;
  {
    const double v = active ? acc[NACC_L - 1] : 0.0;
    fcost = v;
    for (int j = 1; j < G; ++j) fcost += __shfl_down_sync(0xffffffffu, v, j);   // valid in slice 0 of each frame
  }
  // ---- fused statistics (single problem): {model decrease, cost} summed over frames in a fixed order: per-warp
  //      partials in frame order, published with a release-atomic ticket BEFORE the block reduction so the atomic's
  //      round trip overlaps it; the last warp to take a ticket sums the warp partials (fixed tree) at the very end.
  unsigned ticket_old = 0xffffffffu;   // stays so unless this warp is the last of its CTA
  if constexpr (!BATCH) {
    // all shuffles first (independent), then the two fixed-order sums: a warp owns at most 32 frames, usually 4-8
    double wmd = 0.0, wcost = 0.0;
    if (FPW <= 8) {
      double tm[8], tc[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int srcl = min(i * G, 31);
        tm[i] = __shfl_sync(0xffffffffu, md, srcl);
        tc[i] = __shfl_sync(0xffffffffu, fcost, srcl);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) if (i < nf) { wmd += tm[i]; wcost += tc[i]; }
    } else {
      for (int i = 0; i < nf; ++i) {
        wmd += __shfl_sync(0xffffffffu, md, i * G);
        wcost += __shfl_sync(0xffffffffu, fcost, i * G);
      }
    }
    // The CTA's four warps combine in warp order through shared memory, so the grid leaves one partial per CTA: the
    // final sum reads 292 slots in one round of loads instead of 1,167 in three. Warps 1-3 post their partial and arrive
    // at named barrier 1 without waiting; warp 0 waits on it after its own block reduction (below).
    if (lane == 0) s_wpart[wid] = make_double2(wmd, wcost);
    __syncwarp();
    if (wid != 0) asm volatile("bar.arrive 1, %0;" ::"n"(kLinThreads) : "memory");
  }

  CCRS_TCK(6);
  // ---- sum the G slices of each frame in slice order (fixed order -> deterministic) and store SoA ----
  if constexpr (COST_ONLY) {
    if (active && sl == 0) {
      const int prob = BATCH ? pb.frame_problem[f] : 0;
      pb.frame_cost[(BATCH ? cur_of(pb, prob) : cur) ^ which][f] = fcost;
    }
  } else {
    double* const out = pb.blocks[(BATCH ? (active ? cur_of(pb, pb.frame_problem[f]) : 0) : cur) ^ which] + f;
    if constexpr (PAIR) slices_reduce_store_pair<R>(acc, active, lane, fl, sl, G, s_red, s_a2b, out, (size_t)pb.Fs);
    else slices_reduce_store<C>(acc, active, lane, fl, sl, G, s_red, s_a2b, out, (size_t)pb.Fs);
  }

  CCRS_TCK(4);
  if constexpr (!BATCH) {
    if (wid == 0) {
      asm volatile("bar.sync 1, %0;" ::"n"(kLinThreads) : "memory");   // the other warps' partials are posted
      if (lane == 0) {
        double a = 0.0, b = 0.0;
#pragma unroll
        for (int w = 0; w < kLinWarps; ++w) { a += s_wpart[w].x; b += s_wpart[w].y; }
        // no fence: the two partial words validate themselves (armed with kArmBits by the previous launch's last warp /
        // at creation); the relaxed ticket only elects the warp that sums
        reinterpret_cast<double2*>(prm.cta_part)[blockIdx.x] = make_double2(a, b);
        ticket_old = atomicAdd(prm.ticket, 1u);
      }
      const unsigned last = __shfl_sync(0xffffffffu, (unsigned)(ticket_old == gridDim.x - 1), 0);
      if (last) stats_finalize(prm, gridDim.x, lane, phase);
    }
  }
#ifdef CCRS_K2_TIMING
  CCRS_TCK(5);
  if (prm.dbg && lane == 0) {
    long long* o = prm.dbg + (size_t)gw * 12;
    unsigned smid, warpid;
    unsigned long long gt1;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u32 %0, %%warpid;" : "=r"(warpid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
    o[0] = (long long)gt0; o[1] = smid;
    for (int i = 0; i < 6; ++i) o[2 + i] = tck[i];
    o[8] = (long long)gt1; o[9] = warpid; o[10] = tck[6]; o[11] = 0;
  }
#endif
}

// ------------------------------------------------------------------------------------------------
// K1 parity hook: one thread per observation, r and J materialised (HBM-bound: 248 B/obs for EUCM).
// ------------------------------------------------------------------------------------------------
template <int MODEL, bool OF>
__global__ void __launch_bounds__(128) k_eval_rj(ProblemDev pb, const double* __restrict__ intr_dev,
                                                 const double* __restrict__ poses, int apply_loss,
                                                 double* __restrict__ r, double* __restrict__ J, int64_t n_obs) {
  using C = Cfg<MODEL, OF>;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_obs) return;
  const int f = pb.obs_frame[k];
  double ip[kMaxFull];
  if constexpr (OF) { ip[0] = intr_dev[0]; ip[1] = intr_dev[0]; for (int i = 1; i < C::D; ++i) ip[i + 1] = intr_dev[i]; }
  else { for (int i = 0; i < C::D; ++i) ip[i] = intr_dev[i]; }
  FramePose fp;
  pose_from_rvec_tvec(poses + 6 * (size_t)f, fp);
  double fc[12];
  for (int i = 0; i < 9; ++i) fc[i] = fp.R[i];
  for (int i = 0; i < 3; ++i) fc[9 + i] = fp.t[i];
  double au[C::NA], av[C::NA];
  for (int i = 0; i < C::NA; ++i) { au[i] = 0.0; av[i] = 0.0; }
  obs_rows<MODEL, OF, true>(ip, fc, ld_obs(pb.x, k, pb.f32), ld_obs(pb.y, k, pb.f32), ld_obs(pb.z, k, pb.f32), ld_obs(pb.u, k, pb.f32),
                            ld_obs(pb.v, k, pb.f32), apply_loss ? pb.huber_delta : 0.0, au, av);
  r[2 * k] = au[C::N]; r[2 * k + 1] = av[C::N];
  if (J) {
    double* j0 = J + (size_t)(2 * k) * C::N;
    double* j1 = j0 + C::N;
    for (int i = 0; i < C::D; ++i) { j0[i] = C::nzu(i) ? au[i] : 0.0; j1[i] = C::nzv(i) ? av[i] : 0.0; }
    for (int c = 0; c < 3; ++c) {  // d/drvec = d/dphi * J_l
      j0[C::D + c] = au[C::D] * fp.Jl[c] + au[C::D + 1] * fp.Jl[3 + c] + au[C::D + 2] * fp.Jl[6 + c];
      j1[C::D + c] = av[C::D] * fp.Jl[c] + av[C::D + 1] * fp.Jl[3 + c] + av[C::D + 2] * fp.Jl[6 + c];
      j0[C::D + 3 + c] = au[C::D + 3 + c];
      j1[C::D + 3 + c] = av[C::D + 3 + c];
    }
  }
}

// K6 (validation): per-observation reprojection error without the robust loss,
// sqrt(dx^2 + dy^2) of project_one(T * p3d) - p2d  (src/util.rs:733-745). 48 B/obs: HBM-bound.
template <int MODEL, bool OF>
__global__ void __launch_bounds__(128) k_reproj_err(ProblemDev pb, const double* __restrict__ intr_dev,
                                                    const double* __restrict__ poses, double* __restrict__ err, int64_t n_obs) {
  using C = Cfg<MODEL, OF>;
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_obs) return;
  const int f = pb.obs_frame[k];
  double ip[kMaxFull];
  if constexpr (OF) { ip[0] = intr_dev[0]; ip[1] = intr_dev[0]; for (int i = 1; i < C::D; ++i) ip[i + 1] = intr_dev[i]; }
  else { for (int i = 0; i < C::D; ++i) ip[i] = intr_dev[i]; }
  FramePose fp;
  pose_from_rvec_tvec(poses + 6 * (size_t)f, fp);
  double fc[12];
  for (int i = 0; i < 9; ++i) fc[i] = fp.R[i];
  for (int i = 0; i < 3; ++i) fc[9 + i] = fp.t[i];
  const double s = obs_rows<MODEL, OF, false>(ip, fc, ld_obs(pb.x, k, pb.f32), ld_obs(pb.y, k, pb.f32), ld_obs(pb.z, k, pb.f32),
                                              ld_obs(pb.u, k, pb.f32), ld_obs(pb.v, k, pb.f32), 0.0, nullptr, nullptr);
  err[k] = sqrt(s);
}

// ------------------------------------------------------------------------------------------------
// K3. One thread per frame: scale + damp C_f, Cholesky (registers), Y = L^-1 B'^T, S_f = A'_f - Y^T Y,
// g_s = g'_a - Y^T L^-1 g'_p, X = L^-T Y and cg = C^-1 g'_p stored for the back-substitution.
// Single problem: the CTA's frame contributions are summed in thread order in shared memory and one
// partial per CTA is written (summed in CTA order by k_sum_partials). Batch: per-frame contributions are
// written and reduced per problem by k_segreduce.
// ------------------------------------------------------------------------------------------------
constexpr int kSchurThreads = 128;
constexpr int kSchurSplit = 3;      // lanes per value in the last CTA's sum over CTA partials

template <int D, bool BATCH>
__global__ void __launch_bounds__(kSchurThreads) k_schur(const __grid_constant__ SchurParams prm, double* __restrict__ partials) {
  constexpr int N = D + 6, NA = N + 1;
  constexpr int NS = D * (D + 1) / 2;
  constexpr int NRED = NS + 3 * D + 1;
  extern __shared__ double s_red[];  // [NRED][kSchurThreads+1]   (single problem only)
  const ProblemDev& pb = prm.pb;
  // launched as a programmatic dependent of the K2 in front of it: wait until that grid has completed and flushed
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const int f = blockIdx.x * kSchurThreads + threadIdx.x;
  bool valid = f < pb.n_frames;
  if (BATCH && valid && prm.active && !prm.active[pb.frame_problem[f]]) valid = false;   // its problem has stopped
  double red[NRED];
#pragma unroll
  for (int i = 0; i < NRED; ++i) red[i] = 0.0;
  int bad = 0;
#ifdef CCRS_K2_TIMING
  long long tk[6];
  tk[0] = clock64();
#define CCRS_TK3(i) tk[i] = clock64()
#else
#define CCRS_TK3(i)
#endif
  if (valid) {
    const int prob = BATCH ? pb.frame_problem[f] : 0;
    const double* blk = pb.blocks[cur_of(pb, prob) ^ prm.which] + f;
    const size_t Fs = pb.Fs;
    // The frame's whole packed block goes to shared memory in one burst of cp.async (each thread its own column: no
    // barrier): one memory round trip for the kernel instead of one per dependent stage of the elimination, which
    // with one warp per sub-partition nothing else would hide.
    constexpr int NB = NA * (NA + 1) / 2;
    double* const s_col = s_red + threadIdx.x;
#pragma unroll
    for (int e = 0; e < NB; ++e) cp_async8(s_col + e * kSchurThreads, blk + (size_t)e * Fs);
    cp_async_commit();
    cp_async_wait<0>();
    auto H = [&](int i, int j) { return s_col[tri_idx(NA, i, j) * kSchurThreads]; };
    const double u = prm.u_dev ? prm.u_dev[prob] : prm.u_val;
    double sa[D], sp[6];
#pragma unroll
    for (int a = 0; a < D; ++a) sa[a] = prm.intr_scale ? prm.intr_scale[prob * D + a] : 1.0;
#pragma unroll
    for (int i = 0; i < 6; ++i) sp[i] = prm.pose_scale ? prm.pose_scale[(size_t)i * Fs + f] : 1.0;
    // C' (lower, in place Cholesky), damping. no_pose: the frame's pose is not a variable (intrinsics-only problems)
    const bool np = prm.no_pose != 0;
    double L[6][6], gp[6], dd[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
#pragma unroll
      for (int j = 0; j <= i; ++j) L[i][j] = np ? (i == j ? 1.0 : 0.0) : sp[i] * H(D + j, D + i) * sp[j];
      gp[i] = np ? 0.0 : -sp[i] * H(D + i, N);
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
      L[j][j] = il;  // store the reciprocal of the pivot
#pragma unroll
      for (int i = j + 1; i < 6; ++i) {
        double tt = L[i][j];
#pragma unroll
        for (int k = 0; k < j; ++k) tt -= L[i][k] * L[j][k];
        L[i][j] = tt * il;
      }
    }
    CCRS_TK3(1);
    // Y[a] = L^-1 B'[a,:]^T
    double Yv[D][6];
#pragma unroll
    for (int a = 0; a < D; ++a) {
#pragma unroll
      for (int i = 0; i < 6; ++i) {
        double s = np ? 0.0 : sa[a] * H(a, D + i) * sp[i];
#pragma unroll
        for (int k = 0; k < i; ++k) s -= L[i][k] * Yv[a][k];
        Yv[a][i] = s * L[i][i];
      }
    }
    double yg[6];
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      double s = gp[i];
#pragma unroll
      for (int k = 0; k < i; ++k) s -= L[i][k] * yg[k];
      yg[i] = s * L[i][i];
    }
    // reduced-system contributions
    {
      int e = 0;
#pragma unroll
      for (int a = 0; a < D; ++a) {
#pragma unroll
        for (int b = a; b < D; ++b) {
          double s = sa[a] * H(a, b) * sa[b];
#pragma unroll
          for (int i = 0; i < 6; ++i) s -= Yv[a][i] * Yv[b][i];
          red[e++] = s;
        }
      }
#pragma unroll
      for (int a = 0; a < D; ++a) {
        const double ga = -sa[a] * H(a, N);
        double s = ga;
#pragma unroll
        for (int i = 0; i < 6; ++i) s -= Yv[a][i] * yg[i];
        red[NS + a] = s;
        red[NS + D + a] = ga;
        red[NS + 2 * D + a] = sa[a] * H(a, a) * sa[a];
      }
      red[NS + 3 * D] = H(N, N);
    }
    // X = L^-T Y, cg = L^-T yg  -> elim (SoA)
    double* el = prm.elim + f;
#pragma unroll
    for (int a = 0; a < D; ++a) {
#pragma unroll
      for (int i = 5; i >= 0; --i) {
        double s = Yv[a][i];
#pragma unroll
        for (int k = i + 1; k < 6; ++k) s -= L[k][i] * Yv[a][k];
        Yv[a][i] = s * L[i][i];
        el[(size_t)(i * D + a) * Fs] = Yv[a][i];
      }
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
      double s = yg[i];
#pragma unroll
      for (int k = i + 1; k < 6; ++k) s -= L[k][i] * yg[k];
      yg[i] = s * L[i][i];
    }
#pragma unroll
    for (int i = 0; i < 6; ++i) {
      el[(size_t)(6 * D + i) * Fs] = yg[i];
      el[(size_t)(6 * D + 6 + i) * Fs] = gp[i];
      el[(size_t)(6 * D + 12 + i) * Fs] = dd[i];
    }
    if (bad) {  // poison the reduced system so the host sees the Cholesky failure (tiny-solver returns None); the cost
                // (last entry) stays valid: the failure is the factorisation's, not a NaN error
#pragma unroll
      for (int i = 0; i < NRED - 1; ++i) red[i] = nan("");
    }
  }
  CCRS_TK3(2);
  if constexpr (BATCH) {
    if (valid) {
#pragma unroll
      for (int i = 0; i < NRED; ++i) prm.frame_red[(size_t)i * pb.Fs + f] = red[i];
    }
  } else {
    constexpr int LD = kSchurThreads + 1;
    __shared__ int s_last;
    __syncthreads();   // the block staging area becomes the reduction buffer
#pragma unroll
    for (int i = 0; i < NRED; ++i) s_red[i * LD + threadIdx.x] = red[i];
    __syncthreads();
    {  // SPLIT threads per value, each summing a contiguous segment of frames in order; fixed combine
      constexpr int SPLIT = kSchurThreads / NRED;
      constexpr int SEG = (kSchurThreads + SPLIT - 1) / SPLIT;
      __shared__ double s_seg[kSchurThreads];
      if (threadIdx.x < SPLIT * NRED) {
        const int v = threadIdx.x / SPLIT, h = threadIdx.x - SPLIT * v;
        const int j0 = h * SEG, j1 = min(kSchurThreads, j0 + SEG);
        const double* src = s_red + v * LD;
        double s = 0.0;
        for (int j = j0; j < j1; ++j) s += src[j];
        s_seg[threadIdx.x] = s;
      }
      __syncthreads();
      if (threadIdx.x < NRED) {
        double s = s_seg[SPLIT * threadIdx.x];
#pragma unroll
        for (int h = 1; h < SPLIT; ++h) s += s_seg[SPLIT * threadIdx.x + h];
        partials[(size_t)blockIdx.x * NRED + threadIdx.x] = s;
      }
    }
    CCRS_TK3(3);
    __syncthreads();   // partial stores of this CTA happen-before thread 0's release below (cumulativity)
    if (threadIdx.x == 0) {
      unsigned old;
      asm volatile("atom.add.release.gpu.global.u32 %0, [%1], 1;" : "=r"(old) : "l"(prm.ticket) : "memory");
      s_last = (old == gridDim.x - 1);
    }
    __syncthreads();
    CCRS_TK3(4);
    if (s_last) {  // last CTA: sum the CTA partials in CTA order; kSchurSplit interleaved lanes per value, fixed combine
      asm volatile("fence.acq_rel.gpu;" ::: "memory");
      static_assert(kSchurSplit * NRED <= NRED * (kSchurThreads + 1), "staging buffer too small");
      const int nb = gridDim.x;
      for (int idx = threadIdx.x; idx < kSchurSplit * NRED; idx += kSchurThreads) {
        const int v = idx / kSchurSplit, h = idx - v * kSchurSplit;
        double a = 0.0;
        for (int b0 = h; b0 < nb; b0 += 16 * kSchurSplit) {   // 16 loads in flight, then a fixed-order sum
          double t[16];
#pragma unroll
          for (int q = 0; q < 16; ++q) {
            const int b = b0 + q * kSchurSplit;
            t[q] = b < nb ? __ldcg(partials + (size_t)b * NRED + v) : 0.0;
          }
#pragma unroll
          for (int q = 0; q < 16; ++q) a += t[q];
        }
        s_red[idx] = a;
      }
      __syncthreads();
      if (threadIdx.x < NRED) {
        double tot = s_red[kSchurSplit * threadIdx.x];
#pragma unroll
        for (int h = 1; h < kSchurSplit; ++h) tot += s_red[kSchurSplit * threadIdx.x + h];
        if (prm.px.world > 1) tot = peer_exchange(prm.px, threadIdx.x, tot);
        prm.red_out[threadIdx.x] = tot;
        if (prm.host_red) prm.host_red[threadIdx.x] = tot;   // sentinel protocol: no fence
      }
      if (threadIdx.x == 0) *prm.ticket = 0u;
    }
#ifdef CCRS_K2_TIMING
    CCRS_TK3(5);
    if (prm.dbg && (threadIdx.x & 31) == 0 && valid) {
      long long* o = prm.dbg + (size_t)(blockIdx.x * (kSchurThreads / 32) + (threadIdx.x >> 5)) * 8;
      for (int i = 0; i < 6; ++i) o[i] = tk[i];
      o[6] = s_last; o[7] = 0;
    }
#endif
  }
}

// out[v] = sum_b partials[b][v], b ascending (fixed order); optionally published to mapped host memory
__global__ void k_sum_partials(const double* __restrict__ partials, int n_part, int NV, double* __restrict__ out,
                               volatile double* host_out, double seq) {
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v < NV) {
    double s = 0.0;
    for (int b = 0; b < n_part; ++b) s += partials[(size_t)b * NV + v];
    out[v] = s;
    if (host_out) host_out[v] = s;   // sentinel protocol: no fence
  }
}

// out[seg][v] = sum_{f in segment} in[v][f]; one CTA per segment, one WARP per value (eight values in flight, no CTA
// barrier): lane-strided partial sums in frame order, then a fixed butterfly — bit-reproducible.
constexpr int kSegThreads = 256;
__global__ void __launch_bounds__(kSegThreads) k_segreduce(const double* __restrict__ in, int NV, int Fs,
                                                           const int32_t* __restrict__ seg_off, double* __restrict__ out) {
  const int seg = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int b = seg_off[seg], e = seg_off[seg + 1];
  for (int v = warp; v < NV; v += kSegThreads / 32) {
    const double* src = in + (size_t)v * Fs;
    double s = 0.0;
    for (int f = b + lane; f < e; f += 32) s += src[f];
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) out[(size_t)seg * NV + v] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// K4. y_p = cg - X y_a ; dpose = D_p y_p ; trial = pose + dpose. LM model decrease (pose part):
// y^T(2g' - H'y) restricted to this frame = y_p^T g'_p + u * sum_i dd_i y_p,i^2   (uses H_reg y = g').
// ------------------------------------------------------------------------------------------------
template <int D, bool BATCH>
__global__ void __launch_bounds__(128) k_backsub(const __grid_constant__ BacksubParams prm) {
  const ProblemDev& pb = prm.pb;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= pb.n_frames) return;
  const int prob = BATCH ? pb.frame_problem[f] : 0;
  const size_t Fs = pb.Fs;
  const double* el = prm.elim + f;
  const double* ya = prm.y_a + (size_t)prob * D;
  const double u = prm.u_dev ? prm.u_dev[prob] : 0.0;
  const int cur = cur_of(pb, prob);
  const double* src = pb.poses[cur] + 6 * (size_t)f;
  double* dst = pb.poses[prm.in_place ? cur : (cur ^ 1)] + 6 * (size_t)f;
  double md = 0.0;
  if (prm.active && !prm.active[prob]) {  // converged problem of a batch: its poses stay put
    if (!prm.in_place) for (int i = 0; i < 6; ++i) dst[i] = src[i];
    if (prm.frame_md) prm.frame_md[f] = 0.0;
    return;
  }
#pragma unroll
  for (int i = 0; i < 6; ++i) {
    double yp = el[(size_t)(6 * D + i) * Fs];
#pragma unroll
    for (int a = 0; a < D; ++a) yp -= el[(size_t)(i * D + a) * Fs] * ya[a];
    const double sp = prm.pose_scale ? prm.pose_scale[(size_t)i * Fs + f] : 1.0;
    dst[i] = src[i] + sp * yp;
    md += yp * el[(size_t)(6 * D + 6 + i) * Fs] + u * el[(size_t)(6 * D + 12 + i) * Fs] * yp * yp;
  }
  if (prm.frame_md) prm.frame_md[f] = md;
}

// Jacobi scaling (tiny-solver LM, iteration 0): pose scales 1/(1+sqrt(C_ii)); per-frame A_aa diag for the host.
template <int D, bool BATCH>
__global__ void __launch_bounds__(128) k_compute_scale(ProblemDev pb, int which, double* __restrict__ pose_scale,
                                                       double* __restrict__ frame_colsq) {
  constexpr int NA = D + 7;
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= pb.n_frames) return;
  const int prob = BATCH ? pb.frame_problem[f] : 0;
  const double* blk = pb.blocks[cur_of(pb, prob) ^ which] + f;
  const size_t Fs = pb.Fs;
  for (int i = 0; i < 6; ++i) pose_scale[(size_t)i * Fs + f] = 1.0 / (1.0 + sqrt(blk[(size_t)tri_idx(NA, D + i, D + i) * Fs]));
  for (int a = 0; a < D; ++a) frame_colsq[(size_t)a * Fs + f] = blk[(size_t)tri_idx(NA, a, a) * Fs];
}

// per-problem {sum of frame_md, sum of per-frame cost}; one CTA per problem, fixed strided order + fixed tree
__global__ void __launch_bounds__(kSegThreads) k_trial_stats(ProblemDev pb, int rr_idx, int mode,
                                                             const double* __restrict__ frame_md,
                                                             double* __restrict__ stat_out) {
  __shared__ double sh0[kSegThreads], sh1[kSegThreads];
  const int q = blockIdx.x;
  const int b = pb.problem_frame_offsets[q], e = pb.problem_frame_offsets[q + 1];
  const int cur = cur_of(pb, q);
  const double* cost = nullptr;
  if (mode == 0) cost = pb.frame_cost[cur ^ 1];
  else if (mode == 4) cost = pb.frame_cost[cur];
  else if (mode == 1) cost = pb.blocks[cur ^ 1] + (size_t)rr_idx * pb.Fs;
  else if (mode == 2) cost = pb.blocks[cur] + (size_t)rr_idx * pb.Fs;
  double s0 = 0.0, s1 = 0.0;
  for (int f = b + threadIdx.x; f < e; f += kSegThreads) {
    if (frame_md) s0 += frame_md[f];
    if (cost) s1 += cost[f];
  }
  sh0[threadIdx.x] = s0; sh1[threadIdx.x] = s1;
  __syncthreads();
  for (int w = kSegThreads / 2; w > 0; w >>= 1) {
    if (threadIdx.x < w) { sh0[threadIdx.x] += sh0[threadIdx.x + w]; sh1[threadIdx.x] += sh1[threadIdx.x + w]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { stat_out[2 * q] = sh0[0]; stat_out[2 * q + 1] = sh1[0]; }
}

__global__ void k_flip_cur(int32_t* cur, const unsigned char* mask, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (!mask || mask[i])) cur[i] ^= 1;
}

// FP64 FMA throughput microbenchmark: 8 independent chains per thread.
__global__ void __launch_bounds__(256) k_fp64_peak(double* out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double m = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, m, c); a1 = fma(a1, m, c); a2 = fma(a2, m, c); a3 = fma(a3, m, c);
    a4 = fma(a4, m, c); a5 = fma(a5, m, c); a6 = fma(a6, m, c); a7 = fma(a7, m, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

// board-format problems (ccrs_problem_create_board_f32): p3d of observation k = board[corner_id[k]] (board.rs:46-95:
// corner id -> board point), expanded once into the f32 SoA arrays the linearisation kernels read
__global__ void __launch_bounds__(256) k_expand_board(const int32_t* __restrict__ id, const float* __restrict__ board, int n_board,
                                                      int64_t n, float* __restrict__ x, float* __restrict__ y, float* __restrict__ z,
                                                      volatile double* bad_flag) {
  extern __shared__ float s_board[];
  for (int i = threadIdx.x; i < 3 * n_board; i += blockDim.x) s_board[i] = board[i];
  __syncthreads();
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x) {
    int c = id[k];
    if ((unsigned)c >= (unsigned)n_board) { *bad_flag = 1.0; c = 0; }   // reported by ccrs_problem_create_board_f32
    x[k] = s_board[3 * c]; y[k] = s_board[3 * c + 1]; z[k] = s_board[3 * c + 2];
  }
}
cudaError_t launch_expand_board(const int32_t* id, const float* board, int n_board, int64_t n, float* x, float* y, float* z,
                                volatile double* bad_flag, cudaStream_t s) {
  const size_t smem = (size_t)3 * n_board * sizeof(float);
  if (smem > 48 * 1024) return cudaErrorInvalidValue;   // > 4096 board corners: not a calibration board
  const int nb = (int)std::min<int64_t>((n + 255) / 256, 148 * 8);
  k_expand_board<<<nb, 256, smem, s>>>(id, board, n_board, n, x, y, z, bad_flag);
  return cudaGetLastError();
}

__global__ void k_l2_flush(double* buf, size_t n) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) buf[i] = (double)i;
}

// ------------------------------------------------------------------------------------------------
// host-side dispatch
// ------------------------------------------------------------------------------------------------
template <class F>
static auto dispatch_model(int model, int of, F&& f) {
#define CCRS_CASE(M)                                                                         \
  case M:                                                                                    \
    return of ? f(std::integral_constant<int, M>{}, std::true_type{}) : f(std::integral_constant<int, M>{}, std::false_type{});
  switch (model) {
    CCRS_CASE(UCM) CCRS_CASE(EUCM) CCRS_CASE(EUCMT) CCRS_CASE(KB4) CCRS_CASE(OPENCV5) CCRS_CASE(FTHETA)
  }
#undef CCRS_CASE
  return f(std::integral_constant<int, EUCM>{}, std::false_type{});
}

int model_dims(int model, int one_focal, int* D, int* NA, int* NBLK, int* NACC) {
  if (model < 0 || model > 5) return -1;
  return dispatch_model(model, one_focal, [&](auto M, auto OF) {
    using C = Cfg<decltype(M)::value, decltype(OF)::value>;
    if (D) *D = C::D;
    if (NA) *NA = C::NA;
    if (NBLK) *NBLK = C::NBLK;
    // entries of the accumulator -> block-entry table: merged accumulators, or [u-row | v-row] for the pair variant
    if (NACC) *NACC = (C::NACC > kPairThreshold) ? 2 * RowCfg<C>::NACC : C::NACC;
    return 0;
  });
}

bool lin_uses_pairs(int model, int one_focal) {
  return dispatch_model(model, one_focal, [&](auto M, auto OF) { return Cfg<decltype(M)::value, decltype(OF)::value>::NACC > kPairThreshold; });
}

void fill_acc_to_blk(int model, int one_focal, int32_t* table) {
  dispatch_model(model, one_focal, [&](auto M, auto OF) {
    using C = Cfg<decltype(M)::value, decltype(OF)::value>;
    int k = 0;
    if (C::NACC > kPairThreshold) {
      using R = RowCfg<C>;
      for (int a = 0; a < R::NA; ++a)
        for (int b = a; b < R::NA; ++b) {
          table[k] = tri_idx(C::NA, C::ucol(a), C::ucol(b));
          table[R::NACC + k] = tri_idx(C::NA, C::vcol(a), C::vcol(b));
          ++k;
        }
      return 0;
    }
    for (int i = 0; i < C::NA; ++i)
      for (int j = i; j < C::NA; ++j)
        if (C::has(i, j)) table[k++] = tri_idx(C::NA, i, j);
    return 0;
  });
}

static size_t lin_smem_bytes(int FPW, bool batch, bool cost_only) {
  return (size_t)kLinWarps * lin_warp_smem_doubles(FPW, batch, cost_only) * sizeof(double);
}

template <int MODEL, bool OF, bool BATCH, bool COST, bool F32>
static cudaError_t launch_lin_t(const LinParams& prm, int n_ctas, cudaStream_t s) {
  auto kern = k_linearize<MODEL, OF, BATCH, COST, F32>;
  const size_t smem = lin_smem_bytes(prm.FPW, BATCH, COST);
  static bool configured = false;  // per instantiation
  if (!configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    if (e != cudaSuccess) return e;
    configured = true;
  }
  if (prm.ctl) {
    // device-driven loop: programmatic dependent of the K3 in front of it — the CTAs are scheduled while that K3
    // drains, run their observation prefetch and wait at griddepcontrol.wait for its control block
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(n_ctas); cfg.blockDim = dim3(kLinThreads); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, prm);
  }
  kern<<<n_ctas, kLinThreads, smem, s>>>(prm);
  return cudaGetLastError();
}

cudaError_t launch_linearize(int model, int one_focal, bool batch, bool cost_only, const LinParams& prm, int n_ctas,
                             cudaStream_t s) {
  return dispatch_model(model, one_focal, [&](auto M, auto OF) {
    constexpr int m = decltype(M)::value;
    constexpr bool of = decltype(OF)::value;
    if (batch) return cost_only ? launch_lin_t<m, of, true, true, false>(prm, n_ctas, s) : launch_lin_t<m, of, true, false, false>(prm, n_ctas, s);
    if (prm.pb.f32) return cost_only ? launch_lin_t<m, of, false, true, true>(prm, n_ctas, s) : launch_lin_t<m, of, false, false, true>(prm, n_ctas, s);
    return cost_only ? launch_lin_t<m, of, false, true, false>(prm, n_ctas, s) : launch_lin_t<m, of, false, false, false>(prm, n_ctas, s);
  });
}

cudaError_t launch_eval_rj(int model, int one_focal, const ProblemDev& pb, const double* intr_dev, const double* poses,
                           int apply_loss, double* r, double* J, int64_t n_obs, cudaStream_t s) {
  return dispatch_model(model, one_focal, [&](auto M, auto OF) {
    const int nb = (int)((n_obs + 127) / 128);
    k_eval_rj<decltype(M)::value, decltype(OF)::value><<<nb, 128, 0, s>>>(pb, intr_dev, poses, apply_loss, r, J, n_obs);
    return cudaGetLastError();
  });
}

cudaError_t launch_reproj_err(int model, int one_focal, const ProblemDev& pb, const double* intr_dev, const double* poses,
                              double* err, int64_t n_obs, cudaStream_t s) {
  return dispatch_model(model, one_focal, [&](auto M, auto OF) {
    const int nb = (int)((n_obs + 127) / 128);
    k_reproj_err<decltype(M)::value, decltype(OF)::value><<<nb, 128, 0, s>>>(pb, intr_dev, poses, err, n_obs);
    return cudaGetLastError();
  });
}

template <class F>
static cudaError_t dispatch_d(int D, F&& f) {
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

// batch handles: one thread per frame, per-frame contributions reduced per problem by k_segreduce. (The single-problem
// reduction is k_schur2, ccrs_loop.cu.)
cudaError_t launch_schur(int D, const SchurParams& prm, cudaStream_t s) {
  const int nb = (prm.pb.n_frames + kSchurThreads - 1) / kSchurThreads;
  return dispatch_d(D, [&](auto DD) {
    constexpr int d = decltype(DD)::value;
    constexpr size_t kStage = (size_t)(d + 7) * (d + 8) / 2 * kSchurThreads * sizeof(double);   // one packed block per thread
    auto kern = k_schur<d, true>;
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kStage);
      if (e != cudaSuccess) return e;
      configured = true;
    }
    kern<<<nb, kSchurThreads, kStage, s>>>(prm, nullptr);
    return cudaGetLastError();
  });
}

cudaError_t launch_sum_partials(const double* partials, int n_part, int NV, double* out, volatile double* host_out,
                                double seq, cudaStream_t s) {
  if (host_out) {
    if (NV > 1024) return cudaErrorInvalidValue;
    k_sum_partials<<<1, (NV + 31) / 32 * 32, 0, s>>>(partials, n_part, NV, out, host_out, seq);
  } else {
    k_sum_partials<<<(NV + 127) / 128, 128, 0, s>>>(partials, n_part, NV, out, nullptr, 0.0);
  }
  return cudaGetLastError();
}

cudaError_t launch_backsub(int D, const BacksubParams& prm, cudaStream_t s) {
  const bool batch = prm.pb.frame_problem != nullptr;
  const int nb = (prm.pb.n_frames + 127) / 128;
  return dispatch_d(D, [&](auto DD) {
    constexpr int d = decltype(DD)::value;
    if (batch) k_backsub<d, true><<<nb, 128, 0, s>>>(prm);
    else k_backsub<d, false><<<nb, 128, 0, s>>>(prm);
    return cudaGetLastError();
  });
}

cudaError_t launch_compute_scale(int D, const ProblemDev& pb, int which, double* pose_scale, double* frame_colsq,
                                 cudaStream_t s) {
  const bool batch = pb.frame_problem != nullptr;
  const int nb = (pb.n_frames + 127) / 128;
  return dispatch_d(D, [&](auto DD) {
    constexpr int d = decltype(DD)::value;
    if (batch) k_compute_scale<d, true><<<nb, 128, 0, s>>>(pb, which, pose_scale, frame_colsq);
    else k_compute_scale<d, false><<<nb, 128, 0, s>>>(pb, which, pose_scale, frame_colsq);
    return cudaGetLastError();
  });
}

cudaError_t launch_segreduce(const double* in, int NV, int Fs, const int32_t* seg_off, int n_seg, double* out,
                             cudaStream_t s) {
  k_segreduce<<<n_seg, kSegThreads, 0, s>>>(in, NV, Fs, seg_off, out);
  return cudaGetLastError();
}

cudaError_t launch_trial_stats(const ProblemDev& pb, int rr_idx, int mode, const double* frame_md, double* stat_out,
                               cudaStream_t s) {
  k_trial_stats<<<pb.n_problems, kSegThreads, 0, s>>>(pb, rr_idx, mode, frame_md, stat_out);
  return cudaGetLastError();
}

cudaError_t launch_flip_cur(int32_t* cur, const unsigned char* mask_dev, int n_problems, cudaStream_t s) {
  k_flip_cur<<<(n_problems + 127) / 128, 128, 0, s>>>(cur, mask_dev, n_problems);
  return cudaGetLastError();
}

cudaError_t launch_arm(double* p, size_t n, cudaStream_t s) {
  k_arm<<<(unsigned)std::min<size_t>((n + 255) / 256, 1024), 256, 0, s>>>(p, n);
  return cudaGetLastError();
}

cudaError_t launch_fp64_peak(double* out, int n_ctas, int iters, cudaStream_t s) {
  k_fp64_peak<<<n_ctas, 256, 0, s>>>(out, iters);
  return cudaGetLastError();
}

cudaError_t launch_l2_flush(double* buf, size_t n, cudaStream_t s) {
  k_l2_flush<<<148 * 8, 256, 0, s>>>(buf, n);
  return cudaGetLastError();
}

}  // namespace ccrs

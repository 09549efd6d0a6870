// ccrs_api.cu — the C ABI (include/ccrs_b200.h): problem handle, device buffers, kernel sequencing,
// the CUDA implementation of ccrs_backend, NCCL exchange of the reduced system, calib_camera entry point.
#include "../../include/ccrs_b200.h"
#include "ccrs_kernels.cuh"
#include "ccrs_rule.h"

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <dlfcn.h>
#include <atomic>
#include <limits>
#include <map>
#include <mutex>
#include <unordered_map>
#include <vector>

using namespace ccrs;

// ---- minimal NCCL surface, resolved at run time from the libnccl.so.2 the host process already loaded ----
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
namespace {
struct NcclApi {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;   // dtype 0 = ncclInt8
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool ok = false;
};
constexpr int kNcclFloat64 = 8;  // ncclFloat64 / ncclDouble
constexpr int kNcclSum = 0;

NcclApi& nccl() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    api.h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!api.h) api.h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (api.h) {
      api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(api.h, "ncclGetUniqueId");
      api.CommInitRank = (decltype(api.CommInitRank))dlsym(api.h, "ncclCommInitRank");
      api.CommDestroy = (decltype(api.CommDestroy))dlsym(api.h, "ncclCommDestroy");
      api.AllGather = (decltype(api.AllGather))dlsym(api.h, "ncclAllGather");
      api.AllReduce = (decltype(api.AllReduce))dlsym(api.h, "ncclAllReduce");
      api.GetErrorString = (decltype(api.GetErrorString))dlsym(api.h, "ncclGetErrorString");
      api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather && api.AllReduce;
    }
  }
  return api;
}

// one process = one GPU = one communicator, shared by every handle of the process (created once: ncclCommInitRank
// costs ~100 ms, far more than a whole calibration)
struct GlobalComm { ncclComm_t comm = nullptr; int rank = 0, world = 1; } g_comm;

// Process-wide peer-memory exchange buffer (PeerXchg in ccrs_kernels.cuh): [area][parity][rank][kXchgMaxVals] doubles,
// cudaMalloc'd (not pooled: it is exported with cudaIpcGetMemHandle), armed, and every other rank's buffer opened with
// cudaIpcOpenMemHandle. Areas: 0 = K3 reduced system, 1 = K2 {model decrease, cost}. Exchanges alternate parity per area
// (all ranks issue the same sequence), so a rank that races ahead never overwrites a slot its peer has not consumed.
constexpr int kXchgAreas = 3;   // 0: host-driven K3, 1: host-driven K2 statistics, 2: K3 of the device-driven loop
struct GlobalPeer {
  bool ok = false;
  double* local = nullptr;
  double* peer[kXchgMaxRanks] = {};
  unsigned long count[kXchgAreas] = {};
  // area 2 alternates its slot parity on a DEVICE-side count of executed exchanges (slots of the device-driven loop
  // that find nothing to do exchange nothing): one 8-byte word behind the slot areas of the local buffer
  unsigned int* dev_count() const { return reinterpret_cast<unsigned int*>(local + doubles()); }
  size_t doubles() const { return (size_t)kXchgAreas * 2 * kXchgMaxRanks * kXchgMaxVals; }
  size_t bytes() const { return (doubles() + 2) * sizeof(double); }
} g_peer;

thread_local char g_err[512] = "";
int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}
#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t _e = (call);                                                                         \
    if (_e != cudaSuccess) return fail(CCRS_ERR_CUDA, "%s: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
  } while (0)

// Process-wide caching allocators: cudaMalloc / cudaHostAlloc / cudaFree cost 0.1-1 ms each and a calibration call
// creates ~30 buffers, far more than the whole solve (a few hundred microseconds). Freed blocks are kept and reused.
struct BlockPool {
  bool host;
  std::mutex mu;
  std::multimap<size_t, void*> free_blocks;
  std::unordered_map<void*, size_t> live;
  explicit BlockPool(bool h) : host(h) {}
  cudaError_t get(size_t bytes, void** out) {
    bytes = std::max<size_t>((bytes + 255) / 256 * 256, 256);
    {
      std::lock_guard<std::mutex> g(mu);
      auto it = free_blocks.lower_bound(bytes);
      if (it != free_blocks.end() && it->first <= 2 * bytes + (1u << 20)) {
        *out = it->second;
        live[*out] = it->first;
        free_blocks.erase(it);
        return cudaSuccess;
      }
    }
    cudaError_t e = host ? cudaHostAlloc(out, bytes, cudaHostAllocMapped | cudaHostAllocPortable) : cudaMalloc(out, bytes);
    if (e != cudaSuccess) {  // out of memory: drop the cache and retry once
      trim();
      e = host ? cudaHostAlloc(out, bytes, cudaHostAllocMapped | cudaHostAllocPortable) : cudaMalloc(out, bytes);
    }
    if (e == cudaSuccess) { std::lock_guard<std::mutex> g(mu); live[*out] = bytes; }
    return e;
  }
  void put(void* p) {
    if (!p) return;
    std::lock_guard<std::mutex> g(mu);
    auto it = live.find(p);
    if (it == live.end()) return;
    free_blocks.emplace(it->second, p);
    live.erase(it);
  }
  void trim() {
    std::lock_guard<std::mutex> g(mu);
    for (auto& kv : free_blocks) { if (host) cudaFreeHost(kv.second); else cudaFree(kv.second); }
    free_blocks.clear();
  }
};
BlockPool& dev_pool(int device) {
  static BlockPool* pools[64] = {nullptr};
  static std::mutex mu;
  std::lock_guard<std::mutex> g(mu);
  if (!pools[device & 63]) pools[device & 63] = new BlockPool(false);
  return *pools[device & 63];
}
BlockPool& host_pool() { static BlockPool pool(true); return pool; }

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  int dev = 0;
  cudaError_t alloc(size_t count) {
    n = count;
    cudaGetDevice(&dev);
    return dev_pool(dev).get(std::max<size_t>(count, 1) * sizeof(T), (void**)&p);
  }
  void release() { if (p) dev_pool(dev).put(p); p = nullptr; }
};
template <class T>
struct PinBuf {   // page-locked, mapped into the device address space (UVA: same pointer on host and device)
  T* p = nullptr;
  cudaError_t alloc(size_t count) { return host_pool().get(std::max<size_t>(count, 1) * sizeof(T), (void**)&p); }
  void release() { if (p) host_pool().put(p); p = nullptr; }
};
}  // namespace

struct ccrs_problem {
  int model = 0, width = 0, height = 0, one_focal = 0;
  int D = 0, NA = 0, NBLK = 0, NACC = 0, NRED = 0, NOUT = 0;
  int n_frames = 0, n_problems = 1, Fs = 0;
  int64_t n_obs = 0;
  bool batch = false;
  bool f32 = false;               // observation arrays stored as floats (ccrs_problem_create_f32)
  int device = 0;
  int n_sms = 148;
  cudaStream_t stream = nullptr;
  int G = 1, FPW = 32, n_lin_ctas = 0, n_schur_ctas = 0;
  bool use_mma = false;           // K2 with the Gram block on the FP64 tensor path (models with d + 7 >= 14 columns)
  int n_mma_ctas = 0;
  DevBuf<double> frame_stat;      // ... its per-frame {model decrease, cost} slots
  DevBuf<unsigned int> chunk_cnt; // ... frames done per chunk of 16
  double huber = 1.0;
  int64_t launches = 0;

  DevBuf<double> x, y, z, u, v;
  DevBuf<int32_t> corner_id;      // board-format problems: corner id of every observation (expanded to x, y, z on the device)
  DevBuf<float> board_dev;        // ... and the board table [n_board][3]
  const int32_t* h_corner_id = nullptr;   // set by ccrs_problem_create_board_f32 for the duration of the upload
  const float* h_board = nullptr;
  int n_board = 0;
  DevBuf<int32_t> frame_offsets, frame_problem, problem_frame_offsets, obs_frame, cur, acc_to_blk;
  DevBuf<double> poses[2], blocks[2], frame_cost[2];
  DevBuf<double> elim, frame_red, pose_scale, frame_md, cta_part;
#ifdef CCRS_K2_TIMING
  DevBuf<double> k2_dbg;   // [n_warps][10] int64 phase clocks
  DevBuf<double> k3_dbg;   // [n_warps][8]
#endif
  DevBuf<double> red_out, stat_out, gather, intr_dev, ya_dev, u_dev, scale_dev, l2_flush;
  DevBuf<double> rdv;             // one word for the bench's cross-rank rendezvous
  DevBuf<double> ctl_dev;         // LoopCtl of the device-driven loop (single problem)
  DevBuf<double> s2_part;         // K3 (k_schur2) per-CTA partials [n_ctas][NRED]: self-validating slots, armed
  DevBuf<double> bctl_dev, hist_dev;   // batch: per-problem controller state (BatchCtl), error history of problem 0
  PinBuf<double> h_bctl, h_bstatus;    // batch: staging of the controller state; mapped status ring [kRecSlots][2]
  PinBuf<double> h_rec, h_ctl;    // mapped: record ring [kRecSlots][kRecStride]; staging copy of the control block
  double h_scale[9] = {1, 1, 1, 1, 1, 1, 1, 1, 1};   // host copy of the Jacobi scaling of the intrinsic columns (single problem)
  long loop_records = 0;          // records consumed from the ring since creation
  DevBuf<unsigned char> mask_dev;
  DevBuf<unsigned int> tickets;   // [0] K2 statistics, [1] K3 reduction
  PinBuf<double> h_colsq;         // mapped: single problem [D] squared intrinsic column norms (Jacobi scaling)
  PinBuf<double> h_red, h_stat;   // mapped: single problem [NRED + 1] / [4] (last = sequence number); batch: plain D2H targets
  bool have_scale = false, have_obs_frame = false;
  bool fixed_poses = false;       // poses are constants (ccrs_set_fixed_poses)
  // Speculative K3 (single problem, speculative LM): launched right behind the trial-point K2, before its statistics
  // are known, for the outcome that dominates a converging run — step accepted with gain ratio >= 0.937, i.e.
  // u_next = u / 3 exactly. If the controller then asks for exactly that reduction the result is already in flight (the
  // launch call, the launch latency and K3 itself leave the critical path); otherwise the speculative result is
  // ignored and K3 is launched again with the right damping. The decision itself stays on the host.
  struct SpecK3 { bool valid = false; double u = 0.0; int use_scale = 0; double mn = 0.0, mx = 0.0; int cur_after = 0, area = 0; } spec;
  int red_area = 0;               // which half of h_red the last K3 publishes into (alternates per launch)
  double last_mn = 0.0, last_mx = 0.0;
  bool have_last_reduce = false;
  std::vector<int32_t> h_frame_offsets, h_problem_frame_offsets;
  bool last_use_scale = false;    // state of the last reduce(), needed by the back-substitution
  int cur_val = 0;                // single problem: which state buffer is current (host-tracked, passed by value)
  double seq = 0.0;               // sequence number of the last published result
  // back-substitution deferred into the prologue of the next K2 launch
  bool pend = false;
  int pend_in_place = 0;
  bool pend_active = false;
  double pend_ya[9] = {0}, pend_ya_step[9] = {0}, pend_u = 0.0;
  // communicator
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1, deterministic = 1;

  ProblemDev dev() const {
    ProblemDev d{};
    d.x = x.p; d.y = y.p; d.z = z.p; d.u = u.p; d.v = v.p;
    d.f32 = f32 ? 1 : 0;
    d.frame_offsets = frame_offsets.p;
    d.frame_problem = batch ? frame_problem.p : nullptr;
    d.problem_frame_offsets = problem_frame_offsets.p;
    d.obs_frame = obs_frame.p;
    d.cur = batch ? cur.p : nullptr;
    d.cur_val = cur_val;
    for (int i = 0; i < 2; ++i) { d.poses[i] = poses[i].p; d.blocks[i] = blocks[i].p; d.frame_cost[i] = frame_cost[i].p; }
    d.n_frames = n_frames; d.n_problems = n_problems; d.Fs = Fs;
    d.huber_delta = huber;
    return d;
  }
};

namespace {

// Lanes per frame G (1..32; a warp owns 32/G frames): the G that minimises
// (waves of CTAs) x (observations per lane + per-warp overhead) for this device.
void choose_slicing(ccrs_problem* p) {
  const int slots = p->n_sms * kLinCtasPerSm;  // resident CTAs (kLinWarps frame groups each)
  int max_cnt = 1;
  for (int f = 0; f < p->n_frames; ++f) max_cnt = std::max(max_cnt, p->h_frame_offsets[f + 1] - p->h_frame_offsets[f]);
  double best = 1e300;
  int bestG = 1;
  const bool pairs = lin_uses_pairs(p->model, p->one_focal);   // lane-pair K2: the two lanes of a pair share a frame
  for (int G = 1; G <= 32; ++G) {
    if (pairs && (G & 1)) continue;
    const int fpw = 32 / G;
    const int warps = (p->n_frames + fpw - 1) / fpw;
    const int ctas = (warps + kLinWarps - 1) / kLinWarps;
    const int waves = (ctas + slots - 1) / slots;
    const int per_lane = (max_cnt + G - 1) / G;
    // cost model: per-lane observations dominate; a per-warp constant covers prologue (pose maths), pipeline fill,
    // basis change and the shared-memory reduction (in units of main-loop iterations); the reduction sums G slices per
    // entry (+0.1 G) and has an unrolled form only for the slice counts listed in slices_reduce_store (+7 otherwise:
    // tools/k2_sweep.py at 875 frames — G = 16: 13.2 us, G = 32: 14.6 us, where the constant alone picked 32)
    const bool unrolled = G <= 6 || G == 8 || G == 10 || G == 16;
    const double cost = (double)waves * (per_lane + kLinOverheadIters + 0.1 * G + (unrolled ? 0.0 : 7.0));
    if (cost < best - 1e-12) { best = cost; bestG = G; }
  }
  if (const char* e = getenv("CCRS_FORCE_G")) { const int g = atoi(e); if (g >= 1 && g <= 32 && !(pairs && (g & 1))) bestG = g; }
  p->G = bestG;
  p->FPW = 32 / bestG;
  const int warps = (p->n_frames + p->FPW - 1) / p->FPW;
  p->n_lin_ctas = (warps + kLinWarps - 1) / kLinWarps;
  // tensor-core variant: one resident wave of warps, frames handed out dynamically
  const char* em = getenv("CCRS_K2_MMA");
  p->use_mma = lin_mma_available(p->model, p->one_focal) && !(em && atoi(em) == 0);
  p->n_mma_ctas = lin_mma_ctas(p->n_sms, p->n_frames);
}

// K2 launch: the tensor-core variant where it applies (never for the cost-only pass), else the register-accumulator kernel
cudaError_t launch_k2(ccrs_problem* p, bool batch, bool cost_only, LinParams& prm) {
  prm.frame_ctr = reinterpret_cast<unsigned long long*>(p->tickets.p + 4);
  prm.frame_stat = p->frame_stat.p;
  prm.chunk_cnt = p->chunk_cnt.p;
  if (p->use_mma && !cost_only) return launch_linearize_mma(p->model, p->one_focal, batch, prm, p->n_mma_ctas, p->stream);
  return launch_linearize(p->model, p->one_focal, batch, cost_only, prm, p->n_lin_ctas, p->stream);
}

int upload_problem(ccrs_problem* p, int n_problems, const int32_t* problem_frame_offsets, int n_frames,
                   const int32_t* frame_offsets, const void* x, const void* y, const void* z, const void* u,
                   const void* v) {
  p->n_frames = n_frames;
  p->n_problems = n_problems;
  p->n_obs = frame_offsets[n_frames];
  p->Fs = (n_frames + 31) / 32 * 32;
  p->h_frame_offsets.assign(frame_offsets, frame_offsets + n_frames + 1);
  if (problem_frame_offsets) p->h_problem_frame_offsets.assign(problem_frame_offsets, problem_frame_offsets + n_problems + 1);
  else p->h_problem_frame_offsets = {0, n_frames};
  choose_slicing(p);
  p->n_schur_ctas = (n_frames + 127) / 128;
  const size_t N = (size_t)p->n_obs, F = (size_t)n_frames, Fs = (size_t)p->Fs, P = (size_t)n_problems;
  cudaStream_t s = p->stream;
  const size_t esz = p->f32 ? 4 : 8, nel = p->f32 ? (N + 1) / 2 : N;   // floats are packed into the double-typed buffers
  CK(p->x.alloc(nel)); CK(p->y.alloc(nel)); CK(p->z.alloc(nel)); CK(p->u.alloc(nel)); CK(p->v.alloc(nel));
  // start the big copies first so everything below overlaps with them
  if (p->h_corner_id) {
    // board format: 12 bytes per observation cross PCIe (corner id, u, v) instead of 20; p3d = board[id] is expanded
    // into the f32 x, y, z arrays by one small kernel, so the linearisation kernels see the usual layout
    CK(p->corner_id.alloc(N)); CK(p->board_dev.alloc((size_t)3 * p->n_board));
    CK(cudaMemcpyAsync(p->board_dev.p, p->h_board, (size_t)3 * p->n_board * sizeof(float), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->corner_id.p, p->h_corner_id, N * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    // the ids are range-checked where they are expanded (a host pass over a million ids would cost more than the
    // transfer saves): the kernel raises a flag in mapped host memory, read after the upload's final synchronise
    CK(p->h_colsq.alloc(P * p->D + 1));
    p->h_colsq.p[0] = 0.0;
    CK(launch_expand_board(p->corner_id.p, p->board_dev.p, p->n_board, (int64_t)N, reinterpret_cast<float*>(p->x.p),
                           reinterpret_cast<float*>(p->y.p), reinterpret_cast<float*>(p->z.p), p->h_colsq.p, s));
    p->launches++;
  } else {
    CK(cudaMemcpyAsync(p->x.p, x, N * esz, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->y.p, y, N * esz, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->z.p, z, N * esz, cudaMemcpyHostToDevice, s));
  }
  CK(cudaMemcpyAsync(p->u.p, u, N * esz, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(p->v.p, v, N * esz, cudaMemcpyHostToDevice, s));
  CK(p->frame_offsets.alloc(F + 1));
  CK(p->problem_frame_offsets.alloc(P + 1));
  CK(p->cur.alloc(P));
  CK(p->acc_to_blk.alloc(p->NACC));
  CK(p->tickets.alloc(8));        // [0] K2 statistics, [1] K3 reduction, [2] batch: active problems, [3] batch: status ticket, [4..5] K2 frame counter (64 bit)
  CK(cudaMemsetAsync(p->tickets.p, 0, 8 * sizeof(unsigned int), s));
  if (p->use_mma) {
    const size_t n_chunks = (F + 15) / 16;
    CK(p->frame_stat.alloc(2 * Fs)); CK(p->chunk_cnt.alloc(n_chunks));
    CK(launch_arm(p->frame_stat.p, 2 * Fs, s));
    CK(cudaMemsetAsync(p->chunk_cnt.p, 0, n_chunks * sizeof(unsigned int), s));
  }
  for (int i = 0; i < 2; ++i) {
    CK(p->poses[i].alloc(F * 6));
    CK(p->blocks[i].alloc((size_t)p->NBLK * Fs));
    CK(p->frame_cost[i].alloc(Fs));
    CK(cudaMemsetAsync(p->blocks[i].p, 0, (size_t)p->NBLK * Fs * sizeof(double), s));  // structural zeros stay zero
    CK(cudaMemsetAsync(p->poses[i].p, 0, F * 6 * sizeof(double), s));
  }
  CK(p->elim.alloc((size_t)(6 * p->D + 18) * Fs));
  CK(p->frame_red.alloc(std::max((size_t)p->NRED * Fs, (size_t)p->n_schur_ctas * p->NRED)));
  CK(p->pose_scale.alloc(6 * Fs));
  // K2 requests a frame's elimination record and pose scales together with the control block, before it knows whether a
  // step is pending (one memory round trip instead of two): defined values for the first linearisation, which discards them
  CK(cudaMemsetAsync(p->elim.p, 0, (size_t)(6 * p->D + 18) * Fs * sizeof(double), s));
  CK(cudaMemsetAsync(p->pose_scale.p, 0, 6 * Fs * sizeof(double), s));
  CK(p->frame_md.alloc(Fs));
  const size_t n_k2_warps = (size_t)std::max(p->n_lin_ctas, p->n_mma_ctas) * kLinWarps;
  CK(p->cta_part.alloc(2 * n_k2_warps));
  CK(launch_arm(p->cta_part.p, 2 * n_k2_warps, s));
  CK(p->red_out.alloc(P * p->NRED));
  CK(p->stat_out.alloc(P * 2));
  CK(p->intr_dev.alloc(P * p->D)); CK(p->ya_dev.alloc(P * p->D)); CK(p->u_dev.alloc(P)); CK(p->scale_dev.alloc(P * p->D));
  CK(p->mask_dev.alloc(P));
  CK(p->h_red.alloc(2 * (P * p->NRED + 1))); CK(p->h_stat.alloc(P * 2 + 2));
  if (!p->h_colsq.p) CK(p->h_colsq.alloc(P * p->D + 1));
  if (!p->batch) {
    CK(p->ctl_dev.alloc(sizeof(LoopCtl) / 8)); CK(p->h_ctl.alloc(sizeof(LoopCtl) / 8));
    CK(p->h_rec.alloc((size_t)kRecSlots * kRecStride));
    const size_t n_part = (size_t)schur2_ctas(n_frames) * p->NRED;
    CK(p->s2_part.alloc(n_part));
    CK(launch_arm(p->s2_part.p, n_part, s));
  }
  p->h_red.p[p->NRED] = -1.0; p->h_stat.p[2] = -1.0; p->h_stat.p[3] = -1.0;
  CK(cudaMemsetAsync(p->cur.p, 0, P * sizeof(int32_t), s));
  CK(cudaMemsetAsync(p->u_dev.p, 0, P * sizeof(double), s));
  CK(cudaMemcpyAsync(p->frame_offsets.p, frame_offsets, (F + 1) * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(p->problem_frame_offsets.p, p->h_problem_frame_offsets.data(), (P + 1) * 4, cudaMemcpyHostToDevice, s));
  std::vector<int32_t> tbl(p->NACC);
  fill_acc_to_blk(p->model, p->one_focal, tbl.data());
  CK(cudaMemcpyAsync(p->acc_to_blk.p, tbl.data(), tbl.size() * 4, cudaMemcpyHostToDevice, s));
  std::vector<int32_t> fp;
  if (p->batch) {
    fp.resize(F);
    for (int b = 0; b < n_problems; ++b)
      for (int f = problem_frame_offsets[b]; f < problem_frame_offsets[b + 1]; ++f) fp[f] = b;
    CK(p->frame_problem.alloc(F));
    CK(cudaMemcpyAsync(p->frame_problem.p, fp.data(), F * 4, cudaMemcpyHostToDevice, s));
  }
  CK(cudaStreamSynchronize(s));
  if (p->h_corner_id && p->h_colsq.p[0] != 0.0) return fail(CCRS_ERR_INVALID, "a corner id lies outside the board table (%d corners)", p->n_board);
  return 0;
}

// streams are cached per process too (creation costs ~10 us, destruction synchronises)
std::mutex g_stream_mu;
std::vector<std::pair<int, cudaStream_t>> g_streams;
cudaError_t get_stream(int device, cudaStream_t* out) {
  {
    std::lock_guard<std::mutex> g(g_stream_mu);
    for (size_t i = 0; i < g_streams.size(); ++i)
      if (g_streams[i].first == device) { *out = g_streams[i].second; g_streams.erase(g_streams.begin() + i); return cudaSuccess; }
  }
  return cudaStreamCreateWithFlags(out, cudaStreamNonBlocking);
}
void put_stream(int device, cudaStream_t s) { std::lock_guard<std::mutex> g(g_stream_mu); g_streams.emplace_back(device, s); }

std::atomic<unsigned long long> g_seq{0};
double next_seq() { return (double)(++g_seq); }

struct DevInfo { bool known = false; int major = 0, minor = 0, sms = 0; } g_dev_info[64];

int create_common(ccrs_problem** out, int model, int width, int height, int xy_same_focal, int n_problems,
                  const int32_t* problem_frame_offsets, int n_frames, const int32_t* frame_offsets, const void* x,
                  const void* y, const void* z, const void* u, const void* v, double huber_delta, int device_id,
                  bool batch, bool f32 = false, const int32_t* corner_id = nullptr, const float* board = nullptr, int n_board = 0) {
  if (!out || !frame_offsets || !u || !v || n_frames <= 0 || n_problems <= 0 || (!corner_id && (!x || !y || !z)))
    return fail(CCRS_ERR_INVALID, "null pointer or empty problem");
  if (frame_offsets[0] != 0) return fail(CCRS_ERR_INVALID, "frame_offsets[0] must be 0");
  for (int f = 0; f < n_frames; ++f)
    if (frame_offsets[f + 1] == frame_offsets[f]) return fail(CCRS_ERR_INVALID, "frame %d has no observations (drop it: its pose block would be singular)", f);
  if (corner_id && (!board || n_board <= 0)) return fail(CCRS_ERR_INVALID, "board table missing");
  if (model < 0 || model > 5) return fail(CCRS_ERR_INVALID, "unknown model %d", model);
  for (int f = 0; f < n_frames; ++f)
    if (frame_offsets[f + 1] < frame_offsets[f]) return fail(CCRS_ERR_INVALID, "frame_offsets not monotone at %d", f);
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) {
    cudaGetLastError();
    return fail(CCRS_ERR_NO_DEVICE, "no CUDA device: this library has no CPU path");
  }
  if (device_id < 0 || device_id >= n_dev || device_id >= 64) return fail(CCRS_ERR_INVALID, "device %d out of range", device_id);
  DevInfo di;
  {
    static std::mutex dev_info_mu;
    std::lock_guard<std::mutex> dev_info_lock(dev_info_mu);   // handles are created from several host threads
    DevInfo& g = g_dev_info[device_id];
    if (!g.known) {
      CK(cudaDeviceGetAttribute(&g.major, cudaDevAttrComputeCapabilityMajor, device_id));
      CK(cudaDeviceGetAttribute(&g.minor, cudaDevAttrComputeCapabilityMinor, device_id));
      CK(cudaDeviceGetAttribute(&g.sms, cudaDevAttrMultiProcessorCount, device_id));
      g.known = true;
    }
    di = g;
  }
  if (di.major != 10) return fail(CCRS_ERR_NO_DEVICE, "device %d is sm_%d%d; kernels are built for sm_100a only", device_id, di.major, di.minor);
  CK(cudaSetDevice(device_id));
  ccrs_problem* p = new ccrs_problem();
  p->model = model; p->width = width; p->height = height; p->one_focal = xy_same_focal ? 1 : 0;
  p->huber = huber_delta; p->device = device_id; p->batch = batch; p->n_sms = di.sms; p->f32 = f32;
  p->h_corner_id = corner_id; p->h_board = board; p->n_board = n_board;
  model_dims(model, p->one_focal, &p->D, &p->NA, &p->NBLK, &p->NACC);
  p->NRED = nred_of(p->D);
  p->NOUT = p->D * p->D + 3 * p->D + 1;
  cudaError_t e = get_stream(device_id, &p->stream);
  if (e != cudaSuccess) { delete p; return fail(CCRS_ERR_CUDA, "stream: %s", cudaGetErrorString(e)); }
  int st = upload_problem(p, n_problems, problem_frame_offsets, n_frames, frame_offsets, x, y, z, u, v);
  p->h_corner_id = nullptr; p->h_board = nullptr;   // caller-owned: not retained
  if (st != 0) { ccrs_problem_destroy(p); return st; }
  *out = p;
  return 0;
}

// Host-side phase trace of the single-problem LM iteration (ccrs_step_trace): where the time between kernels goes.
//   0 K3 launch call | 1 K3 execution + publish latency (launch return -> sequence number seen) | 2 host: unpack,
//   d x d solve, trial point | 3 K2 launch call | 4 K2 execution + publish latency | 5 host: accept/reject, bookkeeping
struct StepTrace {
  bool on = false;
  double acc[6] = {0, 0, 0, 0, 0, 0};
  long n = 0;
  double t_mark = 0.0;
  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec * 1e6 + ts.tv_nsec * 1e-3; }
  void mark(int slot) { if (!on) return; const double t = now(); if (slot >= 0 && t_mark > 0.0) acc[slot] += t - t_mark; t_mark = t; }
};
StepTrace g_trace;

// speculative K3 (see ccrs_problem::SpecK3): CCRS_SPEC_K3=0 disables it; counters for tests / tools
const bool g_spec_enabled = [] { const char* e = getenv("CCRS_SPEC_K3"); return !(e && atoi(e) == 0); }();
std::atomic<long> g_spec_launched{0}, g_spec_hits{0};

// Results the host needs every iteration come back through mapped pinned host memory WITHOUT any device-side fence
// (a system-scope fence costs ~2-3 us per use): before the launch the host arms the n result words with a sentinel
// bit pattern (a NaN payload no computation produces; the Cholesky-failure poison is the canonical quiet NaN); the
// kernel simply stores its n results; the host spins until none of the n words is the sentinel. Each word is written
// once with a single 8-byte store, so a word is either the sentinel or the final value. Falls back to the stream
// status so a failed launch cannot hang the host.
constexpr uint64_t kSentinelBits = 0x7ff8dead5e471e15ULL;
void arm_payload(volatile double* dst, int n) {
  volatile uint64_t* w = reinterpret_cast<volatile uint64_t*>(dst);
  for (int i = 0; i < n; ++i) w[i] = kSentinelBits;
  std::atomic_thread_fence(std::memory_order_seq_cst);
}
int wait_payload(ccrs_problem* p, volatile double* src, int n) {
  volatile uint64_t* w = reinterpret_cast<volatile uint64_t*>(src);
  for (unsigned long spins = 1;; ++spins) {
    int i = n - 1;
    while (i >= 0 && w[i] != kSentinelBits) --i;
    if (i < 0) return 0;
    if ((spins & 0x3fff) == 0) {
      cudaError_t e = cudaStreamQuery(p->stream);
      if (e == cudaSuccess) {
        i = n - 1;
        while (i >= 0 && w[i] != kSentinelBits) --i;
        if (i < 0) return 0;
        return fail(CCRS_ERR_CUDA, "stream drained but %d result words were never published", n);
      }
      if (e != cudaErrorNotReady) return fail(CCRS_ERR_CUDA, "stream error while waiting: %s", cudaGetErrorString(e));
    }
  }
}

// peer-memory exchange usable for this handle / payload?
bool use_peer(const ccrs_problem* p, size_t count) {
  return p->comm && g_peer.ok && p->deterministic && !p->batch && count <= (size_t)kXchgMaxVals;
}
void fill_peer(const ccrs_problem* p, int area, PeerXchg* px) {
  for (int r = 0; r < kXchgMaxRanks; ++r) px->peer[r] = g_peer.peer[r];
  px->world = p->world; px->rank = p->rank;
  const unsigned long parity = g_peer.count[area]++ & 1ul;
  px->off = (int)(((size_t)area * 2 + parity) * kXchgMaxRanks * kXchgMaxVals);
}

// sum `count` doubles across ranks on the device (stream-ordered), in place; optionally publish to mapped host memory
int exchange(ccrs_problem* p, double* buf, size_t count, volatile double* host_out, double seq) {
  if (host_out) arm_payload(host_out, (int)count);
  if (!p->comm) {
    if (host_out) { CK(launch_sum_partials(buf, 1, (int)count, buf, host_out, seq, p->stream)); p->launches++; }
    return 0;
  }
  NcclApi& n = nccl();
  if (p->deterministic) {
    if (p->gather.n < count * p->world) { p->gather.release(); CK(p->gather.alloc(count * p->world)); }
    int r = n.AllGather(buf, p->gather.p, count, kNcclFloat64, p->comm, p->stream);
    if (r != 0) return fail(CCRS_ERR_COMM, "ncclAllGather: %s", n.GetErrorString ? n.GetErrorString(r) : "?");
    CK(launch_sum_partials(p->gather.p, p->world, (int)count, buf, host_out, seq, p->stream));  // rank order
    p->launches++;
  } else {
    int r = n.AllReduce(buf, buf, count, kNcclFloat64, kNcclSum, p->comm, p->stream);
    if (r != 0) return fail(CCRS_ERR_COMM, "ncclAllReduce: %s", n.GetErrorString ? n.GetErrorString(r) : "?");
    if (host_out) { CK(launch_sum_partials(buf, 1, (int)count, buf, host_out, seq, p->stream)); p->launches++; }
  }
  return 0;
}

int upload_intr(ccrs_problem* p, const double* intr) {
  const size_t n = (size_t)p->n_problems * p->D;
  CK(cudaMemcpyAsync(p->intr_dev.p, intr, n * 8, cudaMemcpyHostToDevice, p->stream));
  return 0;
}

// standalone K4 (used by the public ccrs_backsub and to flush a deferred back-substitution)
int launch_backsub_now(ccrs_problem* p, int in_place, bool want_md) {
  if (!p->batch) {  // the single-problem hot path keeps y_a / u on the host (kernel arguments); upload them for K4
    // K3 (k_schur2) leaves the elimination record WITHOUT the intrinsic Jacobi scaling: the step K4 applies is D_a y_a
    for (int i = 0; i < p->D; ++i) p->pend_ya_step[i] = p->pend_ya[i] * (p->last_use_scale ? p->h_scale[i] : 1.0);
    CK(cudaMemcpyAsync(p->ya_dev.p, p->pend_ya_step, (size_t)p->D * 8, cudaMemcpyHostToDevice, p->stream));
    CK(cudaMemcpyAsync(p->u_dev.p, &p->pend_u, 8, cudaMemcpyHostToDevice, p->stream));
  }
  BacksubParams prm{};
  prm.pb = p->dev();
  prm.elim = p->elim.p;
  prm.y_a = p->ya_dev.p;
  prm.u_dev = p->u_dev.p;
  prm.pose_scale = p->last_use_scale ? p->pose_scale.p : nullptr;
  prm.frame_md = want_md ? p->frame_md.p : nullptr;
  prm.in_place = in_place;
  prm.active = p->pend_active ? p->mask_dev.p : nullptr;
  CK(launch_backsub(p->D, prm, p->stream));
  p->launches++;
  return 0;
}

// record a back-substitution; it runs in the prologue of the next K2 launch (or on flush)
int defer_backsub(ccrs_problem* p, const double* y_a, const double* u, const unsigned char* active, int in_place) {
  const int P = p->n_problems, D = p->D;
  if (p->pend) { p->pend = false; int st = launch_backsub_now(p, p->pend_in_place, false); if (st) return st; }
  p->pend_active = false;
  if (p->batch) {
    CK(cudaMemcpyAsync(p->ya_dev.p, y_a, (size_t)P * D * 8, cudaMemcpyHostToDevice, p->stream));
    if (u) CK(cudaMemcpyAsync(p->u_dev.p, u, (size_t)P * 8, cudaMemcpyHostToDevice, p->stream));
    else CK(cudaMemsetAsync(p->u_dev.p, 0, (size_t)P * 8, p->stream));
    if (active) {
      CK(cudaMemcpyAsync(p->mask_dev.p, active, (size_t)P, cudaMemcpyHostToDevice, p->stream));
      p->pend_active = true;
    }
  } else {
    for (int i = 0; i < D; ++i) p->pend_ya[i] = y_a[i];
    p->pend_u = u ? u[0] : 0.0;
  }
  p->pend = true;
  p->pend_in_place = in_place;
  return 0;
}

int flush_pending(ccrs_problem* p) {
  if (!p->pend) return 0;
  p->pend = false;
  return launch_backsub_now(p, p->pend_in_place, false);
}

// K2 (or its cost-only variant K5). A deferred back-substitution that targets the same point is fused into the
// prologue. Single problem: {model decrease, cost} are reduced in-kernel and published under sequence number *seq_out.
int do_linearize(ccrs_problem* p, const double* intr, int which, bool cost_only, bool publish, double* seq_out) {
  p->spec.valid = false;   // any new linearisation makes a speculative reduction of the previous one stale
  if (p->pend && (p->pend_in_place ? which != 0 : which != 1)) { int st = flush_pending(p); if (st) return st; }
  LinParams prm{};
  prm.pb = p->dev();
  prm.which = which; prm.G = p->G; prm.FPW = p->FPW;
  prm.acc_to_blk = p->acc_to_blk.p;
  if (p->pend) {
    prm.backsub = p->pend_in_place ? 2 : 1;
    prm.elim = p->elim.p;
    prm.pose_scale = p->last_use_scale ? p->pose_scale.p : nullptr;
    prm.ya_dev = p->ya_dev.p; prm.u_dev = p->u_dev.p;
    prm.active = p->pend_active ? p->mask_dev.p : nullptr;
    prm.frame_md = p->frame_md.p;
    // single problem: K3 leaves the elimination record without the intrinsic Jacobi scaling, the step is D_a y_a
    for (int i = 0; i < p->D; ++i) prm.y_a[i] = p->pend_ya[i] * ((!p->batch && p->last_use_scale) ? p->h_scale[i] : 1.0);
    prm.u = p->pend_u;
    p->pend = false;
  } else if (p->batch) {
    CK(cudaMemsetAsync(p->frame_md.p, 0, (size_t)p->Fs * 8, p->stream));  // no step taken: model decrease 0
  }
  if (p->batch) {
    int st = upload_intr(p, intr);
    if (st) return st;
    prm.intr_dev = p->intr_dev.p;
  } else {
    if (p->one_focal) { prm.intr[0] = intr[0]; prm.intr[1] = intr[0]; for (int i = 1; i < p->D; ++i) prm.intr[i + 1] = intr[i]; }
    else for (int i = 0; i < p->D; ++i) prm.intr[i] = intr[i];
    prm.cta_part = p->cta_part.p;
    prm.ticket = p->tickets.p;
    prm.stat_dev = p->stat_out.p;
    p->seq = next_seq();
    prm.seq = p->seq;
    const bool peer = publish && use_peer(p, 2);
    prm.host_stat = (publish && (!p->comm || peer)) ? p->h_stat.p : nullptr;
    if (prm.host_stat) arm_payload(prm.host_stat, 2);
    if (peer) fill_peer(p, 1, &prm.px);
    if (seq_out) *seq_out = p->seq;
  }
#ifdef CCRS_K2_TIMING
  if (!p->k2_dbg.p) CK(p->k2_dbg.alloc((size_t)12 * std::max(p->n_lin_ctas, p->n_mma_ctas) * kLinWarps));
  prm.dbg = reinterpret_cast<long long*>(p->k2_dbg.p);
#endif
  CK(launch_k2(p, p->batch, cost_only, prm));
  p->launches++;
  if (!p->batch && publish && p->comm && !use_peer(p, 2)) {
    int st = exchange(p, p->stat_out.p, 2, p->h_stat.p, p->seq);   // h_stat[0..1] = sums, h_stat[2] = seq
    if (st) return st;
  }
  return 0;
}

// {model decrease, cost} of the last K2 launch. mode: see launch_trial_stats (batch path only).
int fetch_stats(ccrs_problem* p, int batch_mode, double seq, double* out /* [P][2] */) {
  const int P = p->n_problems;
  if (!p->batch) {
    // {md, cost} published by K2's last warp (or by k_sum_partials after the cross-rank exchange)
    int st = wait_payload(p, p->h_stat.p, 2);
    if (st) return st;
    out[0] = p->h_stat.p[0]; out[1] = p->h_stat.p[1];
    return 0;
  }
  CK(launch_trial_stats(p->dev(), p->NBLK - 1, batch_mode, p->frame_md.p, p->stat_out.p, p->stream));
  p->launches++;
  int st = exchange(p, p->stat_out.p, (size_t)P * 2, nullptr, 0.0);
  if (st) return st;
  CK(cudaMemcpyAsync(p->h_stat.p, p->stat_out.p, (size_t)P * 2 * 8, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  std::memcpy(out, p->h_stat.p, (size_t)P * 2 * 8);
  return 0;
}

// per-problem sum of a per-frame array [NV][Fs] -> out_dev[n_problems][NV]
int reduce_frames(ccrs_problem* p, const double* in, int NV, double* out_dev) {
  CK(launch_segreduce(in, NV, p->Fs, p->problem_frame_offsets.p, p->n_problems, out_dev, p->stream));
  p->launches++;
  return 0;
}

// payload area `a` of the single-problem K3 result in mapped host memory
volatile double* red_area_ptr(ccrs_problem* p, int a) { return p->h_red.p + (size_t)a * (p->NRED + 1); }

// single problem: enqueue K3 (+ the NCCL exchange when the peer path is off) publishing into payload area `area`
int launch_k3_single(ccrs_problem* p, int which, double u, int use_scale, double min_diag, double max_diag, int area) {
  Schur2Params prm{};
  prm.pb = p->dev();
  prm.ctl = nullptr;
  prm.which = which;
  prm.intr_scale = use_scale ? p->scale_dev.p : nullptr;
  prm.use_pose_scale = use_scale ? 1 : 0;
  prm.pose_scale = p->pose_scale.p;
  prm.min_diag = min_diag; prm.max_diag = max_diag;
  prm.no_pose = p->fixed_poses ? 1 : 0;
  prm.elim = p->elim.p;
  prm.partials = p->s2_part.p;
  p->last_use_scale = use_scale != 0;
  prm.u_val = u;
  prm.ticket = p->tickets.p + 1;
  prm.red_out = p->red_out.p;
  p->seq = next_seq();
  const bool peer = use_peer(p, (size_t)p->NRED);
  volatile double* host = red_area_ptr(p, area);
  prm.host_red = (p->comm && !peer) ? nullptr : host;
  arm_payload(host, p->NRED);
  if (peer) fill_peer(p, 0, &prm.px);
  CK(launch_schur2(p->D, prm, p->n_frames, true, p->stream));
  p->launches++;
  if (p->comm && !peer) { int st = exchange(p, p->red_out.p, (size_t)p->NRED, host, p->seq); if (st) return st; }
  return 0;
}

// unpack one problem's K3 result: packed upper S -> full row-major, then g_s | g_a | diag_a | sq_err
void unpack_reduced(const ccrs_problem* p, const volatile double* r, double* o) {
  const int D = p->D, NS = D * (D + 1) / 2;
  int e = 0;
  for (int a = 0; a < D; ++a)
    for (int b = a; b < D; ++b) { o[a * D + b] = r[e]; o[b * D + a] = r[e]; ++e; }
  for (int i = 0; i < 3 * D + 1; ++i) o[D * D + i] = r[NS + i];
}

int do_reduce(ccrs_problem* p, int which, const double* u, int use_scale, double min_diag, double max_diag, double* out) {
  const int D = p->D, P = p->n_problems;
  if (use_scale && !p->have_scale) return fail(CCRS_ERR_INVALID, "use_scale without ccrs_compute_scale/ccrs_set_intr_scale");
  int st = flush_pending(p);
  if (st) return st;
  if (p->batch) {
    SchurParams prm{};
    prm.pb = p->dev();
    prm.which = which;
    prm.intr_scale = use_scale ? p->scale_dev.p : nullptr;
    prm.pose_scale = use_scale ? p->pose_scale.p : nullptr;
    prm.min_diag = min_diag; prm.max_diag = max_diag;
    prm.no_pose = p->fixed_poses ? 1 : 0;
    prm.elim = p->elim.p;
    prm.frame_red = p->frame_red.p;
    p->last_use_scale = use_scale != 0;
    if (u) CK(cudaMemcpyAsync(p->u_dev.p, u, (size_t)P * 8, cudaMemcpyHostToDevice, p->stream));
    else CK(cudaMemsetAsync(p->u_dev.p, 0, (size_t)P * 8, p->stream));
    prm.u_dev = p->u_dev.p;
    CK(launch_schur(D, prm, p->stream));
    p->launches++;
    st = reduce_frames(p, p->frame_red.p, p->NRED, p->red_out.p);
    if (st) return st;
    st = exchange(p, p->red_out.p, (size_t)P * p->NRED, nullptr, 0.0);
    if (st) return st;
    CK(cudaMemcpyAsync(p->h_red.p, p->red_out.p, (size_t)P * p->NRED * 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    for (int q = 0; q < P; ++q) unpack_reduced(p, p->h_red.p + (size_t)q * p->NRED, out + (size_t)q * p->NOUT);
    return 0;
  }
  const double uv = u ? u[0] : 0.0;
  int area;
  const bool hit = p->spec.valid && which == 0 && p->spec.cur_after == p->cur_val && p->spec.u == uv &&
                   p->spec.use_scale == use_scale && p->spec.mn == min_diag && p->spec.mx == max_diag;
  p->spec.valid = false;
  g_trace.mark(5);
  if (hit) {
    area = p->spec.area;   // the reduction the controller asks for is the one already in flight
    g_spec_hits++;
  } else {
    area = (p->red_area ^= 1);
    st = launch_k3_single(p, which, uv, use_scale, min_diag, max_diag, area);
    if (st) return st;
  }
  p->last_mn = min_diag; p->last_mx = max_diag; p->have_last_reduce = true;
  g_trace.mark(0);
  st = wait_payload(p, red_area_ptr(p, area), p->NRED);
  if (st) return st;
  g_trace.mark(1);
  if (g_trace.on) g_trace.n++;
  unpack_reduced(p, red_area_ptr(p, area), out);
  return 0;
}

// ---- CUDA implementation of the controller's backend table ------------------------------------------------------
int be_linearize(void* ctx, const double* intr, int which) {
  return do_linearize((ccrs_problem*)ctx, intr, which, false, false, nullptr);
}
int be_compute_scale(void* ctx, int which, double* col_sq) { return ccrs_compute_scale((ccrs_problem*)ctx, which, col_sq); }
int be_set_intr_scale(void* ctx, const double* s) { return ccrs_set_intr_scale((ccrs_problem*)ctx, s); }
int be_reduce(void* ctx, int which, const double* u, int use_scale, double mn, double mx, double* out) {
  return do_reduce((ccrs_problem*)ctx, which, u, use_scale, mn, mx, out);
}
int be_backsub(void* ctx, const double* y_a, const double* u, const unsigned char* active, int in_place) {
  return defer_backsub((ccrs_problem*)ctx, y_a, u, active, in_place);   // fused into the next K2 launch
}
int be_trial_stats(void* ctx, const double* intr_trial, int speculative, double* out) {
  ccrs_problem* p = (ccrs_problem*)ctx;
  double seq = 0.0;
  g_trace.mark(2);
  const double u_step = p->pend_u;   // damping of the step being tried (do_linearize consumes the deferred back-substitution)
  int st = do_linearize(p, intr_trial, 1, !speculative, true, &seq);
  if (st) return st;
  if (speculative && !p->batch && g_spec_enabled && p->have_last_reduce && (!p->comm || use_peer(p, (size_t)p->NRED))) {
    ccrs_problem::SpecK3& sp = p->spec;
    sp.u = u_step * ccrs_rule::kLmMinAcceptFactor;   // ccrs_rule::lm_decide: u *= max(1/3, 1 - (2 rho - 1)^3) on accept
    sp.use_scale = p->last_use_scale ? 1 : 0; sp.mn = p->last_mn; sp.mx = p->last_mx;
    sp.cur_after = p->cur_val ^ 1;   // valid once the controller has accepted the trial point
    sp.area = (p->red_area ^= 1);
    st = launch_k3_single(p, 1, sp.u, sp.use_scale, sp.mn, sp.mx, sp.area);
    if (st) return st;
    sp.valid = true;
    g_spec_launched++;
  }
  g_trace.mark(3);
  st = fetch_stats(p, speculative ? 1 : 0, seq, out);
  g_trace.mark(4);
  return st;
}
int be_accept(void* ctx, const unsigned char* mask) { return ccrs_accept((ccrs_problem*)ctx, mask); }

ccrs_backend cuda_backend(ccrs_problem* p) {
  ccrs_backend be{};
  be.ctx = p; be.d = p->D; be.n_problems = p->n_problems;
  be.linearize = be_linearize; be.compute_scale = be_compute_scale; be.set_intr_scale = be_set_intr_scale;
  be.reduce = be_reduce; be.backsub = be_backsub; be.trial_stats = be_trial_stats; be.accept = be_accept;
  be.allreduce = nullptr;  // exchanged on the device (NCCL) inside reduce / compute_scale / trial_stats
  return be;
}


// Device-side phase trace of the device-driven loop (ccrs_loop_trace): globaltimer stamps carried by the records.
//   0 K2 (first warp past its wait -> last warp done) | 1 K2 done -> K3's last CTA past its wait | 2 K3 per-frame work
//   + CTA sums (last CTA) | 3 K3 tail: cross-CTA sum, exchange, controller rule | 4 record ready -> next K2 running
struct LoopTrace {
  bool on = false;
  long n = 0;
  double acc[5] = {0, 0, 0, 0, 0};
  double wake = 0.0;              // record ready -> first K2 warp past its dependency wait (part of acc[4])
  double fine[7] = {0, 0, 0, 0, 0, 0, 0};   // last CTA of K3: head | block load | per-frame compute | CTA sum + partial store | partial sum (+ exchange) | pre-rule | rule
  double prev_end = -1.0;
  void add(const double* r) {
    const double wrap = 1099511627776.0;   // 2^40 ns
    auto d = [&](double a, double b) { double x = b - a; if (x < -wrap / 2) x += wrap; return x; };
    if (r[REC_T_K2_END] <= 0.0) { prev_end = r[REC_T_END]; return; }
    if (prev_end >= 0.0 && (int)r[REC_DECIDED]) {
      acc[0] += d(r[REC_T_K2_BEGIN], r[REC_T_K2_END]);
      acc[1] += d(r[REC_T_K2_END], r[REC_T_K3_BEGIN]);
      acc[2] += d(r[REC_T_K3_BEGIN], r[REC_T_TAIL]);
      acc[3] += d(r[REC_T_TAIL], r[REC_T_END]);
      acc[4] += d(prev_end, r[REC_T_K2_BEGIN]);
      wake += d(prev_end, r[REC_T_K2_WAKE]);
      fine[0] += d(r[REC_T_K3_BEGIN], r[REC_T_HEAD]); fine[1] += d(r[REC_T_HEAD], r[REC_T_LOAD]);
      fine[2] += d(r[REC_T_LOAD], r[REC_T_COMP]); fine[3] += d(r[REC_T_COMP], r[REC_T_TAIL]);
      fine[4] += d(r[REC_T_TAIL], r[REC_T_SUMMED]); fine[5] += d(r[REC_T_SUMMED], r[REC_T_RULE]);
      fine[6] += d(r[REC_T_RULE], r[REC_T_END]);
      ++n;
    }
    prev_end = r[REC_T_END];
  }
};
LoopTrace g_loop_trace;

// ---- device-driven loop (single problem) -----------------------------------------------------------------------------
// The host enqueues K3, K2, K3, K2, ... a bounded distance ahead of the records it has consumed; every decision of the
// iteration is taken by the last CTA of K3 with the rule of ccrs_rule.h (LoopCtl, ccrs_kernels.cuh). The host keeps
// what north_star keeps on the host — it re-runs the same rule on every published reduced system and checks the
// device's accept / reject decision, damping, solve and trial point bit for bit (audit), fills the summary, and stops
// the loop — but it is no longer on the critical path between two kernels.
const bool g_device_loop = [] { const char* e = getenv("CCRS_DEVICE_LOOP"); return !(e && atoi(e) == 0); }();
const int g_loop_ahead = [] { const char* e = getenv("CCRS_LOOP_AHEAD"); const int v = e ? atoi(e) : 2; return v >= 1 && v <= 8 ? v : 2; }();
std::atomic<long> g_loop_audited{0};

bool device_loop_ok(const ccrs_problem* p, bool lm, const ccrs_options& opt) {
  if (!g_device_loop || p->batch) return false;
  if (p->comm && !use_peer(p, (size_t)p->NRED + 2)) return false;   // NCCL fallback: host-driven
  if (lm && !opt.speculative) return false;                          // classical K5 sequence: host-driven
  return true;
}

struct DeviceLoop {
  ccrs_problem* p;
  bool lm;
  int D;
  long enq = 0, got = 0;          // K3 slots enqueued / records consumed in this loop
  double intr_host[9];            // the host's copy of the current intrinsics (audit of the GN update)
  unsigned char fixed[16];
  bool has_fixed = false, has_bounds = false;
  double lo[9], hi[9];
  ccrs_options opt;
};

volatile double* rec_slot(ccrs_problem* p, long r) { return p->h_rec.p + (size_t)(r % kRecSlots) * kRecStride; }

int loop_begin(DeviceLoop& L, ccrs_problem* p, bool lm, const double* intr, const double* lo, const double* hi,
               const unsigned char* fixed, const ccrs_options& opt) {
  int st = flush_pending(p);
  if (st) return st;
  p->spec.valid = false; p->have_last_reduce = false;
  L.p = p; L.lm = lm; L.D = p->D; L.opt = opt; L.enq = L.got = 0;
  LoopCtl* c = reinterpret_cast<LoopCtl*>(p->h_ctl.p);
  std::memset(c, 0, sizeof(LoopCtl));
  c->mode = lm ? 1 : 0; c->mode_k2 = c->mode; c->D = p->D; c->max_iteration = opt.max_iteration; c->fixed_mode = opt.fixed_mode;
  c->has_bounds = (lo && hi) ? 1 : 0; c->has_fixed = fixed ? 1 : 0;
  c->min_abs = opt.min_abs_decrease; c->min_rel = opt.min_rel_decrease; c->min_error = opt.min_error;
  c->min_diag = opt.lm_min_diag; c->max_diag = opt.lm_max_diag; c->block_huber = lm ? 0.0 : opt.block_huber_delta;
  L.has_bounds = c->has_bounds != 0; L.has_fixed = c->has_fixed != 0;
  for (int i = 0; i < p->D; ++i) {
    c->lo[i] = L.lo[i] = c->has_bounds ? lo[i] : 0.0; c->hi[i] = L.hi[i] = c->has_bounds ? hi[i] : 0.0;
    c->fixed[i] = L.fixed[i] = fixed ? fixed[i] : 0;
    c->intr[i] = c->trial[i] = L.intr_host[i] = intr[i];
    c->scale[i] = 1.0;
  }
  c->phase = PH_LIN0; c->cur = p->cur_val; c->first = lm ? 1 : 0;
  c->u = 1.0 / opt.lm_initial_radius; c->v = ccrs_rule::kLmRejectFactor0;
  CK(cudaMemcpyAsync(p->ctl_dev.p, c, sizeof(LoopCtl), cudaMemcpyHostToDevice, p->stream));
  // records are consumed in order and a slot is re-armed when its record has been read; in flight <= ahead + 1 << ring
  arm_payload(p->h_rec.p, kRecSlots * kRecStride);
  p->last_use_scale = lm;
  return 0;
}

int loop_launch_k2(DeviceLoop& L) {
  ccrs_problem* p = L.p;
  LinParams prm{};
  prm.pb = p->dev();
  prm.ctl = reinterpret_cast<LoopCtl*>(p->ctl_dev.p);
  prm.G = p->G; prm.FPW = p->FPW;
  prm.acc_to_blk = p->acc_to_blk.p;
  prm.elim = p->elim.p;
  prm.pose_scale = L.lm ? p->pose_scale.p : nullptr;
  prm.frame_md = p->frame_md.p;
  prm.cta_part = p->cta_part.p;
  prm.ticket = p->tickets.p;
  prm.stat_dev = p->stat_out.p;
  CK(launch_k2(p, false, false, prm));
  p->launches++;
  return 0;
}

int loop_launch_k3(DeviceLoop& L) {
  ccrs_problem* p = L.p;
  Schur2Params prm{};
  prm.pb = p->dev();
  prm.ctl = reinterpret_cast<LoopCtl*>(p->ctl_dev.p);
  prm.pose_scale = p->pose_scale.p;
  prm.min_diag = L.opt.lm_min_diag; prm.max_diag = L.opt.lm_max_diag;
  prm.no_pose = p->fixed_poses ? 1 : 0;
  prm.elim = p->elim.p;
  prm.partials = p->s2_part.p;
  prm.ticket = p->tickets.p + 1;
  prm.red_out = p->red_out.p;
  prm.rec = p->h_rec.p;
  if (p->comm && p->world > 1) {
    for (int r = 0; r < kXchgMaxRanks; ++r) prm.px.peer[r] = g_peer.peer[r];
    prm.px.world = p->world; prm.px.rank = p->rank;
    prm.px.off = (int)((size_t)2 * 2 * kXchgMaxRanks * kXchgMaxVals);   // area 2, parity added on the device
    prm.xchg_count = g_peer.dev_count();
  }
  CK(launch_schur2(p->D, prm, p->n_frames, true, p->stream));
  p->launches++;
  L.enq++;
  return 0;
}

bool same_bits(double a, double b) { return std::memcmp(&a, &b, 8) == 0; }

// Wait for the next record, audit it against the host rule, fold it into the summary. *done = 1 on the final record.
int loop_consume(DeviceLoop& L, ccrs_summary* sum, double* err_hist, int* done, double* intr_out) {
  ccrs_problem* p = L.p;
  const int D = L.D;
  volatile double* slot = rec_slot(p, L.got);
  int st = wait_payload(p, slot, kRecStride);
  if (st) return st;
  double r[kRecStride];
  for (int i = 0; i < kRecStride; ++i) r[i] = slot[i];
  arm_payload(slot, kRecStride);
  L.got++; p->loop_records++;
  static const bool dbg = getenv("CCRS_LOOP_DEBUG") != nullptr;
  if (dbg)
    fprintf(stderr, "[loop] rec %ld seq %.0f phase %.0f it %.0f iters %.0f acc %.0f rho %.6g u %.6g err %.9g solved %.0f status %.0f stop %.0f cur %.0f\n",
            L.got, r[REC_SEQ], r[REC_PHASE], r[REC_IT], r[REC_ITERATIONS], r[REC_ACCEPTED], r[REC_RHO], r[REC_U], r[REC_CUR_ERR], r[REC_SOLVED],
            r[REC_STATUS], r[REC_STOP], r[REC_CUR]);
  if ((long)r[REC_SEQ] != L.got) return fail(CCRS_ERR_CUDA, "loop record out of sequence (%ld, expected %ld)", (long)r[REC_SEQ], L.got);
  // ---- audit: the device's decisions reproduced with the host build of the same rule ---------------------------------
  using namespace ccrs_rule;
  if (L.lm && r[REC_DECIDED] != 0.0) {
    LmState stt{r[REC_U_BEFORE], r[REC_V_BEFORE], r[REC_CUR_ERR_BEFORE]};
    double rho;
    const int acc = lm_decide(r[REC_SQ_CUR_BEFORE], r[REC_SQ_NEW], r[REC_MD_A_BEFORE] + r[REC_MD_POSE], &stt, &rho);
    if (acc != (int)r[REC_ACCEPTED] || !same_bits(stt.u, r[REC_U]) || !same_bits(stt.v, r[REC_V]) || (!same_bits(rho, r[REC_RHO]) && !(is_nan(rho) && is_nan(r[REC_RHO]))))
      return fail(CCRS_ERR_NUMERIC, "device LM decision differs from the host rule at record %ld (accepted %d vs %d, u %.17g vs %.17g)",
                  L.got, (int)r[REC_ACCEPTED], acc, r[REC_U], stt.u);
    if (acc) for (int i = 0; i < D; ++i) L.intr_host[i] = r[REC_INTR + i];
  }
  if (r[REC_SOLVED] != 0.0) {
    double y[kMaxD], dx[kMaxD], trial[kMaxD], md_a = 0.0;
    const Reduced rv = view(r + REC_OUT, D);
    const int sst = solve_intrinsics(rv, D, r[REC_U_SOLVE], L.opt.lm_min_diag, L.opt.lm_max_diag, L.has_fixed ? L.fixed : nullptr,
                                     L.opt.fixed_mode, y, &md_a);
    bool ok = sst == 0 && same_bits(md_a, r[REC_MD_A]);
    if (ok) {
      for (int i = 0; i < D; ++i) dx[i] = r[REC_SCALE + i] * y[i];
      update_intr(D, L.intr_host, dx, L.has_bounds ? L.lo : nullptr, L.has_bounds ? L.hi : nullptr, L.has_fixed ? L.fixed : nullptr, trial);
      for (int i = 0; i < D; ++i) ok = ok && same_bits(y[i], r[REC_Y + i]) && same_bits(trial[i], r[REC_TRIAL + i]);
    }
    if (!ok) return fail(CCRS_ERR_NUMERIC, "device intrinsic solve differs from the host rule at record %ld", L.got);
    if (!L.lm) for (int i = 0; i < D; ++i) L.intr_host[i] = trial[i];
    g_loop_audited++;
  }
  // ---- summary -----------------------------------------------------------------------------------------------------------
  sum->iterations = (int)r[REC_ITERATIONS];
  sum->final_error = r[REC_FINAL_ERR];
  sum->n_accepted = (int)r[REC_N_ACC]; sum->n_rejected = (int)r[REC_N_REJ];
  if (err_hist && r[REC_HIST_IDX] >= 0.0 && (int)r[REC_HIST_IDX] < L.opt.max_iteration) err_hist[(int)r[REC_HIST_IDX]] = r[REC_HIST_VAL];
  *done = (int)r[REC_PHASE] == PH_DONE;
  if (*done) {
    sum->status = (int)r[REC_STATUS];
    sum->stop_reason = (int)r[REC_STOP];
    p->cur_val = (int)r[REC_CUR];
    for (int i = 0; i < D; ++i) { intr_out[i] = r[REC_INTR + i]; p->h_scale[i] = r[REC_SCALE + i]; }
  }
  if (g_loop_trace.on) g_loop_trace.add(r);
  return 0;
}

int run_device_loop(ccrs_problem* p, bool lm, double* intr, const double* lo, const double* hi, const unsigned char* fixed,
                    const ccrs_options& opt, ccrs_summary* sum, double* err_hist) {
  std::memset(sum, 0, sizeof(*sum));
  if (opt.max_iteration <= 0) return 0;
  DeviceLoop L;
  int st = loop_begin(L, p, lm, intr, lo, hi, fixed, opt);
  if (st) return st;
  st = loop_launch_k2(L);                       // linearise the start point
  int done = 0;
  while (!st && !done) {
    while (!st && L.enq - L.got < g_loop_ahead) { st = loop_launch_k3(L); if (!st) st = loop_launch_k2(L); }
    if (!st) st = loop_consume(L, sum, err_hist, &done, intr);
  }
  if (st) { sum->status = st; cudaStreamSynchronize(p->stream); return st; }
  return sum->status;
}


// ---- device-driven loop for a batch of independent problems ----------------------------------------------------------
// Per iteration the host enqueues K3, k_batch_solve, K2, k_batch_decide (GN: K2, K3, k_batch_solve) and reads one
// number — the count of still-active problems — from a mapped status ring; every per-problem decision (ccrs_rule.h)
// is taken by the two controller kernels, which write straight into the arrays K2 / K3 read.
bool batch_loop_ok(const ccrs_problem* p, bool lm, const ccrs_options& opt) {
  if (!g_device_loop || !p->batch || p->comm) return false;
  if (lm && !opt.speculative) return false;
  if (!lm && opt.block_huber_delta > 0.0) return false;
  return true;
}

int batch_launch_k2(ccrs_problem* p, int which, int backsub, bool use_pose_scale) {
  LinParams prm{};
  prm.pb = p->dev();
  prm.which = which; prm.G = p->G; prm.FPW = p->FPW;
  prm.acc_to_blk = p->acc_to_blk.p;
  prm.backsub = backsub;
  prm.elim = p->elim.p;
  prm.pose_scale = use_pose_scale ? p->pose_scale.p : nullptr;
  prm.ya_dev = p->ya_dev.p; prm.u_dev = p->u_dev.p;
  prm.active = p->mask_dev.p;
  prm.frame_md = p->frame_md.p;
  prm.intr_dev = p->intr_dev.p;
  CK(launch_k2(p, true, false, prm));
  p->launches++;
  return 0;
}

int batch_launch_k3(ccrs_problem* p, bool lm, const ccrs_options& opt) {
  SchurParams prm{};
  prm.pb = p->dev();
  prm.which = 0;
  prm.intr_scale = lm ? p->scale_dev.p : nullptr;
  prm.pose_scale = lm ? p->pose_scale.p : nullptr;
  prm.min_diag = opt.lm_min_diag; prm.max_diag = opt.lm_max_diag;
  prm.no_pose = p->fixed_poses ? 1 : 0;
  prm.active = p->mask_dev.p;
  prm.elim = p->elim.p;
  prm.frame_red = p->frame_red.p;
  prm.u_dev = p->u_dev.p;
  p->last_use_scale = lm;
  CK(launch_schur(p->D, prm, p->stream));
  p->launches++;
  return 0;
}

int run_batch_loop(ccrs_problem* p, bool lm, double* intr, const double* lo, const double* hi, const unsigned char* fixed,
                   const ccrs_options& opt, ccrs_summary* sum, double* err_hist) {
  std::memset(sum, 0, sizeof(*sum));
  const int P = p->n_problems, D = p->D;
  if (opt.max_iteration <= 0) return 0;
  int st = flush_pending(p);
  if (st) return st;
  p->spec.valid = false;
  constexpr size_t kCtlWords = sizeof(BatchCtl) / 8;
  if (!p->bctl_dev.p) {
    CK(p->bctl_dev.alloc((size_t)P * kCtlWords)); CK(p->h_bctl.alloc((size_t)P * kCtlWords));
    CK(p->h_bstatus.alloc((size_t)kRecSlots * 2));
  }
  const bool want_hist = err_hist && opt.max_iteration <= (1 << 20);
  if (want_hist) { p->hist_dev.release(); CK(p->hist_dev.alloc((size_t)opt.max_iteration)); }
  // ---- start: linearise the start point; LM: Jacobi scaling from that linearisation (one-time, host-assisted)
  std::vector<unsigned char> ones((size_t)P, 1);
  CK(cudaMemcpyAsync(p->mask_dev.p, ones.data(), (size_t)P, cudaMemcpyHostToDevice, p->stream));
  CK(cudaMemsetAsync(p->u_dev.p, 0, (size_t)P * 8, p->stream));
  st = upload_intr(p, intr);
  if (!st) st = batch_launch_k2(p, 0, 0, false);
  if (st) return st;
  std::vector<double> scale((size_t)P * D, 1.0);
  if (lm) {
    std::vector<double> colsq((size_t)P * D);
    st = ccrs_compute_scale(p, 0, colsq.data());
    if (st) return st;
    for (size_t i = 0; i < scale.size(); ++i) scale[i] = 1.0 / (1.0 + std::sqrt(colsq[i]));
    st = ccrs_set_intr_scale(p, scale.data());
    if (st) return st;
  }
  BatchCtl* hc = reinterpret_cast<BatchCtl*>(p->h_bctl.p);
  std::vector<double> u0((size_t)P, lm ? 1.0 / opt.lm_initial_radius : 0.0);
  for (int q = 0; q < P; ++q) {
    BatchCtl& c = hc[q];
    std::memset(&c, 0, sizeof(c));
    c.u = u0[q]; c.v = ccrs_rule::kLmRejectFactor0; c.active = 1;
    for (int i = 0; i < D; ++i) { c.intr[i] = c.trial[i] = intr[(size_t)q * D + i]; c.scale[i] = scale[(size_t)q * D + i]; }
  }
  CK(cudaMemcpyAsync(p->bctl_dev.p, hc, (size_t)P * sizeof(BatchCtl), cudaMemcpyHostToDevice, p->stream));
  CK(cudaMemcpyAsync(p->u_dev.p, u0.data(), (size_t)P * 8, cudaMemcpyHostToDevice, p->stream));
  CK(cudaMemsetAsync(p->tickets.p + 2, 0, 2 * sizeof(unsigned int), p->stream));
  CK(cudaStreamSynchronize(p->stream));   // the staging vectors above are pageable / reused
  arm_payload(p->h_bstatus.p, kRecSlots * 2);
  BatchRuleParams rp{};
  rp.pb = p->dev();
  rp.ctl = reinterpret_cast<BatchCtl*>(p->bctl_dev.p);
  rp.lm = lm ? 1 : 0; rp.D = D; rp.NRED = p->NRED; rp.max_iteration = opt.max_iteration; rp.fixed_mode = opt.fixed_mode;
  rp.has_bounds = (lo && hi) ? 1 : 0; rp.has_fixed = fixed ? 1 : 0; rp.rr_idx = p->NBLK - 1;
  rp.min_abs = opt.min_abs_decrease; rp.min_rel = opt.min_rel_decrease; rp.min_error = opt.min_error;
  rp.min_diag = opt.lm_min_diag; rp.max_diag = opt.lm_max_diag;
  for (int i = 0; i < D; ++i) { rp.lo[i] = rp.has_bounds ? lo[i] : 0.0; rp.hi[i] = rp.has_bounds ? hi[i] : 0.0; rp.fixed[i] = fixed ? fixed[i] : 0; }
  rp.frame_red = p->frame_red.p; rp.frame_md = p->frame_md.p;
  rp.intr_dev = p->intr_dev.p; rp.ya_dev = p->ya_dev.p; rp.u_dev = p->u_dev.p; rp.cur = p->cur.p; rp.active = p->mask_dev.p;
  rp.n_active = p->tickets.p + 2; rp.ticket = p->tickets.p + 3;
  rp.host_status = p->h_bstatus.p;
  rp.err_hist0 = want_hist ? p->hist_dev.p : nullptr;
  // ---- the loop: enqueue ahead, read one status word pair per iteration
  long enq = 0, got = 0;
  bool done = false;
  auto enqueue = [&]() -> int {
    int s2 = 0;
    rp.seq = (int)(enq + 1);
    if (lm) {
      s2 = batch_launch_k3(p, true, opt);
      if (!s2) { CK(launch_batch_solve(rp, p->stream)); p->launches++; s2 = batch_launch_k2(p, 1, 1, true); }
      if (!s2) { CK(launch_batch_decide(rp, p->stream)); p->launches++; }
    } else {
      if (enq > 0) s2 = batch_launch_k2(p, 0, 2, false);   // apply the step in place and linearise there
      if (!s2) s2 = batch_launch_k3(p, false, opt);
      if (!s2) { CK(launch_batch_solve(rp, p->stream)); p->launches++; }
    }
    ++enq;
    return s2;
  };
  while (!st && !done) {
    while (!st && enq - got < g_loop_ahead && enq <= opt.max_iteration) st = enqueue();
    if (st) break;
    volatile double* slot = p->h_bstatus.p + (size_t)(got % kRecSlots) * 2;
    st = wait_payload(p, slot, 2);
    if (st) break;
    const double n_active = slot[1];
    arm_payload(slot, 2);
    ++got;
    if (n_active == 0.0 || got > opt.max_iteration) done = true;
  }
  if (st) { cudaStreamSynchronize(p->stream); sum->status = st; return st; }
  CK(cudaMemcpyAsync(hc, p->bctl_dev.p, (size_t)P * sizeof(BatchCtl), cudaMemcpyDeviceToHost, p->stream));
  std::vector<double> hist;
  if (want_hist) { hist.resize((size_t)opt.max_iteration); CK(cudaMemcpyAsync(hist.data(), p->hist_dev.p, hist.size() * 8, cudaMemcpyDeviceToHost, p->stream)); }
  CK(cudaStreamSynchronize(p->stream));
  int worst = 0, iters = 0;
  for (int q = 0; q < P; ++q) {
    const BatchCtl& c = hc[q];
    for (int i = 0; i < D; ++i) intr[(size_t)q * D + i] = c.intr[i];
    if (c.status != 0) worst = c.status;
    iters = std::max(iters, c.iterations);
  }
  sum->iterations = iters; sum->status = worst; sum->stop_reason = hc[0].stop; sum->final_error = hc[0].final_err;
  sum->n_accepted = hc[0].n_acc; sum->n_rejected = hc[0].n_rej;
  if (want_hist) for (int i = 0; i < hc[0].iterations && i < opt.max_iteration; ++i) err_hist[i] = hist[i];
  return worst;
}

}  // namespace

extern "C" {

int ccrs_model_nparams(int model) {
  int D = 0;
  if (model_dims(model, 0, &D, nullptr, nullptr, nullptr) != 0) return -1;
  return D;
}

const char* ccrs_last_error(void) { return g_err; }

int ccrs_problem_create(ccrs_problem** out, int model, int width, int height, int xy_same_focal, int n_frames,
                        const int32_t* frame_offsets, const double* x, const double* y, const double* z,
                        const double* u, const double* v, double huber_delta, int device_id) {
  return create_common(out, model, width, height, xy_same_focal, 1, nullptr, n_frames, frame_offsets, x, y, z, u, v,
                       huber_delta, device_id, false);
}

int ccrs_problem_create_f32(ccrs_problem** out, int model, int width, int height, int xy_same_focal, int n_frames,
                            const int32_t* frame_offsets, const float* x, const float* y, const float* z,
                            const float* u, const float* v, double huber_delta, int device_id) {
  return create_common(out, model, width, height, xy_same_focal, 1, nullptr, n_frames, frame_offsets, x, y, z, u, v,
                       huber_delta, device_id, false, true);
}

int ccrs_problem_create_board_f32(ccrs_problem** out, int model, int width, int height, int xy_same_focal, int n_frames,
                                  const int32_t* frame_offsets, const int32_t* corner_id, const float* u, const float* v,
                                  const float* board_xyz, int n_board, double huber_delta, int device_id) {
  return create_common(out, model, width, height, xy_same_focal, 1, nullptr, n_frames, frame_offsets, nullptr, nullptr, nullptr,
                       u, v, huber_delta, device_id, false, true, corner_id, board_xyz, n_board);
}

int ccrs_problem_update_observations(ccrs_problem* p, const int32_t* frame_offsets, const int32_t* corner_id, const void* x,
                                     const void* y, const void* z, const void* u, const void* v) {
  if (!p || !frame_offsets || !u || !v) return fail(CCRS_ERR_INVALID, "null");
  if (p->batch) return fail(CCRS_ERR_INVALID, "update_observations: single-problem handles only");
  const bool board = p->corner_id.p != nullptr;
  if (board ? !corner_id : (!x || !y || !z)) return fail(CCRS_ERR_INVALID, "update_observations: the handle was created %s", board ? "in board format: corner ids required" : "from x, y, z arrays: they are required");
  const int F = p->n_frames;
  if (frame_offsets[0] != 0 || (int64_t)frame_offsets[F] != p->n_obs)
    return fail(CCRS_ERR_INVALID, "update_observations: the handle holds %lld observations in %d frames; the new detections must have the same totals", (long long)p->n_obs, F);
  for (int f = 0; f < F; ++f)
    if (frame_offsets[f + 1] <= frame_offsets[f]) return fail(CCRS_ERR_INVALID, "frame %d has no observations / offsets not monotone", f);
  CK(cudaSetDevice(p->device));
  cudaStream_t s = p->stream;
  const size_t N = (size_t)p->n_obs, esz = p->f32 ? 4 : 8;
  // any step still deferred into the next linearisation belongs to the old observations
  p->pend = false; p->cur_val = 0; p->spec.valid = false; p->have_last_reduce = false; p->have_scale = false; p->have_obs_frame = false;
  if (board) {
    CK(cudaMemcpyAsync(p->corner_id.p, corner_id, N * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    p->h_colsq.p[0] = 0.0;
    CK(launch_expand_board(p->corner_id.p, p->board_dev.p, p->n_board, (int64_t)N, reinterpret_cast<float*>(p->x.p),
                           reinterpret_cast<float*>(p->y.p), reinterpret_cast<float*>(p->z.p), p->h_colsq.p, s));
    p->launches++;
  } else {
    CK(cudaMemcpyAsync(p->x.p, x, N * esz, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->y.p, y, N * esz, cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync(p->z.p, z, N * esz, cudaMemcpyHostToDevice, s));
  }
  CK(cudaMemcpyAsync(p->u.p, u, N * esz, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(p->v.p, v, N * esz, cudaMemcpyHostToDevice, s));
  p->h_frame_offsets.assign(frame_offsets, frame_offsets + F + 1);
  CK(cudaMemcpyAsync(p->frame_offsets.p, p->h_frame_offsets.data(), (size_t)(F + 1) * 4, cudaMemcpyHostToDevice, s));
  CK(cudaStreamSynchronize(s));   // the caller's buffers are free again
  if (board && p->h_colsq.p[0] != 0.0) return fail(CCRS_ERR_INVALID, "a corner id lies outside the board table (%d corners)", p->n_board);
  return 0;
}

int ccrs_batch_create(ccrs_problem** out, int model, int width, int height, int xy_same_focal, int n_problems,
                      const int32_t* problem_frame_offsets, int n_frames, const int32_t* frame_offsets, const double* x,
                      const double* y, const double* z, const double* u, const double* v, double huber_delta,
                      int device_id) {
  if (!problem_frame_offsets || problem_frame_offsets[0] != 0 || problem_frame_offsets[n_problems] != n_frames)
    return fail(CCRS_ERR_INVALID, "problem_frame_offsets must span [0, n_frames]");
  return create_common(out, model, width, height, xy_same_focal, n_problems, problem_frame_offsets, n_frames,
                       frame_offsets, x, y, z, u, v, huber_delta, device_id, true);
}

int ccrs_problem_destroy(ccrs_problem* p) {
  if (!p) return 0;
  cudaSetDevice(p->device);
  if (p->stream) cudaStreamSynchronize(p->stream);
  p->x.release(); p->y.release(); p->z.release(); p->u.release(); p->v.release();
  p->corner_id.release(); p->board_dev.release();
  p->frame_offsets.release(); p->frame_problem.release(); p->problem_frame_offsets.release(); p->obs_frame.release();
  p->cur.release(); p->acc_to_blk.release(); p->tickets.release();
  for (int i = 0; i < 2; ++i) { p->poses[i].release(); p->blocks[i].release(); p->frame_cost[i].release(); }
  p->elim.release(); p->frame_red.release(); p->pose_scale.release(); p->frame_md.release(); p->cta_part.release();
  p->frame_stat.release(); p->chunk_cnt.release(); p->rdv.release();
  p->red_out.release(); p->stat_out.release(); p->gather.release(); p->intr_dev.release(); p->ya_dev.release();
  p->u_dev.release(); p->scale_dev.release(); p->l2_flush.release(); p->mask_dev.release();
  p->h_red.release(); p->h_stat.release(); p->h_colsq.release();
  p->ctl_dev.release(); p->h_rec.release(); p->h_ctl.release(); p->s2_part.release();
  p->bctl_dev.release(); p->hist_dev.release(); p->h_bctl.release(); p->h_bstatus.release();
  if (p->stream) put_stream(p->device, p->stream);
  delete p;
  return 0;
}

int ccrs_release_cached_memory(void) {
  for (int d = 0; d < 64; ++d) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || d >= n) break;
    cudaSetDevice(d);
    dev_pool(d).trim();
  }
  host_pool().trim();
  std::lock_guard<std::mutex> g(g_stream_mu);
  for (auto& s : g_streams) { cudaSetDevice(s.first); cudaStreamDestroy(s.second); }
  g_streams.clear();
  return 0;
}

int ccrs_problem_dim(const ccrs_problem* p) { return p ? p->D : -1; }
int ccrs_problem_nblk(const ccrs_problem* p) { return p ? p->NBLK : -1; }
int ccrs_problem_n_frames(const ccrs_problem* p) { return p ? p->n_frames : -1; }
int64_t ccrs_problem_n_obs(const ccrs_problem* p) { return p ? p->n_obs : -1; }
int ccrs_problem_n_problems(const ccrs_problem* p) { return p ? p->n_problems : -1; }
int64_t ccrs_launch_count(const ccrs_problem* p) { return p ? p->launches : -1; }

int ccrs_step_trace(int enable, double* avg_us /* [6] or NULL */, int64_t* n_iterations) {
  if (avg_us) for (int i = 0; i < 6; ++i) avg_us[i] = g_trace.n ? g_trace.acc[i] / g_trace.n : 0.0;
  if (n_iterations) *n_iterations = g_trace.n;
  g_trace = StepTrace{};
  g_trace.on = enable != 0;
  return 0;
}

#ifdef CCRS_K2_TIMING
extern "C" int ccrs_debug_k3_timing(ccrs_problem* p, long long* out, int cap_warps) {
  if (!p || !p->k3_dbg.p) return -1;
  const int nw = (p->n_frames + 31) / 32;
  cudaStreamSynchronize(p->stream);
  cudaMemcpy(out, p->k3_dbg.p, (size_t)8 * std::min(nw, cap_warps) * sizeof(long long), cudaMemcpyDeviceToHost);
  return nw;
}
// debug builds only (make timing): per-warp phase clocks of the last K2 launch, [n_warps][10] int64
extern "C" int ccrs_debug_k2_timing(ccrs_problem* p, long long* out, int cap_warps) {
  if (!p || !p->k2_dbg.p) return -1;
  const int nw = (p->use_mma ? p->n_mma_ctas : p->n_lin_ctas) * kLinWarps;
  cudaStreamSynchronize(p->stream);
  cudaMemcpy(out, p->k2_dbg.p, (size_t)12 * std::min(nw, cap_warps) * sizeof(long long), cudaMemcpyDeviceToHost);
  return nw;
}
#endif

int ccrs_set_poses(ccrs_problem* p, const double* poses) {
  if (!p || !poses) return fail(CCRS_ERR_INVALID, "null");
  CK(cudaSetDevice(p->device));
  // resetting the poses also resets the buffer selector and drops any deferred step
  p->pend = false;
  p->cur_val = 0;
  p->spec.valid = false; p->have_last_reduce = false;
  if (p->batch) CK(cudaMemsetAsync(p->cur.p, 0, (size_t)p->n_problems * sizeof(int32_t), p->stream));
  CK(cudaMemcpyAsync(p->poses[0].p, poses, (size_t)p->n_frames * 6 * 8, cudaMemcpyHostToDevice, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  return 0;
}

int ccrs_get_poses(ccrs_problem* p, double* poses) {
  if (!p || !poses) return fail(CCRS_ERR_INVALID, "null");
  CK(cudaSetDevice(p->device));
  int st = flush_pending(p);
  if (st) return st;
  if (!p->batch) {
    CK(cudaMemcpyAsync(poses, p->poses[p->cur_val].p, (size_t)p->n_frames * 48, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    return 0;
  }
  std::vector<int32_t> cur(p->n_problems);
  std::vector<double> b0((size_t)p->n_frames * 6), b1;
  CK(cudaMemcpyAsync(cur.data(), p->cur.p, cur.size() * 4, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaMemcpyAsync(b0.data(), p->poses[0].p, b0.size() * 8, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  bool any1 = false;
  for (int c : cur) any1 |= (c != 0);
  if (any1) {
    b1.resize(b0.size());
    CK(cudaMemcpyAsync(b1.data(), p->poses[1].p, b1.size() * 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
  }
  for (int q = 0; q < p->n_problems; ++q) {
    const std::vector<double>& src = cur[q] ? b1 : b0;
    const size_t a = (size_t)p->h_problem_frame_offsets[q] * 6, b = (size_t)p->h_problem_frame_offsets[q + 1] * 6;
    std::memcpy(poses + a, src.data() + a, (b - a) * 8);
  }
  return 0;
}

// frame index of every observation (thread-per-observation kernels K1 / K6), built on first use
static int ensure_obs_frame(ccrs_problem* p) {
  if (p->have_obs_frame) return 0;
  const size_t N = (size_t)p->n_obs;
  std::vector<int32_t> of(N);
  for (int f = 0; f < p->n_frames; ++f)
    for (int k = p->h_frame_offsets[f]; k < p->h_frame_offsets[f + 1]; ++k) of[k] = f;
  CK(p->obs_frame.alloc(N));
  CK(cudaMemcpyAsync(p->obs_frame.p, of.data(), N * 4, cudaMemcpyHostToDevice, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  p->have_obs_frame = true;
  return 0;
}

int ccrs_eval_rj(ccrs_problem* p, const double* intr, const double* poses, int apply_loss, double* r, double* J) {
  if (!p || !intr || !r) return fail(CCRS_ERR_INVALID, "null");
  if (p->batch) return fail(CCRS_ERR_INVALID, "ccrs_eval_rj is for single-problem handles");
  CK(cudaSetDevice(p->device));
  int st = flush_pending(p);
  if (st) return st;
  const size_t N = (size_t)p->n_obs;
  st = ensure_obs_frame(p);
  if (st) return st;
  const int n = p->D + 6;
  DevBuf<double> d_r, d_J, d_pose;
  CK(d_r.alloc(2 * N));
  if (J) CK(d_J.alloc(2 * N * n));
  st = upload_intr(p, intr);
  if (st) return st;
  const double* pose_ptr = p->poses[p->cur_val].p;
  if (poses) {
    CK(d_pose.alloc((size_t)p->n_frames * 6));
    CK(cudaMemcpyAsync(d_pose.p, poses, (size_t)p->n_frames * 48, cudaMemcpyHostToDevice, p->stream));
    pose_ptr = d_pose.p;
  }
  CK(launch_eval_rj(p->model, p->one_focal, p->dev(), p->intr_dev.p, pose_ptr, apply_loss, d_r.p, J ? d_J.p : nullptr,
                    p->n_obs, p->stream));
  p->launches++;
  CK(cudaMemcpyAsync(r, d_r.p, 2 * N * 8, cudaMemcpyDeviceToHost, p->stream));
  if (J) CK(cudaMemcpyAsync(J, d_J.p, 2 * N * n * 8, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  d_r.release(); d_J.release(); d_pose.release();
  return 0;
}

// validation (src/util.rs:721-795): K6 per-observation errors, then two order statistics and a masked sum by radix
// select on the device; only 3 x 48 bytes of state + the per-CTA partial sums come back.
int ccrs_validation(ccrs_problem* p, const double* intr, const double* poses, double* median, double* avg99,
                    double* errors) {
  if (!p || !intr || !median || !avg99) return fail(CCRS_ERR_INVALID, "null");
  if (p->batch) return fail(CCRS_ERR_INVALID, "ccrs_validation is for single-problem handles");
  CK(cudaSetDevice(p->device));
  int st = flush_pending(p);
  if (st) return st;
  const size_t N = (size_t)p->n_obs;
  if (N == 0) return fail(CCRS_ERR_INVALID, "no observations");
  st = ensure_obs_frame(p);
  if (st) return st;
  st = upload_intr(p, intr);
  if (st) return st;
  DevBuf<double> d_err, d_pose, d_part;
  DevBuf<unsigned int> d_hist;
  DevBuf<unsigned long long> d_state;
  const int n_ctas = p->n_sms * 4;
  CK(d_err.alloc(N)); CK(d_part.alloc(n_ctas)); CK(d_hist.alloc(2 * kSelBins)); CK(d_state.alloc(6));
  const double* pose_ptr = p->poses[p->cur_val].p;
  if (poses) {
    CK(d_pose.alloc((size_t)p->n_frames * 6));
    CK(cudaMemcpyAsync(d_pose.p, poses, (size_t)p->n_frames * 48, cudaMemcpyHostToDevice, p->stream));
    pose_ptr = d_pose.p;
  }
  // reprojection_errors[len / 2]; mean of the first len * 99 / 100 sorted errors (util.rs:771-781)
  const unsigned long long len99 = (unsigned long long)N * 99ull / 100ull;
  SelectState h_st{};
  h_st.rank[0] = (unsigned long long)(N / 2);
  h_st.rank[1] = len99 > 0 ? len99 - 1 : 0;
  CK(cudaMemcpyAsync(d_state.p, &h_st, sizeof(h_st), cudaMemcpyHostToDevice, p->stream));
  CK(cudaMemsetAsync(d_hist.p, 0, 2 * kSelBins * sizeof(unsigned int), p->stream));
  CK(launch_reproj_err(p->model, p->one_focal, p->dev(), p->intr_dev.p, pose_ptr, d_err.p, p->n_obs, p->stream));
  p->launches++;
  CK(launch_select(d_err.p, (int64_t)N, reinterpret_cast<SelectState*>(d_state.p), d_hist.p, d_part.p, n_ctas, p->stream,
                   &p->launches));
  std::vector<double> part(n_ctas);
  CK(cudaMemcpyAsync(&h_st, d_state.p, sizeof(h_st), cudaMemcpyDeviceToHost, p->stream));
  CK(cudaMemcpyAsync(part.data(), d_part.p, (size_t)n_ctas * 8, cudaMemcpyDeviceToHost, p->stream));
  if (errors) CK(cudaMemcpyAsync(errors, d_err.p, N * 8, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  d_err.release(); d_pose.release(); d_part.release(); d_hist.release(); d_state.release();
  double med, t99;
  std::memcpy(&med, &h_st.prefix[0], 8);
  std::memcpy(&t99, &h_st.prefix[1], 8);
  *median = med;
  if (len99 == 0) { *avg99 = 0.0; return 0; }
  double sum_below = 0.0;
  for (int c = 0; c < n_ctas; ++c) sum_below += part[c];      // CTA order: fixed
  const double copies = (double)(len99 - h_st.below[1]);      // elements equal to the 99 % threshold that are kept
  *avg99 = (sum_below + copies * t99) / (double)len99;
  if (std::isnan(*median) || std::isnan(*avg99)) return fail(CCRS_ERR_NUMERIC, "NaN reprojection error");
  return 0;
}

int ccrs_linearize(ccrs_problem* p, const double* intr, int which, double* sq_err) {
  if (!p || !intr) return fail(CCRS_ERR_INVALID, "null");
  CK(cudaSetDevice(p->device));
  double seq = 0.0;
  int st = do_linearize(p, intr, which ? 1 : 0, false, sq_err != nullptr, &seq);
  if (st) return st;
  if (sq_err) {
    std::vector<double> stats((size_t)p->n_problems * 2);
    st = fetch_stats(p, which ? 1 : 2, seq, stats.data());
    if (st) return st;
    for (int q = 0; q < p->n_problems; ++q) sq_err[q] = stats[2 * q + 1];
  }
  return 0;
}

int ccrs_get_frame_blocks(ccrs_problem* p, int which, double* blocks) {
  if (!p || !blocks) return fail(CCRS_ERR_INVALID, "null");
  CK(cudaSetDevice(p->device));
  const size_t n = (size_t)p->NBLK * p->Fs;
  std::vector<double> b[2];
  std::vector<int32_t> cur(p->n_problems, p->cur_val);
  if (p->batch) CK(cudaMemcpyAsync(cur.data(), p->cur.p, cur.size() * 4, cudaMemcpyDeviceToHost, p->stream));
  for (int i = 0; i < 2; ++i) { b[i].resize(n); CK(cudaMemcpyAsync(b[i].data(), p->blocks[i].p, n * 8, cudaMemcpyDeviceToHost, p->stream)); }
  CK(cudaStreamSynchronize(p->stream));
  for (int q = 0; q < p->n_problems; ++q) {
    const std::vector<double>& src = b[cur[q] ^ (which ? 1 : 0)];
    for (int f = p->h_problem_frame_offsets[q]; f < p->h_problem_frame_offsets[q + 1]; ++f)
      for (int e = 0; e < p->NBLK; ++e) blocks[(size_t)f * p->NBLK + e] = src[(size_t)e * p->Fs + f];
  }
  return 0;
}

int ccrs_compute_scale(ccrs_problem* p, int which, double* col_sq) {
  if (!p || !col_sq) return fail(CCRS_ERR_INVALID, "null");
  CK(cudaSetDevice(p->device));
  // frame_red is free between reduce() calls: use it for the per-frame A_aa diagonals [D][Fs]
  CK(launch_compute_scale(p->D, p->dev(), which, p->pose_scale.p, p->frame_red.p, p->stream));
  p->launches++;
  if (!p->batch && !p->comm) {
    // single problem, single GPU: the frame-order sums go straight to mapped host memory (armed words, no memcpy, no
    // stream synchronise) like every other per-iteration result
    arm_payload(p->h_colsq.p, p->D);
    CK(launch_segreduce(p->frame_red.p, p->D, p->Fs, p->problem_frame_offsets.p, 1, p->h_colsq.p, p->stream));
    p->launches++;
    int st1 = wait_payload(p, p->h_colsq.p, p->D);
    if (st1) return st1;
    for (int i = 0; i < p->D; ++i) col_sq[i] = p->h_colsq.p[i];
    return 0;
  }
  int st = reduce_frames(p, p->frame_red.p, p->D, p->red_out.p);
  if (st) return st;
  st = exchange(p, p->red_out.p, (size_t)p->n_problems * p->D, nullptr, 0.0);
  if (st) return st;
  std::vector<double> tmp((size_t)p->n_problems * p->D);
  CK(cudaMemcpyAsync(tmp.data(), p->red_out.p, tmp.size() * 8, cudaMemcpyDeviceToHost, p->stream));
  CK(cudaStreamSynchronize(p->stream));
  std::memcpy(col_sq, tmp.data(), tmp.size() * 8);
  return 0;
}

int ccrs_set_intr_scale(ccrs_problem* p, const double* intr_scale) {
  if (!p) return fail(CCRS_ERR_INVALID, "null");
  CK(cudaSetDevice(p->device));
  if (!intr_scale) { p->have_scale = false; return 0; }
  if (!p->batch) for (int i = 0; i < p->D; ++i) p->h_scale[i] = intr_scale[i];
  CK(cudaMemcpyAsync(p->scale_dev.p, intr_scale, (size_t)p->n_problems * p->D * 8, cudaMemcpyHostToDevice, p->stream));
  p->have_scale = true;
  return 0;
}

int ccrs_reduce(ccrs_problem* p, int which, const double* u, int use_scale, double min_diag, double max_diag, double* out) {
  if (!p || !out) return fail(CCRS_ERR_INVALID, "null");
  CK(cudaSetDevice(p->device));
  return do_reduce(p, which, u, use_scale, min_diag, max_diag, out);
}

int ccrs_backsub(ccrs_problem* p, const double* y_a, const double* u, int in_place, double* model_dec) {
  if (!p || !y_a) return fail(CCRS_ERR_INVALID, "null");
  CK(cudaSetDevice(p->device));
  int st = defer_backsub(p, y_a, u, nullptr, in_place);
  if (st) return st;
  p->pend = false;
  st = launch_backsub_now(p, in_place, model_dec != nullptr);
  if (st) return st;
  if (model_dec) {
    CK(launch_trial_stats(p->dev(), p->NBLK - 1, 3, p->frame_md.p, p->stat_out.p, p->stream));
    p->launches++;
    st = exchange(p, p->stat_out.p, (size_t)p->n_problems * 2, nullptr, 0.0);
    if (st) return st;
    std::vector<double> tmp((size_t)p->n_problems * 2);
    CK(cudaMemcpyAsync(tmp.data(), p->stat_out.p, tmp.size() * 8, cudaMemcpyDeviceToHost, p->stream));
    CK(cudaStreamSynchronize(p->stream));
    for (int q = 0; q < p->n_problems; ++q) model_dec[q] = tmp[2 * q];
  } else {
    CK(cudaStreamSynchronize(p->stream));
  }
  return 0;
}

int ccrs_eval_cost(ccrs_problem* p, const double* intr, int which, double* sq_err) {
  if (!p || !intr || !sq_err) return fail(CCRS_ERR_INVALID, "null");
  CK(cudaSetDevice(p->device));
  double seq = 0.0;
  int st = do_linearize(p, intr, which ? 1 : 0, true, true, &seq);
  if (st) return st;
  std::vector<double> stats((size_t)p->n_problems * 2);
  st = fetch_stats(p, which ? 0 : 4, seq, stats.data());
  if (st) return st;
  for (int q = 0; q < p->n_problems; ++q) sq_err[q] = stats[2 * q + 1];
  return 0;
}

int ccrs_accept(ccrs_problem* p, const unsigned char* mask) {
  if (!p) return fail(CCRS_ERR_INVALID, "null");
  CK(cudaSetDevice(p->device));
  int st = flush_pending(p);
  if (st) return st;
  if (!p->batch) {
    if (!mask || mask[0]) p->cur_val ^= 1;
    return 0;
  }
  const unsigned char* md = nullptr;
  if (mask) {
    CK(cudaMemcpyAsync(p->mask_dev.p, mask, (size_t)p->n_problems, cudaMemcpyHostToDevice, p->stream));
    md = p->mask_dev.p;
  }
  CK(launch_flip_cur(p->cur.p, md, p->n_problems, p->stream));
  p->launches++;
  return 0;
}

// Build the peer-memory exchange after the communicator exists: all-gather the IPC handles over NCCL, open the
// peers' buffers. Any failure (no P2P, different nodes, CCRS_P2P=0) leaves g_peer.ok false: the NCCL path is used.
static void peer_teardown() {
  for (int r = 0; r < kXchgMaxRanks; ++r) {
    if (g_peer.peer[r] && g_peer.peer[r] != g_peer.local) cudaIpcCloseMemHandle(g_peer.peer[r]);
    g_peer.peer[r] = nullptr;
  }
  if (g_peer.local) cudaFree(g_peer.local);
  g_peer = GlobalPeer{};
}
static int peer_setup(int rank, int world, cudaStream_t s) {
  peer_teardown();
  if (const char* e = getenv("CCRS_P2P")) if (atoi(e) == 0) return 0;
  if (world < 2 || world > kXchgMaxRanks) return 0;
  NcclApi& n = nccl();
  if (cudaMalloc((void**)&g_peer.local, g_peer.bytes()) != cudaSuccess) { cudaGetLastError(); g_peer.local = nullptr; return 0; }
  if (launch_arm(g_peer.local, g_peer.doubles(), s) != cudaSuccess) return 0;
  if (cudaMemsetAsync(g_peer.dev_count(), 0, 2 * sizeof(double), s) != cudaSuccess) return 0;
  cudaIpcMemHandle_t mine;
  int have = cudaIpcGetMemHandle(&mine, g_peer.local) == cudaSuccess ? 1 : 0;
  if (!have) cudaGetLastError();
  // every rank must take the same path: gather {have flag, handle} from everybody
  constexpr size_t REC = sizeof(cudaIpcMemHandle_t) + 8;
  unsigned char rec[REC] = {};
  rec[0] = (unsigned char)have;
  std::memcpy(rec + 8, &mine, sizeof(mine));
  unsigned char *d_send = nullptr, *d_recv = nullptr;
  if (cudaMalloc((void**)&d_send, REC) != cudaSuccess || cudaMalloc((void**)&d_recv, REC * world) != cudaSuccess) return 0;
  std::vector<unsigned char> all(REC * world);
  int ok = 1;
  if (cudaMemcpyAsync(d_send, rec, REC, cudaMemcpyHostToDevice, s) != cudaSuccess) ok = 0;
  if (ok && n.AllGather(d_send, d_recv, REC, 0 /* ncclInt8 */, g_comm.comm, s) != 0) ok = 0;
  if (ok && cudaMemcpyAsync(all.data(), d_recv, REC * world, cudaMemcpyDeviceToHost, s) != cudaSuccess) ok = 0;
  if (cudaStreamSynchronize(s) != cudaSuccess) ok = 0;
  cudaFree(d_send); cudaFree(d_recv);
  if (!ok) { cudaGetLastError(); return 0; }
  for (int r = 0; r < world; ++r) if (!all[REC * r]) ok = 0;
  int opened = ok;
  if (ok) {
    for (int r = 0; r < world && opened; ++r) {
      if (r == rank) { g_peer.peer[r] = g_peer.local; continue; }
      cudaIpcMemHandle_t hnd;
      std::memcpy(&hnd, all.data() + REC * r + 8, sizeof(hnd));
      void* ptr = nullptr;
      if (cudaIpcOpenMemHandle(&ptr, hnd, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { cudaGetLastError(); opened = 0; break; }
      g_peer.peer[r] = (double*)ptr;
    }
  }
  // agree on the outcome (a rank that could not open a peer must not leave the others spinning on it)
  double flag = opened ? 1.0 : 0.0, *d_flag = nullptr;
  if (cudaMalloc((void**)&d_flag, 8) != cudaSuccess) return 0;
  cudaMemcpyAsync(d_flag, &flag, 8, cudaMemcpyHostToDevice, s);
  // min over ranks via sum of (1 - flag): all-reduce of a double
  double miss = 1.0 - flag;
  cudaMemcpyAsync(d_flag, &miss, 8, cudaMemcpyHostToDevice, s);
  int r2 = n.AllReduce(d_flag, d_flag, 1, kNcclFloat64, kNcclSum, g_comm.comm, s);
  cudaMemcpyAsync(&miss, d_flag, 8, cudaMemcpyDeviceToHost, s);
  cudaStreamSynchronize(s);
  cudaFree(d_flag);
  g_peer.ok = (r2 == 0 && miss == 0.0);
  return 0;
}

int ccrs_comm_unique_id(void* unique_id_128) {
  NcclApi& n = nccl();
  if (!n.ok) return fail(CCRS_ERR_COMM, "libnccl.so.2 not loadable: %s", dlerror());
  ncclUniqueId id;
  int r = n.GetUniqueId(&id);
  if (r != 0) return fail(CCRS_ERR_COMM, "ncclGetUniqueId failed (%d)", r);
  std::memcpy(unique_id_128, &id, 128);
  return 0;
}

int ccrs_comm_init(ccrs_problem* p, const void* unique_id_128, int rank, int world_size) {
  if (!p) return fail(CCRS_ERR_INVALID, "null");
  if (!unique_id_128) {  // attach the communicator this process already created
    if (!g_comm.comm) return fail(CCRS_ERR_COMM, "no communicator yet: call ccrs_comm_init with a unique id first");
    p->comm = g_comm.comm; p->rank = g_comm.rank; p->world = g_comm.world;
    return 0;
  }
  if (world_size < 1 || rank < 0 || rank >= world_size) return fail(CCRS_ERR_INVALID, "bad comm args");
  NcclApi& n = nccl();
  if (!n.ok) return fail(CCRS_ERR_COMM, "libnccl.so.2 not loadable");
  CK(cudaSetDevice(p->device));
  if (g_comm.comm) { n.CommDestroy(g_comm.comm); g_comm.comm = nullptr; }
  ncclUniqueId id;
  std::memcpy(&id, unique_id_128, 128);
  int r = n.CommInitRank(&g_comm.comm, world_size, id, rank);
  if (r != 0) return fail(CCRS_ERR_COMM, "ncclCommInitRank: %s", n.GetErrorString ? n.GetErrorString(r) : "?");
  g_comm.rank = rank; g_comm.world = world_size;
  p->comm = g_comm.comm; p->rank = rank; p->world = world_size;
  peer_setup(rank, world_size, p->stream);
  return 0;
}

int ccrs_comm_uses_peer_memory(void) { return g_peer.ok ? 1 : 0; }

int ccrs_comm_finalize(void) {
  peer_teardown();
  if (g_comm.comm && nccl().ok) nccl().CommDestroy(g_comm.comm);
  g_comm.comm = nullptr;
  return 0;
}

int ccrs_comm_set_deterministic(ccrs_problem* p, int deterministic) {
  if (!p) return fail(CCRS_ERR_INVALID, "null");
  p->deterministic = deterministic ? 1 : 0;
  return 0;
}

static int timed_solve(ccrs_problem* p, bool lm, double* intr, const double* lo, const double* hi,
                       const unsigned char* fixed, const ccrs_options* opt, ccrs_summary* summary, double* err_hist) {
  if (!p || !intr) return fail(CCRS_ERR_INVALID, "null");
  CK(cudaSetDevice(p->device));
  struct Events {   // destroyed on every return path
    cudaEvent_t a = nullptr, b = nullptr;
    ~Events() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); }
  } ev;
  CK(cudaEventCreate(&ev.a)); CK(cudaEventCreate(&ev.b));
  cudaEvent_t e0 = ev.a, e1 = ev.b;
  CK(cudaEventRecord(e0, p->stream));
  ccrs_backend be = cuda_backend(p);
  ccrs_summary local;
  if (!summary) summary = &local;
  ccrs_options o;
  if (opt) o = *opt; else ccrs_default_options(&o);
  int st;
  if (device_loop_ok(p, lm, o)) st = run_device_loop(p, lm, intr, lo, hi, fixed, o, summary, err_hist);
  else if (batch_loop_ok(p, lm, o)) st = run_batch_loop(p, lm, intr, lo, hi, fixed, o, summary, err_hist);
  else st = lm ? ccrs_controller_lm(&be, intr, lo, hi, fixed, opt, summary, err_hist)
               : ccrs_controller_gn(&be, intr, lo, hi, fixed, opt, summary, err_hist);
  cudaEventRecord(e1, p->stream);
  cudaEventSynchronize(e1);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, e0, e1);
  summary->device_ms = ms;
  if (st == CCRS_ERR_CHOLESKY) fail(st, "Cholesky failure (non-positive pivot)");
  if (st == CCRS_ERR_NUMERIC) fail(st, "NaN error");
  return st;
}

int ccrs_solve_gn(ccrs_problem* p, double* intr, const double* lo, const double* hi, const unsigned char* fixed,
                  const ccrs_options* opt, ccrs_summary* summary, double* err_hist) {
  return timed_solve(p, false, intr, lo, hi, fixed, opt, summary, err_hist);
}
int ccrs_solve_lm(ccrs_problem* p, double* intr, const double* lo, const double* hi, const unsigned char* fixed,
                  const ccrs_options* opt, ccrs_summary* summary, double* err_hist) {
  return timed_solve(p, true, intr, lo, hi, fixed, opt, summary, err_hist);
}

int ccrs_model_bounds(int model, int width, int height, double* lo, double* hi) {
  const int n = ccrs_model_nparams(model);
  if (n < 0 || !lo || !hi) return fail(CCRS_ERR_INVALID, "bad model");
  const double inf = std::numeric_limits<double>::infinity();
  for (int i = 0; i < n; ++i) { lo[i] = -inf; hi[i] = inf; }
  lo[0] = 0; hi[0] = 1e4; lo[1] = 0; hi[1] = 1e4;       // util.rs:36-37
  lo[2] = 0; hi[2] = width; lo[3] = 0; hi[3] = height;  // util.rs:38-39
  // GenericModel::distortion_params_bound() of camera-intrinsic-model 0.8 is not in the reference tree (un-vendored
  // crate): only the bounds the projection itself requires are applied — alpha in (0, 1], beta > 0 of the unified
  // models. Polynomial / tangential coefficients (KB4, OPENCV5, FTHETA k*, EUCMT t1 t2) stay UNBOUNDED rather than being
  // clamped to an invented box: real lenses have |k| > 1 and a silent clamp would converge to a different answer.
  switch (model) {
    case CCRS_UCM: lo[4] = 1e-6; hi[4] = 1.0; break;
    case CCRS_EUCM: case CCRS_EUCMT: lo[4] = 1e-6; hi[4] = 1.0; lo[5] = 1e-6; hi[5] = 100.0; break;
    default: break;
  }
  return 0;
}

int ccrs_calib_camera(int model, int width, int height, int n_frames, const int32_t* frame_offsets, const double* x,
                      const double* y, const double* z, const double* u, const double* v, double* params, double* poses,
                      int xy_same_focal, int disabled_distortions, int fixed_focal, int use_lm, const ccrs_options* opt,
                      ccrs_summary* summary, int device_id) {
  if (!params || !poses) return fail(CCRS_ERR_INVALID, "null");
  const int nfull = ccrs_model_nparams(model);
  if (nfull < 0) return fail(CCRS_ERR_INVALID, "bad model");
  const int shift = xy_same_focal ? 1 : 0;
  const int d = nfull - shift;
  if (disabled_distortions < 0 || disabled_distortions > nfull - 4) return fail(CCRS_ERR_INVALID, "disabled_distortions out of range");
  ccrs_problem* p = nullptr;
  int st = ccrs_problem_create(&p, model, width, height, xy_same_focal, n_frames, frame_offsets, x, y, z, u, v, 1.0, device_id);
  if (st) return st;
  // params.remove_row(1) (util.rs:391-395)
  std::vector<double> intr(d), lo(d), hi(d), flo(nfull), fhi(nfull);
  std::vector<unsigned char> fixed(d, 0);
  ccrs_model_bounds(model, width, height, flo.data(), fhi.data());
  for (int i = 0, j = 0; i < nfull; ++i) {
    if (xy_same_focal && i == 1) continue;
    intr[j] = params[i]; lo[j] = flo[i]; hi[j] = fhi[i]; ++j;
  }
  // set_problem_parameter_disabled (util.rs:50-71): fix + zero the last N distortion parameters
  for (int i = 0; i < disabled_distortions; ++i) { const int idx = nfull - 1 - shift - i; fixed[idx] = 1; intr[idx] = 0.0; }
  st = ccrs_set_poses(p, poses);
  ccrs_summary s1{}, s2{};
  if (!st) st = use_lm ? ccrs_solve_lm(p, intr.data(), lo.data(), hi.data(), fixed.data(), opt, &s1, nullptr)
                       : ccrs_solve_gn(p, intr.data(), lo.data(), hi.data(), fixed.data(), opt, &s1, nullptr);
  if (!st && fixed_focal) {  // util.rs:459-464: fix params[0], reset it to the input focal, optimise again
    fixed[0] = 1;
    intr[0] = params[0];
    st = use_lm ? ccrs_solve_lm(p, intr.data(), lo.data(), hi.data(), fixed.data(), opt, &s2, nullptr)
                : ccrs_solve_gn(p, intr.data(), lo.data(), hi.data(), fixed.data(), opt, &s2, nullptr);
    s1.iterations += s2.iterations; s1.device_ms += s2.device_ms; s1.final_error = s2.final_error;
    s1.stop_reason = s2.stop_reason; s1.n_accepted += s2.n_accepted; s1.n_rejected += s2.n_rejected;
  }
  s1.status = st;
  if (summary) *summary = s1;
  if (!st) {
    // new_params.insert_row(1, new_params[0]) (util.rs:466-470)
    for (int i = 0, j = 0; i < nfull; ++i) {
      if (xy_same_focal && i == 1) { params[1] = intr[0]; continue; }
      params[i] = intr[j++];
    }
    st = ccrs_get_poses(p, poses);
  }
  ccrs_problem_destroy(p);
  return st;
}

int ccrs_init_poses(int n_frames, const int32_t* frame_offsets, const double* x, const double* y, const double* z,
                    const double* xn, const double* yn, double* poses_out, double* cost_out, int device_id) {
  if (!frame_offsets || !x || !y || !z || !xn || !yn || !poses_out || n_frames <= 0) return fail(CCRS_ERR_INVALID, "null / empty");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); return fail(CCRS_ERR_NO_DEVICE, "no CUDA device"); }
  CK(cudaSetDevice(device_id));
  for (int f = 0; f < n_frames; ++f)
    if (frame_offsets[f + 1] - frame_offsets[f] < 4) return fail(CCRS_ERR_INVALID, "frame %d has fewer than 4 points", f);
  const size_t N = (size_t)frame_offsets[n_frames], F = (size_t)n_frames;
  cudaStream_t s = nullptr;
  CK(get_stream(device_id, &s));
  DevBuf<double> dx, dy, dz, dxn, dyn, dpose, dcost;
  DevBuf<int32_t> dfo;
  CK(dx.alloc(N)); CK(dy.alloc(N)); CK(dz.alloc(N)); CK(dxn.alloc(N)); CK(dyn.alloc(N)); CK(dpose.alloc(6 * F)); CK(dcost.alloc(F));
  CK(dfo.alloc(F + 1));
  CK(cudaMemcpyAsync(dfo.p, frame_offsets, (F + 1) * 4, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dx.p, x, N * 8, cudaMemcpyHostToDevice, s)); CK(cudaMemcpyAsync(dy.p, y, N * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dz.p, z, N * 8, cudaMemcpyHostToDevice, s)); CK(cudaMemcpyAsync(dxn.p, xn, N * 8, cudaMemcpyHostToDevice, s));
  CK(cudaMemcpyAsync(dyn.p, yn, N * 8, cudaMemcpyHostToDevice, s));
  CK(launch_pnp(dfo.p, dx.p, dy.p, dz.p, dxn.p, dyn.p, n_frames, dpose.p, dcost.p, s));
  std::vector<double> cost(F);
  CK(cudaMemcpyAsync(poses_out, dpose.p, 6 * F * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(cost.data(), dcost.p, F * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  put_stream(device_id, s);
  dx.release(); dy.release(); dz.release(); dxn.release(); dyn.release(); dpose.release(); dcost.release(); dfo.release();
  if (cost_out) std::memcpy(cost_out, cost.data(), F * 8);
  for (size_t f = 0; f < F; ++f)
    if (std::isnan(cost[f])) return fail(CCRS_ERR_NUMERIC, "no pose with the board in front of the camera for frame %zu", f);
  return 0;
}

int ccrs_loop_trace(int enable, double* avg_us /* [13] or NULL */, int64_t* n_iterations) {
  if (avg_us) {
    for (int i = 0; i < 5; ++i) avg_us[i] = g_loop_trace.n ? g_loop_trace.acc[i] * 1e-3 / g_loop_trace.n : 0.0;
    for (int i = 0; i < 7; ++i) avg_us[5 + i] = g_loop_trace.n ? g_loop_trace.fine[i] * 1e-3 / g_loop_trace.n : 0.0;
    avg_us[12] = g_loop_trace.n ? g_loop_trace.wake * 1e-3 / g_loop_trace.n : 0.0;
  }
  if (n_iterations) *n_iterations = g_loop_trace.n;
  g_loop_trace = LoopTrace{};
  g_loop_trace.on = enable != 0;
  return 0;
}

int ccrs_loop_counters(int64_t* audited_solves) {
  if (audited_solves) *audited_solves = g_loop_audited.load();
  return g_device_loop ? 1 : 0;
}

int ccrs_spec_k3_counters(int64_t* launched, int64_t* hits) {
  if (launched) *launched = g_spec_launched.load();
  if (hits) *hits = g_spec_hits.load();
  return g_spec_enabled ? 1 : 0;
}

int ccrs_set_fixed_poses(ccrs_problem* p, int fixed) {
  if (!p) return fail(CCRS_ERR_INVALID, "null");
  p->fixed_poses = fixed != 0;
  return 0;
}

int ccrs_init_ucm(int width, int height, int n_frames, const int32_t* frame_offsets, const double* x, const double* y,
                  const double* z, const double* u, const double* v, double init_f, double init_alpha, int fixed_focal,
                  double* poses, double* params_out, const ccrs_options* opt, ccrs_summary* summary, int device_id) {
  if (!poses || !params_out || !frame_offsets) return fail(CCRS_ERR_INVALID, "null");
  const double half_w = width / 2.0, half_h = height / 2.0;          // util.rs:293-294
  // stage 1: one-focal UCM [f, cx, cy, alpha]; cx, cy are constants of UCMInitFocalAlphaFactor -> mask 2
  ccrs_problem* p = nullptr;
  int st = ccrs_problem_create(&p, CCRS_UCM, width, height, 1, n_frames, frame_offsets, x, y, z, u, v, 1.0, device_id);
  if (st) return st;
  const double inf = std::numeric_limits<double>::infinity();
  double intr[4] = {init_f, half_w, half_h, init_alpha};
  const double lo[4] = {init_f / 3.0, -inf, -inf, 1e-6}, hi[4] = {init_f * 3.0, inf, inf, 1.0};   // util.rs:337-338
  const unsigned char fixed[4] = {(unsigned char)(fixed_focal ? 1 : 0), 2, 2, 0};                  // util.rs:330-332
  ccrs_summary s1{}, s2{};
  st = ccrs_set_poses(p, poses);
  if (!st) st = ccrs_solve_gn(p, intr, lo, hi, fixed, opt, &s1, nullptr);
  if (!st) st = ccrs_get_poses(p, poses);
  ccrs_problem_destroy(p);
  if (st) { s1.status = st; if (summary) *summary = s1; return st; }
  // stage 2: calib_camera(frames, UCM[f f w/2 h/2 alpha], one_focal = true, 0, fixed_focal)  (util.rs:358-372)
  params_out[0] = intr[0]; params_out[1] = intr[0]; params_out[2] = half_w; params_out[3] = half_h; params_out[4] = intr[3];
  // ... which starts from fresh poses: calib_camera unprojects every detection with the new model and solves the PnP
  // per frame (util.rs:418-439); the stage-1 poses are dropped (util.rs:356). UCM unprojection is closed form:
  // m = ((u - cx)/f, (v - cy)/f), mz = (1 - a^2 r^2) / (a sqrt(1 - (2a - 1) r^2) + 1 - a); normalised point m / mz as f32.
  {
    const double f = intr[0], a = intr[3];
    std::vector<int32_t> fo2(1, 0), src_frame;
    std::vector<double> x2, y2, z2, xn, yn;
    for (int fr = 0; fr < n_frames; ++fr) {
      const size_t before = x2.size();
      for (int k = frame_offsets[fr]; k < frame_offsets[fr + 1]; ++k) {
        const double mx = (u[k] - half_w) / f, my = (v[k] - half_h) / f, r2 = mx * mx + my * my;
        const double disc = 1.0 - (2.0 * a - 1.0) * r2;
        if (disc < 0.0) continue;                                     // outside the model's domain: unproject -> None
        const double mz = (1.0 - a * a * r2) / (a * std::sqrt(disc) + 1.0 - a);
        if (!(std::fabs(mz) > 1e-12)) continue;
        x2.push_back(x[k]); y2.push_back(y[k]); z2.push_back(z[k]);
        xn.push_back((double)(float)(mx / mz)); yn.push_back((double)(float)(my / mz));   // glam::Vec2 (util.rs:425)
      }
      if (x2.size() - before < 10) {                                  // util.rs:431-433: the frame keeps its stage-1 pose here
        x2.resize(before); y2.resize(before); z2.resize(before); xn.resize(before); yn.resize(before);
        continue;
      }
      fo2.push_back((int32_t)x2.size());
      src_frame.push_back(fr);
    }
    if (!src_frame.empty()) {
      std::vector<double> pnp(6 * src_frame.size());
      st = ccrs_init_poses((int)src_frame.size(), fo2.data(), x2.data(), y2.data(), z2.data(), xn.data(), yn.data(), pnp.data(),
                           nullptr, device_id);
      if (st) { s1.status = st; if (summary) *summary = s1; return st; }
      for (size_t i = 0; i < src_frame.size(); ++i)
        for (int j = 0; j < 6; ++j) poses[6 * (size_t)src_frame[i] + j] = pnp[6 * i + j];
    }
  }
  st = ccrs_calib_camera(CCRS_UCM, width, height, n_frames, frame_offsets, x, y, z, u, v, params_out, poses, 1, 0,
                         fixed_focal, 0, opt, &s2, device_id);
  s2.iterations += s1.iterations; s2.device_ms += s1.device_ms;
  if (summary) *summary = s2;
  return st;
}

int ccrs_convert_model(int src_model, const double* src_params, int tgt_model, double* tgt_params, int width, int height,
                       int disabled_distortions, int n_pts, const double* px, const double* py, const double* pz,
                       const ccrs_options* opt_in, ccrs_summary* summary, int device_id) {
  if (!src_params || !tgt_params || !px || !py || !pz || n_pts <= 0) return fail(CCRS_ERR_INVALID, "null / empty");
  const int ns = ccrs_model_nparams(src_model), nt = ccrs_model_nparams(tgt_model);
  if (ns < 0 || nt < 0) return fail(CCRS_ERR_INVALID, "bad model");
  if (disabled_distortions < 0 || disabled_distortions > nt - 4) return fail(CCRS_ERR_INVALID, "disabled_distortions out of range");
  ccrs_summary sum{};
  if (src_model == CCRS_UCM && (tgt_model == CCRS_EUCM || tgt_model == CCRS_EUCMT)) {   // util.rs:230-243
    for (int i = 0; i < 5; ++i) tgt_params[i] = src_params[i];
    tgt_params[5] = 1.0;
    if (tgt_model == CCRS_EUCMT) { tgt_params[6] = 0.0; tgt_params[7] = 0.0; }
    if (summary) *summary = sum;
    return 0;
  }
  const int32_t fo[2] = {0, n_pts};
  std::vector<double> zeros((size_t)n_pts, 0.0), uv((size_t)2 * n_pts), u0(n_pts), v0(n_pts);
  const double pose[6] = {0, 0, 0, 0, 0, 0};
  // p2ds0 = source.project(p3ds) (factors.rs:59): residual of the source model against zero observations
  ccrs_problem* ps = nullptr;
  int st = ccrs_problem_create(&ps, src_model, width, height, 0, 1, fo, px, py, pz, zeros.data(), zeros.data(), 0.0, device_id);
  if (st) return st;
  st = ccrs_eval_rj(ps, src_params, pose, 0, uv.data(), nullptr);
  ccrs_problem_destroy(ps);
  if (st) return st;
  for (int k = 0; k < n_pts; ++k) { u0[k] = uv[2 * k]; v0[k] = uv[2 * k + 1]; }
  ccrs_problem* pt = nullptr;
  st = ccrs_problem_create(&pt, tgt_model, width, height, 0, 1, fo, px, py, pz, u0.data(), v0.data(), 0.0, device_id);
  if (st) return st;
  std::vector<double> intr(tgt_params, tgt_params + nt), lo(nt), hi(nt);
  for (int i = 0; i < 4; ++i) intr[i] = src_params[i];                                     // util.rs:253-255
  std::vector<unsigned char> fixed(nt, 0);
  ccrs_model_bounds(tgt_model, width, height, lo.data(), hi.data());                         // util.rs:262
  for (int i = 0; i < disabled_distortions; ++i) { fixed[nt - 1 - i] = 1; intr[nt - 1 - i] = 0.0; }   // util.rs:263-270
  ccrs_options opt;
  if (opt_in) opt = *opt_in; else ccrs_default_options(&opt);
  opt.block_huber_delta = 1.0;                                                               // util.rs:250
  st = ccrs_set_fixed_poses(pt, 1);
  if (!st) st = ccrs_set_poses(pt, pose);
  if (!st) st = ccrs_solve_gn(pt, intr.data(), lo.data(), hi.data(), fixed.data(), &opt, &sum, nullptr);
  ccrs_problem_destroy(pt);
  sum.status = st;
  if (summary) *summary = sum;
  if (!st) for (int i = 0; i < nt; ++i) tgt_params[i] = intr[i];
  return st;
}

int ccrs_measure_fp64_peak(int device_id, double* tflops) {
  if (!tflops) return fail(CCRS_ERR_INVALID, "null");
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0) { cudaGetLastError(); return fail(CCRS_ERR_NO_DEVICE, "no CUDA device"); }
  CK(cudaSetDevice(device_id));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device_id));
  const int ctas = prop.multiProcessorCount * 8, iters = 1 << 14;
  double* out = nullptr;
  CK(cudaMalloc((void**)&out, (size_t)ctas * 256 * 8));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    CK(cudaEventRecord(e0, 0));
    CK(launch_fp64_peak(out, ctas, iters, 0));
    CK(cudaEventRecord(e1, 0));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 8.0 * iters * (double)ctas * 256.0;
    best = std::max(best, fl / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  cudaFree(out);
  *tflops = best;
  return 0;
}

struct ccrs_lm_state;
ccrs_lm_state* ccrs_lm_state_create(const ccrs_backend* be, double* intr, const ccrs_options* opt, ccrs_summary* sum);
int ccrs_lm_state_step(ccrs_lm_state* S, int* done);
void ccrs_lm_state_destroy(ccrs_lm_state* S);

int ccrs_bench_lm_steps(ccrs_problem* p, const double* intr0, const double* poses0, int warmup, int steps,
                        int reset_every, int flush_l2, double* step_ms, int64_t* timed_launches) {
  if (!p || !intr0 || !poses0 || !step_ms || steps <= 0 || warmup < 0 || reset_every <= 0) return fail(CCRS_ERR_INVALID, "bad args");
  CK(cudaSetDevice(p->device));
  const size_t flush_n = (size_t)64 << 20;  // 512 MB of doubles > 126 MB L2
  if (flush_l2 && !p->l2_flush.p) CK(p->l2_flush.alloc(flush_n));
  ccrs_options opt;
  ccrs_default_options(&opt);
  opt.min_abs_decrease = -1.0; opt.min_rel_decrease = -1.0; opt.min_error = -1.0;  // never stop: every step does full work
  ccrs_summary sum;
  ccrs_backend be = cuda_backend(p);
  std::vector<double> intr((size_t)p->n_problems * p->D);
  ccrs_lm_state* S = nullptr;
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  int st = 0;
  int64_t timed = 0;
  const bool dev_loop = device_loop_ok(p, true, opt);
  DeviceLoop L;
  int done_any = 0;
  opt.max_iteration = 1 << 30;
  for (int i = 0; i < warmup + steps; ++i) {
    if (i % reset_every == 0) {  // back to the initial point: untimed (pose upload, first linearisation, Jacobi scaling)
      if (S) ccrs_lm_state_destroy(S);
      st = ccrs_set_poses(p, poses0);
      if (st) break;
      std::memcpy(intr.data(), intr0, intr.size() * sizeof(double));
      if (dev_loop) {
        // device-driven loop: control block + linearisation of the start point; the first timed step is then
        // K3 (first reduction incl. Jacobi scaling) + K2, every later one K3 (decision + reduction) + K2
        std::memset(&sum, 0, sizeof(sum));
        st = loop_begin(L, p, true, intr.data(), nullptr, nullptr, nullptr, opt);
        if (!st) st = loop_launch_k2(L);
        if (!st) st = loop_launch_k3(L);   // first reduction (Jacobi scaling, first solve): untimed like the start-up of the host loop
        if (!st) { CK(cudaStreamSynchronize(p->stream)); st = loop_consume(L, &sum, nullptr, &done_any, intr.data()); }
        if (st) break;
      } else {
        S = ccrs_lm_state_create(&be, intr.data(), &opt, &sum);
        if (!S) { st = fail(CCRS_ERR_CUDA, "lm_begin failed (%d)", sum.status); break; }
      }
    }
    if (flush_l2) CK(launch_l2_flush(p->l2_flush.p, flush_n, p->stream));
    if (p->comm && p->world > 1) {
      // rendezvous outside the event bracket: without it a rank's timed step absorbs its peers' flush / reset skew
      // (every step waits for all ranks' partial systems)
      if (!p->l2_flush.p) CK(p->l2_flush.alloc(flush_n));
      if (nccl().AllReduce(p->l2_flush.p, p->l2_flush.p, 1, kNcclFloat64, kNcclSum, p->comm, p->stream) != 0)
        return fail(CCRS_ERR_COMM, "bench rendezvous all-reduce failed");
    }
    // Device-driven loop: flush, rendezvous, e0, K3, K2, e1 are enqueued back to back (the flush takes ~100 us, so the
    // timed kernels are queued long before e0 executes): the bracket holds the iteration's device time, not the host's
    // launch latency of a cold queue. The host-driven path needs the host inside the iteration and synchronises first.
    if (!dev_loop) CK(cudaStreamSynchronize(p->stream));
    const int64_t l0 = p->launches;
    int done = 0;
    CK(cudaEventRecord(e0, p->stream));
    if (dev_loop) {
      // one LM iteration = K2 (apply the step, linearise the trial point: the pass over the observations comes first,
      // straight after the L2 flush) + K3 (accept / reject, reduction, d x d solve, next step)
      st = loop_launch_k2(L);
      if (!st) st = loop_launch_k3(L);
    } else {
      st = ccrs_lm_state_step(S, &done);
    }
    if (st) break;
    CK(cudaEventRecord(e1, p->stream));
    CK(cudaEventSynchronize(e1));
    if (dev_loop) { st = loop_consume(L, &sum, nullptr, &done, intr.data()); if (st) break; }
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    if (i >= warmup) { step_ms[i - warmup] = ms; timed += p->launches - l0; }
  }
  if (timed_launches) *timed_launches = timed;
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  if (S) ccrs_lm_state_destroy(S);
  return st;
}

int ccrs_bench_lm_steps_rotating(ccrs_problem** ps, int n_ps, const double* intr0, const double* poses0, int warmup, int steps,
                                 double* total_ms, int64_t* timed_launches, int64_t* executed_steps) {
  if (!ps || n_ps <= 0 || n_ps > 64 || !intr0 || !poses0 || !total_ms || steps <= 0 || warmup < 0) return fail(CCRS_ERR_INVALID, "bad args");
  for (int h = 0; h < n_ps; ++h)
    if (!ps[h] || ps[h]->batch || ps[h]->device != ps[0]->device || ps[h]->n_frames != ps[0]->n_frames || ps[h]->D != ps[0]->D)
      return fail(CCRS_ERR_INVALID, "replicas must be single-problem handles of one shape on one device");
  CK(cudaSetDevice(ps[0]->device));
  ccrs_options opt;
  ccrs_default_options(&opt);
  opt.min_abs_decrease = -1.0; opt.min_rel_decrease = -1.0; opt.min_error = -1.0;   // never stop: every step does full work
  opt.max_iteration = 1 << 30;
  if (!device_loop_ok(ps[0], true, opt)) return fail(CCRS_ERR_INVALID, "the rotating bench needs the device-driven loop");
  // one stream for all replicas: the steps of different replicas run back to back, in enqueue order
  std::vector<cudaStream_t> own(n_ps);
  for (int h = 0; h < n_ps; ++h) { own[h] = ps[h]->stream; CK(cudaStreamSynchronize(own[h])); ps[h]->stream = ps[0]->stream; }
  cudaStream_t s = ps[0]->stream;
  std::vector<DeviceLoop> L(n_ps);
  std::vector<double> intr((size_t)ps[0]->D);
  ccrs_summary sum;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  int st = 0, done = 0;
  int64_t timed = 0;
  double total = 0.0;
  auto restore = [&]() { for (int h = 0; h < n_ps; ++h) ps[h]->stream = own[h]; if (e0) cudaEventDestroy(e0); if (e1) cudaEventDestroy(e1); };
  // Every replica runs LM iterations 1..kItersPerReplica of the problem (the ones a converging solve runs; the old
  // per-step bench reset every 4 too); then all replicas return to the start point, untimed. After every consumed
  // record the LM decisions of that replica are counted: *executed_steps is the number of linearisations the timed
  // slots really executed (across GPUs a mis-speculated reduction re-reduces first and the K2 slot behind it exits at
  // once: such a slot is time spent, not a step).
  constexpr int kItersPerReplica = 4;
  std::vector<int> given(n_ps, 0), decided(n_ps, 0);
  int64_t executed = 0;
  int used = 0, next = 0;   // steps since the last (re)start; replica of the next step
  auto begin_all = [&]() -> int {
    for (int h = 0; h < n_ps && !st; ++h) {   // start point, first linearisation, first reduction (Jacobi scaling, first solve)
      st = ccrs_set_poses(ps[h], poses0);
      std::memset(&sum, 0, sizeof(sum));
      if (!st) st = loop_begin(L[h], ps[h], true, intr0, nullptr, nullptr, nullptr, opt);
      if (!st) st = loop_launch_k2(L[h]);
      if (!st) st = loop_launch_k3(L[h]);
      given[h] = 0; decided[h] = 0;
    }
    if (st) return st;
    CK(cudaStreamSynchronize(s));
    for (int h = 0; h < n_ps && !st; ++h) st = loop_consume(L[h], &sum, nullptr, &done, intr.data());
    used = 0; next = 0;
    return st;
  };
  auto run = [&]() -> int {
    if (cudaEventCreate(&e0) != cudaSuccess || cudaEventCreate(&e1) != cudaSuccess) return fail(CCRS_ERR_CUDA, "event");
    st = begin_all();
    if (st) return st;
    const int per_start = n_ps * kItersPerReplica;
    auto steps_block = [&](int n, bool timed_block) -> int {
      while (n > 0 && !st) {
        if (used == per_start) { st = begin_all(); if (st) return st; }
        const int c = std::min(n, per_start - used);
        const int first = next;
        if (timed_block && ps[0]->comm && ps[0]->world > 1) {
          // rendezvous outside the event bracket: absorbs the ranks' skew (every step waits for all ranks' partial systems)
          if (!ps[0]->rdv.p) CK(ps[0]->rdv.alloc(16));
          if (nccl().AllReduce(ps[0]->rdv.p, ps[0]->rdv.p, 1, kNcclFloat64, kNcclSum, ps[0]->comm, s) != 0)
            return fail(CCRS_ERR_COMM, "bench rendezvous all-reduce failed");
        }
        int64_t l0 = 0;
        for (int h = 0; h < n_ps; ++h) l0 += ps[h]->launches;
        if (timed_block) CK(cudaEventRecord(e0, s));
        for (int i = 0; i < c && !st; ++i) {
          st = loop_launch_k2(L[next]);
          if (!st) st = loop_launch_k3(L[next]);
          next = (next + 1) % n_ps;
        }
        if (st) return st;
        if (timed_block) CK(cudaEventRecord(e1, s));
        CK(cudaStreamSynchronize(s));
        if (timed_block) {
          float ms = 0.f;
          CK(cudaEventElapsedTime(&ms, e0, e1));
          total += ms;
          for (int h = 0; h < n_ps; ++h) timed += ps[h]->launches;
          timed -= l0;
        }
        for (int i = 0, h = first; i < c && !st; ++i, h = (h + 1) % n_ps) {
          st = loop_consume(L[h], &sum, nullptr, &done, intr.data());
          given[h]++;
          if (!st && timed_block) {   // LM decisions taken = linearisations executed (a slot behind a re-reduction exits at once)
            const int dec = sum.n_accepted + sum.n_rejected;
            executed += dec - decided[h];
            decided[h] = dec;
          } else if (!st) {
            decided[h] = sum.n_accepted + sum.n_rejected;
          }
        }
        used += c;
        n -= c;
      }
      return st;
    };
    st = steps_block(warmup, false);
    if (!st) st = steps_block(steps, true);
    return st;
  };
  st = run();
  cudaStreamSynchronize(s);
  restore();
  *total_ms = total;
  if (timed_launches) *timed_launches = timed;
  if (executed_steps) *executed_steps = executed;
  return st;
}

int ccrs_time_linearize(ccrs_problem* p, const double* intr, int reps, int flush_l2, double* avg_ms) {
  if (!p || !intr || !avg_ms || reps <= 0) return fail(CCRS_ERR_INVALID, "bad args");
  CK(cudaSetDevice(p->device));
  const size_t flush_n = (size_t)64 << 20;  // 512 MB of doubles > 126 MB L2
  if (flush_l2 && !p->l2_flush.p) CK(p->l2_flush.alloc(flush_n));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
  double total = 0.0;
  for (int w = 0; w < 3; ++w) { int st = do_linearize(p, intr, 0, false, false, nullptr); if (st) return st; }
  CK(cudaStreamSynchronize(p->stream));
  if (flush_l2) {
    for (int r = 0; r < reps; ++r) {
      CK(launch_l2_flush(p->l2_flush.p, flush_n, p->stream));
      CK(cudaEventRecord(e0, p->stream));
      int st = do_linearize(p, intr, 0, false, false, nullptr);
      if (st) return st;
      CK(cudaEventRecord(e1, p->stream));
      CK(cudaEventSynchronize(e1));
      float ms = 0.f;
      CK(cudaEventElapsedTime(&ms, e0, e1));
      total += ms;
    }
  } else {
    CK(cudaEventRecord(e0, p->stream));
    for (int r = 0; r < reps; ++r) { int st = do_linearize(p, intr, 0, false, false, nullptr); if (st) return st; }
    CK(cudaEventRecord(e1, p->stream));
    CK(cudaEventSynchronize(e1));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, e0, e1));
    total = ms;
  }
  cudaEventDestroy(e0); cudaEventDestroy(e1);
  *avg_ms = total / reps;
  return 0;
}

}  // extern "C"

#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its quoted configuration.

metric   : residual+Jacobian evals/s (and LM iterations/s) at ~1M corner observations, EUCM
workload : configs[3] — ONE EUCM calibration problem of 7,000 frames x 144 corners (1,007,999 obs after the
           image-bounds filter), synthetic (SURVEY.md §8(d) generator, seed 3). At N > 1 the SAME problem is
           frame-sharded over the N ranks ("scaling": "strong", 7000/N frames per GPU), one exchange of the reduced
           intrinsic system per linearisation (fused into the kernels over peer memory; NCCL all-gather + rank-order
           sum when peer memory is unavailable). The weak-scaling figure of round 1 (7,000 frames PER GPU) is kept as
           the extra key `weak_scaling`.
step     : one Levenberg-Marquardt iteration of that problem = reduce (+exchange) -> d x d solve -> pose
           back-substitution -> linearisation of the trial point (speculative LM: the trial cost comes from the
           linearisation itself) -> accept/reject. Stop tests are disabled so each of the K timed steps does the full
           work; every step evaluates residual+Jacobian once per observation. Every 4 steps the state returns
           (untimed) to the perturbed start, so the timed steps are LM iterations 1-4 of the problem (iteration 1
           has the Huber loss active on ~99% of the observations).
value    : total observations over all ranks / time per step, inputs resident in HBM and LARGER THAN L2: the rank's
           shard exists in R replicas (R x shard >= 2.5 x L2) that are visited round-robin, step i on replica i mod R, so a
           replica's arrays have left L2 when its turn comes again; the K steps are enqueued back to back (as a solve
           enqueues them) inside ONE CUDA-event bracket; max over ranks. The isolated step of rounds 1-2 (L2 flushed by a
           512 MB write before every step, one event bracket per step, host synchronisation in between) is kept as
           `isolated_step_l2_flushed`.
e2e      : the same metric through the C-ABI entry point a user calls with HOST buffers: per step one complete
           ccrs_problem_create_f32 (H2D of the f32 observation arrays from pinned memory) + ccrs_set_poses +
           ccrs_solve_lm to convergence + ccrs_get_poses (D2H) + destroy; evals = observations x linearisations.
multi_gpu_check (N > 1, same run): intrinsics bitwise-identical across ranks, equal iteration count and <= 1e-6
           relative difference against a single-rank solve of the same problem on rank 0 — for the peer-memory
           exchange AND for CCRS_P2P=0 (NCCL all-gather + rank-order sum).
batch    : BASELINE configs[4] — independent KB4 calibrations (200 frames each), 512 problems per rank, no
           communication (4,096 problems at N = 8). Extra key `batch` of the default line; `--workload batch` prints
           it as a line of its own.

`--impl reference` times the reference arm: the CPU oracle (oracle/, a restatement of the reference's num-dual +
tiny-solver path: the Rust reference cannot be built here) running the same LM iteration on the host cores.
"""
from __future__ import annotations

import argparse
import hashlib
import importlib
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

FRAMES_TOTAL = 7000
MODEL = "eucm"
WORKLOAD = ("EUCM, one calibration problem of 7000 frames x 144 corners (~1.008M obs), 1024x1024, synthetic seed 3, "
            "Huber(1.0), LM iteration; frame-sharded over the GPUs")
METRIC = "residual+Jacobian evals/s at 1M corner obs (EUCM LM iteration)"
BATCH_PER_RANK = 512
BATCH_FRAMES = 200
BATCH_MODEL = "kb4"
# FP64 work per observation of K2: SURVEY §8(d)'s nominal figure and the count from the kernel's SASS
FLOP_PER_OBS_NOMINAL = 560.0
FLOP_PER_OBS_COUNTED = 382.0


def load_pkg():
    return importlib.import_module("camera-intrinsic-calibration-rs_b200")


def workload_config(n_total: int):
    """Identical in both arms (ours / reference): only what defines the workload."""
    return {"workload": WORKLOAD, "camera_model": MODEL, "frames_total": FRAMES_TOTAL, "obs_total": int(n_total),
            "seed": 3, "loop": "LM iteration (speculative: trial cost from the trial linearisation)",
            "l2": "GPU arm: inputs larger than L2 (replicas of the problem visited round-robin, >= 2.5 x L2 in total); no flush kernel"}


K2_SOURCES = ("ccrs_kernels.cu", "ccrs_lincommon.cuh", "ccrs_kernels.cuh", "ccrs_device.cuh", "ccrs_devutil.cuh", "ccrs_atan_tab.inc")


def _strip_comments(src: str) -> str:
    """C / C++ source without comments and with whitespace collapsed (string literals in these files hold no '//')."""
    import re
    src = re.sub(r"/\*.*?\*/", " ", src, flags=re.S)
    src = re.sub(r"//[^\n]*", " ", src)
    return " ".join(src.split())


def kernel_source_hash() -> str:
    """sha256 over the CODE (comments and whitespace stripped) of K2's translation unit and the headers it includes: ties
    profiles/k2_dram_bytes_per_launch.json to the kernel build that ran."""
    h = hashlib.sha256()
    d = os.path.join(ROOT, "camera-intrinsic-calibration-rs_b200", "csrc")
    for name in K2_SOURCES:
        h.update(name.encode())
        h.update(_strip_comments(open(os.path.join(d, name), encoding="utf-8").read()).encode())
    return h.hexdigest()[:16]


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)), "measured"
        except Exception:
            pass
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks + throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "20",
                                          "-i", str(self.gpu)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            t0 = time.time()
            while not self.rows and time.time() - t0 < 3.0:   # nvidia-smi needs ~0.5 s to produce its first row
                time.sleep(0.01)
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.05)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, col in (("hw_slowdown", 5), ("hw_thermal_slowdown", 6), ("sw_thermal_slowdown", 7), ("sw_power_cap", 8)):
                    if r[col].lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                continue
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def pinned(a: np.ndarray):
    """copy into page-locked host memory (torch is plumbing: allocator only)."""
    import torch
    dt = {np.dtype(np.float64): torch.float64, np.dtype(np.float32): torch.float32, np.dtype(np.int32): torch.int32}[a.dtype]
    t = torch.empty(a.shape, dtype=dt, pin_memory=True)
    n = t.numpy()
    n[...] = a
    return n, t


class Ranks:
    """torch.distributed plumbing: barrier, max over ranks, gather of small arrays."""

    def __init__(self):
        import torch
        self.torch = torch
        self.rank = int(os.environ.get("RANK", "0")); self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.dist = None
        if self.world > 1:
            import torch.distributed as dist
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
            self.dist = dist
        torch.cuda.set_device(self.local)
        self.dev = self.local

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.dist:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max(self, x: float) -> float:
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=f"cuda:{self.dev}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x: float) -> float:
        if not self.dist:
            return x
        t = self.torch.tensor([x], dtype=self.torch.float64, device=f"cuda:{self.dev}")
        self.dist.all_reduce(t, op=self.dist.ReduceOp.SUM)
        return float(t.item())

    def gather_bits(self, a: np.ndarray) -> np.ndarray:
        """[world, len(a)] int64 bit patterns of a float64 vector from every rank."""
        bits = np.ascontiguousarray(a, dtype=np.float64).view(np.int64)
        if not self.dist:
            return bits[None, :].copy()
        t = self.torch.from_numpy(bits.copy()).to(f"cuda:{self.dev}")
        out = [self.torch.empty_like(t) for _ in range(self.world)]
        self.dist.all_gather(out, t)
        return np.stack([o.cpu().numpy() for o in out])

    def finish(self):
        if self.dist:
            self.dist.barrier()
            self.dist.destroy_process_group()


def timed_lm_steps(R: Ranks, prob, intr0, poses0, steps, warmup, n_total, flush_l2=True):
    R.barrier()
    step_ms, launches = prob.bench_lm_steps(intr0, poses0, warmup=warmup, steps=steps, flush_l2=flush_l2)
    R.barrier()
    ms = R.max(float(step_ms.sum())) / steps
    return ms, n_total / (ms * 1e-3), launches


def sharded_problem(pkg, R: Ranks, s, model=MODEL):
    lo, hi = pkg.dist.shard_frames(s.frame_offsets, R.rank, R.world)
    sh = pkg.dist.slice_problem(s, lo, hi)
    prob = pkg.Problem(model, s.width, s.height, sh["frame_offsets"], sh["x"], sh["y"], sh["z"], sh["u"], sh["v"], device=R.dev)
    return prob, sh, lo, hi


def multi_gpu_check(pkg, R: Ranks, s):
    """Same run, same problem: every rank's converged intrinsics bitwise identical, equal iteration count and
    <= 1e-6 relative difference vs a single-rank solve on rank 0 — peer-memory exchange, then CCRS_P2P=0 (NCCL)."""
    out = {"ok": True, "tolerance_rel": 1e-6, "modes": {}}
    ref = {}
    if R.rank == 0:     # single-rank solves of the whole problem (no communicator attached)
        q = pkg.Problem.from_synth(s, device=R.dev)
        for loop in ("lm", "gn"):
            q.set_poses(s.init_poses)
            intr, summ, _ = (q.solve_lm if loop == "lm" else q.solve_gn)(s.init_params)
            ref[loop] = (intr.copy(), int(summ.iterations), int(summ.status))
        q.close()
    for mode in ("peer", "nccl"):
        os.environ["CCRS_P2P"] = "1" if mode == "peer" else "0"
        prob, sh, lo, hi = sharded_problem(pkg, R, s)
        pkg.dist.init_comm(prob, R.rank, R.world)        # new communicator: peer setup honours CCRS_P2P
        uses_peer = int(pkg._abi.load().ccrs_comm_uses_peer_memory())
        res = {"uses_peer_memory": uses_peer}
        for loop in (("lm", "gn") if mode == "peer" else ("lm",)):
            prob.set_poses(s.init_poses[lo:hi])
            intr, summ, _ = (prob.solve_lm if loop == "lm" else prob.solve_gn)(s.init_params)
            bits = R.gather_bits(intr)
            iters = R.gather_bits(np.array([float(summ.iterations), float(summ.status)]))
            bitwise = bool(np.all(bits == bits[0:1]))
            same_iters = bool(np.all(iters == iters[0:1]))
            r = {"bitwise_identical_across_ranks": bitwise, "iterations": int(summ.iterations), "same_iterations_on_all_ranks": same_iters,
                 "status": int(summ.status)}
            if R.rank == 0:
                ri, rit, rst = ref[loop]
                r["iterations_single_rank"] = rit
                r["max_rel_diff_vs_single_rank"] = float(np.max(np.abs(intr - ri) / np.abs(ri)))
                r["ok"] = bool(bitwise and same_iters and rit == int(summ.iterations) and summ.status == 0 and rst == 0 and
                               r["max_rel_diff_vs_single_rank"] <= 1e-6)
            else:
                r["ok"] = bool(bitwise and same_iters)
            res[loop] = r
        expect_peer = 1 if mode == "peer" else 0
        res["ok"] = all(res[k]["ok"] for k in res if isinstance(res[k], dict)) and (mode == "peer" or uses_peer == expect_peer)
        out["modes"][mode] = res
        prob.close()
    os.environ["CCRS_P2P"] = "1"
    ok = R.sum(0.0 if all(m["ok"] for m in out["modes"].values()) else 1.0) == 0.0
    out["ok"] = bool(ok)
    return out


def make_batch(pkg, n_problems: int, first: int = 0, n_distinct: int = 16):
    """n_problems independent KB4 calibrations (200 frames each) concatenated; 16 distinct synthetic problems tiled."""
    probs = [pkg.synth.make_calib(BATCH_MODEL, BATCH_FRAMES, seed=100 + i) for i in range(n_distinct)]
    fo, pfo = [np.zeros(1, dtype=np.int64)], [0]
    xs, ys, zs, us, vs, poses, intr0 = [], [], [], [], [], [], []
    for b in range(first, first + n_problems):
        s = probs[b % n_distinct]
        fo.append(fo[-1][-1] + s.frame_offsets[1:].astype(np.int64))
        pfo.append(pfo[-1] + s.n_frames)
        xs.append(s.x); ys.append(s.y); zs.append(s.z); us.append(s.u); vs.append(s.v)
        poses.append(s.init_poses); intr0.append(s.init_params)
    cat = np.concatenate
    return dict(fo=cat(fo).astype(np.int32), pfo=np.array(pfo, dtype=np.int32), x=cat(xs), y=cat(ys), z=cat(zs), u=cat(us),
                v=cat(vs), poses=cat(poses), intr0=np.stack(intr0), probs=probs)


def bench_batch(pkg, R: Ranks, steps: int, warmup: int):
    """configs[4]: BATCH_PER_RANK independent KB4 calibrations per rank, no communication. A step = one LM solve of
    the rank's batch to convergence (poses reset untimed)."""
    b = make_batch(pkg, BATCH_PER_RANK, first=R.rank * BATCH_PER_RANK)
    t0 = time.perf_counter()
    gp = pkg.Problem(BATCH_MODEL, 1024, 1024, b["fo"], b["x"], b["y"], b["z"], b["u"], b["v"], problem_frame_offsets=b["pfo"], device=R.dev)
    create_s = time.perf_counter() - t0
    n_obs = int(gp.n_obs)
    dev_ms, wall_ms, n_lin = [], [], 0
    launches0 = 0
    for i in range(warmup + steps):
        gp.set_poses(b["poses"])
        R.barrier()
        if i == warmup:
            launches0 = gp.launch_count()
        t0 = time.perf_counter()
        intr, summ, _ = gp.solve_lm(b["intr0"])
        w = (time.perf_counter() - t0) * 1e3
        if i >= warmup:
            dev_ms.append(summ.device_ms); wall_ms.append(w)
            n_lin = summ.iterations + 1
    launches = gp.launch_count() - launches0
    # a sample of problems against solving them on their own
    worst = 0.0
    for k in (0, 5, BATCH_PER_RANK - 1):
        s = b["probs"][(R.rank * BATCH_PER_RANK + k) % len(b["probs"])]
        q = pkg.Problem.from_synth(s, device=R.dev)
        q.set_poses(s.init_poses)
        ref, _, _ = q.solve_lm(s.init_params)
        worst = max(worst, float(np.max(np.abs(intr[k] - ref) / np.abs(ref))))
        q.close()
    gp.close()
    ms = R.max(float(np.median(dev_ms)))
    wall = R.max(float(np.median(wall_ms)))
    total_obs = R.sum(float(n_obs))
    return {"workload": f"{BATCH_PER_RANK * R.world} independent {BATCH_MODEL.upper()} calibrations x {BATCH_FRAMES} frames x 144 corners, "
                        f"{BATCH_PER_RANK} per GPU, no communication (BASELINE configs[4])",
            "n_gpus": R.world, "problems": BATCH_PER_RANK * R.world, "obs_total": int(total_obs),
            "lm_iterations_max": int(n_lin - 1), "linearisations": int(n_lin),
            "ms_per_batch_solve": ms, "wall_ms_per_batch_solve": wall,
            "evals_per_s": total_obs * n_lin / (ms * 1e-3), "calibrations_per_s": BATCH_PER_RANK * R.world / (ms * 1e-3),
            "evals_per_s_wall": total_obs * n_lin / (wall * 1e-3), "gpu_launches": int(launches),
            "max_rel_diff_vs_standalone_solve": R.max(worst), "create_s": create_s, "scaling": "weak (512 problems per GPU)"}


def run_ours(args):
    pkg = load_pkg()
    R = Ranks()
    rank, world, dev = R.rank, R.world, R.dev

    if args.workload == "batch":
        sampler = ClockSampler(dev)
        if rank == 0:
            sampler.start()
        bt = bench_batch(pkg, R, steps=max(3, min(args.steps, 5)), warmup=1)
        clocks = sampler.stop() if rank == 0 else None
        if rank == 0:
            print(json.dumps({"metric": "residual+Jacobian evals/s, batch of independent KB4 calibrations", "value": bt["evals_per_s"],
                              "unit": "evals/s", "n_gpus": world, "steps": max(3, min(args.steps, 5)), "warmup": 1,
                              "ms_per_step": bt["ms_per_batch_solve"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                              "dtype": "f64", "data": "synthetic", "impl": "ours", "config": {"workload": bt["workload"]},
                              "batch": bt, "gpu_launches": bt["gpu_launches"], "clocks": clocks}))
        R.finish()
        return

    # ---- the problem: 7000 frames in total, this rank's contiguous shard (strong scaling) ----------------------
    s = pkg.synth.make_calib(MODEL, FRAMES_TOTAL, seed=3)
    prob, sh, lo, hi = sharded_problem(pkg, R, s)
    poses0 = np.ascontiguousarray(s.init_poses[lo:hi])
    n_local = int(sh["frame_offsets"][-1]); n_total = s.n_obs
    pkg.dist.init_comm(prob, rank, world)
    d = prob.d
    exch = "none (single GPU)" if world == 1 else ("fused into the kernels over peer memory (NVLink P2P stores, rank-order sum)"
                                                 if pkg._abi.load().ccrs_comm_uses_peer_memory() else "NCCL all-gather + rank-order sum")

    # ---- device-resident steps (value): replicas of the shard, together >= 2.5 x L2, visited round-robin ----------
    import torch
    l2_bytes = int(getattr(torch.cuda.get_device_properties(dev), "L2_cache_size", 126 << 20))
    bytes_per_replica = 40 * n_local + (hi - lo) * 8 * (2 * prob.nblk + 6 * d + 18 + 6 + 12)   # observations + blocks + elimination record + poses
    n_rep = int(min(48, max(3, -(-int(2.5 * l2_bytes) // bytes_per_replica))))
    replicas = [prob]
    for _ in range(n_rep - 1):
        q, _, _, _ = sharded_problem(pkg, R, s)
        if world > 1:
            q.comm_init(None)
        replicas.append(q)
    sampler = ClockSampler(dev)
    R.barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    pkg.Problem.bench_lm_steps_rotating(replicas, s.init_params, poses0, warmup=args.warmup, steps=args.steps)   # page-in, pools
    R.barrier()
    total_ms, launches, executed = pkg.Problem.bench_lm_steps_rotating(replicas, s.init_params, poses0, warmup=args.warmup, steps=args.steps)
    R.barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    if executed <= 0:
        raise RuntimeError("no timed LM iteration executed")
    ms_per_step = R.max(total_ms) / executed          # `executed` is the same on every rank (bitwise-identical decisions)
    value = n_total / (ms_per_step * 1e-3)
    for q in replicas[1:]:
        q.close()

    # the isolated step of rounds 1-2: L2 flushed (512 MB write) before every step, one event bracket per step
    iso_ms_per_step, iso_value, _ = timed_lm_steps(R, prob, s.init_params, poses0, args.steps, args.warmup, n_total, flush_l2=True)

    # ---- the LM loop as a caller runs it: no synchronisation or flush between iterations
    prob.set_poses(poses0)
    o = pkg.default_options(max_iteration=40, min_abs_decrease=-1.0, min_rel_decrease=-1.0, min_error=-1.0)
    import ctypes as C
    lib = pkg._abi.load()
    lib.ccrs_loop_trace(1, None, None)          # device stamps carried by the iteration records (globaltimer)
    _, loop_summ, _ = prob.solve_lm(s.init_params, options=o)
    tr_avg = (C.c_double * 13)(); tr_cnt = C.c_int64(0)
    lib.ccrs_loop_trace(0, tr_avg, C.byref(tr_cnt))
    loop_trace = {"iterations_traced": int(tr_cnt.value), "k2_us": tr_avg[0], "k2_to_k3_us": tr_avg[1], "k3_per_frame_us": tr_avg[2],
                  "k3_tail_us": tr_avg[3], "k3_to_k2_us": tr_avg[4]} if tr_cnt.value > 0 else None
    loop_ms = R.max(loop_summ.device_ms) / max(loop_summ.iterations, 1)
    # ... and the loop the reference actually runs (Gauss-Newton, src/util.rs:443-458)
    prob.set_poses(poses0)
    _, gn_summ, _ = prob.solve_gn(s.init_params, options=pkg.default_options(max_iteration=20, min_abs_decrease=-1.0, min_rel_decrease=-1.0, min_error=-1.0))
    gn_ms = R.max(gn_summ.device_ms) / max(gn_summ.iterations, 1)

    # ---- dominant kernel K2 alone, timed live with CUDA events on the handle's stream ------------------
    prob.set_poses(poses0)
    k2_ms = R.max(prob.time_linearize(s.init_params, reps=max(10, args.steps), flush_l2=True))
    k2_ms_warm = R.max(prob.time_linearize(s.init_params, reps=max(10, args.steps), flush_l2=False))
    peaks, peak_kind = measured_peaks()
    nblk = prob.nblk
    bytes_per_obs = 40.0 + (48.0 + 8.0 * nblk) / 144.0            # B_obs = 40 + (48 + 8 n_blk)/144
    n_k2 = R.max(float(n_local))                                   # observations of the launch that was timed (largest shard)
    achieved_gbs = bytes_per_obs * n_k2 / (k2_ms * 1e-3) / 1e9
    fp64_peak = pkg.measure_fp64_peak(dev)
    tf_nominal = FLOP_PER_OBS_NOMINAL * n_k2 / (k2_ms * 1e-3) / 1e12
    tf_counted = FLOP_PER_OBS_COUNTED * n_k2 / (k2_ms * 1e-3) / 1e12
    roofline = {"kernel": "k_linearize<EUCM> (K2)", "bound": "fp64", "achieved": round(tf_nominal, 2), "peak": round(fp64_peak, 2),
                "unit": "TFLOP/s", "frac": round(tf_nominal / fp64_peak, 4), "traffic": None,
                "flop_per_obs": FLOP_PER_OBS_NOMINAL, "flop_source": "SURVEY §8(d) nominal: 360 normal-equation + ~200 model",
                "peak_source": "ccrs_measure_fp64_peak DFMA microbenchmark in this run (MEASURED_PEAKS.json has no FP64 entry; datasheet 37)",
                "counted": {"flop_per_obs": FLOP_PER_OBS_COUNTED, "achieved": round(tf_counted, 2), "frac": round(tf_counted / fp64_peak, 4),
                            "what": "FP64 instructions counted in the kernel's SASS (DFMA=2, DMUL/DADD=1), DESIGN §3"},
                "k2_ms": round(k2_ms, 5), "k2_ms_l2_warm": round(k2_ms_warm, 5), "obs_per_launch": int(n_k2),
                "in_loop": ({"k2_ms": round(loop_trace["k2_us"] * 1e-3, 5),
                             "frac": round(FLOP_PER_OBS_NOMINAL * n_local / (loop_trace["k2_us"] * 1e-6) / 1e12 / fp64_peak, 4),
                             "what": "K2 inside the device-driven LM loop of this rank (L2 warm, device globaltimer stamps: first warp past its dependency wait -> last warp done, no launch latency); `frac` above is the isolated, L2-flushed launch"}
                            if loop_trace else None),
                "hbm": {"achieved": round(achieved_gbs, 1), "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": round(achieved_gbs / peaks["hbm_gbs"], 4),
                        "bytes_per_obs": round(bytes_per_obs, 2), "peak_source": f"{peak_kind} MEASURED_PEAKS.json hbm_gbs (burst: kernel timed alone)"},
                "note": "K2 (EUCM) is bound by the FP64 pipe (CUDA cores: the packed sparse Gram update beats tensor tiles at 13 columns); the HBM fraction is secondary"}
    traffic_file = os.path.join(ROOT, "profiles", "k2_dram_bytes_per_launch.json")
    if os.path.exists(traffic_file) and world == 1:
        try:
            tj = json.load(open(traffic_file))
            if tj.get("kernel_src_sha256") == kernel_source_hash():     # only for the kernel build that was profiled
                roofline["traffic"] = tj.get("bytes_per_launch")
            else:
                roofline["traffic_note"] = "ncu capture is of another kernel build (source hash differs): not reported"
        except Exception:
            pass

    # ---- weak-scaling figure (7000 frames PER GPU, one problem of N x 7000 frames) ------------------------------
    weak = None
    if world > 1 and not args.no_weak:
        sw = pkg.synth.make_calib(MODEL, FRAMES_TOTAL * world, seed=3)
        pw, shw, low, hiw = sharded_problem(pkg, R, sw)
        pw.comm_init(None)
        reps_w = [pw]
        n_rep_w = int(min(48, max(3, -(-int(2.5 * l2_bytes) // (40 * int(shw["frame_offsets"][-1]) + (hiw - low) * 8 * (2 * pw.nblk + 6 * d + 36))))))
        for _ in range(n_rep_w - 1):
            q, _, _, _ = sharded_problem(pkg, R, sw)
            q.comm_init(None)
            reps_w.append(q)
        poses_w = np.ascontiguousarray(sw.init_poses[low:hiw])
        R.barrier()
        pkg.Problem.bench_lm_steps_rotating(reps_w, sw.init_params, poses_w, warmup=args.warmup, steps=args.steps)
        R.barrier()
        w_total, _, w_exec = pkg.Problem.bench_lm_steps_rotating(reps_w, sw.init_params, poses_w, warmup=args.warmup, steps=args.steps)
        R.barrier()
        w_ms = R.max(w_total) / max(w_exec, 1)
        w_value = sw.n_obs / (w_ms * 1e-3)
        weak = {"frames_per_gpu": FRAMES_TOTAL, "obs_total": int(sw.n_obs), "ms_per_step": w_ms, "value": w_value, "lm_iterations_per_s": 1e3 / w_ms,
                "replicas": n_rep_w, "what": "one problem of N x 7000 frames, frame-sharded; same rotating-replica method as `value`"}
        for q in reps_w[1:]:
            q.close()
        pw.close()
        del sw, shw

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------
    # Host data = what the reference holds (src/detected_points.rs:6-17): f32 p2d per detected corner keyed by corner id,
    # p3d = the board point of that id (src/board.rs:46-95). Two host formats are timed: `board` (corner ids + board table,
    # ccrs_problem_create_board_f32: 12 B per observation over PCIe) is the headline; `xyz` (x, y, z, u, v f32 arrays,
    # ccrs_problem_create_f32: 20 B per observation) is the format of round 1.
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    a0, b0 = int(s.frame_offsets[lo]), int(s.frame_offsets[hi])
    hx, _tx = pinned(f32(sh["x"])); hy, _ty = pinned(f32(sh["y"])); hz, _tz = pinned(f32(sh["z"])); hu, _tu = pinned(f32(sh["u"])); hv, _tv = pinned(f32(sh["v"]))
    hid, _ti = pinned(np.ascontiguousarray(s.extra["corner_id"][a0:b0], dtype=np.int32))
    hboard = np.ascontiguousarray(s.extra["board"], dtype=np.float32)
    hfo, _tf = pinned(sh["frame_offsets"]); hp, _tp = pinned(poses0)
    hout, _to = pinned(np.zeros_like(poses0))          # page-locked result buffer for the D2H read
    e2e_steps = max(3, min(args.steps, 10))

    def e2e_run(fmt):
        evals, times = 0, []
        for i in range(2 + e2e_steps):
            R.barrier()
            t0 = time.perf_counter()
            if fmt == "board":
                q = pkg.Problem(MODEL, s.width, s.height, hfo, None, None, None, hu, hv, device=dev, corner_id=hid, board=hboard)
            else:
                q = pkg.Problem(MODEL, s.width, s.height, hfo, hx, hy, hz, hu, hv, device=dev)
            if world > 1:
                q.comm_init(None)
            q.set_poses(hp)
            intr, summ, _ = q.solve_lm(s.init_params)
            out_poses = q.get_poses(out=hout)
            q.close()
            dt = time.perf_counter() - t0   # this rank's call; the ranks leave solve_lm together (every iteration exchanges)
            R.barrier()
            if i >= 2:
                times.append(dt)
                evals += n_total * (1 + summ.iterations)     # initial linearisation + one (speculative) per iteration
        total = R.max(float(np.sum(times)))
        return evals / total, total / e2e_steps * 1e3, intr, summ

    def e2e_reuse():
        """the same calibration on a handle that is kept: new detections -> ccrs_problem_update_observations"""
        q = pkg.Problem(MODEL, s.width, s.height, hfo, None, None, None, hu, hv, device=dev, corner_id=hid, board=hboard)
        if world > 1:
            q.comm_init(None)
        evals, times = 0, []
        for i in range(2 + e2e_steps):
            R.barrier()
            t0 = time.perf_counter()
            q.update_observations(hfo, hu, hv, corner_id=hid)
            q.set_poses(hp)
            _, sm, _ = q.solve_lm(s.init_params)
            q.get_poses(out=hout)
            dt = time.perf_counter() - t0
            R.barrier()
            if i >= 2:
                times.append(dt)
                evals += n_total * (1 + sm.iterations)
        q.close()
        total = R.max(float(np.sum(times)))
        return evals / total, total / e2e_steps * 1e3

    e2e_value, e2e_ms, intr, summ = e2e_run("board")
    e2e_reuse_value, e2e_reuse_ms = e2e_reuse()
    e2e_xyz_value, e2e_xyz_ms, _, _ = e2e_run("xyz")
    h2d = hid.nbytes + hu.nbytes + hv.nbytes + hboard.nbytes + hfo.nbytes + hp.nbytes
    h2d_xyz = hx.nbytes * 5 + hfo.nbytes + hp.nbytes
    d2h = hp.nbytes + d * 8
    d2h_iter = (summ.iterations + 1) * 208 * 8          # one iteration record per executed reduction
    rel_err = float(np.max(np.abs(intr - s.gt_params) / np.abs(s.gt_params)))
    prob.close()

    # ---- batch of independent KB4 calibrations (configs[4]), 512 per rank -------------------------------------
    batch = None
    if not args.no_batch:
        batch = bench_batch(pkg, R, steps=3, warmup=1)

    # ---- multi-GPU correctness in the same run ------------------------------------------------------------------
    check = multi_gpu_check(pkg, R, s) if world > 1 else None

    # ---- CPU baseline (oracle port, bounded sample) on rank 0 at N=1 ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_lm_iteration_rate(pkg, s, sample_frames=FRAMES_TOTAL, iters=3, solves=5)

    if rank == 0:
        cfg = workload_config(n_total)
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "ours", "config": cfg,
            "run": {"frames_per_gpu": int(hi - lo), "obs_per_gpu": int(n_local), "parallelism": f"frame-sharded x{world}", "exchange": exch},
            "lm_iterations_per_s": 1e3 / ms_per_step,
            "replicas": {"n": n_rep, "bytes_per_replica": int(bytes_per_replica), "l2_bytes": l2_bytes},
            "steps_executed": int(executed),
            "isolated_step_l2_flushed": {"ms_per_step": iso_ms_per_step, "value": iso_value,
                                         "what": "rounds 1-2 method: 512 MB flush before every step, one CUDA-event bracket per step, host synchronisation between steps"},
            "lm_loop_l2_warm": {"ms_per_iteration": loop_ms, "iterations": int(loop_summ.iterations), "lm_iterations_per_s": 1e3 / loop_ms,
                                "device_phases_us": loop_trace,
                                "value": n_total / (loop_ms * 1e-3),
                                "what": "ccrs_solve_lm with the stop tests disabled, 40 back-to-back iterations incl. the initial linearisation and Jacobi scaling, CUDA events around the whole loop"},
            "gn_loop_l2_warm": {"ms_per_iteration": gn_ms, "iterations": int(gn_summ.iterations),
                                "what": "ccrs_solve_gn (the loop the reference runs, util.rs:443-458), stop tests disabled, 20 iterations, CUDA events around the whole loop"},
            "wall_ms_bench_calls": wall_ms,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h + d2h_iter),
                    "ms_per_call": e2e_ms, "lm_iterations_per_call": int(summ.iterations),
                    "what": "ccrs_problem_create_board_f32 (H2D from pinned memory of corner ids + f32 p2d + the board table: the reference's FrameFeature / Board data model) + set_poses + ccrs_solve_lm to convergence + get_poses (D2H) + destroy",
                    "converged_rel_err_vs_gt": rel_err,
                    "reused_handle": {"value": e2e_reuse_value, "ms_per_call": e2e_reuse_ms, "h2d_bytes_per_step": int(h2d - hboard.nbytes),
                                      "what": "the handle is kept between calibrations: ccrs_problem_update_observations (H2D of corner ids + f32 p2d) + set_poses + ccrs_solve_lm to convergence + get_poses (D2H)"},
                    "xyz_f32_format": {"value": e2e_xyz_value, "ms_per_call": e2e_xyz_ms, "h2d_bytes_per_step": int(h2d_xyz),
                                       "what": "same call sequence through ccrs_problem_create_f32 (x, y, z, u, v f32 arrays: round 1's format)"}},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        }
        if weak is not None:
            line["weak_scaling"] = weak
        if batch is not None:
            line["batch"] = batch
        if check is not None:
            line["multi_gpu_check"] = check
        print(json.dumps(line))
    R.finish()


def cpu_lm_iteration_rate(pkg, s, sample_frames: int, iters: int, solves: int = 5, threads: int | None = None):
    """Oracle LM iterations on the host cores: the reference arm / cpu_baseline. Median over `solves` solves of `iters`
    LM iterations each (after one warm-up solve). Returns the cpu_baseline object."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle as O
    threads = threads or (os.cpu_count() or 1)
    sub = pkg.dist.slice_problem(s, 0, sample_frames)
    op = O.OracleProblem(pkg.MODELS[MODEL], s.width, s.height, sub["frame_offsets"], sub["x"], sub["y"], sub["z"], sub["u"], sub["v"],
                         n_threads=threads)
    n = int(sub["frame_offsets"][-1])
    poses = s.init_poses[:sample_frames]
    opt = op.default_options(max_iteration=1, min_abs_decrease=-1.0, min_rel_decrease=-1.0, min_error=-1.0)
    op.levenberg_marquardt(s.init_params, poses, options=opt)          # warm-up (thread pool, page faults)
    opt = op.default_options(max_iteration=iters, min_abs_decrease=-1.0, min_rel_decrease=-1.0, min_error=-1.0)
    per_iter = []
    for _ in range(solves):
        t0 = time.perf_counter()
        _, _, res, _ = op.levenberg_marquardt(s.init_params, poses, options=opt)
        # one oracle LM iteration = one dual-number linearisation + one residual-only pass; the initial cost pass is amortised
        per_iter.append((time.perf_counter() - t0) / res.iterations)
    med = float(np.median(per_iter))
    return {"value": n / med, "unit": "evals/s", "cores": threads, "kind": "port",
            "sample": f"median of {solves} solves x {iters} LM iterations of the same EUCM problem restricted to {sample_frames} frames ({n} obs), OpenMP {threads} threads",
            "ms_per_lm_iteration": med * 1e3, "ms_per_lm_iteration_min_max": [float(np.min(per_iter)) * 1e3, float(np.max(per_iter)) * 1e3],
            "lm_iterations_per_s": 1.0 / med}


def run_reference(args):
    """Reference arm: the reference's own algorithm for this path on the host cores. The Rust reference cannot be
    compiled here (no cargo/rustc; tiny-solver / camera-intrinsic-model / num-dual are not vendored), so this is the
    oracle port (oracle/ccrs_oracle.cpp: dual-number autodiff + Huber corrector + Cholesky), all host threads.
    A step = one LM iteration (each timed solve runs 3 of them); the problem is the whole 7000-frame problem at every N."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    pkg = load_pkg()
    s = pkg.synth.make_calib(MODEL, FRAMES_TOTAL, seed=3)
    base = cpu_lm_iteration_rate(pkg, s, sample_frames=FRAMES_TOTAL, iters=3, solves=max(5, args.steps))
    ms = base["ms_per_lm_iteration"]
    n = s.n_obs
    value = n / (ms * 1e-3)
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong" if world > 1 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "impl": "reference", "config": workload_config(n),
        "lm_iterations_per_s": 1e3 / ms, "cpu_baseline": base,
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="lm", choices=["lm", "batch"], help="lm: configs[3] (headline); batch: configs[4] as a line of its own")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-batch", action="store_true", help="skip the batch (configs[4]) extra key")
    ap.add_argument("--no-weak", action="store_true", help="skip the weak-scaling extra key at N > 1")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

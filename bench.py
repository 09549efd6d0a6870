#!/usr/bin/env python
"""bench.py — BASELINE.json's metric on its quoted configuration.

metric   : residual+Jacobian evals/s (and LM iterations/s) at ~1M corner observations, EUCM
workload : configs[3] — EUCM, 7,000 frames x 144 corners (1,007,999 obs after the image-bounds filter) PER GPU,
           synthetic (SURVEY.md §8(d) generator, seed 3), frames sharded across ranks, one exchange of the reduced
           intrinsic system per linearisation (fused into K3/K2: peer-memory stores over NVLink + rank-order sum in
           the kernel's last CTA; NCCL all-gather + rank-order sum when peer memory is unavailable). Weak scaling: every
           rank owns 7,000 frames, the job is one calibration problem of N x 7,000 frames.
step     : one Levenberg-Marquardt iteration of that problem = K3 reduce (+exchange) -> host d x d solve -> K4
           back-substitution -> K2 linearisation of the trial point (speculative LM: the trial cost comes from the
           linearisation itself) -> exchange of (model decrease, cost) -> accept/reject. Stop tests are disabled so
           each of the K timed steps does the full work; every step evaluates residual+Jacobian once per observation.
           Every 4 steps the state returns (untimed) to the perturbed start, so the timed steps are LM iterations 1-4
           of the problem (iteration 1 has the Huber loss active on ~99% of the observations).
value    : total observations over all ranks / time per step, inputs resident in HBM, L2 flushed (512 MB write)
           before every timed step outside the CUDA-event bracket; max over ranks.
e2e      : the same metric through the C-ABI entry point a user calls with HOST buffers: per step one complete
           ccrs_problem_create_f32 (H2D of the f32 observation arrays from pinned memory) + ccrs_set_poses + ccrs_solve_lm
           to convergence + ccrs_get_poses (D2H) + destroy; evals = observations x linearisations performed.

`--impl reference` times the reference arm: the CPU oracle (oracle/, a restatement of the reference's num-dual +
tiny-solver path: the Rust reference cannot be built here) running the same LM iteration on the host cores.
"""
from __future__ import annotations

import argparse
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

WORKLOAD = "EUCM 7000 frames x 144 corners per GPU (~1.008M obs/GPU), 1024x1024, synthetic seed 3, Huber(1.0), LM iteration"
METRIC = "residual+Jacobian evals/s at 1M corner obs (EUCM LM iteration)"
FRAMES_PER_GPU = 7000
MODEL = "eucm"


def load_pkg():
    return importlib.import_module("camera-intrinsic-calibration-rs_b200")


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
            self.n_before = len(self.rows)
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


def run_ours(args):
    import torch
    import torch.distributed as dist
    pkg = load_pkg()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = local

    # ---- synthetic problem: N x 7000 frames, this rank's contiguous shard -------------------------------
    s = pkg.synth.make_calib(MODEL, FRAMES_PER_GPU * world, seed=3)
    lo, hi = pkg.dist.shard_frames(s.frame_offsets, rank, world)
    sh = pkg.dist.slice_problem(s, lo, hi)
    poses0 = np.ascontiguousarray(s.init_poses[lo:hi])
    n_local = int(sh["frame_offsets"][-1]); n_total = s.n_obs
    prob = pkg.Problem(MODEL, s.width, s.height, sh["frame_offsets"], sh["x"], sh["y"], sh["z"], sh["u"], sh["v"], device=dev)
    pkg.dist.init_comm(prob, rank, world)
    d = prob.d
    exch = "none (single GPU)" if world == 1 else ("fused into K2/K3 over peer memory (NVLink P2P stores, rank-order sum)" if pkg._abi.load().ccrs_comm_uses_peer_memory()
                                                 else "NCCL all-gather + rank-order sum")

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident steps (value) --------------------------------------------------------------------
    sampler = ClockSampler(dev)
    barrier()
    if rank == 0:
        sampler.start()
    t_wall0 = time.perf_counter()
    step_ms, launches = prob.bench_lm_steps(s.init_params, poses0, warmup=args.warmup, steps=args.steps, flush_l2=True)
    barrier()
    wall_ms = (time.perf_counter() - t_wall0) * 1e3
    clocks = sampler.stop() if rank == 0 else None
    total_ms = max_over_ranks(float(step_ms.sum()))
    ms_per_step = total_ms / args.steps
    value = n_total / (ms_per_step * 1e-3)

    # L2-warm variant (what a real LM loop sees: the observation arrays stay L2-resident between iterations)
    barrier()
    warm_ms, _ = prob.bench_lm_steps(s.init_params, poses0, warmup=args.warmup, steps=args.steps, flush_l2=False)
    warm_ms_per_step = max_over_ranks(float(warm_ms.sum())) / args.steps

    # ---- the LM loop as a caller runs it: no synchronisation or flush between iterations, so the speculative K3 launch
    #      overlaps the host's bookkeeping (the per-step figure above isolates every iteration between stream syncs)
    prob.set_poses(poses0)
    o = pkg.default_options(max_iteration=40, min_abs_decrease=-1.0, min_rel_decrease=-1.0, min_error=-1.0)
    _, loop_summ, _ = prob.solve_lm(s.init_params, options=o)
    loop_ms = max_over_ranks(loop_summ.device_ms) / max(loop_summ.iterations, 1)

    # ---- dominant kernel K2 alone, timed live with CUDA events on the handle's stream ------------------
    prob.set_poses(poses0)
    k2_ms = prob.time_linearize(s.init_params, reps=max(10, args.steps), flush_l2=True)
    k2_ms_warm = prob.time_linearize(s.init_params, reps=max(10, args.steps), flush_l2=False)
    peaks, peak_kind = measured_peaks()
    nblk = prob.nblk
    bytes_per_obs = 40.0 + (48.0 + 8.0 * nblk) / 144.0            # B_obs = 40 + (48 + 8 n_blk)/144
    flop_per_obs = 560.0                                           # 360 (normal equations) + ~200 (model)
    achieved_gbs = bytes_per_obs * n_local / (k2_ms * 1e-3) / 1e9
    fp64_peak = pkg.measure_fp64_peak(dev)
    achieved_tf = flop_per_obs * n_local / (k2_ms * 1e-3) / 1e12
    roofline = {"kernel": "k_linearize<EUCM> (K2)", "bound": "hbm", "achieved": round(achieved_gbs, 1), "peak": peaks["hbm_gbs"],
                "unit": "GB/s", "frac": round(achieved_gbs / peaks["hbm_gbs"], 4), "traffic": None, "peak_source": f"{peak_kind} MEASURED_PEAKS.json hbm_gbs",
                "k2_ms": round(k2_ms, 5), "k2_ms_l2_warm": round(k2_ms_warm, 5), "bytes_per_obs": round(bytes_per_obs, 2),
                "note": "K2 is FP64-CUDA-core-bound by design (no dense contraction, no tensor cores): see fp64",
                "fp64": {"achieved": round(achieved_tf, 2), "peak": round(fp64_peak, 2), "unit": "TFLOP/s",
                         "frac": round(achieved_tf / fp64_peak, 4), "flop_per_obs": flop_per_obs,
                         "peak_source": "ccrs_measure_fp64_peak DFMA microbenchmark, same run"}}
    traffic_file = os.path.join(ROOT, "profiles", "k2_dram_bytes_per_launch.json")
    if os.path.exists(traffic_file):
        try:
            roofline["traffic"] = json.load(open(traffic_file)).get("bytes_per_launch")
        except Exception:
            pass

    # ---- end to end through the C ABI with host buffers ---------------------------------------------------
    # the reference's FeaturePoint holds f32 (src/detected_points.rs:6-9): the f32 entry point is the natural host format
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    hx, _tx = pinned(f32(sh["x"])); hy, _ty = pinned(f32(sh["y"])); hz, _tz = pinned(f32(sh["z"])); hu, _tu = pinned(f32(sh["u"])); hv, _tv = pinned(f32(sh["v"]))
    hfo, _tf = pinned(sh["frame_offsets"]); hp, _tp = pinned(poses0)
    h2d = hx.nbytes * 5 + hfo.nbytes + hp.nbytes
    d2h = hp.nbytes + d * 8
    e2e_steps = max(3, min(args.steps, 10))
    evals = 0
    e2e_times = []
    for i in range(2 + e2e_steps):
        barrier()
        t0 = time.perf_counter()
        q = pkg.Problem(MODEL, s.width, s.height, hfo, hx, hy, hz, hu, hv, device=dev)
        if world > 1:
            q.comm_init(None)
        q.set_poses(hp)
        intr, summ, _ = q.solve_lm(s.init_params)
        out_poses = q.get_poses()
        n_lin = 1 + summ.iterations          # initial linearisation + one (speculative) per iteration
        d2h_iter = summ.iterations * (prob.nout + 2) * 8
        q.close()
        barrier()
        if i >= 2:
            e2e_times.append(time.perf_counter() - t0)
            evals += n_total * n_lin
    e2e_total = max_over_ranks(float(np.sum(e2e_times)))
    e2e_value = evals / e2e_total
    rel_err = float(np.max(np.abs(intr - s.gt_params) / np.abs(s.gt_params)))

    # ---- CPU baseline (oracle port, bounded sample) on rank 0 at N=1 ---------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        cpu = cpu_lm_iteration_rate(pkg, s, sample_frames=FRAMES_PER_GPU, iters=2)

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "impl": "ours",
            "config": {"workload": WORKLOAD, "camera_model": MODEL, "frames_per_gpu": FRAMES_PER_GPU, "obs_total": int(n_total),
                       "obs_per_gpu": int(n_local), "parallelism": f"frame-sharded x{world}", "exchange": exch,
                       "l2": "flushed (512 MB write) before every timed step, outside the event bracket", "loop": "speculative LM"},
            "lm_iterations_per_s": 1e3 / ms_per_step,
            "l2_warm": {"ms_per_step": warm_ms_per_step, "value": n_total / (warm_ms_per_step * 1e-3), "lm_iterations_per_s": 1e3 / warm_ms_per_step},
            "lm_loop_l2_warm": {"ms_per_iteration": loop_ms, "iterations": int(loop_summ.iterations), "lm_iterations_per_s": 1e3 / loop_ms,
                                "value": n_total / (loop_ms * 1e-3),
                                "what": "ccrs_solve_lm with the stop tests disabled, 40 back-to-back iterations incl. the initial linearisation and Jacobi scaling, CUDA events around the whole loop"},
            "wall_ms_timed_region_incl_flush": wall_ms,
            "e2e": {"value": e2e_value, "unit": "evals/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h + d2h_iter),
                    "ms_per_call": e2e_total / e2e_steps * 1e3, "lm_iterations_per_call": int(summ.iterations),
                    "what": "ccrs_problem_create_f32 (H2D of the f32 FeaturePoint arrays from pinned memory) + set_poses + ccrs_solve_lm to convergence + get_poses (D2H) + destroy",
                    "converged_rel_err_vs_gt": rel_err},
            "gpu_launches": int(launches), "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
        }
        print(json.dumps(line))
    prob.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def cpu_lm_iteration_rate(pkg, s, sample_frames: int, iters: int, threads: int | None = None):
    """Oracle LM iterations on the host cores: the reference arm / cpu_baseline. Returns the cpu_baseline object."""
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
    t0 = time.perf_counter()
    _, _, res, _ = op.levenberg_marquardt(s.init_params, poses, options=opt)
    dt = time.perf_counter() - t0
    # one oracle LM iteration = one dual-number linearisation + one residual-only pass; the initial cost pass is amortised
    per_iter = dt / res.iterations
    return {"value": n / per_iter, "unit": "evals/s", "cores": threads, "kind": "port",
            "sample": f"{res.iterations} LM iterations of the same EUCM problem restricted to {sample_frames} frames ({n} obs), OpenMP {threads} threads",
            "ms_per_lm_iteration": per_iter * 1e3, "lm_iterations_per_s": 1.0 / per_iter}


def run_reference(args):
    """Reference arm: the reference's own algorithm for this path on the host cores. The Rust reference cannot be
    compiled here (no cargo/rustc; tiny-solver / camera-intrinsic-model / num-dual are not vendored), so this is the
    oracle port (oracle/ccrs_oracle.cpp: dual-number autodiff + Huber corrector + Cholesky), all host threads."""
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    pkg = load_pkg()
    s = pkg.synth.make_calib(MODEL, FRAMES_PER_GPU, seed=3)
    times, base = [], None
    for i in range(args.warmup + args.steps):
        base = cpu_lm_iteration_rate(pkg, s, sample_frames=FRAMES_PER_GPU, iters=1)
        if i >= args.warmup:
            times.append(base["ms_per_lm_iteration"])
    ms = float(np.mean(times))
    n = s.n_obs
    value = n / (ms * 1e-3)
    base.update({"value": value, "ms_per_lm_iteration": ms, "lm_iterations_per_s": 1e3 / ms})
    print(json.dumps({
        "metric": METRIC, "value": value, "unit": "evals/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "impl": "reference",
        "config": {"workload": WORKLOAD, "camera_model": MODEL, "frames_per_gpu": FRAMES_PER_GPU, "obs_total": int(n),
                   "note": "CPU arm runs ONE rank's 7000-frame problem on the host cores regardless of N"},
        "lm_iterations_per_s": 1e3 / ms, "cpu_baseline": base,
        "e2e": {"value": value, "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else max(args.warmup, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

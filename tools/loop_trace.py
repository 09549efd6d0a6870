"""Device-side phase trace of the device-driven loop (ccrs_loop_trace): where an LM / GN iteration goes, from the
globaltimer stamps the kernels leave in the iteration records. Usage: python tools/loop_trace.py [frames] [model]
Under torchrun it traces the frame-sharded problem (rank 0 prints)."""
import ctypes as C, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ccrs_b200 as c

n = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
model = sys.argv[2] if len(sys.argv) > 2 else "eucm"
rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
if world > 1:
    import torch, torch.distributed as dist
    torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
    dist.init_process_group("nccl", device_id=torch.device("cuda", int(os.environ["LOCAL_RANK"])))
s = c.synth.make_calib(model, n, seed=3)
lo, hi = c.dist.shard_frames(s.frame_offsets, rank, world)
sh = c.dist.slice_problem(s, lo, hi)
gp = c.Problem(model, s.width, s.height, sh["frame_offsets"], sh["x"], sh["y"], sh["z"], sh["u"], sh["v"], device=int(os.environ.get("LOCAL_RANK", "0")))
c.dist.init_comm(gp, rank, world)
lib = c._abi.load()
names = ["K2 (first warp past its wait -> last warp done)", "K2 done -> K3 last CTA past its wait", "K3 per-frame elimination + CTA sums",
         "K3 tail: cross-CTA sum, exchange, controller rule", "record ready -> next K2 running"]
fine = ["control block + decision", "block load", "per-frame elimination", "CTA sum + partial store", "cross-CTA sum (+ exchange)", "staging", "controller rule"]
out = {"frames_total": n, "frames_this_rank": int(hi - lo), "model": model, "n_gpus": world}
for loop in ("lm", "gn"):
    o = c.default_options(max_iteration=40, min_abs_decrease=-1.0, min_rel_decrease=-1.0, min_error=-1.0)
    solve = gp.solve_lm if loop == "lm" else gp.solve_gn
    gp.set_poses(s.init_poses[lo:hi]); solve(s.init_params, options=o)          # warm-up
    lib.ccrs_loop_trace(1, None, None)
    gp.set_poses(s.init_poses[lo:hi])
    _, summ, _ = solve(s.init_params, options=o)
    avg = (C.c_double * 13)(); cnt = C.c_int64(0)
    lib.ccrs_loop_trace(0, avg, C.byref(cnt))
    out[loop] = {"iterations_traced": int(cnt.value), "device_ms_per_iteration_events": summ.device_ms / max(summ.iterations, 1),
                 "phases_us": {names[i]: round(avg[i], 2) for i in range(5)}, "sum_us": round(sum(avg[:5]), 2),
                 "k3_last_cta_us": {fine[i]: round(avg[5 + i], 2) for i in range(7)},
                 "record_ready_to_k2_past_its_wait_us": round(avg[12], 2)}
if rank == 0:
    print(json.dumps(out, indent=1))
gp.close()
if world > 1:
    dist.barrier(); dist.destroy_process_group()

"""First-contact GPU probe: FP64 peak, K2 timing at the BASELINE sizes, a full LM solve. Prints JSON lines."""
import json, sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ccrs_b200 as c

print(json.dumps({"fp64_peak_tflops": c.measure_fp64_peak(0)}))
for model, nf in [("eucm", 100), ("eucm", 2000), ("eucm", 7000), ("kb4", 7000), ("opencv5", 7000)]:
    s = c.synth.make_calib(model, nf, seed=3)
    t0 = time.time()
    gp = c.Problem.from_synth(s)
    t_create = time.time() - t0
    gp.set_poses(s.init_poses)
    warm = gp.time_linearize(s.init_params, reps=20, flush_l2=False)
    cold = gp.time_linearize(s.init_params, reps=10, flush_l2=True)
    gp.set_poses(s.init_poses)
    t0 = time.time(); intr, summ, hist = gp.solve_lm(s.init_params); t_lm = time.time() - t0
    gp.set_poses(s.init_poses)
    t0 = time.time(); intr2, summ2, hist2 = gp.solve_gn(s.init_params); t_gn = time.time() - t0
    print(json.dumps({"model": model, "frames": s.n_frames, "obs": s.n_obs, "create_s": t_create, "k2_ms_warm": warm, "k2_ms_cold": cold,
                      "k2_gevals_s_warm": s.n_obs / warm / 1e6, "lm_iters": summ.iterations, "lm_ms": summ.device_ms, "lm_wall_ms": t_lm * 1e3,
                      "lm_ms_per_iter": summ.device_ms / summ.iterations, "gn_iters": summ2.iterations, "gn_ms": summ2.device_ms,
                      "gn_ms_per_iter": summ2.device_ms / summ2.iterations, "status": [summ.status, summ2.status],
                      "relerr": float(np.max(np.abs(intr - s.gt_params) / np.abs(s.gt_params))), "launches": gp.launch_count()}))
    gp.close()

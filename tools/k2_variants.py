"""A/B timing + parity of the K2 variants (CCRS_K2_VARIANT) on the BASELINE sizes."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ccrs_b200 as c

for model, nf in [("eucm", 7000), ("eucm", 2000), ("kb4", 7000), ("opencv5", 7000), ("eucm", 100)]:
    s = c.synth.make_calib(model, nf, seed=3)
    ref = None
    for variant in (0, 1):
        os.environ["CCRS_K2_VARIANT"] = str(variant)
        gp = c.Problem.from_synth(s)
        gp.set_poses(s.init_poses)
        sq = gp.linearize(s.init_params)
        B = gp.frame_blocks()
        warm = gp.time_linearize(s.init_params, reps=20, flush_l2=False)
        cold = gp.time_linearize(s.init_params, reps=10, flush_l2=True)
        gp.set_poses(s.init_poses)
        intr, summ, hist = gp.solve_lm(s.init_params)
        ms, launches = gp.bench_lm_steps(s.init_params, s.init_poses, warmup=3, steps=20, flush_l2=True)
        if ref is None:
            ref = (sq, B, intr)
            dmax = 0.0
        else:
            scale = np.abs(ref[1]).max(axis=1, keepdims=True)
            dmax = float(np.max(np.abs(B - ref[1]) / scale))
        print(json.dumps({"model": model, "frames": s.n_frames, "variant": variant, "k2_ms_warm": round(warm, 5), "k2_ms_cold": round(cold, 5),
                          "lm_step_ms": round(float(ms.mean()), 5), "lm_iters": summ.iterations, "blocks_max_rel_diff_vs_v0": dmax,
                          "intr_rel_diff_vs_v0": float(np.max(np.abs(intr - ref[2]) / np.abs(ref[2]))), "sq": float(sq[0])}))
        gp.close()

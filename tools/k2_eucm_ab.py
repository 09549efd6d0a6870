"""A/B of K2 builds on the 1M-observation EUCM problem: K2 alone (cold / warm) and the LM loop (warm), several repeats."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
s = c.synth.make_calib("eucm", 7000, seed=3)
gp = c.Problem.from_synth(s)
gp.set_poses(s.init_poses)
cold = [gp.time_linearize(s.init_params, reps=20, flush_l2=True) * 1e3 for _ in range(3)]
warm = [gp.time_linearize(s.init_params, reps=40, flush_l2=False) * 1e3 for _ in range(3)]
o = c.default_options(max_iteration=40, min_abs_decrease=-1.0, min_rel_decrease=-1.0, min_error=-1.0)
loops = []
for _ in range(3):
    gp.set_poses(s.init_poses)
    _, summ, _ = gp.solve_lm(s.init_params, options=o)
    loops.append(summ.device_ms / summ.iterations * 1e3)
step_ms, _ = gp.bench_lm_steps(s.init_params, s.init_poses, warmup=3, steps=20, flush_l2=True)
print(f"K2 cold {min(cold):.2f} us  warm {min(warm):.2f} us  LM loop warm {min(loops):.2f} us/iteration  flushed step {step_ms.mean()*1e3:.2f} us")
gp.close()

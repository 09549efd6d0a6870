import os, sys, time, faulthandler
faulthandler.dump_traceback_later(60, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ccrs_b200 as c
n = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
s = c.synth.make_calib("eucm", n, seed=3)
gp = c.Problem.from_synth(s)
o = c.default_options(max_iteration=40, min_abs_decrease=-1.0, min_rel_decrease=-1.0, min_error=-1.0)
for rep in range(2):
    gp.set_poses(s.init_poses)
    t0 = time.time()
    intr, summ, hist = gp.solve_lm(s.init_params, options=o)
    print("rep", rep, "wall", time.time() - t0, "iters", summ.iterations, "status", summ.status, "dev_ms", summ.device_ms, flush=True)
gp.close()

"""Debug: the host-driven controllers (CCRS_DEVICE_LOOP=0) on a small problem, stage by stage."""
import os, sys, faulthandler
faulthandler.dump_traceback_later(25, exit=True)
os.environ.setdefault("CCRS_DEVICE_LOOP", "0")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ccrs_b200 as c
s = c.synth.make_calib("eucm", int(sys.argv[1]) if len(sys.argv) > 1 else 100, seed=1, noise_px=0.1)
gp = c.Problem.from_synth(s)
gp.set_poses(s.init_poses)
print("linearize", gp.linearize(s.init_params), flush=True)
print("scale", gp.compute_scale()[0][:3], flush=True)
r = gp.reduce(0); print("reduce", r["sq_err"], flush=True)
gp.set_poses(s.init_poses)
for name in sys.argv[2:] or ["solve_gn", "solve_lm"]:
    gp.set_poses(s.init_poses)
    print("start", name, flush=True)
    intr, summ, hist = getattr(gp, name)(s.init_params)
    print(name, summ.iterations, summ.status, hist, flush=True)
gp.close()

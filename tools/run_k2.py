"""Run K2 a few times on the 1M-observation EUCM problem (ncu target)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
model = sys.argv[1] if len(sys.argv) > 1 else "eucm"
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 7000
s = c.synth.make_calib(model, nf, seed=3)
gp = c.Problem.from_synth(s)
gp.set_poses(s.init_poses)
print(gp.time_linearize(s.init_params, reps=5, flush_l2=False))
gp.close()

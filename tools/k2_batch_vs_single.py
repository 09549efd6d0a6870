import os, sys
import numpy as np
sys.path.insert(0, "/root/repo")
import ccrs_b200 as c
s = c.synth.make_calib("eucm", 7000, seed=3)
for batch in (False, True):
    kw = dict(problem_frame_offsets=np.array([0, s.n_frames], dtype=np.int32)) if batch else {}
    gp = c.Problem("eucm", s.width, s.height, s.frame_offsets, s.x, s.y, s.z, s.u, s.v, **kw)
    gp.set_poses(s.init_poses)
    w = gp.time_linearize(s.init_params, reps=20, flush_l2=False); cold = gp.time_linearize(s.init_params, reps=10, flush_l2=True)
    print("batch" if batch else "single", round(w * 1e3, 2), round(cold * 1e3, 2))
    gp.close()

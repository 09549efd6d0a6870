"""Device time per iteration of the two loop controllers on the bench problem (GN is the loop the reference runs)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
s = c.synth.make_calib("eucm", 7000, seed=3, noise_px=0.1)
gp = c.Problem.from_synth(s)
for name in ("solve_gn", "solve_lm"):
    ms = []
    for rep in range(12):
        gp.set_poses(s.init_poses)
        intr, summ, _ = getattr(gp, name)(s.init_params)
        if rep >= 2: ms.append(summ.device_ms)
    print(f"{name}: {summ.iterations} iterations, device {np.median(ms)*1e3:.1f} us total, {np.median(ms)/summ.iterations*1e3:.1f} us per iteration, status {summ.status}")
gp.close()

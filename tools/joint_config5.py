"""BASELINE config 5, second half: joint cam0 + cam1 extrinsic refinement (200 frames x 2 cameras x 144 corners, KB4)."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
rig = c.synth.make_rig("kb4", 200, 2, seed=4)
gj = c.JointProblem.from_rig(rig)
for rep in range(4):
    t0 = time.perf_counter()
    a, e, p, summ, hist = gj.solve_gn(rig.init_params, rig.init_extr, rig.init_poses)
    wall = time.perf_counter() - t0
print(f"joint GN: {gj.n_obs} obs, {summ.iterations} iterations, status {summ.status}, wall {wall*1e3:.2f} ms "
      f"({wall/summ.iterations*1e6:.0f} us / iteration), device {summ.device_ms:.3f} ms; extrinsic error vs truth "
      f"{np.max(np.abs(e[1] - rig.gt_extr[1])):.2e}")
gj.close()

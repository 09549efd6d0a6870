"""ncu target for the kernels outside the EUCM bench loop: K2 pair variant (KB4), K6 + radix select, k_pnp."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
s = c.synth.make_calib("kb4", 7000, seed=3)
gp = c.Problem.from_synth(s)
gp.set_poses(s.init_poses)
gp.linearize(s.init_params)
print(gp.validation(s.init_params))
R = c.synth.rodrigues(s.gt_poses[:, :3])
fi = np.repeat(np.arange(s.n_frames), np.diff(s.frame_offsets))
Pc = np.einsum("nij,nj->ni", R[fi], np.stack([s.x, s.y, s.z], axis=1)) + s.gt_poses[fi, 3:]
print(np.abs(c.init_poses(s.frame_offsets, s.x, s.y, s.z, Pc[:, 0] / Pc[:, 2], Pc[:, 1] / Pc[:, 2]) - s.gt_poses).max())
gp.close()

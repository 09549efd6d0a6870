"""compute-sanitizer target: small problems through the device-driven loops (single problem LM + GN, batch LM + GN,
board-format create, host-driven step-wise calls, joint GN)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ccrs_b200 as c
s = c.synth.make_calib("eucm", 37, seed=5, drop_fraction=0.2)
f = lambda a: a.astype(np.float32)
gp = c.Problem(s.model, s.width, s.height, s.frame_offsets, None, None, None, f(s.u), f(s.v), corner_id=s.extra["corner_id"], board=s.extra["board"])
for name in ("solve_lm", "solve_gn"):
    gp.set_poses(s.init_poses)
    intr, summ, _ = getattr(gp, name)(s.init_params)
    print(name, summ.iterations, summ.status)
gp.set_poses(s.init_poses); gp.linearize(s.init_params); gp.reduce(0); gp.close()
# tensor-core K2 variant (d >= 8 models) in the device-driven loop: dynamic frame hand-out, per-chunk statistics (37 = 2 x 16 + 5 frames)
sk = c.synth.make_calib("kb4", 37, seed=6, drop_fraction=0.2)
gk = c.Problem.from_synth(sk)
for name in ("solve_lm", "solve_gn"):
    gk.set_poses(sk.init_poses)
    intr, summ, _ = getattr(gk, name)(sk.init_params)
    print("kb4", name, summ.iterations, summ.status)
gk.close()
probs = [c.synth.make_calib("kb4", 12, seed=20 + i) for i in range(5)]
fo = np.concatenate([[0]] + [p.frame_offsets[1:] + sum(q.n_obs for q in probs[:i]) for i, p in enumerate(probs)]).astype(np.int32)
pfo = np.cumsum([0] + [p.n_frames for p in probs]).astype(np.int32)
cat = lambda k: np.concatenate([getattr(p, k) for p in probs])
gb = c.Problem("kb4", 1024, 1024, fo, cat("x"), cat("y"), cat("z"), cat("u"), cat("v"), problem_frame_offsets=pfo)
for name in ("solve_lm", "solve_gn"):
    gb.set_poses(np.concatenate([p.init_poses for p in probs]))
    intr, summ, _ = getattr(gb, name)(np.stack([p.init_params for p in probs]))
    print("batch", name, summ.iterations, summ.status)
gb.close()
rig = c.synth.make_rig("eucm", 12, 2, seed=4)
jp = c.JointProblem.from_rig(rig)
a, e, p, summ, _ = jp.solve_gn(rig.init_params, rig.init_extr, rig.init_poses)
print("joint", summ.iterations, summ.status)
jp.close()

"""ncu target: the device-driven loop on the 1M-observation EUCM problem (K2 = k_linearize, K3 = k_schur2), then two
Gauss-Newton iterations of the same size with KB4 (lane-pair K2). Usage: python tools/ncu_target.py [eucm|kb4|both]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
which = sys.argv[1] if len(sys.argv) > 1 else "both"
if which in ("eucm", "both"):
    s = c.synth.make_calib("eucm", 7000, seed=3)
    gp = c.Problem.from_synth(s)
    gp.set_poses(s.init_poses)
    intr, summ, _ = gp.solve_lm(s.init_params)
    print("eucm lm", summ.iterations, summ.status)
    gp.close()
if which in ("kb4", "both"):
    s = c.synth.make_calib("kb4", 7000, seed=3)
    gp = c.Problem.from_synth(s)
    gp.set_poses(s.init_poses)
    intr, summ, _ = gp.solve_gn(s.init_params, options=c.default_options(max_iteration=2))
    print("kb4 gn", summ.iterations, summ.status)
    gp.close()

"""Sweep lanes-per-frame G and the K2 variant on the 1M-observation EUCM problem."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
model = sys.argv[1] if len(sys.argv) > 1 else "eucm"
s = c.synth.make_calib(model, 7000, seed=3)
for variant, gs in ((0, [0, 4, 5, 6, 8, 9]), (1, [0, 2, 3, 4, 6, 8])):
    for g in gs:
        os.environ["CCRS_K2_VARIANT"] = str(variant)
        if g: os.environ["CCRS_FORCE_G"] = str(g)
        else: os.environ.pop("CCRS_FORCE_G", None)
        gp = c.Problem.from_synth(s)
        gp.set_poses(s.init_poses)
        warm = gp.time_linearize(s.init_params, reps=20, flush_l2=False)
        cold = gp.time_linearize(s.init_params, reps=10, flush_l2=True)
        print(json.dumps({"model": model, "variant": variant, "G": g, "k2_us_warm": round(warm * 1e3, 2), "k2_us_cold": round(cold * 1e3, 2)}))
        gp.close()

"""Sweep lanes-per-frame G (CCRS_FORCE_G; 0 = the library's own choice) for K2 on a model / frame count."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
model = sys.argv[1] if len(sys.argv) > 1 else "eucm"
nf = int(sys.argv[2]) if len(sys.argv) > 2 else 7000
s = c.synth.make_calib(model, nf, seed=3)
for g in (0, 1, 2, 3, 4, 5, 6, 8, 10, 16, 32):
    if g: os.environ["CCRS_FORCE_G"] = str(g)
    else: os.environ.pop("CCRS_FORCE_G", None)
    gp = c.Problem.from_synth(s)
    gp.set_poses(s.init_poses)
    warm = gp.time_linearize(s.init_params, reps=20, flush_l2=False)
    cold = gp.time_linearize(s.init_params, reps=10, flush_l2=True)
    print(json.dumps({"model": model, "frames": nf, "G": g, "k2_us_warm": round(warm * 1e3, 2), "k2_us_cold": round(cold * 1e3, 2)}))
    gp.close()

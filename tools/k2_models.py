"""K2 time per model / focal variant at one frame count: does pipe utilisation track the accumulator count?"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import ccrs_b200 as c
nf = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
for model in ("ucm", "eucm", "eucmt", "kb4", "opencv5", "ftheta"):
    s = c.synth.make_calib(model, nf, seed=3)
    for of in (False, True):
        gp = c.Problem.from_synth(s, xy_same_focal=of)
        gp.set_poses(s.init_poses)
        intr = c.synth.intr_from_full(s.init_params, of)
        warm = gp.time_linearize(intr, reps=20, flush_l2=False)
        cold = gp.time_linearize(intr, reps=10, flush_l2=True)
        print(json.dumps({"model": model, "one_focal": of, "d": gp.d, "obs": gp.n_obs, "k2_us_warm": round(warm * 1e3, 2),
                          "k2_us_cold": round(cold * 1e3, 2)}))
        gp.close()

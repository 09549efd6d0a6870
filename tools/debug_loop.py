"""Debug driver for the device-driven loop: config-4 size solve with the record trace printed (CCRS_LOOP_DEBUG=1)."""
import os, sys, time, faulthandler
faulthandler.dump_traceback_later(60, exit=True)
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ccrs_b200 as c
n = int(sys.argv[1]) if len(sys.argv) > 1 else 7000
loop = sys.argv[2] if len(sys.argv) > 2 else "lm"
s = c.synth.make_calib("eucm", n, seed=3)
gp = c.Problem.from_synth(s)
gp.set_poses(s.init_poses)
print("created", gp.n_frames, gp.n_obs, flush=True)
t0 = time.time()
intr, summ, hist = (gp.solve_lm if loop == "lm" else gp.solve_gn)(s.init_params)
print("solve", loop, time.time() - t0, summ.iterations, summ.status, summ.stop_reason, summ.device_ms, flush=True)
print("rel err vs gt", np.max(np.abs(intr - s.gt_params) / np.abs(s.gt_params)), flush=True)
poses = gp.get_poses()
gp.set_poses(s.init_poses)
intr2, summ2, hist2 = (gp.solve_lm if loop == "lm" else gp.solve_gn)(s.init_params)
print("repeat identical:", np.array_equal(intr, intr2), np.array_equal(hist, hist2), np.array_equal(poses, gp.get_poses()), flush=True)
gp.set_poses(s.init_poses)
sq = gp.linearize(s.init_params)[0]
print("linearize", sq, flush=True)
red = gp.reduce(0)
print("reduce", red["sq_err"][0], flush=True)
gp.close()
# ---- the rest of tests/test_full_size.py::test_config4 (host-driven path on shards) ----
if len(sys.argv) > 3:
    print("shards", flush=True)
    cut = 3123
    k = int(s.frame_offsets[cut])
    for a, b, ka, kb in ((0, cut, 0, k), (cut, s.n_frames, k, s.n_obs)):
        shard = c.Problem("eucm", s.width, s.height, s.frame_offsets[a:b + 1] - s.frame_offsets[a], s.x[ka:kb], s.y[ka:kb], s.z[ka:kb], s.u[ka:kb], s.v[ka:kb])
        print(" created", a, b, flush=True)
        shard.set_poses(s.init_poses[a:b])
        print(" poses", flush=True)
        shard.linearize(s.init_params)
        print(" linearized", flush=True)
        r = shard.reduce(0)
        print(" reduced", r["sq_err"], flush=True)
        shard.close()
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
    import oracle as O
    op = O.OracleProblem.from_synth(s, 1)
    t0 = time.time()
    ref = op.levenberg_marquardt(s.init_params, s.init_poses)
    print("oracle LM", time.time() - t0, ref[2].iterations, flush=True)

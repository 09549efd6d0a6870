"""BASELINE config 5 at full size: 4,096 independent KB4 calibrations x 200 frames x 144 corners (118 M observations)
in one handle on one GPU (sharded 512 per GPU across 8, no communication). Reports solve time and evals/s, and checks a
sample of problems against solving them on their own."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import ccrs_b200 as c

n_problems = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
n_distinct = 16
probs = [c.synth.make_calib("kb4", 200, seed=100 + i) for i in range(n_distinct)]
t0 = time.time()
fo, pfo = [np.zeros(1, dtype=np.int64)], [0]
xs, ys, zs, us, vs, poses, intr0 = [], [], [], [], [], [], []
for b in range(n_problems):
    s = probs[b % n_distinct]
    fo.append(fo[-1][-1] + s.frame_offsets[1:].astype(np.int64))
    pfo.append(pfo[-1] + s.n_frames)
    xs.append(s.x); ys.append(s.y); zs.append(s.z); us.append(s.u); vs.append(s.v)
    poses.append(s.init_poses); intr0.append(s.init_params)
cat = np.concatenate
fo = cat(fo).astype(np.int32); x, y, z, u, v = cat(xs), cat(ys), cat(zs), cat(us), cat(vs)
poses = cat(poses); intr0 = np.stack(intr0)
t_build = time.time() - t0
t0 = time.time()
gp = c.Problem("kb4", 1024, 1024, fo, x, y, z, u, v, problem_frame_offsets=np.array(pfo, dtype=np.int32))
t_create = time.time() - t0
out = {"n_problems": n_problems, "frames": int(gp.n_frames), "obs": int(gp.n_obs), "host_build_s": round(t_build, 2), "create_s": round(t_create, 3)}
for name in ("solve_gn", "solve_lm"):
    gp.set_poses(poses)
    t0 = time.time()
    intr, summ, _ = getattr(gp, name)(intr0)
    wall = time.time() - t0
    n_lin = summ.iterations + (1 if name == "solve_lm" else 0)
    out[name] = {"iterations_max": summ.iterations, "status": summ.status, "device_ms": round(summ.device_ms, 2), "wall_ms": round(wall * 1e3, 2),
                 "evals_per_s": gp.n_obs * n_lin / wall, "problems_per_s": n_problems / wall}
    # sample check: problems solved on their own give the same intrinsics
    worst = 0.0
    for b in (0, 5, n_problems - 1):
        s = probs[b % n_distinct]
        q = c.Problem.from_synth(s)
        q.set_poses(s.init_poses)
        ref, _, _ = getattr(q, name)(s.init_params)
        worst = max(worst, float(np.max(np.abs(intr[b] - ref) / np.abs(ref))))
        q.close()
    out[name]["max_rel_diff_vs_standalone"] = worst
print(json.dumps(out))
gp.close()

"""Regenerates tests/golden/golden_v2.npz: oracle outputs for the §8(f) entry points (validation, init_ucm stage 1,
convert_model, initial poses). Run from the repo root: `python tests/golden/make_golden_v2.py`.
Same status as golden_v1: these pin the ORACLE (the Rust reference cannot run here) so that the GPU box can check the
CUDA path against committed numbers and the checker cannot drift silently."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402
import pnp_oracle  # noqa: E402

pkg = importlib.import_module("camera-intrinsic-calibration-rs_b200")
out = {}
# validation: median / mean of best 99 % at the perturbed start, three models
for model in ("eucm", "kb4", "opencv5"):
    s = pkg.synth.make_calib(model, 25, seed=40, noise_px=0.1, drop_fraction=0.2)
    op = O.OracleProblem.from_synth(s, pkg.MODELS[model])
    med, avg, e = op.validation(s.init_params, s.init_poses)
    out[f"val_{model}"] = np.array([med, avg]); out[f"val_{model}_err_head"] = e[:32]
# init_ucm stage 1: [f, alpha] + poses of two frames
gt = np.array([400.0, 400.0, 512.0, 512.0, 0.6])
s = pkg.synth.make_calib("ucm", 2, seed=41, gt_params=gt, noise_px=0.05)
op = O.OracleProblem.from_synth(s, 0)
for ff in (0, 1):
    fa, poses, res, hist = O.init_ucm_gn(op, 512.0, 512.0, 330.0, 0.5, s.init_poses, fixed_focal=bool(ff))
    out[f"ucm_init_ff{ff}_fa"] = fa; out[f"ucm_init_ff{ff}_poses"] = poses; out[f"ucm_init_ff{ff}_iters"] = res.iterations
    out[f"ucm_init_ff{ff}_hist"] = hist
# convert_model: KB4 -> EUCM and EUCM -> KB4 (one distortion disabled) on the reference's pixel grid
for src, tgt, dis in (("kb4", "eucm", 0), ("eucm", "kb4", 1)):
    sp = np.array(pkg.synth.GT_PARAMS[src], dtype=np.float64)
    nt = len(pkg.synth.GT_PARAMS[tgt])
    init = np.zeros(nt); init[:4] = sp[:4]
    if tgt == "eucm":
        init[4], init[5] = 0.5, 1.0
    rays, valid = pkg.models.unproject(src, sp, pkg.models.conversion_grid(1024, 1024))
    lo, hi = np.full(nt, -np.inf), np.full(nt, np.inf)      # bounds come from the product at test time; wide here
    fixed = np.zeros(nt, dtype=np.uint8)
    for i in range(dis):
        fixed[nt - 1 - i] = 1
    out[f"conv_{src}_{tgt}_rays"] = rays[valid]
    out[f"conv_{src}_{tgt}_init"] = init
    ref, res, hist = O.convert_model_gn(pkg.MODELS[src], sp, pkg.MODELS[tgt], init, rays[valid], None, None, fixed)
    out[f"conv_{src}_{tgt}_params"] = ref; out[f"conv_{src}_{tgt}_iters"] = res.iterations
# initial poses: noisy normalised points, 6 ragged frames
s = pkg.synth.make_calib("eucm", 6, seed=42, drop_fraction=0.4)
R = pkg.synth.rodrigues(s.gt_poses[:, :3])
fi = np.repeat(np.arange(s.n_frames), np.diff(s.frame_offsets))
Pc = np.einsum("nij,nj->ni", R[fi], np.stack([s.x, s.y, s.z], axis=1)) + s.gt_poses[fi, 3:]
rng = np.random.default_rng(7)
xn = (Pc[:, 0] / Pc[:, 2] + rng.normal(scale=1e-3, size=len(Pc))).astype(np.float32).astype(np.float64)
yn = (Pc[:, 1] / Pc[:, 2] + rng.normal(scale=1e-3, size=len(Pc))).astype(np.float32).astype(np.float64)
poses, costs = [], []
for f in range(s.n_frames):
    a, b = s.frame_offsets[f], s.frame_offsets[f + 1]
    rv, t, c = pnp_oracle.solve_frame(np.stack([s.x[a:b], s.y[a:b], s.z[a:b]], axis=1), xn[a:b], yn[a:b], n_starts=32, seed=f)
    poses.append(np.concatenate([rv, t])); costs.append(c)
out["pnp_xn"] = xn; out["pnp_yn"] = yn; out["pnp_poses"] = np.array(poses); out["pnp_cost"] = np.array(costs)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v2.npz"), **out)
print("wrote", len(out), "arrays")

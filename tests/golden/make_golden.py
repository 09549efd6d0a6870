"""Regenerates tests/golden/golden_v1.npz from the CPU oracle (run from the repo root:
`python tests/golden/make_golden.py`). The reference itself (Rust, un-vendored crates) cannot be run here, so
these vectors pin the ORACLE's outputs — which tests/test_oracle_pins.py ties to the reference's own test
fixtures, OpenCV and mpmath — and let the GPU box check the CUDA path without rebuilding anything."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle as O  # noqa: E402

pkg = importlib.import_module("camera-intrinsic-calibration-rs_b200")
out = {}
for model in ["ucm", "eucm", "eucmt", "kb4", "opencv5", "ftheta"]:
    for of in (0, 1):
        s = pkg.synth.make_calib(model, 4, seed=100 + of, drop_fraction=0.7)
        op = O.OracleProblem.from_synth(s, pkg.MODELS[model], xy_same_focal=bool(of))
        intr = pkg.synth.intr_from_full(s.init_params, bool(of))
        r, J = op.eval_rj(intr, s.init_poses, apply_loss=True)
        sq, blk = op.linearize(intr, s.init_poses)
        k = f"{model}_of{of}"
        out[k + "_r"] = r; out[k + "_J"] = J; out[k + "_blk"] = blk; out[k + "_sq"] = sq
# trajectories (BASELINE config 1 size for EUCM; smaller for the sweep)
for model, nf in [("eucm", 100), ("kb4", 40), ("opencv5", 40)]:
    s = pkg.synth.make_calib(model, nf, seed=0)
    op = O.OracleProblem.from_synth(s, pkg.MODELS[model])
    for name, fn in (("gn", op.gauss_newton), ("lm", op.levenberg_marquardt)):
        intr, poses, res, hist = fn(s.init_params, s.init_poses)
        k = f"traj_{model}_{name}"
        out[k + "_intr"] = intr; out[k + "_hist"] = hist; out[k + "_iters"] = res.iterations
        out[k + "_acc_rej"] = np.array([res.n_accepted, res.n_rejected])
        out[k + "_poses_head"] = poses[:5]
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "golden_v1.npz"), **out)
print("wrote", len(out), "arrays")

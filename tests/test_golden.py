"""Committed golden vectors (tests/golden/golden_v1.npz, made by tests/golden/make_golden.py).
CPU: the oracle still reproduces them (no silent drift of the checker). GPU: the CUDA path matches them."""
import os

import numpy as np
import pytest

from helpers import MODEL_NAMES, block_rel_err, rel_err_rows

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v1.npz"))


def _case(pkg, model, of):
    s = pkg.synth.make_calib(model, 4, seed=100 + of, drop_fraction=0.7)
    return s, pkg.synth.intr_from_full(s.init_params, bool(of))


@pytest.mark.parametrize("of", [0, 1])
@pytest.mark.parametrize("model", MODEL_NAMES)
def test_oracle_reproduces_golden(pkg, oracle, model, of):
    s, intr = _case(pkg, model, of)
    op = oracle.OracleProblem.from_synth(s, pkg.MODELS[model], xy_same_focal=bool(of))
    r, J = op.eval_rj(intr, s.init_poses, apply_loss=True)
    k = f"{model}_of{of}"
    assert np.allclose(r, G[k + "_r"], rtol=1e-12, atol=1e-12)
    assert np.max(rel_err_rows(J, G[k + "_J"])) < 1e-11
    sq, blk = op.linearize(intr, s.init_poses)
    assert abs(sq - G[k + "_sq"]) / G[k + "_sq"] < 1e-12
    assert block_rel_err(blk, G[k + "_blk"], op.d + 7) < 1e-11


def test_oracle_reproduces_golden_trajectory(pkg, oracle):
    s = pkg.synth.make_calib("eucm", 100, seed=0)
    op = oracle.OracleProblem.from_synth(s, 1)
    intr, _, res, hist = op.gauss_newton(s.init_params, s.init_poses)
    assert res.iterations == int(G["traj_eucm_gn_iters"])
    assert np.allclose(intr, G["traj_eucm_gn_intr"], rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("of", [0, 1])
@pytest.mark.parametrize("model", MODEL_NAMES)
def test_gpu_matches_golden_rj_and_blocks(pkg, model, of):
    s, intr = _case(pkg, model, of)
    gp = pkg.Problem.from_synth(s, xy_same_focal=bool(of))
    r, J = gp.eval_rj(intr, s.init_poses, apply_loss=True)
    k = f"{model}_of{of}"
    assert np.max(np.abs(r - G[k + "_r"]) / np.maximum(np.abs(G[k + "_r"]), 1e-3)) < 1e-9
    assert np.max(rel_err_rows(J, G[k + "_J"])) < 1e-9
    gp.set_poses(s.init_poses)
    sq = gp.linearize(intr)
    assert abs(sq[0] - G[k + "_sq"]) / G[k + "_sq"] < 1e-12
    assert block_rel_err(gp.frame_blocks(), G[k + "_blk"], gp.d + 7) < 1e-9
    gp.close()


@pytest.mark.gpu
@pytest.mark.parametrize("model,nf", [("eucm", 100), ("kb4", 40), ("opencv5", 40)])
@pytest.mark.parametrize("loop", ["gn", "lm"])
def test_gpu_matches_golden_trajectory(pkg, model, nf, loop):
    s = pkg.synth.make_calib(model, nf, seed=0)
    gp = pkg.Problem.from_synth(s)
    gp.set_poses(s.init_poses)
    intr, summ, hist = (gp.solve_gn if loop == "gn" else gp.solve_lm)(s.init_params)
    k = f"traj_{model}_{loop}"
    assert summ.iterations == int(G[k + "_iters"])                        # equal iteration count
    if loop == "lm":
        assert [summ.n_accepted, summ.n_rejected] == list(G[k + "_acc_rej"])
    assert np.max(np.abs(intr - G[k + "_intr"]) / np.abs(G[k + "_intr"])) < 1e-6   # north_star tolerance
    assert np.allclose(hist, G[k + "_hist"], rtol=1e-6, atol=1e-9)
    assert np.max(np.abs(gp.get_poses()[:5] - G[k + "_poses_head"])) < 1e-7
    gp.close()

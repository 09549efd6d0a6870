"""The PRODUCT's host loop controllers (ccrs_controller_gn / ccrs_controller_lm in libccrs_b200.so) driven on the
CPU through the backend callback table with oracle-computed per-frame work: they must reproduce the oracle's own
GN / LM loops (same iteration count, accept/reject sequence, intrinsics), including bounds and fixed variables."""
import numpy as np
import pytest

from oracle_backend import OracleBackend


def _mk(pkg, oracle, model, nf, seed, **kw):
    s = pkg.synth.make_calib(model, nf, seed=seed, **kw)
    op = oracle.OracleProblem.from_synth(s, pkg.MODELS[model], n_threads=4)
    return s, op


@pytest.mark.parametrize("model", ["eucm", "kb4", "opencv5"])
def test_gn_controller_equals_oracle_loop(pkg, oracle, model):
    s, op = _mk(pkg, oracle, model, 12, 0)
    be = OracleBackend(pkg, op, s.init_poses)
    code, intr, poses, summ, hist = be.run("gn", s.init_params)
    intr_ref, poses_ref, res, hist_ref = op.gauss_newton(s.init_params, s.init_poses)
    assert code == 0 and summ.iterations == res.iterations and summ.stop_reason == res.stop_reason
    assert np.max(np.abs(intr - intr_ref) / np.abs(intr_ref)) < 1e-9
    assert np.max(np.abs(poses - poses_ref)) < 1e-9
    assert np.allclose(hist, hist_ref, rtol=1e-9)


@pytest.mark.parametrize("speculative", [1, 0])
def test_lm_controller_equals_oracle_loop(pkg, oracle, speculative):
    s, op = _mk(pkg, oracle, "eucm", 12, 1, noise_px=0.1)
    be = OracleBackend(pkg, op, s.init_poses)
    code, intr, poses, summ, hist = be.run("lm", s.init_params, options=pkg.default_options(speculative=speculative))
    intr_ref, poses_ref, res, hist_ref = op.levenberg_marquardt(s.init_params, s.init_poses)
    assert code == 0 and summ.iterations == res.iterations
    assert (summ.n_accepted, summ.n_rejected) == (res.n_accepted, res.n_rejected)
    assert np.max(np.abs(intr - intr_ref) / np.abs(intr_ref)) < 1e-8
    assert np.allclose(hist, hist_ref, rtol=1e-8)
    # speculative LM linearises the trial point instead of a residual-only pass: one K2 per iteration, none extra
    n_lin = be.calls.count("linearize")
    assert n_lin == (1 if speculative else 1 + summ.n_accepted)


def test_lm_controller_rejections(pkg, oracle):
    s, op = _mk(pkg, oracle, "eucm", 40, 2)
    intr_bad = s.init_params * np.array([0.7, 0.7, 1.05, 0.95, 0.7, 1.5])
    poses = s.init_poses + np.random.default_rng(1).normal(scale=0.2, size=s.init_poses.shape)
    be = OracleBackend(pkg, op, poses)
    code, intr, _, summ, hist = be.run("lm", intr_bad)
    intr_ref, _, res, hist_ref = op.levenberg_marquardt(intr_bad, poses)
    assert res.n_rejected > 0
    assert (summ.iterations, summ.n_accepted, summ.n_rejected) == (res.iterations, res.n_accepted, res.n_rejected)
    assert np.max(np.abs(intr - intr_ref) / np.abs(intr_ref)) < 1e-6


@pytest.mark.parametrize("fixed_mode", [0, 1])
def test_bounds_and_fixed(pkg, oracle, fixed_mode):
    """set_problem_parameter_bound / set_problem_parameter_disabled (util.rs:29-71)."""
    s, op = _mk(pkg, oracle, "kb4", 12, 4)
    lo, hi = pkg.model_bounds("kb4", s.width, s.height)
    fixed = np.zeros(8, dtype=np.uint8); fixed[-2:] = 1
    intr0 = s.init_params.copy(); intr0[-2:] = 0.0
    be = OracleBackend(pkg, op, s.init_poses)
    code, intr, _, summ, _ = be.run("gn", intr0, lo, hi, fixed, options=pkg.default_options(fixed_mode=fixed_mode))
    intr_ref, _, res, _ = op.gauss_newton(intr0, s.init_poses, lo, hi, fixed, options=op.default_options(fixed_mode=fixed_mode))
    assert code == 0 and summ.iterations == res.iterations
    assert np.all(intr[-2:] == 0.0)
    assert np.max(np.abs(intr[:-2] - intr_ref[:-2]) / np.abs(intr_ref[:-2])) < 1e-8


def test_bounds_clamp_is_active(pkg, oracle):
    s, op = _mk(pkg, oracle, "eucm", 12, 5)
    lo, hi = pkg.model_bounds("eucm", s.width, s.height)
    hi = hi.copy(); hi[0] = s.gt_params[0] * 0.98     # true fx is out of bounds -> clamp binds
    be = OracleBackend(pkg, op, s.init_poses)
    intr0 = s.init_params.copy(); intr0[0] = hi[0] * 0.99
    code, intr, _, summ, _ = be.run("gn", intr0, lo, hi, options=pkg.default_options(max_iteration=8))
    intr_ref, _, res, _ = op.gauss_newton(intr0, s.init_poses, lo, hi, options=op.default_options(max_iteration=8))
    assert intr[0] <= hi[0] and np.allclose(intr, intr_ref, rtol=1e-8)


def test_cholesky_failure_is_reported(pkg, oracle):
    """tiny-solver returns None when the LLT fails: a frame whose points give a rank-deficient pose block."""
    fo = np.array([0, 12], dtype=np.int32)
    z = np.zeros(12)
    x = np.full(12, 0.1); y = np.full(12, 0.2)          # 12 identical points: pose block is singular
    u = np.full(12, 500.0); v = np.full(12, 500.0)
    op = oracle.OracleProblem(1, 1024, 1024, fo, x, y, z, u, v)
    be = OracleBackend(pkg, op, np.array([[0.1, 0.0, 0.0, 0.0, 0.0, 0.5]]))
    code, _, _, summ, _ = be.run("gn", pkg.synth.GT_PARAMS["eucm"])
    _, _, res, _ = op.gauss_newton(pkg.synth.GT_PARAMS["eucm"], np.array([[0.1, 0.0, 0.0, 0.0, 0.0, 0.5]]))
    assert res.status == -2
    assert code == -5 and summ.status == -5   # CCRS_ERR_CHOLESKY


def test_mask_value_2_removes_the_variable_from_the_system(pkg, oracle):
    """init_ucm's [f, alpha] problem (util.rs:295-357): a one-focal UCM whose cx, cy carry mask 2 must follow the
    oracle's dense Gauss-Newton over the reference's own variables ([f, alpha] + two poses), iteration by iteration."""
    gt = np.array([400.0, 400.0, 512.0, 512.0, 0.6])
    s = pkg.synth.make_calib("ucm", 2, seed=12, gt_params=gt, noise_px=0.1)
    fa, poses_ref, res, hist_ref = oracle.init_ucm_gn(oracle.OracleProblem.from_synth(s, 0), 512.0, 512.0, 350.0, 0.55, s.init_poses)
    op1 = oracle.OracleProblem.from_synth(s, 0, xy_same_focal=True, n_threads=2)
    be = OracleBackend(pkg, op1, s.init_poses)
    inf = np.inf
    code, intr, poses, summ, hist = be.run("gn", np.array([350.0, 512.0, 512.0, 0.55]), np.array([350.0 / 3, -inf, -inf, 1e-6]),
                                           np.array([350.0 * 3, inf, inf, 1.0]), np.array([0, 2, 2, 0], dtype=np.uint8))
    assert code == 0 and summ.iterations == res.iterations
    assert intr[1] == 512.0 and intr[2] == 512.0
    assert abs(intr[0] - fa[0]) <= 1e-8 * fa[0] and abs(intr[3] - fa[1]) <= 1e-8 * fa[1]
    assert np.max(np.abs(poses - poses_ref)) < 1e-8
    assert np.allclose(hist, hist_ref, rtol=1e-9)
    # with fixed_mode = 0, mask 1 would have kept cx, cy in the linear system (a different step): mask 2 is not mask 1
    be1 = OracleBackend(pkg, op1, s.init_poses)
    _, intr1, _, _, hist1 = be1.run("gn", np.array([350.0, 512.0, 512.0, 0.55]), None, None, np.array([0, 1, 1, 0], dtype=np.uint8))
    assert intr1[1] == 512.0 and not np.allclose(hist1[:3], hist[:3], rtol=1e-6)


def test_block_huber_changes_the_error_not_the_step(pkg, oracle):
    """ModelConvertFactor is ONE residual block under one HuberLoss (util.rs:246-251): the corrector scales r and J by
    the same factor, so the GN iterates are those of the loss-free problem and only the reported error changes."""
    s, _ = _mk(pkg, oracle, "eucm", 8, 6, noise_px=0.5)
    op = oracle.OracleProblem(1, s.width, s.height, s.frame_offsets, s.x, s.y, s.z, s.u, s.v, huber_delta=0.0, n_threads=2)
    o_plain = pkg.default_options(max_iteration=4, min_abs_decrease=-1.0, min_rel_decrease=-1.0)
    o_block = pkg.default_options(max_iteration=4, min_abs_decrease=-1.0, min_rel_decrease=-1.0, block_huber_delta=1.0)
    _, intr_a, poses_a, _, hist_a = OracleBackend(pkg, op, s.init_poses).run("gn", s.init_params, options=o_plain)
    _, intr_b, poses_b, _, hist_b = OracleBackend(pkg, op, s.init_poses).run("gn", s.init_params, options=o_block)
    assert np.array_equal(intr_a, intr_b) and np.array_equal(poses_a, poses_b)
    # error seen by the stop tests: ||r|| without the loss, (delta^2 s)^(1/4) = sqrt(delta ||r||) with it (s > delta^2)
    assert np.all(hist_a > 1.0) and np.allclose(hist_b, np.sqrt(hist_a), rtol=1e-12)

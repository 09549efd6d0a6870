"""the two small problems of src/optimization that use other factors —
init_ucm (UCMInitFocalAlphaFactor, src/util.rs:284-378, factors.rs:83-120) and convert_model (ModelConvertFactor,
src/util.rs:225-278, factors.rs:11-77) — running on the ReprojectionFactor kernels.
The oracle restates both problems directly (dense normal equations over the reference's own variables)."""
import numpy as np
import pytest

from helpers import MODEL_NAMES, rms_px


# ------------------------------------------------------------------------------------------------------------ CPU side
@pytest.mark.parametrize("model", MODEL_NAMES)
def test_unproject_round_trip(pkg, model):
    """models.unproject is the inverse of the projection on the conversion grid (ModelConvertFactor::new)."""
    prm = np.array(pkg.synth.GT_PARAMS[model], dtype=np.float64)
    grid = pkg.models.conversion_grid(1024, 1024)
    assert len(grid) == 30 * 30 and grid[0].tolist() == [10.0, 10.0] and grid[1].tolist() == [44.0, 10.0]
    rays, valid = pkg.models.unproject(model, prm, grid)
    assert valid.mean() > 0.6
    back = pkg.synth.project(model, prm, rays[valid])
    assert np.max(np.abs(back - grid[valid])) < 1e-6


def test_oracle_init_ucm_recovers_focal_and_alpha(pkg, oracle):
    """two frames of an exact UCM with the principal point at the image centre: stage 1 recovers (f, alpha)."""
    gt = np.array([400.0, 400.0, 512.0, 512.0, 0.6])
    s = pkg.synth.make_calib("ucm", 2, seed=11, gt_params=gt)
    op = oracle.OracleProblem.from_synth(s, 0)
    fa, poses, res, hist = oracle.init_ucm_gn(op, 512.0, 512.0, 330.0, 0.5, s.init_poses)
    assert res.status == 0 and res.iterations < 30
    assert abs(fa[0] - 400.0) < 1e-2 and abs(fa[1] - 0.6) < 1e-4
    # fixed focal: f stays, alpha still moves
    fa2, _, res2, _ = oracle.init_ucm_gn(op, 512.0, 512.0, 330.0, 0.5, s.init_poses, fixed_focal=True)
    assert fa2[0] == 330.0 and fa2[1] != 0.5


def test_oracle_convert_model_kb4_to_eucm(pkg, oracle):
    src = np.array(pkg.synth.GT_PARAMS["kb4"], dtype=np.float64)
    rays, valid = pkg.models.unproject("kb4", src, pkg.models.conversion_grid(1024, 1024))
    tgt0 = np.array([0, 0, 0, 0, 0.5, 1.0], dtype=np.float64)
    tgt0[:4] = src[:4]
    lo, hi = None, None
    tgt, res, hist = oracle.convert_model_gn(3, src, 1, tgt0, rays[valid], lo, hi)
    assert res.status == 0
    fit = pkg.synth.project("eucm", tgt, rays[valid]) - pkg.synth.project("kb4", src, rays[valid])
    assert np.sqrt(np.mean(fit ** 2)) < 0.5        # an EUCM approximates this KB4 to sub-pixel level
    assert hist[-1] < hist[0]


# ------------------------------------------------------------------------------------------------------------ GPU side
def _two_frames(pkg, s):
    frames, init = pkg.synth.to_frame_features(s)
    return frames[0], frames[1], init[0], init[1]


@pytest.mark.gpu
@pytest.mark.parametrize("fixed_focal", [False, True])
def test_gpu_init_ucm_matches_oracle(pkg, oracle, fixed_focal):
    gt = np.array([400.0, 400.0, 512.0, 512.0, 0.6])
    s = pkg.synth.make_calib("ucm", 2, seed=11, gt_params=gt, noise_px=0.05)
    op = oracle.OracleProblem.from_synth(s, 0)
    f0, a0 = 330.0, 0.5
    # oracle: stage 1 directly restated; stage 2 = calib_camera of the two frames with the stage-1 model (util.rs:358-372):
    # fresh PnP poses under that model (the CUDA pose initialisation, checked on its own in test_init_poses.py), then
    # one-focal UCM Gauss-Newton (+ the fixed-focal second pass)
    fa, poses1, res1, _ = oracle.init_ucm_gn(op, 512.0, 512.0, f0, a0, s.init_poses, fixed_focal=fixed_focal)
    assert res1.status == 0
    frames, _ = pkg.synth.to_frame_features(s)
    cam1 = pkg.GenericModel("ucm", np.array([fa[0], fa[0], 512.0, 512.0, fa[1]]), s.width, s.height)
    init = pkg.initial_poses(frames, cam1)
    poses_pnp = np.array([init[f].as_array() for f in range(s.n_frames)])
    assert np.max(np.abs(poses_pnp - s.gt_poses)) < 0.2          # rough (the stage-1 model is approximate) but sane
    op1 = oracle.OracleProblem.from_synth(s, 0, xy_same_focal=True)
    lo, hi = pkg.model_bounds("ucm", s.width, s.height)
    lo1, hi1 = np.delete(lo, 1), np.delete(hi, 1)
    intr = np.array([fa[0], 512.0, 512.0, fa[1]])
    intr, poses2, _, _ = op1.gauss_newton(intr, poses_pnp, lo1, hi1)
    if fixed_focal:
        intr[0] = fa[0]
        intr, poses2, _, _ = op1.gauss_newton(intr, poses2, lo1, hi1, fixed=[1, 0, 0, 0])
    ref = np.insert(intr, 1, intr[0])
    ff0, ff1, rt0, rt1 = _two_frames(pkg, s)
    cam = pkg.init_ucm(ff0, ff1, rt0, rt1, f0, a0, fixed_focal)
    assert cam is not None and cam.model == "ucm"
    assert np.max(np.abs(cam.params - ref) / np.abs(ref)) < 1e-6          # north_star: intrinsics within 1e-6 relative
    if fixed_focal:
        assert cam.params[0] == f0 and cam.params[1] == f0
    else:
        assert abs(cam.params[0] - 400.0) < 5.0 and abs(cam.params[4] - 0.6) < 2e-2   # two noisy frames only


@pytest.mark.gpu
def test_gpu_init_ucm_stage1_is_the_two_parameter_problem(pkg, oracle):
    """the [f, alpha] problem through the step-wise ABI: one-focal UCM with cx, cy masked out (fixed = 2) gives the
    oracle's dense 14 x 14 Gauss-Newton trajectory."""
    gt = np.array([400.0, 400.0, 512.0, 512.0, 0.6])
    s = pkg.synth.make_calib("ucm", 2, seed=12, gt_params=gt, noise_px=0.1)
    op = oracle.OracleProblem.from_synth(s, 0)
    fa, poses_ref, res, hist_ref = oracle.init_ucm_gn(op, 512.0, 512.0, 350.0, 0.55, s.init_poses)
    inf = np.inf
    with pkg.Problem.from_synth(s, xy_same_focal=True) as gp:
        gp.set_poses(s.init_poses)
        intr, summ, hist = gp.solve_gn([350.0, 512.0, 512.0, 0.55], lo=[350.0 / 3, -inf, -inf, 1e-6],
                                       hi=[350.0 * 3, inf, inf, 1.0], fixed=[0, 2, 2, 0])
        assert summ.status == 0 and summ.iterations == res.iterations
        assert intr[1] == 512.0 and intr[2] == 512.0
        assert abs(intr[0] - fa[0]) / fa[0] < 1e-6 and abs(intr[3] - fa[1]) / fa[1] < 1e-6
        assert np.max(np.abs(gp.get_poses() - poses_ref)) < 1e-6
        np.testing.assert_allclose(hist, hist_ref, rtol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("src,tgt,disabled", [("kb4", "eucm", 0), ("eucm", "kb4", 0), ("opencv5", "eucmt", 2),
                                              ("opencv5", "kb4", 1), ("kb4", "ucm", 0), ("eucm", "eucmt", 0)])
def test_gpu_convert_model_matches_oracle(pkg, oracle, src, tgt, disabled):
    sp = np.array(pkg.synth.GT_PARAMS[src], dtype=np.float64)
    tp0 = np.array(pkg.synth.GT_PARAMS[tgt], dtype=np.float64) * 0.0
    nt = len(tp0)
    if tgt in ("ucm", "eucm", "eucmt"):
        tp0[4] = 0.5
        if nt > 5:
            tp0[5] = 1.0
    source = pkg.GenericModel(src, sp, 1024, 1024)
    target = pkg.GenericModel(tgt, tp0.copy(), 1024, 1024)
    pkg.convert_model(source, target, disabled)
    # oracle on the same points, same initial values / bounds / disabled set (util.rs:253-271)
    rays, valid = pkg.models.unproject(src, sp, pkg.models.conversion_grid(1024, 1024))
    lo, hi = pkg.model_bounds(tgt, 1024, 1024)
    init = tp0.copy(); init[:4] = sp[:4]
    fixed = np.zeros(nt, dtype=np.uint8)
    for i in range(disabled):
        fixed[nt - 1 - i] = 1; init[nt - 1 - i] = 0.0
    ref, res, hist = oracle.convert_model_gn(pkg.MODELS[src], sp, pkg.MODELS[tgt], init, rays[valid], lo, hi, fixed)
    assert res.status == 0
    scale = np.maximum(np.abs(ref), 1e-3)            # distortion coefficients may be ~0: absolute floor 1e-9
    assert np.max(np.abs(target.params - ref) / scale) < 1e-6
    for i in range(disabled):
        assert target.params[nt - 1 - i] == 0.0
    # and the conversion did its job: the target reproduces the source better than the starting point did
    uv_src = pkg.synth.project(src, sp, rays[valid])
    start = init.copy()
    rms = lambda prm: np.sqrt(np.mean((pkg.synth.project(tgt, prm, rays[valid]) - uv_src) ** 2))
    assert rms(target.params) < rms(start)


@pytest.mark.gpu
def test_gpu_convert_model_ucm_closed_form(pkg):
    """UCM -> EUCM / EUCMT is a parameter copy with beta = 1 (util.rs:230-243)."""
    sp = np.array(pkg.synth.GT_PARAMS["ucm"], dtype=np.float64)
    for tgt, tail in (("eucm", [1.0]), ("eucmt", [1.0, 0.0, 0.0])):
        target = pkg.GenericModel(tgt, np.zeros(5 + len(tail)), 1024, 1024)
        pkg.convert_model(pkg.GenericModel("ucm", sp, 1024, 1024), target, 0)
        assert target.params.tolist() == sp.tolist() + tail


def test_convert_model_rejects_size_mismatch(pkg):
    a = pkg.GenericModel("ucm", np.array(pkg.synth.GT_PARAMS["ucm"], dtype=np.float64), 1024, 1024)
    b = pkg.GenericModel("eucm", np.zeros(6), 640, 480)
    with pytest.raises(ValueError):
        pkg.convert_model(a, b, 0)

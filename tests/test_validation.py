"""util::validation (src/util.rs:721-795): median and mean-of-best-99 % reprojection error.
CPU: the oracle restatement against a direct numpy evaluation of the reference's formulae.
GPU: K6 + radix select through the C ABI against the oracle (which sorts like the reference)."""
import numpy as np
import pytest

from helpers import MODEL_NAMES


def _numpy_validation(e):
    """util.rs:771-781 verbatim on an error vector."""
    s = np.sort(e)
    n99 = len(s) * 99 // 100
    return s[len(s) // 2], float(np.sum(s[:n99] / n99)) if n99 else 0.0


def test_oracle_validation_follows_reference_formulae(pkg, oracle):
    s = pkg.synth.make_calib("eucm", 12, seed=2, noise_px=0.1)
    op = oracle.OracleProblem.from_synth(s, pkg.MODELS["eucm"])
    med, avg, e = op.validation(s.gt_params, s.gt_poses)
    r = op.eval_r(s.gt_params, s.gt_poses, apply_loss=False).reshape(-1, 2)
    np.testing.assert_allclose(e, np.hypot(r[:, 0], r[:, 1]), rtol=1e-14, atol=0)
    m2, a2 = _numpy_validation(e)
    assert med == m2
    assert abs(avg - a2) <= 1e-15 * max(1.0, abs(a2))
    assert 0.05 < med < 0.2 and avg < 0.2       # sigma = 0.1 px noise + f32 rounding


@pytest.mark.gpu
@pytest.mark.parametrize("model", MODEL_NAMES)
def test_gpu_validation_matches_oracle(pkg, oracle, model):
    s = pkg.synth.make_calib(model, 60, seed=4, noise_px=0.1, drop_fraction=0.15)
    op = oracle.OracleProblem.from_synth(s, pkg.MODELS[model])
    med_ref, avg_ref, e_ref = op.validation(s.init_params, s.init_poses)
    with pkg.Problem.from_synth(s) as gp:
        med, avg, e = gp.validation(s.init_params, s.init_poses, want_errors=True)
        # per-point errors: same tolerance as the residuals they come from
        assert np.max(np.abs(e - e_ref) / np.maximum(e_ref, 1e-3)) < 1e-9
        assert abs(med - med_ref) <= 1e-9 * max(med_ref, 1e-3)
        assert abs(avg - avg_ref) <= 1e-9 * max(avg_ref, 1e-3)
        # the select itself is exact on the device's own errors (bit-for-bit order statistic, fixed-order sum)
        m2, a2 = _numpy_validation(e)
        assert med == m2
        assert abs(avg - a2) <= 1e-13 * a2
        # device pose state instead of a pose argument
        gp.set_poses(s.init_poses)
        assert gp.validation(s.init_params) == (med, avg)


@pytest.mark.gpu
def test_gpu_validation_ties_and_tiny_inputs(pkg, oracle):
    """every observation of a frame repeated: runs of equal keys straddle both ranks; and N < 100 (len99 = N*99/100)."""
    s = pkg.synth.make_calib("kb4", 3, seed=9, noise_px=0.3)
    for rep, take in ((7, None), (1, 37), (1, 1)):
        n = s.n_obs if take is None else take
        idx = np.tile(np.arange(n), rep)
        fo = np.array([0, len(idx)], dtype=np.int32)
        f0 = np.zeros(1, dtype=np.int64)
        args = [a[:n][idx] if take is not None else np.tile(a, rep) for a in (s.x, s.y, s.z, s.u, s.v)]
        # all observations use frame 0's pose: large errors for the others, still a valid ordering problem
        op = oracle.OracleProblem(pkg.MODELS["kb4"], s.width, s.height, fo, *args)
        med_ref, avg_ref, _ = op.validation(s.gt_params, s.gt_poses[:1])
        with pkg.Problem("kb4", s.width, s.height, fo, *args) as gp:
            med, avg = gp.validation(s.gt_params, s.gt_poses[:1])
        assert abs(med - med_ref) <= 1e-9 * max(med_ref, 1e-3)
        assert abs(avg - avg_ref) <= 1e-9 * max(avg_ref, 1e-3)


@pytest.mark.gpu
def test_validation_mirror_after_calibration(pkg, oracle):
    """calib_camera -> validation like the reference binary (src/bin/camera_calibration.rs): sub-pixel errors."""
    s = pkg.synth.make_calib("eucm", 30, seed=6, noise_px=0.05)
    frames, init = pkg.synth.to_frame_features(s)
    cam0 = pkg.GenericModel("eucm", s.init_params.copy(), s.width, s.height)
    out = pkg.calib_camera(frames, cam0, False, 0, False, init)
    assert out is not None
    cam, rtvecs = out
    med, avg = pkg.validation(0, cam, rtvecs, frames)
    assert 0.0 < med < 0.15 and 0.0 < avg < 0.15

"""Error behaviour of the CUDA path through the C ABI: the two cases in which tiny-solver's optimize() returns None
(LLT failure, NaN error) must surface as CCRS_ERR_CHOLESKY / CCRS_ERR_NUMERIC like in the oracle, bad
arguments as CCRS_ERR_INVALID, and a failed solve must leave the handle usable."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _degenerate(pkg):
    """12 identical board points: the frame's pose block is singular (same fixture as tests/test_controller_cpu.py)."""
    fo = np.array([0, 12], dtype=np.int32)
    z = np.zeros(12); x = np.full(12, 0.1); y = np.full(12, 0.2)
    u = np.full(12, 500.0); v = np.full(12, 500.0)
    return fo, x, y, z, u, v, np.array([[0.1, 0.0, 0.0, 0.0, 0.0, 0.5]])


def test_cholesky_failure_is_reported_like_the_oracle(pkg, oracle):
    fo, x, y, z, u, v, pose = _degenerate(pkg)
    prm = np.array(pkg.synth.GT_PARAMS["eucm"], dtype=np.float64)
    _, _, res, _ = oracle.OracleProblem(1, 1024, 1024, fo, x, y, z, u, v).gauss_newton(prm, pose)
    assert res.status == -2                                   # the oracle's "LLT failed" (tiny-solver: None)
    with pkg.Problem("eucm", 1024, 1024, fo, x, y, z, u, v) as gp:
        gp.set_poses(pose)
        _, summ, _ = gp.solve_gn(prm)
        assert summ.status == -5                              # CCRS_ERR_CHOLESKY
        gp.set_poses(pose)
        _, summ, _ = gp.solve_lm(prm)                          # LM's damping may regularise the singular block: no hang,
        assert summ.status in (0, -4, -5)                      # and a definite status either way
        # the handle survives: a step-wise call still works and sees the NaN poison, not a hang
        gp.set_poses(pose)
        gp.linearize(prm)
        red = gp.reduce(0)
        assert np.isnan(red["S"]).any()
    # and the mirror returns None exactly where calib_camera's `?` would (util.rs:457)
    feats = {k: pkg.FeaturePoint((500.0, 500.0), (0.1, 0.2, 0.0)) for k in range(12)}
    frames = [pkg.FrameFeature(0, (1024, 1024), feats)]
    cam = pkg.GenericModel("eucm", prm, 1024, 1024)
    assert pkg.calib_camera(frames, cam, False, 0, False, {0: pkg.RvecTvec((0.1, 0.0, 0.0), (0.0, 0.0, 0.5))}) is None


def test_nan_error_is_reported(pkg):
    s = pkg.synth.make_calib("eucm", 6, seed=1)
    u = s.u.copy(); u[5] = np.nan                              # a NaN detection poisons the error: optimize() -> None
    with pkg.Problem("eucm", s.width, s.height, s.frame_offsets, s.x, s.y, s.z, u, s.v) as gp:
        gp.set_poses(s.init_poses)
        _, summ, _ = gp.solve_gn(s.init_params)
        assert summ.status in (-4, -5)                         # CCRS_ERR_NUMERIC (or the poisoned pivot seen first)
        gp.set_poses(s.init_poses)
        _, summ, _ = gp.solve_lm(s.init_params)
        assert summ.status in (-4, -5)


def test_invalid_arguments(pkg):
    lib = pkg._abi.load()
    s = pkg.synth.make_calib("eucm", 3, seed=2)
    h = C.c_void_p()
    ip = s.frame_offsets.ctypes.data_as(C.POINTER(C.c_int32))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    assert lib.ccrs_problem_create(C.byref(h), 17, 1024, 1024, 0, 3, ip, dp(s.x), dp(s.y), dp(s.z), dp(s.u), dp(s.v), 1.0, 0) == -1
    assert b"model" in lib.ccrs_last_error().lower() or len(lib.ccrs_last_error()) > 0
    assert lib.ccrs_problem_create(C.byref(h), 1, 1024, 1024, 0, 3, ip, None, dp(s.y), dp(s.z), dp(s.u), dp(s.v), 1.0, 0) == -1
    with pkg.Problem.from_synth(s) as gp:
        with pytest.raises(pkg.CcrsError):                     # use_scale before any Jacobi scaling exists
            gp.reduce(0, u=1e-4, use_scale=True)
        med, avg = C.c_double(), C.c_double()
        assert lib.ccrs_validation(gp.h, None, None, C.byref(med), C.byref(avg), None) == -1
    # pose initialisation needs at least 4 points per frame
    fo = np.array([0, 3], dtype=np.int32); a = np.zeros(3)
    with pytest.raises(pkg.CcrsError):
        pkg.init_poses(fo, a, a, a, a, a)
    # convert_model: disabled distortions out of range
    src = np.array(pkg.synth.GT_PARAMS["kb4"], dtype=np.float64); tgt = np.zeros(6); p = np.ones(8)
    assert lib.ccrs_convert_model(3, dp(src), 1, dp(tgt), 1024, 1024, 5, 8, dp(p), dp(p), dp(p), None, None, 0) == -1

"""Joint multi-camera refinement (calib_all_camera_with_extrinsics, src/util.rs:567-715; OtherCamReprojectionFactor,
src/optimization/factors.rs:204-228). CPU part pins the oracle; GPU part checks the CUDA joint path against it."""
import numpy as np
import pytest

from helpers import rel_err_rows


def _rig(pkg, model="eucm", n_frames=24, n_cams=2, seed=4):
    return pkg.synth.make_rig(model, n_frames, n_cams, seed=seed)


# ----------------------------------------------------------------------------- CPU: oracle
def test_oracle_joint_blocks_match_single_camera_factors(pkg, oracle):
    """cam0 blocks == ReprojectionFactor; cam1 blocks == OtherCamReprojectionFactor (isometry chain), both already
    pinned by tests/test_oracle_pins.py."""
    rig = _rig(pkg)
    oj = oracle.OracleJoint(rig, 1)
    r, J = oj.eval_rj(rig.init_params, rig.init_extr, rig.init_poses, apply_loss=True)
    d = oj.d
    for b in range(len(rig.block_cam)):
        c, f = int(rig.block_cam[b]), int(rig.block_frame[b])
        a, e = int(rig.block_offsets[b]), int(rig.block_offsets[b + 1])
        fo = np.array([0, e - a], dtype=np.int32)
        op = oracle.OracleProblem(1, rig.width, rig.height, fo, rig.x[a:e], rig.y[a:e], rig.z[a:e], rig.u[a:e], rig.v[a:e])
        if c == 0:
            r1, J1 = op.eval_rj(rig.init_params[0], rig.init_poses[f][None], apply_loss=True)
            assert np.array_equal(r[2 * a:2 * e], r1)
            assert np.array_equal(J[2 * a:2 * e, :d + 6], J1)
            assert np.all(J[2 * a:2 * e, d + 6:] == 0.0)
        else:
            r1, J1 = op.othercam_rj(rig.init_params[c], rig.init_poses[f][None], rig.init_extr[c], apply_loss=True)
            assert np.array_equal(r[2 * a:2 * e], r1)
            assert np.array_equal(J[2 * a:2 * e], J1)


def test_oracle_othercam_jacobian_finite_differences(pkg, oracle):
    rig = _rig(pkg, n_frames=6)
    oj = oracle.OracleJoint(rig, 1, huber_delta=0.0)
    r0, J = oj.eval_rj(rig.init_params, rig.init_extr, rig.init_poses, apply_loss=False)
    d = oj.d
    cam1 = np.repeat(rig.block_cam, np.diff(rig.block_offsets)) == 1
    rows = np.repeat(cam1, 2)
    eps = 1e-7
    for i in range(6):                                     # extrinsic columns d+6 .. d+11
        e = rig.init_extr.copy(); e[1, i] += eps
        fd = (oj.eval_rj(rig.init_params, e, rig.init_poses, apply_loss=False)[0] - r0) / eps
        assert np.allclose(fd[rows], J[rows, d + 6 + i], rtol=2e-4, atol=2e-4 * np.abs(J[rows, d + 6 + i]).max())


def test_oracle_joint_gn_recovers_ground_truth(pkg, oracle):
    rig = _rig(pkg, n_frames=40)
    oj = oracle.OracleJoint(rig, 1)
    a, e, p, res, hist = oj.gauss_newton(rig.init_params, rig.init_extr, rig.init_poses)
    assert res.status == 0 and res.iterations <= 8
    assert np.max(np.abs(a - rig.gt_params) / np.abs(rig.gt_params)) < 1e-5
    assert np.max(np.abs(e[1] - rig.gt_extr[1])) < 1e-6
    assert np.all(e[0] == 0.0)


def test_make_rig_keeps_only_detected_frames(pkg):
    """a frame no camera detected is not a variable (util.rs:588-600): the generator drops it — with 200 frames and a
    10 % miss rate per camera a couple of frames always go, which used to leave singular pose blocks behind."""
    rig = pkg.synth.make_rig("kb4", 200, 2, seed=4)
    assert rig.n_frames < 200
    assert sorted(set(rig.block_frame.tolist())) == list(range(rig.n_frames))
    assert rig.gt_poses.shape == (rig.n_frames, 6) and rig.init_poses.shape == (rig.n_frames, 6)


# ----------------------------------------------------------------------------- GPU
@pytest.mark.gpu
@pytest.mark.parametrize("one_focal", [False, True])
@pytest.mark.parametrize("model", ["eucm", "kb4", "opencv5", "ucm", "eucmt", "ftheta"])
def test_gpu_joint_eval_rj_matches_autodiff(pkg, oracle, model, one_focal):
    """a3: OtherCamReprojectionFactor residual + Jacobian (2 x (d+12)) within 1e-9 relative."""
    rig = _rig(pkg, model=model, n_frames=12)
    oj = oracle.OracleJoint(rig, pkg.MODELS[model], xy_same_focal=one_focal)
    gj = pkg.JointProblem.from_rig(rig, xy_same_focal=one_focal)
    intr = np.stack([pkg.synth.intr_from_full(q, one_focal) for q in rig.init_params])
    for loss in (True, False):
        r_ref, J_ref = oj.eval_rj(intr, rig.init_extr, rig.init_poses, apply_loss=loss)
        r, J = gj.eval_rj(intr, rig.init_extr, rig.init_poses, apply_loss=loss)
        assert np.max(np.abs(r - r_ref) / np.maximum(np.abs(r_ref), 1e-3)) < 1e-9
        assert np.max(rel_err_rows(J, J_ref)) < 1e-9
    gj.close()


@pytest.mark.gpu
@pytest.mark.parametrize("model,n_cams", [("eucm", 2), ("kb4", 2), ("eucm", 3)])
def test_gpu_joint_gauss_newton_matches_oracle(pkg, oracle, model, n_cams):
    """BASELINE config 5 (joint cam0+cam1 extrinsic refinement): same loop, equal iteration count."""
    rig = pkg.synth.make_rig(model, 60, n_cams, seed=5)
    oj = oracle.OracleJoint(rig, pkg.MODELS[model])
    gj = pkg.JointProblem.from_rig(rig)
    a_ref, e_ref, p_ref, res, hist_ref = oj.gauss_newton(rig.init_params, rig.init_extr, rig.init_poses)
    a, e, p, summ, hist = gj.solve_gn(rig.init_params, rig.init_extr, rig.init_poses)
    assert summ.status == 0 and res.status == 0
    assert summ.iterations == res.iterations and summ.stop_reason == res.stop_reason
    assert np.max(np.abs(a - a_ref) / np.abs(a_ref)) < 1e-6               # north_star tolerance
    assert np.max(np.abs(e - e_ref)) < 1e-8 and np.max(np.abs(p - p_ref)) < 1e-7
    assert np.allclose(hist, hist_ref, rtol=1e-6, atol=1e-9)
    gj.close()


@pytest.mark.gpu
def test_gpu_joint_bounds_fixed_and_mirror(pkg, oracle):
    """calib_all_camera_with_extrinsics mirror: disabled distortion + cam0 fixed focal (util.rs:654-667)."""
    rig = pkg.synth.make_rig("kb4", 30, 2, seed=6)
    cams = [pkg.GenericModel("kb4", rig.init_params[c], rig.width, rig.height) for c in range(2)]
    t_i_0 = [pkg.RvecTvec(tuple(rig.init_extr[c, :3]), tuple(rig.init_extr[c, 3:])) for c in range(2)]
    frames = [[None] * rig.n_frames for _ in range(2)]
    rtvecs = [dict(), dict()]
    R1 = pkg.synth.rodrigues(rig.init_extr[1, :3]); t1 = rig.init_extr[1, 3:]
    import cv2
    for b in range(len(rig.block_cam)):
        c, f = int(rig.block_cam[b]), int(rig.block_frame[b])
        a, e = int(rig.block_offsets[b]), int(rig.block_offsets[b + 1])
        feats = {k: pkg.FeaturePoint((rig.u[a + k], rig.v[a + k]), (rig.x[a + k], rig.y[a + k], rig.z[a + k])) for k in range(e - a)}
        frames[c][f] = pkg.FrameFeature(0, (rig.width, rig.height), feats)
        if c == 0:
            rtvecs[0][f] = pkg.RvecTvec(tuple(rig.init_poses[f, :3]), tuple(rig.init_poses[f, 3:]))
        else:   # per-camera pose T_1_b = T_1_0 * T_0_b (what calib_camera of cam1 would have produced)
            R0 = pkg.synth.rodrigues(rig.init_poses[f, :3]); t0 = rig.init_poses[f, 3:]
            R = R1 @ R0; t = R1 @ t0 + t1
            rtvecs[1][f] = pkg.RvecTvec(tuple(cv2.Rodrigues(R)[0].ravel()), tuple(t))
    out = pkg.calib_all_camera_with_extrinsics(cams, t_i_0, rtvecs, frames, False, 2, True)
    assert out is not None
    cams_out, t_out, board = out
    assert cams_out[0].params[0] == cams[0].params[0]                      # cam0 focal fixed
    assert np.all(cams_out[0].params[-2:] == 0.0) and np.all(cams_out[1].params[-2:] == 0.0)
    assert t_out[0].rvec == (0.0, 0.0, 0.0)
    # oracle with the same bounds / fixed mask and the same start
    lo, hi = pkg.model_bounds("kb4", rig.width, rig.height)
    fixed = np.zeros((2, 8), dtype=np.uint8); fixed[:, -2:] = 1; fixed[0, 0] = 1
    intr0 = rig.init_params.copy(); intr0[:, -2:] = 0.0
    frame_ids = sorted(set(rtvecs[0]) | set(rtvecs[1]))
    assert frame_ids == list(range(rig.n_frames)) or len(frame_ids) <= rig.n_frames
    sub = {f: i for i, f in enumerate(frame_ids)}
    rig2 = pkg.synth.make_rig("kb4", 30, 2, seed=6)
    rig2.block_frame = np.array([sub[int(f)] for f in rig.block_frame], dtype=np.int32); rig2.n_frames = len(frame_ids)
    oj = oracle.OracleJoint(rig2, 3)
    poses0 = np.stack([board_init for board_init in [rig.init_poses[f] for f in frame_ids]])
    a_ref, e_ref, _, res, _ = oj.gauss_newton(intr0, rig.init_extr, poses0, np.tile(lo, (2, 1)), np.tile(hi, (2, 1)), fixed)
    assert res.status == 0
    got = np.stack([c.params for c in cams_out])
    nz = np.abs(a_ref) > 0
    assert np.max(np.abs(got[nz] - a_ref[nz]) / np.abs(a_ref[nz])) < 1e-5   # starts differ by the T_1_b round trip (1e-12)
    assert np.max(np.abs(np.array([*t_out[1].rvec, *t_out[1].tvec]) - e_ref[1])) < 1e-7

"""batched initial board poses (the sqpnp_solve_glam step of calib_camera, src/util.rs:418-439).
The CUDA path (one warp per frame, 32 Newton runs on SO(3)) against oracle/pnp_oracle.py, which restates the SQPnP
objective with explicit matrices and minimises it with scipy; and against the generating poses."""
import numpy as np
import pytest


def _normalised(pkg, s, poses, noise=0.0, seed=0):
    """exact normalised image points of the synthetic problem's board points under `poses`."""
    R = pkg.synth.rodrigues(poses[:, :3])
    fi = np.repeat(np.arange(s.n_frames), np.diff(s.frame_offsets))
    p = np.stack([s.x, s.y, s.z], axis=1)
    Pc = np.einsum("nij,nj->ni", R[fi], p) + poses[fi, 3:]
    xn, yn = Pc[:, 0] / Pc[:, 2], Pc[:, 1] / Pc[:, 2]
    if noise > 0:
        rng = np.random.default_rng(seed)
        xn = xn + rng.normal(scale=noise, size=xn.shape); yn = yn + rng.normal(scale=noise, size=yn.shape)
    return xn, yn


def test_pnp_oracle_recovers_generating_pose(pkg):
    import pnp_oracle
    s = pkg.synth.make_calib("eucm", 3, seed=5)
    xn, yn = _normalised(pkg, s, s.gt_poses)
    for f in range(s.n_frames):
        a, b = s.frame_offsets[f], s.frame_offsets[f + 1]
        p3 = np.stack([s.x[a:b], s.y[a:b], s.z[a:b]], axis=1)
        rv, t, cost = pnp_oracle.solve_frame(p3, xn[a:b], yn[a:b], n_starts=16)
        assert cost < 1e-12                                            # BFGS stops on its gradient tolerance
        assert np.max(np.abs(rv - s.gt_poses[f, :3])) < 1e-6 and np.max(np.abs(t - s.gt_poses[f, 3:])) < 1e-6


@pytest.mark.gpu
def test_gpu_init_poses_exact_data(pkg):
    """exact normalised points: the generating pose is the unique zero of the cost — 2,000 frames in one launch."""
    s = pkg.synth.make_calib("eucm", 2000, seed=6, drop_fraction=0.2)
    xn, yn = _normalised(pkg, s, s.gt_poses)
    poses, cost = pkg.init_poses(s.frame_offsets, s.x, s.y, s.z, xn, yn, want_cost=True)
    assert np.max(np.abs(cost)) < 1e-12          # r^T Omega r cancels to rounding level at the exact pose
    assert np.max(np.abs(poses - s.gt_poses)) < 1e-7
    # bitwise reproducible
    assert np.array_equal(poses, pkg.init_poses(s.frame_offsets, s.x, s.y, s.z, xn, yn))


@pytest.mark.gpu
def test_gpu_init_poses_matches_oracle_on_noisy_data(pkg):
    import pnp_oracle
    s = pkg.synth.make_calib("kb4", 12, seed=7, drop_fraction=0.3)
    xn, yn = _normalised(pkg, s, s.gt_poses, noise=2e-3, seed=1)     # ~0.8 px at f = 380
    xn = xn.astype(np.float32).astype(np.float64); yn = yn.astype(np.float32).astype(np.float64)   # glam::Vec2 (util.rs:425)
    poses, cost = pkg.init_poses(s.frame_offsets, s.x, s.y, s.z, xn, yn, want_cost=True)
    for f in range(s.n_frames):
        a, b = s.frame_offsets[f], s.frame_offsets[f + 1]
        p3 = np.stack([s.x[a:b], s.y[a:b], s.z[a:b]], axis=1)
        rv, t, c = pnp_oracle.solve_frame(p3, xn[a:b], yn[a:b], n_starts=24, seed=f)
        assert abs(cost[f] - c) <= 1e-9 * c                          # the same (global) minimum
        assert np.max(np.abs(poses[f, :3] - rv)) < 1e-6 and np.max(np.abs(poses[f, 3:] - t)) < 1e-6
        assert np.max(np.abs(poses[f] - s.gt_poses[f])) < 0.05       # and it is near the truth


@pytest.mark.gpu
def test_gpu_init_poses_large_rotations(pkg):
    """boards rotated by up to ~180 degrees about the optical axis and tilted: every start basin is exercised,
    including the axis-angle extraction near pi."""
    rng = np.random.default_rng(3)
    board = pkg.synth.aprilgrid_board().astype(np.float64)
    centre = board.mean(axis=0)
    n = 200
    roll = np.stack([np.zeros(n), np.zeros(n), rng.uniform(-np.pi, np.pi, n)], axis=1)
    roll[:8, 2] = [np.pi, -np.pi, np.pi - 1e-9, np.pi - 1e-5, 3.1, -3.1, 0.0, 1e-10]
    tilt = rng.normal(scale=0.25, size=(n, 3)); tilt[:8] = 0.0
    R = pkg.synth.rodrigues(tilt) @ pkg.synth.rodrigues(roll)
    c = np.stack([rng.uniform(-0.2, 0.2, n), rng.uniform(-0.2, 0.2, n), rng.uniform(0.4, 0.9, n)], axis=1)
    t = c - np.einsum("nij,j->ni", R, centre)
    Pc = np.einsum("nij,kj->nki", R, board) + t[:, None, :]
    xn, yn = (Pc[..., 0] / Pc[..., 2]).ravel(), (Pc[..., 1] / Pc[..., 2]).ravel()
    fo = np.arange(n + 1, dtype=np.int32) * len(board)
    p = np.tile(board, (n, 1))
    poses, cost = pkg.init_poses(fo, p[:, 0], p[:, 1], p[:, 2], xn, yn, want_cost=True)
    Rg = pkg.synth.rodrigues(poses[:, :3])
    assert np.max(np.abs(Rg - R)) < 1e-6 and np.max(np.abs(poses[:, 3:] - t)) < 1e-6
    assert np.max(np.abs(cost)) < 1e-12


@pytest.mark.gpu
def test_initial_poses_feed_calib_camera(pkg):
    """the reference's flow: unproject with the initial model -> pose per frame -> calib_camera (util.rs:401-458)."""
    s = pkg.synth.make_calib("eucm", 40, seed=8, noise_px=0.05)
    frames, _ = pkg.synth.to_frame_features(s)
    cam0 = pkg.GenericModel("eucm", s.init_params.copy(), s.width, s.height)
    init = pkg.initial_poses(frames, cam0)
    assert len(init) == s.n_frames
    got = np.array([init[f].as_array() for f in range(s.n_frames)])
    assert np.max(np.abs(got - s.gt_poses)) < 0.1                    # the initial model is 5 % off: rough poses
    out = pkg.calib_camera(frames, cam0, False, 0, False, init)
    assert out is not None
    assert np.max(np.abs(out[0].params - s.gt_params) / np.abs(s.gt_params)) < 1e-3
    # the reference's own signature (no poses passed): the mirror runs the same initialisation itself
    out2 = pkg.calib_camera(frames, cam0, False, 0, False)
    assert out2 is not None and np.array_equal(out2[0].params, out[0].params)

"""K2 for the models with d >= 8 (EUCMT, KB4, OPENCV5, FTHETA): the FP64 tensor-core variant (ccrs_linmma.cu) against
the oracle, against the register (lane-pair) variant of the same library, run twice for bitwise repeatability; the
table-driven atan2 of KB4 / FTHETA at the angles the synthetic boards never reach; the rotating-replica bench entry."""
import os

import numpy as np
import pytest

from helpers import block_rel_err, rel_err_rows

pytestmark = pytest.mark.gpu

BIG = ["eucmt", "kb4", "opencv5", "ftheta"]


def _blocks(pkg, s, one_focal, mma):
    """frame blocks of one linearisation with the tensor-core variant switched on / off (read at create time)."""
    old = os.environ.get("CCRS_K2_MMA")
    os.environ["CCRS_K2_MMA"] = "1" if mma else "0"
    try:
        gp = pkg.Problem.from_synth(s, xy_same_focal=one_focal)
    finally:
        if old is None:
            del os.environ["CCRS_K2_MMA"]
        else:
            os.environ["CCRS_K2_MMA"] = old
    gp.set_poses(s.init_poses)
    sq = gp.linearize(pkg.synth.intr_from_full(s.init_params, one_focal)).copy()
    B = gp.frame_blocks().copy()
    d = gp.d
    gp.close()
    return B, sq, d


@pytest.mark.parametrize("one_focal", [False, True])
@pytest.mark.parametrize("model", BIG)
def test_tensor_core_blocks_match_register_variant_and_oracle(pkg, oracle, model, one_focal):
    # ragged frames (20 % of the corners dropped: rounds of 32 end anywhere, odd counts included), 37 = 2 x 16 + 5 frames
    s = pkg.synth.make_calib(model, 37, seed=11, drop_fraction=0.2)
    B_mma, sq_mma, d = _blocks(pkg, s, one_focal, mma=True)
    B_reg, sq_reg, _ = _blocks(pkg, s, one_focal, mma=False)
    assert block_rel_err(B_mma, B_reg, d + 7) < 1e-11          # two summation orders of the same products
    assert abs(sq_mma[0] - sq_reg[0]) <= 1e-11 * abs(sq_reg[0])
    op = oracle.OracleProblem.from_synth(s, pkg.MODELS[model], xy_same_focal=one_focal)
    _, B_ref = op.linearize(pkg.synth.intr_from_full(s.init_params, one_focal), s.init_poses)
    assert block_rel_err(B_mma, B_ref, d + 7) < 1e-9            # north_star tolerance
    B_again, sq_again, _ = _blocks(pkg, s, one_focal, mma=True)
    assert np.array_equal(B_mma, B_again) and np.array_equal(sq_mma, sq_again)   # dynamic frame hand-out, fixed sums


@pytest.mark.parametrize("model", ["kb4", "ftheta"])
def test_table_driven_atan2_at_wide_angles(pkg, oracle, model):
    """theta = atan2(r, z) from 1e-7 rad to beyond 90 degrees (z < 0), table boundaries included: r and J against the
    oracle's dual numbers (std::atan2) at north_star's 1e-9."""
    th = np.concatenate([np.array([1e-7, 1e-5, 1e-3, 1.0 / 64, 2 * np.arctan(0.5 / 64), 2 * np.arctan(1.5 / 64)]),
                         np.linspace(0.02, 2.6, 60)])
    phi = np.linspace(0.0, 6.0, th.size)
    rho = 2.0
    x = (rho * np.sin(th) * np.cos(phi)).astype(np.float32).astype(np.float64)
    y = (rho * np.sin(th) * np.sin(phi)).astype(np.float32).astype(np.float64)
    z = (rho * np.cos(th)).astype(np.float32).astype(np.float64)
    prm = np.array([400.0, 410.0, 512.0, 500.0, 0.01, -0.002, 0.0005, -0.0001])
    uv = np.stack([oracle.project(pkg.MODELS[model], prm, np.array([a, b, c])) for a, b, c in zip(x, y, z)])
    u = (uv[:, 0] + 0.3).astype(np.float32).astype(np.float64)
    v = (uv[:, 1] - 0.2).astype(np.float32).astype(np.float64)
    fo = np.array([0, th.size], dtype=np.int32)
    pose = np.zeros((1, 6)); pose[0, :3] = [1e-3, -2e-3, 1e-3]
    gp = pkg.Problem(model, 1024, 1024, fo, x, y, z, u, v)
    op = oracle.OracleProblem(pkg.MODELS[model], 1024, 1024, fo, x, y, z, u, v)
    r, J = gp.eval_rj(prm, pose, apply_loss=False)
    r_ref, J_ref = op.eval_rj(prm, pose, apply_loss=False)
    assert np.max(np.abs(r - r_ref) / np.maximum(np.abs(r_ref), 1e-3)) < 1e-9
    assert np.max(rel_err_rows(J, J_ref)) < 1e-9
    gp.close()


def test_rotating_replica_bench_entry(pkg):
    s = pkg.synth.make_calib("eucm", 64, seed=3)
    reps = [pkg.Problem.from_synth(s) for _ in range(3)]
    total_ms, launches, executed = pkg.Problem.bench_lm_steps_rotating(reps, s.init_params, s.init_poses, warmup=2, steps=17)
    assert total_ms > 0.0 and launches == 34 and executed == 17   # K2 + K3 per timed slot; 17 > 3 replicas x 4: one restart inside
    # the handles are usable afterwards (their own streams restored)
    reps[1].set_poses(s.init_poses)
    intr, summ, _ = reps[1].solve_lm(s.init_params)
    assert summ.status == 0 and np.max(np.abs(intr - s.gt_params) / np.abs(s.gt_params)) < 1e-6
    for q in reps:
        q.close()


@pytest.mark.parametrize("fmt", ["board", "f32", "f64"])
def test_update_observations_equals_a_fresh_handle(pkg, fmt):
    """ccrs_problem_update_observations: new detections (same frame structure) into an existing handle give, bit for bit,
    what a handle created from them gives; a different total is refused."""
    s = pkg.synth.make_calib("eucm", 48, seed=9, drop_fraction=0.1)
    rng = np.random.default_rng(1)
    f32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
    u2, v2 = f32(s.u + rng.normal(0, 0.2, s.u.shape)), f32(s.v + rng.normal(0, 0.2, s.v.shape))

    def make(u, v):
        if fmt == "board":
            return pkg.Problem("eucm", s.width, s.height, s.frame_offsets, None, None, None, f32(u), f32(v), corner_id=s.extra["corner_id"], board=s.extra["board"])
        if fmt == "f32":
            return pkg.Problem("eucm", s.width, s.height, s.frame_offsets, f32(s.x), f32(s.y), f32(s.z), f32(u), f32(v))
        return pkg.Problem("eucm", s.width, s.height, s.frame_offsets, s.x, s.y, s.z, np.float64(f32(u)), np.float64(f32(v)))

    gp = make(s.u, s.v)
    gp.set_poses(s.init_poses)
    gp.solve_lm(s.init_params)
    kw = dict(corner_id=s.extra["corner_id"]) if fmt == "board" else dict(x=s.x, y=s.y, z=s.z)
    gp.update_observations(s.frame_offsets, u2, v2, **kw)
    gp.set_poses(s.init_poses)
    intr_a, summ_a, _ = gp.solve_lm(s.init_params)
    poses_a = gp.get_poses()
    fresh = make(u2, v2)
    fresh.set_poses(s.init_poses)
    intr_b, summ_b, _ = fresh.solve_lm(s.init_params)
    assert summ_a.status == 0 and summ_a.iterations == summ_b.iterations
    assert np.array_equal(intr_a, intr_b) and np.array_equal(poses_a, fresh.get_poses())
    fo_short = s.frame_offsets.copy(); fo_short[-1] -= 1          # one observation fewer than the handle holds
    n1 = int(fo_short[-1])
    with pytest.raises(pkg.CcrsError):
        gp.update_observations(fo_short, u2[:n1], v2[:n1], **{k: a[:n1] for k, a in kw.items()})
    gp.close(); fresh.close()

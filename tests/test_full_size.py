"""BASELINE.json's configurations at FULL size on the B200 (configs[1]-[4]; the small versions live in
test_gpu_parity.py). The CPU oracle finishes these in seconds with all host cores, so they are checked directly, plus
the size-independent properties of the path: bitwise determinism, cost == sum of the frame blocks' (r,r) entries,
additivity of the reduced system over frame shards, and recovery of the generating parameters."""
import numpy as np
import pytest

from helpers import MODEL_NAMES, block_rel_err, rel_err_rows, rms_px

pytestmark = pytest.mark.gpu

TOL_RJ, TOL_INTR, TOL_RMS = 1e-9, 1e-6, 1e-4     # north_star tolerances


def test_config2_eucm_2000_frames_lm_to_convergence(pkg, oracle):
    """configs[1]: EUCM, 2,000 frames x 144 corners (~288 k observations), full LM to convergence vs the CPU oracle."""
    s = pkg.synth.make_calib("eucm", 2000, seed=1, noise_px=0.1)
    op = oracle.OracleProblem.from_synth(s, 1)
    with pkg.Problem.from_synth(s) as gp:
        gp.set_poses(s.init_poses)
        intr, summ, hist = gp.solve_lm(s.init_params)
        intr_ref, poses_ref, res, hist_ref = op.levenberg_marquardt(s.init_params, s.init_poses)
        assert summ.status == 0 and summ.iterations == res.iterations            # equal iteration count
        assert (summ.n_accepted, summ.n_rejected) == (res.n_accepted, res.n_rejected)
        assert np.max(np.abs(intr - intr_ref) / np.abs(intr_ref)) < TOL_INTR
        assert abs(rms_px(op, intr, gp.get_poses()) - rms_px(op, intr_ref, poses_ref)) < TOL_RMS
        assert np.max(np.abs(intr - s.gt_params) / np.abs(s.gt_params)) < 1e-3   # and it is the right answer


@pytest.mark.parametrize("model", MODEL_NAMES)
def test_config3_six_models_2000_frames_rj(pkg, oracle, model):
    """configs[2]: every model at 2,000 frames: per-observation residual / Jacobian tolerance."""
    s = pkg.synth.make_calib(model, 2000, seed=2)
    op = oracle.OracleProblem.from_synth(s, pkg.MODELS[model])
    with pkg.Problem.from_synth(s) as gp:
        r_ref, J_ref = op.eval_rj(s.init_params, s.init_poses, apply_loss=True)
        r, J = gp.eval_rj(s.init_params, s.init_poses, apply_loss=True)
        assert np.max(np.abs(r - r_ref) / np.maximum(np.abs(r_ref), 1e-3)) < TOL_RJ
        assert np.max(rel_err_rows(J, J_ref)) < TOL_RJ
        # the fused path at the same size: frame blocks against the oracle's
        gp.set_poses(s.init_poses)
        sq = gp.linearize(s.init_params)
        sq_ref, B_ref = op.linearize(s.init_params, s.init_poses)
        assert abs(sq[0] - sq_ref) / sq_ref < 1e-12
        B = gp.frame_blocks()
        idx = np.linspace(0, s.n_frames - 1, 64).astype(int)           # a spread sample keeps the Python check short
        assert block_rel_err(B[idx], B_ref[idx], gp.d + 7) < TOL_RJ


def test_config4_eucm_7000_frames_properties(pkg, oracle):
    """configs[3] on one GPU: ~1 M observations. LM vs the oracle, determinism, and additivity over frame shards."""
    s = pkg.synth.make_calib("eucm", 7000, seed=3)
    op = oracle.OracleProblem.from_synth(s, 1)
    with pkg.Problem.from_synth(s) as gp:
        assert gp.n_obs > 1_000_000
        gp.set_poses(s.init_poses)
        intr, summ, hist = gp.solve_lm(s.init_params)
        poses = gp.get_poses()
        intr_ref, poses_ref, res, _ = op.levenberg_marquardt(s.init_params, s.init_poses)
        assert summ.status == 0 and summ.iterations == res.iterations
        assert np.max(np.abs(intr - intr_ref) / np.abs(intr_ref)) < TOL_INTR
        assert abs(rms_px(op, intr, poses) - rms_px(op, intr_ref, poses_ref)) < TOL_RMS
        assert np.max(np.abs(intr - s.gt_params) / np.abs(s.gt_params)) < 1e-6   # noise-free data: the generator's values
        # bitwise determinism: no floating-point atomics anywhere, every sum has a fixed order
        gp.set_poses(s.init_poses)
        intr2, summ2, hist2 = gp.solve_lm(s.init_params)
        assert np.array_equal(intr, intr2) and np.array_equal(hist, hist2) and np.array_equal(poses, gp.get_poses())
        # cost == sum over frames of the (r,r) block entry; reduced system symmetric
        gp.set_poses(s.init_poses)
        sq = gp.linearize(s.init_params)[0]
        B = gp.frame_blocks()
        assert abs(B[:, -1].sum() - sq) / sq < 1e-13
        red = gp.reduce(0)
        S = red["S"][0]
        assert np.array_equal(S, S.T) and abs(red["sq_err"][0] - sq) / sq < 1e-13
    # additivity: the reduced systems of two frame shards add up to the whole problem's (what the multi-GPU path relies on)
    cut = 3123
    k = int(s.frame_offsets[cut])
    parts = []
    for a, b, ka, kb in ((0, cut, 0, k), (cut, s.n_frames, k, s.n_obs)):
        with pkg.Problem("eucm", s.width, s.height, s.frame_offsets[a:b + 1] - s.frame_offsets[a], s.x[ka:kb], s.y[ka:kb],
                         s.z[ka:kb], s.u[ka:kb], s.v[ka:kb]) as shard:
            shard.set_poses(s.init_poses[a:b])
            shard.linearize(s.init_params)
            parts.append(shard.reduce(0))
    for key in ("S", "g_s", "g_a", "sq_err"):
        whole, summed = red[key][0], parts[0][key][0] + parts[1][key][0]
        assert np.max(np.abs(whole - summed)) <= 1e-11 * np.max(np.abs(whole)), key


def test_config5_batch_of_kb4_calibrations(pkg, oracle):
    """configs[4], one GPU's share reduced to 256 problems x 200 frames (7.4 M observations): batch == standalone, and a
    sample against the oracle."""
    n_problems, n_distinct = 256, 8
    probs = [pkg.synth.make_calib("kb4", 200, seed=100 + i) for i in range(n_distinct)]
    fo, pfo = [np.zeros(1, dtype=np.int64)], [0]
    xs, ys, zs, us, vs, poses, intr0 = [], [], [], [], [], [], []
    for b in range(n_problems):
        s = probs[b % n_distinct]
        fo.append(fo[-1][-1] + s.frame_offsets[1:].astype(np.int64)); pfo.append(pfo[-1] + s.n_frames)
        xs.append(s.x); ys.append(s.y); zs.append(s.z); us.append(s.u); vs.append(s.v)
        poses.append(s.init_poses); intr0.append(s.init_params)
    cat = np.concatenate
    with pkg.Problem("kb4", 1024, 1024, cat(fo).astype(np.int32), cat(xs), cat(ys), cat(zs), cat(us), cat(vs),
                     problem_frame_offsets=np.array(pfo, dtype=np.int32)) as gp:
        gp.set_poses(cat(poses))
        intr, summ, _ = gp.solve_lm(np.stack(intr0))
        assert summ.status == 0
        for b in range(n_distinct):
            s = probs[b]
            assert np.array_equal(intr[b], intr[b + n_distinct * (n_problems // n_distinct - 1)])   # same data, same bits
            op = oracle.OracleProblem.from_synth(s, 3)
            ref, _, res, _ = op.levenberg_marquardt(s.init_params, s.init_poses)
            assert np.max(np.abs(intr[b] - ref) / np.abs(ref)) < TOL_INTR

"""GPU parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on the same
seeded inputs. Tolerances are north_star's: per-observation r/J <= 1e-9 relative; converged intrinsics
<= 1e-6 relative and RMS reprojection error within 1e-4 px at equal iteration count."""
import numpy as np
import pytest

from helpers import MODEL_NAMES, block_rel_err, rel_err_rows, rms_px

pytestmark = pytest.mark.gpu

TOL_RJ = 1e-9       # north_star: per-observation residual/Jacobian within 1e-9 relative
TOL_INTR = 1e-6     # north_star: final intrinsics within 1e-6 relative
TOL_RMS = 1e-4      # north_star: RMS reprojection error within 1e-4 px


def _setup(pkg, oracle, model, n_frames, seed, one_focal=False, **kw):
    s = pkg.synth.make_calib(model, n_frames, seed=seed, **kw)
    mid = pkg.MODELS[model]
    op = oracle.OracleProblem.from_synth(s, mid, xy_same_focal=one_focal)
    gp = pkg.Problem.from_synth(s, xy_same_focal=one_focal)
    intr0 = pkg.synth.intr_from_full(s.init_params, one_focal)
    return s, op, gp, intr0


@pytest.mark.parametrize("one_focal", [False, True])
@pytest.mark.parametrize("model", MODEL_NAMES)
@pytest.mark.parametrize("apply_loss", [True, False])
def test_eval_rj_matches_autodiff(pkg, oracle, model, one_focal, apply_loss):
    s, op, gp, intr0 = _setup(pkg, oracle, model, 40, seed=3, one_focal=one_focal, drop_fraction=0.2)
    r_ref, J_ref = op.eval_rj(intr0, s.init_poses, apply_loss=apply_loss)
    r, J = gp.eval_rj(intr0, s.init_poses, apply_loss=apply_loss)
    # residuals: relative to max(|ref|, 1e-3 px) — observations are O(1e3) px so 1e-12 px absolute is the noise floor
    assert np.max(np.abs(r - r_ref) / np.maximum(np.abs(r_ref), 1e-3)) < TOL_RJ
    assert np.max(rel_err_rows(J, J_ref)) < TOL_RJ
    gp.close()


@pytest.mark.parametrize("model", MODEL_NAMES)
def test_eval_rj_near_ground_truth(pkg, oracle, model):
    """small residuals (no Huber activity) and small rotations: the regime of the last iterations."""
    s, op, gp, _ = _setup(pkg, oracle, model, 30, seed=5)
    poses = s.gt_poses.copy()
    poses[0, :3] = [1e-9, -2e-9, 3e-9]     # tiny but non-zero rotation (series branch)
    poses[1, :3] = [1e-3, 2e-3, -1e-3]
    r_ref, J_ref = op.eval_rj(s.gt_params, poses, apply_loss=True)
    r, J = gp.eval_rj(s.gt_params, poses, apply_loss=True)
    assert np.max(np.abs(r - r_ref) / np.maximum(np.abs(r_ref), 1e-3)) < TOL_RJ
    assert np.max(rel_err_rows(J, J_ref)) < TOL_RJ
    gp.close()


def test_reference_test_reprojection_factor(pkg, oracle):
    """tests/optimization_test.rs:36-80 through the GPU path: UCM [500,500,320,240,0.5] 640x480,
    p3d (1,2,10), rvec = tvec = 0 (exactly-zero rvec: identity branch), then tvec.x += 0.1."""
    prm = np.array([500.0, 500.0, 320.0, 240.0, 0.5])
    p3d = np.array([1.0, 2.0, 10.0], dtype=np.float32).astype(np.float64)
    uv = oracle.project(0, prm, p3d)
    p2d = uv.astype(np.float32).astype(np.float64)
    fo = np.array([0, 1], dtype=np.int32)
    gp = pkg.Problem("ucm", 640, 480, fo, p3d[:1], p3d[1:2], p3d[2:3], p2d[:1], p2d[1:2])
    op = oracle.OracleProblem(0, 640, 480, fo, p3d[:1], p3d[1:2], p3d[2:3], p2d[:1], p2d[1:2])
    pose = np.zeros((1, 6))
    r, J = gp.eval_rj(prm, pose, apply_loss=False)
    assert np.linalg.norm(r) < 1e-4
    r_ref, J_ref = op.eval_rj(prm, pose, apply_loss=False)
    assert np.max(rel_err_rows(J, J_ref)) < TOL_RJ          # includes the zero rvec-derivative columns
    assert np.all(J[:, 5:8] == 0.0)
    pose[0, 3] = 0.1
    r, _ = gp.eval_rj(prm, pose, apply_loss=False)
    assert np.linalg.norm(r) > 1e-3
    assert np.allclose(r, [4.91153299, -0.04993990], atol=1e-6)  # SURVEY §8(c) golden vector
    gp.close()


@pytest.mark.parametrize("one_focal", [False, True])
@pytest.mark.parametrize("model", MODEL_NAMES)
def test_frame_blocks_match_oracle(pkg, oracle, model, one_focal):
    s, op, gp, intr0 = _setup(pkg, oracle, model, 64, seed=7, one_focal=one_focal, drop_fraction=0.3)
    gp.set_poses(s.init_poses)
    sq = gp.linearize(intr0)
    B = gp.frame_blocks()
    sq_ref, B_ref = op.linearize(intr0, s.init_poses)
    assert abs(sq[0] - sq_ref) / sq_ref < 1e-12
    assert block_rel_err(B, B_ref, gp.d + 7) < TOL_RJ
    # cost-only pass agrees with the (r,r) entries
    c = gp.eval_cost(intr0)
    assert abs(c[0] - sq_ref) / sq_ref < 1e-12
    gp.close()


@pytest.mark.parametrize("n_frames", [1, 3, 37, 130, 700])
def test_slicing_is_size_independent(pkg, oracle, n_frames):
    """different frame counts exercise different lanes-per-frame splits (G) of K2."""
    s, op, gp, intr0 = _setup(pkg, oracle, "eucm", n_frames, seed=11, drop_fraction=0.15)
    gp.set_poses(s.init_poses)
    gp.linearize(intr0)
    B = gp.frame_blocks()
    _, B_ref = op.linearize(intr0, s.init_poses)
    assert block_rel_err(B, B_ref, gp.d + 7) < TOL_RJ
    gp.close()


@pytest.mark.parametrize("model", ["eucm", "kb4", "opencv5"])
def test_single_step_matches_oracle(pkg, oracle, model):
    """reduce + host solve + backsub == the oracle's step (GN and damped/scaled LM step)."""
    s, op, gp, intr0 = _setup(pkg, oracle, model, 50, seed=13)
    d = gp.d
    gp.set_poses(s.init_poses)
    gp.linearize(intr0)
    # Gauss-Newton step
    red = gp.reduce(u=None, use_scale=False)
    y = np.linalg.solve(red["S"][0], red["g_s"][0])
    gp.backsub(y, in_place=False, want_model_dec=False)
    st, di, dp, _ = op.solve_step(intr0, s.init_poses, u=0.0)
    assert st == 0
    assert np.max(np.abs(y - di) / np.maximum(np.abs(di), 1e-12 + 1e-6 * np.abs(intr0))) < 1e-6
    gp.accept()
    poses_new = gp.get_poses()
    assert np.max(np.abs((poses_new - s.init_poses) - dp)) < 1e-9
    # LM step with Jacobi scaling
    gp.set_poses(s.init_poses)
    gp.linearize(intr0)
    col = gp.compute_scale()
    sc_a = 1.0 / (1.0 + np.sqrt(col[0]))
    gp.set_intr_scale(sc_a)
    u = 1e-4
    red = gp.reduce(u=u, use_scale=True)
    S = red["S"][0] + u * np.diag(np.clip(red["diag_a"][0], 1e-6, 1e32))
    y = np.linalg.solve(S, red["g_s"][0])
    md_p = gp.backsub(y, u=u, in_place=False)
    md = md_p[0] + y @ red["g_a"][0] + u * np.sum(np.clip(red["diag_a"][0], 1e-6, 1e32) * y * y)
    # oracle with the same scale vector
    _, B_ref = op.linearize(intr0, s.init_poses)
    NA = d + 7
    from helpers import tri_idx
    sc = np.empty(d + 6 * s.n_frames)
    sc[:d] = sc_a
    for f in range(s.n_frames):
        for i in range(6):
            sc[d + 6 * f + i] = 1.0 / (1.0 + np.sqrt(B_ref[f, tri_idx(NA, d + i, d + i)]))
    st, di, dp, md_ref = op.solve_step(intr0, s.init_poses, u=u, scale=sc)
    assert st == 0
    assert np.max(np.abs(sc_a * y - di) / np.maximum(np.abs(di), 1e-6 * np.abs(intr0))) < 1e-6
    assert abs(md - md_ref) / abs(md_ref) < 1e-8
    gp.accept()
    assert np.max(np.abs((gp.get_poses() - s.init_poses) - dp)) < 1e-9
    gp.close()


@pytest.mark.parametrize("model,n_frames", [("eucm", 100), ("ucm", 60), ("eucmt", 60), ("kb4", 60), ("opencv5", 60), ("ftheta", 60)])
def test_gauss_newton_matches_oracle(pkg, oracle, model, n_frames):
    """BASELINE config 1 (EUCM 100x144) and the six-model sweep: the loop the reference runs (GN)."""
    s, op, gp, intr0 = _setup(pkg, oracle, model, n_frames, seed=0)
    gp.set_poses(s.init_poses)
    intr, summ, hist = gp.solve_gn(intr0)
    intr_ref, poses_ref, res, hist_ref = op.gauss_newton(intr0, s.init_poses)
    assert summ.status == 0 and res.status == 0
    assert summ.iterations == res.iterations                   # equal iteration count
    assert summ.stop_reason == res.stop_reason
    assert np.max(np.abs(intr - intr_ref) / np.abs(intr_ref)) < TOL_INTR
    poses = gp.get_poses()
    assert abs(rms_px(op, intr, poses) - rms_px(op, intr_ref, poses_ref)) < TOL_RMS
    assert np.allclose(hist, hist_ref, rtol=1e-6, atol=1e-9)
    gp.close()


@pytest.mark.parametrize("speculative", [1, 0])
@pytest.mark.parametrize("model,n_frames", [("eucm", 100), ("kb4", 60), ("opencv5", 60)])
def test_levenberg_marquardt_matches_oracle(pkg, oracle, model, n_frames, speculative):
    s, op, gp, intr0 = _setup(pkg, oracle, model, n_frames, seed=1, noise_px=0.1)
    gp.set_poses(s.init_poses)
    intr, summ, hist = gp.solve_lm(intr0, options=pkg.default_options(speculative=speculative))
    intr_ref, poses_ref, res, hist_ref = op.levenberg_marquardt(intr0, s.init_poses)
    assert summ.status == 0 and res.status == 0
    assert summ.iterations == res.iterations
    assert (summ.n_accepted, summ.n_rejected) == (res.n_accepted, res.n_rejected)
    assert np.max(np.abs(intr - intr_ref) / np.abs(intr_ref)) < TOL_INTR
    assert abs(rms_px(op, intr, gp.get_poses()) - rms_px(op, intr_ref, poses_ref)) < TOL_RMS
    assert np.allclose(hist, hist_ref, rtol=1e-6, atol=1e-9)
    gp.close()


def test_lm_rejects_and_recovers(pkg, oracle):
    """a start far enough away that LM rejects steps: same accept/reject sequence as the oracle."""
    s, op, gp, intr0 = _setup(pkg, oracle, "eucm", 40, seed=2)
    intr_bad = intr0 * np.array([0.7, 0.7, 1.05, 0.95, 0.7, 1.5])
    poses = s.init_poses + np.random.default_rng(1).normal(scale=0.2, size=s.init_poses.shape)
    gp.set_poses(poses)
    intr, summ, hist = gp.solve_lm(intr_bad)
    intr_ref, poses_ref, res, hist_ref = op.levenberg_marquardt(intr_bad, poses)
    assert res.n_rejected > 0 and res.status == 0
    assert summ.iterations == res.iterations
    assert (summ.n_accepted, summ.n_rejected) == (res.n_accepted, res.n_rejected)
    assert np.max(np.abs(intr - intr_ref) / np.abs(intr_ref)) < TOL_INTR
    gp.close()


def test_bounds_and_fixed_variables(pkg, oracle):
    """set_variable_bounds / fix_variable semantics (util.rs:29-71): clamp, then fixed indices keep old value."""
    s, op, gp, intr0 = _setup(pkg, oracle, "kb4", 50, seed=4)
    lo, hi = pkg.model_bounds("kb4", s.width, s.height)
    fixed = np.zeros(gp.d, dtype=np.uint8)
    fixed[-2:] = 1                      # --disabled-distortion-num 2 (docs/tutorial.md:16-19)
    intr0 = intr0.copy(); intr0[-2:] = 0.0
    for mode in (0, 1):
        gp.set_poses(s.init_poses)
        intr, summ, _ = gp.solve_gn(intr0, lo, hi, fixed, options=pkg.default_options(fixed_mode=mode))
        intr_ref, _, res, _ = op.gauss_newton(intr0, s.init_poses, lo, hi, fixed, options=op.default_options(fixed_mode=mode))
        assert summ.iterations == res.iterations
        assert np.all(intr[-2:] == 0.0)
        assert np.max(np.abs(intr[:-2] - intr_ref[:-2]) / np.abs(intr_ref[:-2])) < TOL_INTR
    gp.close()


def test_polynomial_coefficients_are_not_clamped(pkg, oracle):
    """A lens whose distortion coefficient exceeds 1 (real wide-angle KB4 / plumb-bob lenses do) must calibrate to that
    value under the model's own bounds: only alpha / beta of the unified models are boxed, polynomial coefficients are
    not (no invented [-1, 1] clamp)."""
    gt = np.array([380.0, 380.0, 512.0, 512.0, 1.5, -0.3, 0.05, -0.004])
    s = pkg.synth.make_calib("kb4", 60, seed=11, gt_params=gt)
    op = oracle.OracleProblem.from_synth(s, pkg.MODELS["kb4"])
    lo, hi = pkg.model_bounds("kb4", s.width, s.height)
    assert np.all(np.isinf(lo[4:])) and np.all(np.isinf(hi[4:]))
    with pkg.Problem.from_synth(s) as gp:
        gp.set_poses(s.init_poses)
        intr, summ, _ = gp.solve_gn(s.init_params, lo, hi)
        ref, _, res, _ = op.gauss_newton(s.init_params, s.init_poses, lo, hi)
        assert summ.status == 0 and summ.iterations == res.iterations
        assert np.max(np.abs(intr - ref) / np.abs(ref)) < TOL_INTR
        assert abs(intr[4] - 1.5) < 1e-4 and np.max(np.abs(intr - gt) / np.abs(gt)) < 1e-3


def test_batch_equals_independent_problems(pkg, oracle):
    """BASELINE config 5 in miniature: a batch of independent KB4 calibrations in one handle."""
    probs = [pkg.synth.make_calib("kb4", nf, seed=20 + i, drop_fraction=0.1) for i, nf in enumerate([12, 30, 7, 21])]
    fo, pfo, xs, ys, zs, us, vs, poses, intr0 = [0], [0], [], [], [], [], [], [], []
    for s in probs:
        fo += list(fo[-1] + s.frame_offsets[1:])
        pfo.append(pfo[-1] + s.n_frames)
        xs.append(s.x); ys.append(s.y); zs.append(s.z); us.append(s.u); vs.append(s.v)
        poses.append(s.init_poses); intr0.append(s.init_params)
    cat = np.concatenate
    gp = pkg.Problem("kb4", 1024, 1024, np.array(fo, dtype=np.int32), cat(xs), cat(ys), cat(zs), cat(us), cat(vs),
                     problem_frame_offsets=np.array(pfo, dtype=np.int32))
    assert gp.n_problems == 4
    gp.set_poses(cat(poses))
    intr0 = np.stack(intr0)
    # blocks of the batch == blocks of each problem on its own
    gp.linearize(intr0)
    B = gp.frame_blocks()
    for i, s in enumerate(probs):
        op = oracle.OracleProblem.from_synth(s, 3)
        _, B_ref = op.linearize(s.init_params, s.init_poses)
        assert block_rel_err(B[pfo[i]:pfo[i + 1]], B_ref, gp.d + 7) < TOL_RJ
    for solver, oname in (("solve_gn", "gauss_newton"), ("solve_lm", "levenberg_marquardt")):
        gp.set_poses(cat(poses))
        intr, summ, _ = getattr(gp, solver)(intr0)
        assert summ.status == 0
        P = gp.get_poses()
        for i, s in enumerate(probs):
            op = oracle.OracleProblem.from_synth(s, 3)
            intr_ref, poses_ref, res, _ = getattr(op, oname)(s.init_params, s.init_poses)
            assert np.max(np.abs(intr[i] - intr_ref) / np.abs(intr_ref)) < TOL_INTR, (solver, i)
            assert np.max(np.abs(P[pfo[i]:pfo[i + 1]] - poses_ref)) < 1e-6
    gp.close()


def test_calib_camera_mirror(pkg, oracle):
    """calib_camera entry point (util.rs:384-490) incl. --one-focal / --disabled-distortion-num / --fixed-focal."""
    s = pkg.synth.make_calib("eucm", 30, seed=6)
    frames, init = [], {}
    for f in range(s.n_frames):
        a, b = s.frame_offsets[f], s.frame_offsets[f + 1]
        feats = {k: pkg.FeaturePoint((s.u[a + k], s.v[a + k]), (s.x[a + k], s.y[a + k], s.z[a + k])) for k in range(b - a)}
        frames.append(pkg.FrameFeature(0, (s.width, s.height), feats))
        init[f] = pkg.RvecTvec(tuple(s.init_poses[f, :3]), tuple(s.init_poses[f, 3:]))
    frames.insert(3, None)  # a frame without detection (Option::None in the reference)
    init = {(k if k < 3 else k + 1): v for k, v in init.items()}
    cam0 = pkg.GenericModel("eucm", s.init_params, s.width, s.height)
    out = pkg.calib_camera(frames, cam0, False, 0, False, init)
    assert out is not None
    cam, rt = out
    assert np.max(np.abs(cam.params - s.gt_params) / np.abs(s.gt_params)) < 1e-5
    assert 3 not in rt and len(rt) == s.n_frames
    # one focal: fy := fx re-inserted (util.rs:466-470)
    cam1, _ = pkg.calib_camera(frames, cam0, True, 0, False, init)
    assert cam1.params[0] == cam1.params[1]
    # fixed focal: second pass with params[0] reset to the input focal (util.rs:459-464)
    cam2, _ = pkg.calib_camera(frames, cam0, False, 0, True, init)
    assert cam2.params[0] == cam0.params[0]
    # oracle equivalents of the two-pass semantics
    op = oracle.OracleProblem.from_synth(s, 1)
    lo, hi = pkg.model_bounds("eucm", s.width, s.height)
    i1, p1, _, _ = op.gauss_newton(s.init_params, s.init_poses, lo, hi)
    i1[0] = s.init_params[0]
    fixed = np.zeros(6, dtype=np.uint8); fixed[0] = 1
    i2, _, _, _ = op.gauss_newton(i1, p1, lo, hi, fixed)
    assert np.max(np.abs(cam2.params - i2) / np.abs(i2)) < TOL_INTR


def test_determinism_bitwise(pkg):
    """fixed-order reductions: two runs give bit-identical blocks and reduced systems."""
    s = pkg.synth.make_calib("eucm", 300, seed=9)
    outs = []
    for _ in range(2):
        gp = pkg.Problem.from_synth(s)
        gp.set_poses(s.init_poses)
        gp.linearize(s.init_params)
        red = gp.reduce()
        outs.append((gp.frame_blocks().copy(), red["S"].copy(), red["g_s"].copy()))
        gp.close()
    for a, b in zip(outs[0], outs[1]):
        assert np.array_equal(a, b)


def test_f32_observation_api_is_bit_identical(pkg):
    """ccrs_problem_create_f32: FeaturePoint's f32 values (detected_points.rs:6-9) widened on load (factors.rs:141-143)
    must give exactly the results of the f64 entry point on the widened arrays."""
    s = pkg.synth.make_calib("eucm", 120, seed=8, drop_fraction=0.2)
    g64 = pkg.Problem.from_synth(s)
    f = lambda a: a.astype(np.float32)
    assert all(np.array_equal(f(a).astype(np.float64), a) for a in (s.x, s.y, s.z, s.u, s.v))   # f32-representable inputs
    g32 = pkg.Problem(s.model, s.width, s.height, s.frame_offsets, f(s.x), f(s.y), f(s.z), f(s.u), f(s.v))
    for g in (g64, g32):
        g.set_poses(s.init_poses)
    assert np.array_equal(g64.linearize(s.init_params), g32.linearize(s.init_params))
    assert np.array_equal(g64.frame_blocks(), g32.frame_blocks())
    r64, J64 = g64.eval_rj(s.init_params, s.init_poses)
    r32, J32 = g32.eval_rj(s.init_params, s.init_poses)
    assert np.array_equal(r64, r32) and np.array_equal(J64, J32)
    i64, s64, _ = g64.solve_lm(s.init_params)
    i32, s32, _ = g32.solve_lm(s.init_params)
    assert np.array_equal(i64, i32) and s64.iterations == s32.iterations
    g64.close(); g32.close()


_SOLVE_CODE = ("import importlib,json,numpy as np;pkg=importlib.import_module('camera-intrinsic-calibration-rs_b200');"
               "import ctypes as C;lib=pkg._abi.load();"
               "s=pkg.synth.make_calib('eucm',100,seed=1,noise_px=0.1);gp=pkg.Problem.from_synth(s);gp.set_poses(s.init_poses);"
               "intr,summ,hist=gp.%s(s.init_params);a,h,n=C.c_int64(0),C.c_int64(0),C.c_int64(0);"
               "lib.ccrs_spec_k3_counters(C.byref(a),C.byref(h));lib.ccrs_loop_counters(C.byref(n));"
               "print(json.dumps([intr.tolist(),hist.tolist(),gp.get_poses().tolist(),int(summ.iterations),a.value,h.value,n.value]))")


def _solve_in_subprocess(loop, **env):
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-c", _SOLVE_CODE % loop], cwd=root, env=dict(os.environ, **env), capture_output=True,
                         text=True, check=True).stdout
    return json.loads(out.strip().splitlines()[-1])


def test_device_loop_is_audited_and_matches_the_host_controllers(pkg, oracle):
    """ccrs_solve_lm / ccrs_solve_gn run the device-driven loop (K3's last CTA executes the controller rule, no host
    round trip): every solve must have been audited by the host (same rule, bit-for-bit) and the trajectory must match
    the oracle and the host-driven controllers (CCRS_DEVICE_LOOP=0) at equal iteration count."""
    import ctypes as C
    s, op, gp, intr0 = _setup(pkg, oracle, "eucm", 100, seed=1, noise_px=0.1)
    lib = pkg._abi.load()
    n0, n1 = C.c_int64(0), C.c_int64(0)
    assert lib.ccrs_loop_counters(C.byref(n0)) == 1
    for loop in ("solve_lm", "solve_gn"):
        gp.set_poses(s.init_poses)
        lib.ccrs_loop_counters(C.byref(n0))
        intr, summ, hist = getattr(gp, loop)(intr0)
        lib.ccrs_loop_counters(C.byref(n1))
        # one audited device solve per iteration (Gauss-Newton's last iteration stops before its solve)
        assert summ.status == 0 and n1.value - n0.value == summ.iterations - (0 if loop == "solve_lm" else 1)
        ref = (op.levenberg_marquardt if loop == "solve_lm" else op.gauss_newton)(intr0, s.init_poses)
        assert summ.iterations == ref[2].iterations and np.max(np.abs(intr - ref[0]) / np.abs(ref[0])) < TOL_INTR
        assert np.max(np.abs(hist - np.asarray(ref[3])[: len(hist)]) / np.asarray(ref[3])[: len(hist)]) < 1e-9
        i2, h2, p2, it2, _, _, n2 = _solve_in_subprocess(loop, CCRS_DEVICE_LOOP="0")
        assert n2 == 0 and it2 == summ.iterations                               # the host-driven controller ran there
        assert np.max(np.abs(intr - np.array(i2)) / np.abs(intr)) < 1e-12
        assert np.max(np.abs(gp.get_poses() - np.array(p2))) < 1e-10
    gp.close()


def test_board_format_is_bit_identical(pkg):
    """ccrs_problem_create_board_f32 (corner ids + board table: the reference's FrameFeature / Board data model) expands
    p3d = board[id] on the device: results must equal the f32 entry point on the expanded arrays bit for bit."""
    s = pkg.synth.make_calib("eucm", 90, seed=12, drop_fraction=0.25)
    f = lambda a: a.astype(np.float32)
    ids, board = s.extra["corner_id"], s.extra["board"]
    assert np.array_equal(board[ids, 0].astype(np.float64), s.x) and np.array_equal(board[ids, 2].astype(np.float64), s.z)
    g32 = pkg.Problem(s.model, s.width, s.height, s.frame_offsets, f(s.x), f(s.y), f(s.z), f(s.u), f(s.v))
    gb = pkg.Problem(s.model, s.width, s.height, s.frame_offsets, None, None, None, f(s.u), f(s.v), corner_id=ids, board=board)
    for g in (g32, gb):
        g.set_poses(s.init_poses)
    assert np.array_equal(g32.linearize(s.init_params), gb.linearize(s.init_params))
    assert np.array_equal(g32.frame_blocks(), gb.frame_blocks())
    i1, s1, h1 = g32.solve_lm(s.init_params)
    i2, s2, h2 = gb.solve_lm(s.init_params)
    assert np.array_equal(i1, i2) and np.array_equal(h1, h2) and np.array_equal(g32.get_poses(), gb.get_poses())
    g32.close(); gb.close()
    bad = ids.copy(); bad[5] = len(board)
    with pytest.raises(pkg.CcrsError):
        pkg.Problem(s.model, s.width, s.height, s.frame_offsets, None, None, None, f(s.u), f(s.v), corner_id=bad, board=board)


def test_speculative_k3_is_used_and_changes_nothing():
    """Host-driven LM (CCRS_DEVICE_LOOP=0): the K3 launched behind the trial K2 (for 'accepted, u_next = u / 3') must be
    consumed on a converging run and the trajectory must be bit-identical to the run without it (same kernels, same
    arguments, only launched earlier)."""
    i1, h1, p1, it1, launched, hits, _ = _solve_in_subprocess("solve_lm", CCRS_DEVICE_LOOP="0")
    assert launched == it1 and hits >= 1
    i2, h2, p2, it2, launched2, _, _ = _solve_in_subprocess("solve_lm", CCRS_DEVICE_LOOP="0", CCRS_SPEC_K3="0")
    assert launched2 == 0 and it1 == it2
    assert np.array_equal(np.array(i1), np.array(i2)) and np.array_equal(np.array(h1), np.array(h2)) and np.array_equal(np.array(p1), np.array(p2))

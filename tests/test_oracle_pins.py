"""Pins for the CPU oracle (which in turn is the checker for the CUDA path).

(1) every fixture the reference's own tests hold for this path:
      tests/optimization_test.rs:36-80   test_reprojection_factor
      tests/util_test.rs:77-110          test_convert_model (parameter order, UCM == EUCM(beta=1))
      tests/types_test.rs:5-20           test_rvec_tvec_conversion (axis-angle convention)
      tests/board_test.rs:4-40           test_board_init (synthetic board geometry)
(2) an independent third party with the same conventions: OpenCV 4.x projectPoints / fisheye.projectPoints —
    values AND Jacobians for OPENCV5 and KB4 ("aka plumb_bob" / "aka OpenCV Fisheye", reference README.md:80-81);
(3) 50-digit mpmath differentiation of the whole residual for all six models;
(4) internal consistency: f64 path == dual path values, dense == per-frame-elimination solve, Huber corrector."""
import numpy as np
import pytest

from helpers import MODEL_NAMES, rel_err_rows


def _one_obs_problem(oracle, model_id, w, h, p3d, p2d, **kw):
    fo = np.array([0, 1], dtype=np.int32)
    return oracle.OracleProblem(model_id, w, h, fo, [p3d[0]], [p3d[1]], [p3d[2]], [p2d[0]], [p2d[1]], **kw)


# ---------------------------------------------------------------- (1) reference fixtures
def test_reference_test_reprojection_factor(oracle):
    prm = np.array([500.0, 500.0, 320.0, 240.0, 0.5])            # optimization_test.rs:41
    p3d = np.array([1.0, 2.0, 10.0])
    uv = oracle.project(0, prm, p3d)
    assert np.allclose(uv, [369.39015319, 338.78030638], atol=1e-8)   # SURVEY §8(c) golden vector (1)
    p2d = uv.astype(np.float32).astype(np.float64)                 # glam::Vec2 is f32 (optimization_test.rs:49)
    assert np.allclose(p2d, [369.39016724, 338.78030396], atol=1e-8)
    op = _one_obs_problem(oracle, 0, 640, 480, p3d, p2d, huber_delta=0.0)
    r = op.eval_r(prm, np.zeros((1, 6)), apply_loss=False)
    assert np.linalg.norm(r) < 1e-4                                 # optimization_test.rs:66-70
    assert abs(np.linalg.norm(r) - 1.425e-5) < 1e-7
    bad = np.zeros((1, 6)); bad[0, 3] = 0.1                         # tvec_bad (optimization_test.rs:73)
    r = op.eval_r(prm, bad, apply_loss=False)
    assert np.linalg.norm(r) > 1e-3
    assert np.allclose(r, [4.91153299, -0.04993990], atol=1e-7)


def test_reference_test_convert_model_param_order(oracle):
    """util_test.rs:77-110 / util.rs:230-243: UCM(fx,fy,cx,cy,alpha) == EUCM(.., alpha, beta=1) == EUCMT(.., 1, 0, 0)."""
    ucm = np.array([500.0, 510.0, 320.0, 240.0, 0.6])
    eucm = np.insert(ucm, 5, 1.0)
    eucmt = np.concatenate([eucm, [0.0, 0.0]])
    rng = np.random.default_rng(0)
    for _ in range(20):
        P = rng.normal(size=3) * [0.5, 0.5, 0.1] + [0, 0, 1.0]
        a, b, c = oracle.project(0, ucm, P), oracle.project(1, eucm, P), oracle.project(2, eucmt, P)
        assert np.array_equal(a, b) and np.array_equal(b, c)
    assert [oracle.lib().oracle_model_nparams(m) for m in range(6)] == [5, 6, 8, 8, 9, 8]


def test_reference_test_rvec_tvec_conversion(oracle, pkg):
    """types_test.rs:5-20: rvec (0.1,0.2,0.3), tvec (1,2,3) <-> Isometry3. The oracle's quaternion path must equal
    the Rodrigues rotation matrix of the same axis-angle vector."""
    rvec, tvec = np.array([0.1, 0.2, 0.3]), np.array([1.0, 2.0, 3.0])
    R = pkg.synth.rodrigues(rvec)
    for p in np.eye(3):
        assert np.allclose(oracle.transform_point(rvec, tvec, p), R @ p + tvec, atol=1e-15)
    # round trip: scaled_axis(R) == rvec
    import cv2
    back, _ = cv2.Rodrigues(R)
    assert np.allclose(back.ravel(), rvec, atol=1e-12)
    # exactly-zero rvec -> identity (nalgebra exp_eps branch)
    assert np.array_equal(oracle.transform_point(np.zeros(3), tvec, np.array([0.3, -0.2, 0.9])), np.array([0.3, -0.2, 0.9]) + tvec)


def test_reference_test_board_init(pkg):
    """board_test.rs:4-40: 6x6 tags, 0.088 m, spacing 0.3 -> 144 corners; corner order TL,TR,BR,BL; y down-negative."""
    b = pkg.synth.aprilgrid_board()
    assert b.shape == (144, 3) and b.dtype == np.float32
    assert np.array_equal(b[0], [0, 0, 0])
    assert np.allclose(b[1], [0.088, 0, 0]) and np.allclose(b[2], [0.088, -0.088, 0]) and np.allclose(b[3], [0, -0.088, 0])
    assert np.allclose(b[4], [0.088 * 1.3, 0, 0], atol=1e-7)
    assert np.all(b[:, 2] == 0)


# ---------------------------------------------------------------- (2) OpenCV
def _random_scene(rng, n=40, spread=0.25):
    pts = np.stack([rng.uniform(-spread, spread, n), rng.uniform(-spread, spread, n), rng.uniform(-0.05, 0.05, n)], axis=1)
    rvec = rng.normal(size=3) * 0.25
    tvec = np.array([0.03, -0.02, 0.8]) + rng.normal(size=3) * 0.02
    return pts, rvec, tvec


def test_opencv5_values_and_jacobian_vs_cv2(oracle):
    import cv2
    rng = np.random.default_rng(1)
    prm = np.array([600.0, 590.0, 512.0, 500.0, -0.1, 0.05, 1e-3, -1e-3, -0.01])   # fx fy cx cy k1 k2 p1 p2 k3
    K = np.array([[prm[0], 0, prm[2]], [0, prm[1], prm[3]], [0, 0, 1]])
    for _ in range(3):
        pts, rvec, tvec = _random_scene(rng)
        img, jac = cv2.projectPoints(pts.reshape(-1, 1, 3), rvec, tvec, K, prm[4:9])
        img = img.reshape(-1, 2)
        fo = np.array([0, len(pts)], dtype=np.int32)
        op = oracle.OracleProblem(4, 1024, 1024, fo, pts[:, 0], pts[:, 1], pts[:, 2], np.zeros(len(pts)), np.zeros(len(pts)), huber_delta=0.0)
        r, J = op.eval_rj(prm, np.concatenate([rvec, tvec])[None], apply_loss=False)
        assert np.max(np.abs(r.reshape(-1, 2) - img)) < 1e-9         # r = projection - 0
        # cv2 jac columns: rvec3 tvec3 fx fy cx cy k1 k2 p1 p2 k3 ; oracle: [fx fy cx cy k1 k2 p1 p2 k3 | rvec | tvec]
        J_cv = np.concatenate([jac[:, 6:10], jac[:, 10:15], jac[:, 0:3], jac[:, 3:6]], axis=1)
        assert np.max(rel_err_rows(J, J_cv)) < 1e-8


def test_kb4_values_and_jacobian_vs_cv2_fisheye(oracle):
    import cv2
    rng = np.random.default_rng(2)
    prm = np.array([380.0, 379.0, 510.0, 514.0, 0.01, -0.002, 3e-4, -4e-5])
    K = np.array([[prm[0], 0, prm[2]], [0, prm[1], prm[3]], [0, 0, 1]])
    for _ in range(3):
        pts, rvec, tvec = _random_scene(rng, spread=0.6)
        img, jac = cv2.fisheye.projectPoints(pts.reshape(-1, 1, 3), rvec.reshape(3, 1), tvec.reshape(3, 1), K, prm[4:8])
        img = img.reshape(-1, 2)
        fo = np.array([0, len(pts)], dtype=np.int32)
        op = oracle.OracleProblem(3, 1024, 1024, fo, pts[:, 0], pts[:, 1], pts[:, 2], np.zeros(len(pts)), np.zeros(len(pts)), huber_delta=0.0)
        r, J = op.eval_rj(prm, np.concatenate([rvec, tvec])[None], apply_loss=False)
        assert np.max(np.abs(r.reshape(-1, 2) - img)) < 1e-9
        # fisheye jac columns: f(2) c(2) k(4) om(3) T(3) alpha(1)
        J_cv = jac[:, :14]
        assert np.max(rel_err_rows(J, J_cv)) < 1e-8


# ---------------------------------------------------------------- (3) mpmath
def _mp_residual(mp, model, q):
    """whole residual in mpmath: q = [params..., rvec(3), tvec(3), p(3)]; follows factors.rs:152-173 literally."""
    n = {"ucm": 5, "eucm": 6, "eucmt": 8, "kb4": 8, "opencv5": 9, "ftheta": 8}[model]
    prm, rv, tv, p = q[:n], q[n:n + 3], q[n + 3:n + 6], q[n + 6:n + 9]
    h = [c / 2 for c in rv]
    nn = mp.sqrt(sum(c * c for c in h))
    s = mp.sin(nn) / nn
    w, v = mp.cos(nn), [c * s for c in h]
    cross = lambda a, b: [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
    t = [2 * c for c in cross(v, p)]
    c2 = cross(v, t)
    P = [t[i] * w + c2[i] + p[i] + tv[i] for i in range(3)]
    x, y, z = P
    fx, fy, cx, cy = prm[:4]
    if model in ("ucm", "eucm", "eucmt"):
        alpha = prm[4]; beta = 1 if model == "ucm" else prm[5]
        r2 = x * x + y * y
        rho = mp.sqrt(beta * r2 + z * z)
        nrm = alpha * rho + (1 - alpha) * z
        mx, my = x / nrm, y / nrm
        if model == "eucmt":
            t1, t2 = prm[6], prm[7]
            rr = mx * mx + my * my
            mx, my = mx + 2 * t1 * mx * my + t2 * (rr + 2 * mx * mx), my + t1 * (rr + 2 * my * my) + 2 * t2 * mx * my
    elif model in ("kb4", "ftheta"):
        r = mp.sqrt(x * x + y * y)
        th = mp.atan2(r, z)
        k = prm[4:8]
        if model == "kb4":
            d = th * (1 + k[0] * th**2 + k[1] * th**4 + k[2] * th**6 + k[3] * th**8)
        else:
            d = th * (1 + k[0] * th + k[1] * th**2 + k[2] * th**3 + k[3] * th**4)
        mx, my = d * x / r, d * y / r
    else:
        k1, k2, p1, p2, k3 = prm[4:9]
        a, b = x / z, y / z
        r2 = a * a + b * b
        rad = 1 + k1 * r2 + k2 * r2**2 + k3 * r2**3
        mx = a * rad + 2 * p1 * a * b + p2 * (r2 + 2 * a * a)
        my = b * rad + p1 * (r2 + 2 * b * b) + 2 * p2 * a * b
    return [fx * mx + cx, fy * my + cy]


@pytest.mark.parametrize("model", MODEL_NAMES)
def test_jacobian_vs_mpmath_50_digits(oracle, pkg, model):
    import mpmath as mp
    mp.mp.dps = 50
    rng = np.random.default_rng(7)
    prm = pkg.synth.GT_PARAMS[model] * (1 + 0.01 * rng.normal(size=len(pkg.synth.GT_PARAMS[model])))
    n = len(prm)
    for trial in range(3):
        rvec = rng.normal(size=3) * (0.3 if trial < 2 else 1e-5)        # incl. a tiny rotation
        tvec = np.array([0.05, -0.03, 0.6]) + rng.normal(size=3) * 0.05
        p = np.array([rng.uniform(-0.3, 0.3), rng.uniform(-0.3, 0.3), rng.uniform(-0.02, 0.02)])
        q0 = [mp.mpf(float(v)) for v in np.concatenate([prm, rvec, tvec, p])]
        J_mp = np.zeros((2, n + 6))
        for c in range(n + 6):
            for row in range(2):
                f = lambda t, c=c, row=row: _mp_residual(mp, model, q0[:c] + [t] + q0[c + 1:])[row]
                J_mp[row, c] = float(mp.diff(f, q0[c]))
        uv_mp = np.array([float(v) for v in _mp_residual(mp, model, q0)])
        op = _one_obs_problem(oracle, pkg.MODELS[model], 1024, 1024, p, [0.0, 0.0], huber_delta=0.0)
        r, J = op.eval_rj(prm, np.concatenate([rvec, tvec])[None], apply_loss=False)
        assert np.max(np.abs(r - uv_mp) / np.maximum(np.abs(uv_mp), 1.0)) < 1e-12
        assert np.max(rel_err_rows(J, J_mp)) < 1e-9


# ---------------------------------------------------------------- (4) internal consistency
@pytest.mark.parametrize("one_focal", [False, True])
@pytest.mark.parametrize("model", MODEL_NAMES)
def test_f64_path_equals_dual_path_and_numpy_model(oracle, pkg, model, one_focal):
    s = pkg.synth.make_calib(model, 6, seed=3)
    op = oracle.OracleProblem.from_synth(s, pkg.MODELS[model], xy_same_focal=one_focal)
    intr = pkg.synth.intr_from_full(s.init_params, one_focal)
    r1 = op.eval_r(intr, s.init_poses, apply_loss=True)
    r2, J = op.eval_rj(intr, s.init_poses, apply_loss=True)
    assert np.allclose(r1, r2, rtol=1e-12, atol=1e-12)   # dual division multiplies by the reciprocal (num-dual), f64 divides
    assert J.shape == (2 * s.n_obs, op.d + 6)
    # third statement of the models (numpy, synth.project)
    full = pkg.synth.full_from_intr(intr, one_focal)
    R = pkg.synth.rodrigues(s.init_poses[:, :3])
    f_of = np.repeat(np.arange(s.n_frames), np.diff(s.frame_offsets))
    P = np.einsum("nij,nj->ni", R[f_of], np.stack([s.x, s.y, s.z], axis=1)) + s.init_poses[f_of, 3:]
    r_np = (pkg.synth.project(model, full, P) - np.stack([s.u, s.v], axis=1)).reshape(-1)
    r_raw = op.eval_r(intr, s.init_poses, apply_loss=False)
    assert np.max(np.abs(r_np - r_raw)) < 1e-9
    # finite differences (coarse, catches sign/ordering mistakes)
    eps = 1e-6
    for c in range(op.d):
        a = intr.copy(); a[c] += eps
        fd = (op.eval_r(a, s.init_poses, apply_loss=False) - r_raw) / eps
        _, J_raw = op.eval_rj(intr, s.init_poses, apply_loss=False)
        assert np.allclose(fd, J_raw[:, c], rtol=1e-3, atol=1e-3 * np.abs(J_raw[:, c]).max())


def test_huber_corrector(oracle, pkg):
    """HuberLoss::new(1.0) + Corrector: inliers untouched, outliers scaled by sqrt(delta/||r||) (util.rs:413)."""
    s = pkg.synth.make_calib("eucm", 4, seed=1)
    op = oracle.OracleProblem.from_synth(s, 1)
    r_raw, J_raw = op.eval_rj(s.init_params, s.init_poses, apply_loss=False)
    r_cor, J_cor = op.eval_rj(s.init_params, s.init_poses, apply_loss=True)
    nrm = np.linalg.norm(r_raw.reshape(-1, 2), axis=1)
    w = np.where(nrm > 1.0, np.sqrt(1.0 / np.maximum(nrm, 1e-300)), 1.0)
    assert (nrm > 1.0).any()
    assert np.allclose(r_cor.reshape(-1, 2), r_raw.reshape(-1, 2) * w[:, None], rtol=1e-15)
    assert np.allclose(J_cor, J_raw * np.repeat(w, 2)[:, None], rtol=1e-15)
    op0 = oracle.OracleProblem.from_synth(s, 1, huber_delta=0.0)
    assert np.array_equal(op0.eval_rj(s.init_params, s.init_poses, apply_loss=True)[0], r_raw)


def test_dense_cholesky_equals_frame_elimination(oracle, pkg):
    """the reference factorises the whole sparse system; the GPU eliminates per frame — same step."""
    s = pkg.synth.make_calib("eucm", 15, seed=2)
    op = oracle.OracleProblem.from_synth(s, 1)
    for u in (0.0, 1e-3):
        _, di0, dp0, md0 = op.solve_step(s.init_params, s.init_poses, u=u, options=op.default_options(solver=0))
        _, di1, dp1, md1 = op.solve_step(s.init_params, s.init_poses, u=u, options=op.default_options(solver=1))
        assert np.allclose(di0, di1, rtol=1e-9, atol=1e-14) and np.allclose(dp0, dp1, rtol=1e-8, atol=1e-13)
        assert abs(md0 - md1) / abs(md1) < 1e-10


def test_config1_gn_converges_to_ground_truth(oracle, pkg):
    """BASELINE config 1: EUCM, 100 frames x 144 corners, 1024x1024 (tests/optimization_test.rs-style)."""
    s = pkg.synth.make_calib("eucm", 100, seed=0)
    assert s.n_obs == 14400 and s.n_frames == 100
    op = oracle.OracleProblem.from_synth(s, 1)
    intr, poses, res, hist = op.gauss_newton(s.init_params, s.init_poses)
    assert res.status == 0 and res.iterations <= 8
    assert np.max(np.abs(intr - s.gt_params) / s.gt_params) < 1e-6
    r = op.eval_r(intr, poses, apply_loss=False)
    assert np.sqrt(np.mean(r**2)) < 1e-4          # f32 rounding of the observations is the floor (~2e-5 px)

"""Committed golden vectors for the §8(f) entry points (tests/golden/golden_v2.npz, made by make_golden_v2.py).
CPU: the oracle still reproduces them. GPU: the CUDA path matches them through the C ABI."""
import os

import numpy as np
import pytest

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "golden_v2.npz"))
GT_UCM = np.array([400.0, 400.0, 512.0, 512.0, 0.6])


def test_oracle_reproduces_golden_v2(pkg, oracle):
    s = pkg.synth.make_calib("kb4", 25, seed=40, noise_px=0.1, drop_fraction=0.2)
    med, avg, e = oracle.OracleProblem.from_synth(s, 3).validation(s.init_params, s.init_poses)
    assert np.allclose([med, avg], G["val_kb4"], rtol=1e-12) and np.allclose(e[:32], G["val_kb4_err_head"], rtol=1e-11)
    s = pkg.synth.make_calib("ucm", 2, seed=41, gt_params=GT_UCM, noise_px=0.05)
    fa, poses, res, _ = oracle.init_ucm_gn(oracle.OracleProblem.from_synth(s, 0), 512.0, 512.0, 330.0, 0.5, s.init_poses)
    assert res.iterations == int(G["ucm_init_ff0_iters"]) and np.allclose(fa, G["ucm_init_ff0_fa"], rtol=1e-10)
    sp = np.array(pkg.synth.GT_PARAMS["kb4"], dtype=np.float64)
    ref, res, _ = oracle.convert_model_gn(3, sp, 1, G["conv_kb4_eucm_init"], G["conv_kb4_eucm_rays"], None, None, np.zeros(6, dtype=np.uint8))
    assert res.iterations == int(G["conv_kb4_eucm_iters"]) and np.allclose(ref, G["conv_kb4_eucm_params"], rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("model", ["eucm", "kb4", "opencv5"])
def test_gpu_validation_matches_golden(pkg, model):
    s = pkg.synth.make_calib(model, 25, seed=40, noise_px=0.1, drop_fraction=0.2)
    with pkg.Problem.from_synth(s) as gp:
        med, avg, e = gp.validation(s.init_params, s.init_poses, want_errors=True)
    assert np.allclose([med, avg], G[f"val_{model}"], rtol=1e-9)
    assert np.allclose(e[:32], G[f"val_{model}_err_head"], rtol=1e-9, atol=1e-12)


@pytest.mark.gpu
@pytest.mark.parametrize("ff", [0, 1])
def test_gpu_init_ucm_stage1_matches_golden(pkg, ff):
    s = pkg.synth.make_calib("ucm", 2, seed=41, gt_params=GT_UCM, noise_px=0.05)
    inf = np.inf
    with pkg.Problem.from_synth(s, xy_same_focal=True) as gp:
        gp.set_poses(s.init_poses)
        intr, summ, hist = gp.solve_gn([330.0, 512.0, 512.0, 0.5], lo=[110.0, -inf, -inf, 1e-6], hi=[990.0, inf, inf, 1.0],
                                       fixed=[1 if ff else 0, 2, 2, 0])
        fa = G[f"ucm_init_ff{ff}_fa"]
        assert summ.iterations == int(G[f"ucm_init_ff{ff}_iters"])
        assert abs(intr[0] - fa[0]) <= 1e-6 * fa[0] and abs(intr[3] - fa[1]) <= 1e-6 * fa[1]
        assert np.max(np.abs(gp.get_poses() - G[f"ucm_init_ff{ff}_poses"])) < 1e-6
        assert np.allclose(hist, G[f"ucm_init_ff{ff}_hist"], rtol=1e-7)


@pytest.mark.gpu
@pytest.mark.parametrize("src,tgt,dis", [("kb4", "eucm", 0), ("eucm", "kb4", 1)])
def test_gpu_convert_model_matches_golden(pkg, src, tgt, dis):
    """ccrs_convert_model on the golden's own points; the product's model bounds are inactive for these fits."""
    import ctypes as C
    sp = np.array(pkg.synth.GT_PARAMS[src], dtype=np.float64)
    rays = np.ascontiguousarray(G[f"conv_{src}_{tgt}_rays"].T)
    tgt_p = G[f"conv_{src}_{tgt}_init"].copy()
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    summ = pkg.Summary()
    code = pkg._abi.load().ccrs_convert_model(pkg.MODELS[src], dp(sp), pkg.MODELS[tgt], dp(tgt_p), 1024, 1024, dis, rays.shape[1],
                                             dp(rays[0]), dp(rays[1]), dp(rays[2]), None, C.byref(summ), 0)
    assert code == 0 and summ.iterations == int(G[f"conv_{src}_{tgt}_iters"])
    ref = G[f"conv_{src}_{tgt}_params"]
    assert np.max(np.abs(tgt_p - ref) / np.maximum(np.abs(ref), 1e-3)) < 1e-6


@pytest.mark.gpu
def test_gpu_init_poses_match_golden(pkg):
    s = pkg.synth.make_calib("eucm", 6, seed=42, drop_fraction=0.4)
    poses, cost = pkg.init_poses(s.frame_offsets, s.x, s.y, s.z, G["pnp_xn"], G["pnp_yn"], want_cost=True)
    assert np.max(np.abs(cost - G["pnp_cost"]) / G["pnp_cost"]) < 1e-8
    assert np.max(np.abs(poses - G["pnp_poses"])) < 1e-6

"""CPU oracle for the initial-pose step (TEST INFRASTRUCTURE ONLY, like everything under oracle/).

The reference calls `sqpnp_simple::sqpnp_solve_glam(&p3ds, &p2ds_z)` per frame (src/util.rs:435-436,
src/optimization/linear.rs:20); the crate (sqpnp_simple 0.2.0, Cargo.toml:42) is not vendored, so this restates the
published SQPnP objective (Terzakis & Lourakis, "A Consistently Fast and Globally Optimal Solution to the
Perspective-n-Point Problem", ECCV 2020, eqs. 4-10) literally — explicit A_i = I_3 (x) p_i^T and Q_i matrices, Omega and
P by dense linear algebra — and minimises it over SO(3) with scipy from many starts. PARITY UNPINNED against the crate
itself (no Rust toolchain here); what is pinned is the objective's defining property: the recovered pose reproduces the
generating pose on exact data, which tests/test_init_poses.py checks for both the oracle and the CUDA path.
"""
from __future__ import annotations

import numpy as np
from scipy.optimize import minimize
from scipy.spatial.transform import Rotation


def omega_and_p(p3d: np.ndarray, xn: np.ndarray, yn: np.ndarray):
    """Omega (9x9) and P (3x9): cost = r^T Omega r, t = P r, r = vec(R) row-major."""
    n = len(p3d)
    SQ = np.zeros((3, 3)); SQA = np.zeros((3, 9)); AQA = np.zeros((9, 9))
    for i in range(n):
        A = np.kron(np.eye(3), p3d[i][None, :])                       # 3 x 9: A r = R p
        Q = np.array([[1.0, 0.0, -xn[i]], [0.0, 1.0, -yn[i]], [-xn[i], -yn[i], xn[i] ** 2 + yn[i] ** 2]])
        SQ += Q; SQA += Q @ A; AQA += A.T @ Q @ A
    P = -np.linalg.solve(SQ, SQA)
    Om = AQA + SQA.T @ P + P.T @ SQA + P.T @ SQ @ P                   # sum (A + P)^T Q (A + P), expanded
    return Om, P


def solve_frame(p3d, xn, yn, n_starts: int = 64, seed: int = 0):
    """(rvec, tvec, cost): global minimiser of the SQPnP cost with positive mean depth."""
    p3d = np.asarray(p3d, dtype=np.float64); xn = np.asarray(xn, dtype=np.float64); yn = np.asarray(yn, dtype=np.float64)
    Om, P = omega_and_p(p3d, xn, yn)
    cost = lambda rv: (lambda r: r @ Om @ r)(Rotation.from_rotvec(rv).as_matrix().reshape(-1))
    rng = np.random.default_rng(seed)
    best = None
    for rv0 in Rotation.random(n_starts, random_state=rng).as_rotvec():
        res = minimize(cost, rv0, method="BFGS", options={"gtol": 1e-14})
        R = Rotation.from_rotvec(res.x).as_matrix()
        t = P @ R.reshape(-1)
        depth = (R @ p3d.mean(axis=0) + t)[2]
        if depth > 0 and (best is None or res.fun < best[2] - 1e-15):
            best = (Rotation.from_matrix(R).as_rotvec(), t, float(res.fun))
    return best

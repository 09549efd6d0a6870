"""Host-side unprojection of the six camera models (pixel -> ray).

In the reference this is `GenericModel::unproject` of the camera-intrinsic-model crate (call site
src/optimization/factors.rs:38): set-up code of `convert_model`, executed once per conversion on a ~30 x 30 pixel grid —
not part of the per-iteration path, so it stays on the host exactly as it stays in the model crate for a Rust caller of
the C ABI (ccrs_convert_model takes the unprojected points). Rays are returned with unit norm; projection is
scale-invariant for every model, so the normalisation does not matter to the caller.
"""
from __future__ import annotations

import numpy as np

from .synth import project

_NAMES = ["ucm", "eucm", "eucmt", "kb4", "opencv5", "ftheta"]


def _name(model) -> str:
    return model if isinstance(model, str) else _NAMES[int(model)]


def _eucm_ray(mx, my, alpha, beta):
    r2 = mx * mx + my * my
    disc = 1.0 - (2.0 * alpha - 1.0) * beta * r2
    valid = disc >= 0.0
    mz = (1.0 - beta * alpha * alpha * r2) / (alpha * np.sqrt(np.where(valid, disc, 1.0)) + (1.0 - alpha))
    return np.stack([mx, my, mz], axis=-1), valid


def _poly_theta(rd, k, odd: bool, iters=20):
    """solve d(theta) = rd for theta, d = theta (1 + k1 t^2 + ...) (KB4, odd) or theta (1 + k1 t + ...) (FTHETA)."""
    th = rd.copy()
    for _ in range(iters):
        if odd:
            t2 = th * th
            d = th * (1 + t2 * (k[0] + t2 * (k[1] + t2 * (k[2] + t2 * k[3]))))
            dd = 1 + t2 * (3 * k[0] + t2 * (5 * k[1] + t2 * (7 * k[2] + t2 * 9 * k[3])))
        else:
            d = th * (1 + th * (k[0] + th * (k[1] + th * (k[2] + th * k[3]))))
            dd = 1 + th * (2 * k[0] + th * (3 * k[1] + th * (4 * k[2] + th * 5 * k[3])))
        th = th - (d - rd) / dd
    return th


def unproject(model, params, p2ds):
    """(rays [n,3] unit norm, valid [n] bool). Invalid pixels (outside the model's domain) are flagged like the
    reference's `Option::None` entries (factors.rs:39-42)."""
    m = _name(model)
    prm = np.asarray(params, dtype=np.float64)
    uv = np.asarray(p2ds, dtype=np.float64).reshape(-1, 2)
    fx, fy, cx, cy = prm[:4]
    mx, my = (uv[:, 0] - cx) / fx, (uv[:, 1] - cy) / fy
    valid = np.ones(len(uv), dtype=bool)
    if m in ("ucm", "eucm"):
        ray, valid = _eucm_ray(mx, my, prm[4], 1.0 if m == "ucm" else prm[5])
    elif m == "eucmt":
        t1, t2 = prm[6], prm[7]
        x, y = mx.copy(), my.copy()
        for _ in range(30):          # invert the tangential map by fixed-point iteration (|t| << 1)
            rr = x * x + y * y
            x = mx - (2 * t1 * x * y + t2 * (rr + 2 * x * x))
            y = my - (t1 * (rr + 2 * y * y) + 2 * t2 * x * y)
        ray, valid = _eucm_ray(x, y, prm[4], prm[5])
    elif m in ("kb4", "ftheta"):
        rd = np.hypot(mx, my)
        th = _poly_theta(rd, prm[4:8], odd=(m == "kb4"))
        valid = (th >= 0) & (th < np.pi)
        s = np.where(rd > 1e-12, np.sin(th) / np.maximum(rd, 1e-300), 1.0)
        ray = np.stack([mx * s, my * s, np.cos(th)], axis=-1)
    elif m == "opencv5":
        k1, k2, p1, p2, k3 = prm[4:9]
        a, b = mx.copy(), my.copy()
        for _ in range(50):
            r2 = a * a + b * b
            rad = 1 + r2 * (k1 + r2 * (k2 + r2 * k3))
            a = (mx - (2 * p1 * a * b + p2 * (r2 + 2 * a * a))) / rad
            b = (my - (p1 * (r2 + 2 * b * b) + 2 * p2 * a * b)) / rad
        ray = np.stack([a, b, np.ones_like(a)], axis=-1)
    else:
        raise ValueError(model)
    ray = ray / np.linalg.norm(ray, axis=1, keepdims=True)
    # keep only pixels the forward model reproduces (guards the iterative inverses and the domain tests)
    back = project(m, prm, ray)
    valid = valid & np.all(np.isfinite(back), axis=1) & (np.linalg.norm(back - uv, axis=1) < 1e-6)
    return ray, valid


def conversion_grid(width: float, height: float) -> np.ndarray:
    """The pixel grid of ModelConvertFactor::new (factors.rs:32-37) with convert_model's edge / step
    (util.rs:245-246): edge = max(w, h) as u32 / 100, step = (max(w, h) / 30.0) as usize."""
    edge = int(max(width, height)) // 100
    step = int(max(width, height) / 30.0)
    rows = range(edge, int(height) - edge, step)
    cols = range(edge, int(width) - edge, step)
    return np.array([(float(c), float(r)) for r in rows for c in cols], dtype=np.float64)

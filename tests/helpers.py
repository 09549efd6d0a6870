"""Shared helpers for the parity tests."""
import numpy as np

MODEL_NAMES = ["ucm", "eucm", "eucmt", "kb4", "opencv5", "ftheta"]


def rel_err_rows(J, Jref):
    """north_star tolerance metric: |J - Jref| / max(|Jref|, 1e-3 * ||row||)."""
    rn = np.linalg.norm(Jref, axis=1, keepdims=True)
    den = np.maximum(np.abs(Jref), 1e-3 * rn)
    den = np.maximum(den, 1e-300)
    return np.abs(J - Jref) / den


def rel_err_vec(r, rref, floor):
    return np.abs(r - rref) / np.maximum(np.abs(rref), floor)


def tri_idx(NA, i, j):
    return i * NA - (i * (i - 1)) // 2 + (j - i)


def unpack_block(b, NA):
    H = np.zeros((NA, NA))
    k = 0
    for i in range(NA):
        for j in range(i, NA):
            H[i, j] = b[k]; H[j, i] = b[k]; k += 1
    return H


def block_rel_err(B, Bref, NA):
    """relative to sqrt(H_ii H_jj) of the reference block (scale-aware)."""
    out = 0.0
    for f in range(B.shape[0]):
        H, Hr = unpack_block(B[f], NA), unpack_block(Bref[f], NA)
        d = np.sqrt(np.maximum(np.diag(Hr), 1e-300))
        out = max(out, np.max(np.abs(H - Hr) / np.outer(d, d)))
    return out


def rms_px(oracle_problem, intr, poses):
    r = oracle_problem.eval_r(intr, poses, apply_loss=False)
    return float(np.sqrt(np.mean(r.reshape(-1, 2) ** 2)))

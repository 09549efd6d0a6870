"""A ccrs_backend (include/ccrs_b200.h) whose per-frame work is done by the CPU ORACLE — test infrastructure.

It lets the CPU test-suite drive the PRODUCT's loop controllers (ccrs_controller_gn / ccrs_controller_lm inside
libccrs_b200.so) without a GPU: single rank, and world_size-2 over gloo with the frames sharded across ranks,
exactly the way the CUDA backend shards them over NCCL."""
import ctypes as C

import numpy as np

from helpers import tri_idx


class OracleBackend:
    def __init__(self, pkg, oracle_problem, poses, allreduce=None):
        self.abi = pkg._abi
        self.op = oracle_problem
        self.d = oracle_problem.d
        self.F = oracle_problem.n_frames
        self.NA = self.d + 7
        self.poses = [np.array(poses, dtype=np.float64).reshape(-1, 6).copy(), None]
        self.poses[1] = self.poses[0].copy()
        self.blocks = [None, None]
        self.cur = 0
        self.pose_scale = None
        self.intr_scale = None
        self.elim = None
        self.md = None
        self._allreduce = allreduce
        self.calls = []
        a = self.abi
        self._cb = dict(
            linearize=a.BE_LINEARIZE(self._linearize), compute_scale=a.BE_COMPUTE_SCALE(self._compute_scale),
            set_intr_scale=a.BE_SET_INTR_SCALE(self._set_intr_scale), reduce=a.BE_REDUCE(self._reduce),
            backsub=a.BE_BACKSUB(self._backsub), trial_stats=a.BE_TRIAL_STATS(self._trial_stats),
            accept=a.BE_ACCEPT(self._accept),
            allreduce=a.BE_ALLREDUCE(self._allreduce_cb) if allreduce is not None else a.BE_ALLREDUCE())
        self.c = a.Backend(None, self.d, 1, **self._cb)

    # ---- helpers ----
    def _H(self, which):
        B = self.blocks[self.cur ^ which]
        NA = self.NA
        H = np.zeros((self.F, NA, NA))
        k = 0
        for i in range(NA):
            for j in range(i, NA):
                H[:, i, j] = B[:, k]; H[:, j, i] = B[:, k]; k += 1
        return H

    def _arr(self, ptr, n):
        return np.ctypeslib.as_array(ptr, shape=(n,))

    # ---- callbacks ----
    def _linearize(self, ctx, intr, which):
        a = self._arr(intr, self.d).copy()
        _, blk = self.op.linearize(a, self.poses[self.cur ^ which])
        self.blocks[self.cur ^ which] = blk
        self.calls.append("linearize")
        return 0

    def _compute_scale(self, ctx, which, col_sq):
        H = self._H(which)
        d = self.d
        self.pose_scale = 1.0 / (1.0 + np.sqrt(np.stack([H[:, d + i, d + i] for i in range(6)], axis=1)))
        self._arr(col_sq, d)[:] = np.stack([H[:, i, i] for i in range(d)], axis=1).sum(axis=0) if self.F else 0.0
        return 0

    def _set_intr_scale(self, ctx, s):
        self.intr_scale = self._arr(s, self.d).copy() if s else None
        return 0

    def _reduce(self, ctx, which, u, use_scale, min_diag, max_diag, out):
        d, F = self.d, self.F
        H = self._H(which)
        uu = self._arr(u, 1)[0] if u else 0.0
        sa = self.intr_scale if use_scale else np.ones(d)
        sp = self.pose_scale if use_scale else np.ones((F, 6))
        self.use_scale = bool(use_scale)
        A = sa[None, :, None] * H[:, :d, :d] * sa[None, None, :]
        Bm = sa[None, :, None] * H[:, :d, d:d + 6] * sp[:, None, :]
        Cm = sp[:, :, None] * H[:, d:d + 6, d:d + 6] * sp[:, None, :]
        ga = -sa[None, :] * H[:, :d, d + 6]
        gp = -sp * H[:, d:d + 6, d + 6]
        dd = np.clip(np.einsum("fii->fi", Cm), min_diag, max_diag)
        Cr = Cm + uu * np.einsum("fi,ij->fij", dd, np.eye(6))
        X = np.linalg.solve(Cr, np.transpose(Bm, (0, 2, 1))) if F else np.zeros((0, 6, d))   # (F,6,d)
        cg = np.linalg.solve(Cr, gp[:, :, None])[:, :, 0] if F else np.zeros((0, 6))
        S = (A - Bm @ X).sum(axis=0)
        gs = (ga - np.einsum("fai,fi->fa", Bm, cg)).sum(axis=0)
        self.elim = (X, cg, gp, dd)
        o = self._arr(out, d * d + 3 * d + 1)
        o[:d * d] = S.reshape(-1)
        o[d * d:d * d + d] = gs
        o[d * d + d:d * d + 2 * d] = ga.sum(axis=0)
        o[d * d + 2 * d:d * d + 3 * d] = np.einsum("fii->i", A)
        o[-1] = H[:, d + 6, d + 6].sum()
        self.calls.append("reduce")
        return 0

    def _backsub(self, ctx, y_a, u, active, in_place):
        d = self.d
        y = self._arr(y_a, d).copy()
        uu = self._arr(u, 1)[0] if u else 0.0
        X, cg, gp, dd = self.elim
        yp = cg - X @ y
        sp = self.pose_scale if self.use_scale else np.ones((self.F, 6))
        new = self.poses[self.cur] + sp * yp
        self.poses[self.cur if in_place else self.cur ^ 1] = new
        self.md = float(np.sum(yp * gp + uu * dd * yp * yp))
        self.calls.append("backsub")
        return 0

    def _trial_stats(self, ctx, intr_trial, speculative, out):
        a = self._arr(intr_trial, self.d).copy()
        if speculative:
            sq, blk = self.op.linearize(a, self.poses[self.cur ^ 1])
            self.blocks[self.cur ^ 1] = blk
        else:
            sq = self.op.sq_error(a, self.poses[self.cur ^ 1])
        o = self._arr(out, 2)
        o[0] = self.md; o[1] = sq
        self.calls.append("trial_stats")
        return 0

    def _accept(self, ctx, mask):
        if (not mask) or mask[0]:
            self.cur ^= 1
        return 0

    def _allreduce_cb(self, ctx, buf, count):
        a = self._arr(buf, count)
        a[:] = self._allreduce(a.copy())
        return 0

    # ---- driving the product controllers ----
    def run(self, which, intr, lo=None, hi=None, fixed=None, options=None):
        lib = self.abi.load()
        a = np.array(intr, dtype=np.float64).copy()
        o = options or self.abi.default_options()
        s = self.abi.Summary()
        hist = np.full(o.max_iteration, np.nan)
        dp = lambda x: x.ctypes.data_as(C.POINTER(C.c_double)) if x is not None else None
        lo_a = np.ascontiguousarray(lo, dtype=np.float64) if lo is not None else None
        hi_a = np.ascontiguousarray(hi, dtype=np.float64) if hi is not None else None
        fx = np.ascontiguousarray(fixed, dtype=np.uint8) if fixed is not None else None
        fn = lib.ccrs_controller_gn if which == "gn" else lib.ccrs_controller_lm
        code = fn(C.byref(self.c), dp(a), dp(lo_a), dp(hi_a),
                  fx.ctypes.data_as(C.POINTER(C.c_ubyte)) if fx is not None else None, C.byref(o), C.byref(s), dp(hist))
        return code, a, self.poses[self.cur].copy(), s, hist[: s.iterations]

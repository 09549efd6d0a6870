"""N>1 host logic on CPU: world_size-2 gloo. Frames are sharded contiguously across ranks; every
rank runs the PRODUCT controller on its shard with the oracle doing the per-frame work, and the reduced intrinsic
system is summed across ranks once per linearisation. All ranks must end with identical intrinsics, equal to the
single-rank result."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))


def _worker(rank, world, port, model, n_frames, which, out_dir):
    sys.path.insert(0, HERE); sys.path.insert(0, os.path.dirname(HERE)); sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle"))
    import importlib
    import oracle as O
    from oracle_backend import OracleBackend
    pkg = importlib.import_module("camera-intrinsic-calibration-rs_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    s = pkg.synth.make_calib(model, n_frames, seed=0, drop_fraction=0.1)
    lo_f, hi_f = pkg.dist.shard_frames(s.frame_offsets, rank, world)
    sh = pkg.dist.slice_problem(s, lo_f, hi_f)
    op = O.OracleProblem(pkg.MODELS[model], s.width, s.height, sh["frame_offsets"], sh["x"], sh["y"], sh["z"], sh["u"], sh["v"], n_threads=2)

    def allreduce(a):
        # deterministic: gather every rank's partial and sum in rank order on every rank
        t = torch.from_numpy(a)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        acc = parts[0].clone()
        for p in parts[1:]:
            acc += p
        return acc.numpy()

    be = OracleBackend(pkg, op, s.init_poses[lo_f:hi_f], allreduce=allreduce)
    code, intr, poses, summ, hist = be.run(which, s.init_params)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), intr=intr, poses=poses, iters=summ.iterations, code=code, lo=lo_f, hi=hi_f)
    dist.destroy_process_group()


@pytest.mark.parametrize("which", ["gn", "lm"])
def test_two_rank_sharded_solve_equals_single_rank(pkg, oracle, tmp_path, which):
    model, n_frames, world = "eucm", 21, 2
    port = 29500 + (os.getpid() % 2000) + (0 if which == "gn" else 1)
    mp.spawn(_worker, args=(world, port, model, n_frames, which, str(tmp_path)), nprocs=world, join=True)
    s = pkg.synth.make_calib(model, n_frames, seed=0, drop_fraction=0.1)
    op = oracle.OracleProblem.from_synth(s, 1)
    ref = (op.gauss_newton if which == "gn" else op.levenberg_marquardt)(s.init_params, s.init_poses)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    assert all(int(x["code"]) == 0 for x in r)
    assert np.array_equal(r[0]["intr"], r[1]["intr"])                  # bitwise identical on every rank
    assert int(r[0]["iters"]) == int(r[1]["iters"]) == ref[2].iterations
    assert np.max(np.abs(r[0]["intr"] - ref[0]) / np.abs(ref[0])) < 1e-8
    poses = np.concatenate([x["poses"] for x in r])
    assert (int(r[0]["lo"]), int(r[0]["hi"]), int(r[1]["hi"])) == (0, int(r[1]["lo"]), s.n_frames)
    assert np.max(np.abs(poses - ref[1])) < 1e-8


def test_shard_frames_balances_observations(pkg):
    fo = np.concatenate([[0], np.cumsum(np.random.default_rng(0).integers(24, 145, size=1000))]).astype(np.int32)
    for world in (1, 2, 3, 4, 8):
        edges = [pkg.dist.shard_frames(fo, r, world) for r in range(world)]
        assert edges[0][0] == 0 and edges[-1][1] == 1000
        assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
        loads = [fo[b] - fo[a] for a, b in edges]
        assert max(loads) - min(loads) <= 2 * 144

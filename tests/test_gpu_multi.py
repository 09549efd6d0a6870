"""Multi-GPU parity (needs >= 2 B200s; skipped otherwise): frames sharded across ranks, the reduced intrinsic system
exchanged once per linearisation inside K2/K3 over peer memory (or with NCCL: all-gather + rank-order sum, or all-reduce). Every rank must end with
bitwise-identical intrinsics that match the single-GPU / oracle result."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _n_gpus():
    try:
        import torch
        return torch.cuda.device_count()
    except Exception:
        return 0


def _worker(rank, world, port, model, n_frames, loop, deterministic, out_dir, p2p=True):
    import importlib
    os.environ["CCRS_P2P"] = "1" if p2p else "0"
    import torch
    import torch.distributed as dist
    sys.path.insert(0, os.path.dirname(HERE))
    pkg = importlib.import_module("camera-intrinsic-calibration-rs_b200")
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    s = pkg.synth.make_calib(model, n_frames, seed=0, drop_fraction=0.1)
    lo, hi = pkg.dist.shard_frames(s.frame_offsets, rank, world)
    sh = pkg.dist.slice_problem(s, lo, hi)
    gp = pkg.Problem(model, s.width, s.height, sh["frame_offsets"], sh["x"], sh["y"], sh["z"], sh["u"], sh["v"], device=rank)
    pkg.dist.init_comm(gp, rank, world, deterministic=deterministic)
    gp.set_poses(s.init_poses[lo:hi])
    sq = gp.linearize(s.init_params)          # global cost on every rank
    intr, summ, hist = (gp.solve_gn if loop == "gn" else gp.solve_lm)(s.init_params)
    np.savez(os.path.join(out_dir, f"rank{rank}.npz"), intr=intr, poses=gp.get_poses(), iters=summ.iterations,
             status=summ.status, sq=sq, lo=lo, hi=hi, peer=pkg._abi.load().ccrs_comm_uses_peer_memory())
    gp.close()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(_n_gpus() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("loop,deterministic,p2p", [("gn", True, True), ("lm", True, True), ("lm", True, False), ("lm", False, True)])
def test_two_gpu_sharded_solve(pkg, oracle, tmp_path, loop, deterministic, p2p):
    """p2p=True: the exchange is fused into K2/K3 over peer memory (rank-order sum); p2p=False / deterministic=False:
    NCCL all-gather + rank-order sum / ncclAllReduce."""
    import torch.multiprocessing as mp
    model, n_frames, world = "eucm", 301, 2
    port = 29600 + (os.getpid() % 1000) + (0 if loop == "gn" else 1) + (0 if deterministic else 2) + (0 if p2p else 4)
    mp.spawn(_worker, args=(world, port, model, n_frames, loop, deterministic, str(tmp_path), p2p), nprocs=world, join=True)
    s = pkg.synth.make_calib(model, n_frames, seed=0, drop_fraction=0.1)
    op = oracle.OracleProblem.from_synth(s, 1)
    ref = (op.gauss_newton if loop == "gn" else op.levenberg_marquardt)(s.init_params, s.init_poses)
    r = [np.load(tmp_path / f"rank{k}.npz") for k in range(world)]
    assert all(int(x["status"]) == 0 for x in r)
    if not p2p:
        assert all(int(x["peer"]) == 0 for x in r)
    assert np.array_equal(r[0]["intr"], r[1]["intr"])                      # identical on every rank
    assert np.array_equal(r[0]["sq"], r[1]["sq"])
    assert abs(r[0]["sq"][0] - op.sq_error(s.init_params, s.init_poses)) / r[0]["sq"][0] < 1e-12
    assert int(r[0]["iters"]) == int(r[1]["iters"]) == ref[2].iterations   # equal iteration count
    assert np.max(np.abs(r[0]["intr"] - ref[0]) / np.abs(ref[0])) < 1e-6   # north_star tolerance
    poses = np.concatenate([x["poses"] for x in r])
    assert np.max(np.abs(poses - ref[1])) < 1e-6

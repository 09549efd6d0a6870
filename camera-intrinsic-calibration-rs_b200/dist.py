"""Frame sharding for multi-GPU runs: frames are independent given the shared intrinsics, so
each rank owns a contiguous range of frames balanced by observation count; the only exchange per linearisation is
the reduced intrinsic system (d*d + 3d + 1 doubles), done on the device by the library (NCCL all-gather + rank-order
sum). torch.distributed is used for plumbing only: rank/world discovery and shipping the ncclUniqueId."""
from __future__ import annotations

import numpy as np


def shard_frames(frame_offsets, rank: int, world: int):
    """Contiguous frame range [lo, hi) of `rank`, balancing observations (ragged frames) not frame counts."""
    fo = np.asarray(frame_offsets, dtype=np.int64)
    n_frames = len(fo) - 1
    if world > n_frames:
        raise ValueError(f"{world} ranks for {n_frames} frames: every rank needs at least one frame (an empty shard cannot be "
                         "created, and its peers would wait for its partial system)")
    total = fo[-1]
    cuts = [int(np.searchsorted(fo, total * r / world, side="left")) for r in range(world + 1)]
    cuts[0], cuts[-1] = 0, n_frames
    for i in range(1, world):       # at least one frame per rank, also when a few frames hold most of the observations
        cuts[i] = min(max(cuts[i], cuts[i - 1] + 1), n_frames - (world - i))
    return cuts[rank], cuts[rank + 1]


def slice_problem(s, lo: int, hi: int) -> dict:
    """Observation arrays of frames [lo, hi) of a synth.SyntheticCalib-like object."""
    a, b = int(s.frame_offsets[lo]), int(s.frame_offsets[hi])
    return dict(frame_offsets=(np.asarray(s.frame_offsets[lo:hi + 1]) - a).astype(np.int32), x=s.x[a:b], y=s.y[a:b],
                z=s.z[a:b], u=s.u[a:b], v=s.v[a:b])


def init_comm(problem, rank: int, world: int, deterministic: bool = True):
    """Create the library's NCCL communicator; the 128-byte unique id travels over torch.distributed."""
    import torch
    import torch.distributed as dist
    from .calib import comm_unique_id
    if world == 1:
        return
    obj = [comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(obj, src=0)
    problem.comm_init(obj[0], rank, world, deterministic)
    torch.cuda.synchronize()

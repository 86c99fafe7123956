"""Frame sharding across GPUs: frames are independent, every rank owns a contiguous
range of frames and its own canvas slice, weights are replicated, and there is NO collective on the
data path.  The only cross-rank traffic is the timing reduction at the end of a measurement.

Mirrors the reference's data parallelism (DistributedSampler + DDP, tools/train.py:130-162): one
process per GPU, each fed its own frames.
"""
from __future__ import annotations

import numpy as np


def frame_range(batch_size: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous, balanced partition of frames [0, batch_size): rank r owns [lo, hi)."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(batch_size, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_points(points: np.ndarray, frame_offsets: np.ndarray, lo: int, hi: int, batch_col: int | None = 0):
    """Rows of frames [lo, hi) with the batch index rebased to the shard, and the shard's frame_offsets."""
    a, b = int(frame_offsets[lo]), int(frame_offsets[hi])
    pts = points[a:b].copy()
    if batch_col is not None and pts.shape[0]:
        pts[:, batch_col] -= lo
    offs = (np.asarray(frame_offsets[lo:hi + 1]) - a).astype(np.int32)
    return pts, offs


def reduce_max(value: float, device=None) -> float:
    """max over ranks (timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(value: float, device=None) -> float:
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())

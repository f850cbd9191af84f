"""Multi-GPU host logic of the decompression path: one process per GPU, clips sharded across ranks.

The path has no exchange step: every (clip, sample_time) request is decoded by the rank that holds the clip, so the
only collectives are the barrier around a timed region and the reductions that turn per-rank numbers into whole-job numbers.
Nothing here decodes anything; it is index arithmetic plus thin wrappers over torch.distributed (NCCL on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def partition_clips(sizes, world: int):
    """Contiguous split of the clip list into `world` shards of about equal compressed bytes.

    Returns (owner, local_index, bounds): owner[c] = rank holding clip c, local_index[c] = its index inside that rank's clip set,
    bounds[r] = (first, last+1) clip of rank r. Contiguous so that a rank uploads one slice of the caller's clip table."""
    sizes = np.asarray(sizes, dtype=np.int64)
    n = len(sizes)
    if world <= 0:
        raise ValueError("world must be positive")
    cumulative = np.concatenate([[0], np.cumsum(sizes)])
    total = int(cumulative[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        cut = int(np.searchsorted(cumulative, target, side="left"))
        if cut > 0 and target - cumulative[cut - 1] < cumulative[min(cut, n)] - target:
            cut -= 1  # the boundary before is closer to the ideal split
        cut = min(max(cut, cuts[-1]), n)
        cuts.append(cut)
    cuts.append(n)
    owner = np.zeros(n, dtype=np.uint32)
    local_index = np.zeros(n, dtype=np.uint32)
    bounds = []
    for r in range(world):
        lo, hi = cuts[r], cuts[r + 1]
        owner[lo:hi] = r
        local_index[lo:hi] = np.arange(hi - lo, dtype=np.uint32)
        bounds.append((lo, hi))
    return owner, local_index, bounds


def route_requests(req_clip, req_time, owner, local_index, rank: int):
    """The requests rank `rank` decodes: (positions in the global request list, clip index inside the rank's clip set, sample times).
    Requests naming a clip outside the table stay with rank 0, which reports them as invalid exactly like a single GPU would."""
    req_clip = np.asarray(req_clip, dtype=np.uint32)
    req_time = np.asarray(req_time, dtype=np.float32)
    valid = req_clip < len(owner)
    request_owner = np.where(valid, owner[np.minimum(req_clip, max(len(owner) - 1, 0))] if len(owner) else 0, 0)
    mine = np.nonzero(request_owner == rank)[0]
    local = np.where(valid[mine], local_index[np.minimum(req_clip[mine], max(len(owner) - 1, 0))] if len(owner) else 0, np.uint32(0xFFFFFFFF))
    return mine, local.astype(np.uint32), req_time[mine]


def scatter_results(num_requests: int, row_shape, positions_per_rank, rows_per_rank, dtype=np.float32):
    """Reassembles per-rank pose rows into the global request order (what a host that wants everything in one place does)."""
    out = np.zeros((num_requests,) + tuple(row_shape), dtype=dtype)
    for positions, rows in zip(positions_per_rank, rows_per_rank):
        out[positions] = rows
    return out


class JobReducer:
    """Whole-job numbers from per-rank ones: time = max over ranks, units = sum over ranks."""

    def __init__(self, device=None):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.device = torch, dist, device
        self.active = dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1

    def barrier(self):
        if self.active:
            self.dist.barrier()

    def _reduce(self, value: float, op):
        if not self.active:
            return float(value)
        t = self.torch.tensor([value], dtype=self.torch.float64, device=self.device)
        self.dist.all_reduce(t, op=op)
        return float(t.item())

    def max(self, value: float) -> float:
        return self._reduce(value, self.dist.ReduceOp.MAX)

    def sum(self, value: float) -> float:
        return self._reduce(value, self.dist.ReduceOp.SUM)

    def throughput(self, units_this_rank: float, seconds_this_rank: float) -> float:
        """Whole-job units per second: every rank's units over the slowest rank's time."""
        return self.sum(units_this_rank) / self.max(seconds_this_rank)

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
    if n < world:
        # a rank without clips cannot build a clip set (aclb200_upload_clips refuses an empty list) and would leave the others
        # waiting in the job's collectives: refuse up front
        raise ValueError(f"{n} clips cannot be sharded over {world} ranks: every rank needs at least one clip")
    cumulative = np.concatenate([[0], np.cumsum(sizes)])
    total = int(cumulative[-1])
    cuts = [0]
    for r in range(1, world):
        target = total * r / world
        cut = int(np.searchsorted(cumulative, target, side="left"))
        if cut > 0 and target - cumulative[cut - 1] < cumulative[min(cut, n)] - target:
            cut -= 1  # the boundary before is closer to the ideal split
        # every shard keeps at least one clip, also when one huge clip swallows several ideal cut points
        cut = min(max(cut, cuts[-1] + 1), n - (world - r))
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


def route_error_jobs(jobs, owner, local_index, rank: int):
    """The compression error jobs (aclb200_error_job records, one per clip to measure: SURVEY 8 f1) rank `rank` runs: the measurement of a
    clip needs nothing but the clip, its raw poses and its skeleton, so it shards with the clips like the decode does. Returns (positions
    in the global job list, the jobs with `clip` rewritten to the index inside the rank's clip set). Raw pose / skeleton offsets are left
    as they are: they index whatever the rank holds in its own device buffers, which the caller lays out per rank."""
    jobs = np.asarray(jobs)
    job_clip = jobs["clip"].astype(np.int64)
    if np.any(job_clip >= len(owner)):
        raise ValueError("an error job names a clip outside the clip table")
    mine = np.nonzero(owner[job_clip] == rank)[0]
    routed = jobs[mine].copy()
    routed["clip"] = local_index[job_clip[mine]]
    return mine, routed


def reduce_worst_error(errors_this_rank, positions_this_rank, num_jobs: int, device=None):
    """Every rank's per job (index, error, sample_time, flags) records gathered into the global job order (one all_reduce of a table that
    is zero outside the rank's own jobs: jobs are disjoint across ranks)."""
    import torch
    import torch.distributed as dist
    table = np.zeros((num_jobs, 4), dtype=np.float64)
    records = np.asarray(errors_this_rank)
    if len(positions_this_rank):
        table[positions_this_rank, 0] = records["index"].astype(np.float64) + 1.0      # 0 = not this rank's job; 0xFFFFFFFF + 1 stays exact in f64
        table[positions_this_rank, 1] = records["error"]
        table[positions_this_rank, 2] = records["sample_time"]
        table[positions_this_rank, 3] = records["flags"]
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        t = torch.from_numpy(table).to(device) if device is not None else torch.from_numpy(table)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        table = t.cpu().numpy()
    out = np.zeros(num_jobs, dtype=np.dtype([("index", np.uint32), ("error", np.float32), ("sample_time", np.float32), ("flags", np.uint32)]))
    out["index"] = (table[:, 0] - 1.0).astype(np.int64).astype(np.uint32)
    out["error"] = table[:, 1]
    out["sample_time"] = table[:, 2]
    out["flags"] = table[:, 3].astype(np.uint32)
    return out


def exchange_plan(generated_bounds, owner_bounds, sizes):
    """Bytes every rank sends to every other rank when clips generated (or loaded) by contiguous ranges `generated_bounds` move to
    the owners `owner_bounds` picked by partition_clips: plan[src][dst] = (first clip, last clip + 1, bytes). Both partitions are
    contiguous and cover the same clip list, so what src holds for dst is one contiguous run (possibly empty)."""
    sizes = np.asarray(sizes, dtype=np.int64)
    cumulative = np.concatenate([[0], np.cumsum(sizes)])
    plan = []
    for g_lo, g_hi in generated_bounds:
        row = []
        for o_lo, o_hi in owner_bounds:
            lo, hi = max(g_lo, o_lo), min(g_hi, o_hi)
            if lo >= hi:
                lo = hi = g_lo
            row.append((int(lo), int(hi), int(cumulative[hi] - cumulative[lo])))
        plan.append(row)
    return plan


def redistribute_clips(local_bytes, rank: int, plan, device=None):
    """The batch split over the interconnect: every rank hands the compressed clips it holds to their owners with ONE
    all_to_all (NCCL over NVLink on GPUs, gloo in the CPU tests). `local_bytes`: uint8 tensor, this rank's clips back to back in
    clip order (no padding). Returns the uint8 tensor of the clips this rank owns, in clip order."""
    import torch
    import torch.distributed as dist
    world = len(plan)
    send_splits = [plan[rank][dst][2] for dst in range(world)]
    recv_splits = [plan[src][rank][2] for src in range(world)]
    assert int(local_bytes.numel()) == sum(send_splits), "local_bytes must hold exactly the clips of this rank's generated range"
    received = torch.empty(sum(recv_splits), dtype=torch.uint8, device=local_bytes.device if device is None else device)
    if world == 1:
        received.copy_(local_bytes)
        return received
    dist.all_to_all_single(received, local_bytes, output_split_sizes=recv_splits, input_split_sizes=send_splits)
    return received


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

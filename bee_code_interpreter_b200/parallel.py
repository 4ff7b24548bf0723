"""Multi-GPU plumbing: one process per GPU (`torch.distributed`, NCCL over NVLink/NVSwitch on
the GPU box, gloo in the CPU tests).

Steady state: every `/v1/execute` request is independent — the reference
spends a fresh pod per request (kubernetes_code_executor.py:93,248-264) — so request i runs
on GPU i mod G and there is NO data-path collective; rank 0 only gathers the scalar results /
stdout.  A collective appears only when ONE op's operand is split across GPUs (an array
larger than a GPU, or inputs already resident on different GPUs): each rank reduces / scans
its contiguous chunk with the normal kernels and the per-rank partials (<= 16 bytes each) are
combined here:
    sum / mean      all_reduce(SUM) of the partial sums (and counts)
    max / min       all_reduce(MAX/MIN); NaN on any rank wins, like numpy
    argmax / argmin all_gather of (value, global index) pairs, first-occurrence rule
    cumsum          all_gather of chunk totals -> exclusive offset of this rank's chunk
"""
from __future__ import annotations

import math
from typing import Any

import numpy as np


def shard(num_requests: int, rank: int, world: int) -> list[int]:
    """Round-robin placement: request i runs on GPU i mod G."""
    return list(range(rank, num_requests, world))


def owner(request_index: int, world: int) -> int:
    return request_index % world


def chunk_bounds(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous chunk [lo, hi) of an n-element operand held by `rank` (split-operand ops)."""
    base, extra = divmod(n, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def _device(dist):
    import torch

    return torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")


def gather_results(local: dict[int, Any], dist, dst: int = 0) -> dict[int, Any] | None:
    """Collects {request index: result} from every rank on `dst` (control plane, not timed)."""
    bucket = [None] * dist.get_world_size() if dist.get_rank() == dst else None
    dist.gather_object(local, bucket, dst=dst)
    if bucket is None:
        return None
    merged: dict[int, Any] = {}
    for part in bucket:
        merged.update(part)
    return merged


def max_over_ranks(value: float, dist) -> float:
    import torch

    t = torch.tensor([value], dtype=torch.float64, device=_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# ---- split-operand combines ------------------------------------------------------------------
def split_sum(local_sum, dist, integer: bool = False):
    import torch

    t = torch.tensor([local_sum], dtype=torch.int64 if integer else torch.float64, device=_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item()) if integer else float(t.item())


def split_mean(local_sum: float, local_n: int, dist) -> float:
    import torch

    t = torch.tensor([float(local_sum), float(local_n)], dtype=torch.float64, device=_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0].item() / t[1].item())


def split_extreme(local_value: float, dist, maximum: bool = True) -> float:
    """numpy.max/min over the whole operand: NaN anywhere makes the result NaN."""
    import torch

    nan = 1.0 if (isinstance(local_value, float) and math.isnan(local_value)) else 0.0
    v = 0.0 if nan else float(local_value)
    t = torch.tensor([v if maximum else -v, nan], dtype=torch.float64, device=_device(dist))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if t[1].item() > 0:
        return float("nan")
    return float(t[0].item()) if maximum else -float(t[0].item())


def split_argext(local_value: float, local_index: int, chunk_offset: int, dist, maximum: bool = True) -> int:
    """numpy.argmax/argmin over the whole operand from per-chunk (value, index) winners:
    first NaN wins, then the extreme value, ties to the lowest global index."""
    import torch

    world = dist.get_world_size()
    mine = torch.tensor([float(local_value), float(chunk_offset + local_index)], dtype=torch.float64, device=_device(dist))
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    vals = np.array([float(t[0].item()) for t in allv])
    idxs = np.array([int(t[1].item()) for t in allv])
    nan = np.isnan(vals)
    if nan.any():
        return int(idxs[nan].min())
    best = vals.max() if maximum else vals.min()
    return int(idxs[vals == best].min())


def split_cumsum_offset(local_total, dist, integer: bool = False):
    """Exclusive prefix of the chunk totals: what this rank adds to its local cumsum."""
    import torch

    world, rank = dist.get_world_size(), dist.get_rank()
    mine = torch.tensor([local_total], dtype=torch.int64 if integer else torch.float64, device=_device(dist))
    allv = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    acc = 0 if integer else 0.0
    for r in range(rank):
        acc += int(allv[r].item()) if integer else float(allv[r].item())
    return acc

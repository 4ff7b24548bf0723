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


def split_sort(local, dist, sort_local, oversample: int = 64):
    """numpy.sort of ONE array whose contiguous chunks live on the ranks ("sort →
    sample-sort with one all-to-all bucket exchange").

    `local` is this rank's chunk as a 1-D torch tensor (CUDA under NCCL, CPU under gloo);
    `sort_local(tensor) -> tensor` is the single-GPU sort (the radix kernel through the C ABI
    on the GPU box).  Returns this rank's bucket of the global order: concatenating the
    results by rank gives numpy.sort(whole array), NaNs last (they all land on the last rank).
    Steps: local sort -> `oversample` regular samples per rank, all_gather -> W-1 splitters ->
    cut points by binary search in the sorted chunk -> all_to_all of counts, then of keys ->
    sort of the received runs.  torch is the plumbing (sample gather, cut search over W-1
    values, NCCL exchange); both sorts are the kernel's."""
    import torch

    world = dist.get_world_size()
    s = sort_local(local)
    n = int(s.numel())
    n_valid = n - int(torch.isnan(s).sum().item()) if s.is_floating_point() else n   # NaNs sit at the end
    k = oversample
    have = torch.tensor([1 if n_valid else 0], dtype=torch.int64, device=s.device)
    if n_valid:
        pick = ((torch.arange(1, k + 1, device=s.device, dtype=torch.int64) * n_valid) // (k + 1)).clamp_(max=n_valid - 1)
        samples = s[pick].contiguous()
    else:
        samples = torch.zeros(k, dtype=s.dtype, device=s.device)
    all_samples = [torch.empty_like(samples) for _ in range(world)]
    all_have = [torch.empty_like(have) for _ in range(world)]
    dist.all_gather(all_samples, samples)
    dist.all_gather(all_have, have)
    pool = [t for t, h in zip(all_samples, all_have) if int(h.item())]
    if pool:
        pooled = torch.sort(torch.cat(pool)).values                      # <= W*k values: plumbing
        at = (torch.arange(1, world, device=s.device, dtype=torch.int64) * pooled.numel()) // world
        splitters = pooled[at]
        cuts = torch.searchsorted(s[:n_valid].contiguous(), splitters, right=True)
    else:
        cuts = torch.zeros(world - 1, dtype=torch.int64, device=s.device)
    bounds = [0] + [int(c) for c in cuts.tolist()] + [n]                 # NaN tail goes with the last bucket
    send_counts = [bounds[i + 1] - bounds[i] for i in range(world)]
    sc = torch.tensor(send_counts, dtype=torch.int64, device=s.device)
    rc = torch.empty_like(sc)
    dist.all_to_all_single(rc, sc)
    recv_counts = [int(c) for c in rc.tolist()]
    recv = torch.empty(sum(recv_counts), dtype=s.dtype, device=s.device)
    if dist.get_backend() == "gloo":
        # gloo has no uneven all_to_all_single: W rounds of point-to-point (test path only)
        ins = list(torch.split(s, send_counts))
        outs = list(torch.split(recv, recv_counts))
        rank = dist.get_rank()
        outs[rank].copy_(ins[rank])
        reqs = []
        for peer in range(world):
            if peer != rank:
                reqs.append(dist.isend(ins[peer].contiguous(), peer))
                reqs.append(dist.irecv(outs[peer], peer))
        for r in reqs:
            r.wait()
    else:
        dist.all_to_all_single(recv, s, output_split_sizes=recv_counts, input_split_sizes=send_counts)
    return sort_local(recv)


def split_matmul_rows(a_rows, b_full, dist, matmul_local, src: int = 0):
    """`A @ B` with A split by rows over the ranks and B held by `src` ("matmul →
    row-split A, replicate B (broadcast), no reduction").  `b_full` is a tensor of B's shape
    on every rank (contents significant on `src` only).  Returns this rank's rows of C."""
    dist.broadcast(b_full, src=src)
    return matmul_local(a_rows, b_full)

"""Split-operand self-check on real GPUs (SURVEY.md §8e "when is a collective used"): ONE operand is split
across the ranks; every rank runs the normal kernels through the C ABI on its chunk and `parallel.py` combines
the per-rank partials over NCCL.  Results are compared with numpy on the whole operand.  Used by
tests/test_gpu_multi.py (torchrun, 2 GPUs) and by bench.py at WORLD_SIZE >= 2, so the scaling run carries
hardware evidence of the collective path (`split_operand_parity` in its JSON line).
Device memory for the exchanged arrays is torch's (the C ABI takes plain device pointers)."""
from __future__ import annotations

import numpy as np

from . import _lib as L
from . import parallel as par


def run(dist, dev, rank: int, world: int) -> dict | None:
    """Every rank calls this inside an initialised NCCL process group with `dev` bound to its GPU.
    Rank 0 returns the flags, the others None."""
    import torch

    NP = {torch.float32: np.float32, torch.float64: np.float64, torch.int64: np.int64, torch.int32: np.int32}
    out = {}

    def sort_local(t):
        t = t.contiguous()
        out = torch.empty_like(t)
        if t.numel():
            torch.cuda.synchronize()               # torch stream -> library stream
            dev.sort(NP[t.dtype], [t.data_ptr()], [out.data_ptr()], [t.numel()])
            dev.sync()
        return out

    def matmul_local(a, b):
        c = torch.empty(a.shape[0], b.shape[1], dtype=torch.float32, device=a.device)
        torch.cuda.synchronize()
        dev.matmul_f32([a.data_ptr()], [b.data_ptr()], [c.data_ptr()], a.shape[0], b.shape[1], a.shape[1])
        dev.sync()
        return c

    n = 40_000_003
    full = np.random.default_rng(99).standard_normal(n)
    full[30_000_001] = full.max() + 1.0
    full[123] = full.min() - 1.0
    lo, hi = par.chunk_bounds(n, rank, world)
    chunk = torch.from_numpy(full[lo:hi].copy()).cuda()
    torch.cuda.synchronize()
    p, m = chunk.data_ptr(), chunk.numel()
    local_sum = float(dev.reduce(L.RED_SUM, np.float64, [p], [m])[0])
    s = par.split_sum(local_sum, dist)
    mean = par.split_mean(local_sum, m, dist)
    mx = par.split_extreme(float(dev.reduce(L.RED_MAX, np.float64, [p], [m])[0]), dist, maximum=True)
    mn = par.split_extreme(float(dev.reduce(L.RED_MIN, np.float64, [p], [m])[0]), dist, maximum=False)
    li = int(dev.argext(L.ARGMAX, np.float64, [p], [m])[0])
    am = par.split_argext(float(full[lo + li]), li, lo, dist, maximum=True)
    li = int(dev.argext(L.ARGMIN, np.float64, [p], [m])[0])
    an = par.split_argext(float(full[lo + li]), li, lo, dist, maximum=False)

    ints = np.random.default_rng(5).integers(-2**40, 2**40, 10_000_001, dtype=np.int64)
    ilo, ihi = par.chunk_bounds(ints.size, rank, world)
    ichunk = torch.from_numpy(ints[ilo:ihi].copy()).cuda()
    iout = torch.empty_like(ichunk)
    torch.cuda.synchronize()
    dev.cumsum(np.int64, [ichunk.data_ptr()], [iout.data_ptr()], [ichunk.numel()])
    total = int(dev.reduce(L.RED_SUM, np.int64, [ichunk.data_ptr()], [ichunk.numel()])[0])
    off = par.split_cumsum_offset(total, dist, integer=True)
    dev.ewise(L.EW_ADD, np.int64, [iout.data_ptr()], None, [off], [iout.data_ptr()], [iout.numel()])
    dev.sync()
    cs_ok = bool(np.array_equal(iout.cpu().numpy(), np.cumsum(ints)[ilo:ihi]))

    fs = np.random.default_rng(6).standard_normal(20_000_001).astype(np.float32)
    fs[[7, 9_000_000, 19_999_999]] = np.nan
    flo, fhi = par.chunk_bounds(fs.size, rank, world)
    mine = par.split_sort(torch.from_numpy(fs[flo:fhi].copy()).cuda(), dist, sort_local)
    sizes = [None] * world
    dist.all_gather_object(sizes, int(mine.numel()))
    want = np.sort(fs)
    start = sum(sizes[:rank])
    sort_ok = bool(np.array_equal(mine.cpu().numpy(), want[start:start + sizes[rank]], equal_nan=True))

    M, K, N = 1024, 512, 768
    a = np.random.default_rng(8).random((M, K), dtype=np.float32)
    b = np.random.default_rng(9).random((K, N), dtype=np.float32)
    rlo, rhi = par.chunk_bounds(M, rank, world)
    bt = torch.from_numpy(b.copy()).cuda() if rank == 0 else torch.zeros(K, N, device="cuda")
    crow = par.split_matmul_rows(torch.from_numpy(a[rlo:rhi].copy()).cuda(), bt, dist, matmul_local)
    ref = a[rlo:rhi].astype(np.float64) @ b.astype(np.float64)
    mm_ok = bool(np.abs(crow.cpu().numpy() - ref).max() <= 1e-5 * np.abs(ref).max())

    flags = [cs_ok, sort_ok, mm_ok]
    allf = [None] * world
    dist.all_gather_object(allf, flags)
    if rank == 0:
        out.update({
            "world": world,
            "sum_ok": abs(s - full.sum()) <= 1e-12 * np.abs(full).sum(),
            "mean_ok": abs(mean - full.mean()) <= 1e-12,
            "max_ok": mx == full.max(), "min_ok": mn == full.min(),
            "argmax_ok": am == int(np.argmax(full)) == 30_000_001, "argmin_ok": an == int(np.argmin(full)) == 123,
            "cumsum_ok": all(f[0] for f in allf), "sort_ok": all(f[1] for f in allf), "matmul_ok": all(f[2] for f in allf),
            "sort_sizes": sizes, "sort_total_ok": sum(sizes) == fs.size,
            "launches": dev.launch_count(),
        })
        out = {k: (bool(v) if isinstance(v, (bool, np.bool_)) else v) for k, v in out.items()}
        return out
    return None

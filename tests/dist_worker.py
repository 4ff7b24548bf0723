"""Worker for tests/test_parallel.py: world_size-2 gloo run of the N>1 host logic.  Local
"device" work is done by tests/fake_device.py (numpy); what is under test is the sharding and
the cross-rank combination in bee_code_interpreter_b200/parallel.py."""
import json
import os
import sys
from pathlib import Path

import numpy as np
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
sys.path.insert(0, str(Path(__file__).resolve().parent))
import fake_device  # noqa: E402

from bee_code_interpreter_b200 import _lib as L  # noqa: E402
from bee_code_interpreter_b200 import parallel as par  # noqa: E402


def main():
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    out = {}

    # steady state: round-robin requests, no collective on the data path
    R = 11
    mine = par.shard(R, rank, world)
    local = {}
    for i in mine:
        x = np.random.default_rng(2000 + i).random(5000, dtype=np.float32)
        p = fake_device.upload(x)
        local[i] = float(fake_device.reduce(L.RED_SUM, np.float32, [p], [x.size], pre=L.PRE_SQUARE)[0])
        fake_device.free(p)
    merged = par.gather_results(local, dist)
    if rank == 0:
        want = {i: float(np.sum(np.square(np.random.default_rng(2000 + i).random(5000, dtype=np.float32)))) for i in range(R)}
        out["round_robin_ok"] = sorted(merged) == list(range(R)) and all(abs(merged[i] - want[i]) <= 1e-5 * want[i] for i in range(R))
        out["owners"] = [par.owner(i, world) for i in range(R)]

    # split operand: one array spans the ranks
    n = 100_003
    full = np.random.default_rng(99).standard_normal(n)
    full[70_001] = full.max() + 1.0        # unique max in rank 1's chunk
    full[123] = full.min() - 1.0           # unique min in rank 0's chunk
    lo, hi = par.chunk_bounds(n, rank, world)
    chunk = full[lo:hi].copy()
    p = fake_device.upload(chunk)
    s = par.split_sum(float(fake_device.reduce(L.RED_SUM, np.float64, [p], [chunk.size])[0]), dist)
    m = par.split_mean(float(fake_device.reduce(L.RED_SUM, np.float64, [p], [chunk.size])[0]), chunk.size, dist)
    mx = par.split_extreme(float(fake_device.reduce(L.RED_MAX, np.float64, [p], [chunk.size])[0]), dist, maximum=True)
    mn = par.split_extreme(float(fake_device.reduce(L.RED_MIN, np.float64, [p], [chunk.size])[0]), dist, maximum=False)
    li = int(fake_device.argext(L.ARGMAX, np.float64, [p], [chunk.size])[0])
    am = par.split_argext(float(chunk[li]), li, lo, dist, maximum=True)
    li = int(fake_device.argext(L.ARGMIN, np.float64, [p], [chunk.size])[0])
    an = par.split_argext(float(chunk[li]), li, lo, dist, maximum=False)
    off = par.split_cumsum_offset(float(chunk.sum()), dist)
    q = fake_device.alloc(chunk.nbytes)
    fake_device.cumsum(np.float64, [p], [q], [chunk.size])
    local_cs = fake_device.download(q, chunk.size, np.float64) + off
    cs_ok = bool(np.allclose(local_cs, np.cumsum(full)[lo:hi], rtol=1e-12, atol=1e-9))
    # NaN on one rank only
    nanv = par.split_extreme(float("nan") if rank == 1 else 1.0, dist, maximum=True)
    nan_arg = par.split_argext(float("nan") if rank == 1 else 5.0, 7, lo, dist, maximum=True)
    ints = par.split_sum(int(rank + 1) * (2**40), dist, integer=True)
    # split sort: sample sort over the ranks (local sorts = the device sort), NaNs last
    import torch
    fs = np.random.default_rng(5).standard_normal(50_001).astype(np.float32)
    fs[[7, 30_000, 49_999]] = np.nan
    flo, fhi = par.chunk_bounds(fs.size, rank, world)

    def sort_local(t):
        h = t.numpy()
        pi, po = fake_device.upload(h), fake_device.alloc(h.nbytes)
        fake_device.sort(h.dtype, [pi], [po], [h.size])
        out = fake_device.download(po, h.size, h.dtype)
        fake_device.free(pi), fake_device.free(po)
        return torch.from_numpy(out)

    mine_sorted = par.split_sort(torch.from_numpy(fs[flo:fhi].copy()), dist, sort_local, oversample=16).numpy()
    parts = [None] * world
    dist.all_gather_object(parts, mine_sorted)
    ks = np.random.default_rng(6).integers(-5, 5, 20_000, dtype=np.int64)       # heavy duplicates
    klo, khi = par.chunk_bounds(ks.size, rank, world)
    kparts = [None] * world
    dist.all_gather_object(kparts, par.split_sort(torch.from_numpy(ks[klo:khi].copy()), dist, sort_local, oversample=8).numpy())
    eparts = [None] * world                                                       # one rank has nothing
    dist.all_gather_object(eparts, par.split_sort(torch.from_numpy(ks[:0] if rank == 0 else ks[:100].copy()), dist, sort_local).numpy())
    # split matmul: rows of A per rank, B broadcast from rank 0
    mat_a = np.random.default_rng(8).random((10, 6), dtype=np.float32)
    bm = np.random.default_rng(9).random((6, 4), dtype=np.float32)
    rlo, rhi = par.chunk_bounds(10, rank, world)
    bt = torch.from_numpy(bm.copy()) if rank == 0 else torch.zeros(6, 4)
    crow = par.split_matmul_rows(torch.from_numpy(mat_a[rlo:rhi].copy()), bt, dist, lambda a, b: a @ b)
    cparts = [None] * world
    dist.all_gather_object(cparts, crow.numpy())
    flags = [cs_ok]
    allf = [None] * world
    dist.all_gather_object(allf, flags)
    if rank == 0:
        out.update({
            "sum_ok": abs(s - full.sum()) <= 1e-9 * np.abs(full).sum(),
            "mean_ok": abs(m - full.mean()) <= 1e-12,
            "max_ok": mx == full.max(), "min_ok": mn == full.min(),
            "argmax_ok": am == int(np.argmax(full)) == 70_001, "argmin_ok": an == int(np.argmin(full)) == 123,
            "cumsum_ok": all(f[0] for f in allf),
            "nan_ok": bool(np.isnan(nanv)) and nan_arg == par.chunk_bounds(n, 1, world)[0] + 7,
            "int_ok": ints == sum((r + 1) * 2**40 for r in range(world)),
            "chunks": [par.chunk_bounds(n, r, world) for r in range(world)],
            "sort_ok": bool(np.array_equal(np.concatenate(parts), np.sort(fs), equal_nan=True)),
            "sort_balanced": min(len(p) for p in parts) > fs.size // (4 * world),
            "sort_dups_ok": bool(np.array_equal(np.concatenate(kparts), np.sort(ks))),
            "sort_empty_rank_ok": bool(np.array_equal(np.concatenate(eparts), np.sort(ks[:100]))) if world == 2 else True,
            "matmul_ok": bool(np.allclose(np.concatenate(cparts), mat_a @ bm, rtol=1e-6)),
        })
        out = {k: (bool(v) if isinstance(v, (bool, np.bool_)) else v) for k, v in out.items()}
        print("RESULT " + json.dumps(out), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()

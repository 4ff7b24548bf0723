#!/usr/bin/env python
"""cumsum forms side by side (CUDA events on the library stream), or one form for ncu:
    python tools/prof_scan.py            # table: legacy / lookback / pipe, i64 + f32 + f64
    BCI_SCAN=pipe python tools/prof_scan.py --once   # a few launches of one form (ncu target)"""
import os

os.environ.setdefault("BCI_TUNING", "1")      # the kernels' A/B switches are honoured only in tuning mode
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import device as dev  # noqa: E402

dev.init(0)
B, n = 32, 10**7
rng = np.random.default_rng(5000)
PEAK = 6569.6


def run(dtype, mode, iters=5):
    if mode == "default":
        os.environ.pop("BCI_SCAN", None)
    else:
        os.environ["BCI_SCAN"] = mode
    odt = dev.cumsum_out_dtype(dtype)
    h = rng.integers(0, 2**31, n).astype(dtype) if np.issubdtype(dtype, np.integer) else rng.random(n).astype(dtype)
    k = [dev.upload(np.roll(h, i * 1001)) for i in range(B)]
    o = [dev.alloc(n * odt.itemsize) for _ in range(B)]
    for _ in range(2):
        dev.cumsum(dtype, k, o, [n] * B)
    dev.sync()
    best = 1e9
    for _ in range(iters):
        dev.event_record(0)
        dev.cumsum(dtype, k, o, [n] * B)
        dev.event_record(1)
        dev.sync()
        best = min(best, dev.event_elapsed(0, 1))
    got = dev.download(o[3], n, odt)
    want = np.cumsum(np.roll(h, 3 * 1001), dtype=np.float64 if odt.kind == "f" else odt)
    ok = np.array_equal(got, want) if odt.kind == "i" else np.allclose(got, want.astype(odt), rtol=1e-6)
    for p in k + o:
        dev.free(p)
    alg = n * B * (np.dtype(dtype).itemsize + odt.itemsize)
    return best, alg / best / 1e6, ok


if "--once" in sys.argv:
    print(run(np.int64, os.environ.get("BCI_SCAN", "default"), iters=2))
else:
    for dtype in (np.int64, np.float64, np.float32, np.int32):
        for mode in ("legacy", "lookback", "pipe"):
            ms, gbs, ok = run(dtype, mode)
            print(f"cumsum {np.dtype(dtype).name:8s} {mode:9s} {ms:8.3f} ms  {gbs:8.1f} GB/s alg  frac {gbs / PEAK:5.2f}  ok={ok}", flush=True)

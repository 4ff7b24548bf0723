#!/usr/bin/env python
"""Sort throughput at C5 shape (f32 1e7 x 32; also i64, argsort), CUDA events on the library stream."""
import os
import sys

os.environ.setdefault("BCI_TUNING", "1")      # the kernels' A/B switches are honoured only in tuning mode
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import device as dev  # noqa: E402

dev.init(0)
B, n = 32, 10**7
rng = np.random.default_rng(4000)
hv = rng.random(n, dtype=np.float32)
hk = rng.integers(0, 2**31, n, dtype=np.int64)


def run(name, fn, dtype, h, out_item, check):
    v = [dev.upload(np.roll(h, i * 1001)) for i in range(B)]
    o = [dev.alloc(n * out_item) for _ in range(B)]
    fn(dtype, v, o, [n] * B)
    dev.sync()
    best = 1e9
    for _ in range(3):
        dev.event_record(0)
        fn(dtype, v, o, [n] * B)
        dev.event_record(1)
        dev.sync()
        best = min(best, dev.event_elapsed(0, 1))
    ok = check(o[5], np.roll(h, 5 * 1001))
    for p in v + o:
        dev.free(p)
    print(f"{name:14s} {best:8.3f} ms  {n * B / best / 1e6:6.1f} Gkeys/s  ok={ok}", flush=True)


run("sort f32", dev.sort, np.float32, hv, 4, lambda p, h: np.array_equal(dev.download(p, n, np.float32), np.sort(h)))
run("sort i64", dev.sort, np.int64, hk, 8, lambda p, h: np.array_equal(dev.download(p, n, np.int64), np.sort(h)))
run("argsort f32", dev.argsort, np.float32, hv, 8, lambda p, h: np.array_equal(h[dev.download(p, n, np.int64)], np.sort(h)))

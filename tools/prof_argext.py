#!/usr/bin/env python
"""argmax/argmin at C5 shape (1e7 x 32), CUDA events on the library stream."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import _lib as L, device as dev  # noqa: E402

dev.init(0)
B, n = 32, 10**7
rng = np.random.default_rng(4000)
for dtype in (np.float32, np.float64, np.int32, np.int64):
    h = rng.random(n).astype(dtype) if np.dtype(dtype).kind == "f" else rng.integers(-2**30, 2**30, n).astype(dtype)
    v = [dev.upload(np.roll(h, i * 1001)) for i in range(B)]
    for op, ref in ((L.ARGMAX, np.argmax), (L.ARGMIN, np.argmin)):
        for _ in range(2):
            r = dev.argext(op, dtype, v, [n] * B)
        best = 1e9
        for _ in range(5):
            dev.flush_l2()
            dev.event_record(0)
            r = dev.argext(op, dtype, v, [n] * B)
            dev.event_record(1)
            dev.sync()
            best = min(best, dev.event_elapsed(0, 1))
        ok = int(r[3]) == int(ref(np.roll(h, 3 * 1001)))
        gbs = n * B * np.dtype(dtype).itemsize / best / 1e6
        print(f"{'argmax' if op == L.ARGMAX else 'argmin'} {np.dtype(dtype).name:8s} {best:7.3f} ms {gbs:8.1f} GB/s frac {gbs / 6569.6:4.2f} ok={ok}", flush=True)
    for p in v:
        dev.free(p)

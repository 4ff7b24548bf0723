#!/usr/bin/env python
"""C3's three reductions at the config's shape (1024 requests x 2^20 float32, two operands), a few launches each:
run under ncu for the kernel-side numbers; prints the event-timed launch-to-result times beside them."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import _lib as L, device as dev  # noqa: E402

dev.init(0)
n, b = 1 << 20, 1024
st = np.random.PCG64(1).state["state"]
g = (st["state"], st["inc"])
X = [dev.alloc(n * 4) for _ in range(b)]
Y = [dev.alloc(n * 4) for _ in range(b)]
dev.rand_pcg64(L.PCG_F32, X, [n] * b, [g] * b)
dev.rand_pcg64(L.PCG_F32, Y, [n] * b, [g] * b)
for name, fn in (("sum_square", lambda: dev.reduce(L.RED_SUM, np.float32, X, [n] * b, pre=L.PRE_SQUARE)),
                 ("mean_mul_add", lambda: dev.reduce_fused(L.RED_MEAN, L.EW_MUL, np.float32, X, Y, None, [n] * b, addends=[1.0] * b)),
                 ("max_abs_sub", lambda: dev.reduce_fused(L.RED_MAX, L.EW_SUB, np.float32, X, Y, None, [n] * b, pre=L.PRE_ABS))):
    ms = []
    for _ in range(4):
        dev.event_record(20)
        fn()
        dev.event_record(21)
        ms.append(dev.event_elapsed(20, 21))
    print(name, "event ms per call:", [round(m, 3) for m in ms], flush=True)

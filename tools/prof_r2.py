#!/usr/bin/env python
"""One launch of every round-2 kernel at a profiling size (run under ncu): fused rand+reduce, PCG64 draws,
fill / cast, and the C3 fused reductions on Generator-born inputs."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import _lib as L, device as dev  # noqa: E402

dev.init(0)
B, n = 16, 10**8
for _ in range(2):
    dev.reduce_rand_f64(L.RED_SUM, [n] * B, list(range(B)), pre=L.PRE_SQUARE)
st = np.random.PCG64(1).state["state"]
g = (st["state"], st["inc"])
m = 1 << 26
bufs = [dev.alloc(m * 8) for _ in range(8)]
for _ in range(2):
    dev.rand_pcg64(L.PCG_F64, bufs, [m] * 8, [g] * 8)
    dev.rand_pcg64(L.PCG_F32, bufs, [2 * m] * 8, [g] * 8)
    dev.rand_pcg64(L.PCG_BOUNDED32_I64, bufs, [m] * 8, [g + (0, 2**31)] * 8)
    dev.fill(L.FILL_ARANGE, np.float64, bufs, [m] * 8, [0.5] * 8, [0.25] * 8, [0.75] * 8)
    dev.fill(L.FILL_CONST, np.float32, bufs, [2 * m] * 8, [1.0] * 8)
half = [dev.alloc(m * 4) for _ in range(8)]
for _ in range(2):
    dev.cast(np.float64, np.float32, bufs, half, [m] * 8)
dev.sync()
print("launches", dev.launch_count())

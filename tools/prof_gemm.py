#!/usr/bin/env python
"""tcgen05 3xTF32 GEMM at C2 shape (4096^3 f32) under different rasterisation groups.
    python tools/prof_gemm.py [groups...]      default: 1 4 8 16 32"""
import os

os.environ.setdefault("BCI_TUNING", "1")      # the kernels' A/B switches are honoured only in tuning mode
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import device as dev  # noqa: E402

dev.init(0)
m, B = 4096, 8
rng = np.random.default_rng(1000)
ha, hb = rng.random((m, m), dtype=np.float32), rng.random((m, m), dtype=np.float32)
a = [dev.upload(ha) for _ in range(B)]
b = [dev.upload(hb) for _ in range(B)]
c = [dev.alloc(m * m * 4) for _ in range(B)]
ref = (ha[:64].astype(np.float64) @ hb.astype(np.float64))
for g in [int(x) for x in sys.argv[1:]] or [1, 4, 8, 16, 32]:
    os.environ["BCI_GEMM_GROUP"] = str(g)
    for _ in range(2):
        dev.matmul_f32(a, b, c, m, m, m)
    dev.sync()
    best = 1e9
    for _ in range(5):
        dev.event_record(0)
        dev.matmul_f32(a, b, c, m, m, m)
        dev.event_record(1)
        dev.sync()
        best = min(best, dev.event_elapsed(0, 1))
    got = dev.download(c[B - 1], 64 * m, np.float32).reshape(64, m)
    err = float(np.abs(got - ref).max() / np.abs(ref).max())
    print(f"group {g:3d}: {best:8.3f} ms  {2 * m**3 * B / best / 1e9:8.1f} TFLOP/s algorithmic  err {err:.2e}", flush=True)

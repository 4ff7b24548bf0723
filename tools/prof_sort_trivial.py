#!/usr/bin/env python
"""Sort rate on keys whose high bytes are constant (copy passes) vs full-entropy keys, int64 and float32."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import device as dev  # noqa: E402

dev.init(0)
n, B = 10**7, 16
rng = np.random.default_rng(0)
cases = {
    "int64 in [0, 2^31)   (4 of 8 passes constant)": rng.integers(0, 2**31, n, dtype=np.int64),
    "int64 in [0, 2^16)   (6 of 8 passes constant)": rng.integers(0, 2**16, n, dtype=np.int64),
    "int64 full range     (no constant pass)": rng.integers(-2**63, 2**63 - 1, n, dtype=np.int64),
    "float32 in [1, 2)    (top byte constant)": (1.0 + rng.random(n, dtype=np.float32)),
    "float32 in [0, 1)    (no constant pass)": rng.random(n, dtype=np.float32),
    "int32 in [0, 256)    (3 of 4 passes constant)": rng.integers(0, 256, n, dtype=np.int32),
}
for name, a in cases.items():
    ins = [dev.upload(a) for _ in range(B)]
    outs = [dev.alloc(a.nbytes) for _ in range(B)]
    ms = []
    for r in range(4):
        dev.event_record(20)
        dev.sort(a.dtype, ins, outs, [n] * B)
        dev.event_record(21)
        ms.append(dev.event_elapsed(20, 21))
    ok = np.array_equal(dev.download(outs[0], n, a.dtype), np.sort(a))
    print(f"{name}: {min(ms[1:]):.3f} ms for {B} x 1e7 keys = {B * n / min(ms[1:]) / 1e6:.1f} Gkeys/s  bit-exact={ok}", flush=True)
    for p in ins + outs:
        dev.free(p)

#!/usr/bin/env python
"""Small driver for ncu: one batched sort (and cumsum/argmax) launch sequence."""
import os
import sys

os.environ.setdefault("BCI_TUNING", "1")      # the kernels' A/B switches are honoured only in tuning mode
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import _lib as L, device as dev

dev.init(0)
B, n = 8, 10**7
rng = np.random.default_rng(4000)
v = [dev.upload(rng.random(n, dtype=np.float32)) for _ in range(B)]
o = [dev.alloc(n * 4) for _ in range(B)]
ns = [n] * B
for _ in range(2):
    dev.sort(np.float32, v, o, ns)
    dev.argext(L.ARGMAX, np.float32, v, ns)
dev.sync()
print("done")

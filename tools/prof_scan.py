#!/usr/bin/env python
"""Small driver for ncu: one batched cumsum launch (int64, 8 x 1e7)."""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import device as dev

dev.init(0)
B, n = 8, 10**7
rng = np.random.default_rng(5000)
k = [dev.upload(rng.integers(0, 2**31, n, dtype=np.int64)) for _ in range(B)]
o = [dev.alloc(n * 8) for _ in range(B)]
for _ in range(3):
    dev.cumsum(np.int64, k, o, [n] * B)
dev.sync()
print("done")

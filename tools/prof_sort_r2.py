#!/usr/bin/env python
"""Two sorts of 8 x 1e7 float32 keys (run under ncu: the second one's middle sweep is the launch to read)."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import device as dev  # noqa: E402

dev.init(0)
n, B = 10**7, 8
a = np.random.default_rng(0).standard_normal(n).astype(np.float32)
ins = [dev.upload(a) for _ in range(B)]
outs = [dev.alloc(a.nbytes) for _ in range(B)]
for _ in range(2):
    dev.sort(np.float32, ins, outs, [n] * B)
dev.sync()
print("ok", np.array_equal(dev.download(outs[0], n, np.float32), np.sort(a)))

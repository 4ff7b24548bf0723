#!/usr/bin/env python
"""One launch of every kernel class at BASELINE-config sizes, for a single
`ncu --set full` capture (profiles/r1_ncu_all_kernels.csv is its summary)."""
import os

os.environ.setdefault("BCI_TUNING", "1")      # the kernels' A/B switches are honoured only in tuning mode
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import _lib as L, device as dev  # noqa: E402

dev.init(0)
rng = np.random.default_rng(0)

# C1: 1e8 f64 x 4
n1, B1 = 10**8, 4
a = [dev.alloc(n1 * 8) for _ in range(B1)]
b = [dev.alloc(n1 * 8) for _ in range(B1)]
dev.rand_f64(a, [n1] * B1, list(range(B1)))
dev.ewise(L.EW_SQUARE, np.float64, a, None, None, b, [n1] * B1)
dev.ewise(L.EW_ADD, np.float64, a, b, None, b, [n1] * B1)
dev.reduce(L.RED_SUM, np.float64, a, [n1] * B1)
dev.reduce(L.RED_SUM, np.float64, a, [n1] * B1, pre=L.PRE_SQUARE)
for p in a + b:
    dev.free(p)

# C3: 2^20 f32 x 256
n3, B3 = 2**20, 256
h = rng.random(n3, dtype=np.float32)
x = [dev.upload(h) for _ in range(B3)]
y = [dev.upload(h[::-1].copy()) for _ in range(B3)]
o = [dev.alloc(n3 * 4) for _ in range(B3)]
ns = [n3] * B3
dev.ewise(L.EW_MUL, np.float32, x, y, None, o, ns)
dev.ewise(L.EW_ADD, np.float32, x, None, [1.0] * B3, o, ns)
dev.reduce(L.RED_MAX, np.float32, x, ns)
dev.reduce_fused(L.RED_MEAN, L.EW_MUL, np.float32, x, y, None, ns, addends=[1.0] * B3)
dev.reduce_fused(L.RED_MAX, L.EW_SUB, np.float32, x, y, None, ns, pre=L.PRE_ABS)
for p in x + y + o:
    dev.free(p)

# C5: 1e7 x 8
n5, B5 = 10**7, 8
v = [dev.upload(rng.random(n5, dtype=np.float32)) for _ in range(B5)]
k = [dev.upload(rng.integers(0, 2**31, n5, dtype=np.int64)) for _ in range(B5)]
vo = [dev.alloc(n5 * 4) for _ in range(B5)]
ko = [dev.alloc(n5 * 8) for _ in range(B5)]
ns = [n5] * B5
dev.argext(L.ARGMAX, np.float32, v, ns)
dev.cumsum(np.int64, k, ko, ns)
os.environ["BCI_SCAN"] = "lookback"
dev.cumsum(np.int64, k, ko, ns)
os.environ.pop("BCI_SCAN")
dev.sort(np.float32, v, vo, ns)
for p in v + k + vo + ko:
    dev.free(p)

# C4: SHA-256, 128 x 8 MiB (same kernel, shorter chains than 64 MiB)
sz, B4 = 8 * 2**20, 128
slab = dev.alloc(sz)
dev.rand_f64([slab], [sz // 8], [7])
dev.sha256_device([slab] * B4, [sz] * B4)
dev.sha256_device([slab + (i % 2048) * 4096 for i in range(16384)], [4096] * 16384)
dev.free(slab)

# C2: matmul 4096^3 x 2 (tcgen05 path), 1 x SIMT
m = 4096
hm = rng.random((m, m), dtype=np.float32)
A = [dev.upload(hm) for _ in range(2)]
Bm = [dev.upload(hm) for _ in range(2)]
Cm = [dev.alloc(m * m * 4) for _ in range(2)]
dev.matmul_f32(A, Bm, Cm, m, m, m)
os.environ["BCI_MATMUL"] = "simt"
dev.matmul_f32(A[:1], Bm[:1], Cm[:1], m, m, m)
dev.sync()
print("done")

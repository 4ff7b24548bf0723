"""Synthetic `/v1/execute` workloads in the shape of BASELINE.json's configs.

C1  examples/benchmark-numpy.py: rand(1e8) f64 -> square -> sum            (the metric)
C2  np.matmul 4096x4096 f32
C3  mixed elementwise+reduction on 2^20-element f32 arrays
C4  64 MiB output file -> SHA-256
C5  sort / argmax / cumsum on 1e7-row arrays
Inputs are seeded per request exactly as SURVEY.md §8d states, so the oracle can regenerate
them.  Scripts are the literal source a client would POST to /v1/execute.
"""
from __future__ import annotations

import numpy as np

# The reference's metric script, examples/benchmark-numpy.py:15-28, restated as the request
# body our benchmark posts (same statements, same printed lines).
BENCHMARK_NUMPY_SOURCE = '''import numpy
import time

def cpu_intensive_computation():
    array_size = {n}
    large_array = numpy.random.rand(array_size)
    result = numpy.sum(numpy.square(large_array))
    return result

start_time = time.time()
result = cpu_intensive_computation()
end_time = time.time()
print("Result:", result)
print("Execution Time:", end_time - start_time, "seconds")
'''

C1_N = 10**8
C1_ALGO_BYTES = 4 * 8 * C1_N  # rand W 0.8 GB + square R/W 1.6 GB + sum R 0.8 GB


def benchmark_numpy_source(n: int = C1_N) -> str:
    return BENCHMARK_NUMPY_SOURCE.format(n=n)


def c2_inputs(i: int, m: int = 4096):
    rng = np.random.default_rng(1000 + i)
    return rng.random((m, m), dtype=np.float32), rng.random((m, m), dtype=np.float32)


def c3_inputs(i: int, n: int = 2**20):
    rng = np.random.default_rng(2000 + i)
    return rng.random(n, dtype=np.float32), rng.random(n, dtype=np.float32)


C3_ALGO_BYTES = 15 * 4 * 2**20  # 15 units of 4 MiB per request


def c4_payload(i: int, nbytes: int = 64 * 2**20) -> bytes:
    return np.random.default_rng(3000 + i).bytes(nbytes)


def c5_inputs(i: int, n: int = 10**7):
    v = np.random.default_rng(4000 + i).random(n, dtype=np.float32)
    k = np.random.default_rng(5000 + i).integers(0, 2**31, n, dtype=np.int64)
    return v, k


C5_ALGO_BYTES = 560_000_000  # sort f32 360 MB + argmax 40 MB + cumsum i64 160 MB


def shard(num_requests: int, rank: int, world: int) -> list[int]:
    """Round-robin placement: request i runs on GPU i mod G."""
    return [i for i in range(num_requests) if i % world == rank]


# ---- request bodies of the other BASELINE configs, as a client would POST them ----------
# Every script prints its results with repr() so that the parity tests can compare stdout with the oracle
# executor (the reference's CPU path) running the SAME source.

C2_SOURCE = '''import numpy as np
rng = np.random.default_rng({seed})
A = rng.random(({m}, {m}), dtype=np.float32)
B = rng.random(({m}, {m}), dtype=np.float32)
C = A @ B
print(repr(float(C.max())), repr(float(C.min())), repr(float(np.mean(C))))
'''

C3_SOURCE = '''import numpy as np
rng = np.random.default_rng({seed})
x = rng.random({n}, dtype=np.float32)
y = rng.random({n}, dtype=np.float32)
r1 = np.sum(np.square(x))
r2 = np.mean(x * y + 1.0)
r3 = np.max(np.abs(x - y))
print(repr(float(r1)), repr(float(r2)), repr(float(r3)))
'''

# the same program on inputs that arrive through the request's `files` map
C3_FILES_SOURCE = '''import numpy as np
x = np.load("x.npy")
y = np.load("y.npy")
r1 = np.sum(np.square(x))
r2 = np.mean(x * y + 1.0)
r3 = np.max(np.abs(x - y))
print(repr(float(r1)), repr(float(r2)), repr(float(r3)))
'''

C4_SOURCE = '''import numpy as np
payload = np.random.default_rng({seed}).bytes({nbytes})
with open("out.bin", "wb") as f:
    f.write(payload)
print(len(payload))
'''

# C4 with the output born on the device: hashed in HBM while it is written out
C4_DEVICE_SOURCE = '''import numpy as np
x = np.random.default_rng({seed}).random({nbytes} // 8)
x.tofile("out.bin")
print(x.nbytes)
'''

C5_SOURCE = '''import numpy as np
v = np.random.default_rng({seed_v}).random({n}, dtype=np.float32)
k = np.random.default_rng({seed_k}).integers(0, 2**31, {n}, dtype=np.int64)
s = np.sort(v)
a = np.argmax(v)
c = np.cumsum(k)
print(repr(float(s[0])), repr(float(s[{n} // 2])), repr(float(s[-1])), int(a), int(c[-1]), int(c[{n} // 3]))
'''


def c2_source(i: int, m: int = 4096) -> str:
    return C2_SOURCE.format(seed=1000 + i, m=m)


def c3_source(i: int, n: int = 2**20) -> str:
    return C3_SOURCE.format(seed=2000 + i, n=n)


def c4_source(i: int, nbytes: int = 64 * 2**20, device: bool = False) -> str:
    return (C4_DEVICE_SOURCE if device else C4_SOURCE).format(seed=3000 + i, nbytes=nbytes)


def c5_source(i: int, n: int = 10**7) -> str:
    return C5_SOURCE.format(seed_v=4000 + i, seed_k=5000 + i, n=n)

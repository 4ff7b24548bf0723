#!/usr/bin/env python
"""Sort / argsort rates at the sizes of the kernel table (32 x 1e7 float32, 16 x 1e7 int64, 32 x 1e7 argsort float32) and
at request-mix shapes (many small requests, ragged sizes), CUDA events on the launching stream; results checked."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from bee_code_interpreter_b200 import device as dev  # noqa: E402

dev.init(0)
rng = np.random.default_rng(0)


def run(name, arrs, arg=False, reps=5):
    dt = arrs[0].dtype.type
    ins = [dev.upload(a) for a in arrs]
    outs = [dev.alloc(a.size * 8 if arg else a.nbytes) for a in arrs]
    ns = [a.size for a in arrs]
    ms = []
    for _ in range(reps):
        dev.event_record(20)
        (dev.argsort if arg else dev.sort)(dt, ins, outs, ns)
        dev.event_record(21)
        ms.append(dev.event_elapsed(20, 21))
    ok = True
    for j in (0, len(arrs) - 1):
        got = dev.download(outs[j], arrs[j].size, np.int64 if arg else dt)
        want = np.argsort(arrs[j], kind="stable") if arg else np.sort(arrs[j])
        ok = ok and np.array_equal(got, want)
    t = min(ms[1:])
    keys = sum(ns)
    print(f"{name}: {t:.3f} ms  {keys / t / 1e6:.1f} Gkeys/s  bit-exact={ok}", flush=True)
    for p in ins + outs:
        dev.free(p)


a32 = rng.standard_normal(10**7).astype(np.float32)
run("sort f32 32 x 1e7", [a32] * 32)
run("sort f32 uniform [0,1) 32 x 1e7  (C5's column)", [rng.random(10**7, dtype=np.float32)] * 32)
run("sort i64 in [0, 1000) 16 x 1e7", [rng.integers(0, 1000, 10**7, dtype=np.int64)] * 16)
run("sort f32 1 x 1e7", [a32])
run("sort i64 16 x 1e7", [rng.integers(-2**63, 2**63 - 1, 10**7, dtype=np.int64)] * 16)
run("sort f64 16 x 1e7", [rng.standard_normal(10**7)] * 16)
run("sort i32 32 x 1e7", [rng.integers(-2**31, 2**31 - 1, 10**7, dtype=np.int32)] * 32)
run("argsort f32 32 x 1e7", [a32] * 32, arg=True)
run("argsort i64 16 x 1e7", [rng.integers(-2**63, 2**63 - 1, 10**7, dtype=np.int64)] * 16, arg=True)
run("sort f32 256 x 1e5", [a32[:100000]] * 256)
run("sort f32 ragged 64 x (1..2e6)", [a32[: int(s)] for s in rng.integers(1, 2 * 10**6, 64)])

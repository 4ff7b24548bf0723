#!/bin/bash
# compute-sanitizer over the FINAL sort sweep (skewed-pass ballots, skipped first/last passes, 20 keys per thread)
set -x
mkdir -p gpurun_out
cat > /tmp/san_sort.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
from bee_code_interpreter_b200 import device as dev
dev.init(0)
rng = np.random.default_rng(5)
def check(dt, arrs):
    ins = [dev.upload(a) for a in arrs]
    outs = [dev.alloc(a.nbytes) for a in arrs]
    idx = [dev.alloc(a.size * 8) for a in arrs]
    ns = [a.size for a in arrs]
    dev.sort(dt, ins, outs, ns)
    dev.argsort(dt, ins, idx, ns)
    for a, o, ix in zip(arrs, outs, idx):
        assert np.array_equal(dev.download(o, a.size, dt), np.sort(a))
        assert np.array_equal(dev.download(ix, a.size, np.int64), np.argsort(a, kind="stable"))
    dev.sort(dt, ins, ins, ns)
    for p in ins + outs + idx:
        dev.free(p)
for dt in (np.float32, np.int64, np.int32, np.float64):
    f = np.issubdtype(dt, np.floating)
    sizes = (17, 321, 5121, 10241, 20001, 70001)
    check(dt, [rng.random(n).astype(dt) if f else rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, n, dtype=dt) for n in sizes])   # skewed top byte (floats)
    check(dt, [(1.0 + rng.random(600_000)).astype(dt) if f else rng.integers(0, 1000, 600_000, dtype=dt) for _ in range(2)])       # skipped passes (batch >= 2^20)
print("sanitizer workload ok")
PY
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_sort.py > gpurun_out/r2san3_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -2 gpurun_out/r2san3_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python /tmp/san_sort.py > gpurun_out/r2san3_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -2 gpurun_out/r2san3_racecheck.log

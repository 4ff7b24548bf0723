#!/bin/bash
# round 2: compute-sanitizer over the sort kernels (memcheck + racecheck on the shared-memory protocol) and the scan
set -x
mkdir -p gpurun_out
cat > /tmp/san_sort.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
from bee_code_interpreter_b200 import device as dev
dev.init(0)
rng = np.random.default_rng(5)
for dt in (np.float32, np.int64, np.int32, np.float64):
    arrs = []
    for n in (1, 17, 257, 4097, 8193, 20001, 70001):
        if np.issubdtype(dt, np.floating):
            arrs.append(rng.standard_normal(n).astype(dt))
        else:
            arrs.append(rng.integers(0, 1000, n, dtype=dt) if n % 2 else rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, n, dtype=dt))
    ins = [dev.upload(a) for a in arrs]
    outs = [dev.alloc(a.nbytes) for a in arrs]
    idx = [dev.alloc(a.size * 8) for a in arrs]
    ns = [a.size for a in arrs]
    dev.sort(dt, ins, outs, ns)
    dev.argsort(dt, ins, idx, ns)
    for a, o, ix in zip(arrs, outs, idx):
        assert np.array_equal(dev.download(o, a.size, dt), np.sort(a))
        assert np.array_equal(dev.download(ix, a.size, np.int64), np.argsort(a, kind="stable"))
    k = [dev.upload(rng.integers(0, 2**31, n, dtype=np.int64)) for n in (5, 4097, 30011)]
    ko = [dev.alloc(n * 8) for n in (5, 4097, 30011)]
    dev.cumsum(np.int64, k, ko, [5, 4097, 30011])
print("sanitizer workload ok")
PY
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python /tmp/san_sort.py > gpurun_out/r2san_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -4 gpurun_out/r2san_memcheck.log
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python /tmp/san_sort.py > gpurun_out/r2san_racecheck.log 2>&1; echo "racecheck rc=$?"; tail -6 gpurun_out/r2san_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 9 python /tmp/san_sort.py > gpurun_out/r2san_synccheck.log 2>&1; echo "synccheck rc=$?"; tail -3 gpurun_out/r2san_synccheck.log

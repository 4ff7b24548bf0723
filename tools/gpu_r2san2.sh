#!/bin/bash
# round 2: racecheck on the sort kernels alone (the scan's mbarrier pipeline is reported as hazards by the tool)
set -x
mkdir -p gpurun_out
cat > /tmp/san_sort.py <<'PY'
import sys
sys.path.insert(0, '.')
import numpy as np
from bee_code_interpreter_b200 import device as dev
dev.init(0)
rng = np.random.default_rng(5)
for dt in (np.float32, np.int64):
    arrs = []
    for n in (17, 4097, 8193, 20001, 70001):
        if np.issubdtype(dt, np.floating):
            arrs.append(rng.standard_normal(n).astype(dt))
        else:
            arrs.append(rng.integers(0, 1000, n, dtype=dt) if n % 2 == 0 else rng.integers(np.iinfo(dt).min, np.iinfo(dt).max, n, dtype=dt))
    ins = [dev.upload(a) for a in arrs]
    outs = [dev.alloc(a.nbytes) for a in arrs]
    idx = [dev.alloc(a.size * 8) for a in arrs]
    ns = [a.size for a in arrs]
    dev.sort(dt, ins, outs, ns)
    dev.argsort(dt, ins, idx, ns)
    for a, o, ix in zip(arrs, outs, idx):
        assert np.array_equal(dev.download(o, a.size, dt), np.sort(a))
        assert np.array_equal(dev.download(ix, a.size, np.int64), np.argsort(a, kind="stable"))
print("sanitizer workload ok")
PY
timeout 1200 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python /tmp/san_sort.py > gpurun_out/r2san_racecheck_sort.log 2>&1; echo "racecheck rc=$?"; tail -3 gpurun_out/r2san_racecheck_sort.log

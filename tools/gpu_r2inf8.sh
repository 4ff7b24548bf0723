#!/bin/bash
# e2e admission at 8 GPUs: requests admitted per GPU vs throughput (headline workload only)
set -x
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max
for inf in 16 24; do
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 2954$((inf/8)) bench.py --gpus 8 --steps 3 --warmup 3 --configs '' --no-cpu-baseline --no-split-check --e2e-inflight $inf > gpurun_out/r2inf8_$inf.json 2> gpurun_out/r2inf8_$inf.err
python - <<PY
import json
d=json.load(open('gpurun_out/r2inf8_$inf.json')); e=d['e2e']
print('inflight/GPU', $inf, 'e2e', e.get('value'), 'p50', e.get('p50_latency_s'), 'workers', e.get('service_workers_per_gpu'), 'per worker', e.get('inflight_per_worker'), e.get('error'))
PY
done

set -x
mkdir -p gpurun_out
BCI_ZYGOTE_POOL=0 timeout 300 python tools/e2e_cost.py --procs 4 --inflight 4 --zygotes 4 --requests 2048 --scripts pass,metric > gpurun_out/r2k_cost_nopin.jsonl 2>&1; cut -c1-400 gpurun_out/r2k_cost_nopin.jsonl
BCI_PIN_ZYGOTES=1 BCI_ZYGOTE_POOL=0 timeout 300 python tools/e2e_cost.py --procs 4 --inflight 4 --zygotes 4 --requests 2048 --scripts pass,metric > gpurun_out/r2k_cost_pin.jsonl 2>&1; cut -c1-400 gpurun_out/r2k_cost_pin.jsonl
timeout 900 python bench.py > gpurun_out/r2k_bench.json 2> gpurun_out/r2k_bench.err; tail -c 300 gpurun_out/r2k_bench.json; tail -3 gpurun_out/r2k_bench.err
BCI_PIN_ZYGOTES=1 timeout 600 python bench.py --configs "" --no-cpu-baseline --steps 2 > gpurun_out/r2k_bench_pin.json 2> gpurun_out/r2k_bench_pin.err; python -c "
import json; d=json.load(open('gpurun_out/r2k_bench_pin.json')); print('pinned e2e', d['e2e']['value'], d['e2e']['p50_latency_s'])"

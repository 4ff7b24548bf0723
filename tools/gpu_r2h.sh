set -x
mkdir -p gpurun_out
cat /proc/sys/kernel/sched_* 2>/dev/null | head; ls /sys/kernel/debug/sched 2>/dev/null | head
for pool in 0 4; do
BCI_ZYGOTE_POOL=$pool timeout 300 python tools/e2e_cost.py --procs 4 --inflight 4 --zygotes 4 --requests 2048 --scripts pass > gpurun_out/r2h_pool$pool.jsonl 2>&1; cat gpurun_out/r2h_pool$pool.jsonl
BCI_ZYGOTE_YIELD=1 BCI_ZYGOTE_POOL=$pool timeout 300 python tools/e2e_cost.py --procs 4 --inflight 4 --zygotes 4 --requests 2048 --scripts pass > gpurun_out/r2h_pool${pool}_yield.jsonl 2>&1; cat gpurun_out/r2h_pool${pool}_yield.jsonl
done
BCI_ZYGOTE_POOL=0 timeout 300 python tools/e2e_cost.py --procs 1 --inflight 16 --zygotes 8 --requests 2048 --scripts pass > gpurun_out/r2h_p1.jsonl 2>&1; cat gpurun_out/r2h_p1.jsonl

set -x
mkdir -p gpurun_out
nproc; grep -c processor /proc/cpuinfo; uname -r
timeout 600 python tools/e2e_cost.py > gpurun_out/r2c_e2e_cost.jsonl 2> gpurun_out/r2c_e2e_cost.err; cat gpurun_out/r2c_e2e_cost.jsonl; tail -3 gpurun_out/r2c_e2e_cost.err
BCI_ZYGOTE_POOL=0 timeout 600 python tools/e2e_cost.py --inflight 16 > gpurun_out/r2c_e2e_cost_nopool.jsonl 2>&1; cat gpurun_out/r2c_e2e_cost_nopool.jsonl
timeout 300 python tools/fork_scale.py --iters 200 --workers 1,4,16,32 > gpurun_out/r2c_fork_scale.log 2>&1; tail -9 gpurun_out/r2c_fork_scale.log | head -8

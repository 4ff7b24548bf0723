set -x
mkdir -p gpurun_out
timeout 300 python tools/broker_probe.py --reconnect 4,4 8,4 16,4 8,8 > gpurun_out/r2d_broker_probe.log 2>&1; grep procs gpurun_out/r2d_broker_probe.log | head -8
timeout 600 python tools/e2e_cost.py --procs 4 --inflight 4,8 --zygotes 2 --requests 1024 > gpurun_out/r2d_e2e_cost_p4.jsonl 2> gpurun_out/r2d_e2e_cost_p4.err; cat gpurun_out/r2d_e2e_cost_p4.jsonl; tail -3 gpurun_out/r2d_e2e_cost_p4.err
timeout 600 python tools/e2e_cost.py --procs 8 --inflight 2,4 --zygotes 2 --requests 512 > gpurun_out/r2d_e2e_cost_p8.jsonl 2> gpurun_out/r2d_e2e_cost_p8.err; cat gpurun_out/r2d_e2e_cost_p8.jsonl

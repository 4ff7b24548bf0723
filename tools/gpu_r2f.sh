set -x
mkdir -p gpurun_out
timeout 300 python tools/e2e_cost.py --procs 4 --inflight 4 --zygotes 2 --requests 2048 --scripts pass > gpurun_out/r2f_pass_pool4.jsonl 2>&1; cat gpurun_out/r2f_pass_pool4.jsonl
BCI_ZYGOTE_POOL=0 timeout 300 python tools/e2e_cost.py --procs 4 --inflight 4 --zygotes 2 --requests 2048 --scripts pass > gpurun_out/r2f_pass_pool0.jsonl 2>&1; cat gpurun_out/r2f_pass_pool0.jsonl
BCI_ZYGOTE_POOL=0 timeout 300 python tools/e2e_cost.py --procs 4 --inflight 4 --zygotes 4 --requests 2048 --scripts pass,metric > gpurun_out/r2f_pool0_z4.jsonl 2>&1; cat gpurun_out/r2f_pool0_z4.jsonl
BCI_ZYGOTE_POOL=8 timeout 300 python tools/e2e_cost.py --procs 4 --inflight 4 --zygotes 4 --requests 2048 --scripts pass,metric > gpurun_out/r2f_pool8_z4.jsonl 2>&1; cat gpurun_out/r2f_pool8_z4.jsonl
timeout 300 python tools/e2e_cost.py --procs 2 --inflight 8 --zygotes 4 --requests 2048 --scripts pass,metric > gpurun_out/r2f_p2.jsonl 2>&1; cat gpurun_out/r2f_p2.jsonl

set -x
mkdir -p gpurun_out
BCI_ZYGOTE_POOL=0 timeout 300 python tools/e2e_cost.py --procs 4 --inflight 4 --zygotes 4 --requests 2048 --scripts pass,metric > gpurun_out/r2g_pool0.jsonl 2>&1; cat gpurun_out/r2g_pool0.jsonl
BCI_ZYGOTE_POOL=4 timeout 300 python tools/e2e_cost.py --procs 4 --inflight 4 --zygotes 4 --requests 2048 --scripts pass,metric > gpurun_out/r2g_pool4.jsonl 2>&1; cat gpurun_out/r2g_pool4.jsonl
BCI_ZYGOTE_POOL=4 timeout 300 python tools/e2e_cost.py --procs 4 --inflight 8 --zygotes 4 --requests 2048 --scripts pass,metric > gpurun_out/r2g_pool4_i8.jsonl 2>&1; cat gpurun_out/r2g_pool4_i8.jsonl
BCI_ZYGOTE_POOL=4 timeout 300 python tools/e2e_cost.py --procs 8 --inflight 4 --zygotes 2 --requests 1024 --scripts pass,metric > gpurun_out/r2g_p8.jsonl 2>&1; cat gpurun_out/r2g_p8.jsonl

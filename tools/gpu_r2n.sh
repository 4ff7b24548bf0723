#!/bin/bash
# usage: gpu_r2n.sh N   — the driver's command line at N GPUs (bench + reference arm)
N=$1
set -x
mkdir -p gpurun_out
cat /sys/fs/cgroup/cpu.max
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; tail -c 300 gpurun_out/r2_bench_n$N.json; tail -3 gpurun_out/r2_bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/r2_bench_ref_n$N.json 2> gpurun_out/r2_bench_ref_n$N.err; cut -c1-200 gpurun_out/r2_bench_ref_n$N.json

set -x
mkdir -p gpurun_out
nvidia-smi -L | wc -l; cat /sys/fs/cgroup/cpu.max; df -h /tmp /dev/shm | tail -2; free -g | head -2
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/r2y_bench_n8.json 2> gpurun_out/r2y_bench_n8.err; tail -c 600 gpurun_out/r2y_bench_n8.json; tail -12 gpurun_out/r2y_bench_n8.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --impl reference --gpus 8 --steps 1 --warmup 0 > gpurun_out/r2y_bench_ref_n8.json 2> gpurun_out/r2y_bench_ref_n8.err; cut -c1-300 gpurun_out/r2y_bench_ref_n8.json

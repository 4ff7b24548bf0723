set -x
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2e_pytest_multi.log 2>&1; tail -5 gpurun_out/r2e_pytest_multi.log
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2e_bench_n2.json 2> gpurun_out/r2e_bench_n2.err; tail -c 1500 gpurun_out/r2e_bench_n2.json; tail -8 gpurun_out/r2e_bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2e_bench_ref_n2.json 2> gpurun_out/r2e_bench_ref_n2.err; tail -c 400 gpurun_out/r2e_bench_ref_n2.json

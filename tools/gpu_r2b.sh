set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2b_pytest.log 2>&1; tail -25 gpurun_out/r2b_pytest.log
timeout 900 python bench.py > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; tail -c 600 gpurun_out/r2b_bench.json; tail -5 gpurun_out/r2b_bench.err
./build/tools/cub_sort_probe > gpurun_out/r2b_cub_sort.log 2>&1; cat gpurun_out/r2b_cub_sort.log
timeout 600 python tools/e2e_sweep.py > gpurun_out/r2b_e2e_sweep.jsonl 2> gpurun_out/r2b_e2e_sweep.err; cat gpurun_out/r2b_e2e_sweep.jsonl; tail -3 gpurun_out/r2b_e2e_sweep.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2b_launches_bench.csv python bench.py --steps 1 --warmup 1 --e2e-requests 0 --no-cpu-baseline > gpurun_out/r2b_ncu_bench.log 2>&1; tail -2 gpurun_out/r2b_ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'reduce_rand_kernel|pcg64_kernel|fill_kernel|cast_kernel' -s 1 -c 13 -o gpurun_out/r2b_prof_new python tools/prof_r2.py > gpurun_out/r2b_ncu_new.log 2>&1; tail -3 gpurun_out/r2b_ncu_new.log
ls -la gpurun_out | tail -12

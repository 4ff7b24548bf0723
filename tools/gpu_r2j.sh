set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2j_pytest.log 2>&1; tail -6 gpurun_out/r2j_pytest.log
timeout 900 python bench.py > gpurun_out/r2j_bench.json 2> gpurun_out/r2j_bench.err; tail -c 400 gpurun_out/r2j_bench.json; tail -5 gpurun_out/r2j_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2j_bench_reference.json 2> gpurun_out/r2j_bench_reference.err; cat gpurun_out/r2j_bench_reference.json | cut -c1-600
timeout 900 python tools/kernel_bench.py --md > gpurun_out/r2j_kernel_table.txt 2> gpurun_out/r2j_kernel_table.err; head -60 gpurun_out/r2j_kernel_table.txt | cut -c1-200; tail -3 gpurun_out/r2j_kernel_table.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2j_launches_bench.csv python bench.py --steps 1 --warmup 1 --e2e-requests 0 --no-cpu-baseline > gpurun_out/r2j_ncu_bench.log 2>&1; tail -2 gpurun_out/r2j_ncu_bench.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'reduce_rand_kernel|pcg64_kernel' -s 1 -c 4 -o gpurun_out/r2j_prof_fused python tools/prof_r2.py > gpurun_out/r2j_ncu_fused.log 2>&1; tail -3 gpurun_out/r2j_ncu_fused.log
ls -la gpurun_out | tail -12

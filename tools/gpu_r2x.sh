#!/bin/bash
# round 2, run x: the whole 1-GPU check on the final code — GPU tests, smoke, bench (+ reference arm), kernel table, launch list
set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2x_pytest.log 2>&1; tail -4 gpurun_out/r2x_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2x_smoke.log 2>&1; tail -2 gpurun_out/r2x_smoke.log
timeout 900 python bench.py > gpurun_out/r2x_bench.json 2> gpurun_out/r2x_bench.err; cut -c1-400 gpurun_out/r2x_bench.json; tail -3 gpurun_out/r2x_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2x_bench_reference.json 2> gpurun_out/r2x_bench_reference.err; cut -c1-300 gpurun_out/r2x_bench_reference.json
timeout 900 python tools/kernel_bench.py --md > gpurun_out/r2x_kernel_table.txt 2> gpurun_out/r2x_kernel_table.err; grep -E "sort|K7|argm" gpurun_out/r2x_kernel_table.txt | cut -c1-160; tail -2 gpurun_out/r2x_kernel_table.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2x_launches_bench.csv python bench.py --steps 1 --warmup 1 --e2e-requests 0 --no-cpu-baseline > gpurun_out/r2x_ncu_bench.log 2>&1; tail -2 gpurun_out/r2x_ncu_bench.log

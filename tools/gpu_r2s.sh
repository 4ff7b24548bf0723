#!/bin/bash
# round 2, run s: sort parity tests + sort rates
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "sort" > gpurun_out/r2s_pytest_sort.log 2>&1; tail -3 gpurun_out/r2s_pytest_sort.log
timeout 600 python tools/sort_bench.py > gpurun_out/r2s_sort_bench.log 2>&1; cat gpurun_out/r2s_sort_bench.log
timeout 300 python tools/prof_sort_trivial.py > gpurun_out/r2s_sort_narrow.log 2>&1; cat gpurun_out/r2s_sort_narrow.log

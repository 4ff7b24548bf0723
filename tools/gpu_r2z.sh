#!/bin/bash
# round 2, run z: sort size-sweep test + the per-kernel table
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu -k "sort" > gpurun_out/r2z_pytest_sort.log 2>&1; tail -3 gpurun_out/r2z_pytest_sort.log
timeout 900 python tools/kernel_bench.py --md > gpurun_out/r2z_kernel_table.txt 2> gpurun_out/r2z_kernel_table.err; grep -E "sort" gpurun_out/r2z_kernel_table.txt | cut -c1-160; tail -2 gpurun_out/r2z_kernel_table.err

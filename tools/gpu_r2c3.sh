#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 300 python tools/prof_c3.py > gpurun_out/r2c3_events.log 2>&1; cat gpurun_out/r2c3_events.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size --clock-control none -k regex:'reduce_kernel' --csv --log-file gpurun_out/r2c3_ncu.csv python tools/prof_c3.py > gpurun_out/r2c3_ncu.log 2>&1; tail -2 gpurun_out/r2c3_ncu.log

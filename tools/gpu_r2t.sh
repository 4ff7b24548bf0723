#!/bin/bash
# round 2, run t: full ncu capture (with source counters) of the sort sweep
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'onesweep_kernel' -s 5 -c 2 -o gpurun_out/r2t_prof_sort python tools/prof_sort_r2.py > gpurun_out/r2t_ncu_sort.log 2>&1; tail -3 gpurun_out/r2t_ncu_sort.log
ls -la gpurun_out | tail -4

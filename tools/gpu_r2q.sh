#!/bin/bash
# round 2, run q: whole-batch pass skipping in the sort — parity tests + the narrow-key probe
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -k "sort" > gpurun_out/r2q_pytest_sort.log 2>&1; tail -3 gpurun_out/r2q_pytest_sort.log
timeout 300 python tools/prof_sort_trivial.py > gpurun_out/r2q_sort_trivial.log 2>&1; cat gpurun_out/r2q_sort_trivial.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/r2q_pytest.log 2>&1; tail -3 gpurun_out/r2q_pytest.log

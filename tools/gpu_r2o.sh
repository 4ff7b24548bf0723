set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; tail -8 gpurun_out/r2o_pytest.log
timeout 900 python tools/kernel_bench.py --md > gpurun_out/r2o_kernel_table.txt 2> gpurun_out/r2o_kernel_table.err; grep -E "sort|cumsum i64 \(K7\)|argmax" gpurun_out/r2o_kernel_table.txt | cut -c1-160; tail -2 gpurun_out/r2o_kernel_table.err

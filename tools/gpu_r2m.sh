set -x
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q > gpurun_out/r2m_pytest.log 2>&1; tail -8 gpurun_out/r2m_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; tail -2 gpurun_out/r2m_smoke.log
timeout 900 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err; tail -c 300 gpurun_out/r2m_bench.json; tail -3 gpurun_out/r2m_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2m_bench_reference.json 2>/dev/null; cut -c1-300 gpurun_out/r2m_bench_reference.json

set -x
timeout 1200 python -m pytest tests -x -q -m gpu > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/r2b_pytest.log
timeout 600 python bench.py > gpurun_out/r2b_bench.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/r2b_bench.log

set -x
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -4 gpurun_out/r2_pytest_gpu.log
timeout 280 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke.log
timeout 900 python bench.py > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.log 2>&1; echo "ref rc=$?"

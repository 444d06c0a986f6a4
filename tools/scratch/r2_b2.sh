timeout 900 python bench.py > gpurun_out/r2d_bench.log 2> gpurun_out/r2d_bench.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2d_bench.err
timeout 600 python -m pytest tests/test_gpu_focr.py -x -q -m gpu 2>&1 | tail -3

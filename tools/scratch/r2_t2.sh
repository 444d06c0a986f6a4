timeout 900 python -m pytest tests/test_gpu_focr.py tests/test_cli.py -x -q -m gpu 2>&1 | tail -15
FOCR_DECODE_LEGACY=1 timeout 900 python -m pytest tests/test_gpu_focr.py -x -q -m gpu 2>&1 | tail -3
timeout 600 python bench.py --steps 2 --no-cpu --no-small --no-config5 > gpurun_out/r2e_bench.log 2> gpurun_out/r2e_bench.err; echo "bench rc=$?"
tail -c 400 gpurun_out/r2e_bench.err

timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/r2f_bench.log 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
tail -c 300 gpurun_out/r2f_bench.err

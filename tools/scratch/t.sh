timeout 900 python bench.py > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err; echo "bench rc=$?"

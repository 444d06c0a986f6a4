set -x
timeout 1500 python -m pytest tests -x -q -m gpu --durations=5 > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2_pytest_gpu.log
timeout 900 python bench.py > gpurun_out/r2_bench.log 2> gpurun_out/r2_bench.err; echo "bench rc=$?"
timeout 280 python tools/prof_run.py > gpurun_out/r2_plain_prof.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_scan_tc_final python tools/prof_run.py > gpurun_out/r2_ncu_a.log 2>&1
echo "scan_tc ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:cand_exact|window_stats|sort_emit|stage_invert|row_cut|select_kernel" -s 6 -c 6 -f -o gpurun_out/r2_prof_tail_final python tools/prof_run.py > gpurun_out/r2_ncu_b.log 2>&1
echo "tail ncu rc=$?"

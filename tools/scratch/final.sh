set -x
timeout 280 python tools/prof_run.py > gpurun_out/r2_plain_prof.log 2>&1 && \
timeout 900 ncu --set full --clock-control none --import-source on -k regex:scan_tc_kernel -s 1 -c 1 -f -o gpurun_out/r2_prof_scan_tc_final python tools/prof_run.py > gpurun_out/r2_ncu_a.log 2>&1
echo "scan_tc ncu rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k "regex:cand_exact|window_stats|sort_emit|stage_invert|row_cut|select_kernel" -s 6 -c 6 -f -o gpurun_out/r2_prof_tail_final python tools/prof_run.py > gpurun_out/r2_ncu_b.log 2>&1
echo "tail ncu rc=$?"
CMD="python bench.py --steps 1 --warmup 3 --pages 32 --no-cpu --no-small --no-focr --no-config5"
timeout 280 $CMD > gpurun_out/r2_plain_bench.log 2>&1 && \
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/r2_ncu_launch.log 2>&1
echo "launch list rc=$?"
cat gpurun_out/r2_plain_prof.log

timeout 300 python tools/tc_modes.py 0 8 > gpurun_out/r2_plain6.log 2>&1 && \
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scan_tc -s 2 -c 1 -f -o gpurun_out/r2_prof_tc python tools/tc_modes.py 0 8 > gpurun_out/r2_ncu6.log 2>&1
echo "rc=$?"; tail -3 gpurun_out/r2_ncu6.log

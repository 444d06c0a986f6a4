timeout 300 python tools/tc_trace.py > gpurun_out/r2_trace2.log 2>&1; echo "trace rc=$?"
tail -8 gpurun_out/r2_trace2.log
timeout 300 python tools/tc_modes.py 0,1,2,8,16,32,57 16 > gpurun_out/r2_modes4.log 2>&1
cat gpurun_out/r2_modes4.log

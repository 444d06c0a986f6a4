timeout 300 python tools/tc_trace.py > gpurun_out/r2_trace1.log 2>&1; echo "trace rc=$?"
cat gpurun_out/r2_trace1.log | tail -20
timeout 300 python tools/tc_modes.py 0,1,2,4,8,16,32,61 16 > gpurun_out/r2_modes2.log 2>&1
cat gpurun_out/r2_modes2.log

timeout 900 python -m pytest tests/test_gpu_ncc.py -x -q -m gpu > gpurun_out/r2_pytest2.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_pytest2.log
timeout 300 python tools/tc_modes.py 0 32 > gpurun_out/r2_modes3.log 2>&1; echo "modes rc=$?"
tail -3 gpurun_out/r2_modes3.log
FOCR_TC_NOMERGE=1 timeout 300 python tools/tc_modes.py 0 32 > gpurun_out/r2_modes3_nomerge.log 2>&1
tail -2 gpurun_out/r2_modes3_nomerge.log

set -x
timeout 900 python -m pytest tests/test_gpu_ncc.py -x -q -m gpu > gpurun_out/r2_pytest1.log 2>&1; echo "pytest rc=$?" 
tail -15 gpurun_out/r2_pytest1.log
timeout 300 python tools/tc_modes.py 0 32 > gpurun_out/r2_modes1.log 2>&1; echo "modes rc=$?"
tail -5 gpurun_out/r2_modes1.log
FOCR_TC_NOMERGE=1 timeout 300 python tools/tc_modes.py 0 32 > gpurun_out/r2_modes1_nomerge.log 2>&1
tail -3 gpurun_out/r2_modes1_nomerge.log

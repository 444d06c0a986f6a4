timeout 900 python -m pytest tests/test_gpu_ncc.py -x -q -m gpu > gpurun_out/r2_pytest5.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/r2_pytest5.log
timeout 300 python tools/tc_modes.py 0 32 > gpurun_out/r2_modes7.log 2>&1; echo "modes rc=$?"
tail -3 gpurun_out/r2_modes7.log
FOCR_TC_NOMERGE=1 timeout 300 python tools/tc_modes.py 0 32 > gpurun_out/r2_modes7_nomerge.log 2>&1
tail -2 gpurun_out/r2_modes7_nomerge.log
export FOCR_B200_LIB=$PWD/font-ocr_b200/libfocr_b200_exp.so
timeout 300 python tools/tc_timeline.py > gpurun_out/r2_tl5.log 2>&1; tail -40 gpurun_out/r2_tl5.log

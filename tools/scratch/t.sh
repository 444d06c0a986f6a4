timeout 300 python tools/prof_run.py 16 0.8 2>&1 | grep "scan\|Error" | tail -2
FOCR_TC_FULLN=1 timeout 300 python tools/prof_run.py 16 0.8 2>&1 | grep "scan\|Error" | tail -2
timeout 900 python -m pytest tests/test_gpu_ncc.py -x -q -m gpu 2>&1 | tail -4

timeout 300 python tools/prof_run.py 16 0.8 2>&1 | grep "scan\|Error" | tail -2

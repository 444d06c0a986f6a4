for sp in 1 2 4 8; do echo "split $sp"; FOCR_EXACT_SPLIT=$sp timeout 300 python tools/prof_run.py 16 0.8 2>&1 | grep "scan\|Error" | tail -1; done

for b in 256 128 64 32; do echo "block $b"; FOCR_EXACT_BLOCK=$b timeout 300 python tools/prof_run.py 16 0.8 2>&1 | grep "scan\|Error" | tail -1; done

for y in 64 128 256 512; do echo "yseg $y"; FOCR_TC_YSEG=$y timeout 300 python tools/prof_run.py 16 0.8 2>&1 | grep "scan" | tail -1; done
for n in 2 3; do echo "nbuf $n"; FOCR_TC_NBUF=$n timeout 300 python tools/prof_run.py 16 0.8 2>&1 | grep "scan" | tail -1; done
echo "nomerge"; FOCR_TC_NOMERGE=1 timeout 300 python tools/prof_run.py 16 0.8 2>&1 | grep "scan" | tail -1

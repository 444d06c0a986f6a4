timeout 300 python tools/scratch/t.py 2>&1 | tail -1
LOCAL_WORLD_SIZE=8 timeout 300 python tools/scratch/t.py 2>&1 | tail -1

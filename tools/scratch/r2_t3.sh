timeout 900 python -m pytest tests/test_gpu_ncc.py -x -q -m gpu 2>&1 | tail -15
timeout 300 python tools/prof_run.py 2>&1 | tail -3

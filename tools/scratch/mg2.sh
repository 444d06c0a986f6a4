timeout 900 python -m pytest tests/test_gpu_ncc.py tests/test_gpu_focr.py -x -q -m gpu -k "multi" 2>&1 | tail -3
bash tools/scratch/mg.sh 2

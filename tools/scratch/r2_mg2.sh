nvidia-smi -L
timeout 600 python -m pytest tests/test_gpu_ncc.py tests/test_gpu_focr.py -x -q -m gpu -k "multi" 2>&1 | tail -5
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_2gpu.log 2> gpurun_out/r2_bench_2gpu.err; echo "bench2 rc=$?"
tail -c 1500 gpurun_out/r2_bench_2gpu.err

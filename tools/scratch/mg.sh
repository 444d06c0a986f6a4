N=$1
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_${N}gpu.log 2> gpurun_out/r2_bench_${N}gpu.err; echo "bench$N rc=$?"
tail -c 400 gpurun_out/r2_bench_${N}gpu.err

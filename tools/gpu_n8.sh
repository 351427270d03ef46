#!/bin/bash
# 8-GPU box: bench.py exactly as the driver launches it
mkdir -p gpurun_out
N=${1:-8}
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -n 4 gpurun_out/r02_bench_n$N.err; head -c 600 gpurun_out/r02_bench_n$N.json

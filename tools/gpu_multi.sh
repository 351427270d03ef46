#!/bin/bash
# N-rank session on one box (gpurun --gpus N): NCCL parity tests, then bench.py exactly as the driver launches it
N=${1:-2}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_multi.py tests/test_reference_scripts.py -m gpu -q -rs -s --timeout 900 -k "two_rank or two_ranks" 2>&1 | tail -40 > gpurun_out/r02_pytest_multi_n$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 \
    > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err
tail -n 3 gpurun_out/r02_bench_n$N.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_n1_samebox.json 2> gpurun_out/r02_bench_n1_samebox.err
tail -n 5 gpurun_out/r02_pytest_multi_n$N.log

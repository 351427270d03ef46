#!/bin/bash
# combined session on a 2-GPU box: risky new-kernel tests first (own process), the whole suite, smoke, NCCL parity, bench N=2 and N=1
mkdir -p gpurun_out
RISKY="conv3x3_group or wgrad_bf16"
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -rs --timeout 300 -k "$RISKY" 2>&1 | tail -30 > gpurun_out/r02_pytest_risky.log
timeout 2400 python -m pytest tests -m gpu -q -rs -s --timeout 900 -k "not ($RISKY)" > gpurun_out/r02_pytest_full.log 2>&1
grep -E "vs oracle|vs reference|vs golden|max-abs|cosine|rel L2|relative|passed|failed|FAILED|SKIPPED|NCCL|orient loss|smooth loss" gpurun_out/r02_pytest_full.log | tail -150 > gpurun_out/r02_pytest_full_summary.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -8 > gpurun_out/r02_smoke.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 \
    > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
timeout 600 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -n 3 gpurun_out/r02_pytest_risky.log gpurun_out/r02_smoke.log gpurun_out/r02_bench_n2.err
tail -n 5 gpurun_out/r02_pytest_full_summary.log

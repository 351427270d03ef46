#!/bin/bash
# multi-GPU check: bench.py under torchrun exactly as the driver launches it (N ranks, NCCL), both workloads
N=${1:-2}
mkdir -p gpurun_out
for WL in gen_fwd train_step; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
      bench.py --gpus $N --steps 5 --warmup 3 --workload $WL --no-cpu-baseline 2>&1 | grep -v "^Network\|^W0\|^\*\*\*" | tail -4 | tee gpurun_out/bench_${WL}_n$N.log
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 1 2>&1 | tail -2 | tee gpurun_out/bench_ref_n$N.log

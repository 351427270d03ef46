#!/bin/bash
# 1-GPU: A/B of the second-stream background encoder, new fused kernels' tests, host-cost probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -s --timeout 600 -k "thin_wgrad or train_mode_vs_golden or training_forward or train_iteration or backward_chain or batch8" > gpurun_out/r02_pytest_delta2.log 2>&1
grep -E "vs oracle|vs reference|passed|failed|Error|cosine 0.99[0-8]" gpurun_out/r02_pytest_delta2.log | tail -12 > gpurun_out/r02_pytest_delta2_summary.log
for ov in 0 1; do
  MICHIGAN_B200_OVERLAP=$ov timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_overlap_$ov.json 2> gpurun_out/r02_bench_overlap_$ov.err
done
timeout 300 python tools/host_time.py > gpurun_out/r02_host_time.log 2>&1
tail -n 5 gpurun_out/r02_pytest_delta2_summary.log; tail -n 2 gpurun_out/r02_host_time.log

#!/bin/bash
# 1-GPU: tests touched since the last full run, host-cost probe, default bench, config-5 train step
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s --timeout 900 -k "use_ig or wgrad_bf16 or train_iteration or backward_chain or training_forward or config1 or train_py_through" > gpurun_out/r02_pytest_delta.log 2>&1
grep -E "vs oracle|vs reference|passed|failed|Error|losses|cosine 0.99[0-8]" gpurun_out/r02_pytest_delta.log | tail -20 > gpurun_out/r02_pytest_delta_summary.log
timeout 300 python tools/host_time.py > gpurun_out/r02_host_time.log 2>&1
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final2_n1.json 2> gpurun_out/r02_bench_final2_n1.err
timeout 600 python bench.py --workload train_step_ig --steps 5 --warmup 2 --no-cpu-baseline > gpurun_out/r02_bench_train_ig.json 2> gpurun_out/r02_bench_train_ig.err
tail -n 6 gpurun_out/r02_pytest_delta_summary.log; tail -n 3 gpurun_out/r02_host_time.log; tail -n 2 gpurun_out/r02_bench_train_ig.err

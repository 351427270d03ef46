#!/bin/bash
# 1-GPU A/B of N-tile choices for the group kernel + config-1 test with the eval policy
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_reference_scripts.py tests/test_gpu_parity.py -m gpu -q -s --timeout 900 -k "config1 or add_feat_zeros or eval_mode" > gpurun_out/r02_pytest_eval2.log 2>&1
grep -E "vs oracle|vs reference|passed|failed|Error" gpurun_out/r02_pytest_eval2.log | tail -8 > gpurun_out/r02_pytest_eval2_summary.log
for cfg in "256 0" "128 0" "256 64" "128 64"; do
  set -- $cfg
  MG_SPADE_BN=$1 MG_CONV3_BN=$2 timeout 600 python bench.py --workload gen_fwd --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_bn_$1_$2.json 2> gpurun_out/r02_bench_bn_$1_$2.err
done
MG_SPADE_BN=128 MG_CONV3_BN=64 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --timeout 600 -k "train_mode_vs_golden or full_size or train_iteration" > gpurun_out/r02_pytest_bn.log 2>&1
grep -E "vs oracle|vs reference|passed|failed|Error|cosine" gpurun_out/r02_pytest_bn.log | tail -12 > gpurun_out/r02_pytest_bn_summary.log
cat gpurun_out/r02_pytest_eval2_summary.log gpurun_out/r02_pytest_bn_summary.log | tail -n 12

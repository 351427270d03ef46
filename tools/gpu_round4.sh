#!/bin/bash
# 1-GPU: previously failing eval-geometry tests, A/B of bf16 conv gradients, ncu --set full capture of the dominant kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_reference_scripts.py -m gpu -q -rs -s --timeout 900 -k "add_feat_zeros or config1" > gpurun_out/r02_pytest_eval.log 2>&1
grep -E "vs oracle|passed|failed|Error" gpurun_out/r02_pytest_eval.log | tail -8 > gpurun_out/r02_pytest_eval_summary.log
MICHIGAN_B200_GRAD16=all timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -s --timeout 600 -k "train_iteration or backward_chain" > gpurun_out/r02_pytest_grad16.log 2>&1
grep -E "cosine|rel L2|err |passed|failed|smooth loss|losses" gpurun_out/r02_pytest_grad16.log | tail -60 > gpurun_out/r02_pytest_grad16_summary.log
timeout 600 python bench.py --workload train_step --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_train_tf32conv.json 2> gpurun_out/r02_bench_train_tf32conv.err
MICHIGAN_B200_GRAD16=all timeout 600 python bench.py --workload train_step --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_train_grad16.json 2> gpurun_out/r02_bench_train_grad16.err
bash tools/ncu_dominant.sh f16 > gpurun_out/r02_ncu_dominant.log 2>&1
cp gpurun_out/prof_spade_f16_summary.txt gpurun_out/r02_ncu_spade_gemm_f16.txt 2>/dev/null
rm -f gpurun_out/prof_spade_f16.ncu-rep
cat gpurun_out/r02_pytest_eval_summary.log; tail -n 3 gpurun_out/r02_pytest_grad16_summary.log

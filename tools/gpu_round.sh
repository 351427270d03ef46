#!/bin/bash
# standard GPU session: parity tests (risky new-kernel tests in their own process: a device trap poisons the CUDA context),
# smoke, bench
mkdir -p gpurun_out
RISKY="conv3x3_group"
timeout 1800 python -m pytest tests -m gpu -q -rs --timeout 900 -k "not ($RISKY)" 2>&1 | tail -150 > gpurun_out/r02_pytest_main.log
timeout 600 python -m pytest tests/test_gpu_kernels.py -m gpu -q -rs --timeout 300 -k "$RISKY" 2>&1 | tail -60 > gpurun_out/r02_pytest_risky.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench.json 2> gpurun_out/r02_bench.err
tail -5 gpurun_out/r02_bench.err
tail -3 gpurun_out/r02_pytest_main.log gpurun_out/r02_pytest_risky.log

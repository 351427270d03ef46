#!/bin/bash
# standard GPU session: parity tests, smoke, bench, launch list
mkdir -p gpurun_out
python __graft_entry__.py smoke 2>&1 | tail -5 | tee gpurun_out/smoke.log
timeout 1200 python -m pytest tests -m gpu -x -q -s 2>&1 | tail -60 | tee gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 10 --warmup 3 2>&1 | tail -5 | tee gpurun_out/bench.log

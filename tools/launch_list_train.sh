#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/launches_train.csv \
    python bench.py --workload train_step --steps 1 --warmup 1 > gpurun_out/launch_bench_train.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_train.csv --top | tee gpurun_out/launches_train_summary.txt | head -70

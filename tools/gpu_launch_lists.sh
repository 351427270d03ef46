#!/bin/bash
# ncu launch lists (gpu__time_duration.sum per launch) of both bench legs on the current build
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_gen_end.csv \
    python bench.py --workload gen_fwd --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launch_bench_gen_end.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_gen_end.csv --top > gpurun_out/r02_launches_gen_fwd_end.txt
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 7200 --csv --log-file gpurun_out/launches_train_end.csv \
    python bench.py --workload train_step --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launch_bench_train_end.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_train_end.csv --top > gpurun_out/r02_launches_train_step_end.txt
head -12 gpurun_out/r02_launches_gen_fwd_end.txt; head -14 gpurun_out/r02_launches_train_step_end.txt

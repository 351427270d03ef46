#!/bin/bash
# per-launch device time of one generator forward (cold-cache, serialised: compare shares)
mkdir -p gpurun_out
MODE=${1:-mixed16}
MICHIGAN_B200_PRECISION=$MODE ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv \
    --log-file gpurun_out/launches_$MODE.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/launch_bench_$MODE.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_$MODE.csv | tee gpurun_out/launches_${MODE}_summary.txt | head -60

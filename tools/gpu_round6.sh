#!/bin/bash
# 2-GPU box: the whole -m gpu suite exactly as the driver runs it (-x), smoke, default bench, train-step launch list
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -x -q -m gpu -rs -s --timeout 900 > gpurun_out/r02_pytest_gpu_final.log 2>&1
grep -E "vs oracle|vs reference|vs golden|max-abs|cosine|rel L2|relative L2|passed|failed|FAILED|SKIPPED|'world'|orient loss|smooth loss|losses" gpurun_out/r02_pytest_gpu_final.log | tail -120 > gpurun_out/r02_pytest_gpu_final_summary.log
timeout 600 python __graft_entry__.py smoke 2>&1 | tail -6 > gpurun_out/r02_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_final_n1.json 2> gpurun_out/r02_bench_final_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_train2.csv \
    python bench.py --workload train_step --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launch_train2.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_train2.csv --top > gpurun_out/r02_launches_train_final_summary.txt
rm -f gpurun_out/r02_launches_train2.csv
tail -n 4 gpurun_out/r02_pytest_gpu_final_summary.log gpurun_out/r02_smoke.log

#!/bin/bash
# A/B of the 3x3 group kernel + launch lists
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_scripts.py tests/test_gpu_orient.py tests/test_gpu_kernels.py -m gpu -q -rs -s --timeout 600 \
  -k "training_forward or add_feat_zeros or config1 or orient or prologue" 2>&1 | tail -60 > gpurun_out/r02_pytest_fix.log
for g in 0 1 2; do
  MG_GROUP3=$g timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_g3_$g.json 2> gpurun_out/r02_bench_g3_$g.err
done
MG_GROUP3=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_gen.csv \
    python bench.py --workload gen_fwd --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launch_gen.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_gen.csv > gpurun_out/r02_launches_gen_summary.txt
MG_GROUP3=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 12000 --csv --log-file gpurun_out/r02_launches_train.csv \
    python bench.py --workload train_step --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/r02_launch_train.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_train.csv --top > gpurun_out/r02_launches_train_summary.txt
rm -f gpurun_out/r02_launches_gen.csv gpurun_out/r02_launches_train.csv
tail -n 4 gpurun_out/r02_pytest_fix.log

#!/bin/bash
# A/B of the early accumulator release in the transposed epilogue (MG_EPI_EARLY)
mkdir -p gpurun_out
{
MG_EPI_EARLY=1 timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "igemm or group or spade or conv or dgrad or wgrad" 2>&1 | tail -2
for k in 1 0; do
MG_EPI_EARLY=$k timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_bench_early_$k.json
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_early_$k.json").read())
t = d.get("train_step", {})
print("epi_early=$k gen", d["ms_per_step"], d["value"], "worst", d["roofline_worst"]["ms_per_launch"], "train", t.get("ms_per_step"), t.get("value"), d["clocks"]["sm_mhz"])
PY
done
} > gpurun_out/r02_ab_epi_early.log 2>&1
cat gpurun_out/r02_ab_epi_early.log

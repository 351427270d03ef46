#!/bin/bash
mkdir -p gpurun_out
bash tools/gpu_whatif.sh > /dev/null 2>&1
grep -E "ms/launch|MG_EPI_TMA|MG_DBG= ?(0|4|8|64) |passed|failed" gpurun_out/r02_whatif_spade_gemm_tma.log | cut -c1-200
timeout 1500 python -m pytest tests -x -q -m gpu -s --timeout 900 > gpurun_out/r02_pytest_gpu_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_final.log
tail -3 gpurun_out/r02_pytest_gpu_final.log
grep -E "config 1|config1|vs reference|max-abs" gpurun_out/r02_pytest_gpu_final.log | head -12 | cut -c1-200
for k in 1 0; do
MG_BN_FILL=$k timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r02_bench_bnfill_$k.json
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_bnfill_$k.json").read())
t = d.get("train_step", {})
print("bn_fill=$k gen", d["ms_per_step"], d["value"], "frac", d["roofline"]["frac"], "train", t.get("ms_per_step"), t.get("value"), d["clocks"]["sm_mhz"])
PY
done

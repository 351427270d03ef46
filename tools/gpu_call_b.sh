#!/bin/bash
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "spade or igemm or group" 2>&1 | tail -3
for k in 0 2 0 2; do
  echo "MG_EPI_TMA=$k"; MG_EPI_TMA=$k MG_TIME=1 timeout 120 python tools/run_kernel.py spade 2>&1 | grep ms/launch
done
MICHIGAN_B200_LIB=michigan_b200/lib/libmichigan_sm100_probes.so timeout 300 python tools/whatif_spade.py f16 2>&1 | grep -E "MG_DBG= ?(0|4|8|64) |MG_DBG=16 " | cut -c1-400
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_reference_scripts.py -x -q -m gpu -s --timeout 900 2>&1 | grep -E "passed|failed|config 1|train-mode vs ref|eval-mode" | cut -c1-200
timeout 600 python bench.py --steps 10 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r02_bench_n1.json
python - <<PY
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read())
t = d.get("train_step", {})
print("gen", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], d["roofline"]["ms_per_launch"], "train", t.get("ms_per_step"), t.get("value"), d["clocks"])
PY
} > gpurun_out/r02_early_release.log 2>&1
cat gpurun_out/r02_early_release.log

#!/bin/bash
# Final 1-GPU validation of the round: the whole -m gpu suite, smoke(), the default bench line, the reference arm
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu -s --timeout 900 > gpurun_out/r02_pytest_gpu_final.log 2>&1
echo "pytest rc=$?" >> gpurun_out/r02_pytest_gpu_final.log
tail -4 gpurun_out/r02_pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/r02_smoke.log
tail -2 gpurun_out/r02_smoke.log
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n1.json").read().strip().splitlines()[-1])
t = d.get("train_step", {})
print("gen", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "frac", d["roofline"]["frac"], "worst", d["roofline_worst"]["frac"])
print("train", t.get("ms_per_step"), t.get("value"), "e2e", t.get("e2e", {}).get("value"), "frac", t.get("roofline", {}).get("frac"))
print("clocks", d.get("clocks"), "cpu", d.get("cpu_baseline", {}).get("value"))
PY
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n1.json 2>/dev/null; tail -c 400 gpurun_out/r02_bench_ref_n1.json
MG_TIME=1 timeout 120 python tools/run_kernel.py stats 2>&1 | grep ms/launch
if [ "$1" = "ncu" ]; then bash tools/ncu_kernels.sh spade seg wgrad16 stats; fi

#!/bin/bash
# quick 2-rank check of the driver's launch line on the final build (bench.py under torch.distributed.run, NCCL)
mkdir -p gpurun_out
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline \
    > gpurun_out/r02_bench_n2_end.json 2> gpurun_out/r02_bench_n2_end.err
echo "rc=$?"; tail -n 2 gpurun_out/r02_bench_n2_end.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r02_bench_n2_end.json").read().strip().splitlines()[-1])
t = d.get("train_step", {})
print("n2 gen", d["ms_per_step"], d["value"], "e2e", d["e2e"]["value"], "train", t.get("ms_per_step"), t.get("value"))
PY

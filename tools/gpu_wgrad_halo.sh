#!/bin/bash
# parity + A/B of the halo-patch weight-gradient schedule (MG_WGRAD_HALO)
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "wgrad" 2>&1 | tail -8
for k in 0 1 0 1; do
  echo "MG_WGRAD_HALO=$k"; MG_WGRAD_HALO=$k MG_TIME=1 timeout 120 python tools/run_kernel.py wgrad16 2>&1 | grep ms/launch
done
MG_WGRAD_HALO=1 timeout 400 python bench.py --workload train_step --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('halo=1', d['ms_per_step'], d['value'], d.get('roofline',{}).get('ms_per_launch'))"
MG_WGRAD_HALO=0 timeout 400 python bench.py --workload train_step --steps 10 --warmup 3 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('halo=0', d['ms_per_step'], d['value'], d.get('roofline',{}).get('ms_per_launch'))"
} > gpurun_out/r02_ab_wgrad_halo.log 2>&1
tail -30 gpurun_out/r02_ab_wgrad_halo.log

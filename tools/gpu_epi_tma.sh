#!/bin/bash
# parity + A/B of the SPADE row-per-lane + TMA-store epilogue (MG_EPI_TMA)
mkdir -p gpurun_out
{
timeout 300 python -m pytest tests/test_gpu_kernels.py -x -q -k "spade or igemm" 2>&1 | tail -8
for k in 0 1 0 1; do
  echo "MG_EPI_TMA=$k"; MG_EPI_TMA=$k MG_TIME=1 timeout 120 python tools/run_kernel.py spade 2>&1 | grep ms/launch
done
for k in 1 0; do
MG_EPI_TMA=$k timeout 400 python bench.py --workload gen_fwd --steps 20 --warmup 5 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('epi_tma=$k', d['ms_per_step'], d['value'], d['e2e']['value'], d['roofline']['ms_per_launch'], d['roofline']['frac'])"
done
} > gpurun_out/r02_ab_epi_tma.log 2>&1
tail -30 gpurun_out/r02_ab_epi_tma.log

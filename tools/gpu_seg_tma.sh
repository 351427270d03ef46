#!/bin/bash
# A/B of the seg-conv TMA-store epilogue (MG_SEG_TMA) + its parity test
mkdir -p gpurun_out
{
python -m pytest tests/test_gpu_kernels.py -x -q -k "seg_conv" 2>&1 | tail -5
for split in "" 1; do for k in 0 1 0 1; do
  echo "MG_SEG_TMA=$k split=$split"; MG_SEG_SPLIT=$split MG_SEG_TMA=$k MG_TIME=1 python tools/run_kernel.py seg 2>&1 | grep ms/launch
done; done
MG_SEG_TMA=1 python bench.py --workload gen_fwd --steps 20 --warmup 5 2>&1 | tail -1
MG_SEG_TMA=0 python bench.py --workload gen_fwd --steps 20 --warmup 5 2>&1 | tail -1
} > gpurun_out/r02_ab_seg_tma.log 2>&1
tail -30 gpurun_out/r02_ab_seg_tma.log

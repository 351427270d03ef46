#!/bin/bash
# what-if probes + role profile of the dominant kernel on the probe build, both epilogues; A/B of the x prefetch (MG_EPI_TMA=2)
mkdir -p gpurun_out
{
for k in 0 1 2 1 2; do
  echo "MG_EPI_TMA=$k"; MG_EPI_TMA=$k MG_TIME=1 timeout 120 python tools/run_kernel.py spade 2>&1 | grep ms/launch
done
for k in 1 2 0; do
  echo "=== probe build, MG_EPI_TMA=$k"
  MG_EPI_TMA=$k MICHIGAN_B200_LIB=michigan_b200/lib/libmichigan_sm100_probes.so timeout 300 python tools/whatif_spade.py f16 2>&1 | tail -20
done
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "spade" 2>&1 | tail -3

} > gpurun_out/r02_whatif_spade_gemm_tma.log 2>&1
cat gpurun_out/r02_whatif_spade_gemm_tma.log | cut -c1-400

#!/bin/bash
# run every probe group in its own process, bounded, log to gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/probe_gpu.txt 2>&1
for g in "$@"; do
  echo "##### $g" 
  timeout 300 python tools/probe_kernels.py $g 2>&1 | tee gpurun_out/probe_$g.log | tail -40
done

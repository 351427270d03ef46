#!/bin/bash
# one `ncu --set full` capture per hot-path kernel (B200_PROFILING.md recipe): tensor-pipe %, DRAM bytes, duration -> profiles/
mkdir -p gpurun_out
for k in "$@"; do
  case $k in
    spade|group|dconv) re="igemm|conv3x3" ;;
    wgrad16|wgrad32) re="wgrad" ;;
    seg) re="seg_mlp" ;;
    stats) re="chan_stats" ;;
    spade_bwd) re="spade_bwd" ;;
  esac
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$re -s 2 -c 1 -f -o gpurun_out/ncu_$k python tools/run_kernel.py $k > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/ncu_$k.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > gpurun_out/r02_ncu_$k.txt
  rm -f gpurun_out/ncu_$k.ncu-rep
  head -n 3 gpurun_out/r02_ncu_$k.txt
done

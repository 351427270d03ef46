#!/bin/bash
# one `ncu --set full` capture of the dominant kernel per operand mode (B200_PROFILING.md recipe)
mkdir -p gpurun_out
for mode in ${@:-f16}; do
  ncu --set full --clock-control none --import-source on -k regex:igemm -s 2 -c 1 -f -o gpurun_out/prof_spade_$mode \
      python tools/run_dominant.py $mode > gpurun_out/ncu_$mode.log 2>&1
  ncu -i gpurun_out/prof_spade_$mode.ncu-rep --page raw --csv 2>/dev/null | python tools/ncu_pick.py > gpurun_out/prof_spade_${mode}_summary.txt
  cat gpurun_out/prof_spade_${mode}_summary.txt
done

#!/bin/bash
# round 2, GPU call 15: ncu counters of the compiled point addition (k_prog_jit) beside the interpreter (k_prog)
set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none -k regex:'k_prog' -c 4 -o gpurun_out/r2f_jit python tools/ncu_targets.py jit > gpurun_out/ncu_f.log 2>&1
ncu -i gpurun_out/r2f_jit.ncu-rep --page raw --csv > gpurun_out/r2f_jit.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2f_jit.csv > gpurun_out/r2_ncu_jit.txt
rm -f gpurun_out/r2f_jit.ncu-rep
cat gpurun_out/r2_ncu_jit.txt

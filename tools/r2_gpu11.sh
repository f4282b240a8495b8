#!/bin/bash
# round 2, GPU call 11 (1 GPU): compiled programs, occupancy variants
set -x
mkdir -p gpurun_out
: > gpurun_out/r2_jit_programs.txt
for MB in 0 3 4 5; do
  MAB_JIT_MINBLOCKS=$MB timeout 300 python tools/bench_jit.py 21 2>&1 | tee -a gpurun_out/r2_jit_programs.txt
done

#!/bin/bash
# round 2, GPU call 30 (1 GPU): chunk rule for the remainder of a queue (MAB_LADDER_TAIL), short batches
set -x
mkdir -p gpurun_out
: > gpurun_out/r2_tail_rule.txt
export LGS="15 16 17 18 19 20 21"
for t in 0 1 2 0 1; do
MAB_LADDER_TAIL=$t timeout 600 python tools/compare_kernels.py 2>&1 | grep -v perkey | sed "s/^/tail=$t  /" | tee -a gpurun_out/r2_tail_rule.txt
done

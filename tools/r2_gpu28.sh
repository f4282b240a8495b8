#!/bin/bash
# round 2, GPU call 28 (1 GPU): step loop unrolled twice (-DMAB_STEP_UNROLL2) against the shipped build, both ladders
set -x
mkdir -p gpurun_out
V=modarith_b200/build/variants
: > gpurun_out/r2_step_unroll.txt
export LGS="18 20 22"
timeout 600 python tools/compare_kernels.py 2>&1 | sed "s/^/shipped  /" | tee -a gpurun_out/r2_step_unroll.txt
MODARITH_B200_LIB=$V/unr2/libmodarith_b200.so timeout 600 python tools/compare_kernels.py 2>&1 | sed "s/^/unroll2  /" | tee -a gpurun_out/r2_step_unroll.txt

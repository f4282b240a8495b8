#!/bin/bash
# round 2, GPU call 25 (1 GPU): X25519 / X448 ladders at fewer resident CTAs per SM than fit (MAB_LADDER_CTAS), shipped
# build (128 registers) and the 146-register build
set -x
mkdir -p gpurun_out
V=modarith_b200/build/variants
: > gpurun_out/r2_ladder_ctas.txt
for c in 4 3 2; do
  MAB_LADDER_CTAS=$c timeout 300 python tools/compare_kernels.py 2>&1 | grep -v perkey | sed "s/^/shipped ctas=$c  /" | tee -a gpurun_out/r2_ladder_ctas.txt
done
for c in 3 2; do
  MAB_LADDER_CTAS=$c MODARITH_B200_LIB=$V/a_mb3/libmodarith_b200.so timeout 300 python tools/compare_kernels.py 2>&1 | grep -v perkey | sed "s/^/a_mb3   ctas=$c  /" | tee -a gpurun_out/r2_ladder_ctas.txt
done
MAB_LADDER_CTAS=1 timeout 300 python tools/compare_kernels.py 2>&1 | grep -v perkey | grep X448 | sed "s/^/shipped ctas=1  /" | tee -a gpurun_out/r2_ladder_ctas.txt

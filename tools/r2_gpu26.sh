#!/bin/bash
# round 2, GPU call 26 (1 GPU): X25519 ladder, 146-register build at three CTAs/SM against the 128-register build at four,
# batch sizes 2^15 .. 2^24
set -x
mkdir -p gpurun_out
V=modarith_b200/build/variants
: > gpurun_out/r2_x25519_mb3_sizes.txt
export LGS="15 16 17 18 19 20 21 22 24"
timeout 600 python tools/compare_kernels.py 2>&1 | grep -v perkey | grep X25519 | sed "s/^/128 regs, 4 CTAs\/SM  /" | tee -a gpurun_out/r2_x25519_mb3_sizes.txt
MODARITH_B200_LIB=$V/a_mb3/libmodarith_b200.so timeout 600 python tools/compare_kernels.py 2>&1 | grep -v perkey | grep X25519 | sed "s/^/146 regs, 3 CTAs\/SM  /" | tee -a gpurun_out/r2_x25519_mb3_sizes.txt

#!/bin/bash
# round 2, GPU call 24 (1 GPU): X25519 ladder at three resident CTAs per SM (146 registers, same step loop) against the
# shipped four (128 registers): does the X448 observation (fewer warps, fewer register-bank conflicts) carry over?
set -x
mkdir -p gpurun_out
V=modarith_b200/build/variants
: > gpurun_out/r2_x25519_occupancy.txt
timeout 300 python tools/compare_kernels.py 2>&1 | grep X25519 | sed "s/^/shipped   /" | tee -a gpurun_out/r2_x25519_occupancy.txt
MODARITH_B200_LIB=$V/a_mb3/libmodarith_b200.so timeout 300 python tools/compare_kernels.py 2>&1 | grep X25519 | sed "s/^/a_mb3     /" | tee -a gpurun_out/r2_x25519_occupancy.txt
timeout 300 python tools/compare_kernels.py 2>&1 | grep X25519 | sed "s/^/shipped   /" | tee -a gpurun_out/r2_x25519_occupancy.txt
MODARITH_B200_LIB=$V/a_mb3/libmodarith_b200.so timeout 300 python tools/compare_kernels.py 2>&1 | grep X25519 | sed "s/^/a_mb3     /" | tee -a gpurun_out/r2_x25519_occupancy.txt

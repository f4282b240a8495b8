#!/bin/bash
# round 2, GPU call 29 (1 GPU): generator knobs on the X25519 ladder again, now that the kernel is compiled for three
# CTAs per SM and has registers to spare (MAB_ROWS=fresh spilled at the 128-register budget)
set -x
mkdir -p gpurun_out
V=modarith_b200/build/variants
: > gpurun_out/r2_x25519_knobs.txt
export LGS="18 20 22"
for rep in 1 2; do
timeout 600 python tools/compare_kernels.py 2>&1 | grep X25519 | grep -v perkey | sed "s/^/shipped  /" | tee -a gpurun_out/r2_x25519_knobs.txt
for t in y_rows y_zero; do
MODARITH_B200_LIB=$V/$t/libmodarith_b200.so timeout 600 python tools/compare_kernels.py 2>&1 | grep X25519 | grep -v perkey | sed "s/^/$t   /" | tee -a gpurun_out/r2_x25519_knobs.txt
done
done

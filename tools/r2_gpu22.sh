#!/bin/bash
# round 2, GPU call 22 (1 GPU): why does the X448 ladder not gain from a third warp per sub-partition?  modmul chain of
# X448 at full occupancy against the model, and the ncu counters of the 3-CTA build beside the shipped 2-CTA build
set -x
mkdir -p gpurun_out
timeout 300 python tools/bench_addon_chain.py X448 X25519 NIST256 2>&1 | tail -4 | tee gpurun_out/r2_x448_chain.txt
V=modarith_b200/build/variants
MODARITH_B200_LIB=$V/mb3/libmodarith_b200.so timeout 900 ncu --set full --clock-control none -k regex:'k_rfc7748_rounds' -s 1 -c 1 -o gpurun_out/r2f_x448_mb3 python tools/ncu_targets.py x448 > gpurun_out/ncu_mb3.log 2>&1
ncu -i gpurun_out/r2f_x448_mb3.ncu-rep --page raw --csv > gpurun_out/r2f_x448_mb3.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2f_x448_mb3.csv | tee gpurun_out/r2_ncu_x448_mb3.txt
rm -f gpurun_out/r2f_x448_mb3.ncu-rep

#!/bin/bash
# round 2, GPU call 23 (1 GPU): X448 ladder at 2 against 3 resident CTAs per SM.  The round-2 "mb3" measurement never
# ran three: the K = 4 result slots need 98 KB of shared memory per CTA.  With two slots per thread (K = 2) three CTAs
# fit; k2_mb2 = K = 2 at 2 CTAs/SM (210 registers), k2_mb3 = K = 2 at 3 CTAs/SM (168 registers)
set -x
mkdir -p gpurun_out
V=modarith_b200/build/variants
: > gpurun_out/r2_x448_occupancy.txt
for t in k2_mb2 k2_mb3 mb3; do
  MODARITH_B200_LIB=$V/$t/libmodarith_b200.so timeout 300 python tools/compare_kernels.py 2>&1 | grep X448 | sed "s/^/$(printf '%-10s' $t)/" | tee -a gpurun_out/r2_x448_occupancy.txt
done
MODARITH_B200_LIB=$V/k2_mb3/libmodarith_b200.so timeout 900 ncu --set full --clock-control none -k regex:'k_rfc7748_rounds' -s 1 -c 1 -o gpurun_out/r2f_x448_k2mb3 python tools/ncu_targets.py x448 > gpurun_out/ncu_k2mb3.log 2>&1
ncu -i gpurun_out/r2f_x448_k2mb3.ncu-rep --page raw --csv > gpurun_out/r2f_x448_k2mb3.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2f_x448_k2mb3.csv | tee gpurun_out/r2_ncu_x448_k2mb3.txt
rm -f gpurun_out/r2f_x448_k2mb3.ncu-rep

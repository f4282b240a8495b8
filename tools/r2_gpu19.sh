#!/bin/bash
# round 2, GPU call 19 (1 GPU): generator-knob variants on the X448 ladder (rows / zero / capture / stash) and the
# carry-capture-on-the-multiplier-pipe variant (MAB_CAPOP=madc) on the P-256 chains; bench of the current tree
set -x
mkdir -p gpurun_out
V=modarith_b200/build/variants
: > gpurun_out/r2_x448_knobs.txt
timeout 300 python tools/compare_kernels.py 2>&1 | grep X448 | sed 's/^/shipped     /' | tee -a gpurun_out/r2_x448_knobs.txt
for t in x_rows x_zero x_cap x_nostash capmadc; do
  MODARITH_B200_LIB=$V/$t/libmodarith_b200.so timeout 300 python tools/compare_kernels.py 2>&1 | grep X448 | sed "s/^/$(printf '%-12s' $t)/" | tee -a gpurun_out/r2_x448_knobs.txt
done
: > gpurun_out/r2_p256_capop.txt
timeout 300 python tools/bench_p256_field.py shipped 2>&1 | tail -6 | tee -a gpurun_out/r2_p256_capop.txt
MODARITH_B200_LIB=$V/capmadc/libmodarith_b200.so timeout 300 python tools/bench_p256_field.py capmadc 2>&1 | tail -6 | tee -a gpurun_out/r2_p256_capop.txt
timeout 300 python tools/bench_p256_field.py shipped 2>&1 | tail -6 | tee -a gpurun_out/r2_p256_capop.txt
timeout 600 python bench.py --no-extra > gpurun_out/r2_bench_n1_cfg.json 2> gpurun_out/r2_bench_n1_cfg.err; tail -c 400 gpurun_out/r2_bench_n1_cfg.json

#!/bin/bash
# round 2, GPU call 3 (1 GPU): tests, queue A/B after the lane-parallel sweep, bench with the new P-256 reduction
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_gpu3_pytest.txt; cat gpurun_out/r2_gpu3_pytest.txt
: > gpurun_out/r2_queue_ab2.txt
for LG in 17 18 20 22; do
  for Q in 1 0; do
    MAB_LADDER_QUEUES=$Q timeout 300 python bench.py --keys $((1 << LG)) --steps 10 --warmup 3 --no-extra --no-cpu-baseline --parity-keys 65536 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print('lg',$LG,'single_queue',$Q,'value %.2f M/s  e2e %.2f M/s  frac %.4f parity %s' % (j['value']/1e6, j['e2e']['value']/1e6, j['roofline']['frac'], j['parity_spot_check']))
" | tee -a gpurun_out/r2_queue_ab2.txt
  done
done
timeout 600 python bench.py > gpurun_out/r2_bench2.json 2> gpurun_out/r2_bench2.err; tail -3 gpurun_out/r2_bench2.err; head -c 300 gpurun_out/r2_bench2.json
# X448 at 3 CTAs/SM (168 registers, small spill) against the shipped 2 CTAs/SM build
timeout 300 python tools/compare_kernels.py 2>&1 | grep X448 | sed 's/^/shipped  /' | tee gpurun_out/r2_x448_variants.txt
MODARITH_B200_LIB=modarith_b200/build/variants/x448_mb3/libmodarith_b200.so timeout 300 python tools/compare_kernels.py 2>&1 | grep X448 | sed 's/^/mb3      /' | tee -a gpurun_out/r2_x448_variants.txt

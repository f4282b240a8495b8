#!/bin/bash
# round 2, GPU call 8 (1 GPU): even one-round chunking A/B on short batches, full GPU suite
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gpu8_pytest.txt; cat gpurun_out/r2_gpu8_pytest.txt
: > gpurun_out/r2_even_ab.txt
for LG in 16 17 18 19 20; do
  for EV in 0 1; do
    MAB_LADDER_EVEN=$EV timeout 300 python bench.py --keys $((1 << LG)) --steps 10 --warmup 3 --no-extra --no-cpu-baseline --parity-keys 65536 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print('lg',$LG,'even',$EV,'value %.2f M/s  e2e %.2f M/s  frac %.4f parity %s' % (j['value']/1e6, j['e2e']['value']/1e6, j['roofline']['frac'], j['parity_spot_check']))
" | tee -a gpurun_out/r2_even_ab.txt
  done
done

#!/bin/bash
# round 2, GPU call 31 (1 GPU): the tree with the short-batch tail rule as default: GPU suite, smoke, bench, short sizes
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gpu_pytest.txt; cat gpurun_out/r2_gpu_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 200 gpurun_out/r2_bench_n1.json
for LG in 16 17 18; do
  timeout 300 python bench.py --keys $((1 << LG)) --steps 10 --warmup 3 --no-extra --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import json,sys
j=json.loads(sys.stdin.read())
print('lg',$LG,'value %.2f M/s  e2e %.2f M/s  frac %.4f parity %s %d' % (j['value']/1e6, j['e2e']['value']/1e6, j['roofline']['frac'], j['parity_spot_check'], j['parity_keys']))
" | tee -a gpurun_out/r2_short_batches.txt
done

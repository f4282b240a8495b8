#!/bin/bash
# round 2, GPU call 10 (1 GPU): compiled field programs (mab_<P>_modprog_jit): tests, bench extra
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_modprog.py -x -q 2>&1 | tail -15 > gpurun_out/r2_gpu10_pytest.txt; cat gpurun_out/r2_gpu10_pytest.txt
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/r2_gpu10_bench.json 2> gpurun_out/r2_gpu10_bench.err
python - <<'PY'
import json
j = json.loads(open('gpurun_out/r2_gpu10_bench.json').read().strip().split('\n')[-1])
print(json.dumps(j['extra'].get('nist256_modprog_point_addition'), indent=1))
print(j['value'], j['e2e'])
PY

#!/bin/bash
# round 2, GPU call 6 (1 GPU): full GPU suite, bench (Jacobian window step, modprog), ecn variants
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_gpu6_pytest.txt; cat gpurun_out/r2_gpu6_pytest.txt
timeout 600 python bench.py > gpurun_out/r2_bench4.json 2> gpurun_out/r2_bench4.err; tail -3 gpurun_out/r2_bench4.err; head -c 300 gpurun_out/r2_bench4.json
python - <<'PY'
import json
d = json.loads(open("gpurun_out/r2_bench4.json").read().strip().splitlines()[-1])
x = d["extra"]
for k in ("nist256_ecnmul", "nist256_ecnmul2", "ed25519_ecnmul", "nist256_modprog_point_addition"):
    print(k, json.dumps(x.get(k))[:400])
PY

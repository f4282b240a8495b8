#!/bin/bash
# round 2, GPU call 17 (1 GPU): Montgomery-friendly plan (ED248, NIST384 add-ons) vs reference builds; chain rates against the fall-back plan
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_extra_modulus.py -x -q 2>&1 | tail -12 > gpurun_out/r2_gpu17_pytest.txt; cat gpurun_out/r2_gpu17_pytest.txt
timeout 600 python tools/bench_addon_chain.py ED248 ED248F NIST384 NIST384F C41417 NIST521 M383 2>&1 | tee gpurun_out/r2_addon_chains2.txt

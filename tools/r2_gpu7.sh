#!/bin/bash
# round 2, GPU call 7 (1 GPU): GPU suite on the halved-doubling build, ecnmul occupancy variants
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gpu7_pytest.txt; cat gpurun_out/r2_gpu7_pytest.txt
timeout 600 python tools/bench_ecn.py 2>&1 | tail -12 | tee gpurun_out/r2_ecn_variants.txt

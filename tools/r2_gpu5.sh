#!/bin/bash
# round 2, GPU call 5 (1 GPU): full GPU suite (modprog included), N=1 sweep, bench
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_gpu5_pytest.txt; cat gpurun_out/r2_gpu5_pytest.txt
timeout 600 python bench.py > gpurun_out/r2_bench3.json 2> gpurun_out/r2_bench3.err; tail -3 gpurun_out/r2_bench3.err; head -c 300 gpurun_out/r2_bench3.json
SWEEP_LGS="20 22 24 26 28" timeout 1500 bash tools/sweep.sh 1 2>&1 | tail -8
cp gpurun_out/sweep_n1.jsonl gpurun_out/r2_sweep_n1.jsonl

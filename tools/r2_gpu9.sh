#!/bin/bash
# round 2, GPU call 9 (1 GPU): FP64-pipe probe beside the integer multiplier pipe; full GPU suite on the tree with the
# even cut reverted; a bench line
set -x
mkdir -p gpurun_out
( cd tools/probe && timeout 300 ./fp64_probe 20000 ) > gpurun_out/r2_fp64_probe.txt 2>&1; cat gpurun_out/r2_fp64_probe.txt
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gpu9_pytest.txt; cat gpurun_out/r2_gpu9_pytest.txt
timeout 600 python bench.py --no-extra > gpurun_out/r2_gpu9_bench.json 2> gpurun_out/r2_gpu9_bench.err; tail -1 gpurun_out/r2_gpu9_bench.json | cut -c1-600

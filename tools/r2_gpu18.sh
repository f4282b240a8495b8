#!/bin/bash
# round 2, GPU call 18 (1 GPU): the whole GPU suite, smoke and the default bench on the final tree
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gpu_pytest.txt; cat gpurun_out/r2_gpu_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 600 gpurun_out/r2_bench_n1.json

#!/bin/bash
# round 2, GPU call 20 (1 GPU): P-256 sqr_w with its carry captures on the multiplier pipe -- the GPU suite, the P-256
# chain figures, the full bench line
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gpu_pytest.txt; cat gpurun_out/r2_gpu_pytest.txt
timeout 300 python tools/bench_p256_field.py sqrw_madc 2>&1 | tail -6 | tee gpurun_out/r2_p256_sqrw.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 300 gpurun_out/r2_bench_n1.json

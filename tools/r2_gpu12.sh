#!/bin/bash
# round 2, GPU call 12 (1 GPU): full GPU suite (compiled programs, add-on modulus), smoke
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_gpu12_pytest.txt; cat gpurun_out/r2_gpu12_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3

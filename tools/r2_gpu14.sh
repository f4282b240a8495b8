#!/bin/bash
# round 2, GPU call 14 (1 GPU): bit-level pseudo-Mersenne plan (C41417, NIST521 add-ons) vs reference builds; chain rates
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_extra_modulus.py -x -q 2>&1 | tail -12 > gpurun_out/r2_gpu14_pytest.txt; cat gpurun_out/r2_gpu14_pytest.txt
timeout 600 python tools/bench_addon_chain.py C41417 C41417F NIST521 NIST521F NIST384 M383 2>&1 | tee gpurun_out/r2_addon_chains.txt

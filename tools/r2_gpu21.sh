#!/bin/bash
# round 2, GPU call 21 (1 GPU): P-256 scalar multiplications with the carry captures of single field functions moved
# to the multiplier pipe (MAB_CAPOP_NIST256_<FN>=madc), best of three launches each, output hashes compared
set -x
mkdir -p gpurun_out
timeout 900 python tools/bench_ecn.py 2>&1 | grep -v "^+" | tee gpurun_out/r2_ecn_capop.txt
timeout 900 python tools/bench_ecn.py 2>&1 | grep -v "^+" | tee -a gpurun_out/r2_ecn_capop.txt

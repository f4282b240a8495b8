#!/bin/bash
# round 2, GPU call 1: new parity tests, issue-rate probes, baseline bench, source-level ncu of the two kernels
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_gpu1_pytest.txt; cat gpurun_out/r2_gpu1_pytest.txt
timeout 300 tools/probe/run.sh 20000 > /dev/null 2> gpurun_out/probe.err; head -50 gpurun_out/issue_probe.txt
timeout 600 python bench.py > gpurun_out/r2_bench0.json 2> gpurun_out/r2_bench0.err; tail -3 gpurun_out/r2_bench0.err; head -c 600 gpurun_out/r2_bench0.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rfc7748_rounds' -s 1 -c 1 -o gpurun_out/r2_ncu_x25519 python tools/ncu_targets.py x25519 > gpurun_out/ncu1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_field' -c 8 -o gpurun_out/r2_ncu_p256 python tools/ncu_targets.py p256 > gpurun_out/ncu2.log 2>&1
ls -la gpurun_out | tail -20

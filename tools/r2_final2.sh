#!/bin/bash
# round 2, final evidence run of the final tree (1 GPU): GPU suite, smoke, both bench arms, ncu --set full of the kernels
# that changed since r2_final.sh (scalar multiplications: capture table; X448 for the record), sanitizer smoke
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gpu_pytest.txt; cat gpurun_out/r2_gpu_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 200 gpurun_out/r2_bench_n1.json
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_n1_reference_arm.json 2>> gpurun_out/r2_bench_n1.err; head -c 300 gpurun_out/r2_bench_n1_reference_arm.json
timeout 600 ncu --set full --clock-control none -k regex:'k_ecnmul' -c 8 -o gpurun_out/r2f_ecn python tools/ncu_targets.py ecn > gpurun_out/ncu_e.log 2>&1
timeout 600 ncu --set full --clock-control none -k regex:'k_rfc7748_rounds' -s 1 -c 1 -o gpurun_out/r2f_x448 python tools/ncu_targets.py x448 > gpurun_out/ncu_b.log 2>&1
for f in ecn x448; do
  ncu -i gpurun_out/r2f_$f.ncu-rep --page raw --csv > gpurun_out/r2f_$f.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/r2f_$f.csv > gpurun_out/r2_ncu_$f.txt
  rm -f gpurun_out/r2f_$f.ncu-rep
done
head -12 gpurun_out/r2_ncu_ecn.txt
timeout 600 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r2_sanitizer_memcheck.log
timeout 600 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r2_sanitizer_racecheck.log

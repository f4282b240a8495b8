#!/bin/bash
# round 2, final evidence run (1 GPU): tests, both bench arms, launch list of the bench command, ncu --set full of
# the kernels the round worked on, sanitizer smoke
set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_final_pytest.txt; cat gpurun_out/r2_final_pytest.txt
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -2 gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference > gpurun_out/r2_bench_n1_reference_arm.json 2>> gpurun_out/r2_bench_n1.err
# launch list of the same bench command (per-launch times under ncu are serialised and cold-cache: shares only)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --parity-keys 4096 > gpurun_out/r2_bench_under_ncu.log 2>&1
python tools/ncu_summary.py --launches gpurun_out/r2_launches.csv > gpurun_out/r2_launches_bench.txt; cat gpurun_out/r2_launches_bench.txt
# full counters: ladders, P-256 field chains, scalar multiplications
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rfc7748_rounds' -s 1 -c 1 -o gpurun_out/r2f_x25519 python tools/ncu_targets.py x25519 > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_rfc7748_rounds' -s 1 -c 1 -o gpurun_out/r2f_x448 python tools/ncu_targets.py x448 > gpurun_out/ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_field' -c 8 -o gpurun_out/r2f_p256 python tools/ncu_targets.py p256 > gpurun_out/ncu_c.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_field' -c 8 -o gpurun_out/r2f_k1 python tools/ncu_targets.py k1 > gpurun_out/ncu_d.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_ecnmul' -c 8 -o gpurun_out/r2f_ecn python tools/ncu_targets.py ecn > gpurun_out/ncu_e.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_prog' -s 2 -c 4 -o gpurun_out/r2f_jit python tools/ncu_targets.py jit > gpurun_out/ncu_f.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_field|k_rfc7748_rounds' -c 6 -o gpurun_out/r2f_addon python tools/ncu_targets.py addon > gpurun_out/ncu_g.log 2>&1
for f in x25519 x448 p256 k1 ecn jit addon; do
  ncu -i gpurun_out/r2f_$f.ncu-rep --page raw --csv > gpurun_out/r2f_$f.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/r2f_$f.csv > gpurun_out/r2_ncu_$f.txt
done
rm -f gpurun_out/r2f_x448.ncu-rep gpurun_out/r2f_p256.ncu-rep gpurun_out/r2f_k1.ncu-rep gpurun_out/r2f_ecn.ncu-rep gpurun_out/r2f_jit.ncu-rep gpurun_out/r2f_addon.ncu-rep
head -30 gpurun_out/r2_ncu_x25519.txt
# sanitizers on the small smoke
timeout 900 compute-sanitizer --tool memcheck python tools/sanitize_smoke.py > gpurun_out/r2_sanitizer_memcheck.log 2>&1; tail -3 gpurun_out/r2_sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python tools/sanitize_smoke.py > gpurun_out/r2_sanitizer_racecheck.log 2>&1; tail -3 gpurun_out/r2_sanitizer_racecheck.log
ls -la gpurun_out | tail -30

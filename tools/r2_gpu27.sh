#!/bin/bash
# round 2, GPU call 27 (1 GPU): tree with the X25519 ladder at three CTAs/SM (146 registers) and the P-256 capture table:
# the GPU suite, smoke, the bench line, ncu --set full of the X25519 ladder and the P-256 chains, the launch list
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2_gpu_pytest.txt; cat gpurun_out/r2_gpu_pytest.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; tail -c 300 gpurun_out/r2_bench_n1.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_rfc7748_rounds' -s 1 -c 1 -o gpurun_out/r2f_x25519 python tools/ncu_targets.py x25519 > gpurun_out/ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none -k regex:'k_field' -c 8 -o gpurun_out/r2f_p256 python tools/ncu_targets.py p256 > gpurun_out/ncu_c.log 2>&1
for f in x25519 p256; do
  ncu -i gpurun_out/r2f_$f.ncu-rep --page raw --csv > gpurun_out/r2f_$f.csv 2>/dev/null
  python tools/ncu_summary.py gpurun_out/r2f_$f.csv > gpurun_out/r2_ncu_$f.txt
done
rm -f gpurun_out/r2f_p256.ncu-rep
head -32 gpurun_out/r2_ncu_x25519.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra --parity-keys 4096 > gpurun_out/r2_bench_under_ncu.log 2>&1
python tools/ncu_summary.py --launches gpurun_out/r2_launches.csv > gpurun_out/r2_launches_bench.txt; cat gpurun_out/r2_launches_bench.txt

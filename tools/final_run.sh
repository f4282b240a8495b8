set -x
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -2 gpurun_out/bench_final.err
python bench.py --impl reference > gpurun_out/bench_ref.json 2>> gpurun_out/bench_final.err
ncu --set full --clock-control none -k regex:'k_rfc7748_rounds|k_ecnmul' -c 8 -o /tmp/prof python tools/profile_targets.py --ladders-only > gpurun_out/ncu_run.log 2>&1
ncu -i /tmp/prof.ncu-rep --page raw --csv > gpurun_out/prof_raw.csv 2>> gpurun_out/ncu_run.log
python tools/ncu_summary.py gpurun_out/prof_raw.csv > gpurun_out/ncu_summary.txt 2>> gpurun_out/ncu_run.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
ls -la gpurun_out | head -20

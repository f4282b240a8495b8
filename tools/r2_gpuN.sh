#!/bin/bash
# round 2, multi-GPU call: usage tools/r2_gpuN.sh N -- multi-device tests, torchrun vs single-process bench, sweep
N=${1:-8}
set -x
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m pytest tests/test_gpu_multi.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2_n${N}_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>gpurun_out/r2_n${N}_torchrun.err | tail -1 > gpurun_out/r2_bench_n${N}_torchrun.json
timeout 600 python bench.py --gpus $N --single-process --steps 5 --warmup 3 2>gpurun_out/r2_n${N}_single.err | tail -1 > gpurun_out/r2_bench_n${N}_single_process.json
python - <<PY
import json
for f in ("gpurun_out/r2_bench_n${N}_torchrun.json", "gpurun_out/r2_bench_n${N}_single_process.json"):
    try:
        j = json.loads(open(f).read())
        print(f, "value %.2f M/s  e2e %.2f M/s  parity %s %s" % (j["value"] / 1e6, j["e2e"]["value"] / 1e6, j["parity_spot_check"], j["parity_keys"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
SWEEP_LGS="${SWEEP_LGS:-20 22 24 26 28}" timeout 1500 bash tools/sweep.sh $N 2>&1 | tail -8
cp gpurun_out/sweep_n$N.jsonl gpurun_out/r2_sweep_n$N.jsonl

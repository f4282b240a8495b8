#!/bin/bash
# round 2, GPU call 16 (2 GPUs): the multi-device tests and both bench forms on the final tree
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py tests/test_gpu_extra_modulus.py -x -q 2>&1 | tail -4 | tee gpurun_out/r2_final_n2_pytest.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 5 --warmup 3 --no-extra --no-cpu-baseline 2>gpurun_out/r2_final_n2_torchrun.err | tail -1 > gpurun_out/r2_final_bench_n2_torchrun.json
timeout 600 python bench.py --gpus 2 --single-process --steps 5 --warmup 3 2>gpurun_out/r2_final_n2_single.err | tail -1 > gpurun_out/r2_final_bench_n2_single_process.json
python - <<PY
import json
for f in ("gpurun_out/r2_final_bench_n2_torchrun.json", "gpurun_out/r2_final_bench_n2_single_process.json"):
    try:
        j = json.loads(open(f).read())
        print(f, "value %.2f M/s  e2e %.2f M/s  parity %s %s" % (j["value"] / 1e6, j["e2e"]["value"] / 1e6, j["parity_spot_check"], j["parity_keys"]))
    except Exception as e:
        print(f, "unreadable", e)
PY
